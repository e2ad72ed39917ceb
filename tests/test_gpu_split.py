"""GPU tests (`-m gpu`) of the fp32-grade split-operand path (csrc/split.cu, gemm_tcgen05.cu NS > 1, engine precision
'bf16x3' / 'bf16x2'): exact integer features x an fp32 operand split into bf16 parts.

Tolerances: a product with n_split parts carries 8 * n_split mantissa bits of the split operand, so against the float64
product of the UNSPLIT operands the error is <= 2^-(8 n + 1) per weight (3 parts: fp32 itself) plus fp32 accumulation
noise; stated per test."""
import math
import os

import networkx as nx
import numpy as np
import pytest
import torch

from gmc_b200 import ops, synth
from gmc_b200.engine import GCNEngine
from gmc_b200.graph import CSRGraph, GraphBatch
from golden_util import GOLDEN, PARAM_NAMES, assert_w1_digest, gcn_shapes, graph_from_edges, seeded_weights
from oracle import ref_step as rs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def bf16_exact_ints(rows, cols, hi, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, hi + 1, (rows, cols), generator=g).to(torch.float32)


@pytest.mark.parametrize("n_split", [1, 2, 3])
@pytest.mark.parametrize("rows,cols", [(1000, 500), (37, 12), (64, 8), (5, 500)])
def test_split_parts_add_up_and_pad_rows_are_zero(rows, cols, n_split):
    torch.manual_seed(rows + cols)
    W = (torch.randn(rows, cols) * torch.logspace(-3, 3, cols)).to(DEV)
    parts = ops.f32_split_bf16(W, n_split)
    sr = ops.split_rows_for(rows)
    assert parts.shape == (n_split * sr, cols)
    total = torch.zeros(rows, cols, dtype=torch.float64, device=DEV)
    for p in range(n_split):
        total += parts[p * sr: p * sr + rows].double()
        assert int(torch.count_nonzero(parts[p * sr + rows: (p + 1) * sr])) == 0
    err = ((total - W.double()).abs() / W.double().abs().clamp_min(1e-30)).max().item()
    assert err <= 2.0 ** -(8 * n_split) * 1.01        # round to nearest per part: 2^-(8n+1), bounded by 2^-8n
    if n_split == 3:
        assert torch.equal(total.float(), W)          # three parts carry every bit of an fp32 value


@pytest.mark.parametrize("n_split", [2, 3])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 500, 1000), (130, 260, 40), (5, 12, 8), (40000, 500, 1000),
                                   (300, 64, 200), (257, 500, 1024)])
def test_split_gemm_nn_matches_float64(M, N, K, n_split):
    A = bf16_exact_ints(M, K, 7, M + K)
    torch.manual_seed(N)
    W = torch.randn(K, N) * 0.05
    want = A.double() @ W.double()
    got = ops.gemm_bf16_split("nn", ops.to_bf16(A.to(DEV)), ops.f32_split_bf16(W.to(DEV), n_split), n_split, K)
    tol = (2.0 ** -17 if n_split == 2 else 4e-7) * max(1.0, math.sqrt(K) / 16)     # 2 parts: 2^-17 per weight, worst case
    assert relerr(got.cpu(), want) < tol


@pytest.mark.parametrize("op", ["nn", "tn"])
@pytest.mark.parametrize("n_split,cluster", [(2, "1"), (2, "2"), (2, "4"), (3, "1"), (3, "3")])
def test_split_gemm_cluster_shapes(op, n_split, cluster, monkeypatch):
    """Every CLM x 1 cluster shape of the split kernels (B chunks shared out by multicast) gives the same product
    (tolerance: the split error bound plus fp32 accumulation over K terms in TMEM)."""
    monkeypatch.setenv("GMC_GEMM_SPLIT_CLUSTER", cluster)
    M, N, K = (3000, 500, 1000) if op == "nn" else (1000, 500, 30000)
    A = bf16_exact_ints(M, K, 7, 17) if op == "nn" else bf16_exact_ints(K, M, 7, 17)
    torch.manual_seed(1)
    B = torch.randn(K, N) * 0.02
    want = (A.double() if op == "nn" else A.double().t()) @ B.double()
    got = ops.gemm_bf16_split(op, ops.to_bf16(A.to(DEV)), ops.f32_split_bf16(B.to(DEV), n_split), n_split, K)
    assert relerr(got.cpu(), want) < (2.0 ** -17 if n_split == 2 else 1e-6) * max(1.0, math.sqrt(K) / 16)


@pytest.mark.parametrize("n_split", [1, 2])
@pytest.mark.parametrize("rows,cols", [(1000, 500), (37, 12), (64, 8)])
def test_fp16_split_parts_add_up(rows, cols, n_split):
    """fp16 parts (11 significant bits each, part p stored scaled by 2^(12 p)): two parts reproduce an fp32 weight to
    2^-22 relative (its residual to one fp16 ulp of the scaled residual), also for weights whose residual would be an
    fp16 subnormal unscaled."""
    torch.manual_seed(rows)
    W = (torch.randn(rows, cols) * torch.logspace(-4, 0, cols)).to(DEV)
    parts = ops.f32_split_f16(W, n_split)
    sr = ops.split_rows_for(rows)
    assert parts.dtype == torch.float16 and parts.shape == (n_split * sr, cols)
    total = sum(parts[p * sr: p * sr + rows].double() * 2.0 ** (-ops.F16_LO_SHIFT * p) for p in range(n_split))
    err = ((total - W.double()).abs() / W.double().abs().clamp_min(2.0 ** -14)).max().item()
    assert err <= 2.0 ** -(11 * n_split) * 1.01
    assert int(torch.count_nonzero(parts[rows: sr])) == 0


@pytest.mark.parametrize("op,M,N,K", [("nn", 1000, 500, 1000), ("nn", 130, 260, 40), ("nn", 40000, 500, 1000), ("nn", 5, 12, 8),
                                      ("tn", 1000, 500, 30000), ("tn", 100, 64, 128)])
def test_f16_split_gemm_matches_float64(op, M, N, K):
    """fp16 integer A x two fp16 parts of an fp32 B (tcgen05.mma kind::f16 takes one 16-bit format for both operands: a
    bf16 A with fp16 parts is refused): 22 mantissa bits of B, i.e. 2^-23 per weight + fp32 accumulation noise."""
    A = bf16_exact_ints(M, K, 7, M + K) if op == "nn" else bf16_exact_ints(K, M, 7, M + K)
    torch.manual_seed(N)
    B = torch.randn(K, N) * 0.05
    want = (A.double() if op == "nn" else A.double().t()) @ B.double()
    A16 = ops.padded_empty_bf16(A.shape[0], A.shape[1], DEV, zero=True, dtype=torch.float16)
    A16.copy_(A.to(DEV))
    with pytest.raises(TypeError):
        ops.gemm_bf16_split(op, ops.to_bf16(A.to(DEV)), ops.f32_split_f16(B.to(DEV), 2), 2, K)
    got = ops.gemm_bf16_split(op, A16, ops.f32_split_f16(B.to(DEV), 2), 2, K)
    assert relerr(got.cpu(), want) < 5e-7 * max(1.0, math.sqrt(K) / 16)
    s, b = torch.rand(M) * 0.2 + 0.05, torch.randn(N) * 0.3
    if op == "nn":
        out = ops.padded_empty(M, N, DEV)
        got = ops.gemm_bf16_split("nn", A16, ops.f32_split_f16(B.to(DEV), 2), 2, K, out=out,
                                  row_scale=s.to(DEV), bias=b.to(DEV), relu=True)
        assert relerr(got.cpu(), torch.relu(s.double()[:, None] * want + b.double())) < 6e-7


@pytest.mark.parametrize("n_split", [2, 3])
@pytest.mark.parametrize("M,N,n_proj", [(1000, 500, 3), (130, 260, 4), (4097, 96, 2), (70, 12, 3)])
def test_split_gemm_fused_projection_is_deterministic_and_exact(M, N, n_proj, n_split):
    K = 256
    A = bf16_exact_ints(M, K, 6, M)
    torch.manual_seed(N)
    W, b, W2 = torch.randn(K, N) * 0.05, torch.randn(N) * 0.2, torch.randn(N, n_proj) * 0.3
    s = torch.rand(M) * 0.2 + 0.05
    Ws = ops.f32_split_bf16(W.to(DEV), n_split)
    Ab, w2p = ops.to_bf16(A.to(DEV)), ops.pad_proj_weights(W2.to(DEV))
    outs = []
    for _ in range(2):
        C = ops.padded_empty(M, N, DEV)
        T = torch.full((M, n_proj), float("nan"), device=DEV)
        ops.gemm_bf16_split("nn", Ab, Ws, n_split, K, out=C, row_scale=s.to(DEV), bias=b.to(DEV), relu=True, proj_w=w2p,
                            proj_out=T, n_proj=n_proj)
        outs.append((C.clone(), T.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    C, T = outs[0]
    assert relerr(T.cpu(), C.cpu().double() @ W2.double()) < 2e-6        # the projection of the rows it wrote
    C_plain = ops.gemm_bf16_split("nn", Ab, Ws, n_split, K, row_scale=s.to(DEV), bias=b.to(DEV), relu=True)
    assert torch.equal(C, C_plain)


@pytest.mark.parametrize("n_split", [2, 3])
@pytest.mark.parametrize("M,N,K", [(1000, 500, 1000), (130, 260, 40), (2000, 500, 1000), (300, 640, 128), (200, 512, 64)])
def test_split_gemm_nn_epilogue_scale_bias_relu(M, N, K, n_split):
    # N = 500 / 260 / 512: bias staged in shared memory; N = 640 (> 512 columns): the global-load form of the same epilogue
    A = bf16_exact_ints(M, K, 8, 3 * M)
    torch.manual_seed(K)
    W, b = torch.randn(K, N) * 0.05, torch.randn(N) * 0.3
    s = torch.rand(M) * 0.2 + 0.05
    want = torch.relu(s.double()[:, None] * (A.double() @ W.double()) + b.double())
    out = ops.padded_empty(M, N, DEV)
    got = ops.gemm_bf16_split("nn", ops.to_bf16(A.to(DEV)), ops.f32_split_bf16(W.to(DEV), n_split), n_split, K, out=out,
                              row_scale=s.to(DEV), bias=b.to(DEV), relu=True)
    assert relerr(got.cpu(), want) < (2.0 ** -17 if n_split == 2 else 5e-7)
    assert float(got.min()) >= 0.0


@pytest.mark.parametrize("n_split", [2, 3])
@pytest.mark.parametrize("M,N,K", [(1000, 500, 3000), (1000, 500, 200000), (128, 128, 4096), (100, 64, 128), (36, 8, 20),
                                   (1000, 500, 1000)])
def test_split_gemm_tn_matches_float64(M, N, K, n_split):
    """dW1 = XI^T (s . dH1pre): split-K over the node dimension, stacked parts with their pad rows."""
    A = bf16_exact_ints(K, M, 7, M + 2 * K)
    torch.manual_seed(M + N)
    B = torch.randn(K, N) * 0.01
    want = A.double().t() @ B.double()
    got = ops.gemm_bf16_split("tn", ops.to_bf16(A.to(DEV)), ops.f32_split_bf16(B.to(DEV), n_split), n_split, K)
    tol = (2.0 ** -17 if n_split == 2 else 5e-7) * max(1.0, math.sqrt(K) / 16)
    assert relerr(got.cpu(), want) < tol


def test_split_gemm_rejects_bad_arguments():
    A = ops.to_bf16(torch.ones(64, 64, device=DEV))
    B = ops.f32_split_bf16(torch.ones(64, 64, device=DEV), 2)
    with pytest.raises(ValueError):
        ops.gemm_bf16_split("nt", A, B, 2, 64)
    with pytest.raises(Exception):
        ops.gemm_bf16_split("nn", A, B, 4, 64)
    with pytest.raises(ValueError):
        ops.gemm_bf16_split("nn", A, B[:64], 2, 64)


@pytest.mark.parametrize("n_split", [2, 3])
@pytest.mark.parametrize("n,n_in,n_out", [(1000, 500, 3), (70, 16, 3), (4097, 128, 4), (333, 24, 2), (70003, 500, 3),
                                          (66000, 128, 2), (65540, 260, 4)])
def test_skinny_bwd_split_equals_the_fp32_kernel(n, n_in, n_out, n_split):
    torch.manual_seed(n + n_in)
    H = torch.relu(torch.randn(n, n_in, device=DEV))
    dT = torch.randn(n, n_out, device=DEV)
    W = torch.randn(n_in, n_out, device=DEV) * 0.2
    s = torch.rand(n, device=DEV) * 0.3 + 0.05
    dH, dW, db = ops.skinny_bwd(dT, W, H)
    S = ops.split_empty(n, n_in, n_split, DEV)
    _, dW2, db2 = ops.skinny_bwd_split(dT, W, H, n_split, S, row_scale=s)
    sr = ops.split_rows_for(n)
    total = sum(S[p * sr: p * sr + n].double() for p in range(n_split))
    want = (s[:, None] * dH).double()
    scale = want.abs().max().item()
    assert (total - want).abs().max().item() <= scale * 2.0 ** -(8 * n_split) * 1.01
    if n < 65536:
        assert torch.equal(dW, dW2) and torch.equal(db, db2)      # unscaled fp32 sums, same reduction order
    else:                                                          # TMA-streamed kernel: its own row partition
        assert relerr(dW2.cpu(), dW.cpu()) < 1e-5 and relerr(db2.cpu(), db.cpu()) < 1e-5
    assert int(torch.count_nonzero(S[n: sr])) == 0                 # pad rows of part 0 untouched


@pytest.mark.parametrize("n,n_in,n_out", [(1000, 500, 3), (70, 16, 3), (70003, 500, 3), (65540, 260, 4)])
def test_skinny_bwd_split_fp16_parts(n, n_in, n_out):
    """fp16 parts of s . dHpre (an fp16 buffer selects them): hi + 2^-12 lo carries 22 bits of every element in fp16's
    normal range, elements below it keep an absolute error of one scaled subnormal quantum; huge values saturate."""
    torch.manual_seed(n + n_in)
    H = torch.relu(torch.randn(n, n_in, device=DEV))
    dT = torch.randn(n, n_out, device=DEV)
    W = torch.randn(n_in, n_out, device=DEV) * 0.2
    s = torch.rand(n, device=DEV) * 0.3 + 0.05
    dH, dW, db = ops.skinny_bwd(dT, W, H)
    sr = ops.split_rows_for(n)
    S = ops.padded_empty_bf16(2 * sr, n_in, DEV, zero=True, dtype=torch.float16)
    _, dW2, db2 = ops.skinny_bwd_split(dT, W, H, 2, S, row_scale=s)
    total = S[:n].double() + S[sr: sr + n].double() * 2.0 ** -ops.F16_LO_SHIFT
    want = (s[:, None] * dH).double()
    err = (total - want).abs()
    assert (err <= want.abs() * 2.0 ** -21 + 2.0 ** -24).all()
    assert relerr(dW2.cpu(), dW.cpu()) < 1e-5 and relerr(db2.cpu(), db.cpu()) < 1e-5
    ops.skinny_bwd_split(dT * 1e7, W, H, 2, S, row_scale=s)                     # saturates, never inf / nan
    assert bool(torch.isfinite(S[:n].float()).all()) and float(S[:n].float().abs().max()) == 65504.0


def regular_batch(n_graphs, n, degs, seed):
    rowptr, colidx, gp = synth.regular_batch_arrays(n_graphs, n, degs, seed=seed)
    return GraphBatch.from_arrays(rowptr, colidx, gp, device=DEV), (rowptr, colidx, gp)


def oracle_items(arrays, n_graphs, width):
    rowptr, colidx, gp = arrays
    items = []
    for g in range(n_graphs):
        lo, hi = int(gp[g]), int(gp[g + 1])
        rp = (rowptr[lo: hi + 1] - rowptr[lo]).astype(np.int32)
        ci = (colidx[rowptr[lo]: rowptr[hi]] - lo).astype(np.int32)
        csr = rs.HostCSR(rp, ci, np.ones(len(ci), dtype=np.float32), hi - lo)
        items.append((csr, rs.dense_adjacency(csr, width)))
    return items


def test_integer_features_times_scale_is_ahat_x():
    batch, _ = regular_batch(6, 128, [5, 6, 7, 8, 6, 7], seed=2)
    xi = ops.IntegerFeatures.from_batch(batch, 160)
    X = ops.densify(batch, 160)
    want = ops.spmm(batch, X)
    got = xi.scale[:, None] * xi.tensor.float()
    assert relerr(got.cpu(), want.cpu()) < 1e-6
    vals = xi.tensor.float()
    assert torch.equal(vals, vals.round()) and float(vals.max()) <= 8.0
    xi16 = ops.IntegerFeatures.from_batch(batch, 160, f16=True)                 # the 'f16x2' operand: same integers
    assert xi16.tensor.dtype == torch.float16 and torch.equal(xi16.tensor.float(), vals)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16x2", "f16x2"])
@pytest.mark.parametrize("mode", ["ste", "soft"])
def test_split_engine_step_matches_the_oracle(precision, mode):
    """Batched loss and all four gradients at fixed weights == the float32 oracle (sum over graphs), rel 1e-4 -- the same
    gate the fp32 / tf32x3 engines meet in tests/test_gpu_api.py."""
    from Training import TrainingNeural as T
    B, n, F, H = 8, 160, 192, 96
    batch, arrays = regular_batch(B, n, [6, 7, 8, 7, 6, 8, 7, 7], seed=5)
    cfg = T.TrainingConfig(n_nodes=F, dim_embedding=F, hidden_dim=H, gemm_precision=precision, loss_mode=mode,
                           batch_graphs=B)
    torch.manual_seed(3)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    with torch.no_grad():
        net.conv1.bias.normal_(0, 0.1)
        net.conv2.bias.normal_(0, 0.1)
    p = rs.GCNParams(*[t.detach().cpu().clone() for t in (net.conv1.weight, net.conv1.bias, net.conv2.weight, net.conv2.bias)])
    losses, grads = rs.batch_loss_and_grads(oracle_items(arrays, B, F), p, mode=mode)
    eng = GCNEngine(net, opt, loss_mode=mode, precision=precision, adjacency_features=True)
    got = eng.loss_and_grads(batch, None)
    assert eng._integer_features(batch, None) is not None            # the tensor-core path ran, not the fallback
    np.testing.assert_allclose(got.cpu().numpy(), losses, rtol=1e-5 if mode == "ste" else 1e-4)
    for k, gten in zip(("W1", "b1", "W2", "b2"), eng.grads()):
        assert relerr(gten.cpu(), grads[k]) < 1e-4, k
    # probabilities: 3 parts reproduce fp32 to accumulation noise
    P_ref = torch.cat([rs.gcn_forward(csr, X, p)["P"] for csr, X in oracle_items(arrays, B, F)])
    assert float((eng.P[: batch.num_nodes].cpu() - P_ref).abs().max()) < (2e-5 if precision == "bf16x2" else 2e-6)


def test_split_engine_falls_back_on_irregular_graphs():
    """No single A_hat coefficient per row -> the standard layer 1 with tf32x3 GEMMs, still fp32-grade."""
    from Training import TrainingNeural as T
    g = nx.barabasi_albert_graph(150, 3, seed=1)
    nx.set_edge_attributes(g, 1, "weight")
    h = CSRGraph.from_networkx(g)
    batch = GraphBatch([h], device=DEV)
    cfg = T.TrainingConfig(n_nodes=160, dim_embedding=160, hidden_dim=64, gemm_precision="bf16x3")
    torch.manual_seed(0)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    eng = GCNEngine(net, opt, precision="bf16x3", adjacency_features=True)
    X = ops.densify(batch, 160, out=ops.padded_empty(batch.num_nodes, 160, DEV))
    assert eng._integer_features(batch, None) is None
    got = eng.loss_and_grads(batch, X)
    csr = rs.csr_from_networkx(g)
    p = rs.GCNParams(*[t.detach().cpu().clone() for t in (net.conv1.weight, net.conv1.bias, net.conv2.weight, net.conv2.bias)])
    losses, grads = rs.batch_loss_and_grads([(csr, rs.dense_adjacency(csr, 160))], p)
    np.testing.assert_allclose(got.cpu().numpy(), losses, rtol=1e-5)
    for k, gten in zip(("W1", "b1", "W2", "b2"), eng.grads()):
        assert relerr(gten.cpu(), grads[k]) < 1e-4, k


# ------------------------------------------------------------------ reference fixtures at BASELINE shapes
def _dataset_from(z, prefix_fmt, count, n):
    from Training import TrainingNeural as T
    ds = {}
    for i in range(count):
        g = graph_from_edges(z[prefix_fmt.format(i)], n)
        h = CSRGraph.from_networkx(g)
        ds[i] = [h, T.AdjacencyFeatures(h, 1000), g, [0, 1, 2]]
    return ds


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "bf16x3", "bf16x2", "f16x2"])
def test_single_step_at_baseline_shapes(precision):
    """Config 1's real shapes (n = 500, F = 1000, H = 500): P, loss and gradients of the engine == the reference's own
    run (baseline_shapes.npz part a), rel 1e-4, on every fp32-grade path."""
    from Training import TrainingNeural as T
    z = np.load(os.path.join(GOLDEN, "baseline_shapes.npz"))
    g = graph_from_edges(z["a_edges"], 500)
    import commons
    ds = {0: [CSRGraph.from_networkx(g), commons.adjacency_tensor(g, 1000), g, [0, 1, 2]]}
    cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500, learning_rate=1e-3, gemm_precision=precision)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    net.load_state_dict(seeded_weights(np.random.default_rng(int(z["a_weight_seed"])), gcn_shapes(1000, 500)))
    eng = T._engine_for(net, opt, cfg)
    step = T._prepare(ds, 1, torch.device(DEV), eng)[0]
    loss = eng.loss_and_grads(step.batch, step.X)
    assert relerr(eng.P[:500].cpu(), z["a_P"]) < 1e-4
    assert abs(loss.sum().item() - float(z["a_loss"])) <= 1e-4 * abs(float(z["a_loss"]))
    assert_w1_digest(eng.gW1.cpu().numpy(), z, "a_grad_conv1.weight", 1e-4)
    for k, gten in zip(PARAM_NAMES[1:], eng.grads()[1:]):
        assert relerr(gten.cpu(), z[f"a_grad_{k}"]) < 1e-4, k


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16x2", "f16x2"])
def test_two_epochs_at_baseline_shapes(precision, monkeypatch, capsys):
    """The 20-graph reference pipeline (complete_training_pipeline.ipynb cell 15) for two epochs through train_model:
    loss history, final weights and evaluate_model == the reference's own run (baseline_shapes.npz part b)."""
    from Training import TrainingNeural as T
    z = np.load(os.path.join(GOLDEN, "baseline_shapes.npz"))
    ds = _dataset_from(z, "b_g{}_edges", int(z["b_num_graphs"]), 500)
    cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500, learning_rate=1e-3, number_epochs=2,
                           patience=20, save_directory=None, gemm_precision=precision)
    real = T.setup_model_and_optimizer
    rng = np.random.default_rng(int(z["b_weight_seed"]))

    def seeded(config):
        net, embed, opt = real(config)
        net.load_state_dict(seeded_weights(rng, gcn_shapes(1000, 500)))
        return net, embed, opt

    monkeypatch.setattr(T, "setup_model_and_optimizer", seeded)
    net, best, epoch, inputs, hist = T.train_model(ds, cfg)
    # 'bf16x2' carries 16 of W1's 24 mantissa bits: single steps meet 1e-4 (test_single_step_at_baseline_shapes), but over
    # 40 sequential steps a near-tied argmax may flip and the trajectory then differs by that one label -- measured here:
    # loss history still within 1e-4, final weights within 5e-3.  'bf16x3' (all 24 bits) is the parity-grade path.
    wtol, ltol = (2e-2, 1e-3) if precision == "bf16x2" else (2e-4, 1e-4)
    np.testing.assert_allclose(hist, z["b_loss_history"], rtol=ltol)
    assert abs(best - float(z["b_best_loss"])) <= ltol * abs(float(z["b_best_loss"]))
    assert_w1_digest(net.conv1.weight.detach().cpu().numpy(), z, "b_final_conv1.weight", wtol)
    for k, prm in list(net.named_parameters())[1:]:
        assert relerr(prm.detach().cpu(), z[f"b_final_{k}"]) < wtol, k
    ev = T.evaluate_model(net, ds, cfg)
    assert abs(ev["total_loss"] - float(z["b_eval_total"])) <= max(ltol, 1e-3 if precision == "bf16x2" else 0) * abs(float(z["b_eval_total"]))
    assert ev["num_samples"] == 20


# ------------------------------------------------------------------ config-3 shape: hard labels of the benchmarked paths
def test_config3_shape_ste_labels_of_headline_and_parity_grade_paths():
    """64 synthetic 7-regular graphs at the benchmark's shape (n = 1000, F = 1000, H = 500), STE loss, against the
    float64 oracle at the same weights.
      * parity-grade path ('bf16x3'): logits within 2e-6 of their scale; identical labels on every node whose float64
        top-2 logit margin exceeds twice that bound; per-graph integer losses equal wherever every node is decided.
      * headline path (bf16 operands + bf16 storage + pre-aggregated layer 1): logits within the bf16 storage bound
        (3 roundings of 2^-9: W1, H1, and the fp32->bf16 A_hat X which is exact for regular graphs only up to 1/d);
        identical labels on every node whose margin exceeds twice the MEASURED logit error; flip rate reported and
        bounded; per-graph losses equal wherever every node is decided."""
    from Training import TrainingNeural as T
    B, n, F, H = 64, 1000, 1000, 500
    rowptr, colidx, gp = synth.regular_batch_arrays(B, n, 7, seed=11)
    batch = GraphBatch.from_arrays(rowptr, colidx, gp, device=DEV)
    cfg = T.TrainingConfig(n_nodes=n, dim_embedding=F, hidden_dim=H, batch_graphs=B)
    torch.manual_seed(5)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    with torch.no_grad():
        net.conv1.bias.normal_(0, 0.05)
        net.conv2.bias.normal_(0, 0.05)
    p64 = rs.GCNParams(*[t.detach().cpu().double() for t in (net.conv1.weight, net.conv1.bias, net.conv2.weight, net.conv2.bias)])
    Z64, lab64, loss64 = [], [], []
    for g in range(B):
        lo, hi = int(gp[g]), int(gp[g + 1])
        rp = (rowptr[lo: hi + 1] - rowptr[lo]).astype(np.int32)
        ci = (colidx[rowptr[lo]: rowptr[hi]] - lo).astype(np.int32)
        csr = rs.HostCSR(rp, ci, np.ones(len(ci), dtype=np.float32), n)
        fwd = rs.gcn_forward(csr, rs.dense_adjacency(csr, F).double(), p64)
        Z = torch.log(fwd["P"])                                   # logits up to a per-row constant: margins are equal
        lab = rs.hard_labels(fwd["P"])
        Z64.append(Z)
        lab64.append(lab)
        loss64.append(-rs.cut_of_labels(csr, lab))
    Z64, lab64 = torch.cat(Z64), torch.cat(lab64).numpy()
    top2 = torch.topk(Z64, 2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).numpy()

    def run(engine, feats):
        loss = engine.loss_and_grads(batch, feats).cpu().numpy()
        P = engine.P[: batch.num_nodes].cpu().double()
        Z = torch.log(P)
        # compare centred logits (softmax is shift invariant per row)
        err = ((Z - Z.mean(1, keepdim=True)) - (Z64 - Z64.mean(1, keepdim=True))).abs().max().item()
        labels = ops.argmax_labels(batch, engine.P[: batch.num_nodes]).cpu().numpy()
        return loss, err, labels

    report = {}
    eng3 = GCNEngine(net, opt, precision="bf16x3", adjacency_features=True)
    headline = GCNEngine(net, opt, precision="bf16", activations="bf16", preaggregate=True)
    XA = ops.PreaggregatedFeatures(ops.preaggregate_features_bf16(batch, F))
    eng_h = GCNEngine(net, opt, precision="f16x2", adjacency_features=True)
    for name, engine, feats, bound in (("bf16x3", eng3, None, 2e-6), ("f16x2", eng_h, None, 2e-6),
                                       ("bf16_preaggregated", headline, XA, 3 * 2.0 ** -8)):
        loss, err, labels = run(engine, feats)
        scale = float((Z64 - Z64.mean(1, keepdim=True)).abs().max())
        assert err <= bound * max(scale, 1.0), (name, err, scale)
        decided = margin > 2.0 * err
        assert (labels[decided] == lab64[decided]).all(), name
        flips = int((labels != lab64).sum())
        graph_decided = decided.reshape(B, n).all(axis=1)
        same_loss = [float(loss[g]) == float(loss64[g]) for g in range(B) if graph_decided[g]]
        assert all(same_loss), name
        # every per-graph loss is the integer cut of the labels the kernel itself chose
        from oracle import postproc as pp
        for g in (0, B // 2, B - 1):
            lo, hi = int(gp[g]), int(gp[g + 1])
            rp = (rowptr[lo: hi + 1] - rowptr[lo]).astype(np.int32)
            ci = (colidx[rowptr[lo]: rowptr[hi]] - lo).astype(np.int32)
            assert float(loss[g]) == -float(pp.cut_value(rp, ci, labels[lo:hi].astype(np.int32)))
        report[name] = dict(logit_err=err, logit_scale=scale, flips=flips, flip_rate=flips / (B * n),
                            undecided_nodes=int((~decided).sum()), graphs_compared=len(same_loss),
                            loss_sum=float(loss.sum()), loss_sum_f64=float(sum(loss64)))
    print("config3-shape label agreement:", report)
    assert report["bf16x3"]["flip_rate"] <= 1e-4 and report["f16x2"]["flip_rate"] <= 1e-4
    assert report["bf16_preaggregated"]["flip_rate"] <= 0.05
    rel = abs(report["bf16_preaggregated"]["loss_sum"] - report["bf16_preaggregated"]["loss_sum_f64"]) / abs(report["bf16_preaggregated"]["loss_sum_f64"])
    assert rel < 5e-3
