"""GPU tests (`-m gpu`) of the bf16 layer-1 activation chain (engine option activations='bf16'):
gmc_gemm_bf16_bf16out -> gmc_spmm_fused_skinny_bf16 -> gmc_skinny_bwd_bf16 -> gmc_spmm_batched_bf16, each called
through the C ABI and checked against a float64 evaluation of the SAME bf16-rounded inputs (so the only differences
are fp32 accumulation order and the final round-to-nearest-even, 2^-8 relative), then the whole engine step against the
fp32-activation engine.  Replaces autograd of TrainingNeural.py:80-83 for the throughput configuration only; the
fp32 / tf32x3 parity paths are untouched (tests/test_gpu_api.py)."""
import numpy as np
import pytest
import torch

from gmc_b200 import ops, synth
from gmc_b200.engine import GCNEngine
from gmc_b200.graph import GraphBatch

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_ULP = 2.0 ** -8          # round-to-nearest-even: half an ulp of an 8-bit significand = 2^-9; one ulp allowed


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float64)


def assert_bf16_close(got: torch.Tensor, want: torch.Tensor, extra: float = 0.0):
    """got (bf16 tensor) equals want (float64) up to one bf16 ulp plus `extra` (absolute accumulation noise)."""
    g, w = got.double().cpu(), want.double().cpu()
    tol = BF16_ULP * w.abs() + extra + 1e-30
    bad = (g - w).abs() > tol
    assert not bool(bad.any()), f"{int(bad.sum())} entries off, worst {float(((g - w).abs() / tol).max()):.2f} x tol"


def regular_batch(n_graphs, n, d, seed):
    rowptr, colidx, gp = synth.regular_batch_arrays(n_graphs, n, d, seed=seed)
    batch = GraphBatch.from_arrays(rowptr, colidx, gp, device=DEV)
    if batch.plan is None:                                  # built automatically only for >= 32 graphs
        batch.build_plan()
    return batch


def ahat_times(batch: GraphBatch, X64: torch.Tensor) -> torch.Tensor:
    """A_hat X in float64 on the host, from the batch's own CSR and per-edge coefficients."""
    rowptr = batch.rowptr.cpu().numpy().astype(np.int64)
    colidx = batch.colidx.cpu().numpy().astype(np.int64)
    coef = batch.coef.cpu().double()
    rows = torch.from_numpy(np.repeat(np.arange(batch.num_nodes), np.diff(rowptr)))
    out = torch.zeros_like(X64)
    out.index_add_(0, rows, X64[torch.from_numpy(colidx)] * coef[:, None])
    return out


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 500, 1000), (1000, 500, 256), (4000, 136, 72), (257, 24, 520)])
def test_gemm_bf16_bf16out(M, N, K):
    torch.manual_seed(M + N + K)
    A, B = torch.randn(M, K), torch.randn(K, N)
    Ab, Bb = ops.to_bf16(A.to(DEV)), ops.to_bf16(B.to(DEV))
    want = bf16_round(A) @ bf16_round(B)
    out = ops.padded_empty_bf16(M, N, DEV, zero=True)
    got = ops.gemm_bf16_bf16out("nn", Ab, Bb, out=out)
    assert got.dtype == torch.bfloat16
    # fp32 accumulation of K products of O(1) values: absolute noise ~ K * 2^-24 * |a||b|
    assert_bf16_close(got, want, extra=K * 2.0 ** -22)
    # the pad columns of the pitched buffer are never written
    full = out._base if out._base is not None else out
    assert float(full[:, N:].abs().max() if full.shape[1] > N else 0.0) == 0.0
    # the fp32-output GEMM rounded afterwards: identical when it runs without split-K (same accumulators, same rounding),
    # within one bf16 ulp when its summation order differs
    ref32 = ops.gemm_bf16("nn", Ab, Bb)
    assert_bf16_close(got, ref32.double(), extra=K * 2.0 ** -22)
    if K <= 64:
        assert torch.equal(got.cpu(), ref32.to(torch.bfloat16).cpu())


def test_gemm_bf16_bf16out_nt_tn_and_errors():
    torch.manual_seed(3)
    A, B = torch.randn(200, 96), torch.randn(72, 96)
    got = ops.gemm_bf16_bf16out("nt", ops.to_bf16(A.to(DEV)), ops.to_bf16(B.to(DEV)))
    assert_bf16_close(got, bf16_round(A) @ bf16_round(B).T, extra=1e-4)
    At, Bt = torch.randn(512, 136), torch.randn(512, 64)
    got = ops.gemm_bf16_bf16out("tn", ops.to_bf16(At.to(DEV)), ops.to_bf16(Bt.to(DEV)))
    assert_bf16_close(got, bf16_round(At).T @ bf16_round(Bt), extra=1e-3)
    with pytest.raises(Exception):                       # ldc must be a multiple of 8 elements
        bad = torch.zeros((200, 70), dtype=torch.bfloat16, device=DEV)[:, :67]
        ops.gemm_bf16_bf16out("nt", ops.to_bf16(A.to(DEV)), ops.to_bf16(B[:67].contiguous().to(DEV)), out=bad)


@pytest.mark.parametrize("C,K", [(500, 3), (128, 3), (64, 2), (20, 8), (512, 4), (36, 1)])
@pytest.mark.parametrize("out_bf16", [True, False])
def test_spmm_fused_skinny_bf16(C, K, out_bf16):
    batch = regular_batch(6, 130, 7, seed=C + K)
    N = batch.num_nodes
    torch.manual_seed(C * 7 + K)
    X = torch.randn(N, C)
    W, bias = torch.randn(C, K) / C ** 0.5, torch.randn(C) * 0.3
    Xb = ops.to_bf16(X.to(DEV))
    want_Y = torch.relu(ahat_times(batch, bf16_round(X)) + bias.double())
    Y, T = ops.spmm_fused_skinny_bf16(batch, Xb, W.to(DEV), bias=bias.to(DEV), relu=True, out_bf16=out_bf16)
    if out_bf16:
        assert Y.dtype == torch.bfloat16
        assert_bf16_close(Y, want_Y, extra=2e-6)
        want_T = Y.double().cpu() @ W.double()              # the projection consumes the rounded activations
        full = Y._base if Y._base is not None else Y
        if full.shape[1] > C:
            assert float(full[:, C:].abs().max()) == 0.0    # pad columns are exact zeros
    else:
        assert Y.dtype == torch.float32
        assert float((Y.double().cpu() - want_Y).abs().max()) < 2e-5
        want_T = want_Y @ W.double()
    assert float((T.double().cpu() - want_T).abs().max()) < 2e-5 * max(1.0, float(want_T.abs().max()))
    # no bias / no relu variant
    Y2, _ = ops.spmm_fused_skinny_bf16(batch, Xb, W.to(DEV), bias=None, relu=False, out_bf16=out_bf16)
    assert_bf16_close(Y2.to(torch.bfloat16), ahat_times(batch, bf16_round(X)), extra=2e-5)


@pytest.mark.parametrize("n_in,K,n", [(500, 3, 5000), (128, 3, 5000), (36, 8, 5000), (512, 1, 5000),
                                      (500, 3, 70001), (256, 8, 66000), (20, 2, 65536)])     # >= 65536 rows: TMA-streamed kernel
def test_skinny_bwd_bf16(n_in, K, n):
    torch.manual_seed(n_in + K)
    H = torch.relu(torch.randn(n, n_in))
    dT, W = torch.randn(n, K), torch.randn(n_in, K)
    Hb = ops.to_bf16(H.to(DEV))
    H64 = bf16_round(H)
    want_dH = (dT.double() @ W.double().T) * (H64 > 0)
    want_dW = H64.T @ dT.double()
    dH, dW, db = ops.skinny_bwd_bf16(dT.to(DEV), W.to(DEV), Hb)
    assert dH.dtype == torch.bfloat16
    assert_bf16_close(dH, want_dH, extra=1e-5)
    assert float((dW.double().cpu() - want_dW).abs().max()) < 1e-4 * float(want_dW.abs().max())
    assert float((db.double().cpu() - want_dH.sum(0)).abs().max()) < 1e-4 * float(want_dH.sum(0).abs().max() + 1)
    # bitwise reproducible (fixed-order two-stage reduction)
    dH2, dW2, db2 = ops.skinny_bwd_bf16(dT.to(DEV), W.to(DEV), Hb)
    assert torch.equal(dW, dW2) and torch.equal(db, db2) and torch.equal(dH, dH2)


@pytest.mark.parametrize("n,d,C", [(1000, 7, 500), (128, 6, 56), (130, 8, 64), (300, 7, 128), (1030, 3, 24), (2000, 7, 40)])
def test_spmm_batched_bf16(n, d, C):
    batch = regular_batch(5, n, d, seed=n + d + C)
    assert batch.plan is not None
    torch.manual_seed(n + C)
    X = torch.randn(batch.num_nodes, C)
    Xb = ops.padded_empty_bf16(batch.num_nodes, C, DEV, zero=True)
    ops.to_bf16(X.to(DEV), out=Xb)
    got = ops.spmm_bf16(batch, Xb)
    assert got.dtype == torch.bfloat16
    assert_bf16_close(got, ahat_times(batch, bf16_round(X)), extra=2e-6)
    # agrees with the fp32 slab kernel fed the same (rounded) values
    ref = ops.spmm(batch, Xb.float().contiguous())
    assert float((got.float() - ref).abs().max()) <= BF16_ULP * float(ref.abs().max())
    # forward form: fp32 bias + ReLU in the epilogue, pad columns stay zero
    if C % 4 == 0:
        bias = torch.randn(C) * 0.3
        got2 = ops.spmm_bf16(batch, Xb, bias=bias.to(DEV), relu=True)
        assert_bf16_close(got2, torch.relu(ahat_times(batch, bf16_round(X)) + bias.double()), extra=2e-6)
        full = got2._base if got2._base is not None else got2
        if full.shape[1] > C:
            assert float(full[:, C:].abs().max()) == 0.0


@pytest.mark.parametrize("n_in,K,n", [(500, 3, 4099), (128, 3, 4099), (36, 4, 4099), (512, 1, 4099), (64, 2, 4099),
                                      (500, 3, 1001), (256, 2, 777), (500, 3, 40000)])    # < 4096 rows: register-staged kernel
def test_skinny_fwd_bf16(n_in, K, n):
    torch.manual_seed(n_in * 3 + K)
    H = torch.randn(n, n_in)
    W = torch.randn(n_in, K) / n_in ** 0.5
    Hb = ops.to_bf16(H.to(DEV))
    got = ops.skinny_fwd_bf16(Hb, W.to(DEV))
    want = bf16_round(H) @ W.double()
    assert float((got.double().cpu() - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))
    with pytest.raises(Exception):
        ops.skinny_fwd_bf16(Hb, torch.randn(n_in, 5, device=DEV))       # n_out > 4: not supported by this kernel


def test_bf16_chain_on_a_ragged_batch():
    """Graphs of different sizes and degrees in one batch (config 4 mixes d = 6, 7, 8; the reference's test sets mix
    n = 50..500): slab SpMM, both skinny kernels and the fused row kernel on the same ragged block-diagonal batch."""
    import networkx as nx
    from gmc_b200.graph import CSRGraph
    specs = [(130, 6, 1), (256, 7, 2), (1000, 7, 3), (128, 8, 4), (514, 7, 5), (300, 6, 6)]
    graphs = []
    for n, d, seed in specs:
        g = nx.random_regular_graph(d=d, n=n, seed=seed)
        nx.set_edge_attributes(g, 1, "weight")
        graphs.append(CSRGraph.from_networkx(g))
    batch = GraphBatch(graphs, device=DEV)
    assert batch.build_plan()
    C, K = 500, 3
    torch.manual_seed(0)
    X = torch.randn(batch.num_nodes, C)
    bias, W = torch.randn(C) * 0.2, torch.randn(C, K) / C ** 0.5
    Xb = ops.padded_empty_bf16(batch.num_nodes, C, DEV, zero=True)
    ops.to_bf16(X.to(DEV), out=Xb)
    want = torch.relu(ahat_times(batch, bf16_round(X)) + bias.double())
    H = ops.spmm_bf16(batch, Xb, bias=bias.to(DEV), relu=True)
    assert_bf16_close(H, want, extra=2e-6)
    Hrow, Trow = ops.spmm_fused_skinny_bf16(batch, Xb, W.to(DEV), bias=bias.to(DEV), relu=True)
    assert_bf16_close(Hrow, want, extra=2e-6)
    T = ops.skinny_fwd_bf16(H, W.to(DEV))
    assert float((T.double().cpu() - H.double().cpu() @ W.double()).abs().max()) < 2e-5
    assert float((Trow.double().cpu() - Hrow.double().cpu() @ W.double()).abs().max()) < 2e-5


def test_spmm_batched_bf16_needs_a_plan():
    from gmc_b200 import _lib
    import networkx as nx
    from gmc_b200.graph import CSRGraph
    g = nx.barabasi_albert_graph(200, 3, seed=1)           # irregular: no ELL plan
    for u, v in g.edges():
        g[u][v]["weight"] = 1
    batch = GraphBatch([CSRGraph.from_networkx(g)])
    assert getattr(batch, "plan", None) is None
    Xb = torch.zeros((200, 64), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(_lib.GmcError):
        ops.spmm_bf16(batch, Xb)


def test_incremental_feature_scatter_equals_densify():
    batch = regular_batch(7, 130, 6, seed=2)
    want = ops.densify_bf16(batch, 136)
    X = ops.padded_empty_bf16(batch.num_nodes, 136, DEV, zero=True)
    ops.scatter_features_bf16(batch, 136, X)
    assert torch.equal(X, want)
    ops.scatter_features_bf16(batch, 136, X, clear=True)
    full = X._base if X._base is not None else X
    assert int(torch.count_nonzero(full)) == 0


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_bf16_activation_engine_step_tracks_the_fp32_activation_step():
    """activations='bf16' vs activations='fp32' (both gemm_precision='bf16') and vs the fp32 engine: same weights,
    same batch.  Storing T1 / H1 / dH1pre / dT1 in bf16 adds 2^-9 relative rounding per stored value; probabilities,
    loss and gradients stay within 2e-2 of the fp32 engine, and a few training steps follow the same trajectory."""
    from Training import TrainingNeural as T
    batch = regular_batch(48, 256, 7, seed=9)
    X = ops.densify(batch, 256)
    out = {}
    for name, precision, act in (("fp32", "fp32", "fp32"), ("bf16", "bf16", "fp32"), ("bf16act", "bf16", "bf16")):
        cfg = T.TrainingConfig(n_nodes=256, dim_embedding=256, hidden_dim=128, gemm_precision=precision, loss_mode="soft")
        torch.manual_seed(11)
        net, embed, opt = T.setup_model_and_optimizer(cfg)
        eng = GCNEngine(net, opt, loss_mode="soft", precision=precision, activations=act)
        loss = eng.loss_and_grads(batch, X).clone()
        grads = [g.cpu().clone() for g in eng.grads()]
        P = eng.P[: batch.num_nodes].cpu().clone()
        if act == "bf16":
            assert eng.bufA is None and eng.bufA16 is not None       # no fp32 [N, hidden] buffers were allocated
            assert torch.equal(eng.loss_and_grads(batch, X).cpu(), loss.cpu())     # deterministic
        losses = [float(eng.train_step(batch, X).sum()) for _ in range(5)]
        out[name] = (loss.cpu(), P, grads, losses, net.conv1.weight.detach().cpu().clone())
    for name in ("bf16", "bf16act"):
        assert relerr(out[name][0], out["fp32"][0]) < 1e-2
        assert float((out[name][1] - out["fp32"][1]).abs().max()) < 1e-2
        for a, b in zip(out[name][2], out["fp32"][2]):
            assert relerr(a, b) < 3e-2
        np.testing.assert_allclose(out[name][3], out["fp32"][3], rtol=1e-2)
        # five Adam steps of lr 1e-3: the first updates are ~ lr * sign(g), so a near-zero gradient entry whose sign
        # differs moves a weight by up to 2 lr per step -- 5 * 2e-3 against max |W1| ~ 0.15
        assert relerr(out[name][4], out["fp32"][4]) < 7e-2
    # the two bf16 variants differ only by the storage rounding of the activations
    for a, b in zip(out["bf16act"][2], out["bf16"][2]):
        assert relerr(a, b) < 2e-2


def test_bf16_activation_forward_variants_agree(monkeypatch):
    """GMC_FWD16=slab (bf16 slab SpMM + bf16 skinny projection) and the fused row kernel compute the same forward."""
    from Training import TrainingNeural as T
    from gmc_b200 import engine as E
    batch = regular_batch(40, 256, 7, seed=4)
    X = ops.densify(batch, 256)
    cfg = T.TrainingConfig(n_nodes=256, dim_embedding=256, hidden_dim=128, gemm_precision="bf16", loss_mode="soft")
    torch.manual_seed(5)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    res = {}
    for slab in (False, True):
        monkeypatch.setattr(E, "_FWD16_SLAB", slab)
        eng = GCNEngine(net, None, loss_mode="soft", precision="bf16", activations="bf16")
        loss = eng.loss_and_grads(batch, X).cpu().clone()
        res[slab] = (loss, eng.Z[: batch.num_nodes].cpu().clone(), [g.cpu().clone() for g in eng.grads()])
    assert relerr(res[True][1], res[False][1]) < 2e-3        # H1 may round differently by one bf16 ulp (fma order)
    assert relerr(res[True][0], res[False][0]) < 2e-3
    for a, b in zip(res[True][2], res[False][2]):
        assert relerr(a, b) < 1e-2


@pytest.mark.parametrize("N", [500, 512, 640])
def test_gemm_bf16_bf16out_bias_relu_epilogue_wide(N):
    """bias / ReLU epilogue of the pair kernel: <= 512 columns read the bias from its shared-memory copy, more columns from
    global memory -- same results."""
    torch.manual_seed(N)
    M, K = 300, 192
    A, B, bias = torch.randn(M, K), torch.randn(K, N) / K ** 0.5, torch.randn(N)
    Ab, Bb = ops.to_bf16(A.to(DEV)), ops.to_bf16(B.to(DEV))
    want = torch.relu(bf16_round(A) @ bf16_round(B) + bias.double())
    got = ops.gemm_bf16_bf16out("nn", Ab, Bb, bias=bias.to(DEV), relu=True)
    assert_bf16_close(got, want, extra=1e-4)


def test_gemm_bf16_bf16out_bias_relu_epilogue():
    torch.manual_seed(8)
    M, N, K = 700, 500, 320
    A, B, bias = torch.randn(M, K), torch.randn(K, N) / K ** 0.5, torch.randn(N)
    Ab, Bb = ops.to_bf16(A.to(DEV)), ops.to_bf16(B.to(DEV))
    want = torch.relu(bf16_round(A) @ bf16_round(B) + bias.double())
    got = ops.gemm_bf16_bf16out("nn", Ab, Bb, bias=bias.to(DEV), relu=True)
    assert_bf16_close(got, want, extra=1e-4)
    got2 = ops.gemm_bf16_bf16out("nn", Ab, Bb, bias=bias.to(DEV))
    assert_bf16_close(got2, bf16_round(A) @ bf16_round(B) + bias.double(), extra=1e-4)
    full = got._base if got._base is not None else got
    assert float(full[:, N:].abs().max()) == 0.0
    # fused skinny projection of the rounded result: P = bf16(C) @ W, reproducible (two n-tiles: two addends commute)
    for n_proj in (3, 1, 4):
        W = torch.randn(N, n_proj) / N ** 0.5
        Wp = ops.pad_proj_weights(W.to(DEV))
        P = torch.full((M, n_proj), 7.0, device=DEV)                       # overwritten, not accumulated
        got3 = ops.gemm_bf16_bf16out("nn", Ab, Bb, bias=bias.to(DEV), relu=True, proj_w=Wp, proj_out=P, n_proj=n_proj)
        assert torch.equal(got3, got)
        wantP = got.double().cpu() @ W.double()
        assert float((P.double().cpu() - wantP).abs().max()) < 2e-5 * max(1.0, float(wantP.abs().max()))
        P2 = torch.empty_like(P)
        ops.gemm_bf16_bf16out("nn", Ab, Bb, bias=bias.to(DEV), relu=True, proj_w=Wp, proj_out=P2, n_proj=n_proj)
        assert torch.equal(P, P2)


def test_fused_projection_engine_equals_separate_pass(monkeypatch):
    from Training import TrainingNeural as T
    from gmc_b200 import engine as E
    batch = regular_batch(36, 256, 7, seed=33)
    XA = ops.PreaggregatedFeatures(ops.preaggregate_features_bf16(batch, 256))
    cfg = T.TrainingConfig(n_nodes=256, dim_embedding=256, hidden_dim=128, loss_mode="soft")
    torch.manual_seed(2)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    res = {}
    for fused in (True, False):
        monkeypatch.setattr(E, "_FUSE_PROJ", fused)
        eng = GCNEngine(net, None, loss_mode="soft", precision="bf16", activations="bf16", preaggregate=True)
        loss = eng.loss_and_grads(batch, XA).cpu().clone()
        res[fused] = (loss, eng.Z[: batch.num_nodes].cpu().clone(), [g.cpu().clone() for g in eng.grads()])
    assert relerr(res[True][1], res[False][1]) < 1e-5          # same rounded H1, fp32 projection in a different order
    assert relerr(res[True][0], res[False][0]) < 1e-5
    for a, b in zip(res[True][2], res[False][2]):
        assert relerr(a, b) < 1e-4


def test_preaggregated_features_equal_ahat_times_adjacency():
    import networkx as nx
    from gmc_b200.graph import CSRGraph
    specs = [(130, 6, 1), (256, 7, 2), (60, 3, 3), (128, 8, 4)]
    graphs = []
    rng = np.random.default_rng(0)
    for n, d, seed in specs:
        g = nx.random_regular_graph(d=d, n=n, seed=seed)
        for u, v in g.edges():
            g[u][v]["weight"] = int(rng.integers(1, 4))
        graphs.append(CSRGraph.from_networkx(g))
    g = nx.barabasi_albert_graph(200, 3, seed=1)             # irregular degrees: no ELL plan needed for this path
    nx.set_edge_attributes(g, 1, "weight")
    graphs.append(CSRGraph.from_networkx(g))
    batch = GraphBatch(graphs, device=DEV)
    F = 264
    X = ops.densify(batch, F)
    want = ahat_times(batch, X.double().cpu())
    XA = ops.preaggregate_features_bf16(batch, F)
    assert_bf16_close(XA, want, extra=1e-7)
    assert torch.equal(XA, ops.preaggregate_features_bf16(batch, F))          # fixed summation order
    full = XA._base if XA._base is not None else XA
    assert float(full[:, F:].abs().max()) == 0.0 if full.shape[1] > F else True


@pytest.mark.parametrize("f16", [False, True])
def test_preaggregate_per_graph_kernel_equals_the_row_parallel_kernel(f16):
    """Unit-weight batches of small graphs take the per-graph kernel (CSR slice staged in shared memory, chosen by the
    caller-supplied max_nodes); it must write the very bits of the row-parallel kernel: regular graphs of several sizes and
    degrees (fast rows), an irregular graph (ordered slow rows), and a dense one whose edges exceed the staging buffer (walked
    from global memory by its CTA)."""
    import ctypes
    import networkx as nx
    from gmc_b200 import _lib
    from gmc_b200.graph import CSRGraph
    graphs = []
    for n, d, seed in [(130, 6, 1), (256, 7, 2), (60, 3, 3), (128, 8, 4), (256, 7, 5)]:
        g = nx.random_regular_graph(d=d, n=n, seed=seed)
        nx.set_edge_attributes(g, 1, "weight")
        graphs.append(CSRGraph.from_networkx(g))
    for g in (nx.barabasi_albert_graph(200, 3, seed=1), nx.random_regular_graph(d=20, n=250, seed=7)):
        nx.set_edge_attributes(g, 1, "weight")
        graphs.append(CSRGraph.from_networkx(g))
    batch = GraphBatch(graphs, device=DEV)
    F = 256                                                        # = max_nodes: the per-graph kernel applies
    assert batch.max_nodes <= F and 250 * 20 > 8 * F + 64          # the d = 20 graph does not fit the edge buffer
    dt = torch.float16 if f16 else torch.bfloat16
    ones = torch.ones(batch.nnz, device=DEV)
    got = ops.integer_features_bf16(batch, F, f16=f16)             # gmc_csr_preaggregate_graphs, unit coefficients
    ref = ops.padded_empty_bf16(batch.num_nodes, F, DEV, zero=True, dtype=dt)
    L, z = _lib.lib(), ctypes.c_void_p()
    args = (batch.rowptr.data_ptr(), batch.colidx.data_ptr(), ones.data_ptr(), z, batch.graph_ptr.data_ptr(), batch.num_graphs,
            batch.num_nodes, F, ref.data_ptr(), ref.stride(0))
    rc = L.gmc_csr_preaggregate_f16(*args, z) if f16 else L.gmc_csr_preaggregate_bf16(*args, z, 0, z)   # row-parallel kernel
    assert rc == 0
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    XA = ops.preaggregate_features_bf16(batch, F)                  # A_hat coefficients instead of ones
    ref2 = ops.padded_empty_bf16(batch.num_nodes, F, DEV, zero=True)
    assert L.gmc_csr_preaggregate_bf16(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), batch.coef.data_ptr(), z,
                                       batch.graph_ptr.data_ptr(), batch.num_graphs, batch.num_nodes, F, ref2.data_ptr(),
                                       ref2.stride(0), z, 0, z) == 0
    assert torch.equal(XA.view(torch.int16), ref2.view(torch.int16))


def test_preaggregate_counting_kernel_and_masked_fallback():
    """Unit weights: rows with <= 8 neighbours of one degree take the byte-counting kernel, the irregular graph's rows
    (and a 12-regular graph's) are flagged and finished by the general kernel in the same call (counting=True; opt-in --
    it measured no faster than the general kernel alone, which is the default and is checked beside it)."""
    import networkx as nx
    from gmc_b200.graph import CSRGraph
    graphs = []
    for n, d, seed in [(130, 6, 1), (1000, 7, 2), (256, 8, 3), (64, 12, 4), (50, 3, 5)]:
        g = nx.random_regular_graph(d=d, n=n, seed=seed)
        nx.set_edge_attributes(g, 1, "weight")
        graphs.append(CSRGraph.from_networkx(g))
    g = nx.barabasi_albert_graph(300, 4, seed=9)
    nx.set_edge_attributes(g, 1, "weight")
    graphs.append(CSRGraph.from_networkx(g))
    batch = GraphBatch(graphs, device=DEV)
    assert batch.wts_f32 is None                                 # unit weights: the counting kernel is eligible
    F = 1000
    want = ahat_times(batch, ops.densify(batch, F).double().cpu())
    XA = ops.preaggregate_features_bf16(batch, F, counting=True)
    assert_bf16_close(XA, want, extra=1e-7)
    assert torch.equal(XA, ops.preaggregate_features_bf16(batch, F, counting=True))
    full = XA._base if XA._base is not None else XA
    assert float(full[:, F:].abs().max()) == 0.0
    XG = ops.preaggregate_features_bf16(batch, F)                # general kernel alone: k c versus c + c + ... (one ulp)
    assert_bf16_close(XG, want, extra=1e-7)
    assert float((XG.float() - XA.float()).abs().max()) <= BF16_ULP * float(want.abs().max())


def test_preaggregated_engine_matches_the_standard_bf16_step():
    """preaggregate=True computes the same layer 1 with the aggregation applied to the features: probabilities, loss and
    gradients agree with the standard bf16-activation step (and with fp32) to bf16 rounding; passing the features
    through ops.PreaggregatedFeatures (built from the graph) or letting the engine aggregate a dense X is the same."""
    from Training import TrainingNeural as T
    batch = regular_batch(40, 256, 7, seed=21)
    X = ops.densify(batch, 256)
    out = {}
    for name, kw in (("fp32", dict(precision="fp32")), ("std16", dict(precision="bf16", activations="bf16")),
                     ("pre16", dict(precision="bf16", activations="bf16", preaggregate=True))):
        cfg = T.TrainingConfig(n_nodes=256, dim_embedding=256, hidden_dim=128, loss_mode="soft")
        torch.manual_seed(3)
        net, embed, opt = T.setup_model_and_optimizer(cfg)
        with torch.no_grad():
            net.conv1.bias.normal_(0, 0.1)
        eng = GCNEngine(net, opt, loss_mode="soft", **kw)
        loss = eng.loss_and_grads(batch, X).cpu().clone()
        rec = (loss, eng.P[: batch.num_nodes].cpu().clone(), [g.cpu().clone() for g in eng.grads()])
        if name == "pre16":
            XA = ops.PreaggregatedFeatures(ops.preaggregate_features_bf16(batch, 256))
            loss2 = eng.loss_and_grads(batch, XA).cpu()
            assert relerr(loss2, loss) < 1e-3
            assert eng.bufA is None                              # no fp32 [N, hidden] buffer, and no SpMM over hidden columns
            losses = [float(eng.train_step(batch, XA).sum()) for _ in range(4)]
            assert all(np.isfinite(losses))
        out[name] = rec
    for ref in ("fp32", "std16"):
        assert relerr(out["pre16"][0], out[ref][0]) < 1e-2
        assert float((out["pre16"][1] - out[ref][1]).abs().max()) < 1e-2
        for a, b in zip(out["pre16"][2], out[ref][2]):
            assert relerr(a, b) < 3e-2
    with pytest.raises(ValueError):
        GCNEngine(net, opt, precision="bf16", activations="fp32", preaggregate=True)


def test_training_api_with_the_throughput_configuration():
    """TrainingConfig(gemm_precision='bf16', activations='bf16', preaggregate_features=True) through train_single_epoch
    (batched steps, dataset produced by the mirror of graphExtender): epoch losses track the fp32 configuration."""
    import contextlib, io
    import networkx as nx
    from DataGenerator import graphExtender as E
    from Training import TrainingNeural as T
    graphs = {i: nx.random_regular_graph(d=5 + i % 3, n=128 + 4 * (i % 5), seed=200 + i) for i in range(36)}
    for g in graphs.values():
        nx.set_edge_attributes(g, 1, "weight")
    with contextlib.redirect_stdout(io.StringIO()):
        ds = E.process_graphs_from_folder(graphs, {i: [5 + i % 3, 40, 77] for i in range(36)}, max_nodes=160)
    runs = {}
    for name, kw in (("fp32", {}), ("std16", dict(gemm_precision="bf16", activations="bf16")),
                     ("pre16", dict(gemm_precision="bf16", activations="bf16", preaggregate_features=True))):
        cfg = T.TrainingConfig(n_nodes=160, dim_embedding=160, hidden_dim=64, batch_graphs=36, learning_rate=1e-3,
                               loss_mode="soft", **kw)
        torch.manual_seed(3)
        net, embed, opt = T.setup_model_and_optimizer(cfg)
        runs[name] = [T.train_single_epoch(ds, net, opt, embed, cfg) for _ in range(4)]
        eng = T._engine_for(net, opt, cfg)
        assert eng.preaggregate == (name == "pre16") and eng.activations == ("fp32" if name == "fp32" else "bf16")
    np.testing.assert_allclose(runs["std16"], runs["fp32"], rtol=1e-2)
    np.testing.assert_allclose(runs["pre16"], runs["fp32"], rtol=1e-2)
    assert runs["fp32"][-1] < runs["fp32"][0]                # and it trains


@pytest.mark.parametrize("preaggregate", [True, False])
def test_bf16_paths_match_the_storage_rounding_emulation(preaggregate):
    """Tight pin of the throughput configuration: the oracle evaluated in float64 with bf16 roundings at exactly the
    engine's storage points (oracle.ref_step.gcn_forward_bf16_storage) reproduces the engine's logits to ~1e-3 of their
    scale -- what is left is fp32 accumulation order and the rare bf16 rounding it flips -- against 1e-2 when the same
    logits are compared with the fp32 forward."""
    from Training import TrainingNeural as T
    from oracle import ref_step as rs
    n, F, Hd = 256, 256, 128
    rowptr, colidx, gp = synth.regular_batch_arrays(34, n, 7, seed=77)
    batch = GraphBatch.from_arrays(rowptr, colidx, gp, device=DEV)
    X = ops.densify(batch, F)
    cfg = T.TrainingConfig(n_nodes=n, dim_embedding=F, hidden_dim=Hd, loss_mode="soft")
    torch.manual_seed(4)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    with torch.no_grad():
        net.conv1.bias.normal_(0, 0.1)
        net.conv2.bias.normal_(0, 0.1)
    eng = GCNEngine(net, None, loss_mode="soft", precision="bf16", activations="bf16", preaggregate=preaggregate)
    Z = eng.forward_logits(batch, X).double().cpu()
    p = rs.GCNParams(*[t.detach().double().cpu() for t in (net.conv1.weight, net.conv1.bias, net.conv2.weight, net.conv2.bias)])
    Xh = X.double().cpu()
    want, exact = [], []
    for g in range(batch.num_graphs):
        lo, hi = int(gp[g]), int(gp[g + 1])
        csr = rs.HostCSR((rowptr[lo: hi + 1] - rowptr[lo]).astype(np.int32), (colidx[rowptr[lo]: rowptr[hi]] - lo).astype(np.int32),
                         np.ones(int(rowptr[hi] - rowptr[lo]), dtype=np.float32), hi - lo)
        want.append(rs.gcn_forward_bf16_storage(csr, Xh[lo:hi], p, preaggregated=preaggregate)["Z"])
        exact.append(rs.gcn_forward(csr, Xh[lo:hi], p)["Z"])
    want, exact = torch.cat(want), torch.cat(exact)
    scale = float(exact.abs().max())
    err_emul = float((Z - want).abs().max()) / scale
    err_fp32 = float((Z - exact).abs().max()) / scale
    assert err_emul < 2e-3, (err_emul, err_fp32)
    assert err_fp32 < 1e-2
    assert float((Z - want).abs().mean()) < 0.2 * float((Z - exact).abs().mean()) + 1e-7     # the emulation explains the gap


def test_bf16_activations_need_bf16_gemms_and_fall_back_without_a_plan():
    from Training import TrainingNeural as T
    import networkx as nx
    from gmc_b200.graph import CSRGraph
    cfg = T.TrainingConfig(n_nodes=200, dim_embedding=200, hidden_dim=64, gemm_precision="bf16", loss_mode="soft")
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    with pytest.raises(ValueError):
        GCNEngine(net, opt, precision="tf32", activations="bf16")
    with pytest.raises(ValueError):
        GCNEngine(net, opt, precision="bf16", activations="fp8")
    g = nx.barabasi_albert_graph(200, 3, seed=2)            # irregular graph: the engine keeps fp32 activations
    for u, v in g.edges():
        g[u][v]["weight"] = 1
    batch = GraphBatch([CSRGraph.from_networkx(g)])
    X = ops.densify(batch, 200)
    a = GCNEngine(net, None, loss_mode="soft", precision="bf16", activations="bf16").loss_and_grads(batch, X).cpu()
    b = GCNEngine(net, None, loss_mode="soft", precision="bf16", activations="fp32").loss_and_grads(batch, X).cpu()
    assert torch.equal(a, b)


def test_empty_and_bad_inputs_through_the_raw_abi():
    """n_rows = 0 is a no-op for every entry point of the bf16 chain; null pointers and misaligned leading dimensions are
    argument errors (negative codes), never launches."""
    import ctypes
    from gmc_b200 import _lib
    L = _lib.lib()
    z = ctypes.c_void_p(0)
    b16 = torch.zeros(64, dtype=torch.bfloat16, device=DEV)
    f32 = torch.zeros(64, device=DEV)
    i32 = torch.zeros(8, dtype=torch.int32, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert L.gmc_gemm_bf16_bf16out(0, p(b16), p(b16), p(b16), 0, 8, 8, 8, 8, 8, z, 0, z, z, 0, 0, z) == 0
    assert L.gmc_spmm_fused_skinny_bf16(p(i32), p(i32), z, z, z, p(b16), p(b16[32:]), 1, 0, 16, 16, 16, z, 0, p(f32), 3,
                                        p(f32), 3, z) == 0
    assert L.gmc_skinny_fwd_bf16(p(b16), 16, p(f32), p(f32), 3, 0, 16, 3, z) == 0
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device=DEV)
    assert L.gmc_skinny_bwd_bf16(p(f32), 3, p(f32), p(b16), 16, p(b16[32:]), 16, p(f32), p(f32[48:]), 0, 16, 3, p(ws),
                                 ws.numel(), z) == 0
    assert L.gmc_spmm_batched_bf16(p(i32), 0, 0, z, p(b16), p(b16[32:]), 0, 16, 16, 16, z, 0, z) == 0
    assert L.gmc_csr_scatter_bf16(p(i32), p(i32), z, p(i32), 0, 0, 16, p(b16), 16, 0, z) == 0
    assert L.gmc_csr_preaggregate_bf16(p(i32), p(i32), p(f32), z, p(i32), 0, 0, 16, p(b16), 16, z, 0, z) == 0
    assert L.gmc_csr_preaggregate_workspace_bytes(0) == 0 and L.gmc_csr_preaggregate_workspace_bytes(77) == 77
    # argument errors
    assert L.gmc_gemm_bf16_bf16out(0, z, p(b16), p(b16), 8, 8, 8, 8, 8, 8, z, 0, z, z, 0, 0, z) < 0            # null A
    assert L.gmc_gemm_bf16_bf16out(0, p(b16), p(b16), p(b16), 8, 8, 8, 8, 8, 12, z, 0, z, z, 0, 0, z) < 0      # ldc % 8 != 0
    assert L.gmc_skinny_fwd_bf16(p(b16), 16, p(f32), p(f32), 5, 4, 16, 5, z) < 0                               # n_out > 4
    assert L.gmc_csr_preaggregate_bf16(p(i32), p(i32), p(f32), z, p(i32), 1, 4, 16, p(b16), 12, z, 0, z) < 0   # ldx % 8 != 0
    assert "gmc_csr_preaggregate_bf16" in _lib.last_error()
    torch.cuda.synchronize()
