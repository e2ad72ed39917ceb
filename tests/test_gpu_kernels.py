"""GPU parity tests (`-m gpu`): every kernel, called through the C ABI (gmc_b200.ops -> ctypes),
against the CPU oracle (oracle/ref_step.py, oracle/postproc.c) on the same seeded inputs.
Tolerances: fp32 paths rel 1e-4 (north_star), usually far tighter; integer paths bit-exact."""
import math
import os

import networkx as nx
import numpy as np
import pytest
import torch

from gmc_b200 import _lib, ops, synth
from gmc_b200.graph import CSRGraph, GraphBatch, ZeroDegreeError
from oracle import postproc as pp
from oracle import ref_step as rs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def nx_regular(n, d, seed, weights=None):
    g = nx.random_regular_graph(d=d, n=n, seed=seed)
    rng = np.random.default_rng(seed)
    for u, v in g.edges():
        g[u][v]["weight"] = int(rng.integers(1, 5)) if weights else 1
    return g


def make_batch(specs, weights=False):
    graphs = [nx_regular(n, d, s, weights) for n, d, s in specs]
    csrs = [rs.csr_from_networkx(g) for g in graphs]
    batch = GraphBatch([CSRGraph.from_networkx(g) for g in graphs])
    return graphs, csrs, batch


def ahat_dense(csr, dtype=torch.float64):
    A = (rs.dense_adjacency(csr, dtype=dtype) != 0).to(dtype)
    dinv = torch.diag(A.sum(1).clamp(min=1).pow(-0.5))
    return dinv @ A @ dinv


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


# ------------------------------------------------------------------ SpMM
@pytest.mark.parametrize("C", [3, 1, 4, 16, 20, 64, 100, 500, 1000, 37])
@pytest.mark.parametrize("use_coef", [True, False])
def test_spmm_matches_dense_ahat(C, use_coef):
    graphs, csrs, batch = make_batch([(40, 5, 1), (64, 7, 2), (30, 6, 3), (50, 8, 4)], weights=True)
    torch.manual_seed(C)
    X = torch.randn(batch.num_nodes, C)
    bias = torch.randn(C)
    want = torch.cat([ahat_dense(c) @ X[s:e].double() for c, (s, e) in
                      zip(csrs, zip(batch.graph_ptr_host[:-1], batch.graph_ptr_host[1:]))])
    got = ops.spmm(batch, X.to(DEV), use_coef=use_coef)
    assert relerr(got.cpu(), want) < 2e-6
    got2 = ops.spmm(batch, X.to(DEV), bias=bias.to(DEV), relu=True, use_coef=use_coef)
    assert relerr(got2.cpu(), torch.relu(want + bias.double())) < 2e-6


@pytest.mark.parametrize("C", [500, 64, 100, 128, 28, 36, 1000])
@pytest.mark.parametrize("specs", [[(1000, 7, 1), (1000, 6, 2)], [(200, 5, 3), (128, 7, 4), (1000, 8, 5), (300, 6, 6)],
                                   [(1800, 3, 7), (130, 4, 8)], [(3000, 4, 9)]])
def test_spmm_slab_kernel_matches_row_kernel_and_dense(C, specs):
    """Graphs with >= 128 nodes take the shared-memory slab kernel (W4 = 7, 4 or 2 by size)."""
    graphs, csrs, batch = make_batch(specs)
    assert batch.build_plan()                                            # opt in to the slab kernel
    torch.manual_seed(C)
    X = torch.randn(batch.num_nodes, C, device=DEV)
    bias = torch.randn(C, device=DEV)
    slab = ops.spmm(batch, X, bias=bias, relu=True)                      # gmc_spmm_batched_f32
    rowk = ops.spmm(batch, X, bias=bias, relu=True, use_coef=False)      # warp-per-row kernel
    assert relerr(slab.cpu(), rowk.cpu()) < 2e-6
    gp = batch.graph_ptr_host
    i = len(csrs) - 1
    want = torch.relu(ahat_dense(csrs[i]) @ X[gp[i]: gp[i + 1]].cpu().double() + bias.cpu().double())
    assert relerr(slab[gp[i]: gp[i + 1]].cpu(), want) < 2e-6
    plain = ops.spmm(batch, X)
    assert relerr(plain.cpu(), ops.spmm(batch, X, use_coef=False).cpu()) < 2e-6


def test_spmm_slab_irregular_degrees_and_nonfinite_isolation():
    # irregular degrees inside one warp (different trip counts per row group) and a NaN in one graph must
    # not leak into rows that do not reference it
    g1 = nx.gnm_random_graph(400, 2400, seed=1)
    g1.remove_nodes_from([n for n, d in g1.degree() if d == 0])
    g1 = nx.convert_node_labels_to_integers(g1)
    g2 = nx.random_regular_graph(d=3, n=256, seed=2)
    for g in (g1, g2):
        nx.set_edge_attributes(g, 1, "weight")
    csrs = [rs.csr_from_networkx(g) for g in (g1, g2)]
    batch = GraphBatch([CSRGraph.from_networkx(g) for g in (g1, g2)])
    assert not batch.build_plan()                          # degrees > 8: no ELL plan, warp-per-row kernel
    torch.manual_seed(0)
    X = torch.randn(batch.num_nodes, 96, device=DEV)
    got = ops.spmm(batch, X)
    want = torch.cat([ahat_dense(c) @ X[s:e].cpu().double() for c, (s, e) in
                      zip(csrs, zip(batch.graph_ptr_host[:-1], batch.graph_ptr_host[1:]))])
    assert relerr(got.cpu(), want) < 2e-6
    n1 = csrs[0].n
    X2 = X.clone()
    X2[0, :] = float("nan")                                # node 0 of graph 1
    got2 = ops.spmm(batch, X2)
    nbrs = set(csrs[0].colidx[csrs[0].rowptr[0]: csrs[0].rowptr[1]].tolist())
    touched = torch.isnan(got2).any(dim=1).cpu().numpy()
    expect = np.zeros(batch.num_nodes, dtype=bool)
    rows = np.repeat(np.arange(n1), np.diff(csrs[0].rowptr))
    expect[rows[csrs[0].colidx == 0]] = True
    assert (touched == expect).all()


def test_spmm_plan_needs_uniform_neighbour_norms():
    """The slab plan keeps ONE coefficient per row (nd[v] * ns[u]), so a graph whose neighbours of some node have
    different degrees must be refused (-> warp-per-row kernel), even when its max degree is <= 8."""
    g = nx.random_regular_graph(d=4, n=200, seed=5)
    g.remove_edge(*next(iter(g.edges())))                   # two nodes of degree 3 among degree-4 nodes
    nx.set_edge_attributes(g, 1, "weight")
    csr = rs.csr_from_networkx(g)
    batch = GraphBatch([CSRGraph.from_networkx(g)])
    assert batch.plan is None and not batch.build_plan()
    torch.manual_seed(0)
    X = torch.randn(200, 64, device=DEV)
    assert relerr(ops.spmm(batch, X).cpu(), ahat_dense(csr) @ X.cpu().double()) < 2e-6
    # a batch of >= 32 regular graphs gets its plan automatically, and mixing degrees ACROSS graphs is fine
    graphs, csrs, big = make_batch([(128 + 2 * (i % 3), 5 + i % 3, i) for i in range(32)])
    assert big.plan is not None
    X = torch.randn(big.num_nodes, 100, device=DEV)
    want = torch.cat([ahat_dense(c) @ X[s:e].cpu().double() for c, (s, e) in
                      zip(csrs, zip(big.graph_ptr_host[:-1], big.graph_ptr_host[1:]))])
    assert relerr(ops.spmm(big, X).cpu(), want) < 2e-6


@pytest.mark.parametrize("C,K", [(500, 3), (64, 2), (128, 8), (20, 3), (512, 5)])
def test_spmm_fused_skinny_projection(C, K):
    graphs, csrs, batch = make_batch([(70, 6, 1), (33 * 2, 7, 2), (1000, 7, 3), (9, 2, 4)])
    torch.manual_seed(C + K)
    X = torch.randn(batch.num_nodes, C, device=DEV)
    W = torch.randn(C, K, device=DEV)
    bias = torch.randn(C, device=DEV)
    Y, T = ops.spmm_fused_skinny(batch, X, W, bias=bias, relu=True)
    want_Y = ops.spmm(batch, X, bias=bias, relu=True, use_coef=False)
    assert torch.equal(Y, want_Y) or relerr(Y.cpu(), want_Y.cpu()) < 1e-6
    assert relerr(T.cpu(), want_Y.cpu().double() @ W.cpu().double()) < 3e-6
    with pytest.raises(_lib.GmcError):
        ops.spmm_fused_skinny(batch, torch.randn(batch.num_nodes, 516, device=DEV), torch.randn(516, 3, device=DEV))


def test_spmm_strided_rows_and_symmetry():
    _, csrs, batch = make_batch([(100, 7, 9)])
    torch.manual_seed(0)
    big = torch.randn(100, 96, device=DEV)
    X = big[:, 16:80]                                   # ld = 96, 64 columns, 16-byte aligned offset
    out_big = torch.zeros(100, 128, device=DEV)
    ops.spmm(batch, X, out=out_big[:, :64])
    want = ahat_dense(csrs[0]) @ X.cpu().double()
    assert relerr(out_big[:, :64].cpu(), want) < 2e-6 and float(out_big[:, 64:].abs().sum()) == 0.0
    # <Y, A X> == <A Y, X>: the backward SpMM is the forward SpMM
    Y = torch.randn(100, 64, device=DEV)
    lhs = (Y * ops.spmm(batch, X.contiguous())).sum().item()
    rhs = (ops.spmm(batch, Y) * X).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * abs(lhs)


def test_zero_degree_raises_like_dgl():
    g = nx.Graph()
    g.add_nodes_from(range(5))
    g.add_edge(0, 1, weight=1)
    with pytest.raises(ZeroDegreeError):
        GraphBatch([CSRGraph.from_networkx(g)])
    b = GraphBatch([CSRGraph.from_networkx(g)], check_degrees=False)       # norms clamp to 1
    assert torch.allclose(b.norm.cpu(), torch.tensor([1.0, 1.0, 1.0, 1.0, 1.0]))


def test_densify_equals_extender_features():
    graphs, csrs, batch = make_batch([(20, 3, 1), (36, 5, 2)], weights=True)
    X = ops.densify(batch, 48).cpu()
    want = torch.cat([rs.dense_adjacency(c, 48) for c in csrs])
    assert torch.equal(X, want)


# ------------------------------------------------------------------ GEMM (fp32 parity path)
@pytest.mark.parametrize("op,M,N,K", [
    ("nn", 300, 500, 1000), ("nn", 129, 7, 33), ("nn", 1, 1, 1), ("nn", 257, 130, 8),
    ("nt", 200, 96, 500), ("nt", 65, 33, 17),
    ("tn", 1000, 500, 3000), ("tn", 100, 64, 70000), ("tn", 37, 5, 9),
])
def test_gemm_fp32_matches_float64(op, M, N, K):
    torch.manual_seed(M + N + K)
    if op == "nn":
        A, B = torch.randn(M, K), torch.randn(K, N)
        want = A.double() @ B.double()
    elif op == "nt":
        A, B = torch.randn(M, K), torch.randn(N, K)
        want = A.double() @ B.double().t()
    else:
        A, B = torch.randn(K, M), torch.randn(K, N)
        want = A.double().t() @ B.double()
    got = ops.gemm(op, A.to(DEV), B.to(DEV))
    assert relerr(got.cpu(), want) < 5e-6 * max(1.0, math.sqrt(K) / 30)
    acc = torch.ones(M, N, device=DEV)
    ops.gemm(op, A.to(DEV), B.to(DEV), out=acc, accumulate=True)
    assert relerr(acc.cpu(), want + 1.0) < 5e-6 * max(1.0, math.sqrt(K) / 30)


def test_gemm_unaligned_leading_dimensions():
    torch.manual_seed(3)
    A = torch.randn(50, 45, device=DEV)[:, 1:44]          # ld 45, offset 1 -> scalar path
    B = torch.randn(43, 21, device=DEV)
    got = ops.gemm("nn", A, B)
    assert relerr(got.cpu(), A.cpu().double() @ B.cpu().double()) < 5e-6


def test_gemm_tn_is_deterministic():
    torch.manual_seed(1)
    A, B = torch.randn(50000, 64, device=DEV), torch.randn(50000, 48, device=DEV)
    a = ops.gemm("tn", A, B)
    b = ops.gemm("tn", A, B)
    assert torch.equal(a, b)


def test_gemm_rejects_bad_arguments():
    A = torch.randn(4, 5, device=DEV)
    with pytest.raises(ValueError):
        ops.gemm("nn", A, torch.randn(6, 3, device=DEV))
    with pytest.raises(TypeError):
        ops.gemm("nn", A.double(), torch.randn(5, 3, device=DEV))
    with pytest.raises(TypeError):
        ops.gemm("nn", A.cpu(), torch.randn(5, 3))


# ------------------------------------------------------------------ skinny layer / colsum
@pytest.mark.parametrize("n_in,n_out", [(500, 3), (16, 3), (24, 2), (128, 8), (50, 3)])
def test_skinny_forward_backward(n_in, n_out):
    torch.manual_seed(n_in)
    n = 777
    H = torch.relu(torch.randn(n, n_in))
    W = torch.randn(n_in, n_out)
    dT = torch.randn(n, n_out)
    got = ops.skinny_fwd(H.to(DEV), W.to(DEV))
    assert relerr(got.cpu(), H.double() @ W.double()) < 3e-6
    if n_in % 4 == 0:
        dH, dW, db = ops.skinny_bwd(dT.to(DEV), W.to(DEV), H.to(DEV))
        want_dH = (dT.double() @ W.double().t()) * (H > 0)
        assert relerr(dH.cpu(), want_dH) < 3e-6
        assert relerr(dW.cpu(), H.double().t() @ dT.double()) < 3e-6
        assert relerr(db.cpu(), want_dH.sum(0)) < 3e-6
        dH2, dW2, db2 = ops.skinny_bwd(dT.to(DEV), W.to(DEV), H.to(DEV))
        assert torch.equal(dW, dW2) and torch.equal(db, db2)             # deterministic reduction


@pytest.mark.parametrize("n,c", [(1000, 3), (5, 3), (70000, 3), (3000, 500), (17, 33)])
def test_colsum(n, c):
    torch.manual_seed(n)
    X = torch.randn(n, c)
    got = ops.colsum(X.to(DEV))
    assert np.abs(got.cpu().numpy() - X.double().sum(0).numpy()).max() < 1e-4 * math.sqrt(n)


# ------------------------------------------------------------------ fused loss
@pytest.mark.parametrize("mode,override,penalty", [("ste", True, 0.0), ("ste", False, 0.0), ("soft", True, 0.0),
                                                   ("soft", False, 7.0), ("ste", True, 5.0), ("soft", True, 3.0)])
@pytest.mark.parametrize("weighted", [False, True])
def test_cut_loss_matches_oracle(mode, override, penalty, weighted):
    specs = [(40, 5, 1), (64, 7, 2), (50, 8, 4)] if weighted else [(40, 5, 1), (64, 7, 2), (30, 6, 3), (34, 8, 4)]
    graphs, csrs, batch = make_batch(specs, weights=weighted)
    torch.manual_seed(5)
    Z = torch.randn(batch.num_nodes, 3) * 2
    loss, P, dZ = ops.cut_loss(batch, Z.to(DEV), mode=mode, override_terminals=override, penalty=penalty, C=1.5)
    gp = batch.graph_ptr_host
    for i, csr in enumerate(csrs):
        Zi = Z[gp[i]: gp[i + 1]]
        Pi = torch.softmax(Zi.double(), dim=1)
        want = rs.ste_loss_and_grads(csr, Pi, C=1.5, override_terminals=override, mode=mode, penalty=penalty)
        assert abs(loss[i].item() - float(want["loss"])) <= 1e-5 * max(1.0, abs(float(want["loss"])))
        assert relerr(P[gp[i]: gp[i + 1]].cpu(), Pi) < 1e-6
        assert np.abs(dZ[gp[i]: gp[i + 1]].cpu().numpy() - want["dZ"].numpy()).max() < 2e-5 * (1 + penalty)
    if mode == "ste" and penalty == 0.0:
        # integrality: loss == -C * integer cut of the hard labels
        labels = ops.argmax_labels(batch, P, force_terminals=override)
        if not weighted or batch.wts_i32 is not None:
            cuts = ops.cut_value(batch, labels)
            assert torch.equal((loss / -1.5).round().long().cpu(), cuts.cpu())


@pytest.mark.parametrize("mode,override,penalty", [("ste", True, 0.0), ("ste", False, 0.0), ("soft", True, 0.0),
                                                   ("soft", False, 7.0), ("ste", True, 5.0), ("soft", True, 3.0)])
@pytest.mark.parametrize("weighted,K", [(False, 3), (True, 3), (False, 5), (False, 2)])
def test_layer2_loss_fused_equals_the_separate_kernels(mode, override, penalty, weighted, K):
    """One launch (one CTA per graph) == spmm + cut_loss + colsum + spmm: Z, P, dZ, per-graph loss, db2, dT2."""
    if K == 2 and override:
        pytest.skip("terminal override needs >= 3 classes")
    specs = [(40, 5, 1), (64, 7, 2), (50, 8, 4), (300, 6, 9)] if weighted else [(40, 5, 1), (64, 7, 2), (4, 3, 3), (34, 8, 4), (1000, 7, 5)]
    graphs, csrs, batch = make_batch(specs, weights=weighted)
    torch.manual_seed(7)
    T2 = (torch.randn(batch.num_nodes, K) * 1.5).to(DEV)
    b2 = (torch.randn(K) * 0.3).to(DEV)
    Z = ops.spmm(batch, T2, bias=b2)
    loss, P, dZ = ops.cut_loss(batch, Z, mode=mode, override_terminals=override, penalty=penalty, C=1.5)
    db2 = ops.colsum(dZ)
    dT2 = ops.spmm(batch, dZ)
    N = batch.num_nodes
    outs = []
    for _ in range(2):
        P2, dZ2, Z2 = (torch.empty(N, K, device=DEV) for _ in range(3))
        dT2b, db2b = torch.empty(N, K, device=DEV), torch.empty(K, device=DEV)
        loss2 = ops.layer2_loss_fused(batch, T2, b2, mode=mode, override_terminals=override, penalty=penalty, C=1.5, P=P2,
                                      dZ=dZ2, dT2=dT2b, db2=db2b, Z=Z2)
        outs.append((loss2.clone(), P2, dZ2, Z2, dT2b, db2b))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)                                   # no atomics: bitwise reproducible
    loss2, P2, dZ2, Z2, dT2b, db2b = outs[0]
    assert relerr(Z2.cpu(), Z.cpu()) < 1e-6 and relerr(P2.cpu(), P.cpu()) < 1e-6
    np.testing.assert_allclose(loss2.cpu().numpy(), loss.cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert float((dZ2 - dZ).abs().max()) < 2e-5 * (1 + penalty)
    assert float((dT2b - dT2).abs().max()) < 2e-5 * (1 + penalty)
    assert float((db2b - db2).abs().max()) < 1e-4 * (1 + penalty) * max(1.0, float(db2.abs().max()))
    # inference form: only P and the loss
    P3 = torch.empty(N, K, device=DEV)
    loss3 = ops.layer2_loss_fused(batch, T2, b2, mode=mode, override_terminals=override, penalty=penalty, C=1.5, P=P3)
    assert torch.equal(P3, P2) and torch.equal(loss3, loss2)


def test_cut_loss_two_class_and_eight_class():
    _, csrs, batch = make_batch([(30, 4, 1)])
    for K in (2, 8):
        torch.manual_seed(K)
        Z = torch.randn(30, K)
        loss, P, dZ = ops.cut_loss(batch, Z.to(DEV), mode="ste", override_terminals=K >= 3)
        want = rs.ste_loss_and_grads(csrs[0], torch.softmax(Z.double(), 1), override_terminals=K >= 3)
        assert abs(loss[0].item() - float(want["loss"])) < 1e-6
        assert np.abs(dZ.cpu().numpy() - want["dZ"].numpy()).max() < 2e-5
    with pytest.raises(_lib.GmcError):
        ops.cut_loss(batch, torch.randn(30, 2, device=DEV), override_terminals=True)


@pytest.mark.parametrize("C", [500, 64, 28, 100])
def test_adjacency_feature_kernels_equal_dense_gemms(C):
    """X W and X^T dT for X = zero-padded adjacency rows, computed as aggregations over the ELL plan (spmm_adj.cu),
    equal the dense products exactly up to fp32 summation order; graphs smaller than the feature width, mixed
    degrees (padded ELL slots) and rows of dW beyond the largest graph (must be 0) included."""
    specs = [(1000, 7, 1), (200, 5, 2), (130, 8, 3), (1000, 6, 4), (400, 3, 5)] + [(128 + 2 * i, 4 + i % 5, 10 + i) for i in range(30)]
    graphs, csrs, batch = make_batch(specs)
    assert batch.plan is not None and ops.adjacency_kernels_apply(batch, 1000)
    F = 1000
    torch.manual_seed(C)
    W = ops.padded_empty(F, C, DEV); W.normal_()
    X = ops.densify(batch, F)
    T = ops.adj_features_fwd(batch, W)
    assert relerr(T.cpu(), X.cpu().double() @ W.cpu().double()) < 2e-6
    dT = torch.randn(batch.num_nodes, C, device=DEV)
    dW = torch.full((F, C), 7.0, device=DEV)
    ops.adj_features_bwd(batch, dT, out=dW)
    want = X.cpu().double().t() @ dT.cpu().double()
    assert relerr(dW.cpu(), want) < 2e-5
    assert float(dW[batch.max_nodes:].abs().sum()) == 0.0 or batch.max_nodes == F
    dW2 = torch.empty_like(dW)
    ops.adj_features_bwd(batch, dT, out=dW2)
    assert torch.equal(dW, dW2)                                              # deterministic


def test_ragged_tiny_graphs_and_empty_calls():
    """Edge cases the reference exercises implicitly: the smallest graphs that can carry three terminals (a triangle,
    K4) next to full-size ones in ONE batch, and empty inputs straight through the C ABI (every entry point must
    return GMC_OK without launching)."""
    import ctypes
    tri = nx.complete_graph(3)
    k4 = nx.complete_graph(4)
    big = nx.random_regular_graph(d=7, n=1000, seed=3)
    mid = nx.random_regular_graph(d=6, n=129, seed=4)
    graphs = [tri, big, k4, mid]
    for g in graphs:
        nx.set_edge_attributes(g, 1, "weight")
    csrs = [rs.csr_from_networkx(g) for g in graphs]
    batch = GraphBatch([CSRGraph.from_networkx(g) for g in graphs])
    gp = batch.graph_ptr_host
    torch.manual_seed(0)
    X = torch.randn(batch.num_nodes, 64, device=DEV)
    want = torch.cat([ahat_dense(c) @ X[s:e].cpu().double() for c, (s, e) in zip(csrs, zip(gp[:-1], gp[1:]))])
    assert relerr(ops.spmm(batch, X).cpu(), want) < 2e-6
    batch.build_plan()                                         # graphs below 128 nodes ride along in the slab kernel
    assert relerr(ops.spmm(batch, X).cpu(), want) < 2e-6
    Z = torch.randn(batch.num_nodes, 3) * 2
    loss, P, dZ = ops.cut_loss(batch, Z.to(DEV), mode="ste", override_terminals=True)
    labels = ops.argmax_labels(batch, P, force_terminals=True)
    cuts = ops.cut_value(batch, labels).cpu()
    for i, csr in enumerate(csrs):
        Pi = torch.softmax(Z[gp[i]: gp[i + 1]].double(), dim=1)
        w = rs.ste_loss_and_grads(csr, Pi, override_terminals=True)
        assert abs(loss[i].item() - float(w["loss"])) < 1e-6 and int(cuts[i]) == int(round(-float(w["loss"])))
        assert np.abs(dZ[gp[i]: gp[i + 1]].cpu().numpy() - w["dZ"].numpy()).max() < 2e-5
    assert int(cuts[0]) == 3 and int(cuts[2]) >= 5               # triangle: all three edges cut; K4: 3 terminals apart
    g_labels, g_cut, _ = ops.greedy_node_move(batch, labels, 3, 200, 3)
    for i, csr in enumerate(csrs):
        wl, wc = pp.greedy_node_move(csr.rowptr, csr.colidx, labels[gp[i]: gp[i + 1]].cpu().numpy(), 3, 200, 3)[:2]
        assert int(g_cut[i]) == int(wc) and (g_labels[gp[i]: gp[i + 1]].cpu().numpy() == wl).all()

    # ---- empty inputs through the raw ABI
    L = _lib.lib()
    z = ctypes.c_void_p(0)
    one = torch.zeros(8, device=DEV)
    i32 = torch.zeros(8, dtype=torch.int32, device=DEV)
    i64 = torch.zeros(8, dtype=torch.int64, device=DEV)
    f64 = torch.zeros(8, dtype=torch.float64, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert L.gmc_spmm_symnorm_f32(p(i32), p(i32), z, z, z, p(one), p(one[4:]), 0, 4, 4, 4, z, 0, z) == 0
    assert L.gmc_gemm_nn(p(one), p(one), p(one), 0, 4, 4, 4, 4, 4, 0, 1, z, 0, z) == 0
    assert L.gmc_gemm_nn(p(one), p(one), p(one), 0, 4, 4, 4, 4, 4, 0, 0, z, 0, z) == 0
    assert L.gmc_cut_value_i32(p(i32), p(i32), p(i32), z, p(i32), 0, 0, p(i64), z) == 0
    assert L.gmc_argmax_labels(p(one), 3, p(i32), 0, 0, 3, 1, p(i32), z) == 0
    assert L.gmc_softmax_cut_loss_fwd_bwd(p(one), 3, p(i32), p(i32), z, p(i32), 0, 0, 3, 0, 1, ctypes.c_float(0.0),
                                          ctypes.c_float(1.0), p(one), p(f64), p(one), z) == 0
    torch.cuda.synchronize()


def test_softmax_fwd_bwd():
    torch.manual_seed(0)
    Z = (torch.randn(1000, 3) * 5).requires_grad_()
    P = torch.softmax(Z, 1)
    dP = torch.randn(1000, 3)
    P.backward(dP)
    got = ops.softmax_fwd(Z.detach().to(DEV))
    assert relerr(got.cpu(), P.detach()) < 1e-6
    assert np.abs(ops.softmax_bwd(got, dP.to(DEV)).cpu().numpy() - Z.grad.numpy()).max() < 1e-6


# ------------------------------------------------------------------ Adam
def test_adam_matches_torch_adam_over_steps():
    torch.manual_seed(0)
    shapes = [(1000, 500), (500,), (500, 3), (3,), (7, 5)]
    ref = [torch.nn.Parameter(torch.randn(*s)) for s in shapes]
    mine = [p.detach().clone().to(DEV) for p in ref]
    m = [torch.zeros_like(p) for p in mine]
    v = [torch.zeros_like(p) for p in mine]
    opt = torch.optim.Adam(ref, lr=1e-3)
    step_dev = torch.zeros(1, dtype=torch.int64, device=DEV)
    state = torch.zeros(8, dtype=torch.int64, device=DEV)        # gmc_adam_multi_devstate: count + cached scalars
    mine2 = [p.clone() for p in mine]
    m2 = [torch.zeros_like(p) for p in mine]
    v2 = [torch.zeros_like(p) for p in mine]
    mine3 = [p.clone() for p in mine]
    m3 = [torch.zeros_like(p) for p in mine]
    v3 = [torch.zeros_like(p) for p in mine]
    for t in range(1, 8):
        grads = [torch.randn(*s) * (10.0 if t % 2 else 0.01) for s in shapes]
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        opt.step()
        gd = [g.to(DEV) for g in grads]
        ops.adam_multi(mine, gd, m, v, lr=1e-3, step=t)
        ops.adam_multi(mine2, gd, m2, v2, lr=1e-3, step_dev=step_dev)
        # the cached state survives a change of hyper-parameters (recomputed, not reused) and a host-side reset
        lr3 = 1e-3 if t != 4 else 2e-3
        if t == 4:
            ops.adam_multi([q.clone() for q in mine3], gd, [q.clone() for q in m3], [q.clone() for q in v3], lr=lr3, step_dev=state)
            state[:1].fill_(t - 1)
        ops.adam_multi(mine3, gd, m3, v3, lr=1e-3, step_dev=state)
        for a, b, c, d in zip(mine, ref, mine2, mine3):
            assert relerr(a.cpu(), b.detach()) < 2e-6
            assert relerr(c.cpu(), b.detach()) < 2e-6
            assert torch.equal(c, d)                                  # same double-precision scalars either way
    assert int(step_dev.item()) == 7 and int(state[0].item()) == 7 and int(state[3].item()) == 0


# ------------------------------------------------------------------ integer post-processing
def test_postproc_fixture_bit_exact():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "postproc.npz"))
    f32 = int(str(z["numpy_version"]).split(".")[0]) >= 2
    for tag in ("s", "m", "l", "w"):
        n = int(z[f"{tag}_n"])
        g = nx.Graph()
        g.add_nodes_from(range(n))
        for (u, v), w in zip(z[f"{tag}_edges"], z[f"{tag}_w"]):
            g.add_edge(int(u), int(v), weight=int(w))
        batch = GraphBatch([CSRGraph.from_networkx(g)])
        P = torch.from_numpy(z[f"{tag}_P"]).to(DEV)
        labels = ops.argmax_labels(batch, P)
        assert labels.cpu().tolist() == z[f"{tag}_simple"].tolist()
        assert int(ops.cut_value(batch, labels).item()) == int(z[f"{tag}_simple_cut"])
        np.random.seed(int(z[f"{tag}_post_seed"]))
        U = torch.from_numpy(np.random.rand(200 * (n - 3))).to(DEV)
        u_ptr = torch.tensor([0, U.numel()], dtype=torch.int64, device=DEV)
        best_labels, best, best_it = ops.sample_best_cut(batch, P, U, u_ptr, 200, compare_f32=f32)
        assert int(best.item()) == int(z[f"{tag}_post_cut"])
        assert best_labels.cpu().tolist() == z[f"{tag}_post_labels"].tolist()


@pytest.mark.parametrize("compare_f32", [True, False])
def test_sample_best_cut_batched_vs_oracle(compare_f32):
    specs = [(50, 6, 1), (3, 2, 2), (120, 7, 3), (77, 8, 4), (4, 3, 5)]
    graphs, csrs, batch = make_batch(specs, weights=True)
    rng = np.random.default_rng(0)
    P = torch.softmax(torch.from_numpy(rng.normal(0, 1.5, size=(batch.num_nodes, 3)).astype(np.float32)), 1)
    iters = 37
    counts = [iters * max(n - 3, 0) for n, _, _ in specs]
    u_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    U = rng.random(int(u_ptr[-1]))
    # adversarial uniforms: exactly on float32 cumulative boundaries
    Pn = P.numpy()
    U[5] = float(Pn[3 + 5 % 47, 0])
    U[6] = float(np.float32(Pn[9, 0]) + np.float32(Pn[9, 1]))
    labels, best, best_it = ops.sample_best_cut(batch, P.to(DEV), torch.from_numpy(U).to(DEV),
                                                torch.from_numpy(u_ptr).to(DEV), iters, compare_f32=compare_f32)
    gp = batch.graph_ptr_host
    for i, csr in enumerate(csrs):
        wl, wc, wi = pp.sample_best_cut(csr.rowptr, csr.colidx, Pn[gp[i]: gp[i + 1]], U[u_ptr[i]: u_ptr[i + 1]], iters,
                                        csr.weights.astype(np.int32), compare_f32=compare_f32)
        assert int(best[i].item()) == wc and int(best_it[i].item()) == wi
        assert labels[gp[i]: gp[i + 1]].cpu().tolist() == wl.tolist()


@pytest.mark.parametrize("K,n_frozen", [(3, 3), (2, 0), (5, 3), (8, 1)])
def test_greedy_node_move_bit_exact(K, n_frozen):
    specs = [(60, 6, 1), (200, 7, 2), (31 * 2, 5, 3), (8, 3, 4)]
    graphs, csrs, batch = make_batch(specs, weights=True)
    rng = np.random.default_rng(K)
    start = rng.integers(0, K, size=batch.num_nodes).astype(np.int32)
    for iters in (0, 1, 7, 200):
        out, cut, moves = ops.greedy_node_move(batch, torch.from_numpy(start).to(DEV), K, iters, n_frozen)
        gp = batch.graph_ptr_host
        for i, csr in enumerate(csrs):
            wl, wc, wm = pp.greedy_node_move(csr.rowptr, csr.colidx, start[gp[i]: gp[i + 1]], K, iters, n_frozen,
                                             csr.weights.astype(np.int32))
            assert int(cut[i].item()) == wc and int(moves[i].item()) == wm
            assert out[gp[i]: gp[i + 1]].cpu().tolist() == wl.tolist()


def test_known_answer_cut_values_on_device():
    # the reference's seeded randomized-maxcut answers (randomizedAlgo.ipynb:L113,L186) through the GPU evaluator
    for n, seed, kw, expected in ((500, 42, dict(max_iterations=1000, threshold=1, patience=50), 1393),
                                  (1000, 42, dict(max_iterations=2000, threshold=1, patience=100), 2741)):
        g = nx.random_regular_graph(d=8, n=n, seed=seed)
        cut, part = pp.py_randomized_k_way_maxcut(g, k=3, random_seed=seed, **kw)
        batch = GraphBatch([CSRGraph.from_networkx(g)])
        labels = torch.tensor([part[i] for i in range(n)], dtype=torch.int32, device=DEV)
        assert int(ops.cut_value(batch, labels).item()) == expected == cut


# ------------------------------------------------------------------ full-size properties (config 3 shape)
def test_full_size_properties_n1000_d7():
    B, n, d = 64, 1000, 7
    batch = synth.regular_batch(B, n, d, seed=11)
    assert batch.num_nodes == B * n and batch.nnz == B * n * d
    # A_hat 1 = 1 for a regular graph (row sums of D^-1/2 A D^-1/2)
    ones = torch.ones(B * n, 500, device=DEV)
    Y = ops.spmm(batch, ones)
    assert float((Y - 1).abs().max()) < 1e-5
    # linearity
    torch.manual_seed(0)
    X1, X2 = torch.randn(B * n, 500, device=DEV), torch.randn(B * n, 500, device=DEV)
    lin = ops.spmm(batch, X1 + 2 * X2) - (ops.spmm(batch, X1) + 2 * ops.spmm(batch, X2))
    assert float(lin.abs().max()) < 1e-4
    # cut of a uniform random labelling: 0 <= cut <= |E| and a checksum of checksums against the oracle evaluator
    rng = np.random.default_rng(1)
    labels = rng.integers(0, 3, size=B * n).astype(np.int32)
    cuts = ops.cut_value(batch, torch.from_numpy(labels).to(DEV)).cpu().numpy()
    assert (cuts >= 0).all() and (cuts <= n * d // 2).all()
    rp, ci = batch.rowptr.cpu().numpy(), batch.colidx.cpu().numpy()
    assert int(cuts.sum()) == pp.cut_value(rp, ci, labels)
    # greedy never decreases the cut, keeps terminals, and is idempotent at its fixed point
    start = torch.from_numpy(labels).to(DEV)
    out, cut2, moves = ops.greedy_node_move(batch, start, 3, 200, 3)
    assert (cut2.cpu().numpy() >= cuts).all()
    gp = batch.graph_ptr_host[:-1]
    assert torch.equal(out.cpu()[gp], start.cpu()[gp]) and torch.equal(out.cpu()[gp + 2], start.cpu()[gp + 2])
    out3, cut3, moves3 = ops.greedy_node_move(batch, out, 3, 10**4, 3)
    out4, cut4, moves4 = ops.greedy_node_move(batch, out3, 3, 200, 3)
    assert int(moves4.sum()) == 0 and torch.equal(out3, out4) and torch.equal(cut3, cut4)


# ------------------------------------------------------------------ gradient exchange over peer memory (csrc/peer.cu)
@pytest.mark.parametrize("world,n", [(2, 502003), (4, 1003), (8, 70001), (3, 17)])
def test_peer_allreduce_protocol_on_one_device(world, n):
    """The one-kernel all-reduce with its two flag barriers, exercised with `world` ranks living on ONE device: every
    rank's kernel runs on its own stream (they wait for each other inside the kernel, so they must be co-resident -- at
    most 64 blocks each), buffers and flags are ordinary allocations of this process standing in for the cudaIpc mappings.
    Result: every buffer holds the rank-ordered sum, bit for bit the same in all of them, over several epochs."""
    import ctypes
    from gmc_b200 import _lib
    lib = _lib.lib()
    pad = (n + 3) // 4 * 4
    bufs = [torch.zeros(pad, device=DEV) for _ in range(world)]
    flags = [torch.zeros(lib.gmc_peer_flag_bytes() // 4, dtype=torch.int32, device=DEV) for _ in range(world)]
    bp = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
    fp = (ctypes.c_void_p * world)(*[f.data_ptr() for f in flags])
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.manual_seed(world * 1000 + n)
    for epoch in range(1, 5):
        data = [torch.randn(n, device=DEV) * (10.0 ** (q % 3)) for q in range(world)]
        for q in range(world):
            bufs[q][:n].copy_(data[q])
        torch.cuda.synchronize()
        for q in range(world):
            _lib.check(lib.gmc_peer_allreduce_f32(bp, fp, world, q, n, epoch, streams[q].cuda_stream), "gmc_peer_allreduce_f32")
        torch.cuda.synchronize()
        want = data[0].clone()
        for q in range(1, world):
            want += data[q]                                    # the kernel's order: ranks 0 .. W-1, one fp32 add each
        for q in range(world):
            assert torch.equal(bufs[q][:n], want), (epoch, q)
        assert all(int(f[32].item()) == 0 for f in flags)      # block counters reset for the next launch


def _peer_worker(rank, world, port, n, out_q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gmc_b200 import dist as gdist
        torch.cuda.set_device(0)
        peer = gdist.PeerAllReduce(n, torch.device("cuda", 0))
        ok = True
        for epoch in range(3):
            torch.manual_seed(100 * epoch + rank)
            mine = torch.randn(n, device="cuda")
            peer.tensor[:n].copy_(mine)
            torch.cuda.synchronize()
            dist.barrier()
            peer.all_reduce_()
            torch.cuda.synchronize()
            want = torch.zeros(n, device="cuda")
            for q in range(world):
                torch.manual_seed(100 * epoch + q)
                want += torch.randn(n, device="cuda")
            ok = ok and bool(torch.equal(peer.tensor[:n], want))
            dist.barrier()
        out_q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_peer_allreduce_across_processes_through_cuda_ipc():
    """dist.PeerAllReduce as the engine uses it: separate PROCESSES exchange cudaIpc handles through torch.distributed
    (gloo here, both ranks on the one GPU of the test box -- NCCL refuses two ranks per device, cudaIpc does not) and run
    the one-kernel exchange on each other's mapped buffers."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n = 2, 50003
    port = 31000 + os.getpid() % 1000
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in range(world):
            results.append(q.get(timeout=120))
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert sorted(results) == [(0, True), (1, True)]


@pytest.mark.parametrize("seed", list(range(8)))
def test_integer_postprocessing_on_random_irregular_weighted_graphs(seed):
    """Differential sweep against oracle/postproc.c on graphs the seeded cases do not cover: irregular degrees (a ring for
    connectivity plus random chords), small integer weights (many tied gains for the greedy tie-break: lowest node, then
    lowest class), sizes from 3 nodes up, K = 2 .. 5 -- cut evaluator, greedy node moves and best-of-N sampling, bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    graphs = []
    for _ in range(6):
        n = int(rng.integers(3, 90))
        g = nx.cycle_graph(n)
        for _ in range(int(rng.integers(0, 3 * n))):
            u, v = int(rng.integers(0, n)), int(rng.integers(0, n))
            if u != v:
                g.add_edge(u, v)
        for u, v in g.edges():
            g[u][v]["weight"] = int(rng.integers(1, 4))
        graphs.append(g)
    csrs = [rs.csr_from_networkx(g) for g in graphs]
    batch = GraphBatch([CSRGraph.from_networkx(g) for g in graphs])
    gp = batch.graph_ptr_host
    K = int(rng.integers(2, 6))
    n_frozen = 3 if K >= 3 else 0
    start = rng.integers(0, K, size=batch.num_nodes).astype(np.int32)
    cuts = ops.cut_value(batch, torch.from_numpy(start).to(DEV)).cpu().numpy()
    for iters in (1, 200):
        out, cut, moves = ops.greedy_node_move(batch, torch.from_numpy(start).to(DEV), K, iters, n_frozen)
        for i, csr in enumerate(csrs):
            w = csr.weights.astype(np.int32)
            assert int(cuts[i]) == pp.cut_value(csr.rowptr, csr.colidx, start[gp[i]: gp[i + 1]], w)
            wl, wc, wm = pp.greedy_node_move(csr.rowptr, csr.colidx, start[gp[i]: gp[i + 1]], K, iters, n_frozen, w)
            assert int(cut[i].item()) == wc and int(moves[i].item()) == wm
            assert out[gp[i]: gp[i + 1]].cpu().tolist() == wl.tolist()
    # best-of-N categorical sampling (3 classes: the reference's assign_partitions), both numpy comparison modes
    P = torch.softmax(torch.from_numpy(rng.normal(0, 2.0, size=(batch.num_nodes, 3)).astype(np.float32)), 1)
    iters = 23
    counts = [iters * max(g.number_of_nodes() - 3, 0) for g in graphs]
    u_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    U = rng.random(int(u_ptr[-1]))
    for f32 in (True, False):
        labels, best, best_it = ops.sample_best_cut(batch, P.to(DEV), torch.from_numpy(U).to(DEV),
                                                    torch.from_numpy(u_ptr).to(DEV), iters, compare_f32=f32)
        for i, csr in enumerate(csrs):
            wl, wc, wi = pp.sample_best_cut(csr.rowptr, csr.colidx, P.numpy()[gp[i]: gp[i + 1]], U[u_ptr[i]: u_ptr[i + 1]],
                                            iters, csr.weights.astype(np.int32), compare_f32=f32)
            assert int(best[i].item()) == wc and int(best_it[i].item()) == wi
            assert labels[gp[i]: gp[i + 1]].cpu().tolist() == wl.tolist()


@pytest.mark.parametrize("burn,n", [(0, 1), (0, 312), (1, 311), (3, 313), (623, 2), (624, 5), (17, 100001), (0, 0),
                                     (1247, 625), (5, 2270000)])
def test_numpy_rand_stream_on_the_device(burn, n):
    """gmc_mt19937_uniform_f64 continues np.random's legacy MT19937 stream bit for bit from any position (pairs that straddle
    a state regeneration, odd word counts, a position of exactly 624) and hands the generator back in the state n scalar
    np.random.rand() calls would have left it in."""
    np.random.seed(1234)
    if burn:
        np.random.randint(0, 2 ** 31, size=burn)              # moves the position by `burn` 32-bit words
    st = np.random.get_state()
    want = np.random.rand(n)
    after = np.random.get_state()
    np.random.set_state(st)
    got = ops.numpy_rand_on_device(n, DEV)
    if n == 0:
        assert got is None
    else:
        assert np.array_equal(got.cpu().numpy(), want)
    now = np.random.get_state()
    assert now[0] == after[0] and np.array_equal(now[1], after[1]) and now[2] == after[2]
    assert np.random.rand() == (np.random.set_state(after) or np.random.rand())
