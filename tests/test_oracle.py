"""CPU tests: the oracle restatement (oracle/ref_step.py, oracle/postproc.{c,py})
pinned against (1) fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz), (2) the six seeded known answers
recorded in the reference's notebooks, (3) the dense D^-1/2 A D^-1/2 form."""
import os

import networkx as nx
import numpy as np
import pytest
import torch

from oracle import postproc as pp
from oracle import ref_step as rs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def graph_from_edges(edges, n, w=None):
    g = nx.Graph()
    g.add_nodes_from(range(int(n)))
    for i, (u, v) in enumerate(edges):
        g.add_edge(int(u), int(v), weight=int(w[i]) if w is not None else 1, capacity=1)
    return g


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


# ------------------------------------------------------------------ known answers
@pytest.mark.parametrize("n,seed,kw,expected", [
    (500, 42, dict(max_iterations=1000, threshold=1, patience=50), 1393),     # randomizedAlgo.ipynb:L113
    (1000, 42, dict(max_iterations=2000, threshold=1, patience=100), 2741),   # :L186
    (1000, 42, dict(max_iterations=2000, threshold=1, patience=100,
                    fixed_terminals={0: 0, 1: 1, 2: 2}), 2738),               # :L275
    (1000, 123, dict(max_iterations=2000, threshold=1, patience=50), 2724),   # :L486-497
    (1000, 123, dict(max_iterations=2000, threshold=1, patience=25), 2719),
    (1000, 123, dict(max_iterations=2000, threshold=1, patience=200), 2742),
])
def test_seeded_known_answers(n, seed, kw, expected):
    g = nx.random_regular_graph(d=8, n=n, seed=seed)
    for u, v in g.edges():
        g[u][v]["weight"] = 1
    cut, part = pp.py_randomized_k_way_maxcut(g, k=3, random_seed=seed, **kw)
    assert cut == expected
    csr = rs.csr_from_networkx(g)
    labels = np.asarray([part[i] for i in range(n)], dtype=np.int32)
    assert pp.cut_value(csr.rowptr, csr.colidx, labels) == expected      # C evaluator
    assert rs.cut_of_labels(csr, labels) == expected                      # float evaluator


def test_known_partition_sizes():
    # randomizedAlgo.ipynb:L113: {0: 153, 2: 176, 1: 171}
    g = nx.random_regular_graph(d=8, n=500, seed=42)
    _, part = pp.py_randomized_k_way_maxcut(g, k=3, max_iterations=1000, threshold=1, patience=50,
                                            random_seed=42)
    counts = np.bincount(list(part.values()), minlength=3)
    assert counts.tolist() == [153, 171, 176]


# ------------------------------------------------------------------ GraphConv / model
def test_graphconv_matches_dense_form():
    g = nx.random_regular_graph(d=5, n=40, seed=1)
    csr = rs.csr_from_networkx(g)
    torch.manual_seed(0)
    X = torch.randn(40, 32, dtype=torch.float64)
    W = torch.randn(32, 8, dtype=torch.float64)
    b = torch.randn(8, dtype=torch.float64)
    A = rs.dense_adjacency(csr, dtype=torch.float64)
    dinv = torch.diag(A.sum(1).clamp(min=1).pow(-0.5))
    want = dinv @ A @ dinv @ X @ W + b
    got = rs.graphconv(csr, X, W, b)
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
    # in <= out branch (aggregate first)
    W2 = torch.randn(32, 64, dtype=torch.float64)
    assert torch.allclose(rs.graphconv(csr, X, W2, None), dinv @ A @ dinv @ X @ W2, rtol=1e-12, atol=1e-12)


def test_graphconv_zero_degree_raises():
    g = nx.Graph()
    g.add_nodes_from(range(4))
    g.add_edge(0, 1, weight=1)
    csr = rs.csr_from_networkx(g)
    with pytest.raises(ValueError):
        rs.graphconv(csr, torch.ones(4, 3), torch.ones(3, 2), None)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_step_matches_reference_fixture(tag):
    z = load("gcn_step.npz")
    n = int(z[f"{tag}_n"])
    csr = rs.csr_from_networkx(graph_from_edges(z[f"{tag}_edges"], n))
    X = rs.dense_adjacency(csr, 1000)
    p = rs.GCNParams(*[torch.from_numpy(z[f"{tag}_{k}"].copy()) for k in
                       ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")])
    fwd = rs.gcn_forward(csr, X, p)
    np.testing.assert_allclose(fwd["P"].numpy(), z[f"{tag}_P"], rtol=1e-5, atol=1e-7)
    lg = rs.ste_loss_and_grads(csr, fwd["P"], C=1.0)
    assert abs(float(lg["loss"]) - float(z[f"{tag}_loss"])) <= 1e-4 * abs(float(z[f"{tag}_loss"]))
    assert float(lg["loss"]) == -rs.cut_of_labels(csr, rs.hard_labels(fwd["P"]))
    np.testing.assert_allclose(lg["g"].numpy(), z[f"{tag}_dP"], rtol=1e-5, atol=1e-6)
    grads = rs.gcn_backward(csr, X, p, fwd, lg["dZ"])
    for k, name in (("W1", "conv1.weight"), ("b1", "conv1.bias"), ("W2", "conv2.weight"), ("b2", "conv2.bias")):
        ref = z[f"{tag}_grad_{name}"]
        scale = np.abs(ref).max() + 1e-12
        assert np.abs(grads[k].numpy() - ref).max() <= 2e-5 * scale, name
    # one Adam step, then the reference's next two sequential steps
    st = rs.AdamState()
    q = p.clone()
    rs.adam_update(q.tensors(), [grads["W1"], grads["b1"], grads["W2"], grads["b2"]], st)
    for t, name in zip(q.tensors(), ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")):
        np.testing.assert_allclose(t.numpy(), z[f"{tag}_step1_{name}"], rtol=1e-5, atol=2e-7)
    losses = [float(lg["loss"])]
    for _ in range(2):
        losses.append(rs.train_step_closed_form(csr, X, q, st))
    np.testing.assert_allclose(losses, z[f"{tag}_losses3"], rtol=1e-4)
    for t, name in zip(q.tensors(), ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")):
        np.testing.assert_allclose(t.numpy(), z[f"{tag}_step3_{name}"], rtol=1e-4, atol=1e-6)


def test_single_step_at_baseline_shapes_matches_reference_fixture():
    """BASELINE config 1 at its real shapes (n = 500, 1000 features, hidden 500): the restatement against the
    reference's own forward, loss and autograd gradients (baseline_shapes.npz part a)."""
    from golden_util import assert_w1_digest, gcn_shapes, seeded_weights
    z = load("baseline_shapes.npz")
    csr = rs.csr_from_networkx(graph_from_edges(z["a_edges"], 500))
    X = rs.dense_adjacency(csr, 1000)
    w = seeded_weights(np.random.default_rng(int(z["a_weight_seed"])), gcn_shapes(1000, 500))
    p = rs.GCNParams(*[w[k] for k in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")])
    fwd = rs.gcn_forward(csr, X, p)
    np.testing.assert_allclose(fwd["P"].numpy(), z["a_P"], rtol=1e-4, atol=1e-7)
    lg = rs.ste_loss_and_grads(csr, fwd["P"], C=1.0)
    assert abs(float(lg["loss"]) - float(z["a_loss"])) <= 1e-4 * abs(float(z["a_loss"]))
    grads = rs.gcn_backward(csr, X, p, fwd, lg["dZ"])
    assert_w1_digest(grads["W1"].numpy(), z, "a_grad_conv1.weight", 1e-4)
    for k, name in (("b1", "conv1.bias"), ("W2", "conv2.weight"), ("b2", "conv2.bias")):
        ref = z[f"a_grad_{name}"]
        assert np.abs(grads[k].numpy() - ref).max() <= 1e-4 * (np.abs(ref).max() + 1e-12), name


def test_two_epochs_at_baseline_shapes_match_reference_fixture():
    """The 20-graph pipeline of complete_training_pipeline.ipynb cell 15 for two epochs (40 sequential Adam steps at
    n = 500, hidden 500): loss history and final weights of the restatement against the reference's train_model run."""
    from golden_util import assert_w1_digest, gcn_shapes, seeded_weights
    z = load("baseline_shapes.npz")
    w = seeded_weights(np.random.default_rng(int(z["b_weight_seed"])), gcn_shapes(1000, 500))
    p = rs.GCNParams(*[w[k] for k in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")])
    st = rs.AdamState()
    items = []
    for i in range(int(z["b_num_graphs"])):
        csr = rs.csr_from_networkx(graph_from_edges(z[f"b_g{i}_edges"], 500))
        items.append((csr, rs.dense_adjacency(csr, 1000)))
    hist = []
    for _ in range(2):
        hist.append(sum(rs.train_step_closed_form(csr, X, p, st) for csr, X in items))
    np.testing.assert_allclose(hist, z["b_loss_history"], rtol=1e-4)
    assert_w1_digest(p.W1.numpy(), z, "b_final_conv1.weight", 2e-4)
    for t, name in ((p.b1, "conv1.bias"), (p.W2, "conv2.weight"), (p.b2, "conv2.bias")):
        ref = z[f"b_final_{name}"]
        assert np.abs(t.numpy() - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-12), name


def test_faithful_port_matches_closed_form():
    g = nx.random_regular_graph(d=6, n=30, seed=3)
    csr = rs.csr_from_networkx(g)
    X = rs.dense_adjacency(csr, 1000)
    port = rs.FaithfulPort(1000, 8, 3, seed=4)
    p = rs.GCNParams(*[t.detach().clone() for t in port.params])
    st = rs.AdamState()
    for _ in range(3):
        a = port.step(csr, X, X)
        b = rs.train_step_closed_form(csr, X, p, st)
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b))
    for t, u in zip(port.params, p.tensors()):
        np.testing.assert_allclose(t.detach().numpy(), u.numpy(), rtol=1e-4, atol=1e-6)


def test_soft_mode_and_penalty_against_autograd():
    g = nx.random_regular_graph(d=5, n=24, seed=9)
    csr = rs.csr_from_networkx(g)
    A = rs.dense_adjacency(csr, dtype=torch.float64)
    torch.manual_seed(1)
    Z = torch.randn(24, 3, dtype=torch.float64, requires_grad=True)
    P = torch.softmax(Z, dim=1)
    s = P.clone()
    eye = torch.eye(3, dtype=torch.float64)
    for t in range(3):
        s[t] = eye[t] + P[t] - P[t].detach()
    pen = sum(torch.dot(s[i], s[j]) for i in range(3) for j in range(i + 1, 3))
    loss = -1.5 * (A * (1 - s @ s.t())).sum() / 2 + 7.0 * pen
    loss.backward()
    lg = rs.ste_loss_and_grads(csr, P.detach(), C=1.5, mode="soft", penalty=7.0)
    assert abs(float(lg["loss"]) - float(loss)) < 1e-9
    assert torch.allclose(lg["dZ"], Z.grad, rtol=1e-9, atol=1e-11)


def test_train_loop_fixture_losses():
    z = load("train_loop.npz")
    G = int(z["num_graphs"])
    data = []
    for i in range(G):
        csr = rs.csr_from_networkx(graph_from_edges(z[f"g{i}_edges"], int(z[f"g{i}_n"])))
        data.append((csr, rs.dense_adjacency(csr, 1000)))
    p = rs.GCNParams(*[torch.from_numpy(z[f"init_{k}"].copy()) for k in
                       ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")])
    st = rs.AdamState()
    hist = []
    for _ in range(len(z["loss_history"])):
        hist.append(sum(rs.train_step_closed_form(csr, X, p, st) for csr, X in data))
    np.testing.assert_allclose(hist, z["loss_history"], rtol=1e-4)
    for t, name in zip(p.tensors(), ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")):
        np.testing.assert_allclose(t.numpy(), z[f"final_{name}"], rtol=2e-4, atol=2e-6)


# ------------------------------------------------------------------ post-processing
@pytest.mark.parametrize("tag", ["s", "m", "l", "w"])
def test_postproc_fixture(tag):
    z = load("postproc.npz")
    n = int(z[f"{tag}_n"])
    g = graph_from_edges(z[f"{tag}_edges"], n, z[f"{tag}_w"])
    csr = rs.csr_from_networkx(g)
    wts = csr.weights.astype(np.int32)
    P = z[f"{tag}_P"]
    f32 = int(str(z["numpy_version"]).split(".")[0]) >= 2
    simple = pp.simple_assignment(P)
    assert simple.tolist() == z[f"{tag}_simple"].tolist()
    assert pp.cut_value(csr.rowptr, csr.colidx, simple, wts) == int(z[f"{tag}_simple_cut"])
    assert pp.py_cut_value(simple.tolist(), g) == int(z[f"{tag}_simple_cut"])
    np.random.seed(int(z[f"{tag}_assign_seed"]))
    U = np.random.rand(n - 3)
    assert pp.assign_partitions(P, U, compare_f32=f32).tolist() == z[f"{tag}_assign"].tolist()
    np.random.seed(int(z[f"{tag}_assign_seed"]))
    if pp.numpy_compares_in_f32() == f32:
        assert pp.py_assign_partitions(P) == z[f"{tag}_assign"].tolist()
    np.random.seed(int(z[f"{tag}_post_seed"]))
    U = np.random.rand(200 * (n - 3))
    np.testing.assert_array_equal(U[:16], z[f"{tag}_U_head"])
    labels, cut, _ = pp.sample_best_cut(csr.rowptr, csr.colidx, P, U, 200, wts, compare_f32=f32)
    assert cut == int(z[f"{tag}_post_cut"])
    assert labels.tolist() == z[f"{tag}_post_labels"].tolist()


def test_assign_partitions_compare_modes_differ_only_at_boundary():
    # r just below float32(p0) in float64 but equal after rounding to float32
    p0 = np.float32(0.3)
    P = np.zeros((4, 3), dtype=np.float32)
    P[3] = [p0, 0.5, 0.2]
    r = float(p0) - 1e-10
    assert np.float32(r) == p0
    assert pp.assign_partitions(P, np.asarray([r]), compare_f32=False)[3] == 0   # numpy 1.x: r < p0 in f64
    assert pp.assign_partitions(P, np.asarray([r]), compare_f32=True)[3] == 1    # NEP 50: f32(r) == p0


def test_greedy_matches_notebook_two_way():
    g = nx.random_regular_graph(d=4, n=36, seed=2)
    csr = rs.csr_from_networkx(g)
    start = np.zeros(36, dtype=np.int32)
    score, sol = pp.py_greedy_2way(start.tolist(), g, num_steps=36)
    labels, cut, moves = pp.greedy_node_move(csr.rowptr, csr.colidx, start, K=2, iters=36, n_frozen=0)
    assert cut == score and labels.tolist() == sol
    assert moves > 0


def test_greedy_properties_three_way():
    g = nx.random_regular_graph(d=7, n=200, seed=5)
    csr = rs.csr_from_networkx(g)
    rng = np.random.default_rng(0)
    start = rng.integers(0, 3, size=200).astype(np.int32)
    start[:3] = [0, 1, 2]
    c0 = pp.cut_value(csr.rowptr, csr.colidx, start)
    prev = c0
    for iters in (1, 5, 50, 200):
        labels, cut, moves = pp.greedy_node_move(csr.rowptr, csr.colidx, start, K=3, iters=iters)
        assert labels[:3].tolist() == [0, 1, 2]
        assert cut >= prev and moves <= iters
        assert cut == pp.cut_value(csr.rowptr, csr.colidx, labels)
        prev = cut
    # idempotent at a local optimum
    again, cut2, moves2 = pp.greedy_node_move(csr.rowptr, csr.colidx, labels, K=3, iters=200)
    if moves < 200:
        assert moves2 == 0 and cut2 == cut and again.tolist() == labels.tolist()


def test_extender_fixture_shapes():
    z = load("extender.npz")
    assert int(z["num_out"]) == 4
    for name in z["kept_names"]:
        X = z[f"{name}_X"]
        assert X.shape == (12, 16)
        assert int(z[f"{name}_nnz"]) == 2 * len(z[f"{name}_edges_out"])
        assert X[:, 12:].sum() == 0 and np.array_equal(X[:, :12], X[:, :12].T)


def test_preaggregated_layer1_is_the_same_function_as_the_reference_order():
    """The identity the B200 throughput path rests on (GCNEngine(preaggregate=True)), checked on the ORACLE in float64:
    GraphConv layer 1 is relu(A_hat (X W1) + b1) in the reference order (TrainingNeural.py:80, DGL multiplies by W first
    because in_feats > out_feats); with XA = A_hat X it is relu(XA W1 + b1), and its weight gradient
    X^T (A_hat dH1pre) equals XA^T dH1pre because A_hat is symmetric.  Also on the reference-generated step fixture."""
    g = nx.random_regular_graph(d=5, n=40, seed=7)
    nx.set_edge_attributes(g, 1, "weight")
    csr = rs.csr_from_networkx(g)
    X = rs.dense_adjacency(csr, 48, dtype=torch.float64)
    torch.manual_seed(0)
    p = rs.GCNParams(torch.randn(48, 20, dtype=torch.float64) * 0.2, torch.randn(20, dtype=torch.float64) * 0.1,
                     torch.randn(20, 3, dtype=torch.float64) * 0.3, torch.randn(3, dtype=torch.float64) * 0.1)
    fwd = rs.gcn_forward(csr, X, p)
    deg = torch.from_numpy(csr.degrees()).double()
    norm = deg.clamp(min=1).pow(-0.5).unsqueeze(1)
    XA = rs._aggregate(csr, X * norm) * norm                         # A_hat X
    H1_pre = torch.relu(XA @ p.W1 + p.b1)
    assert float((H1_pre - fwd["H1"]).abs().max()) < 1e-12
    dZ = torch.randn(40, 3, dtype=torch.float64)
    grads = rs.gcn_backward(csr, X, p, fwd, dZ)
    dT2 = rs._aggregate(csr, dZ * norm) * norm
    dH1pre = (dT2 @ p.W2.t()) * (fwd["pre1"] > 0).double()
    assert float((XA.t() @ dH1pre - grads["W1"]).abs().max()) < 1e-12
    assert float((dH1pre.sum(0) - grads["b1"]).abs().max()) < 1e-12


def test_bf16_storage_emulation_brackets_the_fp32_forward():
    """The storage-rounding emulation used by the GPU bf16 parity tests: both orders agree with the exact forward to
    bf16 rounding (a few 2^-9 relative steps through two layers) and with each other more closely than with fp32."""
    g = nx.random_regular_graph(d=7, n=64, seed=3)
    nx.set_edge_attributes(g, 1, "weight")
    csr = rs.csr_from_networkx(g)
    X = rs.dense_adjacency(csr, 64, dtype=torch.float64)
    torch.manual_seed(1)
    p = rs.GCNParams(torch.randn(64, 32, dtype=torch.float64) * 0.2, torch.randn(32, dtype=torch.float64) * 0.1,
                     torch.randn(32, 3, dtype=torch.float64) * 0.3, torch.randn(3, dtype=torch.float64) * 0.1)
    exact = rs.gcn_forward(csr, X, p)["Z"]
    pre = rs.gcn_forward_bf16_storage(csr, X, p, preaggregated=True)["Z"]
    std = rs.gcn_forward_bf16_storage(csr, X, p, preaggregated=False)["Z"]
    scale = float(exact.abs().max())
    assert 0 < float((pre - exact).abs().max()) < 2e-2 * scale
    assert 0 < float((std - exact).abs().max()) < 2e-2 * scale
    H = rs.gcn_forward_bf16_storage(csr, X, p)["H1"]
    assert torch.equal(H, H.to(torch.bfloat16).to(torch.float64))          # H1 is exactly representable in bf16


@pytest.mark.parametrize("burn,n", [(0, 1), (0, 312), (1, 311), (3, 313), (623, 2), (624, 5), (17, 20001), (1247, 625)])
def test_mt19937_restatement_is_numpys_generator(burn, n):
    """oracle/mt19937.py (the device generator's algorithm: three-sweep regeneration, carry across blocks) against np.random
    itself: same doubles, same state and position afterwards."""
    from oracle import mt19937 as mt
    np.random.seed(4321)
    if burn:
        np.random.randint(0, 2 ** 31, size=burn)
    st = np.random.get_state()
    want = np.random.rand(n)
    after = np.random.get_state()
    got, key, pos = mt.uniform(st[1], int(st[2]), n)
    assert np.array_equal(got, want)
    assert np.array_equal(key, after[1]) and pos == after[2]
    np.random.set_state(st)
