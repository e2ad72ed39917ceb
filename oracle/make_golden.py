"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the
UNMODIFIED reference (/root/reference, imported through oracle/ref_import.py with
the dgl/matplotlib shims) on small seeded inputs.  Run in the build container:

    python -m oracle.make_golden

The committed fixtures are what travels to the GPU box (the Python reference
cannot).  Inputs are stored next to outputs so that neither side of a parity test
needs the reference tree.
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import sys

import networkx as nx
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def edges_of(g: nx.Graph) -> np.ndarray:
    return np.asarray(sorted((min(u, v), max(u, v)) for u, v in g.edges()), dtype=np.int32)


def set_weights(net, rng: np.random.Generator, bias_scale: float = 0.05):
    """Deterministic (numpy-seeded) weights incl. NON-zero biases so that the bias
    path is exercised; shapes/keys follow TrainingNeural.py:72-77."""
    sd = net.state_dict()
    new = {}
    for k, v in sd.items():
        if k.endswith("weight"):
            fan_in, fan_out = v.shape
            a = np.sqrt(6.0 / (fan_in + fan_out))
            new[k] = torch.from_numpy(rng.uniform(-a, a, size=tuple(v.shape)).astype(np.float32))
        else:
            new[k] = torch.from_numpy(rng.normal(0, bias_scale, size=tuple(v.shape)).astype(np.float32))
    net.load_state_dict(new)
    return {k: v.numpy().copy() for k, v in new.items()}


def gen_gcn_step(ref):
    """One graph, fixed weights: forward, override+STE loss, autograd grads, one and
    three Adam steps -- TrainingNeural.py:371-388 executed by the reference itself."""
    T = ref.training
    rng = np.random.default_rng(7)
    out = {}
    for tag, n, d, seed, hidden in (("a", 60, 7, 11, 16), ("b", 90, 6, 12, 24)):
        g = ref.creator.generate_graph(n=n, d=d, graph_type="reg", random_seed=seed)
        with quiet():
            ds = ref.extender.process_graphs_from_folder({0: g}, {0: [5, 9, 17]}, max_nodes=1000)
        dgl_g, X, nx_g, terms = ds[0]
        cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=hidden, learning_rate=1e-3)
        torch.manual_seed(0)
        net, embed, opt = T.setup_model_and_optimizer(cfg)
        w = set_weights(net, rng)
        net.train()
        P = net(dgl_g, X)
        P.retain_grad()
        s = T.apply_max_to_one_hot(T.override_fixed_nodes(P))
        loss = T.compute_loss(s, X, cfg.A, cfg.C, cfg.penalty)
        opt.zero_grad()
        loss.backward()
        out[f"{tag}_edges"] = edges_of(nx_g)
        out[f"{tag}_n"] = np.int32(n)
        for k, v in w.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_P"] = P.detach().numpy().copy()
        out[f"{tag}_loss"] = np.float64(loss.item())
        out[f"{tag}_dP"] = P.grad.numpy().copy()
        for k, prm in net.named_parameters():
            out[f"{tag}_grad_{k}"] = prm.grad.numpy().copy()
        opt.step()
        for k, prm in net.named_parameters():
            out[f"{tag}_step1_{k}"] = prm.detach().numpy().copy()
        # two more reference steps on the same graph (sequential Adam semantics)
        losses = [loss.item()]
        for _ in range(2):
            losses.append(T.train_single_epoch(ds, net, opt, embed, cfg))
        out[f"{tag}_losses3"] = np.asarray(losses, dtype=np.float64)
        for k, prm in net.named_parameters():
            out[f"{tag}_step3_{k}"] = prm.detach().numpy().copy()
        with torch.no_grad():
            ev = T.evaluate_model(net, ds, cfg)
        out[f"{tag}_eval_total"] = np.float64(ev["total_loss"])
    np.savez_compressed(os.path.join(OUT, "gcn_step.npz"), **out)


def gen_train_loop(ref):
    """train_model on 3 graphs x 6 epochs with numpy-seeded weights: loss history,
    early-stop bookkeeping and return tuple -- TrainingNeural.py:392-484."""
    T = ref.training
    rng = np.random.default_rng(21)
    random.seed(3)
    graphs, terms = {}, {}
    for i, (n, d) in enumerate(((40, 6), (50, 7), (44, 8))):
        graphs[i] = ref.creator.generate_graph(n=n, d=d, graph_type="reg", random_seed=100 + i)
        terms[i] = ref.creator.generate_unique_terminals(n, 3)
    out = {f"g{i}_edges_in": edges_of(g) for i, g in graphs.items()}
    out.update({f"g{i}_terminals_in": np.asarray(t, dtype=np.int32) for i, t in terms.items()})
    with quiet():
        ds = ref.extender.process_graphs_from_folder(graphs, terms, max_nodes=1000)
    out["num_graphs"] = np.int32(len(ds))
    for i, item in ds.items():
        out[f"g{i}_edges"] = edges_of(item[2])
        out[f"g{i}_n"] = np.int32(item[2].number_of_nodes())
    cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=12, learning_rate=1e-3,
                           number_epochs=6, patience=20, save_directory=None)
    # seed the weights by patching setup (the loop itself is the reference's)
    real_setup = T.setup_model_and_optimizer
    captured = {}

    def patched(config):
        net, embed, opt = real_setup(config)
        captured["w"] = set_weights(net, rng)
        return net, embed, opt

    T.setup_model_and_optimizer = patched
    try:
        with quiet():
            net, best_loss, epoch, inputs, hist = T.train_model(ds, cfg)
    finally:
        T.setup_model_and_optimizer = real_setup
    for k, v in captured["w"].items():
        out[f"init_{k}"] = v
    out["loss_history"] = np.asarray(hist, dtype=np.float64)
    out["best_loss"] = np.float64(best_loss)
    out["final_epoch"] = np.int32(epoch)
    for k, prm in net.named_parameters():
        out[f"final_{k}"] = prm.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "train_loop.npz"), **out)


def gen_postproc(ref):
    """simple_partition_assignment / calculate_cut_value / assign_partitions /
    post_processing_optimization -- TestingNeuralNetwork.py:18-122, np.random seeded."""
    Te = ref.testing
    out = {"numpy_version": np.asarray(np.__version__)}
    rng = np.random.default_rng(5)
    cases = (("s", 50, 6, 1), ("m", 120, 7, 2), ("l", 300, 8, 3), ("w", 64, 6, 4))
    for tag, n, d, seed in cases:
        g = ref.creator.generate_graph(n=n, d=d, graph_type="reg", random_seed=seed)
        if tag == "w":  # non-unit integer weights
            for u, v in g.edges():
                g[u][v]["weight"] = int(rng.integers(1, 5))
        logits = rng.normal(0, 1.2, size=(n, 3)).astype(np.float32)
        P = torch.softmax(torch.from_numpy(logits), dim=1)
        simple = Te.simple_partition_assignment(P)
        simple_cut = Te.calculate_cut_value(simple, g)
        np.random.seed(1234 + seed)
        one = Te.assign_partitions(P.numpy())
        np.random.seed(99 + seed)
        best, score = Te.post_processing_optimization(P, g, iterations=200)
        np.random.seed(99 + seed)
        U = np.random.rand(200 * (n - 3))       # same stream as 200*(n-3) scalar draws
        after = np.random.rand()
        np.random.seed(99 + seed)
        Te.post_processing_optimization(P, g, iterations=200)
        assert after == np.random.rand(), "vector draw must leave the same RNG state as scalar draws"
        e = edges_of(g)
        out[f"{tag}_edges"] = e
        out[f"{tag}_w"] = np.asarray([g[int(u)][int(v)]["weight"] for u, v in e], dtype=np.int32)
        out[f"{tag}_n"] = np.int32(n)
        out[f"{tag}_P"] = P.numpy().copy()
        out[f"{tag}_simple"] = np.asarray(simple, dtype=np.int32)
        out[f"{tag}_simple_cut"] = np.int64(simple_cut)
        out[f"{tag}_assign_seed"] = np.int64(1234 + seed)
        out[f"{tag}_assign"] = np.asarray(one, dtype=np.int32)
        out[f"{tag}_post_seed"] = np.int64(99 + seed)
        out[f"{tag}_post_labels"] = np.asarray(best, dtype=np.int32)
        out[f"{tag}_post_cut"] = np.int64(score)
        out[f"{tag}_U_head"] = U[:16].copy()
    np.savez_compressed(os.path.join(OUT, "postproc.npz"), **out)


def gen_extender(ref):
    """process_graphs_from_folder terminal normalisation (graphExtender.py:68-122):
    the four swap cases, the skip case, mutation of caller's terminal lists."""
    out = {}
    cases = {"none": [7, 4, 9], "has2": [9, 2, 5], "has1": [8, 1, 6], "has0": [0, 10, 3],
             "skip01": [0, 1, 7], "skip_all": [2, 0, 1]}
    graphs, terms = {}, {}
    for i, (name, t) in enumerate(cases.items()):
        graphs[name] = ref.creator.generate_graph(n=12, d=3, graph_type="reg", random_seed=40 + i)
        terms[name] = list(t)
        out[f"{name}_edges_in"] = edges_of(graphs[name])
        out[f"{name}_terminals_in"] = np.asarray(t, dtype=np.int32)
    with quiet():
        ds = ref.extender.process_graphs_from_folder(graphs, terms, max_nodes=16)
    out["num_out"] = np.int32(len(ds))
    kept = [name for name in cases if not name.startswith("skip")]
    for i, name in enumerate(kept):
        dgl_g, X, nx_g, t = ds[i]
        out[f"{name}_edges_out"] = edges_of(nx_g)
        out[f"{name}_X"] = X.numpy().copy()
        out[f"{name}_nnz"] = np.int32(dgl_g.number_of_edges())
        out[f"{name}_terminals_after"] = np.asarray(terms[name], dtype=np.int32)
    out["kept_names"] = np.asarray(kept)
    np.savez_compressed(os.path.join(OUT, "extender.npz"), **out)


def gen_testing(ref):
    """test_multiple_graphs end to end with a numpy-seeded model and np.random.seed(0)
    (TestingNeuralNetwork.py:124-295)."""
    T, Te = ref.training, ref.testing
    rng = np.random.default_rng(33)
    random.seed(8)
    graphs, terms = {}, {}
    sizes = [30, 50]
    for size in sizes:
        for i in range(2):
            name = f"test_n{size}_{i}"
            graphs[name] = ref.creator.generate_graph(n=size, d=6 + i, graph_type="reg",
                                                      random_seed=size * 1000 + i)
            terms[name] = ref.creator.generate_unique_terminals(size, 3)
    out = {f"{k}_edges_in": edges_of(g) for k, g in graphs.items()}
    out.update({f"{k}_terminals_in": np.asarray(t, dtype=np.int32) for k, t in terms.items()})
    out["names"] = np.asarray(list(graphs.keys()))
    with quiet():
        ds = ref.extender.process_graphs_from_folder(graphs, terms, max_nodes=1000)
    cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=10)
    net, _, _ = T.setup_model_and_optimizer(cfg)
    w = set_weights(net, rng)
    net.eval()
    for k, v in w.items():
        out[f"model_{k}"] = v
    np.random.seed(0)
    with quiet():
        results, by_size = Te.test_multiple_graphs(net, ds, sizes, post_processing_iterations=50,
                                                   verbose=False)
    out["num_results"] = np.int32(len(results))
    for i, r in enumerate(results):
        out[f"r{i}_edges"] = edges_of(ds[i][2])
        out[f"r{i}_n"] = np.int32(r["nodes"])
        out[f"r{i}_size"] = np.int32(r["graph_size"])
        out[f"r{i}_simple_cut"] = np.int64(r["simple_cut"])
        out[f"r{i}_post_cut"] = np.int64(r["post_cut"])
        out[f"r{i}_simple_assignment"] = np.asarray(r["simple_assignment"], dtype=np.int32)
        out[f"r{i}_post_assignment"] = np.asarray(r["post_assignment"], dtype=np.int32)
        out[f"r{i}_P"] = r["node_probabilities"].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "testing.npz"), **out)


def w1_digest(W: np.ndarray) -> dict:
    """Compact but covering summary of a [1000, 500] matrix (2 MB in full): every 16th row verbatim, all row sums, all
    column sums and the Frobenius norm -- each entry enters two of the linear functionals."""
    W = np.asarray(W, dtype=np.float64)
    return {"rows16": W[::16].astype(np.float32), "rowsum": W.sum(1), "colsum": W.sum(0),
            "fro": np.float64(np.sqrt((W * W).sum()))}


def gen_baseline_shapes(ref):
    """BASELINE.json config 1 at its real shapes, executed by the reference itself: n = 500, 1000 features, hidden 500.
    (a) one graph: forward, override + STE loss, autograd gradients (TrainingNeural.py:371-385);
    (b) the 20-graph pipeline of complete_training_pipeline.ipynb cell 15 for two epochs through train_model
        (:392-484): loss history + final weights;
    (c) config 2: test_multiple_graphs with 200 post-processing iterations on one graph per size 50 .. 500
        (TestingNeuralNetwork.py:188-295) with the model (b) trained.
    Weights are numpy-seeded (set_weights with default_rng(seed)); the tests regenerate them from the same seeds, so
    only outputs and the (swapped) edge lists are stored."""
    T, Te = ref.training, ref.testing
    out = {}
    # ---- (a) single step
    rng = np.random.default_rng(101)
    out["a_weight_seed"] = np.int64(101)
    g = ref.creator.generate_graph(n=500, d=7, graph_type="reg", random_seed=1000)
    with quiet():
        ds = ref.extender.process_graphs_from_folder({0: g}, {0: [17, 250, 433]}, max_nodes=1000)
    dgl_g, X, nx_g, _ = ds[0]
    cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500, learning_rate=1e-3)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    set_weights(net, rng)
    net.train()
    P = net(dgl_g, X)
    s = T.apply_max_to_one_hot(T.override_fixed_nodes(P))
    loss = T.compute_loss(s, X, cfg.A, cfg.C, cfg.penalty)
    opt.zero_grad()
    loss.backward()
    out["a_edges"] = edges_of(nx_g).astype(np.int16)
    out["a_P"] = P.detach().numpy().copy()
    out["a_loss"] = np.float64(loss.item())
    for k, prm in net.named_parameters():
        if k == "conv1.weight":
            for kk, v in w1_digest(prm.grad.numpy()).items():
                out[f"a_grad_{k}_{kk}"] = v
        else:
            out[f"a_grad_{k}"] = prm.grad.numpy().copy()
    # ---- (b) 20 graphs x 2 epochs
    random.seed(0)
    graphs, terms = {}, {}
    for i in range(20):
        graphs[i] = ref.creator.generate_graph(n=500, d=random.randint(6, 8), graph_type="reg", random_seed=1000 + i)
        terms[i] = ref.creator.generate_unique_terminals(500, 3)
    with quiet():
        ds = ref.extender.process_graphs_from_folder(graphs, terms, max_nodes=1000)
    out["b_num_graphs"] = np.int32(len(ds))
    for i, item in ds.items():
        out[f"b_g{i}_edges"] = edges_of(item[2]).astype(np.int16)
    cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500, learning_rate=1e-3, number_epochs=2,
                           patience=20, save_directory=None)
    rng_b = np.random.default_rng(202)
    out["b_weight_seed"] = np.int64(202)
    real_setup = T.setup_model_and_optimizer

    def patched(config):
        net, embed, opt = real_setup(config)
        set_weights(net, rng_b)
        return net, embed, opt

    T.setup_model_and_optimizer = patched
    try:
        with quiet():
            net, best_loss, epoch, inputs, hist = T.train_model(ds, cfg)
    finally:
        T.setup_model_and_optimizer = real_setup
    out["b_loss_history"] = np.asarray(hist, dtype=np.float64)
    out["b_best_loss"] = np.float64(best_loss)
    for k, prm in net.named_parameters():
        if k == "conv1.weight":
            for kk, v in w1_digest(prm.detach().numpy()).items():
                out[f"b_final_{k}_{kk}"] = v
        else:
            out[f"b_final_{k}"] = prm.detach().numpy().copy()
    with torch.no_grad():
        ev = T.evaluate_model(net, ds, cfg)
    out["b_eval_total"] = np.float64(ev["total_loss"])
    # ---- (c) config 2 on the trained model: one graph per size, 200 iterations
    sizes = [50, 100, 200, 300, 500]
    random.seed(5)
    tg, tt = {}, {}
    for size in sizes:
        name = f"test_n{size}_0"
        tg[name] = ref.creator.generate_graph(n=size, d=random.randint(6, 8), graph_type="reg", random_seed=size * 1000)
        tt[name] = ref.creator.generate_unique_terminals(size, 3)
    with quiet():
        tds = ref.extender.process_graphs_from_folder(tg, tt, max_nodes=1000)
    net.eval()
    np.random.seed(0)
    with quiet():
        results, _ = Te.test_multiple_graphs(net, tds, sizes, post_processing_iterations=200, verbose=False)
    out["c_num_results"] = np.int32(len(results))
    for i, r in enumerate(results):
        out[f"c_r{i}_edges"] = edges_of(tds[i][2]).astype(np.int16)
        out[f"c_r{i}_n"] = np.int32(r["nodes"])
        out[f"c_r{i}_simple_cut"] = np.int64(r["simple_cut"])
        out[f"c_r{i}_post_cut"] = np.int64(r["post_cut"])
        out[f"c_r{i}_simple_assignment"] = np.asarray(r["simple_assignment"], dtype=np.int8)
        out[f"c_r{i}_post_assignment"] = np.asarray(r["post_assignment"], dtype=np.int8)
        out[f"c_r{i}_P"] = r["node_probabilities"].astype(np.float32)
    out["c_rng_after"] = np.float64(np.random.rand())
    np.savez_compressed(os.path.join(OUT, "baseline_shapes.npz"), **out)


def main():
    if not ref_import.available():
        raise SystemExit("reference tree not available; fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref = ref_import.load()
    gen_gcn_step(ref)
    gen_train_loop(ref)
    gen_postproc(ref)
    gen_extender(ref)
    gen_testing(ref)
    gen_baseline_shapes(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
