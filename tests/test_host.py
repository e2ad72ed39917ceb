"""CPU tests (`-m "not gpu"`): host logic of the drop-in, the C-ABI library's exported symbols,
and the no-CPU-fallback guarantee.  No compute call is made here."""
import os
import pickle
import re
import subprocess
import sys
import types

import networkx as nx
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from gmc_b200 import _lib, dist as gdist, synth  # noqa: E402
from gmc_b200.graph import CSRGraph  # noqa: E402
from oracle import ref_step as rs  # noqa: E402


def graph_from_edges(edges, n, w=None):
    g = nx.Graph()
    g.add_nodes_from(range(int(n)))
    for i, (u, v) in enumerate(edges):
        g.add_edge(int(u), int(v), weight=int(w[i]) if w is not None else 1, capacity=1)
    return g


# ------------------------------------------------------------------ C ABI
def header_symbols():
    text = open(os.path.join(ROOT, "include", "gcnmaxcut.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gmc_[a-z0-9_]+)\s*\(", text)))


def test_header_binding_and_library_agree():
    hdr = header_symbols()
    assert len(hdr) >= 25
    assert hdr == _lib.exported_symbols(), "ctypes SIGNATURES must list exactly the header's entry points"
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python gcn-max-cut_b200/build.py"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (gmc_[a-z0-9_]+)", out))
    assert set(hdr) <= exported, f"missing from the .so: {sorted(set(hdr) - exported)}"


def test_library_loads_and_reports_abi():
    lib = _lib.lib()
    assert lib.gmc_abi_version() == 1
    assert lib.gmc_gemm_workspace_bytes(0, 128, 128, 128, 0) == 0


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    from Training import TrainingNeural as T
    from Testing import TestingNeuralNetwork as Te
    with pytest.raises(_lib.GmcError):
        T.setup_model_and_optimizer(T.TrainingConfig(n_nodes=16))
    with pytest.raises(_lib.GmcError):
        T.train_single_epoch({}, torch.nn.Linear(1, 1), None, None, T.TrainingConfig(n_nodes=16))
    g = nx.random_regular_graph(3, 10, seed=0)
    with pytest.raises(_lib.GmcError):
        Te.post_processing_optimization(np.full((10, 3), 1 / 3, dtype=np.float32), g, 5)
    with pytest.raises(_lib.GmcError):
        Te.simple_partition_assignment(torch.full((10, 3), 1 / 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gcn-max-cut_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.replace("oracle/postproc.c", "").replace("oracle/", "") or \
                    "import oracle" not in text and "from oracle" not in text, f
                assert "from oracle" not in text and "import oracle" not in text, f


# ------------------------------------------------------------------ both import spellings
def test_both_import_spellings_resolve():
    import importlib
    a = importlib.import_module("Training.TrainingNeural")
    b = importlib.import_module("python.Training.TrainingNeural")
    for m in (a, b):
        cfg = m.TrainingConfig()
        assert (cfg.n_nodes, cfg.dim_embedding, cfg.hidden_dim, cfg.number_classes) == (1000, 1000, 500, 3)
        assert (cfg.learning_rate, cfg.number_epochs, cfg.tolerance, cfg.patience) == (0.001, 1000, 1e-4, 20)
        assert (cfg.A, cfg.C, cfg.penalty, cfg.save_directory, cfg.save_frequency) == (0.0, 1.0, 1000.0, None, 100)
        for name in ("train_from_pickle", "train_model", "train_single_epoch", "setup_model_and_optimizer",
                     "evaluate_model", "load_neural_model", "save_neural_model", "GCNSoftmax", "train_multi_class",
                     "get_gnn", "hyperParameters", "train1", "train_2wayNeural", "FIndAC", "GetOptimalNetValue",
                     "calculateAllCut", "LoadNeuralModel", "compute_loss", "calculate_HC_vectorized",
                     "override_fixed_nodes", "apply_max_to_one_hot", "terminal_independence_penalty"):
            assert hasattr(m, name), name
    cfg = a.TrainingConfig(n_nodes=64, dim_embedding=None)
    assert cfg.dim_embedding == 64 and cfg.hidden_dim == 32
    for mod in ("Testing.TestingNeuralNetwork", "python.Testing.TestingNeuralNetwork",
                "DataGenerator.graphExtender", "python.DataGenerator.graphExtender",
                "DataGenerator.GraphCreator", "commons", "python.commons"):
        importlib.import_module(mod)


def test_config_pickles_under_reference_module_paths():
    from Training import TrainingNeural as A
    from python.Training import TrainingNeural as B
    for m in (A, B):
        cfg = m.TrainingConfig(n_nodes=77, learning_rate=0.5)
        blob = pickle.dumps(cfg)
        assert m.__name__.encode() in blob
        back = pickle.loads(blob)
        assert back == cfg and back.hidden_dim == 38


# ------------------------------------------------------------------ loss helper semantics (CPU torch)
def test_loss_helpers_match_reference_semantics():
    from Training import TrainingNeural as T
    torch.manual_seed(0)
    g = nx.random_regular_graph(5, 20, seed=1)
    csr = rs.csr_from_networkx(g)
    A = rs.dense_adjacency(csr, 1000)
    Z = torch.randn(20, 3, requires_grad=True)
    P = torch.softmax(Z, dim=1)
    s = T.apply_max_to_one_hot(T.override_fixed_nodes(P))
    loss = T.compute_loss(s, A, 0.0, 1.0, 1000.0)
    loss.backward()
    lg = rs.ste_loss_and_grads(csr, P.detach())
    assert abs(float(loss) - float(lg["loss"])) < 1e-4
    assert torch.allclose(Z.grad, lg["dZ"], atol=1e-6)
    assert torch.equal(T.max_to_one_hot(torch.tensor([0.2, 0.5, 0.3])).detach(), torch.tensor([0.0, 1.0, 0.0]))
    with pytest.raises(ValueError):
        T.extend_matrix_torch(torch.zeros(5, 5), 4)
    assert T.find_ac_parameters(g) == (6, 2.5)
    assert len(T.generate_terminal_permutations({"a": 0, "b": 1, "c": 2})) == 6
    assert T.hyperParameters(n=80)[-2:] == (80, 40)


# ------------------------------------------------------------------ training-loop host logic
class _FakeNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(2))


def _loop(losses, **cfg_kw):
    from Training import TrainingNeural as T
    cfg = T.TrainingConfig(n_nodes=8, **cfg_kw)
    it = iter(losses)
    saved = []
    net = _FakeNet()
    embed = torch.nn.Embedding(8, 8)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    out = T._run_training_loop(cfg, lambda: next(it), net, opt, embed, saver=lambda obj, path: saved.append((path, obj)))
    return out, saved


def test_early_stopping_and_history(capsys):
    # loss rises / stalls for `patience` consecutive epochs -> stop (reference :430-437)
    (net, best, epoch, inputs, hist), saved = _loop([-10.0, -12.0, -11.0, -11.0, -11.00001, -5.0, -20.0],
                                                    number_epochs=50, patience=3)
    assert hist == [-10.0, -12.0, -11.0, -11.0, -11.00001]
    assert epoch == 4 and best == -12.0
    assert "Early stopping at epoch 4" in capsys.readouterr().out
    # counter resets on improvement
    (_, best, epoch, _, hist), _ = _loop([-1.0, -0.5, -2.0, -1.5, -3.0, -2.5, -4.0], number_epochs=7, patience=2)
    assert epoch == 6 and len(hist) == 7 and best == -4.0


def test_checkpoint_names_and_keys(capsys):
    (net, best, epoch, inputs, hist), saved = _loop([-1.0, -2.0, -3.0, -4.0, -5.0], number_epochs=5, patience=20,
                                                    save_directory="m.pth", save_frequency=2)
    paths = [p for p, _ in saved]
    assert paths == ["./epoch_0_loss_-1.0000_m.pth", "./epoch_2_loss_-3.0000_m.pth", "./epoch_4_loss_-5.0000_m.pth",
                     "./final_m.pth"]
    for _, ckpt in saved:
        assert set(ckpt) == {"epoch", "model", "optimizer", "loss_history", "inputs", "config"}
    out = capsys.readouterr().out
    assert "Epoch: 0, Cumulative Loss: -1.000000" in out and "Best loss: -5.000000" in out
    assert "Final model saved to ./final_m.pth" in out
    assert inputs.shape == (8, 8)


# ------------------------------------------------------------------ graph handles / extender / creator
def test_csrgraph_matches_oracle_csr_and_dgl_surface():
    g = nx.random_regular_graph(4, 30, seed=3)
    for u, v in g.edges():
        g[u][v]["weight"] = 1 + (u + v) % 3
    h = CSRGraph.from_networkx(g)
    o = rs.csr_from_networkx(g)
    assert np.array_equal(h.rowptr, o.rowptr) and np.array_equal(h.colidx, o.colidx)
    assert np.array_equal(h.weights, o.weights)
    assert h.number_of_nodes() == 30 and h.number_of_edges() == 2 * g.number_of_edges()
    assert h.to("cuda") is h
    back = pickle.loads(pickle.dumps(h))
    assert np.array_equal(back.colidx, h.colidx) and back._batch is None
    # non-contiguous labels are relabelled in sorted order
    g2 = nx.relabel_nodes(g, {i: 10 * i + 5 for i in g.nodes})
    h2 = CSRGraph.from_networkx(g2)
    assert np.array_equal(h2.colidx, h.colidx)


def test_graph_extender_matches_reference_fixture(capsys):
    from DataGenerator import graphExtender as E
    z = np.load(os.path.join(GOLDEN, "extender.npz"))
    names = ["none", "has2", "has1", "has0", "skip01", "skip_all"]
    graphs = {k: graph_from_edges(z[f"{k}_edges_in"], 12) for k in names}
    terms = {k: z[f"{k}_terminals_in"].tolist() for k in names}
    ds = E.process_graphs_from_folder(graphs, terms, max_nodes=16)
    out = capsys.readouterr().out
    assert "Skipped items: 2" in out and "Terminal swapped 0" in out and "Graph finished: 4" in out
    assert len(ds) == int(z["num_out"]) and list(ds.keys()) == [0, 1, 2, 3]
    for i, name in enumerate(z["kept_names"]):
        handle, X, nx_g, t = ds[i]
        got = np.asarray(sorted((min(u, v), max(u, v)) for u, v in nx_g.edges()), dtype=np.int32)
        assert np.array_equal(got, z[f"{name}_edges_out"]), name
        assert np.array_equal(X.numpy(), z[f"{name}_X"]) and X.dtype == torch.float32
        assert handle.number_of_edges() == int(z[f"{name}_nnz"]) and t == [0, 1, 2]
        assert terms[str(name)] == z[f"{name}_terminals_after"].tolist()       # caller's list mutated alike
        assert nx_g is graphs[str(name)]                                        # mutated in place
    with pytest.raises(ValueError):
        E.extend_matrix_torch_2(torch.zeros(5, 5), 4)
    # max_nodes < n is swallowed into a print + partial result, like the reference (:126-129)
    ds2 = E.process_graphs_from_folder({"a": graph_from_edges(z["none_edges_in"], 12)}, {"a": [7, 4, 9]}, max_nodes=8)
    assert ds2 == {} and "Exception occurred at graph 0" in capsys.readouterr().out


def test_extender_batch_pickles(tmp_path, capsys):
    from DataGenerator import graphExtender as E, GraphCreator as C
    import commons
    graphs = {i: C.generate_graph(n=10, d=3, random_seed=i) for i in range(5)}
    terms = {i: [5, 6, 7] for i in range(5)}
    prefix = str(tmp_path / "proc")
    rest = E.process_graphs_from_folder(graphs, terms, max_nodes=12, save_batch_size=2, output_filename_prefix=prefix)
    assert sorted(os.listdir(tmp_path)) == ["proc_2.pkl", "proc_4.pkl"] and list(rest.keys()) == [4]
    first = commons.open_file(prefix + "_2.pkl")
    assert list(first.keys()) == [0, 1] and isinstance(first[0][0], CSRGraph) and first[0][1].shape == (10, 12)


def test_graph_creator_matches_networkx_and_validates():
    from DataGenerator import GraphCreator as C
    import random
    g = C.generate_graph(n=50, d=7, graph_type="reg", random_seed=1234)
    want = nx.random_regular_graph(d=7, n=50, seed=1234)
    assert sorted(map(tuple, map(sorted, g.edges()))) == sorted(map(tuple, map(sorted, want.edges())))
    assert list(g.nodes()) == list(range(50))
    assert all(d == {"weight": 1, "capacity": 1} for _, _, d in g.edges(data=True))
    for bad in (dict(n=5, d=None), dict(n=0, d=1), dict(n=5, d=5), dict(n=5, p=1.5, graph_type="prob")):
        with pytest.raises(ValueError):
            C.generate_graph(**bad)
    with pytest.raises(NotImplementedError):
        C.generate_graph(n=5, d=2, graph_type="nope")
    random.seed(0)
    t = C.generate_unique_terminals(500, 3)
    random.seed(0)
    assert t == random.sample(range(500), 3)
    random.seed(1)
    gs, ts = C.generate_graph_dataset(4, 20, 30, 3, 5)
    assert len(gs) == 4 and all(len(set(v)) == 3 for v in ts.values())


def test_open_file_adopts_legacy_dgl_handles(tmp_path):
    import commons
    # fabricate a pickle that references a class of a module which is not installed
    fake = types.ModuleType("dgl")
    sub = types.ModuleType("dgl.heterograph")

    class DGLGraph:
        def __init__(self):
            self.payload = [1, 2, 3]

    DGLGraph.__module__ = "dgl.heterograph"
    DGLGraph.__qualname__ = "DGLGraph"
    sub.DGLGraph = DGLGraph
    fake.heterograph = sub
    sys.modules["dgl"], sys.modules["dgl.heterograph"] = fake, sub
    try:
        g = nx.random_regular_graph(3, 8, seed=0)
        nx.set_edge_attributes(g, 1, "weight")
        path = str(tmp_path / "legacy.pkl")
        commons.save_object({0: [DGLGraph(), torch.zeros(8, 10), g, [0, 1, 2]]}, path)
    finally:
        del sys.modules["dgl"], sys.modules["dgl.heterograph"]
    data = commons.open_file(path)
    handle = data[0][0]
    assert isinstance(handle, CSRGraph) and handle.number_of_edges() == 24


def test_commons_adjacency_helpers():
    import commons
    g = nx.random_regular_graph(3, 10, seed=2)
    for u, v in g.edges():
        g[u][v]["weight"] = 2
    dense = commons.qubo_dict_to_torch(g, commons.gen_adj_matrix(g), torch_dtype=torch.float32)
    fast = commons.adjacency_tensor(g, 14)
    assert torch.equal(fast[:, :10], dense) and fast[:, 10:].abs().sum() == 0
    assert torch.equal(dense, torch.from_numpy(nx.to_numpy_array(g, nodelist=range(10), weight="weight")).float())
    with pytest.raises(ValueError):
        commons.adjacency_tensor(g, 9)


# ------------------------------------------------------------------ testing-module host helpers
def test_assign_partitions_and_cut_value_match_reference_fixture():
    from Testing import TestingNeuralNetwork as Te
    z = np.load(os.path.join(GOLDEN, "postproc.npz"))
    same_mode = (int(str(z["numpy_version"]).split(".")[0]) >= 2) == Te._numpy_compares_in_f32()
    for tag in ("s", "w"):
        n = int(z[f"{tag}_n"])
        g = graph_from_edges(z[f"{tag}_edges"], n, z[f"{tag}_w"])
        assert Te.calculate_cut_value(z[f"{tag}_simple"].tolist(), g) == int(z[f"{tag}_simple_cut"])
        if same_mode:
            np.random.seed(int(z[f"{tag}_assign_seed"]))
            assert Te.assign_partitions(z[f"{tag}_P"]) == z[f"{tag}_assign"].tolist()
    assert Te.calculate_cut_value([0, 1], g) >= 0           # short assignment: out-of-range nodes ignored


def test_size_bucketing_rules():
    from Testing import TestingNeuralNetwork as Te
    g = nx.path_graph(52)
    assert Te._size_bucket("test_n50_3", g, [50, 100]) == ("test_n50_3", 50)
    assert Te._size_bucket("weird", g, [50, 100]) == ("weird", 52)
    assert Te._size_bucket(7, g, [50, 100]) == ("graph_7", 50)
    assert Te._size_bucket(7, nx.path_graph(60), [50, 100]) == ("graph_7", 60)


# ------------------------------------------------------------------ synthetic generator / sharding
def test_regular_batch_arrays_are_simple_regular_block_diagonal():
    B, n, d = 12, 60, 7
    rowptr, colidx, gp = synth.regular_batch_arrays(B, n, d, seed=5)
    assert np.array_equal(np.diff(rowptr), np.full(B * n, d)) and gp.tolist() == [i * n for i in range(B + 1)]
    rows = np.repeat(np.arange(B * n), d)
    assert (rows != colidx).all() and (rows // n == colidx // n).all()
    keys = rows.astype(np.int64) * (B * n) + colidx
    assert len(np.unique(keys)) == len(keys)
    rev = colidx.astype(np.int64) * (B * n) + rows
    assert np.array_equal(np.sort(keys), np.sort(rev))      # symmetric
    a = synth.regular_batch_arrays(3, 20, 3, seed=1)[1]
    b = synth.regular_batch_arrays(3, 20, 3, seed=1)[1]
    assert np.array_equal(a, b)
    mixed = synth.regular_batch_arrays(6, 40, [6, 7, 8, 6, 7, 8], seed=2)
    assert np.diff(mixed[0]).reshape(6, 40)[:, 0].tolist() == [6, 7, 8, 6, 7, 8]
    with pytest.raises(ValueError):
        synth.regular_graph_edges(1, 11, 3)


def test_shard_bounds_partition_everything():
    for total, world in ((32768, 8), (10, 3), (5, 8), (0, 2)):
        spans = [gdist.shard_bounds(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert gdist.flat_grad_layout([500000, 500, 1500, 3]) == ([0, 500000, 500500, 502000], 502004)


def _gloo_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "gcn-max-cut_b200"))
    from gmc_b200 import dist as gd
    gd.init_from_env(backend="gloo")
    # every rank owns a shard of "graphs" whose gradient contribution is its index; the all-reduced flat
    # buffer must equal the full sum irrespective of the sharding
    lo, hi = gd.shard_bounds(11, rank, world)
    flat = torch.zeros(8, dtype=torch.float32)
    for g in range(lo, hi):
        flat += torch.arange(8, dtype=torch.float32) * (g + 1)
    loss = torch.tensor([float(sum(range(lo, hi)))], dtype=torch.float64)
    gd.all_reduce_sum_(flat)
    gd.all_reduce_sum_(loss)
    t = torch.tensor([float(rank)])
    gd.all_reduce_max_(t)
    gd.barrier()
    q.put((rank, flat.tolist(), loss.item(), t.item()))
    dist.destroy_process_group()


def test_data_parallel_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = (torch.arange(8, dtype=torch.float32) * sum(range(1, 12))).tolist()
    for _, flat, loss, mx in res:
        assert flat == want and loss == float(sum(range(11))) and mx == 1.0


def _shard_worker(rank, world, port, q):
    """world-size-2 gloo: train_single_epoch's step preparation shards every global batch contiguously over the ranks."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from Training import TrainingNeural as T
    from gmc_b200 import synth
    rp, ci, gp = synth.regular_batch_arrays(11, 16, 3, seed=4)
    ds = synth.RegularGraphDataset(rp, ci, gp, 16, with_networkx=False)
    steps = T._prepare(ds, 4, "cpu", None, stream=True)          # global batches of 4, 4, 3 graphs
    out = []
    for st in steps:
        hb = st.host
        out.append(None if hb is None else (st.n_graphs, hb.rowptr.tolist(), hb.colidx.tolist(), hb.graph_ptr.tolist(), hb.regular))
    touched = sorted(k for k in ds.keys() if dict.__getitem__(ds, k) is not None)
    # the peer-memory gradient exchange is an NCCL-on-GPU, opt-in affair: under gloo every rank agrees on "no" without a
    # collective (an engine built here would use torch.distributed's all_reduce)
    from gmc_b200 import dist as gdist
    os.environ["GMC_PEER_ALLREDUCE"] = "1"
    assert gdist.make_peer_allreduce(1024, "cpu") is None
    q.put((rank, out, touched))
    dist.barrier()
    dist.destroy_process_group()


def test_train_steps_are_sharded_by_rank_world2_gloo():
    import torch.multiprocessing as mp
    from gmc_b200 import synth
    from gmc_b200.dist import shard_bounds
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30000 + os.getpid() % 1000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (out, touched)) for r, out, touched in (q.get(timeout=180) for _ in procs))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rp, ci, gp = synth.regular_batch_arrays(11, 16, 3, seed=4)
    chunks = [(0, 4), (4, 8), (8, 11)]
    seen = {0: [], 1: []}
    for step, (lo, hi) in enumerate(chunks):
        for rank in (0, 1):
            a, b = shard_bounds(hi - lo, rank, 2)
            got = res[rank][0][step]
            assert got[0] == b - a
            g0, g1 = lo + a, lo + b
            want_rp = (rp[g0 * 16: g1 * 16 + 1] - rp[g0 * 16]).tolist()
            want_ci = (ci[rp[g0 * 16]: rp[g1 * 16]] - g0 * 16).tolist()
            assert got[1] == want_rp and got[2] == want_ci and got[3] == [16 * i for i in range(b - a + 1)] and got[4]
            seen[rank] += list(range(g0, g1))
    # every graph belongs to exactly one rank, and a rank never dereferenced another rank's items
    assert sorted(seen[0] + seen[1]) == list(range(11))
    assert res[0][1] == sorted(seen[0]) and res[1][1] == sorted(seen[1])


# ------------------------------------------------------------------ bench.py argument rules / config fields
def _bench_args(argv, monkeypatch):
    import importlib
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    for k in ("GMC_BENCH_PRECISION", "GMC_BENCH_ACTIVATIONS", "GMC_BENCH_LAYER1"):
        monkeypatch.delenv(k, raising=False)
    return bench.parse_args()


def test_bench_defaults_name_the_throughput_configuration(monkeypatch):
    a = _bench_args([], monkeypatch)
    assert (a.gpus, a.steps, a.warmup) == (1, 20, 3) and a.warmup >= 3
    assert (a.precision, a.activations, a.layer1, a.workload) == ("bf16", "bf16", "preaggregated", "config3")
    assert (a.graphs_per_gpu, a.nodes, a.degree, a.features, a.hidden, a.classes) == (4096, 1000, 7, 1000, 500, 3)


@pytest.mark.parametrize("argv,want", [
    (["--precision", "tf32"], ("tf32", "fp32", "standard")),            # bf16 storage needs bf16 GEMM operands
    (["--activations", "fp32"], ("bf16", "fp32", "standard")),          # pre-aggregation is built on the bf16 chain
    (["--layer1", "standard"], ("bf16", "bf16", "standard")),
    (["--feature-source", "embedding"], ("bf16", "bf16", "standard")),  # trainable features: nothing to pre-aggregate
    (["--precision", "fp32", "--layer1", "preaggregated"], ("fp32", "fp32", "standard")),
    (["--workload", "config5"], ("tf32", "fp32", "standard")),         # config 5 trains embeddings
])
def test_bench_option_rules(argv, want, monkeypatch):
    a = _bench_args(argv, monkeypatch)
    assert (a.precision, a.activations, a.layer1) == want


def test_training_config_extension_fields_default_to_reference_behaviour():
    from Training import TrainingNeural as T
    cfg = T.TrainingConfig()
    assert (cfg.batch_graphs, cfg.gemm_precision, cfg.activations, cfg.preaggregate_features,
            cfg.adjacency_kernels, cfg.loss_mode, cfg.feature_source) == (1, "fp32", "fp32", False, False, "ste", "adjacency")
    blob = pickle.dumps(T.TrainingConfig(activations="bf16", gemm_precision="bf16", preaggregate_features=True))
    back = pickle.loads(blob)
    assert back.activations == "bf16" and back.preaggregate_features is True


def test_feature_verification_is_memoised_per_item_and_notices_writes():
    """check_adjacency_features (host tensors stay on the host): a verified (graph, tensor) pair is not verified again while
    the tensor is untouched -- a test harness revisits the same dataset items every pass -- and any write torch can see
    (version counter) or another tensor object triggers the full check again; a pickled handle carries no memo."""
    import pickle
    import networkx as nx
    import torch
    from gmc_b200 import graph as G
    g = nx.random_regular_graph(d=4, n=30, seed=1)
    nx.set_edge_attributes(g, 1, "weight")
    h = G.CSRGraph.from_networkx(g)
    X = torch.zeros(30, 40)
    for u, v in g.edges():
        X[u, v] = X[v, u] = 1.0
    assert G.check_adjacency_features(h, X) == 40
    memo = h._features_verified
    assert memo is not None and memo[0]() is X and memo[1] == X._version
    assert G.check_adjacency_features(h, X) == 40 and h._features_verified is memo      # served from the memo
    X[0, 39] = 5.0                                   # a write: version bumps, the entry is not an edge -> refused
    with pytest.raises(NotImplementedError):
        G.check_adjacency_features(h, X)
    X[0, 39] = 0.0
    assert G.check_adjacency_features(h, X) == 40    # verified afresh at the new version
    assert h._features_verified[1] == X._version
    Y = X.clone()
    Y[3, 3] = 1.0                                    # another tensor object: its own check
    with pytest.raises(NotImplementedError):
        G.check_adjacency_features(h, Y)
    h2 = pickle.loads(pickle.dumps(h))
    assert h2._features_verified is None and h2._nx_edges is None and np.array_equal(h2.colidx, h.colidx)
