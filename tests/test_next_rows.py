"""Tests of the SURVEY 8(f) "next" rows: batched testing harness, GPU randomized baseline, text loader."""
import os
import random

import networkx as nx
import numpy as np
import pytest
import torch

from oracle import postproc as pp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------ host-only
def test_randint_word_stream_matches_python_random():
    from RandomAlgorithm import RandomizedMaxCut as R
    for k in (2, 3, 5, 8, 200):
        random.seed(1000 + k)
        want = [random.randint(0, k - 1) for _ in range(3000)]
        after = random.random()
        random.seed(1000 + k)
        stream = R._WordStream()
        got = R._labels_from_stream(stream, k, 777).tolist() + R._labels_from_stream(stream, k, 3000 - 777).tolist()
        stream.commit()
        assert got == want and random.random() == after


def test_text_graph_loader_roundtrip_and_errors(tmp_path, capsys):
    from DataGenerator import textGraphLoader as L
    from DataGenerator.textGraphLoader import TextGraphLoader
    g = nx.random_regular_graph(3, 10, seed=1)
    for u, v in g.edges():
        g[u][v]["weight"] = 1 + (u + v) % 3
    L.write_graph_to_text(g, [4, 7, 2], str(tmp_path / "a.txt"))
    (tmp_path / "b.txt").write_text("[0, 1, 2]\n0 1\n\n1 2 2.5\nbad\nx y\n2 0 1\n")
    (tmp_path / "c.txt").write_text("0 1\n")
    (tmp_path / "empty.txt").write_text("")
    graph, terms = TextGraphLoader.load_graph_from_text(str(tmp_path / "a.txt"))
    assert terms == [4, 7, 2]
    assert sorted(map(tuple, map(sorted, graph.edges()))) == sorted(map(tuple, map(sorted, g.edges())))
    assert all(graph[u][v]["weight"] == g[u][v]["weight"] and graph[u][v]["capacity"] == 1.0 for u, v in g.edges())
    gb, tb = TextGraphLoader.load_graph_from_text(str(tmp_path / "b.txt"))
    out = capsys.readouterr().out
    assert gb.number_of_edges() == 3 and gb[1][2]["weight"] == 2.5 and gb[0][1]["weight"] == 1.0 and tb == [0, 1, 2]
    assert "Warning: Invalid line 5" in out and "Warning: Could not parse line 6" in out
    with pytest.raises(ValueError):
        TextGraphLoader.load_graph_from_text(str(tmp_path / "c.txt"))
    with pytest.raises(ValueError):
        TextGraphLoader.load_graph_from_text(str(tmp_path / "empty.txt"))
    graphs, terminals = TextGraphLoader.load_all_graphs(str(tmp_path))
    assert sorted(graphs) == ["a.txt", "b.txt"] and terminals["a.txt"] == [4, 7, 2]
    assert "Successfully loaded 2 graphs" in capsys.readouterr().out
    with pytest.raises(ValueError):
        TextGraphLoader.load_all_graphs(str(tmp_path / "nope"))
    # loader output feeds the extender unchanged
    from DataGenerator import graphExtender as E
    ds = E.process_graphs_from_folder({"a.txt": graph}, {"a.txt": terms}, max_nodes=16)
    assert len(ds) == 1 and ds[0][1].shape == (10, 16)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,kw,expected", [
    (500, 42, dict(max_iterations=1000, threshold=1, patience=50), 1393),     # randomizedAlgo.ipynb:L113
    (1000, 42, dict(max_iterations=2000, threshold=1, patience=100), 2741),   # :L186
    (1000, 42, dict(max_iterations=2000, threshold=1, patience=100,
                    fixed_terminals={0: 0, 1: 1, 2: 2}), 2738),               # :L275
    (1000, 123, dict(max_iterations=2000, threshold=1, patience=50), 2724),   # :L486-497
    (1000, 123, dict(max_iterations=2000, threshold=1, patience=25), 2719),
    (1000, 123, dict(max_iterations=2000, threshold=1, patience=200), 2742),
])
def test_gpu_randomized_baseline_reproduces_known_answers(n, seed, kw, expected):
    from RandomAlgorithm import RandomizedMaxCut as R
    g = R.create_random_regular_graph(n=n, degree=8, random_seed=seed)
    cut, part = R.randomized_k_way_maxcut(g, k=3, random_seed=seed, **kw)
    state_gpu = random.getstate()
    want_cut, want_part = pp.py_randomized_k_way_maxcut(g, k=3, random_seed=seed, **kw)
    assert cut == expected == want_cut
    assert part == want_part and list(part.keys()) == list(want_part.keys())
    assert random.getstate() == state_gpu            # generator left exactly where the reference leaves it
    assert R.calculate_cut_value(g, part) == expected


@pytest.mark.gpu
def test_gpu_randomized_baseline_partition_sizes_and_evaluate():
    from RandomAlgorithm import RandomizedMaxCut as R
    g = R.create_random_regular_graph(n=500, degree=8, random_seed=42)
    _, part = R.randomized_k_way_maxcut(g, k=3, max_iterations=1000, threshold=1, patience=50, random_seed=42)
    assert np.bincount(list(part.values()), minlength=3).tolist() == [153, 171, 176]       # ipynb:L113
    graphs = [R.create_random_regular_graph(n=200, degree=6, random_seed=100 + i) for i in range(3)]
    random.seed(5)
    res = R.evaluate_algorithm_on_graphs(graphs, k=3, max_iterations=200, threshold=1, patience=20)
    random.seed(5)
    want = [pp.py_randomized_k_way_maxcut(g2, 3, 200, 1, 20)[0] for g2 in graphs]
    assert res["cut_values"] == want


@pytest.mark.gpu
def test_batched_testing_harness_equals_per_graph_loop():
    from DataGenerator import graphExtender as E
    from Testing import TestingNeuralNetwork as Te
    from Training import TrainingNeural as T
    import contextlib, io
    random.seed(3)
    graphs, terms = {}, {}
    for size in (40, 60):
        for i in range(3):
            name = f"test_n{size}_{i}"
            graphs[name] = nx.random_regular_graph(d=5 + i, n=size, seed=size + i)
            nx.set_edge_attributes(graphs[name], 1, "weight")
            terms[name] = random.sample(range(size), 3)
    graphs["test_n77_0"] = nx.random_regular_graph(d=4, n=78, seed=1)      # size not configured -> skipped
    nx.set_edge_attributes(graphs["test_n77_0"], 1, "weight")
    terms["test_n77_0"] = [5, 6, 7]
    with contextlib.redirect_stdout(io.StringIO()):
        ds = E.process_graphs_from_folder(graphs, terms, max_nodes=128)
    cfg = T.TrainingConfig(n_nodes=128, dim_embedding=128, hidden_dim=32)
    torch.manual_seed(1)
    net, _, _ = T.setup_model_and_optimizer(cfg)
    net.eval()
    np.random.seed(11)
    loop, by_size_loop = Te.test_multiple_graphs(net, ds, [40, 60], post_processing_iterations=30, verbose=False)
    after_loop = np.random.rand()
    np.random.seed(11)
    fast, by_size_fast = Te.test_multiple_graphs_batched(net, ds, [40, 60], post_processing_iterations=30,
                                                         greedy_iterations=50)
    assert np.random.rand() == after_loop
    assert len(loop) == len(fast) == 6
    for a, b in zip(loop, fast):
        for key in ("simple_cut", "simple_assignment", "post_cut", "post_assignment", "improvement", "graph_name",
                    "graph_size", "nodes", "edges", "terminals"):
            assert a[key] == b[key], key
        np.testing.assert_allclose(a["node_probabilities"], b["node_probabilities"], rtol=1e-5, atol=1e-7)
        assert b["greedy_cut"] >= b["post_cut"] and b["greedy_assignment"][:3] == [0, 1, 2]
    for s in (40, 60):
        assert by_size_loop[s]["simple"]["cut_values"] == by_size_fast[s]["simple"]["cut_values"]
        assert by_size_loop[s]["post_processed"]["cut_values"] == by_size_fast[s]["post_processed"]["cut_values"]
