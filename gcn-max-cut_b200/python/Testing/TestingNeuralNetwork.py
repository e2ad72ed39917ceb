"""TestingNeuralNetwork -- B200 drop-in for the inference + post-processing half of the
reference's python/Testing/TestingNeuralNetwork.py (:18-295).

`test_single_graph` / `test_multiple_graphs` / `post_processing_optimization` /
`simple_partition_assignment` run on libgcnmaxcut's integer kernels (gmc_argmax_labels,
gmc_cut_value_i32, gmc_sample_best_cut) and return exactly what the reference returns -- the
same result-dict schema, the same integer cut values and, for a given numpy RNG state, the same
sampled partitions (the uniforms are drawn on the host with `np.random.rand` in the reference's
call order, so the global RNG ends in the same state).  `greedy_local_search` adds the
north-star node-move search.  The two pure-Python helpers `assign_partitions` and
`calculate_cut_value` keep their list/networkx signatures for callers that use them directly.
"""
import os  # noqa: F401
import random  # noqa: F401
from time import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

try:
    from python import commons as _commons
except ImportError:  # notebook spelling
    import commons as _commons

from gmc_b200 import _lib as _gmc_lib
from gmc_b200 import ops as _ops
from gmc_b200.graph import CSRGraph, GraphBatch
from gmc_b200.model import as_batch as _as_batch


def _numpy_compares_in_f32() -> bool:
    """`python_float < np.float32` is a float32 comparison under NEP 50 (numpy >= 2) and a
    float64 one under the reference's pinned numpy 1.x; follow the numpy that is running."""
    return int(np.__version__.split(".")[0]) >= 2


# ---------------------------------------------------------------- host helpers (API parity)
def assign_partitions(node_probs: np.ndarray) -> List[int]:
    """One categorical sampling: nodes 0,1,2 -> 0,1,2, every other node takes the first class whose
    running probability sum exceeds a fresh np.random.rand() (reference :18-46)."""
    labels = [0, 1, 2]
    for row in node_probs[3:]:
        r = np.random.rand()
        acc = 0
        pick = len(row) - 1
        for k, p in enumerate(row):
            acc += p
            if r < acc:
                pick = k
                break
        labels.append(pick)
    return labels


def calculate_cut_value(partition_assignment: List[int], graph) -> int:
    """Total weight of edges whose end points carry different labels (reference :48-64)."""
    m = len(partition_assignment)
    value = 0
    for u, v, data in graph.edges(data=True):
        if u < m and v < m and partition_assignment[u] != partition_assignment[v]:
            value += data.get("weight", 1)
    return value


# ---------------------------------------------------------------- device plumbing
_NX_BATCHES: Dict[int, tuple] = {}


def _batch_of(nx_graph, handle=None) -> GraphBatch:
    """One-graph device batch for a networkx graph (cached on identity + edge count)."""
    if isinstance(handle, (CSRGraph, GraphBatch)):
        return _as_batch(handle)
    key = id(nx_graph)
    hit = _NX_BATCHES.get(key)
    sig = (nx_graph.number_of_nodes(), nx_graph.number_of_edges())
    if hit is None or hit[0] is not nx_graph or hit[1] != sig:
        if len(_NX_BATCHES) > 512:
            _NX_BATCHES.clear()
        hit = (nx_graph, sig, GraphBatch([CSRGraph.from_networkx(nx_graph)], check_degrees=False))
        _NX_BATCHES[key] = hit
    return hit[2]


def _device_probs(node_probabilities) -> torch.Tensor:
    dev = _gmc_lib.require_cuda()
    if isinstance(node_probabilities, torch.Tensor):
        return node_probabilities.detach().to(device=dev, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.asarray(node_probabilities), dtype=torch.float32).to(dev).contiguous()


# draws per call from which the uniforms of the sampling post-processing come from the device generator (GMC_DEVICE_RNG=0: never)
_DEVICE_RNG_MIN = 20000 if os.environ.get("GMC_DEVICE_RNG", "1") != "0" else (1 << 62)


def _edge_count(handle, nx_graph) -> int:
    """len(nx_graph.edges()) -- an O(n) Python walk in networkx, half of a batched test pass once everything else is on the
    device -- remembered on the dataset item's graph handle for the networkx object it was counted on."""
    memo = getattr(handle, "_nx_edges", None)
    if memo is not None and memo[0] == id(nx_graph):
        return memo[1]
    count = len(nx_graph.edges())
    try:
        handle._nx_edges = (id(nx_graph), count)
    except AttributeError:
        pass
    return count


def _sample_best(batch: GraphBatch, probs: torch.Tensor, iterations: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """P1 on the device for every graph of `batch`; uniforms drawn per graph in dataset order with
    np.random.rand (same stream as the reference's scalar draws)."""
    sizes = batch.sizes
    counts = [iterations * max(int(n) - 3, 0) for n in sizes]
    u_ptr = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=u_ptr[1:])
    total = int(u_ptr[-1])
    dev = probs.device
    U_d = None
    if total >= _DEVICE_RNG_MIN:
        # the same np.random.rand stream, generated on the device (MT19937 is 10 ms of host time per 2 M draws, plus the
        # upload); the host generator is left where the scalar draws would have left it
        U_d = _ops.numpy_rand_on_device(total, dev)
    if U_d is None:
        U = np.random.rand(total) if total else np.zeros(0, dtype=np.float64)
        U_d = torch.from_numpy(np.ascontiguousarray(U, dtype=np.float64)).to(dev)
    if U_d.numel() == 0:
        U_d = torch.zeros(1, dtype=torch.float64, device=dev)
    labels, best, _ = _ops.sample_best_cut(batch, probs, U_d, torch.from_numpy(u_ptr).to(dev), iterations,
                                           _numpy_compares_in_f32())
    return labels, best


def post_processing_optimization(node_probabilities, graph, iterations: int = 200) -> Tuple[List[int], int]:
    """Best of `iterations` random samplings (reference :66-98) on the GPU."""
    if iterations <= 0:
        return None, -float("inf")
    batch = _batch_of(graph)
    probs = _device_probs(node_probabilities)
    if probs.shape[0] != batch.num_nodes:
        raise ValueError("node_probabilities rows must equal the number of graph nodes")
    labels, best = _sample_best(batch, probs, iterations)
    return labels.cpu().numpy().tolist(), int(best.item())


def simple_partition_assignment(node_probabilities) -> List[int]:
    """argmax per node with nodes 0,1,2 forced to 0,1,2 (reference :100-122)."""
    probs = _device_probs(node_probabilities)
    n = probs.shape[0]
    gp = torch.tensor([0, n], dtype=torch.int32, device=probs.device)
    labels = torch.empty(n, dtype=torch.int32, device=probs.device)
    _gmc_lib.check(_gmc_lib.lib().gmc_argmax_labels(probs.data_ptr(), probs.shape[1], gp.data_ptr(), 1, n,
                                                    probs.shape[1], 1, labels.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "gmc_argmax_labels")
    return labels.cpu().numpy().tolist()


def greedy_local_search(partition_assignment, graph, iterations: int = 200, num_classes: int = 3,
                        frozen_terminals: int = 3) -> Tuple[List[int], int]:
    """North-star post-processing: greedy best-improvement node moves (k-way `greedy_maxcut`,
    "Other Algorithms/huerestics_multi-max.ipynb":L5817-5854) from a given assignment."""
    batch = _batch_of(graph)
    dev = batch.device
    start = torch.as_tensor(np.asarray(partition_assignment, dtype=np.int32)).to(dev)
    labels, cut, _ = _ops.greedy_node_move(batch, start, num_classes, iterations, frozen_terminals)
    return labels.cpu().numpy().tolist(), int(cut.item())


# ---------------------------------------------------------------- harness
def test_single_graph(model, dgl_graph, adjacency_matrix, nx_graph, terminals: List[int],
                      post_processing_iterations: int = 200) -> Dict[str, Any]:
    """Argmax and post-processed cuts for one graph (reference :124-186); errors are swallowed into
    {'success': False, 'error': ...} like the reference."""
    try:
        with torch.no_grad():
            node_probabilities = model(dgl_graph, adjacency_matrix)
        batch = _batch_of(nx_graph, dgl_graph)
        probs = _device_probs(node_probabilities)

        simple_start_time = time()
        labels = _ops.argmax_labels(batch, probs, force_terminals=True)
        simple_cut_value = int(_ops.cut_value(batch, labels).item())
        simple_assignment = labels.cpu().numpy().tolist()
        simple_time = time() - simple_start_time

        post_start_time = time()
        if post_processing_iterations > 0:
            best_labels, best = _sample_best(batch, probs, post_processing_iterations)
            post_assignment, post_cut_value = best_labels.cpu().numpy().tolist(), int(best.item())
        else:
            post_assignment, post_cut_value = None, -float("inf")
        post_time = time() - post_start_time

        improvement = post_cut_value - simple_cut_value
        improvement_percent = (improvement / simple_cut_value * 100) if simple_cut_value > 0 else 0
        return {
            "success": True,
            "nodes": len(nx_graph.nodes()),
            "edges": len(nx_graph.edges()),
            "simple_cut": simple_cut_value,
            "simple_time": simple_time,
            "simple_assignment": simple_assignment,
            "post_cut": post_cut_value,
            "post_time": post_time,
            "post_assignment": post_assignment,
            "improvement": improvement,
            "improvement_percent": improvement_percent,
            "terminals": terminals,
            "node_probabilities": node_probabilities.detach().cpu().numpy(),
        }
    except Exception as e:  # noqa: BLE001 - reference behaviour (:180-186)
        return {"success": False, "error": str(e),
                "nodes": len(nx_graph.nodes()) if nx_graph else 0,
                "edges": len(nx_graph.edges()) if nx_graph else 0}


def _size_bucket(key, nx_graph, graph_sizes: List[int]) -> Tuple[str, int]:
    """Graph name and size category (reference :229-245): string keys 'test_n{size}_...' are
    parsed, integer keys are matched to the closest configured size within +-5 nodes."""
    if isinstance(key, str):
        try:
            return key, int(key.split("_")[1][1:])
        except (IndexError, ValueError):
            return key, len(nx_graph.nodes())
    size = len(nx_graph.nodes())
    closest = min(graph_sizes, key=lambda s: abs(s - size))
    if abs(closest - size) <= 5:
        size = closest
    return f"graph_{key}", size


def test_multiple_graphs(model, processed_graphs: Dict, graph_sizes: List[int], post_processing_iterations: int = 200,
                         verbose: bool = True) -> Tuple[List[Dict], Dict]:
    """Run test_single_graph over a processed dataset and bucket results by size (reference :188-295)."""
    if verbose:
        print("Testing neural network performance...")
        print("=" * 60)
    test_results: List[Dict] = []
    results_by_size = {size: {"simple": {"cut_values": [], "times": []},
                              "post_processed": {"cut_values": [], "times": []}} for size in graph_sizes}
    total_graphs = len(processed_graphs)
    processed_count = 0
    if verbose:
        print(f"Sample keys from processed_graphs: {list(processed_graphs.keys())[:3]}")

    for key, (dgl_graph, adjacency_matrix, nx_graph, terminals) in processed_graphs.items():
        processed_count += 1
        graph_name, graph_size = _size_bucket(key, nx_graph, graph_sizes)
        if verbose:
            print(f"\\nProcessing graph {processed_count}/{total_graphs}: {graph_name}")
            print(f"  Nodes: {len(nx_graph.nodes())}, Edges: {len(nx_graph.edges())}, Size category: {graph_size}")
        if graph_size not in graph_sizes:
            if verbose:
                print(f"  Skipping: graph size {graph_size} not in test configuration")
            continue

        result = test_single_graph(model, dgl_graph, adjacency_matrix, nx_graph, terminals, post_processing_iterations)
        if result["success"]:
            result.update({"graph_name": graph_name, "graph_size": graph_size})
            test_results.append(result)
            bucket = results_by_size[graph_size]
            bucket["simple"]["cut_values"].append(result["simple_cut"])
            bucket["simple"]["times"].append(result["simple_time"])
            bucket["post_processed"]["cut_values"].append(result["post_cut"])
            bucket["post_processed"]["times"].append(result["post_time"])
            if verbose:
                print(f"  Simple GCN:      Cut = {result['simple_cut']}, Time = {result['simple_time']:.4f}s")
                print(f"  Post-processed:  Cut = {result['post_cut']}, Time = {result['post_time']:.4f}s")
                print(f"  Improvement:     {result['improvement']:+d} ({result['improvement_percent']:+.1f}%)")
        elif verbose:
            print(f"  ✗ Error processing graph: {result['error']}")
        if verbose and processed_count % 10 == 0:
            progress = (processed_count / total_graphs) * 100
            print(f"\\n--- Progress: {processed_count}/{total_graphs} ({progress:.1f}%) ---")

    if verbose:
        print(f"\\n{'=' * 60}")
        print("Neural network testing completed!")
        print(f"Successfully processed: {len(test_results)}/{total_graphs} graphs")
    return test_results, results_by_size


def test_multiple_graphs_batched(model, processed_graphs: Dict, graph_sizes: List[int],
                                 post_processing_iterations: int = 200, verbose: bool = False,
                                 greedy_iterations: int = 0) -> Tuple[List[Dict], Dict]:
    """Same results as test_multiple_graphs (identical cuts, assignments and final numpy RNG state), computed
    for ALL selected graphs at once: one block-diagonal forward pass, one argmax + integer-cut launch, one
    best-of-N sampling launch (uniforms drawn per graph in dataset order, as the per-graph loop would).
    The per-graph `simple_time` / `post_time` fields hold the batch time divided by the number of graphs.
    With greedy_iterations > 0 each result also carries `greedy_cut` / `greedy_assignment`: the north-star
    node-move local search started from the post-processed assignment."""
    from gmc_b200.engine import GCNEngine
    from gmc_b200.model import to_device_features
    dev = _gmc_lib.require_cuda()
    results_by_size = {size: {"simple": {"cut_values": [], "times": []},
                              "post_processed": {"cut_values": [], "times": []}} for size in graph_sizes}
    chosen = []
    for key, (handle, X, nx_graph, terminals) in processed_graphs.items():
        name, size = _size_bucket(key, nx_graph, graph_sizes)
        if size in graph_sizes:
            chosen.append((name, size, handle, X, nx_graph, terminals))
        elif verbose:
            print(f"  Skipping {name}: graph size {size} not in test configuration")
    if not chosen:
        return [], results_by_size
    handles = [h if isinstance(h, CSRGraph) else CSRGraph.from_networkx(g) for _, _, h, _, g, _ in chosen]
    batch = GraphBatch(handles, device=dev)
    # features are verified per item to be the adjacency rows of their graph and then rebuilt on the device from the
    # batch (no concatenation of the items' dense host tensors); anything else keeps the uploaded tensors
    from gmc_b200.graph import check_adjacency_features
    try:
        widths = {check_adjacency_features(h, X) for h, (_, _, _, X, _, _) in zip(handles, chosen)}
    except NotImplementedError:
        widths = set()
    if len(widths) == 1:
        width = widths.pop()
        X_all = _ops.densify(batch, width, out=_ops.padded_empty(batch.num_nodes, width, dev))
    else:
        feats = [to_device_features(X, dev) for _, _, _, X, _, _ in chosen]
        X_all = feats[0] if len(feats) == 1 else torch.cat(feats, dim=0)
    engine = GCNEngine(model, None, precision=getattr(model.conv1, "gemm_precision", "fp32"))
    with torch.no_grad():
        probs = engine.forward(batch, X_all).clone()

    t0 = time()
    labels = _ops.argmax_labels(batch, probs, force_terminals=True)
    simple_cuts = _ops.cut_value(batch, labels).cpu().numpy()
    simple_labels = labels.cpu().numpy()
    simple_time = (time() - t0) / len(chosen)

    t0 = time()
    if post_processing_iterations > 0:
        best_labels, best = _sample_best(batch, probs, post_processing_iterations)
        post_cuts, post_labels = best.cpu().numpy(), best_labels.cpu().numpy()
    else:
        post_cuts, post_labels = None, None
    post_time = (time() - t0) / len(chosen)
    greedy = None
    if greedy_iterations > 0 and post_labels is not None:
        g_labels, g_cut, _ = _ops.greedy_node_move(batch, best_labels, probs.shape[1], greedy_iterations, 3)
        greedy = (g_labels.cpu().numpy(), g_cut.cpu().numpy())

    probs_host = probs.cpu().numpy()
    gp = batch.graph_ptr_host
    test_results: List[Dict] = []
    for i, (name, size, _h, _X, nx_graph, terminals) in enumerate(chosen):
        lo, hi = int(gp[i]), int(gp[i + 1])
        s_cut = int(simple_cuts[i])
        p_cut = int(post_cuts[i]) if post_cuts is not None else -float("inf")
        improvement = p_cut - s_cut
        result = {
            "success": True, "nodes": len(nx_graph.nodes()), "edges": _edge_count(_h, nx_graph),
            "simple_cut": s_cut, "simple_time": simple_time, "simple_assignment": simple_labels[lo:hi].tolist(),
            "post_cut": p_cut, "post_time": post_time,
            "post_assignment": post_labels[lo:hi].tolist() if post_labels is not None else None,
            "improvement": improvement,
            "improvement_percent": (improvement / s_cut * 100) if s_cut > 0 else 0,
            "terminals": terminals, "node_probabilities": probs_host[lo:hi].copy(),
            "graph_name": name, "graph_size": size,
        }
        if greedy is not None:
            result["greedy_cut"] = int(greedy[1][i])
            result["greedy_assignment"] = greedy[0][lo:hi].tolist()
        test_results.append(result)
        bucket = results_by_size[size]
        bucket["simple"]["cut_values"].append(s_cut)
        bucket["simple"]["times"].append(simple_time)
        bucket["post_processed"]["cut_values"].append(p_cut)
        bucket["post_processed"]["times"].append(post_time)
        if verbose:
            print(f"  {name}: simple {s_cut}, post-processed {p_cut} ({improvement:+d})")
    return test_results, results_by_size


def analyze_results(test_results: List[Dict], results_by_size: Dict, graph_sizes: List[int]) -> Dict[str, Any]:
    """Aggregate statistics over test results (subset of reference :297-382: the numeric summary;
    report/plot helpers are presentation code and out of scope)."""
    if not test_results:
        return {"error": "No successful test results to analyze"}
    simple = np.asarray([r["simple_cut"] for r in test_results], dtype=np.float64)
    post = np.asarray([r["post_cut"] for r in test_results], dtype=np.float64)
    st = np.asarray([r["simple_time"] for r in test_results], dtype=np.float64)
    pt = np.asarray([r["post_time"] for r in test_results], dtype=np.float64)
    imp = post - simple
    per_size = {}
    for size in graph_sizes:
        b = results_by_size.get(size)
        if b and b["simple"]["cut_values"]:
            per_size[size] = {
                "count": len(b["simple"]["cut_values"]),
                "simple_avg_cut": float(np.mean(b["simple"]["cut_values"])),
                "post_avg_cut": float(np.mean(b["post_processed"]["cut_values"])),
                "simple_avg_time": float(np.mean(b["simple"]["times"])),
                "post_avg_time": float(np.mean(b["post_processed"]["times"])),
            }
    return {
        "total_graphs": len(test_results),
        "simple_avg_cut": float(simple.mean()), "post_avg_cut": float(post.mean()),
        "avg_improvement": float(imp.mean()),
        "avg_improvement_percent": float(np.mean([r["improvement_percent"] for r in test_results])),
        "graphs_improved": int((imp > 0).sum()),
        "simple_avg_time": float(st.mean()), "post_avg_time": float(pt.mean()),
        "time_overhead_factor": float(pt.mean() / st.mean()) if st.mean() > 0 else float("inf"),
        "by_size": per_size,
    }
