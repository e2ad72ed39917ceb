"""graphExtender -- B200 drop-in for the reference's python/DataGenerator/graphExtender.py.

Produces the dataset tuples the training / testing hot path consumes:
    dataset[i] = [graph_handle, X float32 [n, max_nodes], nx.Graph, [0, 1, 2]]     (reference :114)
with identical terminal normalisation (in-place node swaps, the same four cases, the same skip
rule, the same mutation of the caller's graph and terminal list; reference :68-97), prints and
batch pickles.  What changed underneath: the graph handle is a gmc_b200 CSRGraph built straight
from the edge list instead of a DGLGraph, and X is filled from the edge list in O(|E|) instead
of through an O(n^2) Python dictionary (commons.py:38-77) -- same values, bit for bit.
"""
try:
    from python.commons import *  # noqa: F401,F403  (reference spelling, :1)
    from python.commons import adjacency_tensor
except ImportError:
    from commons import *  # noqa: F401,F403
    from commons import adjacency_tensor

import traceback
from typing import Dict, List, Optional, Tuple  # noqa: F401

import torch

from gmc_b200.graph import AdjacencyFeatures

TORCH_DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
TORCH_DTYPE = torch.float32


def swap_graph_nodes(graph, mapping):
    """Relabel `graph` in place so that every key of `mapping` ends up carrying the label
    `mapping` sends it to, via fresh temporary labels (reference :8-26)."""
    first_free = max(graph.nodes) + 1
    parked = {old: first_free + i for i, old in enumerate(mapping)}
    nx.relabel_nodes(graph, mapping=parked, copy=False)
    came_from = {new: old for old, new in mapping.items()}
    nx.relabel_nodes(graph, mapping={parked[old]: came_from[old] for old in mapping}, copy=False)


def extend_matrix_torch_2(matrix, N, torch_dtype=None, torch_device=None):
    """[n, n] -> [n, N] with zero columns appended; ValueError if N < n (reference :28-48)."""
    n = matrix.shape[0]
    if N < n:
        raise ValueError("N should be greater than or equal to the original matrix size.")
    out = torch.zeros((n, N), dtype=torch_dtype or matrix.dtype, device=torch_device or matrix.device)
    out[:, :n] = matrix
    return out


def _terminal_swap_plan(terminals: List[int]) -> Optional[Dict[int, int]]:
    """The reference's four normalisation cases (:72-97).  Returns the swap mapping, or None when
    the graph must be skipped (two or more of {0,1,2} already terminals).  Sorts `terminals` in
    place in the same cases the reference does."""
    has = [t in terminals for t in (0, 1, 2)]
    if not any(has):
        t = terminals
        return {t[0]: 0, t[1]: 1, t[2]: 2, 0: t[0], 1: t[1], 2: t[2]}
    if has == [False, False, True]:
        terminals.sort()
        return {terminals[1]: 0, terminals[2]: 1, 0: terminals[1], 1: terminals[2]}
    if has == [False, True, False]:
        terminals.sort()
        return {terminals[1]: 0, terminals[2]: 2, 0: terminals[1], 2: terminals[2]}
    if has == [True, False, False]:
        terminals.sort()
        return {terminals[1]: 1, terminals[2]: 2, 1: terminals[1], 2: terminals[2]}
    return None


def process_graphs_from_folder(all_graphs: Dict, all_terminals: Dict, max_nodes: int,
                               save_batch_size: Optional[int] = None,
                               output_filename_prefix: str = "processed_graphs", dense_features: bool = True) -> Dict:
    """Normalise terminals to nodes 0,1,2 and emit dataset tuples (reference :50-132).
    dense_features=False (extension) puts an AdjacencyFeatures stand-in into slot [1] instead of the dense
    [n, max_nodes] tensor (4 MB per graph at max_nodes = 1000): same 4-tuple format, the training / testing entry
    points rebuild the features on the device from the graph, `.dense()` yields the reference's tensor."""
    datasetItem = {}
    i = 0
    skipped = 0
    filename, graph, terminals = None, None, None
    try:
        for filename, graph in all_graphs.items():
            terminals = all_terminals[filename]
            plan = _terminal_swap_plan(terminals)
            if plan is None:
                skipped += 1
                continue
            swap_graph_nodes(graph, plan)
            print(f"Terminal swapped {i}")

            handle = dgl.from_networkx(nx_graph=graph).to(TORCH_DEVICE)
            if dense_features:
                full_matrix = adjacency_tensor(graph, max_nodes, TORCH_DTYPE)
            else:
                if max_nodes < graph.number_of_nodes():
                    raise ValueError("N should be greater than or equal to the original matrix size.")
                full_matrix = AdjacencyFeatures(handle, max_nodes)

            datasetItem[i] = [handle, full_matrix, graph, [0, 1, 2]]
            i += 1
            if save_batch_size and (i % save_batch_size == 0):
                batch_filename = f"{output_filename_prefix}_{i}.pkl"
                save_object(datasetItem, batch_filename)
                print(f"Saved batch to {batch_filename}")
                datasetItem = {}
            print(f"Graph finished: {i}")
    except Exception:  # noqa: BLE001 - reference behaviour: print and return what was built (:126-129)
        print(f"Exception occurred at graph {i}, filename {filename}, terminals {terminals}")
        print(f"Graph nodes: {graph.number_of_nodes() if graph is not None else 'n/a'}")
        print(traceback.format_exc())
    print(f"Skipped items: {skipped}")
    return datasetItem


def load_and_process_graphs(graphs_filename: str, terminals_filename: str, max_nodes: int, output_filename: str,
                            save_batch_size: Optional[int] = None) -> None:
    """Pickle in, pickle out (reference :134-161)."""
    print(f"Loading graphs from {graphs_filename}")
    all_graphs = open_file(graphs_filename)
    print(f"Loading terminals from {terminals_filename}")
    all_terminals = open_file(terminals_filename)
    print(f"Processing {len(all_graphs)} graphs with max_nodes={max_nodes}")
    processed = process_graphs_from_folder(all_graphs, all_terminals, max_nodes, save_batch_size=save_batch_size,
                                           output_filename_prefix=output_filename.replace(".pkl", ""))
    if processed:
        print(f"Saving final dataset to {output_filename}")
        save_object(processed, output_filename)


def save_processed_graphs(processed_data: Dict, filename: str) -> None:
    save_object(processed_data, filename)
    print(f"Saved processed graphs to {filename}")
