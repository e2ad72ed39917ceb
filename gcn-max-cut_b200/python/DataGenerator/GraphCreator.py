"""GraphCreator -- seeded networkx graph generators with the reference's API
(python/DataGenerator/GraphCreator.py:31-183).  Input generation only: nothing here is on the
GPU hot path; it exists so configs 1-2 (BASELINE.json) can be rebuilt without the reference tree.
"""
import pickle
import random
from typing import Dict, List, Optional, Tuple

import networkx as nx

_REGULAR = ("reg", "reg_random")
_BINOMIAL = ("prob", "erdos")


def save_object(obj, filename: str) -> None:
    with open(filename, "wb") as fh:
        pickle.dump(obj, fh, pickle.HIGHEST_PROTOCOL)


def load_object(filename: str):
    with open(filename, "rb") as fh:
        return pickle.load(fh)


def generate_graph(n: int, d: Optional[int] = None, p: Optional[float] = None, graph_type: str = "reg",
                   random_seed: int = 0, edge_weight: int = 1, edge_capacity: int = 1) -> nx.Graph:
    """'reg' (seeded d-regular), 'reg_random' (unseeded), 'prob' (fast G(n,p)), 'erdos' (G(n,p));
    nodes 0..n-1 in sorted order, every edge carries `weight` and `capacity` (reference :31-92)."""
    if graph_type in _REGULAR and d is None:
        raise ValueError("Degree 'd' must be provided for regular graphs")
    if graph_type in _BINOMIAL and p is None:
        raise ValueError("Probability 'p' must be provided for probabilistic graphs")
    if n < 1:
        raise ValueError("Number of nodes must be positive")
    if d is not None and d >= n:
        raise ValueError("Degree must be less than number of nodes")
    if p is not None and not 0 <= p <= 1:
        raise ValueError("Probability must be between 0 and 1")

    makers = {
        "reg": lambda: nx.random_regular_graph(d=d, n=n, seed=random_seed),
        "reg_random": lambda: nx.random_regular_graph(d=d, n=n),
        "prob": lambda: nx.fast_gnp_random_graph(n, p, seed=random_seed),
        "erdos": lambda: nx.erdos_renyi_graph(n, p, seed=random_seed),
    }
    if graph_type not in makers:
        raise NotImplementedError(f"Graph type {graph_type} not supported")
    raw = nx.relabel.convert_node_labels_to_integers(makers[graph_type]())

    graph = nx.Graph()
    graph.add_nodes_from(sorted(raw.nodes()))
    graph.add_edges_from(raw.edges)
    nx.set_edge_attributes(graph, edge_weight, "weight")
    nx.set_edge_attributes(graph, edge_capacity, "capacity")
    return graph


def generate_unique_terminals(n: int, num_terminals: int = 3) -> List[int]:
    """`num_terminals` distinct nodes from Python's global `random` stream (reference :93-109)."""
    if n < num_terminals:
        raise ValueError(f"Graph size ({n}) must be >= number of terminals ({num_terminals})")
    return random.sample(range(n), num_terminals)


def generate_graph_dataset(num_graphs: int, min_nodes: int, max_nodes: int, min_degree: int, max_degree: int,
                           graph_type: str = "reg", num_terminals: int = 3, edge_weight: int = 1,
                           edge_capacity: int = 1) -> Tuple[Dict[int, nx.Graph], Dict[int, List[int]]]:
    """Random sizes/degrees from the global `random` stream, graph i seeded with i; odd n*d draws
    are retried, at most 2*num_graphs failed attempts (reference :112-183)."""
    graphs: Dict[int, nx.Graph] = {}
    terminals: Dict[int, List[int]] = {}
    made, failed = 0, 0
    while made < num_graphs and failed < num_graphs * 2:
        try:
            nodes = random.randint(min_nodes, max_nodes)
            degree = random.randint(min_degree, max_degree)
            if graph_type in _REGULAR and (nodes * degree) % 2:
                failed += 1
                continue
            graphs[made] = generate_graph(n=nodes, d=degree, graph_type=graph_type, random_seed=made,
                                          edge_weight=edge_weight, edge_capacity=edge_capacity)
            terminals[made] = generate_unique_terminals(nodes, num_terminals)
            made += 1
        except (ValueError, nx.NetworkXError):
            failed += 1
    if made < num_graphs:
        print(f"Warning: Only generated {made} valid graphs out of {num_graphs} requested")
    return graphs, terminals


def save_graphs_to_pickle(graphs: Dict[int, nx.Graph], filename: str) -> None:
    save_object(graphs, filename)


def save_terminals_to_pickle(terminals: Dict[int, List[int]], filename: str) -> None:
    save_object(terminals, filename)
