"""textGraphLoader -- the on-disk text graph format next to the hot path.

Format (reference python/DataGenerator/prepareData.ipynb, cell 2, class TextGraphLoader):
    line 1      : [t1, t2, t3]            terminal node ids
    other lines : from_node to_node [w]   one undirected edge per line, optional weight
Blank lines are skipped; lines with fewer than two fields or non-integer nodes are reported with a
"Warning:" print and skipped; a missing / malformed terminal line raises ValueError -- the same
behaviour as the notebook class.  The result feeds graphExtender.process_graphs_from_folder unchanged.
"""
from pathlib import Path
from typing import Dict, List, Tuple

import networkx as nx


class TextGraphLoader:
    """Loads graphs with terminal information from text files."""

    @staticmethod
    def load_graph_from_text(file_path: str, edge_weight: float = 1.0,
                             edge_capacity: float = 1.0) -> Tuple[nx.Graph, List[int]]:
        with open(file_path, "r") as fh:
            lines = fh.readlines()
        if not lines:
            raise ValueError(f"File {file_path} is empty")
        head = lines[0].strip()
        if not (head.startswith("[") and head.endswith("]")):
            raise ValueError(f"Invalid terminal format in {file_path}")
        terminals = [int(tok.strip()) for tok in head[1:-1].split(",")]

        graph = nx.Graph()
        for number, raw in enumerate(lines[1:], 2):
            text = raw.strip()
            if not text:
                continue
            fields = text.split()
            if len(fields) < 2:
                print(f"Warning: Invalid line {number} in {file_path}: {text}")
                continue
            try:
                u, v = int(fields[0]), int(fields[1])
                w = float(fields[2]) if len(fields) >= 3 else edge_weight
            except ValueError:
                print(f"Warning: Could not parse line {number} in {file_path}: {text}")
                continue
            graph.add_edge(u, v, weight=w, capacity=edge_capacity)
        return graph, terminals

    @staticmethod
    def load_all_graphs(directory: str, file_extension: str = ".txt") -> Tuple[Dict[str, nx.Graph], Dict[str, List[int]]]:
        root = Path(directory)
        if not root.exists():
            raise ValueError(f"Directory {directory} does not exist")
        graphs: Dict[str, nx.Graph] = {}
        terminals: Dict[str, List[int]] = {}
        files = list(root.glob(f"*{file_extension}"))
        if not files:
            print(f"Warning: No {file_extension} files found in {directory}")
            return graphs, terminals
        print(f"Loading {len(files)} graph files...")
        for path in files:
            try:
                graph, terms = TextGraphLoader.load_graph_from_text(str(path))
            except Exception as exc:  # noqa: BLE001 - notebook behaviour: report and continue
                print(f"Error loading {path}: {exc}")
                continue
            graphs[path.name] = graph
            terminals[path.name] = terms
        print(f"Successfully loaded {len(graphs)} graphs")
        return graphs, terminals


def write_graph_to_text(graph: nx.Graph, terminals: List[int], file_path: str) -> None:
    """Inverse of load_graph_from_text (used by tests and for exporting synthetic sets)."""
    with open(file_path, "w") as fh:
        fh.write("[" + ", ".join(str(int(t)) for t in terminals) + "]\n")
        for u, v, data in graph.edges(data=True):
            w = data.get("weight", 1)
            fh.write(f"{u} {v} {w:g}\n")
