"""commons -- star-import hub of the B200 drop-in (mirror of the reference's python/commons.py).

The reference's modules start with `from python.commons import *` (TrainingNeural.py:25,
graphExtender.py:1) and expect this namespace to provide torch / nn / F / nx / np / dgl /
GraphConv / chain / permutations / ... (commons.py:1-20) plus the pickle and adjacency helpers
(commons.py:22-77).  Here `GraphConv` and `dgl.from_networkx` are the libgcnmaxcut-backed
implementations from gmc_b200; nothing in this tree imports DGL.

Both import spellings resolve:
    sys.path += [<repo>/gcn-max-cut_b200]          ->  from python.commons import *
    sys.path += [<repo>/gcn-max-cut_b200/python]   ->  from commons import *
"""
import os
import pickle
import random
import sys
import types
from collections import OrderedDict, defaultdict
from itertools import chain, combinations, islice, permutations
from time import time

import networkx as nx
import numpy as np
import torch
import torch as th
import torch.nn as nn
import torch.nn.functional as F

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:          # make `gmc_b200` importable under the notebook spelling too
    sys.path.insert(0, _PKG_ROOT)

from gmc_b200.graph import CSRGraph, GraphBatch, from_networkx as _from_networkx  # noqa: E402
from gmc_b200.model import GraphConv  # noqa: E402

try:  # plotting is out of scope for the hot path; keep the names when matplotlib exists
    import matplotlib  # noqa: F401
    import matplotlib.pyplot as plt  # noqa: F401
except Exception:  # pragma: no cover - matplotlib is optional here
    matplotlib = None
    plt = None

# `dgl.from_networkx(...)` / `dgl.nn.pytorch.GraphConv` call sites keep working against this shim
dgl = types.SimpleNamespace(
    from_networkx=_from_networkx,
    nn=types.SimpleNamespace(pytorch=types.SimpleNamespace(GraphConv=GraphConv)),
    DGLGraph=CSRGraph,
    __version__="gmc_b200",
)


# ---------------------------------------------------------------------------- pickle I/O
def save_object(obj, filename):
    """Pickle `obj` to `filename` with the highest protocol (reference commons.py:22-24)."""
    with open(filename, "wb") as fh:
        pickle.dump(obj, fh, pickle.HIGHEST_PROTOCOL)


class _ForeignObject:
    """Stand-in for classes/functions of modules that are not installed (legacy pickles hold real
    dgl.DGLGraph objects, graphExtender.py:102-114).  Swallows any construction/state protocol."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return _ForeignObject()

    def __setstate__(self, state):
        pass

    def __reduce__(self):  # pragma: no cover - never re-pickled
        return (_ForeignObject, ())


class _TolerantUnpickler(pickle.Unpickler):
    _FOREIGN_ROOTS = ("dgl",)

    def find_class(self, module, name):
        if module.split(".")[0] in self._FOREIGN_ROOTS:
            try:
                return super().find_class(module, name)
            except Exception:
                return _ForeignObject
        return super().find_class(module, name)


def _adopt_legacy_handles(obj):
    """Replace foreign graph handles in dataset tuples `[graph, X, nx_graph, terminals]` by a
    CSRGraph rebuilt from the networkx graph stored next to them."""
    if isinstance(obj, dict):
        for key, item in obj.items():
            if isinstance(item, (list, tuple)) and len(item) == 4 and isinstance(item[2], nx.Graph) \
                    and not isinstance(item[0], CSRGraph):
                rebuilt = [_from_networkx(item[2]), item[1], item[2], item[3]]
                obj[key] = rebuilt if isinstance(item, list) else tuple(rebuilt)
    return obj


def open_file(filename):
    """Unpickle `filename` (reference commons.py:26-36).  Datasets written by the reference with
    real DGL graphs load without DGL: their graph handles are rebuilt as CSRGraph."""
    with open(filename, "rb") as fh:
        try:
            data = pickle.load(fh)
        except ModuleNotFoundError:
            fh.seek(0)
            data = _TolerantUnpickler(fh).load()
    return _adopt_legacy_handles(data)


# ---------------------------------------------------------------------------- adjacency helpers
def gen_adj_matrix(nx_G):
    """Dense symmetric adjacency as a {(u, v): weight} mapping with explicit zeros for every
    ordered node pair (reference commons.py:65-77)."""
    adj = defaultdict(int)
    for u, v in nx_G.edges:
        w = nx_G[u][v]["weight"]
        adj[(u, v)] = w
        adj[(v, u)] = w
    nodes = list(nx_G.nodes)
    for u in nodes:
        for v in nodes:
            adj.setdefault((u, v), 0)
    return adj


def qubo_dict_to_torch(nx_G, Q, torch_dtype=None, torch_device=None):
    """{(row, col): value} -> [n, n] tensor (reference commons.py:38-63), vectorised fill."""
    n = len(nx_G.nodes)
    mat = torch.zeros(n, n)
    if len(Q):
        keys = np.fromiter((c for rc in Q.keys() for c in rc), dtype=np.int64, count=2 * len(Q)).reshape(-1, 2)
        vals = np.fromiter(Q.values(), dtype=np.float32, count=len(Q))
        mat[torch.from_numpy(keys[:, 0]), torch.from_numpy(keys[:, 1])] = torch.from_numpy(vals)
    if torch_dtype is not None:
        mat = mat.type(torch_dtype)
    if torch_device is not None:
        mat = mat.to(torch_device)
    return mat


def adjacency_tensor(nx_G, width=None, torch_dtype=torch.float32):
    """Fast path used by graphExtender: padded adjacency rows [n, width] straight from the edge
    list (O(|E|)), equal to qubo_dict_to_torch(gen_adj_matrix(G)) zero-padded to `width`.
    Node labels must be 0..n-1 (graphExtender relabels terminals in place, so they are)."""
    n = nx_G.number_of_nodes()
    width = n if width is None else width
    if width < n:
        raise ValueError("N should be greater than or equal to the original matrix size.")
    mat = torch.zeros((n, width), dtype=torch_dtype)
    m = nx_G.number_of_edges()
    if m:
        e = np.empty((m, 2), dtype=np.int64)
        w = np.empty(m, dtype=np.float32)
        for i, (u, v, data) in enumerate(nx_G.edges(data=True)):
            e[i, 0], e[i, 1], w[i] = u, v, data["weight"]
        if e.min() < 0 or e.max() >= n:
            raise ValueError("adjacency_tensor needs node labels 0..n-1")
        r = torch.from_numpy(np.concatenate([e[:, 0], e[:, 1]]))
        c = torch.from_numpy(np.concatenate([e[:, 1], e[:, 0]]))
        mat[r, c] = torch.from_numpy(np.concatenate([w, w])).to(torch_dtype)
    return mat
