"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Minimal stand-in for the two pieces of `dgl==2.0.0` (envList.txt:36) that the
reference's hot path touches, plus an inert `matplotlib` stub, so that the
reference's *own, unmodified* modules can be imported from /root/reference in
the build container (where neither wheel is installable: no network).

    dgl.from_networkx      <- python/DataGenerator/graphExtender.py:102
    dgl.nn.pytorch.GraphConv <- python/commons.py:12, python/Training/TrainingNeural.py:76-77

`dgl` is a third-party dependency that is NOT vendored under /root/reference.
The restatement below follows the published DGL 2.0.0 algorithm
(python/dgl/nn/pytorch/conv/graphconv.py upstream):

  * norm='both', weight=True, bias=True, allow_zero_in_degree=False
  * feat_src = feat * out_degrees().clamp(min=1)^-0.5          (source norm first)
  * if in_feats > out_feats:  feat_src = feat_src @ W, then aggregate
    else:                     aggregate, then @ W
  * aggregation = sum over in-edges (copy_u / sum), no self loops added
  * rst = rst * in_degrees().clamp(min=1)^-0.5, then + bias
  * weight [in, out] xavier_uniform_, bias zeros
  * DGLError if any node has in-degree 0

`from_networkx`: nodes relabelled with ordering='sorted', undirected graphs
expanded to both directions (2|E| edges; complete_training_pipeline.ipynb:L442).

No reference test pins GraphConv numerics ("parity unpinned at the DGL
boundary"); this shim is self-checked against the dense
D^-1/2 A D^-1/2 X W + b form in tests/test_oracle.py.
"""
from __future__ import annotations

import sys
import types

import networkx as nx
import torch
import torch.nn as nn


class DGLError(Exception):
    pass


class ShimGraph:
    """Edge-list graph with the handful of DGLGraph methods the reference calls."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int):
        self._src = src.to(torch.int64)
        self._dst = dst.to(torch.int64)
        self._n = int(num_nodes)

    # -- DGLGraph surface used by the reference ---------------------------
    def number_of_nodes(self) -> int:
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self) -> int:
        return int(self._src.numel())

    num_edges = number_of_edges

    def to(self, device):
        return self

    def edges(self):
        return self._src, self._dst

    def in_degrees(self) -> torch.Tensor:
        return torch.bincount(self._dst, minlength=self._n)

    def out_degrees(self) -> torch.Tensor:
        return torch.bincount(self._src, minlength=self._n)

    def local_scope(self):  # pragma: no cover - API completeness
        import contextlib
        return contextlib.nullcontext()


def from_networkx(nx_graph, node_attrs=None, edge_attrs=None, **_ignored) -> ShimGraph:
    """dgl.from_networkx restated: sorted relabel, both directions for undirected."""
    if not nx_graph.is_directed():
        nx_graph = nx_graph.to_directed()
    g = nx.convert_node_labels_to_integers(nx_graph, ordering="sorted")
    n = g.number_of_nodes()
    if g.number_of_edges() == 0:
        e = torch.zeros((0, 2), dtype=torch.int64)
    else:
        e = torch.tensor(list(g.edges()), dtype=torch.int64)
    return ShimGraph(e[:, 0], e[:, 1], n)


class GraphConv(nn.Module):
    """dgl.nn.pytorch.GraphConv(in, out) with all defaults (norm='both')."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True,
                 activation=None, allow_zero_in_degree=False):
        super().__init__()
        if norm != "both" or not weight or activation is not None:
            raise NotImplementedError("shim covers the reference's default GraphConv only")
        self._in_feats = in_feats
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.empty(out_feats)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph: ShimGraph, feat: torch.Tensor) -> torch.Tensor:
        in_deg = graph.in_degrees()
        if not self._allow_zero_in_degree and bool((in_deg == 0).any()):
            raise DGLError("There are 0-in-degree nodes in the graph, output for those nodes "
                           "will be invalid.")
        src, dst = graph.edges()
        out_norm = graph.out_degrees().to(feat.dtype).clamp(min=1).pow(-0.5)
        feat_src = feat * out_norm.unsqueeze(1)

        def aggregate(h):
            out = torch.zeros((graph.number_of_nodes(), h.shape[1]), dtype=h.dtype, device=h.device)
            out.index_add_(0, dst, h.index_select(0, src))
            return out

        if self._in_feats > self._out_feats:
            rst = aggregate(torch.matmul(feat_src, self.weight))
        else:
            rst = torch.matmul(aggregate(feat_src), self.weight)
        in_norm = in_deg.to(feat.dtype).clamp(min=1).pow(-0.5)
        rst = rst * in_norm.unsqueeze(1)
        if self.bias is not None:
            rst = rst + self.bias
        return rst


def _make_module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install() -> None:
    """Register `dgl`, `dgl.nn.pytorch` and an inert `matplotlib` in sys.modules."""
    if "dgl" not in sys.modules:
        pytorch = _make_module("dgl.nn.pytorch", GraphConv=GraphConv)
        nn_mod = _make_module("dgl.nn", pytorch=pytorch)
        dgl = _make_module("dgl", from_networkx=from_networkx, DGLError=DGLError,
                           DGLGraph=ShimGraph, nn=nn_mod, __version__="2.0.0-shim")
        sys.modules["dgl"] = dgl
        sys.modules["dgl.nn"] = nn_mod
        sys.modules["dgl.nn.pytorch"] = pytorch
    try:
        import matplotlib  # noqa: F401
    except ModuleNotFoundError:
        class _Inert:
            def __getattr__(self, _name):
                return _Inert()

            def __call__(self, *a, **k):
                return _Inert()

            def __iter__(self):
                return iter(())

        def _module_getattr(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Inert()

        pyplot = _make_module("matplotlib.pyplot")
        pyplot.__getattr__ = _module_getattr  # type: ignore[attr-defined]
        mpl = _make_module("matplotlib", pyplot=pyplot)
        mpl.__getattr__ = _module_getattr  # type: ignore[attr-defined]
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = pyplot
