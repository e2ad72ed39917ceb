"""Graph handles of the B200 path.

`CSRGraph`   -- host-side, one graph; occupies the DGLGraph slot of the reference's
                dataset tuple `[graph, X, nx_graph, [0,1,2]]` (graphExtender.py:114) and
                duck-types the DGLGraph methods the reference and its notebooks call
                (`number_of_nodes`, `number_of_edges`, `.to(device)`; TrainingNeural.py:271,
                complete_training_pipeline.ipynb cell 10).
`GraphBatch` -- device-side block-diagonal batch of many graphs: int32 CSR, the values
                of A_hat = D^-1/2 A D^-1/2 per edge, float/int edge weights, graph_ptr.

Layout follows dgl.from_networkx (graphExtender.py:102-103): nodes relabelled in sorted
order, undirected edges stored in both directions (nnz = 2|E|).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence

import numpy as np


class CSRGraph:
    """In-edge CSR of an undirected graph (both directions stored), host arrays."""

    __slots__ = ("rowptr", "colidx", "weights", "n", "_device", "_batch", "_features_verified", "_nx_edges")

    def __init__(self, rowptr: np.ndarray, colidx: np.ndarray, weights: Optional[np.ndarray], n: int):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        if weights is None:
            weights = np.ones(self.colidx.shape[0], dtype=np.float32)
        self.weights = np.ascontiguousarray(weights, dtype=np.float32)
        self.n = int(n)
        self._device = "cpu"
        self._batch = None           # lazily built one-graph GraphBatch (device), never pickled
        self._features_verified = None   # check_adjacency_features memo (weakref to the tensor, version, width), never pickled
        self._nx_edges = None            # (id of the networkx graph, its undirected edge count), never pickled
        if self.rowptr.shape[0] != self.n + 1 or self.rowptr[-1] != self.colidx.shape[0]:
            raise ValueError("inconsistent CSR arrays")

    def __getstate__(self):
        return {"rowptr": self.rowptr, "colidx": self.colidx, "weights": self.weights, "n": self.n}

    def __setstate__(self, state):
        self.rowptr, self.colidx, self.weights, self.n = state["rowptr"], state["colidx"], state["weights"], state["n"]
        self._device = "cpu"
        self._batch = None
        self._features_verified = None
        self._nx_edges = None

    # ---- DGLGraph surface -------------------------------------------------------------
    def number_of_nodes(self) -> int:
        return self.n

    num_nodes = number_of_nodes

    def number_of_edges(self) -> int:
        return int(self.colidx.shape[0])

    num_edges = number_of_edges

    def to(self, device):
        """DGL moves the structure; ours is uploaded lazily per batch, so this only records intent."""
        self._device = str(device)
        return self

    @property
    def device(self):
        return self._device

    def in_degrees(self):
        import torch
        return torch.from_numpy(np.diff(self.rowptr).astype(np.int64))

    out_degrees = in_degrees          # symmetric by construction

    def degrees(self) -> np.ndarray:
        return np.diff(self.rowptr)

    def __repr__(self):
        return f"CSRGraph(num_nodes={self.n}, num_edges={self.number_of_edges()})"

    # ---- constructors -----------------------------------------------------------------
    @staticmethod
    def from_edges(n: int, edges: np.ndarray, weights: Optional[np.ndarray] = None) -> "CSRGraph":
        """`edges` [m,2] undirected pairs (each listed once); vectorised symmetrise + sort."""
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        m = edges.shape[0]
        w = np.ones(m, dtype=np.float32) if weights is None else np.asarray(weights, dtype=np.float32)
        u, v = edges[:, 0], edges[:, 1]
        loops = u == v
        rows = np.concatenate([v, u[~loops]])
        cols = np.concatenate([u, v[~loops]])
        ww = np.concatenate([w, w[~loops]])
        order = np.lexsort((cols, rows))
        rows, cols, ww = rows[order], cols[order], ww[order]
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        rowptr = np.cumsum(rowptr)
        return CSRGraph(rowptr.astype(np.int32), cols.astype(np.int32), ww, n)

    @staticmethod
    def from_networkx(nx_graph) -> "CSRGraph":
        """dgl.from_networkx equivalent (sorted relabel, both directions), edge attr 'weight' kept."""
        nodes = sorted(nx_graph.nodes())
        n = len(nodes)
        identity = n == 0 or (nodes[0] == 0 and nodes[-1] == n - 1)
        index = None if identity else {u: i for i, u in enumerate(nodes)}
        m = nx_graph.number_of_edges()
        e = np.empty((m, 2), dtype=np.int64)
        w = np.empty(m, dtype=np.float32)
        for i, (a, b, data) in enumerate(nx_graph.edges(data=True)):
            e[i, 0] = a if identity else index[a]
            e[i, 1] = b if identity else index[b]
            w[i] = data.get("weight", 1)
        return CSRGraph.from_edges(n, e, w)


class AdjacencyFeatures:
    """Stand-in for element [1] of a dataset item (`[graph, X, nx_graph, terminals]`, graphExtender.py:114) whose
    features are the zero-padded adjacency rows of its graph -- what graphExtender always writes -- without holding
    the dense [n, max_nodes] matrix (4 MB per graph at max_nodes = 1000).  `process_graphs_from_folder(...,
    dense_features=False)` produces it; `.dense()` materialises the reference's tensor on demand."""
    __slots__ = ("graph", "n_cols")

    def __init__(self, graph: CSRGraph, n_cols: int):
        self.graph, self.n_cols = graph, int(n_cols)

    @property
    def shape(self):
        return (self.graph.n, self.n_cols)

    def dense(self):
        import torch
        g = self.graph
        out = torch.zeros(g.n, self.n_cols, dtype=torch.float32)
        rows = np.repeat(np.arange(g.n), np.diff(g.rowptr))
        out[torch.from_numpy(rows), torch.from_numpy(g.colidx.astype(np.int64))] = torch.from_numpy(g.weights)
        return out


NOT_ADJACENCY = ("dataset features are not the zero-padded adjacency rows of the graph; "
                  "the fused max-cut loss takes its edge weights from the graph structure")


def check_adjacency_features(handle: CSRGraph, X) -> int:
    """The reference's loss reads its weights from `adjacency_matrix` (= the features, :380); the fused kernel reads
    them from the graph, which is the same thing for every dataset graphExtender produces.  Verified per item where the
    item lives (host tensors on the host), against the graph's own nnz entries: every edge position holds its weight
    and nothing else is non-zero.  Anything else is refused rather than silently diverging.  Returns the feature width."""
    if isinstance(X, AdjacencyFeatures):
        if X.graph is not handle and (X.graph.n != handle.n or not np.array_equal(X.graph.colidx, handle.colidx)
                                      or not np.array_equal(X.graph.rowptr, handle.rowptr)):
            raise NotImplementedError(NOT_ADJACENCY)
        return X.n_cols
    import torch
    if not torch.is_tensor(X) or X.dim() != 2 or X.shape[0] != handle.n:
        raise NotImplementedError(NOT_ADJACENCY)
    # verified before and untouched since (same tensor object, same autograd version): a test harness revisits the same
    # dataset items every pass, and the check is 20 % of a config-2 pass
    memo = getattr(handle, "_features_verified", None)
    if memo is not None and memo[0]() is X and memo[1] == X._version:
        return memo[2]
    nnz = handle.number_of_edges()
    if nnz and int(handle.colidx.max()) >= X.shape[1]:
        raise NotImplementedError(NOT_ADJACENCY)
    rows = torch.from_numpy(np.repeat(np.arange(handle.n), np.diff(handle.rowptr))).to(X.device)
    cols = torch.from_numpy(handle.colidx.astype(np.int64)).to(X.device)
    want = torch.from_numpy(handle.weights).to(device=X.device, dtype=X.dtype)
    # duplicate-free CSR: nnz matching entries + exactly count_nonzero(weights) non-zeros overall == equality
    if not torch.equal(X[rows, cols], want) or int(torch.count_nonzero(X)) != int(np.count_nonzero(handle.weights)):
        raise NotImplementedError(NOT_ADJACENCY)
    try:
        import weakref
        handle._features_verified = (weakref.ref(X), X._version, int(X.shape[1]))
    except (AttributeError, TypeError):                          # a handle with __slots__ / an object that cannot be weak-referenced
        pass
    return int(X.shape[1])


def from_networkx(nx_graph, **_ignored) -> CSRGraph:
    """Module-level spelling so that `dgl.from_networkx(nx_graph=...)` call sites keep working."""
    return CSRGraph.from_networkx(nx_graph)


class ZeroDegreeError(RuntimeError):
    """Mirror of DGLError('There are 0-in-degree nodes in the graph ...') raised by GraphConv."""


class GraphBatch:
    """Block-diagonal device batch.  All arrays are torch CUDA tensors:

        rowptr  int32 [N+1]    colidx int32 [nnz]     graph_ptr int32 [B+1]
        coef    f32 [nnz]  = w? * deg(u)^-1/2 * deg(v)^-1/2   (GraphConv ignores edge weights:
                              the reference calls conv(g, feat) without edge_weight, so w == 1 here)
        wts_f32 f32 [nnz] or None   edge weights for the loss (adjacency values, commons.py:65-77)
        wts_i32 i32 [nnz] or None   the same weights for the integer cut evaluator
    """

    def __init__(self, graphs: Sequence[CSRGraph], device=None, check_degrees: bool = True):
        import torch
        from . import _lib, ops
        self.device = device if device is not None else _lib.require_cuda()
        graphs = list(graphs)
        self.num_graphs = len(graphs)
        sizes = np.asarray([g.n for g in graphs], dtype=np.int64)
        nnzs = np.asarray([g.number_of_edges() for g in graphs], dtype=np.int64)
        self.sizes = sizes
        gp = np.zeros(self.num_graphs + 1, dtype=np.int64)
        np.cumsum(sizes, out=gp[1:])
        ep = np.zeros(self.num_graphs + 1, dtype=np.int64)
        np.cumsum(nnzs, out=ep[1:])
        self.num_nodes = int(gp[-1])
        self.nnz = int(ep[-1])
        if self.num_nodes >= 2**31 - 1 or self.nnz >= 2**31 - 1:
            raise ValueError("batch too large for int32 CSR indices; shard the graphs")
        rowptr = np.empty(self.num_nodes + 1, dtype=np.int32)
        colidx = np.empty(self.nnz, dtype=np.int32)
        weights = np.empty(self.nnz, dtype=np.float32)
        rowptr[0] = 0
        for i, g in enumerate(graphs):
            rowptr[gp[i] + 1: gp[i + 1] + 1] = g.rowptr[1:] + ep[i]
            colidx[ep[i]: ep[i + 1]] = g.colidx + gp[i]
            weights[ep[i]: ep[i + 1]] = g.weights
        self._finish(rowptr, colidx, weights, gp.astype(np.int32), check_degrees)

    @classmethod
    def from_arrays(cls, rowptr, colidx, graph_ptr, weights=None, device=None, check_degrees=True) -> "GraphBatch":
        """Adopt already block-diagonal host arrays (synthetic generators, config 3-5)."""
        from . import _lib
        self = cls.__new__(cls)
        self.device = device if device is not None else _lib.require_cuda()
        graph_ptr = np.asarray(graph_ptr, dtype=np.int64)
        self.num_graphs = len(graph_ptr) - 1
        self.sizes = np.diff(graph_ptr)
        self.num_nodes = int(graph_ptr[-1])
        self.nnz = int(len(colidx))
        w = np.ones(self.nnz, dtype=np.float32) if weights is None else np.asarray(weights, dtype=np.float32)
        self._finish(np.asarray(rowptr, dtype=np.int32), np.asarray(colidx, dtype=np.int32), w,
                     graph_ptr.astype(np.int32), check_degrees)
        return self

    @classmethod
    def from_device_arrays(cls, rowptr, colidx, graph_ptr, sizes, max_nodes: Optional[int] = None,
                           norm=None, coef=None) -> "GraphBatch":
        """Adopt block-diagonal CSR arrays that already live on the device (a streamed batch: the arrays were copied
        from pinned host memory on a copy stream).  Unit edge weights; degrees must have been validated by the caller
        (no zero-degree rows) -- nothing here synchronises with the host.  `norm` / `coef` are optional preallocated
        outputs.  No ELL plan is built (its construction reads a flag back)."""
        from . import ops
        self = cls.__new__(cls)
        self.device = rowptr.device
        self.sizes = np.asarray(sizes, dtype=np.int64)
        self.num_graphs = int(self.sizes.shape[0])
        self.num_nodes = int(self.sizes.sum())
        self.nnz = int(colidx.numel())
        self.rowptr, self.colidx, self.graph_ptr = rowptr, colidx, graph_ptr
        self.unit_weights, self.wts_f32, self.wts_i32, self.integer_weights = True, None, None, True
        self.norm, _ = ops.degree_norm(rowptr, self.num_nodes, count_zero=False, out=norm)
        self.coef = ops.edge_coef(rowptr, colidx, None, self.norm, self.norm, self.num_nodes, out=coef)
        self.max_nodes = int(max_nodes if max_nodes is not None else (self.sizes.max() if self.num_graphs else 0))
        self.plan = None
        return self

    def _finish(self, rowptr, colidx, weights, graph_ptr, check_degrees):
        import torch
        from . import ops
        dev = self.device
        self.rowptr = torch.from_numpy(rowptr).to(dev)
        self.colidx = torch.from_numpy(colidx).to(dev)
        self.graph_ptr = torch.from_numpy(graph_ptr).to(dev)
        unit = bool(np.all(weights == 1.0))
        self.unit_weights = unit
        self.wts_f32 = None if unit else torch.from_numpy(weights).to(dev)
        self.integer_weights = True
        if unit:
            self.wts_i32 = None
        else:
            wi = np.rint(weights).astype(np.int32)
            self.wts_i32 = torch.from_numpy(wi).to(dev) if np.array_equal(wi.astype(np.float32), weights) else None
            self.integer_weights = self.wts_i32 is not None
        self.norm, zero = ops.degree_norm(self.rowptr, self.num_nodes)
        if check_degrees and zero > 0:
            raise ZeroDegreeError(
                "There are 0-in-degree nodes in the graph, output for those nodes will be invalid. "
                f"({zero} nodes; DGL GraphConv raises the same with allow_zero_in_degree=False)")
        self.coef = ops.edge_coef(self.rowptr, self.colidx, None, self.norm, self.norm, self.num_nodes)
        self.max_nodes = int(self.sizes.max()) if self.num_graphs else 0
        # ELL plan for the shared-memory slab SpMM (TMA-staged, 0.60 of the HBM roofline at config 3 vs 0.50-0.54
        # for the L2-bound warp-per-row kernel, profiles/r01_spmm_slab_notes.md).  One CTA owns a (graph, slab)
        # item, so it only pays off when the batch has enough graphs to fill the GPU: built automatically for
        # batches of >= 32 graphs; GMC_SPMM_SLAB=0 disables it, =1 forces it for any batch.
        self.plan = None
        mode = os.environ.get("GMC_SPMM_SLAB", "auto")
        if mode == "1" or (mode != "0" and self.num_graphs >= 32):
            self.build_plan()

    def build_plan(self) -> bool:
        """Build the ELL plan that lets ops.spmm take the shared-memory slab kernel (needs max degree <= 8 and,
        per node, neighbours of equal degree -- regular graphs)."""
        from . import ops
        if self.max_nodes >= 128 and self.num_nodes > 0:
            self.plan = ops.spmm_plan(self.rowptr, self.colidx, self.norm, self.norm, self.graph_ptr, self.num_graphs,
                                      self.num_nodes)
        return self.plan is not None

    # per-graph views --------------------------------------------------------------------
    def node_slice(self, g: int) -> slice:
        gp = self.graph_ptr_host
        return slice(int(gp[g]), int(gp[g + 1]))

    @property
    def graph_ptr_host(self) -> np.ndarray:
        if not hasattr(self, "_gp_host"):
            gp = np.zeros(self.num_graphs + 1, dtype=np.int64)
            np.cumsum(self.sizes, out=gp[1:])
            self._gp_host = gp
        return self._gp_host
