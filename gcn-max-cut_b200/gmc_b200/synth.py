"""Synthetic inputs for the BASELINE.json configurations 3-5: many random d-regular simple graphs
as ONE block-diagonal CSR, generated vectorised with numpy (networkx's random_regular_graph would
take minutes for 4 096 x n=1000 and cannot do n = 10^6 in reasonable time).

Construction: a d-regular simple graph on an even number of nodes as the union of d random perfect
matchings; a matching that repeats an edge of the earlier ones is redrawn (rejection, vectorised
over all graphs of the batch).  Not exactly uniform over d-regular graphs, but every output is a
simple d-regular graph, which is all the benchmark shape requires.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def _matchings(rng: np.random.Generator, B: int, n: int) -> np.ndarray:
    """[B, n/2, 2] random perfect matchings, endpoints ordered (lo, hi)."""
    perm = np.argsort(rng.random((B, n), dtype=np.float32), axis=1).astype(np.int64)
    a, b = perm[:, 0::2], perm[:, 1::2]
    return np.stack([np.minimum(a, b), np.maximum(a, b)], axis=2)


def regular_graph_edges(B: int, n: int, d: int, seed: int = 0, max_rounds: int = 2000) -> np.ndarray:
    """Edges [B, n*d/2, 2] (local node ids, lo < hi) of B independent simple d-regular graphs."""
    if n % 2:
        raise ValueError("the matching construction needs an even number of nodes")
    if d < 1 or d >= n:
        raise ValueError("need 1 <= d < n")
    rng = np.random.default_rng(seed)
    half = n // 2
    keys = np.empty((B, d * half), dtype=np.int64)          # edge keys lo*n+hi, matching k in [k*half,(k+1)*half)
    for k in range(d):
        todo = np.arange(B)
        for _ in range(max_rounds):
            m = _matchings(rng, len(todo), n)
            cand = m[:, :, 0] * n + m[:, :, 1]
            if k == 0:
                ok = np.ones(len(todo), dtype=bool)
            else:
                # one global searchsorted: shift every graph's keys into its own disjoint range
                shift = (np.arange(len(todo), dtype=np.int64) * (n * n))[:, None]
                prev = (np.sort(keys[todo, : k * half], axis=1) + shift).ravel()
                probe = (cand + shift).ravel()
                pos = np.minimum(np.searchsorted(prev, probe), prev.shape[0] - 1)
                ok = ~(prev[pos] == probe).reshape(cand.shape).any(axis=1)
            keys[todo[ok], k * half: (k + 1) * half] = cand[ok]
            todo = todo[~ok]
            if len(todo) == 0:
                break
        else:  # pragma: no cover
            raise RuntimeError("regular graph generation did not converge")
    return np.stack([keys // n, keys % n], axis=2)


def block_diagonal_csr(edges: np.ndarray, n: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """edges [B, m, 2] local ids -> (rowptr int32 [B*n+1], colidx int32 [2*B*m], graph_ptr int32 [B+1]) with both
    directions stored and neighbours sorted within each row."""
    B, m, _ = edges.shape
    base = (np.arange(B, dtype=np.int64) * n)[:, None]
    u = (edges[:, :, 0] + base).ravel()
    v = (edges[:, :, 1] + base).ravel()
    rows = np.concatenate([u, v])
    cols = np.concatenate([v, u])
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    counts = np.bincount(rows, minlength=B * n)
    rowptr = np.zeros(B * n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    graph_ptr = (np.arange(B + 1, dtype=np.int64) * n)
    return rowptr.astype(np.int32), cols.astype(np.int32), graph_ptr.astype(np.int32)


def regular_batch_arrays(B: int, n: int, d, seed: int = 0):
    """Block-diagonal CSR of B random regular graphs.  `d` is an int or a sequence of per-graph
    degrees (config 4: d_g = 6 + (g mod 3)); graphs keep their order."""
    if np.isscalar(d):
        return block_diagonal_csr(regular_graph_edges(B, n, int(d), seed), n)
    degs = np.asarray(d, dtype=np.int64)
    if degs.shape[0] != B:
        raise ValueError("need one degree per graph")
    rowptrs, cols = [], []
    # generate per degree class, then interleave back into graph order
    per_graph = {}
    for dv in np.unique(degs):
        idx = np.nonzero(degs == dv)[0]
        e = regular_graph_edges(len(idx), n, int(dv), seed + 7919 * int(dv))
        for j, g in enumerate(idx):
            per_graph[int(g)] = e[j]
    nnz_off = 0
    rowptr = np.zeros(B * n + 1, dtype=np.int64)
    col_chunks = []
    for g in range(B):
        rp, ci, _ = block_diagonal_csr(per_graph[g][None], n)
        rowptr[g * n + 1: (g + 1) * n + 1] = rp[1:].astype(np.int64) + nnz_off
        col_chunks.append(ci.astype(np.int64) + g * n)
        nnz_off += len(ci)
    graph_ptr = np.arange(B + 1, dtype=np.int64) * n
    return rowptr.astype(np.int32), np.concatenate(col_chunks).astype(np.int32), graph_ptr.astype(np.int32)


def regular_batch(B: int, n: int, d, seed: int = 0, device=None):
    """GraphBatch of B random regular graphs on the current CUDA device."""
    from .graph import GraphBatch
    rowptr, colidx, graph_ptr = regular_batch_arrays(B, n, d, seed)
    return GraphBatch.from_arrays(rowptr, colidx, graph_ptr, device=device)


class RegularGraphDataset(dict):
    """A graphExtender-format dataset `{key: [graph, X, nx.Graph, [0, 1, 2]]}` (graphExtender.py:114) over the graphs of
    a block-diagonal CSR (`regular_batch_arrays`), materialised item by item on first access.

    Keys are `first_key .. first_key + B - 1` of a dataset with `total` keys: under data parallelism every rank holds
    the dataset with only ITS shard backed by arrays (train_single_epoch dereferences nothing else); asking for another
    rank's item raises KeyError.  X is the AdjacencyFeatures stand-in (the dense [n, width] tensor is 4 MB per graph)."""

    def __init__(self, rowptr, colidx, graph_ptr, width: int, first_key: int = 0, total: Optional[int] = None,
                 with_networkx: bool = True):
        n_local = len(graph_ptr) - 1
        total = n_local + first_key if total is None else int(total)
        super().__init__((k, None) for k in range(total))
        self._arrays = (np.asarray(rowptr), np.asarray(colidx), np.asarray(graph_ptr))
        self._first, self._n_local, self._width, self._nx = int(first_key), n_local, int(width), bool(with_networkx)

    def _build(self, key: int):
        from .graph import AdjacencyFeatures, CSRGraph
        g = key - self._first
        if not 0 <= g < self._n_local:
            raise KeyError(f"graph {key} belongs to another rank's shard")
        rowptr, colidx, gp = self._arrays
        lo, hi = int(gp[g]), int(gp[g + 1])
        rp = (rowptr[lo: hi + 1] - rowptr[lo]).astype(np.int32)
        ci = (colidx[rowptr[lo]: rowptr[hi]] - lo).astype(np.int32)
        handle = CSRGraph(rp, ci, None, hi - lo)
        nx_graph = None
        if self._nx:
            import networkx as nx
            rows = np.repeat(np.arange(hi - lo), np.diff(rp))
            keep = rows < ci
            nx_graph = nx.Graph()
            nx_graph.add_nodes_from(range(hi - lo))
            nx_graph.add_edges_from(zip(rows[keep].tolist(), ci[keep].tolist()), weight=1, capacity=1)
        return [handle, AdjacencyFeatures(handle, self._width), nx_graph, [0, 1, 2]]

    def __getitem__(self, key):
        item = super().__getitem__(key)
        if item is None:
            item = self._build(key)
            super().__setitem__(key, item)
        return item

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]
