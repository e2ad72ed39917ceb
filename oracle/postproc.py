"""TEST INFRASTRUCTURE ONLY -- ctypes driver for oracle/postproc.c plus
pure-Python restatements (small cases) of the same reference functions.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use this.
Citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes
import os
import random
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_postproc.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -shared oracle/postproc.c -> oracle/_build/liboracle_postproc.so"""
    src = os.path.join(_HERE, "postproc.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, src])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        _lib.orc_cut_value.restype = ctypes.c_int64
        _lib.orc_cut_value.argtypes = [i32p, i32p, i32p, i32p, ctypes.c_int32]
        _lib.orc_assign_partitions.restype = None
        _lib.orc_assign_partitions.argtypes = [f32p, ctypes.c_int32, ctypes.c_int32, f64p, ctypes.c_int32, i32p]
        _lib.orc_sample_best_cut.restype = ctypes.c_int64
        _lib.orc_sample_best_cut.argtypes = [i32p, i32p, i32p, f32p, ctypes.c_int32, ctypes.c_int32, f64p,
                                             ctypes.c_int32, ctypes.c_int32, i32p, i32p]
        _lib.orc_greedy_node_move.restype = ctypes.c_int64
        _lib.orc_greedy_node_move.argtypes = [i32p, i32p, i32p, i32p, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int32, ctypes.c_int32, i32p, i32p]
        _lib.orc_simple_assignment.restype = None
        _lib.orc_simple_assignment.argtypes = [f32p, ctypes.c_int32, ctypes.c_int32, i32p]
    return _lib


def _p(a: Optional[np.ndarray], ty):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ty))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def numpy_compares_in_f32() -> bool:
    """True when `python_float < np.float32` is evaluated in float32 (NEP 50, numpy >= 2)."""
    return int(np.__version__.split(".")[0]) >= 2


# ---------------------------------------------------------------- C-backed
def cut_value(rowptr, colidx, labels, wts=None) -> int:
    rowptr, colidx, labels = _i32(rowptr), _i32(colidx), _i32(labels)
    w = None if wts is None else _i32(wts)
    return int(lib().orc_cut_value(_p(rowptr, ctypes.c_int32), _p(colidx, ctypes.c_int32),
                                   _p(w, ctypes.c_int32), _p(labels, ctypes.c_int32), len(labels)))


def assign_partitions(P: np.ndarray, U: np.ndarray, compare_f32: Optional[bool] = None) -> np.ndarray:
    P = np.ascontiguousarray(P, dtype=np.float32)
    U = np.ascontiguousarray(U, dtype=np.float64)
    n, K = P.shape
    out = np.empty(n, dtype=np.int32)
    cf = numpy_compares_in_f32() if compare_f32 is None else compare_f32
    lib().orc_assign_partitions(_p(P, ctypes.c_float), n, K, _p(U, ctypes.c_double), int(cf),
                                _p(out, ctypes.c_int32))
    return out


def sample_best_cut(rowptr, colidx, P, U, iters: int, wts=None,
                    compare_f32: Optional[bool] = None) -> Tuple[np.ndarray, int, int]:
    rowptr, colidx = _i32(rowptr), _i32(colidx)
    w = None if wts is None else _i32(wts)
    P = np.ascontiguousarray(P, dtype=np.float32)
    U = np.ascontiguousarray(U, dtype=np.float64)
    n, K = P.shape
    out = np.zeros(n, dtype=np.int32)
    best_it = ctypes.c_int32(-1)
    cf = numpy_compares_in_f32() if compare_f32 is None else compare_f32
    best = lib().orc_sample_best_cut(_p(rowptr, ctypes.c_int32), _p(colidx, ctypes.c_int32),
                                     _p(w, ctypes.c_int32), _p(P, ctypes.c_float), n, K,
                                     _p(U, ctypes.c_double), iters, int(cf),
                                     _p(out, ctypes.c_int32), ctypes.byref(best_it))
    return out, int(best), int(best_it.value)


def greedy_node_move(rowptr, colidx, labels, K: int = 3, iters: int = 200, n_frozen: int = 3,
                     wts=None) -> Tuple[np.ndarray, int, int]:
    rowptr, colidx, labels = _i32(rowptr), _i32(colidx), _i32(labels)
    w = None if wts is None else _i32(wts)
    out = np.empty_like(labels)
    moves = ctypes.c_int32(0)
    cut = lib().orc_greedy_node_move(_p(rowptr, ctypes.c_int32), _p(colidx, ctypes.c_int32),
                                     _p(w, ctypes.c_int32), _p(labels, ctypes.c_int32), len(labels), K,
                                     iters, n_frozen, _p(out, ctypes.c_int32), ctypes.byref(moves))
    return out, int(cut), int(moves.value)


def simple_assignment(P: np.ndarray) -> np.ndarray:
    P = np.ascontiguousarray(P, dtype=np.float32)
    out = np.empty(P.shape[0], dtype=np.int32)
    lib().orc_simple_assignment(_p(P, ctypes.c_float), P.shape[0], P.shape[1], _p(out, ctypes.c_int32))
    return out


# ---------------------------------------------------------------- pure Python
def py_cut_value(labels: Sequence[int], nx_graph) -> int:
    """TestingNeuralNetwork.py:48-64 on the networkx edge list."""
    total = 0
    m = len(labels)
    for u, v, data in nx_graph.edges(data=True):
        if u < m and v < m and labels[u] != labels[v]:
            total += data.get("weight", 1)
    return total


def py_assign_partitions(P: np.ndarray, draw=np.random.rand) -> List[int]:
    """TestingNeuralNetwork.py:18-46, numpy scalar semantics preserved by
    iterating the float32 rows exactly as the reference does."""
    out = [0, 1, 2]
    for row in P[3:]:
        r = draw()
        acc = 0
        chosen = len(row) - 1
        for i, p in enumerate(row):
            acc += p
            if r < acc:
                chosen = i
                break
        out.append(chosen)
    return out


def py_greedy_2way(labels: Sequence[int], nx_graph, num_steps: int) -> Tuple[int, List[int]]:
    """huerestics_multi-max.ipynb:L5817-5854 (cell 48) restated with full
    re-evaluation of the objective per candidate, exactly as the notebook does
    (O(n*|E|) per iteration; small graphs only)."""
    cur = list(labels)
    n = len(cur)

    def obj(sol):
        return sum(1 for u, v in nx_graph.edges() if sol[u] != sol[v])

    cur_score = obj(cur)
    for it in range(n):
        if it >= num_steps:
            break
        scores = []
        for v in range(n):
            cand = list(cur)
            cand[v] = (cand[v] + 1) % 2
            scores.append(obj(cand))
        best = max(scores)
        idx = scores.index(best)
        if best > cur_score:
            cur_score = best
            cur[idx] = (cur[idx] + 1) % 2
        else:
            break
    return cur_score, cur


def py_randomized_k_way_maxcut(nx_graph, k: int = 3, max_iterations: int = 1000, threshold: int = 0,
                               patience: int = 10, fixed_terminals: Optional[Dict[int, int]] = None,
                               random_seed: Optional[int] = None) -> Tuple[int, Optional[Dict[int, int]]]:
    """python/RandomAlgorithm/RandomizedMaxCut.py:63-122: uniform random labels per
    iteration from Python's `random` stream, keep when cut > best + threshold,
    stop after `patience` non-improving iterations.  This is the generator of the
    six seeded known answers (randomizedAlgo.ipynb:L113,L186,L275,L486-497)."""
    if random_seed is not None:
        random.seed(random_seed)
    best_cut, best_part, stale = 0, None, 0
    nodes = list(nx_graph.nodes())
    for _ in range(max_iterations):
        if fixed_terminals:
            part = dict(fixed_terminals)
            free = [u for u in nodes if u not in fixed_terminals]
        else:
            part, free = {}, nodes
        for u in free:
            part[u] = random.randint(0, k - 1)
        cut = 0
        for u, v, data in nx_graph.edges(data=True):
            if part[u] != part[v]:
                cut += data.get("weight", 1)
        if cut > best_cut + threshold:
            best_cut, best_part, stale = cut, dict(part), 0
        else:
            stale += 1
        if stale >= patience:
            break
    return best_cut, best_part
