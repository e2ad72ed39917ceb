"""RandomizedMaxCut -- the reference's randomized k-way max-cut baseline
(python/RandomAlgorithm/RandomizedMaxCut.py:23-122) with the cut evaluation on the GPU
(SURVEY.md 8(f) rank 3).  Same API, same results for a given seed -- including the six seeded
known answers recorded in randomizedAlgo.ipynb (1393 / 2741 / 2738 / 2724 / 2719 / 2742) -- and
Python's global `random` generator is left in exactly the state the reference leaves it in.

How: `random.randint(0, k-1)` is `getrandbits(k.bit_length())` with rejection, i.e. the top bits of
successive MT19937 words.  The same words are produced in bulk by numpy's MT19937 seeded with
Python's generator state, turned into labels vectorised, evaluated `chunk` iterations at a time by
`gmc_cut_value_multi_u8`, and the sequential patience rule is replayed on the integer cut values.
Afterwards Python's generator is advanced to the first unconsumed word.
"""
import random
import time
from typing import Dict, List, Optional, Tuple

import networkx as nx
import numpy as np

_CHUNK = 128


def create_random_regular_graph(n: int, degree: int = 8, random_seed: Optional[int] = None) -> nx.Graph:
    """Random `degree`-regular graph with weight=1 on every edge (reference :23-45)."""
    G = nx.random_regular_graph(d=degree, n=n, seed=random_seed) if random_seed is not None \
        else nx.random_regular_graph(d=degree, n=n)
    nx.set_edge_attributes(G, 1, "weight")
    return G


def calculate_cut_value(graph, partition: Dict[int, int]) -> int:
    """Total weight of edges whose end points lie in different partitions (reference :48-60)."""
    total = 0
    for u, v, data in graph.edges(data=True):
        if partition[u] != partition[v]:
            total += data.get("weight", 1)
    return total


# ---------------------------------------------------------------------------- MT19937 word stream
class _WordStream:
    """Bulk view of the 32-bit words Python's `random` module would produce next."""

    def __init__(self):
        version, internal, gauss = random.getstate()
        self._version, self._gauss = version, gauss
        self._rs = np.random.RandomState()
        self._rs.set_state(("MT19937", np.asarray(internal[:-1], dtype=np.uint32), int(internal[-1])))
        self._start = self._rs.get_state()
        self._buf = np.zeros(0, dtype=np.uint32)
        self.taken = 0                       # words handed out so far

    def words(self, count: int) -> np.ndarray:
        """Return (without consuming) the next `count` words after the `taken` ones."""
        need = self.taken + count - len(self._buf)
        if need > 0:
            extra = np.frombuffer(self._rs.bytes(4 * need), dtype="<u4")
            self._buf = np.concatenate([self._buf, extra])
        return self._buf[self.taken: self.taken + count]

    def consume(self, count: int) -> None:
        self.taken += count

    def commit(self) -> None:
        """Advance Python's generator past exactly the consumed words."""
        rs = np.random.RandomState()
        rs.set_state(self._start)
        if self.taken:
            rs.bytes(4 * self.taken)
        _, key, pos = rs.get_state()[:3]
        random.setstate((self._version, tuple(int(x) for x in key) + (int(pos),), self._gauss))


def _labels_from_stream(stream: _WordStream, k: int, count: int) -> np.ndarray:
    """`count` values of random.randint(0, k-1), consuming exactly the words Python would."""
    bits = k.bit_length()                    # _randbelow_with_getrandbits(n=k)
    out = np.empty(count, dtype=np.uint8)
    filled = 0
    while filled < count:
        want = count - filled
        block = max(256, int(want * (1 << bits) / k * 1.1) + 16)
        w = stream.words(block)
        vals = (w >> np.uint32(32 - bits)).astype(np.int64)
        ok = np.nonzero(vals < k)[0]
        if len(ok) >= want:
            last = ok[want - 1]
            out[filled:] = vals[ok[:want]]
            stream.consume(int(last) + 1)
            filled = count
        else:
            out[filled: filled + len(ok)] = vals[ok]
            filled += len(ok)
            stream.consume(block)
    return out


def randomized_k_way_maxcut(graph, k: int = 3, max_iterations: int = 1000, threshold: int = 0, patience: int = 10,
                            fixed_terminals: Optional[Dict[int, int]] = None,
                            random_seed: Optional[int] = None) -> Tuple[int, Dict[int, int]]:
    """Randomized k-way max-cut with early stopping (reference :63-122): uniform random labels per
    iteration, keep when cut > best + threshold, stop after `patience` non-improving iterations."""
    import torch
    from gmc_b200 import _lib, ops
    from gmc_b200.graph import CSRGraph, GraphBatch

    if random_seed is not None:
        random.seed(random_seed)
    if k < 1 or k > 255:
        raise ValueError("k must be in 1..255")
    dev = _lib.require_cuda()
    nodes = list(graph.nodes())
    order = sorted(nodes)
    index = {u: i for i, u in enumerate(order)}                  # CSR position (from_networkx sorts labels)
    fixed = dict(fixed_terminals) if fixed_terminals else {}
    free = [u for u in nodes if u not in fixed] if fixed else nodes
    free_pos = np.asarray([index[u] for u in free], dtype=np.int64)
    n = len(nodes)
    base = np.zeros(n, dtype=np.uint8)
    for u, p in fixed.items():
        if u in index:
            base[index[u]] = p
    batch = GraphBatch([CSRGraph.from_networkx(graph)], device=dev, check_degrees=False)

    stream = _WordStream()
    best_cut, best_labels, stale = 0, None, 0
    done = 0
    stop = False
    while done < max_iterations and not stop:
        chunk = min(_CHUNK, max_iterations - done)
        # labels for `chunk` iterations; remember how many words each iteration consumed
        labels = np.tile(base, (chunk, 1))
        marks = []
        for it in range(chunk):
            labels[it, free_pos] = _labels_from_stream(stream, k, len(free))
            marks.append(stream.taken)
        cuts = ops.cut_value_multi(batch, torch.from_numpy(labels).to(dev)).cpu().numpy()
        for it in range(chunk):
            c = int(cuts[it])
            if c > best_cut + threshold:
                best_cut, best_labels, stale = c, labels[it].copy(), 0
            else:
                stale += 1
            if stale >= patience:
                stream.taken = marks[it]                         # un-consume the iterations never run
                stop = True
                break
        done += chunk
    stream.commit()
    best_partition = None
    if best_labels is not None:
        # reference builds the dict as fixed terminals first, then the remaining nodes in graph order
        best_partition = dict(fixed)
        for u in free:
            best_partition[u] = int(best_labels[index[u]])
    return best_cut, best_partition


def evaluate_algorithm_on_graphs(graphs: List, k: int = 3, max_iterations: int = 1000, threshold: int = 0,
                                 patience: int = 10, fixed_terminals: Optional[Dict[int, int]] = None) -> Dict:
    """Mean cut / total time over a list of graphs (reference :125-160)."""
    cut_values = []
    start = time.time()
    for graph in graphs:
        best, _ = randomized_k_way_maxcut(graph, k, max_iterations, threshold, patience, fixed_terminals)
        cut_values.append(best)
    return {"mean_cut_value": np.mean(cut_values), "total_time": time.time() - start, "cut_values": cut_values}
