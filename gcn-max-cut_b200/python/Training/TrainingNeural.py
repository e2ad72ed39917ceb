"""TrainingNeural -- B200 drop-in for the reference's python/Training/TrainingNeural.py.

Same public surface (TrainingConfig, GCNSoftmax, setup_model_and_optimizer, train_single_epoch,
train_model, train_from_pickle, train_multi_class, evaluate_model, load_neural_model,
save_neural_model, the loss helpers and the legacy aliases), same prints, same checkpoint
layout.  The arithmetic of the hot loop (reference :371-388) runs in gmc_b200.engine.GCNEngine
on hand-written sm_100a kernels; there is no CPU fallback -- without a CUDA device every entry
point that would compute raises gmc_b200._lib.GmcError.

Opt-in extensions (defaults reproduce the reference exactly):
    TrainingConfig.batch_graphs      graphs per optimiser step (1 = the reference's sequential
                                     per-graph Adam steps; >1 = block-diagonal mini-batches with
                                     loss = sum of per-graph losses)
    TrainingConfig.gemm_precision    'fp32' (FFMA, parity) | 'tf32' | 'tf32x3' (tcgen05) | 'bf16x3' / 'bf16x2' (fp32-grade on
                                     bf16 tensor cores: exact integer features x split weights, csrc/split.cu) | 'bf16'
    TrainingConfig.stream_dataset    upload every step's graphs from pinned host memory (the reference moves each graph
                                     to the device inside its loop, :371-373) instead of caching device batches
    TrainingConfig.eval_batch_graphs graphs per block-diagonal forward pass in evaluate_model
    TrainingConfig.loss_mode         'ste' (reference live path) | 'soft' (north-star objective)
    TrainingConfig.use_terminal_penalty   enable the penalty the reference left commented (:308)
    TrainingConfig.adjacency_kernels  with feature_source='adjacency' and batch_graphs >= 32: the dataset features are
                                     verified to be the zero-padded adjacency rows (they always are for graphExtender
                                     output), so layer 1's X W1 and X^T dT1 run as aggregations over the graph
                                     structure instead of dense GEMMs -- same results up to fp32 summation order
    TrainingConfig.feature_source    'adjacency' (reference live path: zero-padded adjacency rows are the
                                     features, :373) | 'embedding' (north-star: the learned nn.Embedding is the
                                     input, `inputs = embed.weight[:n]`, as in the legacy trainer
                                     python/utils.py:184; it then receives gradients and Adam updates)
"""
try:
    from python.commons import *  # noqa: F401,F403  (reference spelling, :25)
except ImportError:  # notebook spelling: <pkg>/python on sys.path
    from commons import *  # noqa: F401,F403

import random  # noqa: F401
import weakref
from dataclasses import dataclass
from itertools import permutations
from time import time
from typing import Callable, Dict, List, Optional, Tuple  # noqa: F401

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

from gmc_b200 import _lib as _gmc_lib
from gmc_b200.engine import GCNEngine
from gmc_b200.graph import NOT_ADJACENCY as _NOT_ADJACENCY
from gmc_b200.graph import AdjacencyFeatures, CSRGraph, GraphBatch, check_adjacency_features
from gmc_b200.model import GCNSoftmax, to_device_features
from gmc_b200.optim import FusedAdam

import os as _os

# per-graph optimiser steps are replayed from captured CUDA graphs (GCNEngine.train_step_graphed); GMC_CUDA_GRAPHS=0
# keeps the eager launch sequence
_USE_CUDA_GRAPHS = _os.environ.get("GMC_CUDA_GRAPHS", "1") != "0"
_EPOCH_GRAPH = _os.environ.get("GMC_EPOCH_GRAPH", "1") != "0"     # one captured graph per epoch instead of one per step

TORCH_DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
TORCH_DTYPE = torch.float32


@dataclass
class TrainingConfig:
    """Hyper-parameter record; field-for-field the reference's (:36-60) plus opt-in extensions."""
    n_nodes: int = 1000
    dim_embedding: Optional[int] = None      # defaults to n_nodes
    hidden_dim: Optional[int] = None         # defaults to dim_embedding // 2
    dropout: float = 0.0
    number_classes: int = 3

    learning_rate: float = 0.001
    number_epochs: int = 1000
    tolerance: float = 1e-4
    patience: int = 20
    prob_threshold: float = 0.5

    A: float = 0.0
    C: float = 1.0
    penalty: float = 1000.0

    save_directory: Optional[str] = None
    save_frequency: int = 100

    # ---- extensions (not in the reference; defaults keep its behaviour) ----
    batch_graphs: int = 1
    gemm_precision: str = "fp32"
    loss_mode: str = "ste"
    use_terminal_penalty: bool = False
    feature_source: str = "adjacency"
    adjacency_kernels: bool = False          # batched steps only: X W1 / X^T dT1 as aggregations (csrc/spmm_adj.cu)
    activations: str = "fp32"                # 'bf16' (with gemm_precision='bf16'): layer-1 activations stored in bf16
    preaggregate_features: bool = False      # with bf16 / bf16: layer 1 as relu((A_hat X) W1 + b1), A_hat X built once
    stream_dataset: bool = False             # keep the dataset in pinned host memory and upload every step's graphs (H2D on a
                                             # copy stream, one step ahead) instead of caching device-resident batches
    eval_batch_graphs: int = 64              # evaluate_model: graphs per block-diagonal forward pass
    per_graph_embeddings: bool = False       # feature_source='embedding': every dataset graph owns its [n_g, dim_embedding]
                                             # table (initialised from embed.weight[:n_g]) with its own Adam state, so
                                             # batch_graphs > 1 works; False = the legacy single shared table (utils.py:184)

    def __post_init__(self):
        if self.feature_source not in ("adjacency", "embedding"):
            raise ValueError(f"feature_source must be 'adjacency' or 'embedding', got {self.feature_source!r}")
        if self.dim_embedding is None:
            self.dim_embedding = self.n_nodes
        if self.hidden_dim is None:
            self.hidden_dim = self.dim_embedding // 2


# --------------------------------------------------------------------------------------------
# loss helpers kept for API compatibility (device-agnostic torch; the training loop does not
# call them -- the fused kernel computes the same quantities in closed form)
# --------------------------------------------------------------------------------------------
def override_fixed_nodes(h):
    """Rows 0,1,2 -> e0,e1,e2 in value, identity in gradient (reference :87-94)."""
    k = min(3, h.shape[0])
    eye = torch.eye(h.shape[1], dtype=h.dtype, device=h.device)[:k]
    head = eye + h[:k] - h[:k].detach()
    return torch.cat([head, h[k:]], dim=0)


def max_to_one_hot(tensor):
    """one_hot(argmax) with a straight-through gradient (reference :96-102)."""
    hot = torch.zeros_like(tensor)
    hot[torch.argmax(tensor)] = 1.0
    return hot + tensor - tensor.detach()


def apply_max_to_one_hot(output):
    """Row-wise max_to_one_hot (reference :104-106), vectorised."""
    idx = torch.argmax(output, dim=1, keepdim=True)
    hot = torch.zeros_like(output).scatter_(1, idx, 1.0)
    return hot + output - output.detach()


def extend_matrix_torch_training(matrix, N):
    size = matrix.shape[0]
    if N <= size:
        return matrix
    out = torch.zeros((N, N), dtype=matrix.dtype, device=matrix.device)
    out[:size, :size] = matrix
    return out


def extend_matrix_torch(matrix, N, torch_dtype=None, torch_device=None):
    """Zero-pad the columns of a square matrix to N (reference :137-152)."""
    size = matrix.shape[0]
    if N < size:
        raise ValueError("N should be greater than or equal to the original matrix size.")
    out = torch.zeros(size, N, dtype=matrix.dtype, device=matrix.device)
    out[:, :size] = matrix
    if torch_dtype is not None:
        out = out.type(torch_dtype)
    if torch_device is not None:
        out = out.to(torch_device)
    return out


def calculate_HC_vectorized(s, adjacency_matrix):
    """sum(A * (1 - pad(s s^T))) / 2 (reference :154-176).  The reference pads to a hard-coded
    1000 columns (:171); padding to A's own width is identical whenever the reference runs."""
    same = s @ s.T
    padded = extend_matrix_torch(same, adjacency_matrix.shape[1])
    return torch.sum(adjacency_matrix.to(s.device) * (1 - padded)) / 2


def terminal_independence_penalty(s, terminal_nodes: List[int]):
    """sum_{i<j in T} s_i . s_j (reference :178-195)."""
    total = 0
    for a in range(len(terminal_nodes)):
        for b in range(a + 1, len(terminal_nodes)):
            total = total + torch.dot(s[terminal_nodes[a]], s[terminal_nodes[b]])
    return total


def find_ac_parameters(graph):
    """(max_degree + 1, max_degree / 2) (reference :197-210)."""
    top = max(d for _, d in graph.degree())
    return top + 1, top / 2


def generate_terminal_permutations(terminal_dict: Dict):
    keys = list(terminal_dict.keys())
    return [dict(zip(keys, perm)) for perm in permutations(terminal_dict.values())]


def calculate_all_cut_legacy(q_torch, s):
    """Legacy per-column cut (reference :227-251)."""
    if len(s) == 0:
        return 0
    total = 0
    for i in range(s.shape[1]):
        col = s[:, i].unsqueeze(0)
        total = total + (q_torch * (col != col.t()).float()).sum() / 2
    return total / 2


def evaluate_optimal_partitioning(net, dgl_graph, inputs, adjacency_matrix, terminal_dict: Dict):
    """Legacy helper (reference :253-289): best thresholded cut over terminal permutations."""
    net.eval()
    best = float("inf")
    if dgl_graph.number_of_nodes() < 30:
        inputs = torch.ones((dgl_graph.number_of_nodes(), 30))
    with torch.no_grad():
        for _ in generate_terminal_permutations(terminal_dict):
            probs = override_fixed_nodes(net(dgl_graph, inputs))
            hard = (probs >= 0.5).float()
            value = calculate_all_cut_legacy(adjacency_matrix.to(hard.device), hard)
            if value < best:
                best = value
    return best


def compute_loss(s, adjacency_matrix, A: float = 0, C: float = 1, penalty: float = 1000):
    """C * (-HC); A and penalty are accepted and ignored exactly like the reference (:291-309)."""
    return C * (-1 * calculate_HC_vectorized(s, adjacency_matrix))


# --------------------------------------------------------------------------------------------
# model / optimiser
# --------------------------------------------------------------------------------------------
def setup_model_and_optimizer(config: TrainingConfig):
    """(net, embed, optimizer) as the reference (:311-339).  `embed` is created, registered with
    the optimiser and never receives a gradient -- it only travels into checkpoints as 'inputs'."""
    device = _gmc_lib.require_cuda()
    net = GCNSoftmax(in_feats=config.dim_embedding, hidden_size=config.hidden_dim,
                     num_classes=config.number_classes, dropout=config.dropout, device=device)
    net = net.type(TORCH_DTYPE).to(device)
    net.set_gemm_precision(getattr(config, "gemm_precision", "fp32"))
    embed = nn.Embedding(config.n_nodes, config.dim_embedding)
    embed = embed.type(TORCH_DTYPE).to(device)
    optimizer = FusedAdam(chain(net.parameters(), embed.parameters()), lr=config.learning_rate)
    return net, embed, optimizer


def _engine_for(net, optimizer, config: TrainingConfig) -> GCNEngine:
    precision = getattr(config, "gemm_precision", "fp32")
    adjacency = getattr(config, "feature_source", "adjacency") == "adjacency"
    key = (id(optimizer), float(config.C), getattr(config, "loss_mode", "ste"),
           bool(getattr(config, "use_terminal_penalty", False)), float(config.penalty), precision,
           bool(getattr(config, "adjacency_kernels", False)) and adjacency,
           getattr(config, "activations", "fp32"),
           bool(getattr(config, "preaggregate_features", False)) and adjacency)
    cached = _ENGINES.get(net)
    if cached is not None and cached[0] == key:
        return cached[1]
    if optimizer is not None and not isinstance(optimizer, FusedAdam):
        raise TypeError("the B200 training loop needs the FusedAdam returned by setup_model_and_optimizer")
    # adjacency_features: _prepare verifies that every item's features ARE the zero-padded adjacency rows of its graph
    # before a batch reaches the engine, which is what the integer-feature GEMM path ('bf16x2' / 'bf16x3') relies on
    engine = GCNEngine(net, optimizer, C=config.C, loss_mode=key[2], override_terminals=True,
                       penalty=config.penalty if key[3] else 0.0, precision=precision, adjacency_kernels=key[6],
                       activations=key[7], preaggregate=key[8], adjacency_features=adjacency)
    _ENGINES[net] = (key, engine)
    return engine


_ENGINES = weakref.WeakKeyDictionary()     # net -> (settings key, GCNEngine); never pickled with the model


class _PreparedItem:
    __slots__ = ("batch", "X", "host", "n_graphs")

    def __init__(self, batch, X, host=None, n_graphs=None):
        self.batch, self.X, self.host = batch, X, host
        self.n_graphs = n_graphs if n_graphs is not None else (batch.num_graphs if batch is not None else 0)


def _graph_handle(item) -> CSRGraph:
    handle, _, nx_graph, _ = item
    if isinstance(handle, CSRGraph):
        return handle
    if isinstance(handle, GraphBatch):
        raise TypeError("dataset items hold single graphs; pass GraphBatch objects to GCNEngine directly")
    return from_networkx(nx_graph)  # foreign handle (e.g. legacy DGL object): rebuild from networkx


_check_item_features = check_adjacency_features


def _check_features_are_adjacency(batch: GraphBatch, X: torch.Tensor) -> None:
    """Device-side form of the same check for an already batched dense X (kept for callers that hold one)."""
    from gmc_b200 import ops
    want = ops.densify(batch, X.shape[1])
    if not torch.equal(want, X):
        raise NotImplementedError(_NOT_ADJACENCY)


def _engine_mode(engine: Optional[GCNEngine]) -> tuple:
    if engine is None:
        return ("dense", None)
    if engine.split_fwd:
        return ("integer", engine.F)
    if engine.preaggregate:
        return ("preaggregated", engine.F)
    if engine.adjacency_kernels:
        return ("sparse", engine.F)
    return ("dense", None)


def _device_features(batch: GraphBatch, width: int, engine: Optional[GCNEngine], mode: tuple):
    """The step's feature operand, REBUILT on the device from the (verified) graph structure in the form the engine's
    layer 1 wants -- never a concatenation of the items' dense host tensors."""
    from gmc_b200 import ops
    kind = mode[0]
    if kind == "integer":
        xi = engine._integer_features(batch, None)
        if xi is not None:
            return xi
    elif kind == "preaggregated":
        return ops.PreaggregatedFeatures(ops.preaggregate_features_bf16(batch, width))
    elif kind == "sparse" and ops.adjacency_kernels_apply(batch, width):
        return None
    return ops.densify(batch, width, out=ops.padded_empty(batch.num_nodes, width, batch.device))


def _dist_world() -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _prepare(dataset: Dict, batch_graphs: int, device, engine: Optional[GCNEngine] = None,
             stream: bool = False, build_features: bool = True) -> List[_PreparedItem]:
    """One (GraphBatch, features) pair per optimiser step, cached on the dataset dict's identity.

    * every item's features are verified to be the adjacency rows of its graph (host-side, O(nnz) gather + one
      count_nonzero); the device operand is then rebuilt from the graph (`_device_features`);
    * under torch.distributed each step's chunk of `batch_graphs` graphs is sharded contiguously over the ranks
      (gmc_b200.dist.shard_bounds): `batch_graphs` is the GLOBAL batch, so N ranks take the same optimiser steps as one;
    * stream=True keeps each step's block-diagonal CSR in PINNED HOST memory instead (uploaded every step by
      train_single_epoch, as the reference moves every graph to the device inside its loop, :371-373)."""
    from gmc_b200 import dist as gdist
    rank, world = _dist_world()
    mode = _engine_mode(engine)
    cache = _PREPARED.get(id(dataset))
    keys = list(dataset.keys())
    sig = (keys, batch_graphs, mode, rank, world, bool(stream), bool(build_features))
    if cache is not None and cache[0] is dataset and cache[1] == sig:
        return cache[2]
    steps: List[_PreparedItem] = []
    bg = max(1, batch_graphs)
    for lo in range(0, len(keys), bg):
        chunk_keys = keys[lo: lo + bg]
        if world > 1:
            a, b = gdist.shard_bounds(len(chunk_keys), rank, world)
            chunk_keys = chunk_keys[a:b]
        chunk = [dataset[k] for k in chunk_keys]          # only this rank's items are ever dereferenced
        if not chunk:
            steps.append(_PreparedItem(None, None, n_graphs=0))        # this rank only joins the step's all-reduce
            continue
        handles = [_graph_handle(it) for it in chunk]
        if not build_features:                             # learned embeddings are the input: item[1] is not read
            steps.append(_PreparedItem(GraphBatch(handles, device=device), None))
            continue
        widths = {_check_item_features(h, it[1]) for h, it in zip(handles, chunk)}
        if len(widths) != 1:
            raise ValueError(f"items of one step have different feature widths: {sorted(widths)}")
        width = widths.pop()
        if stream:
            steps.append(_PreparedItem(None, None, host=_HostBatch(handles, width), n_graphs=len(handles)))
            continue
        batch = GraphBatch(handles, device=device)
        steps.append(_PreparedItem(batch, _device_features(batch, width, engine, mode)))
    if len(_PREPARED) > 8:
        _PREPARED.clear()
    _PREPARED[id(dataset)] = (dataset, sig, steps)
    return steps


_PREPARED: Dict[int, tuple] = {}


class _HostBatch:
    """Block-diagonal CSR of one step's graphs in pinned host memory (unit weights, degrees validated here so the
    device side never has to read a flag back)."""

    def __init__(self, handles: List[CSRGraph], width: int):
        sizes = np.asarray([h.n for h in handles], dtype=np.int64)
        nnzs = np.asarray([h.number_of_edges() for h in handles], dtype=np.int64)
        gp = np.zeros(len(handles) + 1, dtype=np.int64)
        np.cumsum(sizes, out=gp[1:])
        ep = np.zeros(len(handles) + 1, dtype=np.int64)
        np.cumsum(nnzs, out=ep[1:])
        if gp[-1] >= 2 ** 31 - 1 or ep[-1] >= 2 ** 31 - 1:
            raise ValueError("batch too large for int32 CSR indices; lower batch_graphs")
        pin = torch.cuda.is_available()              # (host-only unit tests build the same batches unpinned)
        self.rowptr = torch.empty(int(gp[-1]) + 1, dtype=torch.int32, pin_memory=pin)
        self.colidx = torch.empty(int(ep[-1]), dtype=torch.int32, pin_memory=pin)
        self.graph_ptr = torch.from_numpy(gp.astype(np.int32))
        if pin:
            self.graph_ptr = self.graph_ptr.pin_memory()
        rp, ci = self.rowptr.numpy(), self.colidx.numpy()
        rp[0] = 0
        uniform = True
        for i, h in enumerate(handles):
            if not np.all(h.weights == 1.0):
                raise NotImplementedError("stream_dataset=True supports unit edge weights (what GraphCreator writes)")
            deg = np.diff(h.rowptr)
            if h.n and deg.min() == 0:
                from gmc_b200.graph import ZeroDegreeError
                raise ZeroDegreeError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid.")
            uniform = uniform and (h.n == 0 or deg.min() == deg.max())
            rp[gp[i] + 1: gp[i + 1] + 1] = h.rowptr[1:] + ep[i]
            ci[ep[i]: ep[i + 1]] = h.colidx + gp[i]
        self.sizes, self.width = sizes, int(width)
        self.num_nodes, self.nnz = int(gp[-1]), int(ep[-1])
        self.max_nodes = int(sizes.max()) if len(handles) else 0
        self.regular = bool(uniform)                  # every graph regular => one A_hat coefficient per row
        self.nbytes = (self.rowptr.numel() + self.colidx.numel() + self.graph_ptr.numel()) * 4


class _Streamer:
    """Double-buffered upload of _HostBatch steps: the H2D copy of step i+1 runs on a copy stream while step i
    computes; device buffers are reused across steps and epochs (no allocation in the loop)."""

    def __init__(self, device):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [dict(cap=(0, 0, 0)), dict(cap=(0, 0, 0))]
        self.ones = None
        self.feat = None
        self.loss_host = None
        self.prefetched = None      # (host batch, step index whose slot holds it): upload begun by the previous epoch
        self.speculative = True
        self.stats = {"h2d_bytes": 0, "d2h_bytes": 0, "steps": 0}

    def _slot(self, i: int, hb: _HostBatch) -> dict:
        sl = self.slots[i]
        need = (hb.rowptr.numel(), hb.colidx.numel(), hb.graph_ptr.numel())
        if any(n > c for n, c in zip(need, sl["cap"])):
            dev, i32, f32 = self.device, torch.int32, torch.float32
            cap = tuple(max(n, c) for n, c in zip(need, sl["cap"]))
            sl.update(cap=cap, rowptr=torch.empty(cap[0], dtype=i32, device=dev), colidx=torch.empty(cap[1], dtype=i32, device=dev),
                      graph_ptr=torch.empty(cap[2], dtype=i32, device=dev), norm=torch.empty(cap[0], dtype=f32, device=dev),
                      coef=torch.empty(cap[1], dtype=f32, device=dev), scale=torch.empty(cap[0], dtype=f32, device=dev),
                      copied=torch.cuda.Event(), consumed=None)
        return sl

    def issue(self, i: int, hb: _HostBatch) -> None:
        sl = self._slot(i, hb)
        with torch.cuda.stream(self.copy_stream):
            if sl["consumed"] is not None:
                self.copy_stream.wait_event(sl["consumed"])          # the step that last read this slot has finished
            sl["rowptr"][: hb.rowptr.numel()].copy_(hb.rowptr, non_blocking=True)
            sl["colidx"][: hb.colidx.numel()].copy_(hb.colidx, non_blocking=True)
            sl["graph_ptr"][: hb.graph_ptr.numel()].copy_(hb.graph_ptr, non_blocking=True)
            sl["copied"].record(self.copy_stream)
        self.stats["h2d_bytes"] += hb.nbytes

    def batch(self, i: int, hb: _HostBatch) -> GraphBatch:
        sl = self.slots[i]
        torch.cuda.current_stream().wait_event(sl["copied"])
        n, nnz = hb.num_nodes, hb.nnz
        return GraphBatch.from_device_arrays(sl["rowptr"][: n + 1], sl["colidx"][:nnz], sl["graph_ptr"][: len(hb.sizes) + 1],
                                             hb.sizes, hb.max_nodes, norm=sl["norm"][:n], coef=sl["coef"][:nnz])

    def features(self, i: int, hb: _HostBatch, batch: GraphBatch, engine: GCNEngine, mode: tuple):
        from gmc_b200 import ops
        kind = mode[0]
        if kind in ("integer", "preaggregated"):
            want = torch.float16 if (kind == "integer" and engine.split_f16) else torch.bfloat16
            if (self.feat is None or self.feat.dtype != want or self.feat.shape[0] < hb.num_nodes
                    or self.feat.shape[1] != hb.width):
                self.feat = ops.padded_empty_bf16(hb.num_nodes, hb.width, self.device, zero=True, dtype=want)
        if kind == "integer" and hb.regular and hb.max_nodes <= hb.width and engine.H % 4 == 0 and engine.K <= 4:
            if self.ones is None or self.ones.numel() < hb.nnz:
                self.ones = torch.ones(hb.nnz, dtype=torch.float32, device=self.device)
            scale, _ = ops.row_scale(batch, count_nonuniform=False, out=self.slots[i]["scale"][: hb.num_nodes])
            xi = ops.integer_features_bf16(batch, hb.width, out=self.feat[: hb.num_nodes], ones=self.ones)
            return ops.IntegerFeatures(xi, scale)
        if kind == "preaggregated":
            return ops.PreaggregatedFeatures(ops.preaggregate_features_bf16(batch, hb.width, out=self.feat[: hb.num_nodes]))
        if self.feat is None or self.feat.dtype != torch.float32 or self.feat.shape[0] < hb.num_nodes or self.feat.shape[1] != hb.width:
            self.feat = ops.padded_empty(hb.num_nodes, hb.width, self.device)
        return ops.densify(batch, hb.width, out=self.feat[: hb.num_nodes])

    def done(self, i: int) -> None:
        ev = torch.cuda.Event()
        ev.record()
        self.slots[i]["consumed"] = ev

    def read_loss(self, losses: torch.Tensor, out: List[torch.Tensor]) -> None:
        """D2H of the step's per-graph losses into pinned memory (asynchronous; summed after the epoch's sync)."""
        host = torch.empty(losses.numel(), dtype=losses.dtype).pin_memory()
        host.copy_(losses, non_blocking=True)
        out.append(host)
        self.stats["d2h_bytes"] += losses.numel() * losses.element_size()
        self.stats["steps"] += 1


_STREAMERS = weakref.WeakKeyDictionary()      # GCNEngine -> _Streamer


def _train_epoch_streamed(engine: GCNEngine, steps: List[_PreparedItem], device) -> float:
    st = _STREAMERS.get(engine)
    if st is None:
        st = _Streamer(device)
        _STREAMERS[engine] = st
    mode = _engine_mode(engine)
    live = [s for s in steps if s.host is not None]
    host_losses: List[torch.Tensor] = []
    j = 0
    if live:
        if st.prefetched is not None and st.prefetched[0] is live[0].host:
            j = st.prefetched[1]                       # the previous epoch already started this upload (slot parity j & 1)
        else:
            st.issue(0, live[0].host)
    st.prefetched = None
    first = j
    for step in steps:
        if step.host is None:
            engine.train_step_empty()
            continue
        slot = j & 1
        if j - first + 1 < len(live):
            st.issue(slot ^ 1, live[j - first + 1].host)
        elif st.speculative:
            # last step of the epoch: start uploading the first step of the NEXT epoch over the same dataset, so the
            # copy overlaps this step and the epoch-end read-back (epochs revisit the dataset in the same order, :371)
            st.issue(slot ^ 1, live[0].host)
            st.prefetched = (live[0].host, j + 1)
        batch = st.batch(slot, step.host)
        feats = st.features(slot, step.host, batch, engine, mode)
        losses = engine.train_step(batch, feats)
        st.read_loss(losses, host_losses)
        st.done(slot)
        j += 1
    torch.cuda.current_stream().synchronize()
    return float(sum(float(h.sum()) for h in host_losses))


def train_single_epoch(dataset: Dict, net, optimizer, embed, config: TrainingConfig,
                       dataset_files: Optional[List[str]] = None) -> float:
    """One pass over the dataset: forward, override, STE, loss, backward, Adam per graph, in dict
    order (reference :341-390).  Returns the summed loss.  The per-graph `.item()` host sync of the
    reference (:388) is replaced by one device-side accumulation read back once per epoch.
    Under torch.distributed every rank trains its shard of each step's graphs and the returned loss
    is the sum over all ranks."""
    device = _gmc_lib.require_cuda()
    net.train()
    if float(getattr(net, "dropout_frac", 0.0) or 0.0) > 0.0:
        raise NotImplementedError(
            "TrainingConfig.dropout > 0: the fused training step has no dropout between the GraphConv layers "
            "(reference :82 applies F.dropout in training mode); train through the autograd path "
            "(loss = compute_loss(apply_max_to_one_hot(override_fixed_nodes(net(g, x))), ...); loss.backward(); "
            "optimizer.step()), which honours it, or set dropout=0.0 (the reference default)")
    engine = _engine_for(net, optimizer, config)
    if dataset_files is None:
        dataset_files = ["./nx_test_generated_graph_n200_300_d8_12_t500.pkl"]
    rank, world = _dist_world()
    embedding_mode = getattr(config, "feature_source", "adjacency") == "embedding"
    streamed = bool(getattr(config, "stream_dataset", False)) and not embedding_mode
    total = torch.zeros((), dtype=torch.float64, device=device)
    for dataset_file in dataset_files:
        current = dataset if isinstance(dataset, dict) else open_file(dataset_file)
        steps = _prepare(current, int(getattr(config, "batch_graphs", 1)), device, engine, stream=streamed,
                         build_features=not embedding_mode)
        if streamed:
            total += _train_epoch_streamed(engine, steps, device)
            continue
        if (_USE_CUDA_GRAPHS and _EPOCH_GRAPH and not embedding_mode and world == 1
                and all(st.batch is not None for st in steps)):
            # the whole epoch as one captured CUDA graph (the dataset dict is revisited in the same order, :371)
            epoch_total = engine.train_epoch_graphed([(st.batch, st.X) for st in steps])
            if epoch_total is not None:
                total += epoch_total
                continue
        for i_step, step in enumerate(steps):
            if step.batch is None:
                engine.train_step_empty()
            elif embedding_mode and getattr(config, "per_graph_embeddings", False):
                total += _graph_embeddings(embed, steps, current, config).step(engine, i_step, step).sum()
            elif embedding_mode:
                total += _embedding_step(engine, step, embed).sum()
            elif _USE_CUDA_GRAPHS:
                total += engine.train_step_graphed(step.batch, step.X).sum()   # launch-bound per-graph steps: graph replay
            else:
                total += engine.train_step(step.batch, step.X).sum()
    if world > 1:
        from gmc_b200 import dist as gdist
        gdist.all_reduce_sum_(total)
    return float(total.item())


class GraphEmbeddings:
    """Per-graph learned node embeddings (the north-star's input: "embeddings are per-graph, so they stay local").
    One fp32 table row per dataset node of THIS rank's shard, rows of a step's graphs contiguous, 128-byte row pitch;
    every graph starts from embed.weight[:n_g] (so a one-graph dataset reproduces the legacy `inputs = embed.weight`,
    python/utils.py:184) and owns its Adam moments and step count.  Nothing here is ever all-reduced."""

    def __init__(self, embed, steps: List[_PreparedItem], lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        from gmc_b200 import ops
        weight = embed.weight.data
        self.F = int(weight.shape[1])
        ld = ops.pad_cols(self.F)
        self.offsets, total, biggest = [], 0, 0
        for st in steps:
            n = st.batch.num_nodes if st.batch is not None else 0
            self.offsets.append(total)
            total += n
            biggest = max(biggest, n)
        dev = weight.device
        self.param = torch.zeros((total, ld), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        self.grad = torch.zeros((biggest, ld), dtype=torch.float32, device=dev)
        self.steps_taken = [0] * len(steps)
        self.lr, self.betas, self.eps = float(lr), tuple(betas), float(eps)
        for st, off in zip(steps, self.offsets):
            if st.batch is None:
                continue
            gp = st.batch.graph_ptr_host
            for g in range(st.batch.num_graphs):
                lo, hi = int(gp[g]), int(gp[g + 1])
                if hi - lo > weight.shape[0]:
                    raise ValueError(f"graph has {hi - lo} nodes but the embedding table has {weight.shape[0]} rows")
                self.param[off + lo: off + hi, : self.F] = weight[: hi - lo]

    def table(self, i_step: int, n_nodes: int) -> torch.Tensor:
        """[n_nodes, F] view of the rows of step `i_step`'s graphs (block-diagonal order)."""
        off = self.offsets[i_step]
        return self.param[off: off + n_nodes, : self.F]

    def step(self, engine: GCNEngine, i_step: int, prepared: _PreparedItem) -> torch.Tensor:
        from gmc_b200 import ops
        n = prepared.batch.num_nodes
        off = self.offsets[i_step]
        X = self.param[off: off + n, : self.F]
        dX = self.grad[:n, : self.F]
        rows = (self.param[off: off + n], self.grad[:n], self.exp_avg[off: off + n], self.exp_avg_sq[off: off + n])
        self.steps_taken[i_step] += 1
        count = self.steps_taken[i_step]

        def update():
            # with bf16 GEMM operands the same pass leaves the bf16 copy of the rows for the next visit (it survives only
            # when the next step trains the same rows: single-step datasets)
            sh = engine.feature_shadow(X)
            ops.adam_multi([rows[0]], [rows[1]], [rows[2]], [rows[3]], lr=self.lr, beta1=self.betas[0], beta2=self.betas[1],
                           eps=self.eps, step=count, shadows=[sh] if sh is not None else None)
            if sh is not None:
                engine.note_feature_shadow(X)

        return engine.train_step_features(prepared.batch, X, dX, update)


def _graph_embeddings(embed, steps, dataset, config) -> GraphEmbeddings:
    hit = _GRAPH_EMBEDDINGS.get(embed)
    sig = (id(dataset), len(steps), tuple(st.n_graphs for st in steps))
    if hit is None or hit[0] != sig or hit[1] is not dataset:
        hit = (sig, dataset, GraphEmbeddings(embed, steps, config.learning_rate))
        _GRAPH_EMBEDDINGS[embed] = hit
    return hit[2]


def graph_embeddings(embed) -> Optional[GraphEmbeddings]:
    """The per-graph embedding store a per_graph_embeddings=True run trained for `embed` (None before the first step)."""
    hit = _GRAPH_EMBEDDINGS.get(embed)
    return hit[2] if hit is not None else None


_GRAPH_EMBEDDINGS = weakref.WeakKeyDictionary()       # nn.Embedding -> (signature, dataset, GraphEmbeddings)


def _embedding_features(batch: GraphBatch, embed) -> torch.Tensor:
    """inputs = embed.weight[:n] (python/utils.py:184 generalised to the multi-graph loop): one graph per step, its
    node i reads embedding row i."""
    if batch.num_graphs != 1:
        raise NotImplementedError("feature_source='embedding' with the single shared table trains one graph per step "
                                  "(batch_graphs=1): the table is indexed by node id, which a block-diagonal batch "
                                  "would alias; set per_graph_embeddings=True for per-graph tables")
    weight = embed.weight
    if batch.num_nodes > weight.shape[0]:
        raise ValueError(f"graph has {batch.num_nodes} nodes but the embedding table has {weight.shape[0]} rows")
    return weight.data[: batch.num_nodes]


def _embedding_step(engine: GCNEngine, step, embed) -> torch.Tensor:
    weight = embed.weight
    state = _EMBED_STATE.get(embed)                   # keyed by the module (tensors do not hash/compare as keys)
    if state is None or state[0].shape != weight.shape or state[0].device != weight.device:
        state = [torch.zeros_like(weight.data), 0]
        _EMBED_STATE[embed] = state
    grad, prev_rows = state
    n = step.batch.num_nodes
    if prev_rows > n:
        grad[n: prev_rows].zero_()                    # rows a larger previous graph wrote
    state[1] = n
    return engine.train_step(step.batch, _embedding_features(step.batch, embed), feature_param=weight,
                             feature_grad=grad)


_EMBED_STATE = weakref.WeakKeyDictionary()            # nn.Embedding -> [gradient buffer, rows written last step]


def _early_stop_update(epoch: int, loss: float, prev_loss: float, counter: int, config: TrainingConfig):
    """Patience bookkeeping of the reference (:430-437): returns (counter, stop)."""
    if epoch > 0 and (loss > prev_loss or abs(prev_loss - loss) <= config.tolerance):
        counter += 1
        return counter, counter >= config.patience
    return 0, False


def _run_training_loop(config: TrainingConfig, epoch_fn: Callable[[], float], net, optimizer, embed,
                       saver=torch.save) -> Tuple:
    """Epoch loop, early stopping, best-loss tracking, periodic + final checkpoints (reference
    :412-484).  Split out so the host logic is testable without a GPU."""
    best_loss = float("inf")
    best_model_state = None
    loss_history: List[float] = []
    patience_counter = 0
    prev_loss = float("inf")
    start_time = time()
    epoch = -1
    for epoch in range(config.number_epochs):
        cumulative_loss = epoch_fn()
        loss_history.append(cumulative_loss)
        patience_counter, stop = _early_stop_update(epoch, cumulative_loss, prev_loss, patience_counter, config)
        if stop:
            print(f"Early stopping at epoch {epoch}")
            break
        if cumulative_loss < best_loss:
            best_loss = cumulative_loss
            best_model_state = net.state_dict()      # aliases live tensors, as in the reference (:442)
        prev_loss = cumulative_loss
        if epoch % config.save_frequency == 0:
            print(f"Epoch: {epoch}, Cumulative Loss: {cumulative_loss:.6f}")
            if config.save_directory:
                saver(_checkpoint(net, optimizer, embed, epoch, loss_history, config),
                      f"./epoch_{epoch}_loss_{cumulative_loss:.4f}_{config.save_directory}")
    if best_model_state is not None:
        net.load_state_dict(best_model_state)        # no-op by aliasing: the last weights are returned
    training_time = time() - start_time
    print(f"Training completed in {training_time:.2f} seconds")
    print(f"Best loss: {best_loss:.6f}")
    if config.save_directory:
        final_filename = f"./final_{config.save_directory}"
        saver(_checkpoint(net, optimizer, embed, epoch, loss_history, config), final_filename)
        print(f"Final model saved to {final_filename}")
    return net, best_loss, epoch, embed.weight, loss_history


def _checkpoint(net, optimizer, embed, epoch, loss_history, config) -> Dict:
    out = {"epoch": epoch, "model": net.state_dict(), "optimizer": optimizer.state_dict(),
           "loss_history": loss_history, "inputs": embed.weight, "config": config}
    store = graph_embeddings(embed)
    if store is not None:                              # extension key, only when per-graph tables were trained
        out["graph_inputs"] = {"param": store.param, "offsets": list(store.offsets), "dim": store.F}
    return out


def train_model(dataset: Dict, config: TrainingConfig, dataset_files: Optional[List[str]] = None) -> Tuple:
    """Main training function (reference :392-484); returns
    (net, best_loss, final_epoch, embed.weight, loss_history)."""
    print(f"Starting training with {config.number_epochs} epochs")
    print(f"Model: {config.n_nodes} nodes, {config.number_classes} classes")
    print(f"Device: {TORCH_DEVICE}")
    net, embed, optimizer = setup_model_and_optimizer(config)
    return _run_training_loop(
        config, lambda: train_single_epoch(dataset, net, optimizer, embed, config, dataset_files),
        net, optimizer, embed)


def train_from_pickle(dataset_filename: str, model_name: str, n_nodes: int = 1000, **kwargs) -> Tuple:
    """Train from a graphExtender pickle (reference :486-513)."""
    config = TrainingConfig(**{"n_nodes": n_nodes, "save_directory": f"{model_name}.pth", **kwargs})
    print(f"Loading dataset from {dataset_filename}")
    dataset = open_file(dataset_filename)
    return train_model(dataset, config)


def train_multi_class(dataset_filename: str, model_name: str, num_classes: int = 3, **kwargs) -> Tuple:
    """reference :515-535"""
    params = {"number_classes": num_classes, "save_directory": f"{model_name}.pth", **kwargs}
    return train_from_pickle(dataset_filename, model_name, **params)


def evaluate_model(model, dataset: Dict, config: TrainingConfig) -> Dict:
    """No-grad forward + override + STE + loss per graph (reference :537-570)."""
    device = _gmc_lib.require_cuda()
    model.eval()
    cached = _ENGINES.get(model)
    engine = cached[1] if cached is not None else _engine_for(model, None, config)
    total = torch.zeros((), dtype=torch.float64, device=device)
    num_samples = 0
    # block-diagonal batches: per-graph losses are independent of their batch mates, so the sums equal the reference's
    # per-graph loop (:553-562); never sharded -- every rank evaluates the whole dataset, as the reference would
    per_batch = max(1, int(getattr(config, "eval_batch_graphs", 64)))
    for step in _prepare_eval(dataset, per_batch, device, engine):
        total += engine.evaluate(step.batch, step.X).sum()
        num_samples += step.batch.num_graphs
    total_loss = float(total.item())
    return {"average_loss": total_loss / num_samples if num_samples > 0 else 0,
            "total_loss": total_loss, "num_samples": num_samples}


def _prepare_eval(dataset: Dict, per_batch: int, device, engine: GCNEngine) -> List[_PreparedItem]:
    """_prepare without data-parallel sharding (cached separately from the training steps)."""
    mode = _engine_mode(engine)
    keys = list(dataset.keys())
    sig = (keys, per_batch, mode)
    cache = _PREPARED_EVAL.get(id(dataset))
    if cache is not None and cache[0] is dataset and cache[1] == sig:
        return cache[2]
    items = [dataset[k] for k in keys]
    steps: List[_PreparedItem] = []
    for lo in range(0, len(items), per_batch):
        chunk = items[lo: lo + per_batch]
        handles = [_graph_handle(it) for it in chunk]
        widths = {_check_item_features(h, it[1]) for h, it in zip(handles, chunk)}
        if len(widths) != 1:
            raise ValueError(f"items of one evaluation batch have different feature widths: {sorted(widths)}")
        batch = GraphBatch(handles, device=device)
        steps.append(_PreparedItem(batch, _device_features(batch, widths.pop(), engine, mode)))
    if len(_PREPARED_EVAL) > 8:
        _PREPARED_EVAL.clear()
    _PREPARED_EVAL[id(dataset)] = (dataset, sig, steps)
    return steps


_PREPARED_EVAL: Dict[int, tuple] = {}


def load_neural_model(model_path: str, config: TrainingConfig):
    """(net, inputs, loaded_config) from a checkpoint (reference :572-609).  Reference-written
    files pickle their config as `python.Training.TrainingNeural.TrainingConfig` or
    `Training.TrainingNeural.TrainingConfig`; both resolve to this module's dataclass."""
    import torch.serialization
    device = _gmc_lib.require_cuda()
    try:
        torch.serialization.add_safe_globals([TrainingConfig])
        checkpoint = torch.load(model_path, map_location=device)
    except Exception:
        checkpoint = torch.load(model_path, map_location=device, weights_only=False)
    net, embed, _ = setup_model_and_optimizer(config)
    net.load_state_dict(checkpoint["model"])
    return net, checkpoint.get("inputs", embed.weight), checkpoint.get("config", config)


def save_neural_model(model, optimizer, embed, epoch: int, loss_history: List, config: TrainingConfig,
                      model_path: str):
    """reference :611-634"""
    torch.save(_checkpoint(model, optimizer, embed, epoch, loss_history, config), model_path)
    print(f"Model saved to {model_path}")


# --------------------------------------------------------------------------------------------
# legacy compatibility (reference :636-733)
# --------------------------------------------------------------------------------------------
def get_gnn_legacy(n_nodes: int, gnn_hypers: Dict, opt_params: Dict, torch_device, torch_dtype):
    config = TrainingConfig(n_nodes=n_nodes, dim_embedding=gnn_hypers["dim_embedding"],
                            hidden_dim=gnn_hypers["hidden_dim"], dropout=gnn_hypers["dropout"],
                            number_classes=gnn_hypers["number_classes"], learning_rate=opt_params["lr"])
    return setup_model_and_optimizer(config)


def hyperparameters_legacy(n: int = 80, d: int = 3, p=None, graph_type: str = "reg", number_epochs: int = int(1e5),
                           learning_rate: float = 1e-4, prob_threshold: float = 0.5, tol: float = 1e-4,
                           patience: int = 100):
    dim_embedding = n
    return (n, d, p, graph_type, number_epochs, learning_rate, prob_threshold, tol, patience, dim_embedding,
            int(dim_embedding / 2))


def train_legacy_wrapper(model_name: str, filename: str = "./testData/nx_generated_graph_n80_d3_t200.pkl",
                         n: int = 80):
    return train_from_pickle(filename, model_name, n_nodes=n, learning_rate=0.001, patience=20)


def train_2way_neural_legacy(model_name: str, filename: str = "./testData/prepareDS.pkl"):
    return train_multi_class(filename, model_name, num_classes=2, n_nodes=4096, learning_rate=0.001, patience=20,
                             number_epochs=500)


get_gnn = get_gnn_legacy
hyperParameters = hyperparameters_legacy
train1 = train_legacy_wrapper
train_2wayNeural = train_2way_neural_legacy
FIndAC = find_ac_parameters
GetOptimalNetValue = evaluate_optimal_partitioning
calculateAllCut = calculate_all_cut_legacy
LoadNeuralModel = load_neural_model
