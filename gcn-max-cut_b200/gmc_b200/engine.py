"""GCNEngine -- the fused training / inference step of the B200 path.

One explicit kernel sequence per step (no autograd, no torch operators):

    fwd   T1 = X W1                       gmc_gemm_nn            (tensor / CUDA cores)
          H1 = relu(A_hat T1 + b1)        gmc_spmm_symnorm_f32   (bias + ReLU fused)
          T2 = H1 W2                      gmc_skinny_fwd_f32
          Z  = A_hat T2 + b2              gmc_spmm_symnorm_f32
    loss  P, loss_g, dZ                   gmc_softmax_cut_loss_fwd_bwd  (one pass over edges)
    bwd   db2 = colsum(dZ)                gmc_colsum_f32
          dT2 = A_hat dZ                  gmc_spmm_symnorm_f32   (A_hat symmetric)
          dH1pre, dW2, db1                gmc_skinny_bwd_f32     (one pass over H1)
          dT1 = A_hat dH1pre              gmc_spmm_symnorm_f32
          dW1 = X^T dT1                   gmc_gemm_tn            (split-K over all nodes)
         [dX  = dT1 W1^T                  gmc_gemm_nt            only if features are trainable]
    dp    all-reduce(sum) of [dW1|db1|dW2|db2] over NCCL        (only when world_size > 1)
    opt   Adam on W1,b1,W2,b2 [,X]        gmc_adam_multi

It replaces the body of the reference's inner training loop,
python/Training/TrainingNeural.py:371-388 (forward :373, override :374, STE :377, loss :380,
backward :385, optimizer.step :386) and evaluate_model's body (:553-562).

Buffers: two [N,H] activations are enough -- bufA holds T1 then dH1pre, bufB holds H1 then dT1.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops
from .graph import GraphBatch
from .optim import FusedAdam


# forward aggregation: the fused row kernel (aggregate + bias + ReLU + H1 W2 in one pass) unless GMC_FWD_SLAB=1 asks
# for the slab SpMM followed by the separate skinny projection
_FWD_SLAB = os.environ.get("GMC_FWD_SLAB", "0") == "1"
# bf16 activations: GMC_FWD16 = 'slab' (default: bf16 slab SpMM + bf16 skinny projection, 2.4 + 1.3 ms at config 3) |
# 'row' (fused row kernel with bf16 gathers, 4.95 ms: ~775 warp instructions per row, issue-bound)
_FWD16_SLAB = os.environ.get("GMC_FWD16", "slab") == "slab"
# pre-aggregated layer 1: T2 = H1 W2 folded into the GEMM epilogue (GMC_FUSE_PROJ=0: separate skinny_fwd pass over H1)
_FUSE_PROJ = os.environ.get("GMC_FUSE_PROJ", "1") == "1"
# second layer + loss + its backward in one launch per batch (csrc/tail_fused.cu, one CTA per graph); GMC_FUSED_TAIL=0
# keeps the separate spmm_k / cut_loss / colsum / spmm_k kernels
_FUSED_TAIL = os.environ.get("GMC_FUSED_TAIL", "1") == "1"
# split-operand path: the layer-1 GEMM's projection partials are added by the tail kernel while it loads T2 (no reduce launch)
_DEFER_PROJ = os.environ.get("GMC_DEFER_PROJ", "1") == "1"


def _pad4(n: int) -> int:
    return (n + 3) & ~3


class OpTimer:
    """CUDA-event stopwatch per engine op (events on torch's current stream, which is the stream
    every gmc_* call is launched on).  Used by bench.py for the live roofline numbers."""

    def __init__(self):
        self.pending = []            # (name, start_event, stop_event)
        self.total_ms: Dict[str, float] = {}
        self.calls: Dict[str, int] = {}

    def begin(self, name: str):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return (name, ev)

    def end(self, token) -> None:
        stop = torch.cuda.Event(enable_timing=True)
        stop.record()
        self.pending.append((token[0], token[1], stop))

    def collect(self) -> None:
        """Call after a synchronize: folds all finished intervals into the totals."""
        for name, a, b in self.pending:
            self.total_ms[name] = self.total_ms.get(name, 0.0) + a.elapsed_time(b)
            self.calls[name] = self.calls.get(name, 0) + 1
        self.pending = []

    def reset(self) -> None:
        self.pending, self.total_ms, self.calls = [], {}, {}


class GCNEngine:
    def __init__(self, net, optimizer: Optional[FusedAdam] = None, *, C: float = 1.0, loss_mode: str = "ste",
                 override_terminals: bool = True, penalty: float = 0.0, precision: str = "fp32",
                 process_group=None, adjacency_kernels: bool = False, activations: str = "fp32",
                 preaggregate: bool = False, adjacency_features: bool = False, data_parallel: bool = True):
        self.device = _lib.require_cuda()
        self.net = net
        self.optimizer = optimizer
        self.C, self.loss_mode, self.override, self.penalty = float(C), loss_mode, bool(override_terminals), float(penalty)
        if loss_mode not in _lib.LOSS_MODES:
            raise ValueError(f"unknown loss_mode {loss_mode!r}")
        if precision not in _lib.ENGINE_PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}")
        self.precision = precision
        # precision 'bf16x3' / 'bf16x2' -- the fp32-GRADE tensor-core path (csrc/split.cu).  Layer 1 runs pre-aggregated on
        # the integer features XI (2-step path counts, exact in bf16; A_hat X = s . XI with one scale per row), and the
        # fp32 operand of each GEMM is split into bf16 parts: W1 into 3 (2) parts forward -- all 24 (16) mantissa bits --
        # and s . dH1pre into 2 parts backward (16 bits: 4e-6 relative per element, far inside the 1e-4 parity bar; the
        # forward decides the hard labels, so it gets the exact operand).  H1, logits, loss, gradients, Adam: fp32.
        # Needs adjacency features (IntegerFeatures, or adjacency_features=True: the caller guarantees that the
        # features are the zero-padded 0/1 adjacency rows, as _prepare verifies) and one A_hat coefficient per row
        # (regular graphs); other batches take the standard layer 1 with tf32x3 GEMMs (also fp32-grade).
        # 'f16x2': W1 as TWO fp16 parts (11 significant bits each = 22 of fp32's 24; the low part stored scaled by 2^12 and
        # scaled back in the epilogue) against the integer features held as fp16 (one 16-bit format per tcgen05.mma; a
        # bf16 A with an fp16 B faults on B200) -- fp32-grade at the cost of 'bf16x2'.  Backward: s . dH1pre as two fp16
        # parts as well (22 bits; values saturate at +-65504 and lose relative accuracy below 2^-14 -- the max-cut
        # gradients sit at 1e-4 .. 1).
        self.split_fwd = {"bf16x2": 2, "bf16x3": 3, "f16x2": 2}.get(precision, 0)
        self.split_f16 = precision == "f16x2"
        self.split_bwd = int(os.environ.get("GMC_SPLIT_BWD", "2")) if self.split_fwd else 0
        if self.split_bwd not in (0, 2, 3):
            raise ValueError("GMC_SPLIT_BWD must be 2 or 3")
        self.adjacency_features = bool(adjacency_features)
        self._std_precision = "tf32x3" if self.split_fwd else precision
        self.W1s = None
        self.bufS = None
        # activations='bf16' (needs precision='bf16'): the four [n_nodes, hidden] layer-1 tensors -- T1 = X W1,
        # H1 = relu(A_hat T1 + b1), dH1pre, dT1 -- are STORED in bf16 (standard mixed-precision practice: bf16 storage,
        # fp32 arithmetic).  Each is written once and read once (T1: d times by the gather), so every layer-1 pass
        # moves half the bytes; logits, loss, all reductions, gradients and Adam stay fp32.  Batches without an ELL
        # plan (irregular graphs) keep fp32 activations.
        if activations not in ("fp32", "bf16"):
            raise ValueError(f"unknown activations {activations!r}")
        if activations == "bf16" and precision != "bf16":
            raise ValueError("activations='bf16' needs precision='bf16'")
        self.activations = activations
        # preaggregate=True (needs precision='bf16', activations='bf16'): GraphConv layer 1 is evaluated as
        # H1 = relu((A_hat X) W1 + b1) instead of relu(A_hat (X W1) + b1).  The two are the same matrix product; the
        # difference is WHERE the aggregation runs: on the [N, F] features, which do not depend on the weights (once per
        # (graph, features) pair -- for the reference's adjacency features, once per graph), instead of on the [N, H]
        # activations in every forward and every backward pass.  Layer 1 is then ONE GEMM with a bias / ReLU epilogue,
        # dW1 = (A_hat X)^T dH1pre ONE GEMM, and no [N, H] SpMM is left in the step.  Works for any batch (no ELL plan
        # needed).  Not available with trainable features (dX): their aggregate changes every step.
        self.preaggregate = bool(preaggregate)
        if self.preaggregate and (precision != "bf16" or activations != "bf16"):
            raise ValueError("preaggregate=True needs precision='bf16' and activations='bf16'")
        # adjacency_kernels: the caller guarantees that the features ARE the zero-padded unit-weight adjacency rows of
        # the batch (the reference's live input, TrainingNeural.py:373).  X W1 and X^T dT1 then run as aggregations
        # over the batch's ELL plan (csrc/spmm_adj.cu) whenever the batch qualifies, and X itself is never read.
        self.adjacency_kernels = bool(adjacency_kernels)
        self.pg = process_group
        W1, b1, W2, b2 = self.params()
        for p in (W1, b1, W2, b2):
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.GmcError("GCNEngine needs contiguous CUDA float32 parameters")
        self.F, self.H = W1.shape
        self.K = W2.shape[1]
        if self.K > 8:
            raise ValueError("number_classes > 8 is not supported by the fused kernels")
        # flat gradient buffer [dW1 | db1 | dW2 | db2], segments 16-byte aligned
        sizes = [W1.numel(), b1.numel(), W2.numel(), b2.numel()]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += _pad4(s)
        # data parallel on one node: the buffer lives in memory every rank has mapped and the exchange is one kernel over
        # NVLink (dist.PeerAllReduce, csrc/peer.cu); otherwise a torch tensor and NCCL
        from . import dist as gdist
        # (construction is a COLLECTIVE under torch.distributed -- every rank builds its engines in the same order;
        # data_parallel=False builds a purely local engine, e.g. a one-rank reference inside a distributed job)
        self.data_parallel = bool(data_parallel) and optimizer is not None      # inference engines exchange nothing
        self._peer = gdist.make_peer_allreduce(tot, self.device, self.pg) if self.data_parallel else None
        self.grads_flat = self._peer.tensor[:tot] if self._peer is not None else torch.zeros(tot, dtype=torch.float32,
                                                                                             device=self.device)
        self.gW1 = self.grads_flat[offs[0]: offs[0] + sizes[0]].view_as(W1)
        self.gb1 = self.grads_flat[offs[1]: offs[1] + sizes[1]]
        self.gW2 = self.grads_flat[offs[2]: offs[2] + sizes[2]].view_as(W2)
        self.gb2 = self.grads_flat[offs[3]: offs[3] + sizes[3]]
        self.ws = ops.Workspace()
        self.W1p = ops.padded_empty(self.F, self.H, self.device, zero=True)   # 128-byte pitched copy of W1 per step
        # precision 'bf16': bf16 copies of W1 (per step), of the features (once per feature tensor) and of dT1
        # (written directly by the slab SpMM) feed gmc_gemm_bf16; everything else stays fp32
        self.W1b = ops.padded_empty_bf16(self.F, self.H, self.device, zero=True) if precision == "bf16" else None
        self._xb_cache: Dict[tuple, torch.Tensor] = {}
        self._trainable_features = False      # set for the duration of a loss_and_grads call that asks for dX
        self._x16 = None                      # per-step bf16 copy of trainable features
        self._x16_fresh = False
        self._x16_shadow_of = None            # (ptr, version, shape) of the parameter whose bf16 copy Adam wrote into _x16
        self._x16_use_shadow = False
        self._xa_cache: Dict[tuple, torch.Tensor] = {}
        self._w2p = None
        self.bufB16 = None
        self.bufA16 = None
        self._cap_nodes = 0
        self._cap_graphs = 0
        # bumped whenever _ensure replaces an engine buffer: captured CUDA graphs bake raw addresses in, so a graph
        # captured under an older generation (or an older workspace allocation) must never be replayed
        self._buffer_generation = 0
        self.bufA = self.bufB = self.T2 = self.Z = self.P = self.dZ = self.dT2 = None
        self.loss = None
        self.timer: Optional[OpTimer] = None
        self.launch_count = 0          # gmc_* kernels launched (bench.py's gpu_launches claim)
        # CUDA-graph replay of whole training steps (train_step_graphed): one captured graph per (batch, features)
        self._graphs: Dict[tuple, dict] = {}
        self._epoch_graphs: Dict[tuple, dict] = {}
        self._loss_out = None                 # train_epoch_graphed: where the step being captured writes its losses
        self._t2parts = None                  # projection partials of the split-operand layer-1 GEMM (deferred reduce)
        self._step_dev: Optional[torch.Tensor] = None      # device-side Adam step counter shared by all graphs
        self._step_dev_value = 0                           # what that counter holds, mirrored on the host
        self.graph_replays = 0

    def _op(self, name: str, launches: int, fn, *args, **kwargs):
        self.launch_count += launches
        if self.timer is None:
            return fn(*args, **kwargs)
        tok = self.timer.begin(name)
        out = fn(*args, **kwargs)
        self.timer.end(tok)
        return out

    # ------------------------------------------------------------------ plumbing
    def params(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        n = self.net
        return n.conv1.weight, n.conv1.bias, n.conv2.weight, n.conv2.bias

    def grads(self) -> List[torch.Tensor]:
        return [self.gW1, self.gb1, self.gW2, self.gb2]

    def _ensure(self, n_nodes: int, n_graphs: int, b16: bool = False, split: bool = False) -> None:
        dev, f32 = self.device, torch.float32
        if n_nodes > self._cap_nodes:
            cap = n_nodes
            self.T2 = torch.empty((cap, self.K), dtype=f32, device=dev)
            self.Z = torch.empty((cap, self.K), dtype=f32, device=dev)
            self.P = torch.empty((cap, self.K), dtype=f32, device=dev)
            self.dZ = torch.empty((cap, self.K), dtype=f32, device=dev)
            self.dT2 = torch.empty((cap, self.K), dtype=f32, device=dev)
            self._cap_nodes = cap
            self._buffer_generation += 1
        if split:
            # split-operand path: H1 fp32 in bufB, the stacked bf16 parts of s . dH1pre in bufS (zeroed: its pad rows
            # must stay finite)
            if self.bufB is None or self.bufB.shape[0] < n_nodes:
                self.bufB = ops.padded_empty(n_nodes, self.H, dev)
                self._buffer_generation += 1
            need = self.split_bwd * ops.split_rows_for(n_nodes)
            if self.bufS is None or self.bufS.shape[0] < need:
                self.bufS = ops.padded_empty_bf16(need, self.H, dev, zero=True,
                                                  dtype=torch.float16 if self.split_f16 else torch.bfloat16)
                self._buffer_generation += 1
        elif not b16 and (self.bufA is None or self.bufA.shape[0] < n_nodes):
            # row pitch padded to 128 B: TMA boxes / 128-bit gathers never straddle cache lines
            self.bufA = ops.padded_empty(n_nodes, self.H, dev)
            self.bufB = ops.padded_empty(n_nodes, self.H, dev)
            self._buffer_generation += 1
        if self.precision == "bf16" and (self.bufB16 is None or self.bufB16.shape[0] < n_nodes):
            # zero-initialised: the pad columns (hidden .. pitch) stay zero, every kernel that writes it preserves that
            self.bufB16 = ops.padded_empty_bf16(n_nodes, self.H, dev, zero=True)
            self._buffer_generation += 1
        if b16 and (self.bufA16 is None or self.bufA16.shape[0] < n_nodes):
            # bf16 activations: two bf16 buffers carry the whole layer-1 chain (A16 = T1 then dH1pre, B16 = H1 then dT1)
            self.bufA16 = ops.padded_empty_bf16(n_nodes, self.H, dev, zero=True)
            self._buffer_generation += 1
        if n_graphs > self._cap_graphs:
            self.loss = torch.empty(n_graphs, dtype=torch.float64, device=self.device)
            self._cap_graphs = n_graphs
            self._buffer_generation += 1

    def _integer_features(self, batch: GraphBatch, X) -> Optional["ops.IntegerFeatures"]:
        """IntegerFeatures of the batch for the split-operand path, or None when the batch / features do not qualify
        (the caller then runs the standard tf32x3 layer 1)."""
        if not self.split_fwd:
            return None
        if isinstance(X, ops.IntegerFeatures):
            if X.tensor.shape[0] != batch.num_nodes or X.tensor.shape[1] != self.F:
                raise ValueError(f"integer features must be [{batch.num_nodes}, {self.F}], got {tuple(X.tensor.shape)}")
            want = torch.float16 if self.split_f16 else torch.bfloat16
            if X.tensor.dtype != want:
                raise ValueError(f"precision {self.precision!r} takes {want} integer features, got {X.tensor.dtype} "
                                 "(IntegerFeatures.from_batch(..., f16=True) for 'f16x2')")
            return X
        if not (X is None or self.adjacency_features):
            return None
        key = (self.F, self.split_f16)
        hit = getattr(batch, "_xi", None)
        if hit is None or hit[0] != key:
            xi = None
            if batch.unit_weights and batch.max_nodes <= self.F and self.H % 4 == 0 and self.K <= 4:
                scale, bad = ops.row_scale(batch)
                if bad == 0:
                    xi = ops.IntegerFeatures(ops.integer_features_bf16(batch, self.F, f16=self.split_f16), scale)
            hit = (key, xi)
            batch._xi = hit
        return hit[1]

    def _sparse_layer1(self, batch: GraphBatch) -> bool:
        return self.adjacency_kernels and ops.adjacency_kernels_apply(batch, self.F)

    def _b16_activations(self, batch: GraphBatch) -> bool:
        return (self.activations == "bf16" and not self._sparse_layer1(batch)
                and ops.bf16_activations_apply(batch, self.H))

    def _preaggregated(self, batch: GraphBatch, X) -> torch.Tensor:
        """XA = A_hat X in bf16: taken as is from an ops.PreaggregatedFeatures, else computed once per (batch, X) and
        cached on their identity and version (fp32 SpMM over the F feature columns, rounded once)."""
        if isinstance(X, ops.PreaggregatedFeatures):
            XA = X.tensor
            if XA.shape[0] != batch.num_nodes or XA.shape[1] != self.F:
                raise ValueError(f"pre-aggregated features must be [{batch.num_nodes}, {self.F}], got {tuple(XA.shape)}")
            return XA
        X = self._features(batch, X)
        key = (id(batch), X.data_ptr(), tuple(X.shape), X.stride(0), X._version, X.dtype)
        hit = self._xa_cache.get(key)
        # the entry holds the batch AND the tensor it was computed from: a freed X whose address the caching allocator
        # hands to the next batch's features (same shape, _version 0) must not hit
        if hit is None or hit[0] is not batch or hit[1] is not X:
            if len(self._xa_cache) >= 64:
                self._xa_cache.clear()
            X32 = X.float() if X.dtype == torch.bfloat16 else X
            hit = (batch, X, ops.to_bf16(ops.spmm(batch, X32)))
            del X32
            self._xa_cache[key] = hit
        return hit[2]

    def _features(self, batch: GraphBatch, X: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if X is None:
            if not self._sparse_layer1(batch):
                raise ValueError("features may only be omitted when the adjacency kernels apply to the batch")
            return None
        if X.dim() != 2 or X.shape[0] != batch.num_nodes or X.shape[1] != self.F:
            raise ValueError(f"features must be [{batch.num_nodes}, {self.F}], got {tuple(X.shape)}")
        if X.dtype == torch.bfloat16:
            if self.precision != "bf16" or not X.is_cuda:
                raise ValueError("bf16 feature tensors need precision='bf16' and a CUDA tensor")
            return X
        if not X.is_cuda:
            from .model import to_device_features
            X = to_device_features(X, self.device)
        return X

    def _bf16_features(self, X: torch.Tensor) -> torch.Tensor:
        """bf16 copy of a (static) fp32 feature tensor, converted once and cached on its identity and version.  Trainable
        features (a dX was asked for: they change every step, through raw pointers) are converted afresh per step into
        one reused buffer, once per step (the forward converts, the backward reuses)."""
        if X.dtype == torch.bfloat16:
            return X
        if self._trainable_features:
            if self._x16 is None or self._x16.shape[0] < X.shape[0] or self._x16.shape[1] != X.shape[1]:
                self._x16 = ops.padded_empty_bf16(X.shape[0], X.shape[1], X.device, zero=True)
                self._buffer_generation += 1
                self._x16_fresh = False
            view = self._x16[: X.shape[0]]
            if not self._x16_fresh:
                ops.to_bf16(X, out=view)
                self._x16_fresh = True
            return view
        key = (X.data_ptr(), tuple(X.shape), X.stride(0), X._version)
        hit = self._xb_cache.get(key)
        if hit is None or hit[0] is not X:                  # identity, not address: see _preaggregated
            if len(self._xb_cache) >= 64:
                self._xb_cache.clear()
            hit = (X, ops.to_bf16(X))
            self._xb_cache[key] = hit
        return hit[1]

    def invalidate_feature_caches(self) -> None:
        """Forget every derived copy of feature tensors (bf16 copies, A_hat X).  Call after writing features through
        raw pointers (gmc_* kernels such as ops.densify(out=X) do not bump torch's version counter)."""
        self._xb_cache.clear()
        self._xa_cache.clear()
        self._x16_shadow_of = None

    # ------------------------------------------------------------------ forward
    def forward_logits(self, batch: GraphBatch, X: torch.Tensor) -> torch.Tensor:
        """Z (pre-softmax) for the batch; leaves H1 in bufB."""
        N = batch.num_nodes
        _, b2 = self._forward_t2(batch, X), self.params()[3]
        self._op("spmm_k", 1, ops.spmm, batch, self.T2[:N], out=self.Z[:N], bias=b2.data)
        return self.Z[:N]

    def _forward_t2(self, batch: GraphBatch, X: torch.Tensor, defer_proj: bool = False):
        """Layer 1 and the projection of layer 2: T2 = relu(A_hat X W1 + b1) W2 in self.T2; leaves H1 in bufB / bufB16.
        defer_proj (split-operand path): the GEMM's per-n-tile projection partials may stay un-reduced -- returns
        (buffer, n_parts) for layer2_loss_fused(t2_parts=...), which adds them while it loads T2; None otherwise."""
        N = batch.num_nodes
        W1, b1, W2, b2 = self.params()
        if self.preaggregate:
            if self.K > 4 or self.H > 512 or self.H % 4:
                raise NotImplementedError("preaggregate=True needs number_classes <= 4 and hidden_dim <= 512 (multiple of 4)")
            XA = self._preaggregated(batch, X)
            self._ensure(N, batch.num_graphs, b16=True)
            B16 = self.bufB16[:N]
            ops.to_bf16(W1.data, out=self.W1b)
            if _FUSE_PROJ and self.H <= 512:
                # the whole first layer AND the projection of the second: H1 = relu(XA W1 + b1) and T2 = H1 W2 in one
                # GEMM (bias / ReLU / rounding / projection of the rounded rows in the epilogue): H1 is not read again
                if self._w2p is None:
                    self._w2p = torch.zeros(((self.H + 63) // 64 * 64, 4), dtype=torch.float32, device=self.device)
                ops.pad_proj_weights(W2.data, out=self._w2p)
                self._op("gemm_nn_layer1", 1, ops.gemm_bf16_bf16out, "nn", XA, self.W1b, out=B16, bias=b1.data, relu=True,
                         proj_w=self._w2p, proj_out=self.T2[:N], n_proj=self.K)
            else:
                # the whole first layer: H1 = relu(XA W1 + b1), bias and ReLU in the GEMM epilogue, bf16 out
                self._op("gemm_nn_layer1", 1, ops.gemm_bf16_bf16out, "nn", XA, self.W1b, out=B16, bias=b1.data, relu=True)
                self._op("skinny_fwd", 1, ops.skinny_fwd_bf16, B16, W2.data, out=self.T2[:N])
            return self.T2[:N]
        XI = self._integer_features(batch, X)
        if XI is not None:
            # fp32-grade layer 1 on the tensor cores: H1 = relu(s . (XI (W1_hi + W1_lo + W1_lo2)) + b1), fp32 out
            self._ensure(N, batch.num_graphs, split=True)
            Bf = self.bufB[:N]
            if self.W1s is None:
                self.W1s = ops.split_empty(self.F, self.H, self.split_fwd, self.device)
                if self.split_f16:
                    self.W1s = self.W1s.view(torch.float16)
            if self.split_f16:
                ops.f32_split_f16(W1.data, self.split_fwd, out=self.W1s)
            else:
                ops.f32_split_bf16(W1.data, self.split_fwd, out=self.W1s)
            if _FUSE_PROJ:
                # ... and T2 = H1 W2 from the fp32 rows while they are in registers (per-tile partials + ordered reduce)
                if self._w2p is None:
                    self._w2p = torch.zeros(((self.H + 63) // 64 * 64, 4), dtype=torch.float32, device=self.device)
                ops.pad_proj_weights(W2.data, out=self._w2p)
                if defer_proj and _DEFER_PROJ:
                    tiles = ops.split_proj_tiles(self.H, self.split_fwd)
                    if self._t2parts is None or self._t2parts.numel() < tiles * N * 4:
                        self._t2parts = torch.empty(tiles * max(N, self._cap_nodes) * 4, dtype=torch.float32, device=self.device)
                        self._buffer_generation += 1
                    self._op("gemm_nn_layer1", 1, ops.gemm_bf16_split, "nn", XI.tensor, self.W1s, self.split_fwd, self.F,
                             out=Bf, row_scale=XI.scale, bias=b1.data, relu=True, proj_w=self._w2p, n_proj=self.K,
                             proj_parts=self._t2parts)
                    return (self._t2parts, tiles)
                self._op("gemm_nn_layer1", 2, ops.gemm_bf16_split, "nn", XI.tensor, self.W1s, self.split_fwd, self.F, out=Bf,
                         row_scale=XI.scale, bias=b1.data, relu=True, workspace=self.ws, proj_w=self._w2p,
                         proj_out=self.T2[:N], n_proj=self.K)
            else:
                self._op("gemm_nn_layer1", 1, ops.gemm_bf16_split, "nn", XI.tensor, self.W1s, self.split_fwd, self.F, out=Bf,
                         row_scale=XI.scale, bias=b1.data, relu=True, workspace=self.ws)
                self._op("skinny_fwd", 1, ops.skinny_fwd, Bf, W2.data, out=self.T2[:N])
            return self.T2[:N]
        X = self._features(batch, X)
        if self._b16_activations(batch):
            self._ensure(N, batch.num_graphs, b16=True)
            A16, B16 = self.bufA16[:N], self.bufB16[:N]
            ops.to_bf16(W1.data, out=self.W1b)
            self._op("gemm_nn_xw1", 1, ops.gemm_bf16_bf16out, "nn", self._bf16_features(X), self.W1b, out=A16)   # T1, bf16
            if _FWD16_SLAB and self.K <= 4:
                # aggregation through the shared-memory slab kernel, then the skinny projection as its own pass over H1
                self._op("spmm_h_fwd", 1, ops.spmm_bf16, batch, A16, out=B16, bias=b1.data, relu=True)           # H1 bf16
                self._op("skinny_fwd", 1, ops.skinny_fwd_bf16, B16, W2.data, out=self.T2[:N])                  # T2
            else:
                self._op("spmm_h_fused", 1, ops.spmm_fused_skinny_bf16, batch, A16, W2.data, out=B16, proj=self.T2[:N],
                         bias=b1.data, relu=True)                                                               # H1 bf16, T2
            return self.T2[:N]
        self._ensure(N, batch.num_graphs)
        A, Bf = self.bufA[:N], self.bufB[:N]
        ops.copy2d(self.W1p, W1.data)
        if self._sparse_layer1(batch):
            self._op("adj_fwd_xw1", 1, ops.adj_features_fwd, batch, self.W1p, out=A)
        elif self.precision == "bf16":
            ops.to_bf16(W1.data, out=self.W1b)
            self._op("gemm_nn_xw1", 1, ops.gemm_bf16, "nn", self._bf16_features(X), self.W1b, out=A, workspace=self.ws)
        else:
            self._op("gemm_nn_xw1", 1, ops.gemm, "nn", X, self.W1p, out=A, precision=self._std_precision, workspace=self.ws)
        if 16 <= self.H <= 512 and self.H % 4 == 0 and not (_FWD_SLAB and getattr(batch, "plan", None) is not None):
            # aggregation + bias + ReLU + the skinny projection H1 W2 in one pass over the row
            self._op("spmm_h_fused", 1, ops.spmm_fused_skinny, batch, A, W2.data, out=Bf, proj=self.T2[:N],
                     bias=b1.data, relu=True, workspace=self.ws)
        else:
            self._op("spmm_h", 1, ops.spmm, batch, A, out=Bf, bias=b1.data, relu=True)
            self._op("skinny_fwd", 1, ops.skinny_fwd, Bf, W2.data, out=self.T2[:N])
        return self.T2[:N]

    def _fused_tail(self, batch: GraphBatch) -> bool:
        return _FUSED_TAIL and ops.layer2_loss_fused_applies(batch, self.K)

    def forward(self, batch: GraphBatch, X: torch.Tensor) -> torch.Tensor:
        """Softmax probabilities P [N,K] (a view into engine memory; clone to keep)."""
        N = batch.num_nodes
        if self._fused_tail(batch):
            self._forward_t2(batch, X)
            self._op("layer2_loss", 1, ops.layer2_loss_fused, batch, self.T2[:N], self.params()[3].data, self.loss_mode,
                     self.override, self.penalty, self.C, P=self.P[:N], loss=self.loss[: batch.num_graphs], Z=self.Z[:N])
            return self.P[:N]
        Z = self.forward_logits(batch, X)
        self._op("cut_loss", 1, ops.cut_loss, batch, Z, self.loss_mode, self.override, self.penalty, self.C,
                 need_P=True, need_dZ=False, P=self.P[:N], loss=self.loss[: batch.num_graphs])
        return self.P[:N]

    def evaluate(self, batch: GraphBatch, X: torch.Tensor) -> torch.Tensor:
        """Per-graph loss (float64 [B], device) without gradients -- evaluate_model body (:553-562)."""
        self.forward(batch, X)
        return self.loss[: batch.num_graphs]

    # ------------------------------------------------------------------ backward
    def loss_and_grads(self, batch: GraphBatch, X: torch.Tensor, dX: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Forward + loss + full backward at the current weights.  Gradients land in
        self.grads_flat (overwritten); returns per-graph loss (float64 [B], device view)."""
        if self.preaggregate:
            if dX is not None:
                raise ValueError("preaggregate=True folds the aggregation into fixed features; trainable features (dX) "
                                 "need the standard path")
        elif self._integer_features(batch, X) is None:
            X = self._features(batch, X)
        if dX is not None and self._sparse_layer1(batch):
            raise ValueError("adjacency_kernels promises fixed adjacency features; trainable features need the dense path")
        N, B = batch.num_nodes, batch.num_graphs
        W1, b1, W2, b2 = self.params()
        self._trainable_features, self._x16_fresh = dX is not None, False
        if dX is not None and self._x16_use_shadow:
            # train_step found the bf16 copy the previous step's feature Adam left behind still valid for this table
            self._x16_fresh = self._x16 is not None and self._x16.shape[0] >= X.shape[0] and self._x16.shape[1] == X.shape[1]
        self._x16_use_shadow = False
        try:
            return self._loss_and_grads(batch, X, dX, N, B)
        finally:
            self._trainable_features = False

    def _loss_and_grads(self, batch: GraphBatch, X, dX, N: int, B: int) -> torch.Tensor:
        W1, b1, W2, b2 = self.params()
        if self._fused_tail(batch):
            # Z = A_hat T2 + b2, softmax / override / STE / loss, dZ, db2, dT2 = A_hat dZ: one launch (+ the db2 reduce)
            parts = self._forward_t2(batch, X, defer_proj=True)
            parts = parts if isinstance(parts, tuple) else None
            loss = self.loss[:B] if self._loss_out is None else self._loss_out
            self._op("layer2_loss", 2 if B > 1 else 1, ops.layer2_loss_fused, batch, self.T2[:N], b2.data, self.loss_mode,
                     self.override, self.penalty, self.C, P=self.P[:N], loss=loss, dT2=self.dT2[:N], db2=self.gb2, Z=self.Z[:N],
                     dZ=self.dZ[:N], workspace=self.ws, t2_parts=parts)
        else:
            Z = self.forward_logits(batch, X)
            loss = self.loss[:B] if self._loss_out is None else self._loss_out
            self._op("cut_loss", 1, ops.cut_loss, batch, Z, self.loss_mode, self.override, self.penalty, self.C,
                     need_P=True, need_dZ=True, P=self.P[:N], dZ=self.dZ[:N], loss=loss)
            self._op("colsum_db2", 2, ops.colsum, self.dZ[:N], out=self.gb2, workspace=self.ws)
            self._op("spmm_k", 1, ops.spmm, batch, self.dZ[:N], out=self.dT2[:N])
        if self.preaggregate:
            XA = self._preaggregated(batch, X)
            A16, B16 = self.bufA16[:N], self.bufB16[:N]
            self._op("skinny_bwd", 2, ops.skinny_bwd_bf16, self.dT2[:N], W2.data, B16, dH=A16, dW=self.gW2,
                     dbias=self.gb1, workspace=self.ws)                                              # dH1pre, bf16
            self._op("gemm_tn_dw1", 2, ops.gemm_bf16, "tn", XA, A16, out=self.gW1, workspace=self.ws)  # (A_hat X)^T dH1pre
            return loss
        XI = self._integer_features(batch, X)
        if XI is not None:
            if dX is not None:
                raise ValueError("the integer-feature path has fixed adjacency features; trainable features need tf32 / fp32")
            ns = self.split_bwd
            S = self.bufS[: ns * ops.split_rows_for(N)]
            self._op("skinny_bwd", 2, ops.skinny_bwd_split, self.dT2[:N], W2.data, self.bufB[:N], ns, S,
                     row_scale=XI.scale, dW=self.gW2, dbias=self.gb1, workspace=self.ws)             # s . dH1pre, bf16 parts
            self._op("gemm_tn_dw1", 2, ops.gemm_bf16_split, "tn", XI.tensor, S, ns, N, out=self.gW1, workspace=self.ws)
            return loss
        if self._b16_activations(batch):
            A16, B16 = self.bufA16[:N], self.bufB16[:N]
            self._op("skinny_bwd", 2, ops.skinny_bwd_bf16, self.dT2[:N], W2.data, B16, dH=A16, dW=self.gW2,
                     dbias=self.gb1, workspace=self.ws)                                              # dH1pre, bf16
            self._op("spmm_h", 1, ops.spmm_bf16, batch, A16, out=B16)                                # dT1, bf16
            self._op("gemm_tn_dw1", 2, ops.gemm_bf16, "tn", self._bf16_features(X), B16, out=self.gW1, workspace=self.ws)
            if dX is not None:                                                                       # dX = dT1 W1^T, fp32 out
                self._op("gemm_nt_dx", 1, ops.gemm_bf16, "nt", B16, self.W1b, out=dX, workspace=self.ws)
            return loss
        A, Bf = self.bufA[:N], self.bufB[:N]
        self._op("skinny_bwd", 2, ops.skinny_bwd, self.dT2[:N], W2.data, Bf, dH=A, dW=self.gW2, dbias=self.gb1,
                 workspace=self.ws)
        if self.precision == "bf16" and not self._sparse_layer1(batch):
            dT1b = self._op("spmm_h", 1, ops.spmm_bf16out, batch, A, out=self.bufB16[:N])            # dT1, bf16
            self._op("gemm_tn_dw1", 2, ops.gemm_bf16, "tn", self._bf16_features(X), dT1b, out=self.gW1, workspace=self.ws)
            if dX is not None:
                self._op("gemm_nt_dx", 1, ops.gemm_bf16, "nt", dT1b, self.W1b, out=dX, workspace=self.ws)
            return loss
        self._op("spmm_h", 1, ops.spmm, batch, A, out=Bf)                                   # dT1
        if self._sparse_layer1(batch):
            self._op("adj_bwd_dw1", 2, ops.adj_features_bwd, batch, Bf, out=self.gW1, workspace=self.ws)
        else:
            self._op("gemm_tn_dw1", 2, ops.gemm, "tn", X, Bf, out=self.gW1, precision=self._std_precision, workspace=self.ws)
        if dX is not None:
            self._op("gemm_nt_dx", 1, ops.gemm, "nt", Bf, self.W1p, out=dX, precision=self._std_precision,
                     workspace=self.ws)
        return loss

    def allreduce_grads(self) -> None:
        """Data-parallel exchange: ONE NCCL all-reduce of the 502 003-float gradient buffer."""
        import torch.distributed as dist
        if not self.data_parallel:
            return
        if self.pg is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            # timed as its own op: the interval ends when the LAST rank's gradients have arrived, so it contains the
            # wait for slower ranks (bench.py reports it per rank)
            if self._peer is not None:
                self._op("allreduce", 1, self._peer.all_reduce_)
            else:
                self._op("allreduce", 0, dist.all_reduce, self.grads_flat, op=dist.ReduceOp.SUM, group=self.pg)

    def apply_adam(self, feature_param: Optional[torch.Tensor] = None,
                   feature_grad: Optional[torch.Tensor] = None) -> None:
        if self.optimizer is None:
            raise RuntimeError("GCNEngine was built without an optimizer")
        self._op("adam", 1, self.optimizer.fused_step, list(self.params()), self.grads())
        if feature_param is not None:
            # per-graph / per-node embeddings never leave the rank: no all-reduce, their own Adam launch.  With bf16 GEMM
            # operands the same pass also writes the bf16 copy the next step's GEMMs read (30 instead of 28 bytes per
            # element, against a separate 6-byte conversion pass over the table): valid for the next loss_and_grads as long
            # as nobody else writes the parameter in between (its autograd version is unchanged).
            shadow = None
            x16 = self._x16
            if (self.precision == "bf16" and x16 is not None and feature_param.dim() == 2 and feature_param.is_contiguous()
                    and hasattr(self.optimizer, "fused_step")):
                base = x16._base if x16._base is not None else x16
                if base.is_contiguous() and base.shape == feature_param.shape:
                    shadow = base
            if shadow is not None:
                self._op("adam_features", 1, self.optimizer.fused_step, [feature_param], [feature_grad], shadows=[shadow])
                self._x16_shadow_of = self._shadow_key(feature_param)
            else:
                self._op("adam_features", 1, self.optimizer.fused_step, [feature_param], [feature_grad])
                self._x16_shadow_of = None

    # ------------------------------------------------------------------ CUDA-graph replay
    def _adam_state(self):
        opt = self.optimizer
        params = list(self.params())
        group = None
        for g in opt.param_groups:
            if any(p is q for q in g["params"] for p in params[:1]):
                group = g
        if group is None:
            raise RuntimeError("the model parameters are not registered with the optimiser")
        states = [opt._state_for(p) for p in params]
        return params, group, states

    def _train_step_capturable(self, batch: GraphBatch, X: torch.Tensor) -> torch.Tensor:
        """train_step with nothing host-dependent baked in: Adam reads its step count from device memory."""
        loss = self.loss_and_grads(batch, X)
        params, group, states = self._adam_state()
        b1, b2 = group["betas"]
        ops.adam_multi([p.data for p in params], self.grads(), [s["exp_avg"] for s in states],
                       [s["exp_avg_sq"] for s in states], lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"],
                       step_dev=self._step_dev)
        self.launch_count += 1
        return loss

    def train_step_graphed(self, batch: GraphBatch, X: torch.Tensor) -> torch.Tensor:
        """train_step replayed from a captured CUDA graph.  The reference trains one small graph per optimiser step
        (TrainingNeural.py:371-388): ~15 kernels of a few microseconds each, so the loop is launch-bound; every
        (batch, features) pair of a dataset is visited once per epoch, which makes it a natural unit to capture.
        First visit: eager step (sizes buffers, primes the kernels' one-time attribute calls).  Second visit: capture
        + replay.  Later visits: one cudaGraphLaunch.  Bit-identical to train_step (same kernels, same order);
        falls back to the eager step under data parallelism or while a timer is attached."""
        import torch.distributed as dist
        if self.timer is not None or self.pg is not None or (dist.is_available() and dist.is_initialized()
                                                             and dist.get_world_size() > 1):
            return self.train_step(batch, X)
        params, group, states = self._adam_state()
        xkey = (X.data_ptr(), tuple(X.shape)) if torch.is_tensor(X) else id(X)     # feature wrappers / None: identity
        key = (id(batch), xkey, group["lr"], tuple(group["betas"]), group["eps"])
        entry = self._graphs.get(key)
        if entry is None or entry["batch"] is not batch or entry["X"] is not X:
            if len(self._graphs) >= 4096:
                self._graphs.clear()
            self._graphs[key] = {"batch": batch, "X": X, "graph": None, "loss": None}   # keeps both alive
            return self.train_step(batch, X)
        generation = (self._buffer_generation, self.ws.generation)
        if entry["graph"] is not None and entry["generation"] != generation:
            # an engine buffer or the workspace was reallocated since the capture (a larger batch came by): the graph
            # holds freed addresses -- drop it; this visit runs eagerly (it may grow buffers again), the next re-captures
            entry["graph"], entry["loss"] = None, None
            return self.train_step(batch, X)
        if entry["graph"] is None:
            steps = {int(s["step"].item()) for s in states}
            if len(steps) != 1:
                return self.train_step(batch, X)             # parameters with different histories: stay eager
            if self._step_dev is None:
                self._step_dev = torch.zeros(8, dtype=torch.int64, device=self.device)
                self._step_dev_value = 0
            launches = self.launch_count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                entry["loss"] = self._train_step_capturable(batch, X)
            entry["graph"], entry["launches"] = graph, self.launch_count - launches
            entry["generation"] = (self._buffer_generation, self.ws.generation)
            self.launch_count = launches
        # eager steps (first visits of other items) advance only the host counters: resynchronise the device one
        host_step = int(states[0]["step"].item())
        if host_step != self._step_dev_value:
            self._step_dev[:1].fill_(host_step)
        entry["graph"].replay()
        self._step_dev_value = host_step + 1
        self.graph_replays += 1
        self.launch_count += entry["launches"]
        for st in states:
            st["step"] += 1
        return entry["loss"]

    def train_epoch_graphed(self, items) -> Optional[torch.Tensor]:
        """All optimiser steps of one epoch -- `items` = [(batch, X), ...] in dataset order -- replayed from ONE captured
        CUDA graph; returns the summed loss as a float64 device scalar, or None when the epoch has to run step by step
        (data parallel, a timer attached, too many items).  The reference revisits the same dataset dict in the same
        order every epoch (TrainingNeural.py:371), so the whole epoch is a fixed kernel sequence: one cudaGraphLaunch per
        epoch instead of one per graph (config 1: 80 -> 64 us per 500-node step; the per-step form leaves the GPU idle
        between replays and adds two eager reduction launches per step).  First visit of a dataset: eager steps (sizes the
        buffers, primes one-time kernel attributes); second: capture + replay.  Same kernels in the same order as
        train_step: bit-identical weights and losses."""
        import torch.distributed as dist
        if (self.timer is not None or self.pg is not None or not items or len(items) > 512
                or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)):
            return None
        params, group, states = self._adam_state()

        def xkey(X):
            return (X.data_ptr(), tuple(X.shape)) if torch.is_tensor(X) else id(X)

        key = (tuple((id(b), xkey(X)) for b, X in items), group["lr"], tuple(group["betas"]), group["eps"])
        entry = self._epoch_graphs.get(key)
        same = entry is not None and all(eb is b and eX is X for (eb, eX), (b, X) in zip(entry["items"], items))
        generation = (self._buffer_generation, self.ws.generation)
        if same and entry["graph"] is not None and entry["generation"] != generation:
            entry["graph"], entry["total"] = None, None      # buffers moved since the capture: this visit runs eagerly
            same = False
        if not same:
            if len(self._epoch_graphs) >= 16:
                self._epoch_graphs.clear()
            self._epoch_graphs[key] = {"items": list(items), "graph": None, "total": None}   # keeps the items alive
            total = torch.zeros((), dtype=torch.float64, device=self.device)
            for b, X in items:
                total += self.train_step(b, X).sum()
            return total
        if entry["graph"] is None:
            if len({int(s["step"].item()) for s in states}) != 1:
                return None                                   # parameters with different histories: stay per step
            if self._step_dev is None:
                self._step_dev = torch.zeros(8, dtype=torch.int64, device=self.device)
                self._step_dev_value = 0
            launches = self.launch_count
            offsets = [0]
            for b, _ in items:
                offsets.append(offsets[-1] + b.num_graphs)
            entry["losses"] = torch.zeros(offsets[-1], dtype=torch.float64, device=self.device)
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    for i, (b, X) in enumerate(items):
                        self._loss_out = entry["losses"][offsets[i]: offsets[i + 1]]
                        self._train_step_capturable(b, X)
                    entry["total"] = entry["losses"].sum()
            finally:
                self._loss_out = None
            entry["graph"], entry["launches"] = graph, self.launch_count - launches
            entry["generation"] = (self._buffer_generation, self.ws.generation)
            self.launch_count = launches
        host_step = int(states[0]["step"].item())
        if host_step != self._step_dev_value:
            self._step_dev[:1].fill_(host_step)
        entry["graph"].replay()
        self._step_dev_value = host_step + len(items)
        self.graph_replays += 1
        self.launch_count += entry["launches"]
        for st in states:
            st["step"] += len(items)
        return entry["total"].clone()

    def train_step_features(self, batch: GraphBatch, X: torch.Tensor, dX: torch.Tensor, update) -> torch.Tensor:
        """One optimiser step with trainable features owned by the caller: dL/dX lands in `dX`, the shared weights take
        their (all-reduced) Adam step, then `update()` runs the caller's update of the rows X views (per-graph embedding
        tables never leave the rank: no all-reduce for them)."""
        self._x16_use_shadow = self._x16_shadow_of == self._shadow_key(X)
        loss = self.loss_and_grads(batch, X, dX=dX)
        self.allreduce_grads()
        self.apply_adam()
        self._op("adam_features", 1, update)
        return loss

    @staticmethod
    def _shadow_key(X: torch.Tensor):
        return (X.data_ptr(), X._version, (int(X.shape[0]), int(X.stride(0))))

    def feature_shadow(self, X: torch.Tensor) -> Optional[torch.Tensor]:
        """For callers that own the Adam update of trainable features (train_step_features): the bf16 buffer
        [rows, row pitch] the engine's GEMMs read X from, or None (no bf16 operands, pitch mismatch, not allocated yet).
        Fill it in the update -- ops.adam_multi(..., shadows=[buffer]) writes bf16(parameter) at the same element index --
        and call note_feature_shadow(X): the next step over the same, untouched X skips its conversion pass."""
        if self.precision != "bf16" or self._x16 is None or not torch.is_tensor(X) or X.dim() != 2:
            return None
        base = self._x16._base if self._x16._base is not None else self._x16
        if base.dim() != 2 or not base.is_contiguous() or X.stride(0) != base.shape[1] or X.shape[0] > base.shape[0]:
            return None
        return base[: X.shape[0]]

    def note_feature_shadow(self, X: torch.Tensor) -> None:
        self._x16_shadow_of = self._shadow_key(X)

    def train_step_empty(self) -> None:
        """The optimiser step of a rank whose shard of the step's graphs is empty (data parallel, fewer graphs than
        ranks in the last chunk): zero gradients into the all-reduce, then the same Adam update as everyone else."""
        self.grads_flat.zero_()
        self.allreduce_grads()
        self.apply_adam()

    def train_step(self, batch: GraphBatch, X: torch.Tensor, feature_param: Optional[torch.Tensor] = None,
                   feature_grad: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimiser step on the whole batch (loss = sum of per-graph losses).

        Learned node-embedding input (the north-star's feature mode; reference precedent python/utils.py:184,
        `inputs = embed.weight`): pass the parameter that X is a view of (`feature_param`, registered with the
        optimiser) and a same-shaped gradient buffer (`feature_grad`).  dL/dX = dT1 W1^T is written into the
        rows/columns of `feature_grad` that X covers (everything else must be zero) and the parameter takes its
        own fused Adam update."""
        dX = None
        if feature_param is not None:
            if feature_grad is None or feature_grad.shape != feature_param.shape or not feature_grad.is_contiguous():
                raise ValueError("feature_grad must be a contiguous buffer shaped like feature_param")
            if X.data_ptr() != feature_param.data_ptr() or X.shape[0] > feature_param.shape[0] \
                    or X.stride(0) != feature_param.stride(0):
                raise ValueError("X must be a leading-rows / leading-columns view of feature_param")
            dX = feature_grad[: X.shape[0], : X.shape[1]]
            # bf16 copy written by the previous step's feature Adam: still this table, untouched since (torch's version
            # counter; writes through .data or raw pointers are invisible to it -- invalidate_feature_caches() after those)
            self._x16_use_shadow = self._x16_shadow_of == self._shadow_key(feature_param)
        loss = self.loss_and_grads(batch, X, dX=dX)
        self.allreduce_grads()
        self.apply_adam(feature_param, feature_grad)
        return loss
