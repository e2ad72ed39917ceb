"""GraphConv / GCNSoftmax as torch modules whose forward AND backward run on libgcnmaxcut.

Mirrors the reference's model surface (python/Training/TrainingNeural.py:69-85 and
dgl.nn.pytorch.GraphConv with its defaults): same constructor arguments, parameter names
(`conv1.weight [in,out]`, `conv1.bias`, `conv2.weight`, `conv2.bias`), initialisation
(xavier_uniform_ / zeros) and `forward(g, inputs)` signature, so state_dicts are
interchangeable with reference checkpoints.

This module is the *generic* path (any caller can use autograd through it).  Training via
`train_single_epoch` / `train_model` uses gmc_b200.engine instead, which fuses the loss and
skips autograd entirely.
"""
from __future__ import annotations

from typing import Optional, Union

import torch
import torch.nn as nn

from . import _lib, ops
from .graph import AdjacencyFeatures, CSRGraph, GraphBatch

GraphLike = Union[CSRGraph, GraphBatch]


def as_batch(g: GraphLike) -> GraphBatch:
    """Device batch for a graph handle; single graphs cache their one-graph batch."""
    if isinstance(g, GraphBatch):
        return g
    if isinstance(g, CSRGraph):
        if g._batch is None:
            g._batch = GraphBatch([g])
        return g._batch
    if hasattr(g, "number_of_nodes") and hasattr(g, "edges"):
        raise TypeError("got a foreign graph object (DGLGraph?); rebuild it with gmc_b200.graph.from_networkx "
                        "or load the pickle through commons.open_file, which converts legacy DGL handles")
    raise TypeError(f"unsupported graph handle {type(g)!r}")



def to_device_features(x: torch.Tensor, device) -> torch.Tensor:
    """Reference datasets hold CPU float32 tensors (graphExtender.py:110-114); upload once and cache."""
    if isinstance(x, AdjacencyFeatures):
        # implied features (graphExtender dense_features=False): densified on the device from the graph
        key = ("adjacency", id(x))
        hit = _FEATURE_CACHE.get(key)
        if hit is None or hit[0] is not x:
            if len(_FEATURE_CACHE) > 1024:
                _FEATURE_CACHE.clear()
            batch = as_batch(x.graph)
            hit = (x, ops.densify(batch, x.n_cols, out=ops.padded_empty(batch.num_nodes, x.n_cols, batch.device)))
            _FEATURE_CACHE[key] = hit
        return hit[1]
    if x.is_cuda:
        return x if x.dtype == torch.float32 else x.float()
    key = (x.data_ptr(), tuple(x.shape), x._version)
    hit = _FEATURE_CACHE.get(key)
    if hit is None or hit[0] is not x:
        if len(_FEATURE_CACHE) > 1024:
            _FEATURE_CACHE.clear()
        hit = (x, x.to(device=device, dtype=torch.float32, non_blocking=False))
        _FEATURE_CACHE[key] = hit
    return hit[1]


_FEATURE_CACHE = {}


class _GraphConvFn(torch.autograd.Function):
    """Y = act(A_hat (X W) + b)  [in > out]   or   act((A_hat X) W + b)  [in <= out]"""

    @staticmethod
    def forward(ctx, X, W, b, batch: GraphBatch, relu: bool, precision: str):
        n_in, n_out = W.shape
        mult_first = n_in > n_out
        Wc = W.contiguous()
        if mult_first:
            T = ops.skinny_fwd(X, Wc) if n_out <= 8 else ops.gemm("nn", X, Wc, precision=precision)
            Y = ops.spmm(batch, T, bias=b, relu=relu)
            ctx.save_for_backward(X, Wc, Y if relu else None)
        else:
            S = ops.spmm(batch, X)
            Y = ops.gemm("nn", S, Wc, precision=precision)
            Y = _bias_act(Y, b, relu)
            ctx.save_for_backward(S, Wc, Y if relu else None)
        ctx.batch, ctx.relu, ctx.precision, ctx.mult_first, ctx.has_bias = batch, relu, precision, mult_first, b is not None
        return Y

    @staticmethod
    def backward(ctx, dY):
        A, W, Y = ctx.saved_tensors
        batch, precision = ctx.batch, ctx.precision
        dY = dY.contiguous()
        if ctx.relu:
            dY = _relu_mask(dY, Y)
        db = ops.colsum(dY) if ctx.has_bias and ctx.needs_input_grad[2] else None
        dX = dW = None
        if ctx.mult_first:
            dT = ops.spmm(batch, dY)                          # A_hat symmetric: same SpMM
            if ctx.needs_input_grad[1]:
                dW = ops.gemm("tn", A, dT, precision=precision)
            if ctx.needs_input_grad[0]:
                dX = ops.gemm("nt", dT, W, precision=precision)
        else:
            if ctx.needs_input_grad[1]:
                dW = ops.gemm("tn", A, dY, precision=precision)
            if ctx.needs_input_grad[0]:
                dX = ops.spmm(batch, ops.gemm("nt", dY, W, precision=precision))
        return dX, dW, db, None, None, None


def _bias_act(Y, b, relu):
    # only reached for in_feats <= out_feats layers (never in the reference's 1000->500->3 net)
    if b is not None:
        Y = Y + b
    return torch.relu(Y) if relu else Y


def _relu_mask(dY, Y):
    return ops.relu_bwd(dY, Y)


class _SoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Z):
        P = ops.softmax_fwd(Z)
        ctx.save_for_backward(P)
        return P

    @staticmethod
    def backward(ctx, dP):
        (P,) = ctx.saved_tensors
        return ops.softmax_bwd(P, dP)


class GraphConv(nn.Module):
    """dgl.nn.pytorch.GraphConv(in_feats, out_feats) with DGL's defaults (norm='both', weight, bias)."""

    def __init__(self, in_feats: int, out_feats: int, norm: str = "both", weight: bool = True, bias: bool = True,
                 activation=None, allow_zero_in_degree: bool = False):
        super().__init__()
        if norm != "both" or not weight:
            raise NotImplementedError("only norm='both' with a weight matrix is implemented (the reference's usage)")
        self._in_feats, self._out_feats = in_feats, out_feats
        self._activation = activation
        self._allow_zero_in_degree = allow_zero_in_degree
        self.gemm_precision = "fp32"
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph: GraphLike, feat: torch.Tensor, fuse_relu: bool = False) -> torch.Tensor:
        dev = _lib.require_cuda()
        if not self.weight.is_cuda:
            raise _lib.GmcError("GraphConv parameters live on the CPU; move the module to CUDA "
                                "(the hot path has no CPU implementation)")
        batch = as_batch(graph)
        x = to_device_features(feat, dev)
        y = _GraphConvFn.apply(x, self.weight, self.bias, batch, fuse_relu, self.gemm_precision)
        if self._activation is not None:
            y = self._activation(y)
        return y

    def extra_repr(self):
        return f"in={self._in_feats}, out={self._out_feats}, normalization=both"


class GCNSoftmax(nn.Module):
    """Two GraphConv layers + softmax; reference python/Training/TrainingNeural.py:69-85."""

    def __init__(self, in_feats: int, hidden_size: int, num_classes: int, dropout: float, device):
        super().__init__()
        self.dropout_frac = dropout
        self.conv1 = GraphConv(in_feats, hidden_size).to(device)
        self.conv2 = GraphConv(hidden_size, num_classes).to(device)

    def set_gemm_precision(self, precision: str) -> None:
        if precision not in _lib.ENGINE_PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}")
        # 'bf16' is an engine (training-step) mode with resident bf16 operands; this generic autograd path keeps fp32
        # operands and runs them through the one-pass TF32 kernel instead
        # ('bf16x3' / 'bf16x2' are fp32-grade engine modes on integer features: the generic path uses tf32x3 for them)
        generic = {"bf16": "tf32", "bf16x3": "tf32x3", "bf16x2": "tf32x3", "f16x2": "tf32x3"}.get(precision, precision)
        self.conv1.gemm_precision = generic
        self.conv2.gemm_precision = generic

    def forward(self, g: GraphLike, inputs: torch.Tensor) -> torch.Tensor:
        # ReLU is fused into the first SpMM's epilogue (F.relu, TrainingNeural.py:81)
        h = self.conv1(g, inputs, fuse_relu=True)
        if self.dropout_frac > 0.0 and self.training:
            # dropout defaults to 0.0 (TrainingConfig.dropout, :43) -> identity on the hot path
            h = torch.nn.functional.dropout(h, p=self.dropout_frac, training=True)
        z = self.conv2(g, h)
        return _SoftmaxFn.apply(z)
