"""Thin torch-tensor wrappers over the C ABI (one gmc_* call each, on torch's current stream).

torch is used for device memory and streams only; no torch operator computes anything on
the hot path.  Every wrapper validates dtype/device/contiguity and raises -- nothing here
falls back to a PyTorch implementation.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    return t


def _rowmajor(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """Accept 2-D tensors whose rows are contiguous (stride(1) == 1); return (tensor, ld)."""
    _f32(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D tensor")
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    if ld < t.shape[1]:
        t = t.contiguous()
        ld = t.shape[1]
    return t, ld


class Workspace:
    """Grow-only device scratch (the C ABI never allocates)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.generation = 0          # bumped on every reallocation (captured CUDA graphs hold the old address)

    def get(self, nbytes: int, device) -> Tuple[Optional[int], int]:
        if nbytes == 0:
            return None, 0
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self.generation += 1
        return self.buf.data_ptr(), self.buf.numel()


_default_ws = Workspace()
# GMC_SLAB_FUSED=1 sends the forward aggregation + projection through the slab kernel with the projection in its
# epilogue (gmc_spmm_batched_fused_skinny_f32).  Measured at config 3: 6.13 ms against 5.29 ms for the warp-per-row
# fused kernel -- the 12 shuffles + 12 FMAs per row and slab cost more than the L2 traffic they save -- so it is opt-in.
_SLAB_FUSED = os.environ.get("GMC_SLAB_FUSED", "0") == "1"


# ---------------------------------------------------------------- graph preparation
def degree_norm(rowptr: torch.Tensor, n_rows: int, count_zero: bool = True,
                out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[int]]:
    """(norm, zero-degree rows).  count_zero=False skips the count and its host read-back (no synchronisation: for
    callers that validated the degrees on the host already)."""
    norm = out if out is not None else torch.empty(n_rows, dtype=torch.float32, device=rowptr.device)
    if not count_zero:
        check(lib().gmc_degree_norm_f32(rowptr.data_ptr(), n_rows, norm.data_ptr(), None, _stream()), "gmc_degree_norm_f32")
        return norm, None
    zero = torch.zeros(1, dtype=torch.int32, device=rowptr.device)
    check(lib().gmc_degree_norm_f32(rowptr.data_ptr(), n_rows, norm.data_ptr(), zero.data_ptr(), _stream()),
          "gmc_degree_norm_f32")
    return norm, int(zero.item())


def edge_coef(rowptr, colidx, vals, norm_src, norm_dst, n_rows: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    coef = out if out is not None else torch.empty(colidx.numel(), dtype=torch.float32, device=colidx.device)
    check(lib().gmc_edge_coef_f32(rowptr.data_ptr(), colidx.data_ptr(), _ptr(vals), _ptr(norm_src), _ptr(norm_dst),
                                  n_rows, coef.data_ptr(), _stream()), "gmc_edge_coef_f32")
    return coef


def spmm_plan(rowptr, colidx, norm_src, norm_dst, graph_ptr, n_graphs: int, n_rows: int) -> Optional[torch.Tensor]:
    """ELL plan for the slab SpMM of A_hat = diag(norm_dst) A diag(norm_src) (None when some row has more
    than 8 neighbours, a graph has more than 65534 nodes, or a node's neighbours do not share one norm_src
    value -- i.e. the plan serves regular graphs)."""
    nbytes = lib().gmc_spmm_plan_bytes(n_rows)
    plan = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=colidx.device)
    over = torch.zeros(1, dtype=torch.int32, device=colidx.device)
    check(lib().gmc_spmm_plan_build(rowptr.data_ptr(), colidx.data_ptr(), norm_src.data_ptr(), norm_dst.data_ptr(),
                                    graph_ptr.data_ptr(),
                                    n_graphs, n_rows, plan.data_ptr(), over.data_ptr(), _stream()),
          "gmc_spmm_plan_build")
    return None if int(over.item()) else plan


def densify(batch, n_cols: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Padded adjacency rows X[N, n_cols] of a GraphBatch (device-side graphExtender)."""
    if out is None:
        out = torch.empty((batch.num_nodes, n_cols), dtype=torch.float32, device=batch.device)
    out, ld = _rowmajor(out, "out")
    check(lib().gmc_csr_densify_f32(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), _ptr(batch.wts_f32),
                                    batch.graph_ptr.data_ptr(), batch.num_graphs, batch.num_nodes, n_cols,
                                    out.data_ptr(), ld, _stream()), "gmc_csr_densify_f32")
    return out


def pad_cols(n: int, multiple: int = 32) -> int:
    """Leading dimension (floats) padded to 128 bytes."""
    return (n + multiple - 1) // multiple * multiple


def padded_empty(rows: int, cols: int, device, zero: bool = False) -> torch.Tensor:
    """[rows, cols] fp32 view whose row pitch is a multiple of 128 bytes."""
    make = torch.zeros if zero else torch.empty
    return make((rows, pad_cols(cols)), dtype=torch.float32, device=device)[:, :cols]


def copy2d(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    dst, lddst = _rowmajor(dst, "dst")
    src, ldsrc = _rowmajor(src, "src")
    if dst.shape != src.shape:
        raise ValueError("copy2d: shape mismatch")
    check(lib().gmc_copy2d_f32(dst.data_ptr(), lddst, src.data_ptr(), ldsrc, src.shape[0], src.shape[1], _stream()),
          "gmc_copy2d_f32")
    return dst


# ---------------------------------------------------------------- (a) SpMM
def spmm(batch, X: torch.Tensor, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
         relu: bool = False, use_coef: bool = True) -> torch.Tensor:
    """Y = act(A_hat X + bias).  use_coef=False recomputes the norms per edge from batch.norm."""
    X, ldx = _rowmajor(X, "X")
    n, c = X.shape
    if n != batch.num_nodes:
        raise ValueError(f"X has {n} rows, batch has {batch.num_nodes} nodes")
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=X.device)
    out, ldy = _rowmajor(out, "out")
    if bias is not None:
        _f32(bias, "bias")
    if use_coef:
        # block-diagonal fast path: shared-memory staged slabs when the graphs fit, else warp-per-row
        check(lib().gmc_spmm_batched_f32(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), batch.coef.data_ptr(),
                                         batch.graph_ptr.data_ptr(), batch.num_graphs, batch.max_nodes,
                                         _ptr(getattr(batch, "plan", None)), X.data_ptr(), out.data_ptr(), n, c,
                                         ldx, ldy, _ptr(bias), int(relu), _stream()),
              "gmc_spmm_batched_f32")
        return out
    check(lib().gmc_spmm_symnorm_f32(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), None, batch.norm.data_ptr(),
                                     batch.norm.data_ptr(), X.data_ptr(), out.data_ptr(), n, c, ldx, ldy, _ptr(bias),
                                     int(relu), _stream()), "gmc_spmm_symnorm_f32")
    return out


def spmm_fused_skinny(batch, X: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor] = None,
                      proj: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, relu: bool = True,
                      workspace: Optional[Workspace] = None):
    """(Y, T) with Y = act(A_hat X + bias) and T = Y W in one pass (n_cols <= 512, W [n_cols, n_out<=8]).  Batches
    with an ELL plan and n_out <= 4 take the TMA-staged slab kernel with the projection in its epilogue."""
    X, ldx = _rowmajor(X, "X")
    W = _f32(W, "W").contiguous()
    n, c = X.shape
    n_out = W.shape[1]
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=X.device)
    out, ldy = _rowmajor(out, "out")
    if proj is None:
        proj = torch.empty((n, n_out), dtype=torch.float32, device=X.device)
    proj, ldt = _rowmajor(proj, "proj")
    if (getattr(batch, "plan", None) is not None and n_out <= 4 and batch.max_nodes >= 128 and c >= 16 and c % 4 == 0
            and ldx % 4 == 0 and ldy % 4 == 0 and _SLAB_FUSED):
        ws = workspace or _default_ws
        wptr, wbytes = ws.get(lib().gmc_spmm_batched_fused_workspace_bytes(n, c), X.device)
        rc = lib().gmc_spmm_batched_fused_skinny_f32(batch.graph_ptr.data_ptr(), batch.num_graphs, batch.max_nodes,
                                                     batch.plan.data_ptr(), X.data_ptr(), out.data_ptr(), n, c, ldx, ldy,
                                                     _ptr(bias), int(relu), W.data_ptr(), n_out, proj.data_ptr(), ldt,
                                                     wptr, wbytes, _stream())
        if rc == 0:
            return out, proj
        if rc != -2:                                   # GMC_ERR_UNSUPPORTED: fall through to the row kernel
            check(rc, "gmc_spmm_batched_fused_skinny_f32")
    check(lib().gmc_spmm_fused_skinny_f32(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), batch.coef.data_ptr(), None,
                                          None, X.data_ptr(), out.data_ptr(), n, c, ldx, ldy, _ptr(bias), int(relu),
                                          W.data_ptr(), n_out, proj.data_ptr(), ldt, _stream()),
          "gmc_spmm_fused_skinny_f32")
    return out, proj


# ---------------------------------------------------------------- (b) GEMM
_OPS = {"nn": 0, "nt": 1, "tn": 2}


def gemm(op: str, A: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False,
         precision: str = "fp32", workspace: Optional[Workspace] = None) -> torch.Tensor:
    """nn: A[M,K] B[K,N];  nt: A[M,K] B[N,K];  tn: A[K,M] B[K,N]  ->  C[M,N]"""
    A, lda = _rowmajor(A, "A")
    B, ldb = _rowmajor(B, "B")
    if op == "nn":
        M, K = A.shape; K2, N = B.shape
    elif op == "nt":
        M, K = A.shape; N, K2 = B.shape
    elif op == "tn":
        K, M = A.shape; K2, N = B.shape
    else:
        raise ValueError(f"unknown gemm op {op!r}")
    if K != K2:
        raise ValueError(f"gemm {op}: inner dimensions differ ({K} vs {K2})")
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an output tensor")
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    out, ldc = _rowmajor(out, "out")
    prec = _lib.PRECISIONS[precision]
    ws = workspace or _default_ws
    need = lib().gmc_gemm_workspace_bytes(_OPS[op], M, N, K, prec)
    wptr, wbytes = ws.get(need, A.device)
    fn = getattr(lib(), "gmc_gemm_" + op)
    check(fn(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldc, int(accumulate), prec, wptr, wbytes,
             _stream()), "gmc_gemm_" + op)
    return out


# ---------------------------------------------------------------- bf16 operands
def _bf16_rowmajor(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    if t.dtype != torch.bfloat16 or not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
        raise TypeError(f"{name}: expected a row-major CUDA bfloat16 matrix")
    return t, t.stride(0)


def _b16_rowmajor(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """bf16 or IEEE fp16 (the operand format of the 'f16x2' GEMMs)"""
    if t.dtype not in (torch.bfloat16, torch.float16) or not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
        raise TypeError(f"{name}: expected a row-major CUDA bfloat16 / float16 matrix")
    return t, t.stride(0)


def padded_empty_bf16(rows: int, cols: int, device, zero: bool = False, dtype=torch.bfloat16) -> torch.Tensor:
    """[rows, cols] bf16 (or fp16) view whose row pitch is a multiple of 128 bytes (64 elements)."""
    make = torch.zeros if zero else torch.empty
    return make((rows, (cols + 63) // 64 * 64), dtype=dtype, device=device)[:, :cols]


def to_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 copy (round to nearest even) of a row-major fp32 matrix, 128-byte row pitch."""
    src, lds = _rowmajor(src, "src")
    if out is None:
        out = padded_empty_bf16(src.shape[0], src.shape[1], src.device)
    out, ldd = _bf16_rowmajor(out, "out")
    check(lib().gmc_f32_to_bf16(src.data_ptr(), lds, out.data_ptr(), ldd, src.shape[0], src.shape[1], _stream()),
          "gmc_f32_to_bf16")
    return out


def densify_bf16(batch, n_cols: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 twin of densify: padded adjacency rows X[N, n_cols] (0/1 entries are exact in bf16)."""
    if out is None:
        out = padded_empty_bf16(batch.num_nodes, n_cols, batch.device)
    out, ld = _bf16_rowmajor(out, "out")
    check(lib().gmc_csr_densify_bf16(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), _ptr(batch.wts_f32),
                                     batch.graph_ptr.data_ptr(), batch.num_graphs, batch.num_nodes, n_cols,
                                     out.data_ptr(), ld, _stream()), "gmc_csr_densify_bf16")
    return out


def scatter_features_bf16(batch, n_cols: int, X: torch.Tensor, clear: bool = False) -> torch.Tensor:
    """Incremental densify_bf16 on a reused buffer: write the batch's adjacency entries into an all-zero X
    (clear=False) or zero exactly those entries again (clear=True).  No memset of the [N, n_cols] matrix."""
    X, ld = _bf16_rowmajor(X, "X")
    if X.shape[0] != batch.num_nodes or X.shape[1] != n_cols:
        raise ValueError(f"X must be [{batch.num_nodes}, {n_cols}]")
    check(lib().gmc_csr_scatter_bf16(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), _ptr(batch.wts_f32),
                                     batch.graph_ptr.data_ptr(), batch.num_graphs, batch.num_nodes, n_cols,
                                     X.data_ptr(), ld, int(clear), _stream()), "gmc_csr_scatter_bf16")
    return X


class PreaggregatedFeatures:
    """Marks a bf16 matrix as XA = A_hat X (already aggregated over the graph) for GCNEngine(preaggregate=True)."""
    __slots__ = ("tensor",)

    def __init__(self, tensor: torch.Tensor):
        _bf16_rowmajor(tensor, "XA")
        self.tensor = tensor


def preaggregate_features_bf16(batch, n_cols: int, out: Optional[torch.Tensor] = None,
                               workspace: Optional[Workspace] = None, counting: bool = False) -> torch.Tensor:
    """XA = A_hat X in bf16 for X = the zero-padded adjacency rows of the batch, straight from the graph (the dense X
    is never formed)."""
    if out is None:
        out = padded_empty_bf16(batch.num_nodes, n_cols, batch.device, zero=True)
    out, ld = _bf16_rowmajor(out, "out")
    if out.shape[0] != batch.num_nodes or out.shape[1] != n_cols:
        raise ValueError(f"out must be [{batch.num_nodes}, {n_cols}]")
    wptr, wbytes = None, 0
    if counting:        # byte-counting kernel for the regular rows + masked general kernel (no faster at config 3: opt-in)
        ws = workspace or _default_ws
        wptr, wbytes = ws.get(lib().gmc_csr_preaggregate_workspace_bytes(batch.num_nodes), batch.device)
    if not counting:
        check(lib().gmc_csr_preaggregate_graphs(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), batch.coef.data_ptr(),
                                                _ptr(batch.wts_f32), batch.graph_ptr.data_ptr(), batch.num_graphs,
                                                int(batch.max_nodes), batch.num_nodes, n_cols, out.data_ptr(), ld, 0,
                                                _stream()), "gmc_csr_preaggregate_graphs")
        return out
    check(lib().gmc_csr_preaggregate_bf16(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), batch.coef.data_ptr(),
                                          _ptr(batch.wts_f32), batch.graph_ptr.data_ptr(), batch.num_graphs,
                                          batch.num_nodes, n_cols, out.data_ptr(), ld, wptr, wbytes, _stream()),
          "gmc_csr_preaggregate_bf16")
    return out


def spmm_bf16out(batch, X: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A_hat X rounded to bf16 (slab kernel; falls back to the fp32 SpMM + conversion when the batch has no plan)."""
    X, ldx = _rowmajor(X, "X")
    n, c = X.shape
    if out is None:
        out = padded_empty_bf16(n, c, X.device)
    out, ldy = _bf16_rowmajor(out, "out")
    if getattr(batch, "plan", None) is not None and batch.max_nodes >= 128 and c >= 16 and c % 4 == 0:
        check(lib().gmc_spmm_batched_bf16out(batch.graph_ptr.data_ptr(), batch.num_graphs, batch.max_nodes,
                                             batch.plan.data_ptr(), X.data_ptr(), out.data_ptr(), n, c, ldx, ldy,
                                             _stream()), "gmc_spmm_batched_bf16out")
        return out
    return to_bf16(spmm(batch, X), out=out)


def gemm_bf16(op: str, A: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False,
              workspace: Optional[Workspace] = None) -> torch.Tensor:
    """gemm with bf16 operands and fp32 output.  nn: A[M,K] B[K,N];  nt: A[M,K] B[N,K];  tn: A[K,M] B[K,N]."""
    A, lda = _bf16_rowmajor(A, "A")
    B, ldb = _bf16_rowmajor(B, "B")
    if op == "nn":
        M, K = A.shape; K2, N = B.shape
    elif op == "nt":
        M, K = A.shape; N, K2 = B.shape
    elif op == "tn":
        K, M = A.shape; K2, N = B.shape
    else:
        raise ValueError(f"unknown gemm op {op!r}")
    if K != K2:
        raise ValueError(f"gemm_bf16 {op}: inner dimensions differ ({K} vs {K2})")
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an output tensor")
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    out, ldc = _rowmajor(out, "out")
    ws = workspace or _default_ws
    wptr, wbytes = ws.get(lib().gmc_gemm_bf16_workspace_bytes(_OPS[op], M, N, K), A.device)
    check(lib().gmc_gemm_bf16(_OPS[op], A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldc,
                              int(accumulate), wptr, wbytes, _stream()), "gmc_gemm_bf16")
    return out


# ---------------------------------------------------------------- fp32-grade split-operand path (csrc/split.cu)
def split_rows_for(k: int) -> int:
    """Row stride between the stacked bf16 parts of a K-row operand (K rounded up to the 64-row k-block)."""
    return (int(k) + 63) // 64 * 64


def split_empty(rows: int, cols: int, n_split: int, device) -> torch.Tensor:
    """Zeroed [n_split * split_rows_for(rows), cols] bf16 buffer (128-byte row pitch) for a split operand."""
    return padded_empty_bf16(n_split * split_rows_for(rows), cols, device, zero=True)


def f32_split_bf16(src: torch.Tensor, n_split: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [K, N] -> n_split stacked bf16 parts hi, lo [, lo2] with src = sum of the parts up to 2^-(8 n_split + 1)
    relative; part p at rows [p * split_rows_for(K), ...); the pad rows of every part are written as zeros."""
    src, lds = _rowmajor(src, "src")
    k, n = src.shape
    sr = split_rows_for(k)
    if out is None:
        out = split_empty(k, n, n_split, src.device)
    out, ldd = _bf16_rowmajor(out, "out")
    if out.shape[0] != n_split * sr or out.shape[1] != n:
        raise ValueError(f"out must be [{n_split * sr}, {n}]")
    check(lib().gmc_f32_split_bf16(src.data_ptr(), lds, out.data_ptr(), ldd, k, n, n_split, sr, _stream()),
          "gmc_f32_split_bf16")
    return out


F16_LO_SHIFT = 12       # fp16 parts: part p is stored multiplied by 2^(12 p)


def f32_split_f16(src: torch.Tensor, n_split: int = 2, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [K, N] -> n_split stacked fp16 parts (11 significant bits each; part p scaled by 2^(12 p)), held in a
    2-byte tensor of dtype float16; same stacking as f32_split_bf16."""
    src, lds = _rowmajor(src, "src")
    k, n = src.shape
    sr = split_rows_for(k)
    if out is None:
        out = torch.zeros((n_split * sr, (n + 63) // 64 * 64), dtype=torch.float16, device=src.device)[:, :n]
    if out.dtype != torch.float16 or not out.is_cuda or out.stride(1) != 1 or out.shape != (n_split * sr, n):
        raise ValueError(f"out must be a row-major CUDA float16 [{n_split * sr}, {n}] matrix")
    check(lib().gmc_f32_split_f16(src.data_ptr(), lds, out.data_ptr(), out.stride(0), k, n, n_split, sr, F16_LO_SHIFT,
                                  _stream()), "gmc_f32_split_f16")
    return out


def row_scale(batch, count_nonuniform: bool = True, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[int]]:
    """(s, nonuniform): s[v] = the A_hat coefficient of row v's edges, nonuniform = rows whose edges differ (such a
    batch cannot take the integer-feature path).  count_nonuniform=False skips the count and its host read-back."""
    s = out if out is not None else torch.empty(batch.num_nodes, dtype=torch.float32, device=batch.device)
    if not count_nonuniform:
        check(lib().gmc_row_scale_f32(batch.rowptr.data_ptr(), batch.coef.data_ptr(), batch.num_nodes, s.data_ptr(),
                                      None, _stream()), "gmc_row_scale_f32")
        return s, None
    bad = torch.zeros(1, dtype=torch.int32, device=batch.device)
    check(lib().gmc_row_scale_f32(batch.rowptr.data_ptr(), batch.coef.data_ptr(), batch.num_nodes, s.data_ptr(),
                                  bad.data_ptr(), _stream()), "gmc_row_scale_f32")
    return s, int(bad.item())


def integer_features_bf16(batch, n_cols: int, out: Optional[torch.Tensor] = None,
                          ones: Optional[torch.Tensor] = None, f16: bool = False) -> torch.Tensor:
    """XI = sum over neighbours of the zero-padded 0/1 adjacency rows (2-step path counts): (A_hat X) = s . XI for a
    batch with one coefficient per row; small integers, exact in bf16 and in fp16 (f16=True, or an fp16 `out`: the
    operand of the 'f16x2' GEMMs).  Unit edge weights only.  `ones` is an optional caller-owned fp32 vector of >= nnz ones
    (streamed batches reuse one instead of allocating per step)."""
    if not batch.unit_weights:
        raise ValueError("integer features need unit edge weights")
    if ones is None:
        ones = getattr(batch, "_unit_coef", None)
        if ones is None or ones.numel() != batch.nnz:
            ones = torch.ones(batch.nnz, dtype=torch.float32, device=batch.device)
            batch._unit_coef = ones
    elif ones.numel() < batch.nnz or ones.dtype != torch.float32:
        raise ValueError("ones must be an fp32 vector with at least nnz entries")
    if out is None:
        out = padded_empty_bf16(batch.num_nodes, n_cols, batch.device, zero=True,
                                dtype=torch.float16 if f16 else torch.bfloat16)
    out, ld = _b16_rowmajor(out, "out")
    check(lib().gmc_csr_preaggregate_graphs(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), ones.data_ptr(), None,
                                            batch.graph_ptr.data_ptr(), batch.num_graphs, int(batch.max_nodes),
                                            batch.num_nodes, n_cols, out.data_ptr(), ld,
                                            int(out.dtype == torch.float16), _stream()), "gmc_csr_preaggregate_graphs")
    return out


class IntegerFeatures:
    """XI (bf16 integers; fp16 for 'f16x2') + the per-row scale s with A_hat X = s . XI, for
    GCNEngine(precision='bf16x2' | 'bf16x3' | 'f16x2')."""
    __slots__ = ("tensor", "scale")

    def __init__(self, tensor: torch.Tensor, scale: torch.Tensor):
        _b16_rowmajor(tensor, "XI")
        if scale.dtype != torch.float32 or scale.numel() != tensor.shape[0] or not scale.is_contiguous():
            raise ValueError("scale must be a contiguous fp32 vector with one entry per row")
        self.tensor, self.scale = tensor, scale

    @classmethod
    def from_batch(cls, batch, n_cols: int, f16: bool = False) -> "IntegerFeatures":
        s, bad = row_scale(batch)
        if bad:
            raise ValueError(f"{bad} rows have neighbours of different degrees: no single A_hat coefficient per row")
        return cls(integer_features_bf16(batch, n_cols, f16=f16), s)


def gemm_bf16_split(op: str, A: torch.Tensor, B_split: torch.Tensor, n_split: int, k: int,
                    out: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
                    bias: Optional[torch.Tensor] = None, relu: bool = False, accumulate: bool = False,
                    workspace: Optional[Workspace] = None, proj_w: Optional[torch.Tensor] = None,
                    proj_out: Optional[torch.Tensor] = None, n_proj: int = 0,
                    proj_parts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = op(A) (B_0 + ... + B_{n_split-1}) with bf16 A (exact) and the stacked bf16 parts of an fp32 B [k, N]
    (f32_split_bf16 / skinny_bwd_split); fp32 accumulation and output.  op 'nn': A [M, k]; 'tn': A [k, M].
    fp16 parts (f32_split_f16, or skinny_bwd_split into an fp16 buffer) need an fp16 A: tcgen05.mma kind::f16 takes ONE
    16-bit format for both operands.
    Optional fp32 epilogue C[m, n] = act(row_scale[m] * acc + bias[n]) and fused projection proj_out = C @ W for a
    padded weight matrix from pad_proj_weights (deterministic: per-tile partials in the workspace, added in order)."""
    A, lda = _b16_rowmajor(A, "A")
    B_split, ldb = _b16_rowmajor(B_split, "B_split")
    if A.dtype != B_split.dtype:
        raise TypeError(f"gemm_bf16_split: A is {A.dtype} and the parts are {B_split.dtype}; both operands of one "
                        "tcgen05.mma must share a 16-bit format")
    lo_shift = F16_LO_SHIFT if B_split.dtype == torch.float16 else 0
    if op == "nn":
        M, K = A.shape
    elif op == "tn":
        K, M = A.shape
    else:
        raise ValueError("gemm_bf16_split: op must be 'nn' or 'tn'")
    if K != k:
        raise ValueError(f"gemm_bf16_split {op}: inner dimensions differ ({K} vs {k})")
    sr = split_rows_for(k)
    if B_split.shape[0] < n_split * sr:
        raise ValueError(f"B_split must hold {n_split} parts of {sr} rows")
    N = B_split.shape[1]
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an output tensor")
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    out, ldc = _rowmajor(out, "out")
    if row_scale is not None and (_f32(row_scale, "row_scale").numel() != M or not row_scale.is_contiguous()):
        raise ValueError(f"row_scale must be a contiguous fp32 vector of {M} elements")
    if bias is not None and (_f32(bias, "bias").numel() != N or not bias.is_contiguous()):
        raise ValueError(f"bias must be a contiguous fp32 vector of {N} elements")
    ldp = 0
    if proj_w is not None:
        rows = (N + 63) // 64 * 64
        if proj_w.shape != (rows, 4) or proj_w.dtype != torch.float32 or not proj_w.is_contiguous():
            raise ValueError(f"proj_w must come from pad_proj_weights: contiguous fp32 [{rows}, 4]")
        if not 1 <= n_proj <= 4:
            raise ValueError("the fused projection needs 1 <= n_proj <= 4")
        if proj_parts is not None:
            # deferred: the per-n-tile partials stay in `proj_parts` (fp32 [split_proj_tiles(N, n_split) * M, 4]) for
            # layer2_loss_fused(t2_parts=...) -- no reduce launch, no proj_out
            tiles = split_proj_tiles(N, n_split)
            if (proj_out is not None or proj_parts.dtype != torch.float32 or not proj_parts.is_contiguous()
                    or proj_parts.numel() < tiles * M * 4):
                raise ValueError(f"proj_parts must be a contiguous fp32 buffer of {tiles} x {M} float4 (and proj_out None)")
        else:
            if proj_out is None:
                raise ValueError("the fused projection needs proj_out (or proj_parts)")
            proj_out, ldp = _rowmajor(proj_out, "proj_out")
            if proj_out.shape[0] != M or proj_out.shape[1] < n_proj:
                raise ValueError(f"proj_out must be [{M}, >= {n_proj}]")
    if proj_parts is not None and proj_w is not None:
        wptr, wbytes = proj_parts.data_ptr(), proj_parts.numel() * 4
    else:
        ws = workspace or _default_ws
        wptr, wbytes = ws.get(lib().gmc_gemm_bf16_split_workspace_bytes(_OPS[op], M, N, K, n_split, int(proj_w is not None)),
                              A.device)
    check(lib().gmc_gemm_bf16_split(_OPS[op], A.data_ptr(), B_split.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldc,
                                    n_split, sr, _ptr(row_scale), _ptr(bias), int(relu), _ptr(proj_w), _ptr(proj_out), ldp,
                                    int(n_proj), int(accumulate), lo_shift, wptr, wbytes, _stream()), "gmc_gemm_bf16_split")
    return out


def split_proj_tiles(n_cols: int, n_split: int) -> int:
    """n-tiles of a split-operand GEMM = projection partials per row (128 real columns per tile for two parts, 64 for three)."""
    rn = 128 if n_split == 2 else 64
    return (int(n_cols) + rn - 1) // rn


def skinny_bwd_split(dT: torch.Tensor, W: torch.Tensor, H: torch.Tensor, n_split: int, dH_split: torch.Tensor,
                     row_scale: Optional[torch.Tensor] = None, dW: Optional[torch.Tensor] = None,
                     dbias: Optional[torch.Tensor] = None, workspace: Optional[Workspace] = None):
    """skinny_bwd with an fp32 H whose dHpre leaves as n_split stacked bf16 parts of row_scale . dHpre -- or fp16 parts
    when dH_split is a float16 buffer (part p scaled by 2^(F16_LO_SHIFT p), values saturate at +-65504)."""
    dT, lddt = _rowmajor(dT, "dT")
    H, ldh = _rowmajor(H, "H")
    W = _f32(W, "W").contiguous()
    n, n_in = H.shape
    n_out = W.shape[1]
    dH_split, lddh = _b16_rowmajor(dH_split, "dH_split")
    lo_shift = F16_LO_SHIFT if dH_split.dtype == torch.float16 else 0
    sr = split_rows_for(n)
    if dH_split.shape[0] < n_split * sr or dH_split.shape[1] != n_in:
        raise ValueError(f"dH_split must be [>= {n_split * sr}, {n_in}]")
    if dW is None:
        dW = torch.empty((n_in, n_out), dtype=torch.float32, device=H.device)
    if dbias is None:
        dbias = torch.empty(n_in, dtype=torch.float32, device=H.device)
    ws = workspace or _default_ws
    wptr, wbytes = ws.get(lib().gmc_skinny_bwd_workspace_bytes(n_in, n_out), H.device)
    check(lib().gmc_skinny_bwd_split(dT.data_ptr(), lddt, W.data_ptr(), H.data_ptr(), ldh, _ptr(row_scale),
                                     dH_split.data_ptr(), lddh, sr, n_split, lo_shift, dW.data_ptr(), dbias.data_ptr(), n, n_in,
                                     n_out, wptr, wbytes, _stream()), "gmc_skinny_bwd_split")
    return dH_split, dW, dbias


# ---------------------------------------------------------------- bf16 layer-1 activations
def pad_proj_weights(W: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n_in, n_out <= 4] fp32 -> zero-padded [n_in rounded up to 64, 4] (the layout of the fused GEMM projection)."""
    W = _f32(W, "W")
    n_in, n_out = W.shape
    if n_out > 4:
        raise ValueError("the fused projection takes at most 4 output columns")
    rows = (n_in + 63) // 64 * 64
    if out is None:
        out = torch.zeros((rows, 4), dtype=torch.float32, device=W.device)
    if out.shape != (rows, 4) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous fp32 [{rows}, 4] tensor")
    copy2d(out[:n_in, :n_out], W.contiguous())
    return out


def gemm_bf16_bf16out(op: str, A: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None,
                      bias: Optional[torch.Tensor] = None, relu: bool = False, proj_w: Optional[torch.Tensor] = None,
                      proj_out: Optional[torch.Tensor] = None, n_proj: int = 0) -> torch.Tensor:
    """gemm_bf16 whose result is rounded to bf16 in the epilogue (no split-K, no accumulate); optional fp32 bias
    over the columns and ReLU before the rounding; optional fused projection proj_out = bf16(C) @ W for a padded weight
    matrix from pad_proj_weights (proj_out [M, >= n_proj] fp32 is overwritten)."""
    A, lda = _bf16_rowmajor(A, "A")
    B, ldb = _bf16_rowmajor(B, "B")
    if op == "nn":
        M, K = A.shape; K2, N = B.shape
    elif op == "nt":
        M, K = A.shape; N, K2 = B.shape
    elif op == "tn":
        K, M = A.shape; K2, N = B.shape
    else:
        raise ValueError(f"unknown gemm op {op!r}")
    if K != K2:
        raise ValueError(f"gemm_bf16_bf16out {op}: inner dimensions differ ({K} vs {K2})")
    if out is None:
        out = padded_empty_bf16(M, N, A.device, zero=True)
    out, ldc = _bf16_rowmajor(out, "out")
    if bias is not None and (_f32(bias, "bias").numel() != N or not bias.is_contiguous()):
        raise ValueError(f"bias must be a contiguous fp32 vector of {N} elements")
    ldp = 0
    if proj_w is not None:
        rows = (N + 63) // 64 * 64
        if proj_w.shape != (rows, 4) or proj_w.dtype != torch.float32 or not proj_w.is_contiguous():
            raise ValueError(f"proj_w must come from pad_proj_weights: contiguous fp32 [{rows}, 4]")
        if proj_out is None or not 1 <= n_proj <= 4:
            raise ValueError("the fused projection needs proj_out and 1 <= n_proj <= 4")
        proj_out, ldp = _rowmajor(proj_out, "proj_out")
        if proj_out.shape[0] != M or proj_out.shape[1] < n_proj:
            raise ValueError(f"proj_out must be [{M}, >= {n_proj}]")
    check(lib().gmc_gemm_bf16_bf16out(_OPS[op], A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldc,
                                      _ptr(bias), int(relu), _ptr(proj_w), _ptr(proj_out), ldp, int(n_proj), _stream()),
          "gmc_gemm_bf16_bf16out")
    return out


def spmm_fused_skinny_bf16(batch, X: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor] = None,
                           proj: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                           relu: bool = True, out_bf16: bool = True):
    """spmm_fused_skinny with a bf16 X; Y is bf16 (default) or fp32.  Returns (Y, T)."""
    X, ldx = _bf16_rowmajor(X, "X")
    W = _f32(W, "W").contiguous()
    n, c = X.shape
    n_out = W.shape[1]
    if out is None:
        out = padded_empty_bf16(n, c, X.device, zero=True) if out_bf16 else padded_empty(n, c, X.device)
    if out.dtype == torch.bfloat16:
        out, ldy = _bf16_rowmajor(out, "out")
        y_bf16 = 1
    else:
        out, ldy = _rowmajor(out, "out")
        y_bf16 = 0
    if proj is None:
        proj = torch.empty((n, n_out), dtype=torch.float32, device=X.device)
    proj, ldt = _rowmajor(proj, "proj")
    if bias is not None:
        _f32(bias, "bias")
    check(lib().gmc_spmm_fused_skinny_bf16(batch.rowptr.data_ptr(), batch.colidx.data_ptr(), batch.coef.data_ptr(),
                                           None, None, X.data_ptr(), out.data_ptr(), y_bf16, n, c, ldx, ldy,
                                           _ptr(bias), int(relu), W.data_ptr(), n_out, proj.data_ptr(), ldt,
                                           _stream()), "gmc_spmm_fused_skinny_bf16")
    return out, proj


def skinny_bwd_bf16(dT: torch.Tensor, W: torch.Tensor, H: torch.Tensor, dH: Optional[torch.Tensor] = None,
                    dW: Optional[torch.Tensor] = None, dbias: Optional[torch.Tensor] = None,
                    workspace: Optional[Workspace] = None):
    """skinny_bwd with bf16 H and dH."""
    dT, lddt = _rowmajor(dT, "dT")
    H, ldh = _bf16_rowmajor(H, "H")
    W = _f32(W, "W").contiguous()
    n, n_in = H.shape
    n_out = W.shape[1]
    if dH is None:
        dH = padded_empty_bf16(n, n_in, H.device, zero=True)
    dH, lddh = _bf16_rowmajor(dH, "dH")
    if dW is None:
        dW = torch.empty((n_in, n_out), dtype=torch.float32, device=H.device)
    if dbias is None:
        dbias = torch.empty(n_in, dtype=torch.float32, device=H.device)
    ws = workspace or _default_ws
    wptr, wbytes = ws.get(lib().gmc_skinny_bwd_workspace_bytes(n_in, n_out), H.device)
    check(lib().gmc_skinny_bwd_bf16(dT.data_ptr(), lddt, W.data_ptr(), H.data_ptr(), ldh, dH.data_ptr(), lddh,
                                    dW.data_ptr(), dbias.data_ptr(), n, n_in, n_out, wptr, wbytes, _stream()),
          "gmc_skinny_bwd_bf16")
    return dH, dW, dbias


def spmm_bf16(batch, X: torch.Tensor, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
              relu: bool = False) -> torch.Tensor:
    """act(A_hat X + bias) with bf16 X and bf16 result (slab kernel; the batch needs an ELL plan)."""
    X, ldx = _bf16_rowmajor(X, "X")
    n, c = X.shape
    if n != batch.num_nodes:
        raise ValueError(f"X has {n} rows, batch has {batch.num_nodes} nodes")
    if out is None:
        out = padded_empty_bf16(n, c, X.device, zero=True)
    out, ldy = _bf16_rowmajor(out, "out")
    check(lib().gmc_spmm_batched_bf16(batch.graph_ptr.data_ptr(), batch.num_graphs, batch.max_nodes,
                                      _ptr(getattr(batch, "plan", None)), X.data_ptr(), out.data_ptr(), n, c, ldx, ldy,
                                      _ptr(_f32(bias, "bias")) if bias is not None else None, int(relu), _stream()),
          "gmc_spmm_batched_bf16")
    return out


def skinny_fwd_bf16(H: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """T = H W with a bf16 H [n, n_in <= 512] and fp32 W [n_in, n_out <= 4]."""
    H, ldh = _bf16_rowmajor(H, "H")
    W = _f32(W, "W").contiguous()
    n, n_in = H.shape
    n_out = W.shape[1]
    if out is None:
        out = torch.empty((n, n_out), dtype=torch.float32, device=H.device)
    out, ldt = _rowmajor(out, "out")
    check(lib().gmc_skinny_fwd_bf16(H.data_ptr(), ldh, W.data_ptr(), out.data_ptr(), ldt, n, n_in, n_out, _stream()),
          "gmc_skinny_fwd_bf16")
    return out


def bf16_activations_apply(batch, hidden: int) -> bool:
    """True when a batch can run the bf16-activation step: ELL plan (regular graphs, degree <= 8, 128..3600 nodes per graph)."""
    return (getattr(batch, "plan", None) is not None and 128 <= batch.max_nodes <= 3600 and 16 <= hidden <= 512
            and hidden % 4 == 0)


def adjacency_kernels_apply(batch, n_w_rows: int) -> bool:
    """Can X W / X^T dT for X = zero-padded unit-weight adjacency rows be computed as aggregations for this batch?"""
    return (getattr(batch, "plan", None) is not None and bool(getattr(batch, "unit_weights", False))
            and batch.max_nodes <= min(n_w_rows, 1024) and n_w_rows <= 1030)


def adj_features_fwd(batch, W: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """T = X W for X = the batch's zero-padded adjacency rows [N, W.shape[0]], without forming X (row gather of W)."""
    W, ldw = _rowmajor(W, "W")
    n_w_rows, c = W.shape
    if out is None:
        out = torch.empty((batch.num_nodes, c), dtype=torch.float32, device=W.device)
    out, ldt = _rowmajor(out, "out")
    check(lib().gmc_adj_features_fwd_f32(batch.plan.data_ptr(), batch.graph_ptr.data_ptr(), batch.num_graphs,
                                         batch.max_nodes, W.data_ptr(), ldw, n_w_rows, out.data_ptr(), ldt,
                                         batch.num_nodes, c, _stream()), "gmc_adj_features_fwd_f32")
    return out


def adj_features_bwd(batch, dT: torch.Tensor, out: torch.Tensor, workspace: Optional[Workspace] = None) -> torch.Tensor:
    """out = X^T dT for the same X ([out.shape[0], dT.shape[1]]); deterministic."""
    dT, lddt = _rowmajor(dT, "dT")
    out, lddw = _rowmajor(out, "out")
    n_w_rows, c = out.shape
    if dT.shape != (batch.num_nodes, c):
        raise ValueError("adj_features_bwd: dT must be [num_nodes, out.shape[1]]")
    ws = workspace or _default_ws
    wptr, wbytes = ws.get(lib().gmc_adj_features_bwd_workspace_bytes(n_w_rows, c), dT.device)
    check(lib().gmc_adj_features_bwd_f32(batch.plan.data_ptr(), batch.graph_ptr.data_ptr(), batch.num_graphs,
                                         batch.max_nodes, dT.data_ptr(), lddt, batch.num_nodes, c, out.data_ptr(), lddw,
                                         n_w_rows, wptr, wbytes, _stream()), "gmc_adj_features_bwd_f32")
    return out


def skinny_fwd(H: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    H, ldh = _rowmajor(H, "H")
    W = _f32(W, "W").contiguous()
    n, n_in = H.shape
    n_out = W.shape[1]
    if out is None:
        out = torch.empty((n, n_out), dtype=torch.float32, device=H.device)
    out, ldt = _rowmajor(out, "out")
    check(lib().gmc_skinny_fwd_f32(H.data_ptr(), ldh, W.data_ptr(), out.data_ptr(), ldt, n, n_in, n_out, _stream()),
          "gmc_skinny_fwd_f32")
    return out


def skinny_bwd(dT: torch.Tensor, W: torch.Tensor, H: torch.Tensor, dH: Optional[torch.Tensor] = None,
               dW: Optional[torch.Tensor] = None, dbias: Optional[torch.Tensor] = None,
               workspace: Optional[Workspace] = None):
    dT, lddt = _rowmajor(dT, "dT")
    H, ldh = _rowmajor(H, "H")
    W = _f32(W, "W").contiguous()
    n, n_in = H.shape
    n_out = W.shape[1]
    if dH is None:
        dH = torch.empty((n, n_in), dtype=torch.float32, device=H.device)
    dH, lddh = _rowmajor(dH, "dH")
    if dW is None:
        dW = torch.empty((n_in, n_out), dtype=torch.float32, device=H.device)
    if dbias is None:
        dbias = torch.empty(n_in, dtype=torch.float32, device=H.device)
    ws = workspace or _default_ws
    wptr, wbytes = ws.get(lib().gmc_skinny_bwd_workspace_bytes(n_in, n_out), H.device)
    check(lib().gmc_skinny_bwd_f32(dT.data_ptr(), lddt, W.data_ptr(), H.data_ptr(), ldh, dH.data_ptr(), lddh,
                                   dW.data_ptr(), dbias.data_ptr(), n, n_in, n_out, wptr, wbytes, _stream()),
          "gmc_skinny_bwd_f32")
    return dH, dW, dbias


def colsum(X: torch.Tensor, out: Optional[torch.Tensor] = None, workspace: Optional[Workspace] = None) -> torch.Tensor:
    X, ldx = _rowmajor(X, "X")
    n, c = X.shape
    if out is None:
        out = torch.empty(c, dtype=torch.float32, device=X.device)
    ws = workspace or _default_ws
    wptr, wbytes = ws.get(lib().gmc_colsum_workspace_bytes(c), X.device)
    check(lib().gmc_colsum_f32(X.data_ptr(), ldx, n, c, out.data_ptr(), wptr, wbytes, _stream()), "gmc_colsum_f32")
    return out


# ---------------------------------------------------------------- (c) fused loss
def cut_loss(batch, Z: torch.Tensor, mode: str = "ste", override_terminals: bool = True, penalty: float = 0.0,
             C: float = 1.0, need_P: bool = True, need_dZ: bool = True, P: Optional[torch.Tensor] = None,
             dZ: Optional[torch.Tensor] = None, loss: Optional[torch.Tensor] = None):
    """Returns (loss_per_graph float64 [B], P or None, dZ or None)."""
    Z, ldz = _rowmajor(Z, "Z")
    n, K = Z.shape
    if need_P and P is None:
        P = torch.empty((n, K), dtype=torch.float32, device=Z.device)
    if need_dZ and dZ is None:
        dZ = torch.empty((n, K), dtype=torch.float32, device=Z.device)
    if loss is None:
        loss = torch.empty(batch.num_graphs, dtype=torch.float64, device=Z.device)
    for t, nm in ((P, "P"), (dZ, "dZ")):
        if t is not None and not t.is_contiguous():
            raise ValueError(f"{nm} must be contiguous [N, K]")
    check(lib().gmc_softmax_cut_loss_fwd_bwd(Z.data_ptr(), ldz, batch.rowptr.data_ptr(), batch.colidx.data_ptr(),
                                             _ptr(batch.wts_f32), batch.graph_ptr.data_ptr(), batch.num_graphs, n, K,
                                             _lib.LOSS_MODES[mode], int(override_terminals), float(penalty), float(C),
                                             _ptr(P) if need_P else None, loss.data_ptr(),
                                             _ptr(dZ) if need_dZ else None, _stream()),
          "gmc_softmax_cut_loss_fwd_bwd")
    return loss, (P if need_P else None), (dZ if need_dZ else None)


def layer2_loss_fused_applies(batch, n_classes: int) -> bool:
    """One CTA per graph with 3 * n * K floats of shared memory: graphs of up to 6000 nodes at 3 classes."""
    return 0 < batch.max_nodes and 3 * batch.max_nodes * n_classes * 4 <= 216 * 1024


def layer2_loss_fused(batch, T2: torch.Tensor, bias2: Optional[torch.Tensor], mode: str = "ste", override_terminals: bool = True,
                      penalty: float = 0.0, C: float = 1.0, P: Optional[torch.Tensor] = None,
                      dZ: Optional[torch.Tensor] = None, loss: Optional[torch.Tensor] = None,
                      dT2: Optional[torch.Tensor] = None, db2: Optional[torch.Tensor] = None,
                      Z: Optional[torch.Tensor] = None, workspace: Optional[Workspace] = None,
                      t2_parts: Optional[Tuple[torch.Tensor, int]] = None) -> torch.Tensor:
    """Z = A_hat T2 + b2, softmax / override / STE / max-cut loss, dZ, db2 = colsum(dZ), dT2 = A_hat dZ in one launch
    (one CTA per graph).  P / dZ / Z (contiguous [N, K]), dT2 ([N, >= K]) and db2 ([K]) are optional outputs; returns the
    per-graph loss (float64 [B]).  t2_parts = (buffer, n_parts): T2 is READ from the projection partials a
    gemm_bf16_split(..., proj_parts=buffer) call left behind (added in tile order) and only written to the T2 argument."""
    T2, ldt = _rowmajor(T2, "T2")
    n, K = T2.shape
    if n != batch.num_nodes:
        raise ValueError(f"T2 has {n} rows, batch has {batch.num_nodes} nodes")
    if loss is None:
        loss = torch.empty(batch.num_graphs, dtype=torch.float64, device=T2.device)
    for t, nm in ((P, "P"), (dZ, "dZ"), (Z, "Z")):
        if t is not None and (not t.is_contiguous() or t.shape != (n, K) or t.dtype != torch.float32):
            raise ValueError(f"{nm} must be a contiguous fp32 [N, K] tensor")
    lddt = 0
    if dT2 is not None:
        dT2, lddt = _rowmajor(dT2, "dT2")
    wptr, wbytes = None, 0
    if db2 is not None:
        ws = workspace or _default_ws
        wptr, wbytes = ws.get(lib().gmc_layer2_loss_fused_workspace_bytes(batch.num_graphs, K), T2.device)
    if t2_parts is not None:
        parts, n_parts = t2_parts
        if K > 4 or parts.dtype != torch.float32 or not parts.is_contiguous() or parts.numel() < n_parts * n * 4:
            raise ValueError(f"t2_parts needs <= 4 classes and a contiguous fp32 buffer of {n_parts} x {n} float4")
        check(lib().gmc_layer2_loss_fused_parts(parts.data_ptr(), int(n_parts), n, T2.data_ptr(), ldt,
                                                batch.rowptr.data_ptr(), batch.colidx.data_ptr(),
                                                batch.coef.data_ptr(), _ptr(batch.wts_f32), batch.graph_ptr.data_ptr(),
                                                batch.num_graphs, batch.max_nodes, n, K, _ptr(bias2), _lib.LOSS_MODES[mode],
                                                int(override_terminals), float(penalty), float(C), _ptr(Z), _ptr(P),
                                                loss.data_ptr(), _ptr(dZ), _ptr(dT2), lddt, _ptr(db2), wptr, wbytes, _stream()),
              "gmc_layer2_loss_fused_parts")
        return loss
    check(lib().gmc_layer2_loss_fused(T2.data_ptr(), ldt, batch.rowptr.data_ptr(), batch.colidx.data_ptr(),
                                      batch.coef.data_ptr(), _ptr(batch.wts_f32), batch.graph_ptr.data_ptr(),
                                      batch.num_graphs, batch.max_nodes, n, K, _ptr(bias2), _lib.LOSS_MODES[mode],
                                      int(override_terminals), float(penalty), float(C), _ptr(Z), _ptr(P),
                                      loss.data_ptr(), _ptr(dZ), _ptr(dT2), lddt, _ptr(db2), wptr, wbytes, _stream()),
          "gmc_layer2_loss_fused")
    return loss


def softmax_fwd(Z: torch.Tensor) -> torch.Tensor:
    Z, ldz = _rowmajor(Z, "Z")
    n, K = Z.shape
    P = torch.empty((n, K), dtype=torch.float32, device=Z.device)
    check(lib().gmc_softmax_fwd_f32(Z.data_ptr(), ldz, n, K, P.data_ptr(), _stream()), "gmc_softmax_fwd_f32")
    return P


def relu_bwd(dY: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """dX = (Y > 0) ? dY : 0 in one launch (the autograd path's ReLU backward)."""
    dY, lddy = _rowmajor(dY, "dY")
    Y, ldy = _rowmajor(Y, "Y")
    if dY.shape != Y.shape:
        raise ValueError("relu_bwd: shapes differ")
    out = padded_empty(dY.shape[0], dY.shape[1], dY.device)
    check(lib().gmc_relu_bwd_f32(dY.data_ptr(), lddy, Y.data_ptr(), ldy, out.data_ptr(), out.stride(0), dY.shape[0],
                                 dY.shape[1], _stream()), "gmc_relu_bwd_f32")
    return out


def softmax_bwd(P: torch.Tensor, dP: torch.Tensor) -> torch.Tensor:
    P = _f32(P, "P").contiguous()
    dP = _f32(dP, "dP").contiguous()
    dZ = torch.empty_like(P)
    check(lib().gmc_softmax_bwd_f32(P.data_ptr(), dP.data_ptr(), P.shape[0], P.shape[1], dZ.data_ptr(), _stream()),
          "gmc_softmax_bwd_f32")
    return dZ


# ---------------------------------------------------------------- (d) Adam
def _ptr_array(ts: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


def adam_multi(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], exp_avg: Sequence[torch.Tensor],
               exp_avg_sq: Sequence[torch.Tensor], lr: float, beta1: float = 0.9, beta2: float = 0.999,
               eps: float = 1e-8, step: Optional[int] = None, step_dev: Optional[torch.Tensor] = None,
               shadows: Optional[Sequence[Optional[torch.Tensor]]] = None) -> None:
    """In-place Adam on up to 16 tensors per launch.  Pass either `step` (1-based host count) or
    `step_dev` (int64 device scalar holding the number of steps taken; incremented on device) -- or an 8-word int64
    state whose word 0 is that count (gmc_adam_multi_devstate: no increment launch, bias corrections precomputed)."""
    n = len(params)
    for group in (params, grads, exp_avg, exp_avg_sq):
        if len(group) != n:
            raise ValueError("adam_multi: list lengths differ")
        for t in group:
            _f32(t, "adam tensor")
            if not t.is_contiguous():
                raise ValueError("adam_multi: tensors must be contiguous")
    for lo in range(0, n, 16):
        hi = min(n, lo + 16)
        sizes = (ctypes.c_int64 * (hi - lo))(*[p.numel() for p in params[lo:hi]])
        args = (hi - lo, _ptr_array(params[lo:hi]), _ptr_array(grads[lo:hi]), _ptr_array(exp_avg[lo:hi]),
                _ptr_array(exp_avg_sq[lo:hi]), sizes, float(lr), float(beta1), float(beta2), float(eps))
        if step_dev is not None:
            if hi < n:
                raise ValueError("adam_multi with step_dev supports at most 16 tensors per call")
            if step_dev.dtype != torch.int64 or not step_dev.is_cuda or not step_dev.is_contiguous() or step_dev.numel() not in (1, 8):
                raise ValueError("step_dev must be a contiguous CUDA int64 tensor of 1 (count) or 8 (state) elements")
            if step_dev.numel() == 8:
                check(lib().gmc_adam_multi_devstate(*args, step_dev.data_ptr(), _stream()), "gmc_adam_multi_devstate")
            else:
                check(lib().gmc_adam_multi_devstep(*args, step_dev.data_ptr(), _stream()), "gmc_adam_multi_devstep")
        else:
            if step is None or step < 1:
                raise ValueError("adam_multi: step must be >= 1")
            if shadows is not None and any(sh is not None for sh in shadows[lo:hi]):
                # bf16 copies of the updated parameters (same element index), written by the same pass
                arr = (ctypes.c_void_p * (hi - lo))()
                for i, (sh, p) in enumerate(zip(shadows[lo:hi], params[lo:hi])):
                    if sh is None:
                        continue
                    if sh.dtype != torch.bfloat16 or not sh.is_contiguous() or sh.numel() != p.numel() or not sh.is_cuda:
                        raise ValueError("adam_multi: a shadow must be a contiguous CUDA bf16 tensor with the parameter's numel")
                    arr[i] = sh.data_ptr()
                a = args[:5] + (arr,) + args[5:]
                check(lib().gmc_adam_multi_shadow(*a, int(step), _stream()), "gmc_adam_multi_shadow")
            else:
                check(lib().gmc_adam_multi(*args, int(step), _stream()), "gmc_adam_multi")


# ---------------------------------------------------------------- (e) post-processing
def argmax_labels(batch, P: torch.Tensor, force_terminals: bool = True) -> torch.Tensor:
    P, ldp = _rowmajor(P, "P")
    n, K = P.shape
    labels = torch.empty(n, dtype=torch.int32, device=P.device)
    check(lib().gmc_argmax_labels(P.data_ptr(), ldp, batch.graph_ptr.data_ptr(), batch.num_graphs, n, K,
                                  int(force_terminals), labels.data_ptr(), _stream()), "gmc_argmax_labels")
    return labels


def _int_weights(batch) -> Optional[torch.Tensor]:
    if batch.unit_weights:
        return None
    if batch.wts_i32 is None:
        raise ValueError("integer cut kernels need integer edge weights (the reference writes weight=1, "
                         "GraphCreator.py:88-90)")
    return batch.wts_i32


def cut_value(batch, labels: torch.Tensor) -> torch.Tensor:
    if labels.dtype != torch.int32 or not labels.is_cuda or not labels.is_contiguous():
        raise TypeError("labels: expected a contiguous CUDA int32 tensor")
    out = torch.empty(batch.num_graphs, dtype=torch.int64, device=labels.device)
    check(lib().gmc_cut_value_i32(labels.data_ptr(), batch.rowptr.data_ptr(), batch.colidx.data_ptr(),
                                  _ptr(_int_weights(batch)), batch.graph_ptr.data_ptr(), batch.num_graphs,
                                  batch.num_nodes, out.data_ptr(), _stream()), "gmc_cut_value_i32")
    return out


def cut_value_multi(batch, labels_u8: torch.Tensor) -> torch.Tensor:
    """Cuts of T labelings (uint8 [T, n]) of a ONE-graph batch; returns int64 [T]."""
    if batch.num_graphs != 1:
        raise ValueError("cut_value_multi evaluates many labelings of a single graph")
    if labels_u8.dtype != torch.uint8 or not labels_u8.is_cuda or not labels_u8.is_contiguous() or labels_u8.dim() != 2:
        raise TypeError("labels: expected a contiguous CUDA uint8 tensor [T, n]")
    T, n = labels_u8.shape
    if n != batch.num_nodes:
        raise ValueError("labels have the wrong number of nodes")
    out = torch.empty(T, dtype=torch.int64, device=labels_u8.device)
    check(lib().gmc_cut_value_multi_u8(labels_u8.data_ptr(), batch.rowptr.data_ptr(), batch.colidx.data_ptr(),
                                       _ptr(_int_weights(batch)), n, T, out.data_ptr(), _stream()),
          "gmc_cut_value_multi_u8")
    return out


def numpy_rand_on_device(n: int, device) -> Optional[torch.Tensor]:
    """n doubles of the GLOBAL numpy generator's np.random.rand stream, produced on the device (gmc_mt19937_uniform_f64);
    the host generator is advanced to exactly where n scalar draws would have left it.  None when the global generator is
    not the legacy MT19937 (the caller then draws on the host)."""
    import numpy as np
    st = np.random.get_state()
    if st[0] != "MT19937" or n <= 0:
        return None
    key = torch.from_numpy(np.ascontiguousarray(st[1], dtype=np.uint32).view(np.int32)).to(device)
    out = torch.empty(n, dtype=torch.float64, device=device)
    new_key = torch.empty(624, dtype=torch.int32, device=device)
    new_pos = torch.empty(1, dtype=torch.int32, device=device)
    check(lib().gmc_mt19937_uniform_f64(key.data_ptr(), int(st[2]), int(n), out.data_ptr(), new_key.data_ptr(),
                                        new_pos.data_ptr(), _stream()), "gmc_mt19937_uniform_f64")
    host_key = new_key.cpu().numpy().view(np.uint32)              # synchronises: the state must be back before anyone draws
    np.random.set_state((st[0], host_key, int(new_pos.item()), st[3], st[4]))
    return out


def sample_best_cut(batch, P: torch.Tensor, U: torch.Tensor, u_ptr: torch.Tensor, iters: int,
                    compare_f32: bool):
    """P1.  U float64 device uniforms, u_ptr int64 [B+1] offsets.  Returns (labels, cut, best_iter)."""
    P, ldp = _rowmajor(P, "P")
    n, K = P.shape
    if U.dtype != torch.float64 or u_ptr.dtype != torch.int64:
        raise TypeError("U must be float64 and u_ptr int64")
    dev = P.device
    cuts = torch.empty(batch.num_graphs * iters, dtype=torch.int64, device=dev)
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    best = torch.empty(batch.num_graphs, dtype=torch.int64, device=dev)
    best_it = torch.empty(batch.num_graphs, dtype=torch.int32, device=dev)
    check(lib().gmc_sample_best_cut(P.data_ptr(), ldp, U.data_ptr(), u_ptr.data_ptr(), batch.rowptr.data_ptr(),
                                    batch.colidx.data_ptr(), _ptr(_int_weights(batch)), batch.graph_ptr.data_ptr(),
                                    batch.num_graphs, n, K, iters, int(compare_f32), cuts.data_ptr(),
                                    labels.data_ptr(), best.data_ptr(), best_it.data_ptr(), _stream()),
          "gmc_sample_best_cut")
    return labels, best, best_it


def greedy_node_move(batch, labels: torch.Tensor, n_classes: int = 3, iters: int = 200, n_frozen: int = 3):
    """P2.  Returns (labels_out int32 [N], cut int64 [B], moves int32 [B])."""
    if labels.dtype != torch.int32 or not labels.is_cuda or not labels.is_contiguous():
        raise TypeError("labels: expected a contiguous CUDA int32 tensor")
    dev = labels.device
    out = torch.empty_like(labels)
    cut = torch.empty(batch.num_graphs, dtype=torch.int64, device=dev)
    moves = torch.empty(batch.num_graphs, dtype=torch.int32, device=dev)
    check(lib().gmc_greedy_node_move(labels.data_ptr(), batch.rowptr.data_ptr(), batch.colidx.data_ptr(),
                                     _ptr(_int_weights(batch)), batch.graph_ptr.data_ptr(), batch.num_graphs,
                                     batch.num_nodes, n_classes, iters, n_frozen, out.data_ptr(), cut.data_ptr(),
                                     moves.data_ptr(), _stream()), "gmc_greedy_node_move")
    return out, cut, moves
