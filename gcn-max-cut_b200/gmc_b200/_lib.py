"""ctypes binding of libgcnmaxcut.so (C ABI declared in include/gcnmaxcut.h).

The product path has NO CPU fallback: `lib()` raises if the shared library is
missing and `require_cuda()` raises if no CUDA device is visible.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p
from typing import Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("GMC_LIB", os.path.join(PKG_ROOT, "lib", "libgcnmaxcut.so"))

GMC_GEMM_FP32, GMC_GEMM_TF32, GMC_GEMM_TF32X3 = 0, 1, 2
GMC_LOSS_STE, GMC_LOSS_SOFT = 0, 1
PRECISIONS = {"fp32": GMC_GEMM_FP32, "tf32": GMC_GEMM_TF32, "tf32x3": GMC_GEMM_TF32X3}
# engine-level precision (GCNEngine / TrainingConfig.gemm_precision): the three above plus "bf16" (bf16 operands
# through gmc_gemm_bf16; not a gmc_gemm_* precision code because its operands are a different type)
# and "bf16x2" / "bf16x3" (fp32-grade: exact integer features x an fp32 operand split into 2 / 3 bf16 parts, csrc/split.cu)
# "f16x2": the same with W1 split into two fp16 parts (22 mantissa bits) -- fp32-grade at the cost of bf16x2
ENGINE_PRECISIONS = tuple(PRECISIONS) + ("bf16", "bf16x2", "bf16x3", "f16x2")
LOSS_MODES = {"ste": GMC_LOSS_STE, "soft": GMC_LOSS_SOFT}


class GmcError(RuntimeError):
    """Non-zero return of a gmc_* entry point."""


P = c_void_p  # every device pointer / stream travels as void*

# name -> (restype, [argtypes])   -- must list EVERY symbol of include/gcnmaxcut.h
SIGNATURES: Dict[str, tuple] = {
    "gmc_abi_version": (c_int, []),
    "gmc_last_error": (c_char_p, []),
    "gmc_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "gmc_degree_norm_f32": (c_int, [P, c_int64, P, P, P]),
    "gmc_edge_coef_f32": (c_int, [P, P, P, P, P, c_int64, P, P]),
    "gmc_csr_densify_f32": (c_int, [P, P, P, P, c_int32, c_int64, c_int32, P, c_int64, P]),
    "gmc_copy2d_f32": (c_int, [P, c_int64, P, c_int64, c_int64, c_int32, P]),
    "gmc_spmm_symnorm_f32": (c_int, [P, P, P, P, P, P, P, c_int64, c_int32, c_int64, c_int64, P, c_int32, P]),
    "gmc_spmm_plan_bytes": (c_size_t, [c_int64]),
    "gmc_spmm_plan_build": (c_int, [P, P, P, P, P, c_int32, c_int64, P, P, P]),
    "gmc_spmm_batched_f32": (c_int, [P, P, P, P, c_int32, c_int32, P, P, P, c_int64, c_int32, c_int64, c_int64, P, c_int32, P]),
    "gmc_spmm_fused_skinny_f32": (c_int, [P, P, P, P, P, P, P, c_int64, c_int32, c_int64, c_int64, P, c_int32, P, c_int32,
                                          P, c_int64, P]),
    "gmc_spmm_batched_fused_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "gmc_spmm_batched_fused_skinny_f32": (c_int, [P, c_int32, c_int32, P, P, P, c_int64, c_int32, c_int64, c_int64, P,
                                                  c_int32, P, c_int32, P, c_int64, P, c_size_t, P]),
    "gmc_spmm_batched_bf16out": (c_int, [P, c_int32, c_int32, P, P, P, c_int64, c_int32, c_int64, c_int64, P]),
    "gmc_gemm_bf16_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int64, c_int64]),
    "gmc_gemm_bf16": (c_int, [c_int32, P, P, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int32, P,
                              c_size_t, P]),
    "gmc_gemm_bf16_bf16out": (c_int, [c_int32, P, P, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, P, c_int32,
                                      P, P, c_int64, c_int32, P]),
    "gmc_spmm_fused_skinny_bf16": (c_int, [P, P, P, P, P, P, P, c_int32, c_int64, c_int32, c_int64, c_int64, P, c_int32,
                                           P, c_int32, P, c_int64, P]),
    "gmc_skinny_bwd_bf16": (c_int, [P, c_int64, P, P, c_int64, P, c_int64, P, P, c_int64, c_int32, c_int32, P, c_size_t, P]),
    "gmc_spmm_batched_bf16": (c_int, [P, c_int32, c_int32, P, P, P, c_int64, c_int32, c_int64, c_int64, P, c_int32, P]),
    "gmc_skinny_fwd_bf16": (c_int, [P, c_int64, P, P, c_int64, c_int64, c_int32, c_int32, P]),
    "gmc_f32_to_bf16": (c_int, [P, c_int64, P, c_int64, c_int64, c_int32, P]),
    "gmc_csr_densify_bf16": (c_int, [P, P, P, P, c_int32, c_int64, c_int32, P, c_int64, P]),
    "gmc_csr_scatter_bf16": (c_int, [P, P, P, P, c_int32, c_int64, c_int32, P, c_int64, c_int32, P]),
    "gmc_csr_preaggregate_workspace_bytes": (c_size_t, [c_int64]),
    "gmc_csr_preaggregate_bf16": (c_int, [P, P, P, P, P, c_int32, c_int64, c_int32, P, c_int64, P, c_size_t, P]),
    "gmc_adj_features_fwd_f32": (c_int, [P, P, c_int32, c_int32, P, c_int64, c_int32, P, c_int64, c_int64, c_int32, P]),
    "gmc_adj_features_bwd_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "gmc_adj_features_bwd_f32": (c_int, [P, P, c_int32, c_int32, P, c_int64, c_int64, c_int32, P, c_int64, c_int32, P,
                                         c_size_t, P]),
    "gmc_f32_split_bf16": (c_int, [P, c_int64, P, c_int64, c_int64, c_int32, c_int32, c_int64, P]),
    "gmc_row_scale_f32": (c_int, [P, P, c_int64, P, P, P]),
    "gmc_gemm_bf16_split_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int64, c_int64, c_int32, c_int32]),
    "gmc_gemm_bf16_split": (c_int, [c_int32, P, P, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int32, c_int64,
                                    P, P, c_int32, P, P, c_int64, c_int32, c_int32, c_int32, P, c_size_t, P]),
    "gmc_f32_split_f16": (c_int, [P, c_int64, P, c_int64, c_int64, c_int32, c_int32, c_int64, c_int32, P]),
    "gmc_skinny_bwd_split": (c_int, [P, c_int64, P, P, c_int64, P, P, c_int64, c_int64, c_int32, c_int32, P, P, c_int64, c_int32,
                                     c_int32, P, c_size_t, P]),
    "gmc_csr_preaggregate_graphs": (c_int, [P, P, P, P, P, c_int32, c_int32, c_int64, c_int32, P, c_int64, c_int32, P]),
    "gmc_csr_preaggregate_f16": (c_int, [P, P, P, P, P, c_int32, c_int64, c_int32, P, c_int64, P]),
    "gmc_gemm_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int64, c_int64, c_int32]),
    "gmc_gemm_nn": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, P, c_size_t, P]),
    "gmc_gemm_nt": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, P, c_size_t, P]),
    "gmc_gemm_tn": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, P, c_size_t, P]),
    "gmc_skinny_fwd_f32": (c_int, [P, c_int64, P, P, c_int64, c_int64, c_int32, c_int32, P]),
    "gmc_skinny_bwd_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "gmc_skinny_bwd_f32": (c_int, [P, c_int64, P, P, c_int64, P, c_int64, P, P, c_int64, c_int32, c_int32, P, c_size_t, P]),
    "gmc_colsum_workspace_bytes": (c_size_t, [c_int32]),
    "gmc_colsum_f32": (c_int, [P, c_int64, c_int64, c_int32, P, P, c_size_t, P]),
    "gmc_softmax_cut_loss_fwd_bwd": (c_int, [P, c_int64, P, P, P, P, c_int32, c_int64, c_int32, c_int32, c_int32,
                                             c_float, c_float, P, P, P, P]),
    "gmc_layer2_loss_fused_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "gmc_layer2_loss_fused": (c_int, [P, c_int64, P, P, P, P, P, c_int32, c_int32, c_int64, c_int32, P, c_int32, c_int32,
                                      c_float, c_float, P, P, P, P, P, c_int64, P, P, c_size_t, P]),
    "gmc_layer2_loss_fused_parts": (c_int, [P, c_int32, c_int64, P, c_int64, P, P, P, P, P, c_int32, c_int32, c_int64, c_int32, P, c_int32, c_int32,
                                      c_float, c_float, P, P, P, P, P, c_int64, P, P, c_size_t, P]),
    "gmc_softmax_fwd_f32": (c_int, [P, c_int64, c_int64, c_int32, P, P]),
    "gmc_relu_bwd_f32": (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int64, c_int32, P]),
    "gmc_softmax_bwd_f32": (c_int, [P, P, c_int64, c_int32, P, P]),
    "gmc_adam_multi": (c_int, [c_int32, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                               POINTER(c_int64), c_double, c_double, c_double, c_double, c_int64, P]),
    "gmc_adam_multi_shadow": (c_int, [c_int32, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                      POINTER(c_void_p), POINTER(c_int64), c_double, c_double, c_double, c_double, c_int64, P]),
    "gmc_adam_multi_devstep": (c_int, [c_int32, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                       POINTER(c_void_p), POINTER(c_int64), c_double, c_double, c_double, c_double,
                                       P, P]),
    "gmc_adam_multi_devstate": (c_int, [c_int32, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                        POINTER(c_void_p), POINTER(c_int64), c_double, c_double, c_double, c_double,
                                        P, P]),
    "gmc_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "gmc_peer_free": (c_int, [P]),
    "gmc_peer_flag_bytes": (c_size_t, []),
    "gmc_ipc_handle_bytes": (c_int32, []),
    "gmc_ipc_get_handle": (c_int, [P, P]),
    "gmc_ipc_open_handle": (c_int, [P, POINTER(c_void_p)]),
    "gmc_ipc_close_handle": (c_int, [P]),
    "gmc_peer_allreduce_f32": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int32, c_int32, c_int64, c_uint32, P]),
    "gmc_mt19937_uniform_f64": (c_int, [P, c_int32, c_int64, P, P, P, P]),
    "gmc_argmax_labels": (c_int, [P, c_int64, P, c_int32, c_int64, c_int32, c_int32, P, P]),
    "gmc_cut_value_i32": (c_int, [P, P, P, P, P, c_int32, c_int64, P, P]),
    "gmc_cut_value_multi_u8": (c_int, [P, P, P, P, c_int32, c_int32, P, P]),
    "gmc_sample_best_cut": (c_int, [P, c_int64, P, P, P, P, P, P, c_int32, c_int64, c_int32, c_int32, c_int32, P, P,
                                    P, P, P]),
    "gmc_greedy_node_move": (c_int, [P, P, P, P, P, c_int32, c_int64, c_int32, c_int32, c_int32, P, P, P, P]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libgcnmaxcut.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GmcError(
                f"{LIB_PATH} not found: build it with `python gcn-max-cut_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the GCN max-cut hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError == ABI mismatch, surfaced as is
            fn.restype = res
            fn.argtypes = args
        if handle.gmc_abi_version() != 1:
            raise GmcError(f"ABI version mismatch: library reports {handle.gmc_abi_version()}, binding expects 1")
        _lib = handle
    return _lib


def exported_symbols() -> List[str]:
    return sorted(SIGNATURES)


def last_error() -> str:
    msg = lib().gmc_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        kind = "argument/usage error" if rc < 0 else "CUDA error"
        raise GmcError(f"{what or 'gmc call'} failed ({kind} {rc}): {last_error()}")


def require_cuda():
    """Return the CUDA device to run on, or raise: the hot path never runs on the CPU."""
    import torch
    if not torch.cuda.is_available():
        raise GmcError("no CUDA device visible: the GCN max-cut hot path is CUDA-only (sm_100a); "
                       "there is deliberately no CPU fallback")
    lib()
    return torch.device("cuda", torch.cuda.current_device())
