"""GPU tests (`-m gpu`) of the tcgen05 / TMA TF32 GEMM path against an exact emulation:
the MMA truncates fp32 operands to TF32 (10 explicit mantissa bits), multiplies exactly and
accumulates in fp32 -- so with operands truncated on the host the float64 product must agree to
fp32 accumulation error.  Tolerance stated per test."""
import math

import numpy as np
import pytest
import torch

from gmc_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def tf32_trunc(t: torch.Tensor) -> torch.Tensor:
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def reference(op, A, B):
    A, B = tf32_trunc(A).double(), tf32_trunc(B).double()
    if op == "nn":
        return A @ B
    if op == "nt":
        return A @ B.t()
    return A.t() @ B


SHAPES = [
    ("nn", 128, 256, 32), ("nn", 128, 256, 64), ("nn", 256, 512, 128), ("nn", 1000, 500, 1000), ("nn", 130, 260, 40),
    ("nn", 5, 12, 8), ("nn", 40000, 500, 1000),
    ("nt", 128, 256, 32), ("nt", 200, 96, 500), ("nt", 1000, 1000, 500), ("nt", 77, 1000, 12),
    ("tn", 128, 256, 32), ("tn", 128, 256, 4096), ("tn", 1000, 500, 3000), ("tn", 1000, 500, 200000), ("tn", 100, 64, 128),
    ("tn", 36, 8, 20),
]


@pytest.mark.parametrize("op,M,N,K", SHAPES)
def test_tf32_gemm_matches_truncated_float64(op, M, N, K):
    torch.manual_seed(M * 7 + N * 3 + K)
    if op == "nn":
        A, B = torch.randn(M, K), torch.randn(K, N)
    elif op == "nt":
        A, B = torch.randn(M, K), torch.randn(N, K)
    else:
        A, B = torch.randn(K, M), torch.randn(K, N)
    want = reference(op, A, B)
    got = ops.gemm(op, A.to(DEV), B.to(DEV), precision="tf32")
    tol = 2e-6 * max(1.0, math.sqrt(K) / 8)
    assert relerr(got.cpu(), want) < tol
    # and it is a TF32 product, i.e. within ~2^-10 of the exact fp32 one
    exact = reference(op, A, B) if False else (A.double() @ B.double() if op == "nn" else
                                               A.double() @ B.double().t() if op == "nt" else A.double().t() @ B.double())
    assert relerr(got.cpu(), exact) < 4e-3
    acc = torch.full((M, N), 2.0, device=DEV)
    ops.gemm(op, A.to(DEV), B.to(DEV), out=acc, accumulate=True, precision="tf32")
    assert relerr(acc.cpu(), want + 2.0) < tol


def test_tf32_gemm_strided_operands_and_output():
    torch.manual_seed(1)
    Abig = torch.randn(300, 1024, device=DEV)
    A = Abig[:, 8:1008]                                  # ld 1024, 32-byte offset
    B = torch.randn(1000, 512, device=DEV)[:, :500]      # ld 512
    Cbig = torch.zeros(300, 640, device=DEV)
    ops.gemm("nn", A, B, out=Cbig[:, 64:564], precision="tf32")
    want = reference("nn", A.cpu(), B.cpu())
    assert relerr(Cbig[:, 64:564].cpu(), want) < 1e-5
    assert float(Cbig[:, :64].abs().sum()) == 0.0 and float(Cbig[:, 564:].abs().sum()) == 0.0


def test_tf32_exact_on_adjacency_features():
    # 0/1 features are exactly representable, so only W is truncated (the parity argument in DESIGN.md)
    torch.manual_seed(2)
    X = (torch.rand(512, 1000) < 0.007).float()
    W = torch.randn(1000, 500) * 0.05
    got = ops.gemm("nn", X.to(DEV), W.to(DEV), precision="tf32")
    assert relerr(got.cpu(), X.double() @ tf32_trunc(W).double()) < 2e-6
    assert relerr(got.cpu(), X.double() @ W.double()) < 1.5e-3


@pytest.mark.parametrize("op,M,N,K", [("nn", 1000, 500, 1000), ("nn", 132, 260, 40), ("nt", 200, 96, 500),
                                      ("tn", 1000, 500, 3000), ("tn", 1000, 500, 200000)])
def test_tf32x3_is_fp32_grade(op, M, N, K):
    """3xTF32 (hi*hi + lo*hi + hi*lo on the tensor cores): agrees with the float64 product of the UNtruncated
    operands to fp32 accumulation error, i.e. ~1000x closer than one TF32 pass."""
    torch.manual_seed(M + N + K)
    if op == "nn":
        A, B = torch.randn(M, K), torch.randn(K, N)
        exact = A.double() @ B.double()
    elif op == "nt":
        A, B = torch.randn(M, K), torch.randn(N, K)
        exact = A.double() @ B.double().t()
    else:
        A, B = torch.randn(K, M), torch.randn(K, N)
        exact = A.double().t() @ B.double()
    got = ops.gemm(op, A.to(DEV), B.to(DEV), precision="tf32x3")
    one_pass = ops.gemm(op, A.to(DEV), B.to(DEV), precision="tf32")
    tol = 3e-6 * max(1.0, math.sqrt(K) / 8)
    assert relerr(got.cpu(), exact) < tol
    # (for very long K the fp32 accumulation of the split-K partials, common to both, narrows the gap)
    assert relerr(one_pass.cpu(), exact) > (20 if K <= 4096 else 5) * relerr(got.cpu(), exact)
    acc = torch.full((M, N), -1.0, device=DEV)
    ops.gemm(op, A.to(DEV), B.to(DEV), out=acc, accumulate=True, precision="tf32x3")
    assert relerr(acc.cpu(), exact - 1.0) < tol


@pytest.mark.parametrize("op,M,N,K", [("nn", 128, 256, 64), ("nn", 1000, 500, 1000), ("nn", 40000, 500, 1000),
                                      ("nn", 130, 260, 72), ("nt", 128, 256, 64), ("nt", 200, 96, 500),
                                      ("nt", 1000, 1000, 500), ("tn", 128, 256, 64), ("tn", 1000, 500, 3000),
                                      ("tn", 1000, 500, 200000), ("tn", 100, 64, 128)])
def test_bf16_gemm_matches_float64_of_rounded_operands(op, M, N, K):
    """bf16 operands (round to nearest even), exact products, fp32 accumulation in TMEM: the float64 product of the
    ROUNDED operands must agree to fp32 accumulation error -- for K-major and MN-major operands alike."""
    torch.manual_seed(M + 3 * N + 7 * K)
    if op == "nn":
        A, B = torch.randn(M, K), torch.randn(K, N)
    elif op == "nt":
        A, B = torch.randn(M, K), torch.randn(N, K)
    else:
        A, B = torch.randn(K, M), torch.randn(K, N)
    Ab, Bb = ops.to_bf16(A.to(DEV)), ops.to_bf16(B.to(DEV))
    assert torch.equal(Ab.cpu(), A.to(torch.bfloat16)) and torch.equal(Bb.cpu(), B.to(torch.bfloat16))
    Ad, Bd = Ab.cpu().double(), Bb.cpu().double()
    want = Ad @ Bd if op == "nn" else Ad @ Bd.t() if op == "nt" else Ad.t() @ Bd
    got = ops.gemm_bf16(op, Ab, Bb)
    tol = 2e-6 * max(1.0, math.sqrt(K) / 8)
    assert relerr(got.cpu(), want) < tol
    acc = torch.full((M, N), 3.0, device=DEV)
    ops.gemm_bf16(op, Ab, Bb, out=acc, accumulate=True)
    assert relerr(acc.cpu(), want + 3.0) < tol


def test_tf32_rejects_unaligned():
    A = torch.randn(64, 33, device=DEV)
    with pytest.raises(_lib.GmcError):
        ops.gemm("nn", A, torch.randn(33, 16, device=DEV), precision="tf32")


def test_tf32_is_deterministic_with_split_k():
    torch.manual_seed(3)
    A, B = torch.randn(150000, 1000, device=DEV), torch.randn(150000, 500, device=DEV)
    a = ops.gemm("tn", A, B, precision="tf32")
    b = ops.gemm("tn", A, B, precision="tf32")
    assert torch.equal(a, b)
    assert relerr(a.cpu(), reference("tn", A.cpu(), B.cpu())) < 1e-4
