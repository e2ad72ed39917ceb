"""Standalone timings of the bf16-activation kernels at the config-3 shape (B graphs x n=1000, d=7, H=500, K=3)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops, synth
B = int(os.environ.get("B", "2048"))
REP = int(os.environ.get("REP", "5"))
PEAK = 6556.5
batch = synth.regular_batch(B, 1000, 7, seed=1)
N, H, K = batch.num_nodes, 500, 3
nnz = batch.nnz
A16 = ops.padded_empty_bf16(N, H, "cuda", zero=True); A16.copy_(torch.randn(N, H, device="cuda").to(torch.bfloat16))
B16 = ops.padded_empty_bf16(N, H, "cuda", zero=True)
W2 = torch.randn(H, K, device="cuda") / H ** 0.5
b1 = torch.randn(H, device="cuda") * 0.1
T2 = torch.empty(N, K, device="cuda")
dT2 = torch.randn(N, K, device="cuda")
NCU = os.environ.get("NCU", "0") == "1"           # under ncu: one launch per kernel, no warm-up
def timeit(fn, n=REP):
    if NCU:
        fn(); torch.cuda.synchronize(); return 1.0
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def report(name, ms, nbytes):
    print(f"{name:34s} {ms:7.3f} ms  {nbytes/ms/1e6:6.0f} GB/s  frac {nbytes/ms/1e6/PEAK:.3f}", flush=True)
idx = 4.0 * nnz + 4.0 * (N + 1)
report("fused row fwd (bf16 in, bf16 out)", timeit(lambda: ops.spmm_fused_skinny_bf16(batch, A16, W2, out=B16, proj=T2, bias=b1, relu=True)), 4.0 * N * H + idx + 4.0 * N * K)
Y32 = ops.padded_empty(N, H, "cuda")
report("fused row fwd (bf16 in, fp32 out)", timeit(lambda: ops.spmm_fused_skinny_bf16(batch, A16, W2, out=Y32, proj=T2, bias=b1, relu=True)), 6.0 * N * H + idx + 4.0 * N * K)
del Y32
report("slab fwd bias+relu (bf16, bf16)", timeit(lambda: ops.spmm_bf16(batch, A16, out=B16, bias=b1, relu=True)), 4.0 * N * H + idx)
report("skinny_fwd bf16", timeit(lambda: ops.skinny_fwd_bf16(B16, W2, out=T2)), 2.0 * N * H + 4.0 * N * K)
report("slab bwd (bf16, bf16)", timeit(lambda: ops.spmm_bf16(batch, A16, out=B16)), 4.0 * N * H + idx)
dW = torch.empty(H, K, device="cuda"); db = torch.empty(H, device="cuda")
C16 = ops.padded_empty_bf16(N, H, "cuda", zero=True)
report("skinny_bwd bf16", timeit(lambda: ops.skinny_bwd_bf16(dT2, W2, B16, dH=C16, dW=dW, dbias=db)), 4.0 * N * H + 4.0 * N * K)
