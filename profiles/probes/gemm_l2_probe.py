# Is the bf16 nn GEMM bound by the DRAM stream of X or by operand delivery into the SMs?  Same kernel, M small enough
# for X + T1 to stay L2-resident (repeated launches) against the config-3 M.
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
W = ops.padded_empty_bf16(1000, 500, dev, zero=True); W.copy_(torch.randn(1000, 500, device=dev).to(torch.bfloat16))
for M in (148 * 128 * 2, 148 * 128 * 4, 148 * 128 * 8, 148 * 128 * 32, 148 * 128 * 128):
    X = ops.padded_empty_bf16(M, 1000, dev, zero=True)
    X.copy_((torch.rand(M, 1000, device=dev) < 0.007).to(torch.bfloat16))
    T = ops.padded_empty_bf16(M, 500, dev, zero=True)
    ms = timeit(lambda: ops.gemm_bf16_bf16out("nn", X, W, out=T))
    fl = 2.0 * M * 1000 * 500
    print(f"M={M:8d}  X {M*2048/1e6:7.1f} MB  {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TF/s", flush=True)
    del X, T
