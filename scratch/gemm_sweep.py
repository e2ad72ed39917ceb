import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops
dev = "cuda"
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
M = 1024000
for N in (256, 500):
    for K in (256, 512, 1000, 2000):
        A = torch.randn(M, K, device=dev); B = torch.randn(K, N, device=dev); C = torch.empty(M, N, device=dev)
        ms = timeit(lambda: ops.gemm("nn", A, B, out=C, precision="tf32"))
        tiles = (M // 128) * ((N + 255) // 256)
        print(f"nn M={M} N={N} K={K}: {ms:7.3f} ms  {2*M*N*K/ms/1e9:7.1f} TF/s  per-tile-per-SM {ms*1e3*148/tiles:6.2f} us  A+C GB/s {(M*K+M*N)*4/ms/1e6:7.0f}")
        del A, B, C
Kn = 1024000
A = torch.randn(Kn, 1000, device=dev); B = torch.randn(Kn, 500, device=dev)
ms = timeit(lambda: ops.gemm("tn", A, B, precision="tf32"))
print(f"tn K={Kn} 1000x500: {ms:7.3f} ms {2*Kn*1000*500/ms/1e9:7.1f} TF/s")
