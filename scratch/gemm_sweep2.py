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
print("mode 1CTA" if os.environ.get("GMC_GEMM_1CTA") == "1" else "mode 2CTA")
A = torch.randn(M, 512, device=dev); B = torch.randn(1024, 512, device=dev); C = torch.empty(M, 1024, device=dev)
ms = timeit(lambda: ops.gemm("nt", A, B, out=C, precision="tf32"))
print(f"nt M={M} N=1024 K=512 (both K-major): {ms:7.3f} ms {2*M*1024*512/ms/1e9:7.1f} TF/s")
del A, B, C
A = torch.randn(M, 1024, device=dev); B = torch.randn(1024, 512, device=dev); C = torch.empty(M, 512, device=dev)
ms = timeit(lambda: ops.gemm("nn", A, B, out=C, precision="tf32"))
print(f"nn M={M} N=512 K=1024: {ms:7.3f} ms {2*M*1024*512/ms/1e9:7.1f} TF/s")
Bt = B.t().contiguous()
ms = timeit(lambda: ops.gemm("nt", A, Bt, out=C, precision="tf32"))
print(f"nt M={M} N=512 K=1024 (same product, B K-major): {ms:7.3f} ms {2*M*1024*512/ms/1e9:7.1f} TF/s")
