# config-3 shaped bf16 GEMMs (pitched operands like the engine): nn X[N,1000]*W1[1000,500] -> bf16, tn X^T dT1 -> fp32.
# GMC_GEMM_CLUSTER = 1 | 2 | 4 | 8 | 2x2 | 4x2 selects the cluster shape of both.
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops
dev = "cuda"
G = int(os.environ.get("B", "4096"))
N = G * 1000
def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
X = ops.padded_empty_bf16(N, 1000, dev, zero=True)
X.copy_((torch.rand(N, 1000, device=dev) < 0.007).to(torch.bfloat16))
W = ops.padded_empty_bf16(1000, 500, dev, zero=True); W.copy_(torch.randn(1000, 500, device=dev).to(torch.bfloat16))
T = ops.padded_empty_bf16(N, 500, dev, zero=True)
ws = ops.Workspace()
fl = 2.0 * N * 1000 * 500
ms = timeit(lambda: ops.gemm_bf16_bf16out("nn", X, W, out=T))
print(f"cluster {os.environ.get('GMC_GEMM_CLUSTER', 'default')}: nn {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s", end="  ")
dW = torch.empty(1000, 500, device=dev)
ms = timeit(lambda: ops.gemm_bf16("tn", X, T, out=dW, workspace=ws))
print(f"tn {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")
