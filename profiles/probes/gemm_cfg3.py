# config-3 shaped GEMMs (pitched operands like the engine): nn X[N,1000]*W1[1000,500], tn X^T dT1, nt dT1*W1^T
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops
dev = "cuda"
G = int(os.environ.get("B", "4096"))
N = G * 1000
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
X = ops.padded_empty(N, 1000, dev); X.normal_()
W = ops.padded_empty(1000, 500, dev); W.normal_()
T = ops.padded_empty(N, 500, dev)
ws = ops.Workspace()
fl = 2.0 * N * 1000 * 500
ms = timeit(lambda: ops.gemm("nn", X, W, out=T, precision="tf32", workspace=ws))
print(f"nn: {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")
# check vs truncated reference on a slice
ref = (X[:4096].double() @ W.double())
err = float((T[:4096].double() - ref).abs().max() / ref.abs().max())
print(f"nn relerr {err:.2e}")
dW = torch.empty(1000, 500, device=dev)
ms = timeit(lambda: ops.gemm("tn", X, T, out=dW, precision="tf32", workspace=ws))
print(f"tn: {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s")
if G <= 1024:
    ref = X.double().t() @ T.double()
    print(f"tn relerr {float((dW.double()-ref).abs().max()/ref.abs().max()):.2e}")
