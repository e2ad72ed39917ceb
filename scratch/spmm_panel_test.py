# hypothesis test: is the slab kernel limited by the strided (112 B of every 2 KB) DRAM pattern?
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops, synth
B = int(os.environ.get("B", "2048"))
W = int(os.environ.get("W", "28"))
batch = synth.regular_batch(B, 1000, 7, seed=1)
assert batch.build_plan()
N = batch.num_nodes
P = 504 // W
Xp = [torch.randn(N, W, device="cuda") for _ in range(P)]      # panel-major: contiguous rows of W floats
Yp = [torch.empty(N, W, device="cuda") for _ in range(P)]
X = ops.padded_empty(N, 500, "cuda"); X.normal_()
Y = ops.padded_empty(N, 500, "cuda")
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
bytes_ = 8.0 * N * 500 + 4.0 * batch.nnz + 4.0 * (N + 1)
ms = timeit(lambda: ops.spmm(batch, X, out=Y))
print(f"row-major  [N,500]: {ms:.3f} ms  frac {bytes_/ms/1e6/6556.5:.3f}")
def panels():
    for p in range(P):
        ops.spmm(batch, Xp[p], out=Yp[p])
ms = timeit(panels)
print(f"panel-major {P}x[N,{W}]: {ms:.3f} ms  frac {bytes_/ms/1e6/6556.5:.3f}")
