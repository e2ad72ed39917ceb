import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops, synth
B = int(os.environ.get("B", "2048"))
batch = synth.regular_batch(B, 1000, 7, seed=1)
N = batch.num_nodes
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
ref = Y.clone()
print(f"row kernel : {ms:.3f} ms  {bytes_/ms/1e6:.0f} GB/s  frac {bytes_/ms/1e6/6556.5:.3f}")
assert batch.build_plan()
ms = timeit(lambda: ops.spmm(batch, X, out=Y))
print(f"slab kernel: {ms:.3f} ms  {bytes_/ms/1e6:.0f} GB/s  frac {bytes_/ms/1e6/6556.5:.3f}  maxdiff {float((Y-ref).abs().max()):.2e}")
