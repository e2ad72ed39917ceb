import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops, synth
B = int(os.environ.get("B", "4096"))
batch = synth.regular_batch(B, 1000, 7, seed=1)
XA = ops.padded_empty_bf16(batch.num_nodes, 1000, "cuda", zero=True)
for _ in range(2): ops.preaggregate_features_bf16(batch, 1000, out=XA)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): ops.preaggregate_features_bf16(batch, 1000, out=XA)
b.record(); torch.cuda.synchronize()
print(f"preaggregate B={B}: {a.elapsed_time(b)/5:.3f} ms")
