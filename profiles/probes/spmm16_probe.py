import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import torch
from gmc_b200 import ops, synth
from gmc_b200.graph import GraphBatch
rowptr, colidx, gp = synth.regular_batch_arrays(4096, 1000, 7, seed=3)
batch = GraphBatch.from_arrays(rowptr, colidx, gp, device="cuda")
X = ops.padded_empty_bf16(batch.num_nodes, 500, "cuda", zero=True); X.normal_()
Y = ops.padded_empty_bf16(batch.num_nodes, 500, "cuda", zero=True)
ops.spmm_bf16(batch, X, out=Y); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.spmm_bf16(batch, X, out=Y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ref = ops.spmm(batch, X.float()[:, :500].contiguous())
err = float((Y.float() - ref).abs().max())
print(os.environ.get("GMC_SLAB16_VARIANT", "default"), "spmm bf16 %.3f ms  %.2f TB/s  frac %.3f  maxerr %.4f" % (ms, 8.323e9 / ms / 1e9, 8.323e9 / ms / 1e9 / 6.5565, err))
