import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import torch
from gmc_b200 import ops, synth
from gmc_b200.graph import GraphBatch
B = 4096
rowptr, colidx, gp = synth.regular_batch_arrays(B, 1000, 7, seed=3)
batch = GraphBatch.from_arrays(rowptr, colidx, gp, device="cuda")
out = ops.padded_empty_bf16(batch.num_nodes, 1000, "cuda", zero=True)
def t_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("preaggregate bf16 %.3f ms" % t_ms(lambda: ops.preaggregate_features_bf16(batch, 1000, out=out)))
print("integer features  %.3f ms" % t_ms(lambda: ops.integer_features_bf16(batch, 1000, out=out)))
print("degree_norm+coef  %.3f ms" % t_ms(lambda: (ops.degree_norm(batch.rowptr, batch.num_nodes, count_zero=False, out=batch.norm), ops.edge_coef(batch.rowptr, batch.colidx, None, batch.norm, batch.norm, batch.num_nodes, out=batch.coef))))
print("memset 8.2 GB     %.3f ms" % t_ms(lambda: out.zero_()))
chk = out.float().sum().item(); print(chk)
flat = torch.empty(batch.num_nodes * 1024, dtype=torch.bfloat16, device="cuda")
print("contiguous fill 8.4 GB  %.3f ms" % t_ms(lambda: flat.zero_()))
import ctypes
from gmc_b200 import _lib
print("cudaMemsetAsync 8.4 GB  %.3f ms" % t_ms(lambda: torch.cuda.cudart().cudaMemset(flat.data_ptr(), 0, flat.numel() * 2) if False else flat.fill_(1.0)))
src = torch.empty_like(flat)
print("copy 8.4 GB -> 8.4 GB   %.3f ms" % t_ms(lambda: flat.copy_(src)))
