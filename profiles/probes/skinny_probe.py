import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import torch
from gmc_b200 import ops
from gmc_b200.ops import Workspace
dev = "cuda"
N, H, K = 4096 * 1000, 500, 3
H1 = ops.padded_empty_bf16(N, H, dev, zero=True); H1.normal_().relu_()
dH = ops.padded_empty_bf16(N, H, dev, zero=True)
dT2 = torch.randn(N, K, device=dev); W2 = torch.randn(H, K, device=dev) * 0.1
dW2 = torch.empty(H, K, device=dev); db1 = torch.empty(H, device=dev)
ws = Workspace()
def f(): ops.skinny_bwd_bf16(dT2, W2, H1, dH=dH, dW=dW2, dbias=db1, workspace=ws)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("GMC_LIB", "default"), "skinny_bwd bf16 %.3f ms" % (e0.elapsed_time(e1) / 10), float(dW2.abs().sum()), float(db1.abs().sum()))
