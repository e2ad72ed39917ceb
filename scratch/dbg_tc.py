import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
import torch
from gmc_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def run(op, M, N, K, kind):
    if op == "nn": sa, sb = (M, K), (K, N)
    elif op == "nt": sa, sb = (M, K), (N, K)
    else: sa, sb = (K, M), (K, N)
    if kind == "ones":
        A, B = torch.ones(sa), torch.ones(sb)
    elif kind == "rowid":   # A(m,k) = m+1 for k==0 else 0 ; B = 1 for k==0
        A, B = torch.zeros(sa), torch.zeros(sb)
        if op in ("nn", "nt"): A[:, 0] = torch.arange(M) + 1.0
        else: A[0, :] = torch.arange(M) + 1.0
        if op in ("nn", "tn"): B[0, :] = torch.arange(N) * 0.001 + 1.0
        else: B[:, 0] = torch.arange(N) * 0.001 + 1.0
    else:
        A, B = torch.randn(sa), torch.randn(sb)
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(op, A.to(dev), B.to(dev), out=out, precision="tf32")
    torch.cuda.synchronize()
    o = out.cpu()
    Ad, Bd = A.double(), B.double()
    want = Ad @ Bd if op == "nn" else Ad @ Bd.t() if op == "nt" else Ad.t() @ Bd
    print(f"--- {op} M={M} N={N} K={K} {kind}: nan={int(torch.isnan(o).sum())} zeros={int((o==0).sum())} "
          f"maxerr={float((o.double()-want).abs().nan_to_num(1e9).max()):.4g}")
    print("   got ", [round(float(x), 3) for x in o[0, :6]], [round(float(x), 3) for x in o[min(M-1, 65), :3]])
    print("   want", [round(float(x), 3) for x in want[0, :6]], [round(float(x), 3) for x in want[min(M-1, 65), :3]])
for op in ("nt", "nn", "tn"):
    for kind in ("ones", "rowid", "rand"):
        run(op, 128, 256, 32, kind)
run("nn", 128, 256, 64, "rand")
run("nn", 256, 512, 32, "rand")
