"""Can skinny_bwd (HBM-bound) run concurrently with the weight-gradient GEMM (tensor-bound)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import torch
from gmc_b200 import ops
from gmc_b200.ops import Workspace

dev = "cuda"
N, F, H, K = 4096 * 1000, 1000, 500, 3
half = N // 2
XA = ops.padded_empty_bf16(N, F, dev, zero=True); XA.bernoulli_(0.05)
H1 = ops.padded_empty_bf16(N, H, dev, zero=True); H1.normal_().relu_()
dH = ops.padded_empty_bf16(N, H, dev, zero=True); dH.normal_()
dH2 = ops.padded_empty_bf16(N, H, dev, zero=True)
dT2 = torch.randn(N, K, device=dev)
W2 = torch.randn(H, K, device=dev) * 0.1
gW1 = torch.empty(F, H, device=dev)
dW2 = torch.empty(H, K, device=dev); db1 = torch.empty(H, device=dev)
ws_a, ws_b = Workspace(), Workspace()
s2 = torch.cuda.Stream()

def t_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def gemm(rows=slice(0, N), acc=False): ops.gemm_bf16("tn", XA[rows], dH[rows], out=gW1, accumulate=acc, workspace=ws_a)
def skinny(rows=slice(0, N), out=dH2): ops.skinny_bwd_bf16(dT2[rows], W2, H1[rows], dH=out[rows], dW=dW2, dbias=db1, workspace=ws_b)

print("gemm tn alone      %.3f ms" % t_ms(gemm))
print("skinny alone       %.3f ms" % t_ms(skinny))
print("sequential         %.3f ms" % t_ms(lambda: (skinny(), gemm())))

def concurrent():
    ev = torch.cuda.Event(); ev.record()
    with torch.cuda.stream(s2):
        s2.wait_event(ev)
        skinny()
        done = torch.cuda.Event(); done.record()
    gemm()
    torch.cuda.current_stream().wait_event(done)
print("concurrent (indep) %.3f ms" % t_ms(concurrent))

def chunked(C):
    # skinny(c) on s2, gemm(c) on main after skinny(c): skinny(c+1) overlaps gemm(c); real dependency through dH2
    step = (N // C + 127) // 128 * 128
    ev0 = torch.cuda.Event(); ev0.record()
    evs = []
    with torch.cuda.stream(s2):
        s2.wait_event(ev0)
        for c in range(C):
            r = slice(c * step, min(N, (c + 1) * step))
            ops.skinny_bwd_bf16(dT2[r], W2, H1[r], dH=dH2[r], dW=dW2, dbias=db1, workspace=ws_b)
            e = torch.cuda.Event(); e.record(); evs.append(e)
    for c in range(C):
        r = slice(c * step, min(N, (c + 1) * step))
        torch.cuda.current_stream().wait_event(evs[c])
        ops.gemm_bf16("tn", XA[r], dH2[r], out=gW1, accumulate=c > 0, workspace=ws_a)
for C in (2, 4, 8):
    print("chunked x%d         %.3f ms" % (C, t_ms(lambda: chunked(C))))
