"""bf16 GEMMs of the headline step at config-3 shape: cluster kernel vs cta_group::2 (GMC_GEMM_BF16_2CTA); scratch."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-max-cut_b200"), os.path.join(ROOT, "gcn-max-cut_b200", "python")):
    sys.path.insert(0, p)
import torch
from gmc_b200 import ops, synth
from gmc_b200.graph import GraphBatch

B = int(os.environ.get("PROBE_GRAPHS", "4096"))
n, F, H = 1000, 1000, 500
rowptr, colidx, gp = synth.regular_batch_arrays(B, n, 7, seed=0)
batch = GraphBatch.from_arrays(rowptr, colidx, gp, device="cuda")
N = batch.num_nodes
XA = ops.preaggregate_features_bf16(batch, F)

def timed(fn, k=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k

torch.manual_seed(0)
W1 = torch.randn(F, H, device="cuda") * 0.03
b1 = torch.randn(H, device="cuda") * 0.05
W2 = torch.randn(H, 3, device="cuda") * 0.1
W1b = ops.to_bf16(W1)
w2p = ops.pad_proj_weights(W2)
H16 = ops.padded_empty_bf16(N, H, "cuda", zero=True)
T2 = torch.empty(N, 3, device="cuda")
ws = ops.Workspace()
gW1 = torch.empty(F, H, device="cuda")
flops = 2.0 * N * F * H
# interleaved A/B of the two kernels (the library reads GMC_GEMM_BF16_2CTA per call): clock / power state is shared
ab = {"nn_layer1": {"0": [], "nn": []}, "nn_plain": {"0": [], "nn": []}, "tn": {"0": [], "tn": []}}
for rep in range(4):
    for mode in ("0", "nn"):
        os.environ["GMC_GEMM_BF16_2CTA"] = mode
        ab["nn_layer1"][mode].append(timed(lambda: ops.gemm_bf16_bf16out("nn", XA, W1b, out=H16, bias=b1, relu=True, proj_w=w2p, proj_out=T2, n_proj=3), k=6, warm=1))
        ab["nn_plain"][mode].append(timed(lambda: ops.gemm_bf16_bf16out("nn", XA, W1b, out=H16), k=6, warm=1))
    for mode in ("0", "tn"):
        os.environ["GMC_GEMM_BF16_2CTA"] = mode
        ab["tn"][mode].append(timed(lambda: ops.gemm_bf16("tn", XA, H16, out=gW1, workspace=ws), k=6, warm=1))
os.environ["GMC_GEMM_BF16_2CTA"] = os.environ.get("PROBE_MODE", "0")
print(json.dumps({"interleaved_ms": ab}))
out = {"mode": os.environ.get("GMC_GEMM_BF16_2CTA", "0")}
ms = timed(lambda: ops.gemm_bf16_bf16out("nn", XA, W1b, out=H16, bias=b1, relu=True, proj_w=w2p, proj_out=T2, n_proj=3))
out["nn_layer1_ms"], out["nn_layer1_tflops"] = ms, flops / ms / 1e9
ms = timed(lambda: ops.gemm_bf16_bf16out("nn", XA, W1b, out=H16))
out["nn_plain_ms"], out["nn_plain_tflops"] = ms, flops / ms / 1e9
ms = timed(lambda: ops.gemm_bf16("tn", XA, H16, out=gW1, workspace=ws))
out["tn_ms"], out["tn_tflops"] = ms, flops / ms / 1e9
# correctness spot check against float64 on a slice
ref = torch.relu(XA[:4096].double() @ W1b.double() + b1.double())
ops.gemm_bf16_bf16out("nn", XA, W1b, out=H16, bias=b1, relu=True, proj_w=w2p, proj_out=T2, n_proj=3)
out["nn_max_err_vs_f64"] = float((H16[:4096].double() - ref).abs().max())
out["proj_max_err"] = float((T2[:4096].double() - H16[:4096].double() @ W2.double()).abs().max())
ops.gemm_bf16("tn", XA, H16, out=gW1, workspace=ws)
ref_tn = XA[:200000].double().t() @ H16[:200000].double()
g_small = ops.gemm_bf16("tn", XA[:200000], H16[:200000], workspace=ws)
out["tn_rel_err_200k"] = float((g_small.double() - ref_tn).abs().max() / ref_tn.abs().max())
print(json.dumps(out))
