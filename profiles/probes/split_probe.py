"""Times the split-operand GEMMs and the parity-grade engine step at config-3 shape (scratch; not product code)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-max-cut_b200"), os.path.join(ROOT, "gcn-max-cut_b200", "python")):
    sys.path.insert(0, p)
import torch
from gmc_b200 import ops, synth
from gmc_b200.engine import GCNEngine, OpTimer
from gmc_b200.graph import GraphBatch
from Training import TrainingNeural as T

B = int(os.environ.get("PROBE_GRAPHS", "4096"))
n, F, H = 1000, 1000, 500
rowptr, colidx, gp = synth.regular_batch_arrays(B, n, 7, seed=0)
batch = GraphBatch.from_arrays(rowptr, colidx, gp, device="cuda")
N = batch.num_nodes

def timed(fn, k=5, warm=2):
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

out = {}
xi = ops.IntegerFeatures.from_batch(batch, F)
W1 = torch.randn(F, H, device="cuda") * 0.03
b1 = torch.randn(H, device="cuda") * 0.05
Hbuf = ops.padded_empty(N, H, "cuda")
flops = 2.0 * N * F * H
for ns in (2, 3):
    Ws = ops.f32_split_bf16(W1, ns)
    ms = timed(lambda: ops.gemm_bf16_split("nn", xi.tensor, Ws, ns, F, out=Hbuf, row_scale=xi.scale, bias=b1, relu=True))
    out[f"nn_split{ns}_ms"] = ms
    out[f"nn_split{ns}_tflops_eff"] = flops / ms / 1e9
    for cl in (("1", "2", "4") if ns == 2 else ("1", "3")):
        os.environ["GMC_GEMM_SPLIT_CLUSTER"] = cl
        ms = timed(lambda: ops.gemm_bf16_split("nn", xi.tensor, Ws, ns, F, out=Hbuf, row_scale=xi.scale, bias=b1, relu=True))
        out[f"nn_split{ns}_cl{cl}_ms"] = ms
        del os.environ["GMC_GEMM_SPLIT_CLUSTER"]
W1b = ops.to_bf16(W1)
H16 = ops.padded_empty_bf16(N, H, "cuda", zero=True)
out["nn_plain_bf16_ms"] = timed(lambda: ops.gemm_bf16_bf16out("nn", xi.tensor, W1b, out=H16, bias=b1, relu=True))
ws = ops.Workspace()
gW1 = torch.empty(F, H, device="cuda")
for ns in (2, 3):
    S = ops.split_empty(N, H, ns, "cuda")
    S.normal_()
    ms = timed(lambda: ops.gemm_bf16_split("tn", xi.tensor, S, ns, N, out=gW1, workspace=ws))
    out[f"tn_split{ns}_ms"] = ms
    for cl in (("1", "2", "4") if ns == 2 else ("1", "3")):
        os.environ["GMC_GEMM_SPLIT_CLUSTER"] = cl
        out[f"tn_split{ns}_cl{cl}_ms"] = timed(lambda: ops.gemm_bf16_split("tn", xi.tensor, S, ns, N, out=gW1, workspace=ws))
        del os.environ["GMC_GEMM_SPLIT_CLUSTER"]
    del S
out["tn_plain_bf16_ms"] = timed(lambda: ops.gemm_bf16("tn", xi.tensor, H16, out=gW1, workspace=ws))
dT2 = torch.randn(N, 3, device="cuda")
W2 = torch.randn(H, 3, device="cuda")
for ns in (2, 3):
    S = ops.split_empty(N, H, ns, "cuda")
    out[f"skinny_bwd_split{ns}_ms"] = timed(lambda: ops.skinny_bwd_split(dT2, W2, Hbuf, ns, S, row_scale=xi.scale, workspace=ws))
    del S
out["skinny_fwd_f32_ms"] = timed(lambda: ops.skinny_fwd(Hbuf, W2))
del Hbuf, H16
torch.cuda.empty_cache()
for prec in ("bf16x3", "bf16x2"):
    cfg = T.TrainingConfig(n_nodes=n, dim_embedding=F, hidden_dim=H, gemm_precision=prec, batch_graphs=B)
    torch.manual_seed(0)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    eng = GCNEngine(net, opt, precision=prec, adjacency_features=True)
    for _ in range(3):
        eng.train_step(batch, xi)
    torch.cuda.synchronize()
    eng.timer = OpTimer()
    ms = timed(lambda: eng.train_step(batch, xi), k=10, warm=0)
    eng.timer.collect()
    out[f"step_{prec}_ms"] = ms
    out[f"step_{prec}_graph_epochs_per_s"] = B / ms * 1e3
    out[f"step_{prec}_ops"] = {k: eng.timer.total_ms[k] / eng.timer.calls[k] for k in eng.timer.total_ms}
    del eng, net, opt
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
