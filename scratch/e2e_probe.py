"""Phase timing of bench.py's end-to-end step (CUDA events + host clock) to see where the non-kernel time goes."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gcn-max-cut_b200", "python"))
import torch
from gmc_b200 import ops, synth
from gmc_b200.engine import GCNEngine
from gmc_b200.graph import GraphBatch
from Training import TrainingNeural as T
B, n, F, H = int(os.environ.get("B", "4096")), 1000, 1000, 500
rowptr, colidx, gp = synth.regular_batch_arrays(B, n, 7, seed=0)
h = [torch.from_numpy(a).pin_memory() for a in (rowptr, colidx, gp)]
batch = GraphBatch.from_arrays(rowptr, colidx, gp, device="cuda")
N = batch.num_nodes
cfg = T.TrainingConfig(n_nodes=n, dim_embedding=F, hidden_dim=H, gemm_precision="bf16", batch_graphs=B)
net, embed, opt = T.setup_model_and_optimizer(cfg)
PRE = os.environ.get("PREAGG", "1") == "1"
eng = GCNEngine(net, opt, precision="bf16", activations="bf16", preaggregate=PRE)
X = ops.padded_empty_bf16(N, F, "cuda", zero=True)
d = [torch.empty_like(batch.rowptr), torch.empty_like(batch.colidx), torch.empty_like(batch.graph_ptr)]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e = [ev()]
    for a, b in zip(d, h): a.copy_(b, non_blocking=True)
    e.append(ev())
    b2 = GraphBatch.__new__(GraphBatch)
    b2.device, b2.num_graphs, b2.sizes, b2.num_nodes, b2.nnz = batch.device, B, batch.sizes, N, batch.nnz
    b2.rowptr, b2.colidx, b2.graph_ptr, b2.max_nodes = d[0], d[1], d[2], batch.max_nodes
    b2.unit_weights, b2.wts_f32, b2.wts_i32, b2.integer_weights = True, None, None, True
    b2.norm, _z = ops.degree_norm(d[0], N); e.append(ev())
    b2.coef = ops.edge_coef(d[0], d[1], None, b2.norm, b2.norm, N); e.append(ev())
    if PRE:
        b2.plan = None; e.append(ev())
        ops.preaggregate_features_bf16(b2, F, out=X); e.append(ev())
        loss = eng.train_step(b2, ops.PreaggregatedFeatures(X)); e.append(ev())
        e.append(ev())
    else:
        b2.plan = ops.spmm_plan(d[0], d[1], b2.norm, b2.norm, d[2], B, N); e.append(ev())
        ops.scatter_features_bf16(b2, F, X); e.append(ev())
        loss = eng.train_step(b2, X); e.append(ev())
        ops.scatter_features_bf16(b2, F, X, clear=True); e.append(ev())
    out = loss.cpu(); e.append(ev())
    torch.cuda.synchronize(); t1 = time.perf_counter()
    names = ["h2d", "norm+sync", "coef", "plan+sync", "preagg" if PRE else "scatter", "train_step", "clear", "d2h"]
    print(f"iter {it}: wall {1e3*(t1-t0):.2f} ms | " + " ".join(f"{nm}={e[i].elapsed_time(e[i+1]):.2f}" for i, nm in enumerate(names)), flush=True)
