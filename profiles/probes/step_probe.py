"""One training step per precision at 1024 graphs (for ncu captures of the GEMM kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import torch
from gmc_b200 import ops, synth
from gmc_b200.engine import GCNEngine
from gmc_b200.graph import GraphBatch
from Training import TrainingNeural as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16", "f16x2", "bf16x3"]
rowptr, colidx, gp = synth.regular_batch_arrays(B, 1000, 7, seed=3)
batch = GraphBatch.from_arrays(rowptr, colidx, gp, device="cuda")
cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500)
for prec in precs:
    torch.manual_seed(0)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    if prec == "bf16":
        eng = GCNEngine(net, opt, precision="bf16", activations="bf16", preaggregate=True)
        feats = ops.PreaggregatedFeatures(ops.preaggregate_features_bf16(batch, 1000))
    else:
        eng = GCNEngine(net, opt, precision=prec, adjacency_features=True)
        feats = ops.IntegerFeatures.from_batch(batch, 1000, f16=prec == "f16x2")
    for _ in range(2):
        loss = eng.train_step(batch, feats)
    torch.cuda.synchronize()
    print(prec, float(loss.sum()))
    del eng, feats
    torch.cuda.empty_cache()
