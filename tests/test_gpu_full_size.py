"""BASELINE config 3 at its FULL size (4 096 graphs x n = 1000, F = 1000 -> H = 500 -> K = 3) on the benchmarked path
(bf16 operands and activations, pre-aggregated layer 1, fused tail), checked through properties that do not need an oracle
run of that size:
  * STE loss: every per-graph loss is MINUS the integer cut of the hard labels the kernel chose (C oracle-free: the
    integer cut kernel is itself bit-exact against oracle/postproc.c and the reference fixtures);
  * graphs are independent units: the gradients of the whole batch are the sum of the gradients of its two halves, and
    the per-graph losses of a half are the losses the whole batch reported for those graphs -- bit for bit;
  * the same step twice gives the same bits (fixed summation orders everywhere on the path).
The oracle pins the same path at small sizes (tests/test_gpu_bf16_activations.py, tests/test_gpu_split.py)."""
import numpy as np
import pytest
import torch

from gmc_b200 import ops, synth
from gmc_b200.engine import GCNEngine
from gmc_b200.graph import GraphBatch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sub_batch(rowptr, colidx, gp, g0, g1):
    lo, hi = int(gp[g0]), int(gp[g1])
    e0, e1 = int(rowptr[lo]), int(rowptr[hi])
    rp = (rowptr[lo: hi + 1] - e0).astype(np.int32)
    ci = (colidx[e0:e1] - lo).astype(np.int32)
    return GraphBatch.from_arrays(rp, ci, (gp[g0: g1 + 1] - lo).astype(np.int32), device=DEV)


_ARRAYS = {}


def _config3_arrays():
    if "a" not in _ARRAYS:
        _ARRAYS["a"] = synth.regular_batch_arrays(4096, 1000, 7, seed=11)
    return _ARRAYS["a"]


@pytest.mark.parametrize("path", ["bf16_preaggregated", "f16x2"])
def test_config3_full_size_properties_of_the_benchmarked_paths(path):
    """path = the headline configuration of bench.py, or its fp32-grade `parity_grade` configuration."""
    from Training import TrainingNeural as T
    B, n, F = 4096, 1000, 1000
    rowptr, colidx, gp = _config3_arrays()
    batch = GraphBatch.from_arrays(rowptr, colidx, gp, device=DEV)
    cfg = T.TrainingConfig(n_nodes=n, dim_embedding=F, hidden_dim=500, gemm_precision="bf16")
    torch.manual_seed(0)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    if path == "f16x2":
        eng = GCNEngine(net, opt, precision="f16x2", adjacency_features=True)
        features = lambda b: ops.IntegerFeatures.from_batch(b, F, f16=True)
    else:
        eng = GCNEngine(net, opt, precision="bf16", activations="bf16", preaggregate=True)
        features = lambda b: ops.PreaggregatedFeatures(ops.preaggregate_features_bf16(b, F))
    XA = features(batch)
    loss = eng.loss_and_grads(batch, XA).clone()
    grads = eng.grads_flat.clone()
    # (1) loss = -cut(labels), exactly, for every graph
    labels = ops.argmax_labels(batch, eng.P[: batch.num_nodes])
    cut = ops.cut_value(batch, labels)
    assert torch.equal(loss, -cut.to(torch.float64))
    assert 0 < int(cut.min()) and int(cut.max()) <= 3500             # 3 500 edges per graph
    # (3) the same step again: same bits
    loss2 = eng.loss_and_grads(batch, XA)
    assert torch.equal(loss2, loss) and torch.equal(eng.grads_flat, grads)
    del XA
    torch.cuda.empty_cache()
    # (2) halves: per-graph losses carry over bit for bit, gradients add up
    total = torch.zeros_like(grads, dtype=torch.float64)
    for g0, g1 in ((0, B // 2), (B // 2, B)):
        half = _sub_batch(rowptr, colidx, gp, g0, g1)
        XAh = features(half)
        lh = eng.loss_and_grads(half, XAh)
        assert torch.equal(lh, loss[g0:g1])
        total += eng.grads_flat.double()
        del XAh, half
        torch.cuda.empty_cache()
    scale = float(grads.abs().max())
    assert float((total - grads.double()).abs().max()) <= 2e-5 * scale   # fp32 sums over 4.1 M rows in two different splits
