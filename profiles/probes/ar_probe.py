import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
from gmc_b200 import dist as gdist
rank, local, world = gdist.init_from_env()
dev = torch.device("cuda", local)
n = 502004
peer = gdist.PeerAllReduce(n, dev)
t_nccl = torch.randn(n, device=dev)
def bench(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1000 * e0.elapsed_time(e1) / reps
a = bench(lambda: dist.all_reduce(t_nccl))
b = bench(lambda: peer.all_reduce_())
# correctness: ranks contribute rank+1
peer.tensor.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
peer.all_reduce_(); torch.cuda.synchronize()
ok = bool((peer.tensor[:n] == world * (world + 1) / 2).all())
res = torch.tensor([a, b, float(ok)], device=dev, dtype=torch.float64)
out = [torch.empty_like(res) for _ in range(world)]
dist.all_gather(out, res)
if rank == 0:
    print("world", world, "NCCL us/call", [round(float(o[0]), 1) for o in out])
    print("peer  us/call", [round(float(o[1]), 1) for o in out], "correct", [bool(o[2]) for o in out])
