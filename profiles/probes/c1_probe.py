import contextlib, io, random, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import torch
from DataGenerator import GraphCreator as C, graphExtender as E
from Training import TrainingNeural as T

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
random.seed(0)
n_graphs = 20
graphs = {i: C.generate_graph(n=500, d=random.randint(6, 8), graph_type="reg", random_seed=1000 + i) for i in range(n_graphs)}
terms = {i: C.generate_unique_terminals(500, 3) for i in range(n_graphs)}
with contextlib.redirect_stdout(io.StringIO()):
    ds = E.process_graphs_from_folder(graphs, terms, max_nodes=1000)
cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500, learning_rate=1e-3, gemm_precision=prec)
torch.manual_seed(0)
net, embed, opt = T.setup_model_and_optimizer(cfg)
for _ in range(4):
    T.train_single_epoch(ds, net, opt, embed, cfg)
torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
for _ in range(K):
    T.train_single_epoch(ds, net, opt, embed, cfg)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"API loop: {1e6 * dt / (K * n_graphs):.1f} us per graph step")
eng = T._engine_for(net, opt, cfg)
entry = [e for e in eng._epoch_graphs.values() if e["graph"] is not None][0]
print("epoch graph launches:", entry["launches"], "items", len(entry["items"]))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ev0.record()
for _ in range(K):
    entry["graph"].replay()
ev1.record()
torch.cuda.synchronize()
print(f"epoch graph replay: device {1e3 * ev0.elapsed_time(ev1) / (K * n_graphs):.1f} us per graph step")
