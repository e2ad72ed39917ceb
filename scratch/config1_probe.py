# config 1 (reference pipeline): 20 graphs n=500, per-graph Adam steps through the drop-in API
import sys, os, time, random, io, contextlib
ROOT = os.path.join(os.path.dirname(__file__), "..")
for p in (ROOT, os.path.join(ROOT, "gcn-max-cut_b200"), os.path.join(ROOT, "gcn-max-cut_b200", "python")):
    sys.path.insert(0, p)
import torch
from DataGenerator import GraphCreator as C, graphExtender as E
from Training import TrainingNeural as T
random.seed(0)
graphs = {i: C.generate_graph(n=500, d=random.randint(6, 8), graph_type="reg", random_seed=1000 + i) for i in range(20)}
terms = {i: C.generate_unique_terminals(500, 3) for i in range(20)}
with contextlib.redirect_stdout(io.StringIO()):
    ds = E.process_graphs_from_folder(graphs, terms, max_nodes=1000)
for prec in ("fp32", "tf32", "tf32x3"):
    cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500, learning_rate=1e-3, gemm_precision=prec)
    torch.manual_seed(0)
    net, embed, opt = T.setup_model_and_optimizer(cfg)
    for _ in range(3):
        T.train_single_epoch(ds, net, opt, embed, cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E_ = 20
    for _ in range(E_):
        loss = T.train_single_epoch(ds, net, opt, embed, cfg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{prec}: {len(ds) * E_ / dt:.0f} graph-epochs/s, {1e3 * dt / (len(ds) * E_):.3f} ms per graph step, loss {loss:.1f}")
