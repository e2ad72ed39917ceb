import cProfile, pstats, io, os, sys, random, contextlib, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/gcn-max-cut_b200", ROOT + "/gcn-max-cut_b200/python"):
    sys.path.insert(0, p)
import numpy as np, torch
from DataGenerator import GraphCreator as C, graphExtender as E
from Testing import TestingNeuralNetwork as Te
from Training import TrainingNeural as T
random.seed(0)
sizes = [50, 100, 200, 300, 500]
graphs, terms = {}, {}
i = 0
for n in sizes:
    for j in range(10):
        graphs[i] = C.generate_graph(n=n, d=random.randint(6, 8), graph_type="reg", random_seed=n * 1000 + j)
        terms[i] = C.generate_unique_terminals(n, 3)
        i += 1
with contextlib.redirect_stdout(io.StringIO()):
    ds = E.process_graphs_from_folder(graphs, terms, max_nodes=1000)
cfg = T.TrainingConfig(n_nodes=1000, dim_embedding=1000, hidden_dim=500)
torch.manual_seed(0)
net, embed, opt = T.setup_model_and_optimizer(cfg)
def one():
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        return Te.test_multiple_graphs_batched(net, ds, sizes, 200, verbose=False)
for _ in range(3): one()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): one()
torch.cuda.synchronize()
print("ms per pass", (time.perf_counter() - t0) * 100)
t0 = time.perf_counter(); x = np.random.rand(200 * sum(n - 3 for n in sizes) * 10); print("rand ms", (time.perf_counter() - t0) * 1e3, x.size)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): one()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4500])
