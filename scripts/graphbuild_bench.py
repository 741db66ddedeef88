import sys, time
sys.path.insert(0, '/root/repo')
import torch
from g_adaptivity_b200 import GNN, synth
for md, B in (((30, 30), 256), ((50, 50), 1024), ((100, 100), 64)):
    opt = synth.default_opt(md, device="cuda:0", gad_store_alpha=False)
    ds = synth.SyntheticDataset(2, md)
    model = GNN(ds, opt).to("cuda:0")
    data = synth.make_batch(md, B, seed=0)
    dd = data.to("cuda:0")
    torch.cuda.synchronize()
    for rep in range(3):
        model._graphs.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = model._graph(dd, torch.device("cuda:0"))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    print(md, B, "nodes", g.N, "edges", g.E, "graph build (device-resident inputs): %.2f ms" % (1e3 * (t1 - t0)), "tiles", g.T)
