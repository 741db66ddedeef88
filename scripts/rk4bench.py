#!/usr/bin/env python
"""Deep RK4 forward (BASELINE cfg 4 shape) on one mesh: module call time per F-evaluation.
python scripts/rk4bench.py --mesh 200 200 --steps 64"""
import argparse, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=int, nargs=2, default=[200, 200])
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--no-cluster", action="store_true", help="streaming wide-row kernels instead of the cluster kernel")
a = ap.parse_args()
from g_adaptivity_b200 import GNN, synth
md = tuple(a.mesh)
opt = synth.default_opt(md, device="cuda:0", gad_store_alpha=False, ode_method="rk4", num_layers=a.steps,
                        gad_no_cluster=a.no_cluster)
ds = synth.SyntheticDataset(2, md)
torch.manual_seed(42)
model = GNN(ds, opt).to("cuda:0").eval()
data = synth.make_batch(md, a.batch, seed=0).to("cuda:0")
with torch.no_grad():
    sess = model.inference_session(data)
    for _ in range(3):
        sess()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        sess()
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
g = model.last_graph if model.last_graph is not None else sess.graph
print({"mesh": md, "batch": a.batch, "cluster": getattr(sess.graph, "clf_C", None), "wide_rows": sess.graph.wide_in is not None, "ms_per_call": round(ms, 4),
       "us_per_feval": round(1e3 * ms / (4 * a.steps), 3)})
