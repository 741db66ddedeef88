#!/usr/bin/env python
"""Phase timeline of the cluster-resident train kernel (Args::trace marks of k_cl_train):
entry | inputs + first cluster barrier | forward + loss | backward.   python scripts/cltrace.py --mesh 100 100 --batch 64"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=int, nargs="+", default=[100, 100])
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
from g_adaptivity_b200 import GNN, _lib, synth
from g_adaptivity_b200.trainer import DeformerTrainer
md = tuple(a.mesh)
opt = synth.default_opt(md, device="cuda:0", gad_store_alpha=False)
ds = synth.SyntheticDataset(len(md), md)
torch.manual_seed(42)
model = GNN(ds, opt).to("cuda:0")
tr = DeformerTrainer(model, use_cuda_graph=False)
sids = [tr.add_batch(synth.make_batch(md, a.batch, seed=100 + r)) for r in range(2)]
lib = _lib.load()
trace = torch.zeros((2048, 64), dtype=torch.int64, device="cuda:0")
g = tr.slots[0].graph
with torch.cuda.stream(tr.stream):
    for rep in range(4):
        for sid in sids:
            s = tr.slots[sid]
            d = tr._train_desc(s, 2)
            if rep == 3 and sid == sids[-1]:
                trace.zero_()
                d.trace = trace.data_ptr()
            _lib.check(lib.gad_train_step_cluster(C.byref(d), g.cl_C, tr.stream.cuda_stream), "train")
tr.synchronize()
t = trace.cpu().numpy()
grid = int((t[:, 0] > 0).sum())
t = t[:grid]
t0 = t[:, 0].min()
print(f"grid {grid} CTAs (clusters of {g.cl_C}, slab {g.cl_S} nodes)")
print("entry spread (us): med %.2f max %.2f" % (np.median(t[:, 0] - t0) / 1e3, (t[:, 0].max() - t0) / 1e3))
for k, n in enumerate(["entry -> inputs + first barrier", "forward + loss", "backward"]):
    dd = (t[:, k + 1] - t[:, k]) / 1e3
    print(f"  {n:32s} med {np.median(dd):7.2f}  p10 {np.percentile(dd, 10):7.2f}  p90 {np.percentile(dd, 90):7.2f} us")
print("CTA end (us after first entry): med %.2f max %.2f" % (np.median(t[:, 3] - t0) / 1e3, (t[:, 3].max() - t0) / 1e3))
