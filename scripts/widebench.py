#!/usr/bin/env python
"""Eager training steps on large meshes (streaming path) -- for ncu launch lists / timing.
python scripts/widebench.py --mesh 100 100 --batch 64 --steps 5 [--graph]"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=int, nargs="+", default=[100, 100])
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
from g_adaptivity_b200 import GNN, synth
from g_adaptivity_b200.trainer import DeformerTrainer
md = tuple(a.mesh)
opt = synth.default_opt(md, device="cuda:0", gad_store_alpha=False)
ds = synth.SyntheticDataset(len(md), md)
torch.manual_seed(42)
model = GNN(ds, opt).to("cuda:0")
tr = DeformerTrainer(model, use_cuda_graph=a.graph)
sids = [tr.add_batch(synth.make_batch(md, a.batch, seed=100 + r)) for r in range(2)]
for i in range(4):
    tr.step(sids[i % 2])
tr.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(tr.stream)
for i in range(a.steps):
    tr.step(sids[i % 2])
e1.record(tr.stream)
tr.synchronize()
N = tr.slots[0].N
us = 1e3 * e0.elapsed_time(e1) / a.steps
print(json.dumps({"mesh": list(md), "batch": a.batch, "nodes": N, "step_us": round(us, 2), "gnodes_per_s": round(N / us / 1e3, 3),
                  "wide": tr.slots[0].graph.wide_in is not None, "cluster": (tr.slots[0].graph.cl_C if tr._cluster(tr.slots[0]) else 0)}))
