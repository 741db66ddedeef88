#!/usr/bin/env python
"""Where the time of the reference's loop shape goes on the module seam (cfg 2, a fresh pinned host Batch per
iteration): model(data) -> F.l1_loss -> backward() -> torch.optim.Adam.step().   python scripts/seam_profile.py"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from g_adaptivity_b200 import GNN, synth

dev = torch.device("cuda", 0)
md, B = (30, 30), 256
opt = synth.default_opt(md)
opt.update(device="cuda:0", gad_store_alpha=False, gad_shared_topology=True)
ds = synth.SyntheticDataset(2, md)
torch.manual_seed(42)
model = GNN(ds, opt).to(dev).train()
optim = torch.optim.Adam(model.parameters(), lr=1e-3, fused=("--fused" in sys.argv))
base = synth.make_batch(md, B, seed=0)
base.pin_memory()
batches = []
for _ in range(40):
    b = base.clone()
    b.pin_memory()
    batches.append(b)


def step(data, marks=None):
    def mark(name):
        if marks is not None:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))
    mark("start")
    optim.zero_grad(set_to_none=True)
    out = model(data)
    mark("forward")
    loss = F.l1_loss(out, data.x_phys.to(dev, non_blocking=True))
    mark("loss")
    loss.backward()
    mark("backward")
    optim.step()
    mark("adam")


for d in batches[:5]:
    step(d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for d in batches[5:25]:
    step(d)
torch.cuda.synchronize()
print(f"loop: {1e3 * (time.perf_counter() - t0) / 20:.3f} ms per step (async)")
acc = {}
for d in batches[25:35]:
    marks = []
    step(d, marks)
    for (n0, t0_), (n1, t1_) in zip(marks[:-1], marks[1:]):
        acc[n1] = acc.get(n1, 0.0) + (t1_ - t0_)
print("synchronised phases (ms):", {k: round(1e3 * v / 10, 3) for k, v in acc.items()})
pr = cProfile.Profile()
pr.enable()
for d in batches[35:40]:
    step(d)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

# the same loop shape on the fused step: trainer.train_batch(data)
from g_adaptivity_b200.trainer import DeformerTrainer
torch.manual_seed(42)
model2 = GNN(ds, dict(opt)).to(dev).train()
tr = DeformerTrainer(model2)
for d in batches[:6]:
    tr.train_batch(d)
tr.synchronize()
t0 = time.perf_counter()
for d in batches[6:36]:
    loss = tr.train_batch(d)
tr.synchronize()
print(f"trainer.train_batch loop: {1e3 * (time.perf_counter() - t0) / 30:.3f} ms per step, loss {float(loss.item()):.5f}")
