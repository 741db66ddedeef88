"""Where the host time of the reference-shaped training iteration goes: model(data) + F.l1_loss + backward() + Adam.step()
on a fresh pinned host Batch (cfg 2 shape), cProfile over many iterations."""
import cProfile
import os
import pstats
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from g_adaptivity_b200 import GNN, synth  # noqa: E402

md, B, dev = (30, 30), 256, "cuda:0"
opt = synth.default_opt(md, device=dev, gad_store_alpha=False, gad_shared_topology=True)
ds = synth.SyntheticDataset(2, md)
base = synth.make_batch(md, B, seed=0)
iters = 200
batches = []
for _ in range(iters + 10):
    b = base.clone()
    b.pin_memory()
    batches.append(b)
torch.manual_seed(42)
model = GNN(ds, dict(opt)).to(dev).train()
optim = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
pr = cProfile.Profile()
for k, d in enumerate(batches):
    if k == 10:
        torch.cuda.synchronize()
        pr.enable()
    optim.zero_grad(set_to_none=True)
    F.l1_loss(model(d), d.x_phys.to(dev, non_blocking=True)).backward()
    optim.step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
