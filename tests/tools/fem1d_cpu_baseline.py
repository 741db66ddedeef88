#!/usr/bin/env python
"""CPU baseline of the 1-D FEM row: the oracle restatement of torch_FEM_1D (what the reference calls once
per mesh in a Python loop, src/GNN.py:307-327) forward + mse + autograd backward, per mesh.
    python tests/tools/fem1d_cpu_baseline.py [--nodes 200] [--meshes 8]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F

from oracle import fem1d_oracle as F1

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=200)
ap.add_argument("--meshes", type=int, default=8)
a = ap.parse_args()
rng = np.random.default_rng(0)
n, K, Q = a.nodes, 101, 101
t0 = time.perf_counter()
for b in range(a.meshes):
    x = np.linspace(0, 1, n)
    x[1:-1] += (rng.random(n - 2) - 0.5) * 0.3 / (n - 1)
    xb = torch.from_numpy(x.astype(np.float32)).requires_grad_(True)
    cs, ss = [torch.tensor(0.5, dtype=torch.float32)], [torch.tensor(0.1, dtype=torch.float32)]
    _, sol, *_ = F1.torch_fem_1d(xb, torch.linspace(0, 1, Q), cs, ss, load_quad_points=K)
    F.mse_loss(sol, torch.zeros(Q)).backward()
ms = 1e3 * (time.perf_counter() - t0) / a.meshes
print(f"{ms:.2f} ms per mesh of {n} nodes ({os.cpu_count()} host cores, torch {torch.get_num_threads()} threads)")
