#!/usr/bin/env python
"""CPU cost of the 2-D differentiable FEM solve per mesh at the reference's default quadrature sizes
(load_quad_points = eval_quad_points = 101, src/params.py:68-70): the line-by-line restatement of torch_FEM_2D
(what the reference loops over per mesh, src/GNN.py:327-335) and the vectorised formulation + hand adjoint.
python tests/tools/fem2d_cpu_baseline.py [n ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import warnings
import torch
from oracle import fem2d_oracle as O, fem2d_fast as Fz
from oracle.ref_harness.make_golden_fem2d import case_inputs

warnings.filterwarnings("ignore")
K = Q = 101
for n in [int(a) for a in sys.argv[1:]] or [11, 15]:
    cells, bc, pts, centers, scales = case_inputs(n, 2, 0.3, 0)
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    mesh = torch.tensor(pts, requires_grad=True)
    c_list = [torch.from_numpy(c.copy()) for c in centers]
    s_list = [torch.from_numpy(s.copy()) for s in scales]
    tgt = O.u_true(torch.stack([X, Y]), c_list, s_list)
    t = time.time()
    coeffs, sol = O.torch_fem_2d(cells, bc, mesh, [X, Y], K, c_list, s_list)
    torch.nn.functional.mse_loss(sol, tgt).backward()
    t1 = time.time() - t
    t = time.time()
    Fz.fem2d_forward_backward(torch.from_numpy(cells), torch.from_numpy(bc), torch.tensor(pts), [X, Y], K,
                              torch.from_numpy(centers), torch.from_numpy(scales), lambda s: 2 * (s - tgt) / s.numel())
    t2 = time.time() - t
    print(f"{n}x{n} mesh, {torch.get_num_threads()} threads: restatement fwd+bwd {t1:.2f} s; vectorised fwd+adjoint {t2:.3f} s")
