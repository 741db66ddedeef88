#!/usr/bin/env python
"""Check + timing of the 2-D FEM kernels (csrc/fem2d.cu) on a B200 -- the script of their first GPU run
(profiles/r01_v11_fem2d_first_gpu_run.log):

    python scripts/fem2d_check.py

Compares gad_fem2d_fwd / gad_fem2d_bwd with the fixtures minted from the reference's difFEM_2d.py
(tests/golden_fem2d) -- forward 1e-5, gradient 5e-5, the bars the host harness of the same arithmetic meets --
and times a batch of 256 meshes of 30x30 at the reference's default quadrature sizes.  The fixture checks are also
tests/test_fem2d_gpu.py."""
import glob, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from g_adaptivity_b200 import fem2d
from oracle import fem2d_fast as Fz

dev = torch.device("cuda:0")
ok = True
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden_fem2d", "fem2d_*.pt"))):
    fx = torch.load(path)
    N, Q = fx["mesh"].shape[0], int(fx["eval_points"])
    topo = fem2d.Fem2DTopology(fx["cells"].numpy(), fx["bc_nodes"].numpy(), N, dev)
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    coords = fx["mesh"].to(dev).unsqueeze(0).clone().requires_grad_(True)
    sol, coeffs, iters = fem2d.fem2d_solve(coords, topo, fx["centers"].unsqueeze(0), fx["scales"].unsqueeze(0),
                                           X.reshape(-1).to(dev), Y.reshape(-1).to(dev), int(fx["load_quad_points"]))
    tgt = Fz.u_true(torch.stack([X, Y], dim=-1), fx["centers"], fx["scales"]).reshape(1, -1).to(dev)
    loss = F.mse_loss(sol, tgt)
    loss.backward()
    sc, sg = fx["coeffs"].abs().max().item(), fx["grad_mesh"].abs().max().item()
    e_c = (coeffs[0].cpu() - fx["coeffs"].reshape(-1)).abs().max().item() / sc
    e_s = (sol[0].cpu() - fx["sol"].reshape(-1)).abs().max().item() / sc
    e_g = (coords.grad[0].cpu() - fx["grad_mesh"]).abs().max().item() / sg
    good = e_c <= 1e-5 and e_s <= 1e-5 and e_g <= 5e-5
    ok &= good
    print(f"{fx['name']}: coeffs {e_c:.2e} sol {e_s:.2e} grad {e_g:.2e} cg {int(iters[0])} {'ok' if good else 'FAIL'}")

# timing: 256 meshes of 30x30, default quadrature sizes
from oracle.ref_harness.make_golden_fem2d import case_inputs
import numpy as np
n, B, Q = 30, 256, 101
cells, bc, pts, centers, scales = case_inputs(n, 2, 0.3, 0)
topo = fem2d.Fem2DTopology(cells, bc, n * n, dev)
x0 = torch.linspace(0, 1, Q)
X, Y = torch.meshgrid(x0, x0, indexing="ij")
coords = torch.tensor(pts, device=dev).unsqueeze(0).repeat(B, 1, 1).requires_grad_(True)
cen = torch.from_numpy(centers).unsqueeze(0).repeat(B, 1, 1)
scl = torch.from_numpy(scales).unsqueeze(0).repeat(B, 1, 1)
for it in range(3):
    torch.cuda.synchronize()
    t = time.time()
    sol, coeffs, iters = fem2d.fem2d_solve(coords, topo, cen, scl, X.reshape(-1).to(dev), Y.reshape(-1).to(dev), 101)
    sol.square().mean().backward()
    torch.cuda.synchronize()
    print(f"256 x 30x30, fwd+bwd: {1e3 * (time.time() - t):.2f} ms, cg iterations {int(iters.max())}")
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
