"""GPU parity of the 2-D FEM kernels (csrc/fem2d.cu through the C ABI gad_fem2d_fwd / gad_fem2d_bwd and
g_adaptivity_b200/fem2d.py) against the fixtures minted from the reference's own firedrake_difFEM/difFEM_2d.py
(tests/golden_fem2d, oracle/ref_harness/make_golden_fem2d.py): coefficients and solution 1e-5, gradient with respect
to the mesh points 5e-5 -- the uniform mesh included, where every grid point lies on an element edge and the
result depends on reproducing the reference's tie decisions."""
import glob
import os

import pytest
import torch
import torch.nn.functional as F

from g_adaptivity_b200 import fem2d
from oracle import fem2d_fast as Fz

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden_fem2d", "fem2d_*.pt")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[6:-3] for p in GOLDEN])
def test_cuda_fem2d_matches_reference_fixture(path):
    dev = torch.device("cuda:0")
    fx = torch.load(path)
    N, Q = fx["mesh"].shape[0], int(fx["eval_points"])
    topo = fem2d.Fem2DTopology(fx["cells"].numpy(), fx["bc_nodes"].numpy(), N, dev)
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    B = 3                                                   # the same mesh three times: every CTA must agree bit for bit
    coords = fx["mesh"].to(dev).unsqueeze(0).repeat(B, 1, 1).clone().requires_grad_(True)
    cen = fx["centers"].unsqueeze(0).repeat(B, 1, 1)
    scl = fx["scales"].unsqueeze(0).repeat(B, 1, 1)
    sol, coeffs, iters = fem2d.fem2d_solve(coords, topo, cen, scl, X.reshape(-1).to(dev), Y.reshape(-1).to(dev),
                                           int(fx["load_quad_points"]))
    tgt = Fz.u_true(torch.stack([X, Y], dim=-1), fx["centers"], fx["scales"]).reshape(1, -1).to(dev)
    F.mse_loss(sol[0:1], tgt).backward()
    sc, sg = fx["coeffs"].abs().max().item(), fx["grad_mesh"].abs().max().item()
    assert (coeffs[0].cpu() - fx["coeffs"].reshape(-1)).abs().max().item() <= 1e-5 * sc
    assert (sol[0].cpu() - fx["sol"].reshape(-1)).abs().max().item() <= 1e-5 * sc
    assert (coords.grad[0].cpu() - fx["grad_mesh"]).abs().max().item() <= 5e-5 * sg
    assert torch.equal(coeffs[1], coeffs[0]) and torch.equal(sol[2], sol[0]) and int(iters[0]) == int(iters[1]) > 0
    assert coords.grad[1].abs().max().item() == 0.0        # no cotangent reached meshes 1, 2


def test_fem2d_needs_cuda():
    fx = torch.load(GOLDEN[0])
    with pytest.raises(RuntimeError):
        fem2d.FEM2DFunction.apply(fx["mesh"].unsqueeze(0), None, fx["centers"].unsqueeze(0), fx["scales"].unsqueeze(0),
                                  torch.zeros(4), torch.zeros(4), 9)
