"""GPU parity of the 2-D FEM kernels (csrc/fem2d.cu through the C ABI gad_fem2d_fwd / gad_fem2d_bwd and
g_adaptivity_b200/fem2d.py) against the fixtures minted from the reference's own firedrake_difFEM/difFEM_2d.py
(tests/golden_fem2d, oracle/ref_harness/make_golden_fem2d.py): coefficients and solution 1e-5, gradient with respect
to the mesh points 5e-5 -- the uniform mesh included, where every grid point lies on an element edge and the
result depends on reproducing the reference's tie decisions."""
import glob
import os

import numpy as np
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


@pytest.mark.parametrize("n,Q,K,B", [(7, 13, 9, 3), (6, 21, 9, 2)])
def test_pde_loss_through_the_deformer_2d(n, Q, K, B):
    """`loss_type='pde_loss'` on 2-D meshes -- the reference's default loss (params.py:109; src/GNN.py:307-342,
    src/run_GNN.py:109-110): the model returns (coeffs, x_phys, sol) with `sol` in the fine mesh's node order
    (`reshape_grid_to_fd_tensor` with `dataset.mapping_tensor_fine`, GNN.py:333); mse(sol, u_true_fine)
    back-propagates through the batched FEM solve AND the deformer into the Linear parameters.
    Against the oracle deformer (fp32, as the reference runs) + the oracle's 2-D FEM per mesh + autograd."""
    import copy
    from g_adaptivity_b200 import GNN, synth
    from oracle import gnn_oracle
    md = (n, n)
    opt = synth.default_opt(md, eval_quad_points=Q, load_quad_points=K)
    ds = synth.SyntheticDataset(2, md, eval_quad_points=Q)
    data = synth.make_batch(md, B, seed=5, eval_quad_points=Q, with_u_true_fine=True)
    torch.manual_seed(42)
    ref = gnn_oracle.GNNRef(ds, copy.deepcopy(opt))
    xp = ref(data)                                             # [B*n*n, 2]
    topo = synth.MeshTopology(md)
    bc = np.nonzero(topo.boundary_nodes)[0]
    x0 = torch.linspace(0, 1, Q)
    X, Y = torch.meshgrid(x0, x0, indexing="ij")
    _, order = torch.sort(ds.mapping_tensor_fine)
    sols, coefs = [], []
    for b in range(B):
        cen = torch.from_numpy(np.stack(data.pde_params["centers"][b])).float()
        scl = torch.from_numpy(np.stack(data.pde_params["scales"][b])).float()
        c, s = Fz.fem2d_fast(topo.cells, bc, xp[b * n * n:(b + 1) * n * n], [X, Y], K, cen, scl)
        coefs.append(c)
        sols.append(s.reshape(-1)[order])                      # reshape_grid_to_fd_tensor (utils_data.py:143-159)
    sol_ref, coef_ref = torch.cat(sols), torch.cat(coefs)
    loss_ref = F.mse_loss(sol_ref, data.u_true_fine_tensor)
    loss_ref.backward()

    gopt = copy.deepcopy(opt)
    gopt.update(device="cuda", loss_type="pde_loss")
    model = GNN(ds, gopt).to("cuda")
    model.load_state_dict(ref.state_dict())
    model.train()
    coeffs, x_phys, sol = model(data)
    assert coeffs.shape == (B * n * n, 1) and x_phys.shape == (B * n * n, 2) and sol.shape == (B * Q * Q,)
    x_phys.retain_grad()
    loss = F.mse_loss(sol, data.u_true_fine_tensor.cuda())
    loss.backward()
    # ---- forward: deformer + FEM against the oracle chain
    sc = coef_ref.abs().max().item()
    assert (x_phys.detach().cpu() - xp.detach()).abs().max().item() <= 1e-5 * xp.abs().max().item()
    assert (coeffs.detach().cpu() - coef_ref.detach()).abs().max().item() <= 2e-5 * sc
    assert (sol.detach().cpu() - sol_ref.detach()).abs().max().item() <= 2e-5 * sc
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    # ---- backward, link by link.  The reference's FEM gradient is a discontinuous function of the mesh points
    # (hat-function derivatives jump across element edges, and cubature / evaluation points sit exactly on
    # edges and vertices: DESIGN section 11), so a 1-ulp difference in x_phys between two deformer
    # implementations can flip tie decisions and move it by O(1e-2).  Each link is therefore checked on
    # IDENTICAL inputs: (1) the FEM cotangent the kernels hand to the deformer == autograd through the oracle FEM
    # evaluated on the GPU's own x_phys, bit for bit the same coordinates;
    gx = x_phys.detach().cpu().clone().requires_grad_(True)
    sols2 = []
    for b in range(B):
        cen = torch.from_numpy(np.stack(data.pde_params["centers"][b])).float()
        scl = torch.from_numpy(np.stack(data.pde_params["scales"][b])).float()
        _, s = Fz.fem2d_fast(topo.cells, bc, gx[b * n * n:(b + 1) * n * n], [X, Y], K, cen, scl)
        sols2.append(s.reshape(-1)[order])
    F.mse_loss(torch.cat(sols2), data.u_true_fine_tensor).backward()
    cot = x_phys.grad.detach().cpu()
    assert (cot - gx.grad).abs().max().item() <= 5e-5 * gx.grad.abs().max().item()
    # (2) the deformer's backward of THAT cotangent == autograd through the oracle deformer.  The parameter
    # gradients are ~1e-3 of the cotangent (cancellation over the nodes), which amplifies rounding accordingly.
    ref.zero_grad(set_to_none=True)
    ref(data).backward(cot)
    scale = max(p.grad.abs().max().item() for n_, p in ref.named_parameters() if p.grad is not None and "lin_key.bias" not in n_)
    for (n_, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        if q.grad is None or "lin_key.bias" in n_:
            continue
        err = (p.grad.cpu() - q.grad).abs().max().item() / scale
        assert err <= 2e-3, (n_, err)
    assert not torch.equal(order, torch.arange(Q * Q))         # the fine-mesh order is a real permutation here
