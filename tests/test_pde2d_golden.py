"""The whole 2-D `loss_type='pde_loss'` route -- the reference's DEFAULT loss (params.py:109) -- against fixtures minted
from the reference's own `GNN.forward` (src/GNN.py:190-342: deformer, per-mesh `torch_FEM_2D`,
`reshape_grid_to_fd_tensor` with `mapping_tensor_fine`) + `F.mse_loss` + autograd, by
oracle/ref_harness/make_golden_pde2d.py.

CPU: the oracle composition (gnn_oracle.GNNRef with its pde_loss tail) reproduces the fixtures.
GPU: `g_adaptivity_b200.GNN(loss_type='pde_loss')` -- deformer kernel + batched FEM kernels -- against the same fixtures."""
import copy
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden_pde2d", "*.pt")))


def _rebuild(fx):
    from g_adaptivity_b200 import synth
    md = tuple(fx["mesh_dims_list"][0])
    B, Q, K = len(fx["mesh_dims_list"]), fx["eval_quad_points"], fx["load_quad_points"]
    opt = synth.default_opt(md, **fx["opt_overrides"])
    ds = synth.SyntheticDataset(2, md, eval_quad_points=Q)
    data = synth.make_batch(md, B, seed=fx["seed"], eval_quad_points=Q, with_u_true_fine=True)
    # the fixture's own inputs win over whatever synth generates today
    for k, v in fx["inputs"].items():
        setattr(data, k, v.clone())
    data.pde_params = {"centers": [[c.numpy() for c in row] for row in fx["centers"]],
                       "scales": [[s.numpy() for s in row] for row in fx["scales"]]}
    assert torch.equal(ds.mapping_tensor_fine, fx["mapping_tensor_fine"])
    return opt, ds, data, md, B, Q


def test_fixtures_present():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
@pytest.mark.parametrize("fem", ["reference", "fast"])
def test_oracle_reproduces_reference_pde_loss_2d(path, fem):
    from oracle import gnn_oracle
    fx = torch.load(path)
    opt, ds, data, md, B, Q = _rebuild(fx)
    opt["oracle_fem2d"] = fem
    ref = gnn_oracle.GNNRef(ds, copy.deepcopy(opt))
    ref.load_state_dict(fx["state_dict"])
    ref.train()
    coeffs, x_phys, sol = ref(data)
    loss = F.mse_loss(sol, data.u_true_fine_tensor)
    loss.backward()
    sc = fx["coeffs"].abs().max().item()
    tol = 1e-6 if fem == "reference" else 2e-5
    assert (x_phys.detach() - fx["x_phys"]).abs().max().item() <= 1e-6
    assert (coeffs.detach() - fx["coeffs"]).abs().max().item() <= tol * sc
    assert (sol.detach() - fx["sol"]).abs().max().item() <= tol * sc
    assert abs(loss.item() - fx["loss"]) <= 10 * tol * abs(fx["loss"])
    if fem != "reference":
        return          # the vectorised formulation's gradient is checked against the line-by-line one in test_fem2d_oracle
    scale = max(g.abs().max().item() for k, g in fx["grads"].items() if "lin_key.bias" not in k)
    for k, p in ref.named_parameters():
        if k not in fx["grads"] or "lin_key.bias" in k:
            continue
        assert (p.grad - fx["grads"][k]).abs().max().item() <= 1e-4 * scale, k


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_cuda_pde_loss_2d_matches_reference_fixture(path):
    from g_adaptivity_b200 import GNN
    fx = torch.load(path)
    opt, ds, data, md, B, Q = _rebuild(fx)
    n = md[0]
    opt.update(device="cuda")
    model = GNN(ds, opt).to("cuda")
    model.load_state_dict(fx["state_dict"])
    model.train()
    coeffs, x_phys, sol = model(data)
    assert coeffs.shape == fx["coeffs"].shape and x_phys.shape == fx["x_phys"].shape and sol.shape == fx["sol"].shape
    loss = F.mse_loss(sol, data.u_true_fine_tensor.cuda())
    loss.backward()
    sc = fx["coeffs"].abs().max().item()
    assert (x_phys.detach().cpu() - fx["x_phys"]).abs().max().item() <= 1e-5 * fx["x_phys"].abs().max().item()
    assert (coeffs.detach().cpu() - fx["coeffs"]).abs().max().item() <= 2e-5 * sc
    assert (sol.detach().cpu() - fx["sol"]).abs().max().item() <= 2e-5 * sc
    assert abs(loss.item() - fx["loss"]) <= 1e-4 * abs(fx["loss"])
    # Parameter gradients.  The reference's FEM gradient is a discontinuous function of the mesh points (DESIGN
    # section 11: cubature and evaluation points sit exactly on element edges, and a 1-ulp difference in x_phys can
    # flip a tie), so against a fixture computed on the REFERENCE's x_phys the bar is loose; the tight, link-by-link
    # check on identical inputs is test_fem2d_gpu.py::test_pde_loss_through_the_deformer_2d.
    scale = max(g.abs().max().item() for k, g in fx["grads"].items() if "lin_key.bias" not in k)
    worst = 0.0
    for k, p in model.named_parameters():
        if k not in fx["grads"] or "lin_key.bias" in k:
            continue
        worst = max(worst, (p.grad.cpu() - fx["grads"][k]).abs().max().item() / scale)
    assert worst <= 5e-2, worst
