"""Mint tests/golden_pde2d/*.pt by running the REFERENCE's own `GNN.forward` with `loss_type='pde_loss'` on 2-D
batches: deformer (src/GNN.py:190-306) -> per-mesh `torch_FEM_2D` (firedrake_difFEM/difFEM_2d.py:345-372) ->
`reshape_grid_to_fd_tensor` with `dataset.mapping_tensor_fine` (src/utils_data.py:143-159; GNN.py:307-342) ->
`F.mse_loss(sol, data.u_true_fine_tensor)` (src/run_GNN.py:109-110) -> autograd.

    python -m oracle.ref_harness.make_golden_pde2d        # build container only

Every file involved is the reference's, executed in place: GNN.py / GRAND_plus.py / utils_data.py through
load_reference.py (PyG shim), difFEM_2d.py through make_golden_fem2d.load_reference_fem2d (Firedrake reduced to the
cell-node map and the Dirichlet nodes; torchquad's Simpson rule restated -- that part stays unpinned).  The fixture
stores the batch, the state_dict and what the reference returned: coeffs, x_phys, sol (fine-mesh order), the loss and
the parameter gradients."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _REPO)
GOLDEN_DIR = os.path.join(_REPO, "tests", "golden_pde2d")

# name, n, meshes, eval points per side, load_quad_points, seed
CASES = [("pde2d_6x6_b2", 6, 2, 9, 121, 4), ("pde2d_7x7_b3", 7, 3, 13, 225, 5)]


def main():
    from g_adaptivity_b200 import synth
    from oracle.ref_harness import load_reference, make_golden_fem2d
    fem = make_golden_fem2d.load_reference_fem2d()
    gnn_mod, _, _ = load_reference.load()
    gnn_mod.torch_FEM_2D = fem.torch_FEM_2D          # the name GNN.py imported from firedrake_difFEM.difFEM_2d (:9)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, n, B, Q, K, seed in CASES:
        md = (n, n)
        opt = synth.default_opt(md, loss_type="pde_loss", eval_quad_points=Q, load_quad_points=K)
        ds = synth.SyntheticDataset(2, md, eval_quad_points=Q)
        data = synth.make_batch(md, B, seed=seed, eval_quad_points=Q, with_u_true_fine=True)
        torch.manual_seed(opt["seed"])
        with load_reference.quiet():
            model = gnn_mod.GNN(ds, opt)
        model.train()
        coeffs, x_phys, sol = model(data)
        loss = F.mse_loss(sol, data.u_true_fine_tensor)
        loss.backward()
        inputs = {k: getattr(data, k).clone() for k in
                  ("edge_index", "batch", "x_comp", "x_phys", "f_tensor", "uu_tensor", "u_true_tensor", "u_true_fine_tensor",
                   "to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask")}
        fx = {"name": name, "mesh_dims_list": [list(md)] * B, "eval_quad_points": Q, "load_quad_points": K, "seed": seed,
              "opt_overrides": {"loss_type": "pde_loss", "eval_quad_points": Q, "load_quad_points": K},
              "inputs": inputs, "corner_nodes": [torch.from_numpy(np.asarray(c, dtype=np.int64)) for c in data.corner_nodes],
              "centers": torch.from_numpy(np.asarray(data.pde_params["centers"], dtype=np.float32)),
              "scales": torch.from_numpy(np.asarray(data.pde_params["scales"], dtype=np.float32)),
              "mapping_tensor_fine": ds.mapping_tensor_fine.clone(),
              "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
              "coeffs": coeffs.detach().clone(), "x_phys": x_phys.detach().clone(), "sol": sol.detach().clone(),
              "loss": float(loss.item()),
              "grads": {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}}
        torch.save(fx, os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: coeffs {tuple(coeffs.shape)} x_phys {tuple(x_phys.shape)} sol {tuple(sol.shape)} loss {loss.item():.4e}")


if __name__ == "__main__":
    main()
