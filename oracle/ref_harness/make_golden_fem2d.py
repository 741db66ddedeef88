"""Mint tests/golden_fem2d/*.pt by running the REFERENCE's own 2-D differentiable FEM
(/root/reference/firedrake_difFEM/difFEM_2d.py, executed in place) on small synthetic meshes.

    python -m oracle.ref_harness.make_golden_fem2d        # build container only

The file imports firedrake, torchquad, torchdiffeq, matplotlib and `firedrake_difFEM.solve_poisson` at
module top; none is installed here.  Stand-ins:
  * firedrake: `FunctionSpace` / `DirichletBC(...).nodes` / `mesh.coordinates.cell_node_map().values` answer
    from a plain (cells, boundary nodes) pair -- the only three things `torch_FEM_2D` asks of Firedrake;
  * torchquad: `Simpson().integrate` is oracle.fem2d_oracle.simpson_2d, the restatement of torchquad's
    composite Simpson rule (the package is an un-vendored, un-pinned dependency: see the oracle's header --
    this part of the parity is UNPINNED, the fixtures pin everything else of the reference file);
  * the rest: empty modules.
Each fixture stores cells, boundary nodes, mesh points, Gaussian centres / scales, quadrature sizes and
what `torch_FEM_2D` + `F.mse_loss(sol, u_true)` + autograd produced (fp32, as the reference runs)."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _REPO)
REF_FILE = "/root/reference/firedrake_difFEM/difFEM_2d.py"
GOLDEN_DIR = os.path.join(_REPO, "tests", "golden_fem2d")


class _Mesh:
    """What torch_FEM_2D reads from a Firedrake mesh."""

    def __init__(self, cells: np.ndarray, bc_nodes: np.ndarray):
        self._cells, self.bc_nodes = cells, bc_nodes
        self.coordinates = self

    def cell_node_map(self):
        return types.SimpleNamespace(values=self._cells)


def load_reference_fem2d():
    from oracle import fem2d_oracle
    names = ("matplotlib", "matplotlib.pyplot", "firedrake", "torchquad", "torchdiffeq", "firedrake_difFEM",
             "firedrake_difFEM.solve_poisson")
    for name in names:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    fd = sys.modules["firedrake"]
    for n in ("TestFunction", "TrialFunction", "Function", "SpatialCoordinate", "UnitSquareMesh", "inner", "grad", "dx",
              "div", "exp", "triplot", "tripcolor", "solve", "sqrt", "assemble", "tricontour"):
        setattr(fd, n, None)
    fd.FunctionSpace = lambda mesh, family, degree: mesh
    fd.DirichletBC = lambda V, value, where: types.SimpleNamespace(nodes=V.bc_nodes)
    tq = sys.modules["torchquad"]

    class Simpson:
        def integrate(self, fn, dim, N, integration_domain, backend):
            assert dim == 2 and backend == "torch"
            return fem2d_oracle.simpson_2d(fn, N, integration_domain)

    tq.Simpson = Simpson
    tq.Trapezoid = tq.Gaussian = tq.set_up_backend = tq.utils = None
    sys.modules["torchdiffeq"].odeint = sys.modules["torchdiffeq"].odeint_adjoint = None
    sys.modules["firedrake_difFEM"].solve_poisson = sys.modules["firedrake_difFEM.solve_poisson"]
    sys.modules["firedrake_difFEM.solve_poisson"].poisson2d_fmultigauss_bcs = None
    spec = importlib.util.spec_from_file_location("ref_difFEM_2d", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert os.path.realpath(mod.__file__).startswith("/root/reference/")
    return mod


CASES = [  # name, n (nodes per side), gaussians, load_quad_points, eval points per side, jitter, seed
    ("uniform_n4_g1", 4, 1, 121, 9, 0.0, 0),
    ("jitter_n5_g1", 5, 1, 225, 11, 0.3, 1),
    ("jitter_n6_g2", 6, 2, 441, 13, 0.3, 2),
    ("even_quad_n5_g2", 5, 2, 100, 8, 0.25, 3),       # N per dimension 10 -> reduced to 9
]


def case_inputs(n, G, jitter, seed):
    from g_adaptivity_b200 import synth
    topo = synth.MeshTopology((n, n))
    rng = np.random.default_rng(seed)
    pts = topo.coords.astype(np.float64).copy()
    interior = ~topo.boundary_nodes
    pts[interior] += (rng.random((int(interior.sum()), 2)) - 0.5) * jitter / (n - 1)
    centers = rng.uniform(0.3, 0.7, size=(G, 2))
    scales = rng.uniform(0.2, 0.4, size=(G, 2))
    return topo.cells.astype(np.int64), np.nonzero(topo.boundary_nodes)[0].astype(np.int64), pts.astype(np.float32), centers, scales


def main():
    ref = load_reference_fem2d()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, n, G, K, Q, jit, seed in CASES:
        cells, bc_nodes, pts, centers, scales = case_inputs(n, G, jit, seed)
        mesh_points = torch.tensor(pts, requires_grad=True)
        x0 = torch.linspace(0, 1, Q)
        X, Y = torch.meshgrid(x0, x0, indexing="ij")          # src/GNN.py:185-188
        # src/GNN.py:311-312: centres / scales arrive as float64 numpy arrays
        c_list = [torch.from_numpy(c.copy()) for c in centers]
        s_list = [torch.from_numpy(s.copy()) for s in scales]
        opt = {"load_quad_points": K, "device": "cpu"}
        coeffs, _, sol = ref.torch_FEM_2D(opt, _Mesh(cells, bc_nodes), mesh_points, [X, Y], n, c_list, s_list)
        target = ref.u_true_exact_2d(torch.stack([X, Y]), c_list, s_list)
        loss = F.mse_loss(sol, target)
        loss.backward()
        torch.save({"name": name, "n": n, "cells": torch.from_numpy(cells), "bc_nodes": torch.from_numpy(bc_nodes),
                    "mesh": mesh_points.detach().clone(), "centers": torch.from_numpy(centers),
                    "scales": torch.from_numpy(scales), "load_quad_points": K, "eval_points": Q,
                    "coeffs": coeffs.detach().clone(), "sol": sol.detach().clone(), "loss": float(loss.item()),
                    "grad_mesh": mesh_points.grad.detach().clone()}, os.path.join(GOLDEN_DIR, f"fem2d_{name}.pt"))
        print(f"{name}: n={n} loss={loss.item():.3e} |grad|max={mesh_points.grad.abs().max().item():.3e} "
              f"max|coeffs-u_true(nodes)|={(coeffs.detach().reshape(-1) - ref.u_true_exact_2d(mesh_points.detach().t().contiguous(), c_list, s_list)).abs().max().item():.3e}")


if __name__ == "__main__":
    main()
