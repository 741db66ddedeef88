"""Mint tests/golden/fem1d_*.pt by running the REFERENCE's own 1-D differentiable FEM
(/root/reference/firedrake_difFEM/difFEM_1d.py, executed in place) on synthetic meshes.

    python -m oracle.ref_harness.make_golden_fem1d        # build container only

The file imports matplotlib and `src.utils_main` (plot helpers) at module top; both are absent /
unneeded here and are replaced by empty stand-ins.  Each fixture stores the mesh points, the
Gaussian centres / scales, the quadrature sizes and what `torch_FEM_1D` + `F.mse_loss` +
autograd produced: coeffs, sol, loss, d loss / d mesh_points (fp32, as the reference runs)."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_FILE = "/root/reference/firedrake_difFEM/difFEM_1d.py"
GOLDEN_DIR = os.path.join(_REPO, "tests", "golden_fem1d")


def load_reference_fem1d():
    for name in ("matplotlib", "matplotlib.pyplot", "src", "src.utils_main"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["src"].utils_main = sys.modules["src.utils_main"]
    sys.modules["src.utils_main"].plot_training_evol = lambda *a, **k: None
    sys.modules["src.utils_main"].plot_mesh_evol = lambda *a, **k: None
    spec = importlib.util.spec_from_file_location("ref_difFEM_1d", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert os.path.realpath(mod.__file__).startswith("/root/reference/")
    return mod


CASES = [  # name, n, gaussians, load_quad_points, eval_quad_points, jitter, seed
    ("uniform_n21_g1", 21, 1, 101, 101, 0.0, 0),
    ("jitter_n21_g2", 21, 2, 101, 101, 0.35, 1),
    ("jitter_n51_g2", 51, 2, 101, 101, 0.35, 2),
    ("burgers_n200_g1", 200, 1, 101, 101, 0.3, 3),
    ("small_n5_g1", 5, 1, 11, 33, 0.2, 4),
]


def main():
    ref = load_reference_fem1d()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, n, G, K, Q, jit, seed in CASES:
        rng = np.random.default_rng(seed)
        x = np.linspace(0.0, 1.0, n)
        x[1:-1] += (rng.random(n - 2) - 0.5) * jit / (n - 1)
        centers = rng.uniform(0.25, 0.75, size=G).astype(np.float32)
        scales = rng.uniform(0.1, 0.4, size=G).astype(np.float32)
        mesh = torch.tensor(x, dtype=torch.float32, requires_grad=True)
        quad = torch.linspace(0, 1, Q)
        c_list = [torch.tensor(c) for c in centers]
        s_list = [torch.tensor(s) for s in scales]
        opt = {"load_quad_points": K, "stiff_quad_points": 3}
        coeffs, mesh_out, sol, BC1, BC2 = ref.torch_FEM_1D(opt, mesh, quad, n, c_list, s_list)
        loss = F.mse_loss(sol, ref.u_true_exact_1d(quad, c_list, s_list))
        loss.backward()
        torch.save({"name": name, "mesh": mesh.detach().clone(), "centers": torch.from_numpy(centers),
                    "scales": torch.from_numpy(scales), "load_quad_points": K, "eval_quad_points": Q,
                    "coeffs": coeffs.detach().reshape(-1), "sol": sol.detach(), "loss": float(loss.item()),
                    "grad_mesh": mesh.grad.detach().clone()}, os.path.join(GOLDEN_DIR, f"fem1d_{name}.pt"))
        print(f"{name}: n={n} loss={loss.item():.3e} |grad|max={mesh.grad.abs().max().item():.3e}")


if __name__ == "__main__":
    main()
