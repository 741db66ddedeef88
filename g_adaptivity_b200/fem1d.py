"""Batched differentiable 1-D FEM solve on the GPU: the `pde_loss` tail of `GNN.forward` for 1-D
meshes (`src/GNN.py:307-342` calling `torch_FEM_1D`, firedrake_difFEM/difFEM_1d.py:211-238, once per
mesh in a Python loop).  One launch for the whole batch forward, one for the backward
(csrc/fem1d.cu); no CPU fallback."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib


def _pde_params_to_tensors(pde_params, B: int, device) -> tuple:
    """`data.pde_params['centers'][b]` is a list (one entry per Gaussian) of 1-element arrays."""
    def stack(key):
        try:                                                   # rectangular: one vectorised conversion
            arr = np.asarray(pde_params[key][:B], dtype=np.float32)
            arr = arr.reshape(B, -1)
        except ValueError as e:
            raise ValueError("every mesh of a batch must carry the same number of Gaussians") from e
        return torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=True)
    return stack("centers"), stack("scales")


class FEM1DFunction(torch.autograd.Function):
    """sol [B*Q], coeffs [B*(n-2)] = FEM(x_phys [B*n]); differentiable with respect to x_phys
    through `sol` (as in the reference, no gradient flows through the Dirichlet values)."""

    @staticmethod
    def forward(ctx, x, centers, scales, quad, n, K):
        if x.device.type != "cuda":
            raise RuntimeError("fem1d needs CUDA tensors (there is no CPU fallback)")
        lib = _lib.load()
        xf = x.detach().reshape(-1).float().contiguous()
        B = xf.numel() // n
        assert B * n == xf.numel() and centers.shape[0] == B
        G, Q = int(centers.shape[1]), int(quad.numel())
        sol = torch.empty(B * Q, dtype=torch.float32, device=x.device)
        coeffs = torch.empty(B * (n - 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.gad_fem1d_fwd(_lib.ptr(xf), _lib.ptr(centers), _lib.ptr(scales), _lib.ptr(quad), B, n, G, K, Q,
                                         _lib.ptr(sol), _lib.ptr(coeffs), torch.cuda.current_stream(x.device).cuda_stream),
                       "gad_fem1d_fwd")
        ctx.save_for_backward(xf, centers, scales, quad)
        ctx.meta = (B, n, G, K, Q, x.shape)
        ctx.mark_non_differentiable(coeffs)
        return sol, coeffs

    @staticmethod
    def backward(ctx, g_sol, _g_coeffs):
        xf, centers, scales, quad = ctx.saved_tensors
        B, n, G, K, Q, shape = ctx.meta
        lib = _lib.load()
        g = g_sol.reshape(-1).float().contiguous()
        g_x = torch.empty(B * n, dtype=torch.float32, device=xf.device)
        with torch.cuda.device(xf.device):
            _lib.check(lib.gad_fem1d_bwd(_lib.ptr(xf), _lib.ptr(centers), _lib.ptr(scales), _lib.ptr(quad), _lib.ptr(g), B, n,
                                         G, K, Q, _lib.ptr(g_x), torch.cuda.current_stream(xf.device).cuda_stream),
                       "gad_fem1d_bwd")
        return g_x.view(shape), None, None, None, None, None


def fem1d_solve(x_phys: torch.Tensor, centers: torch.Tensor, scales: torch.Tensor, quad_points: torch.Tensor,
                num_meshpoints: int, load_quad_points: int = 101):
    """Batched `torch_FEM_1D`: x_phys [B*n] (or [B*n, 1]) -> (coeffs [B*(n-2), 1], sol [B*Q])."""
    dev = x_phys.device
    quad = quad_points.to(device=dev, dtype=torch.float32).contiguous()
    sol, coeffs = FEM1DFunction.apply(x_phys, centers.to(dev).float().contiguous(), scales.to(dev).float().contiguous(), quad,
                                      int(num_meshpoints), int(load_quad_points))
    return coeffs.unsqueeze(-1), sol
