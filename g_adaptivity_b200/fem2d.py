"""Batched differentiable 2-D FEM solve (`loss_type='pde_loss'` on 2-D meshes; reference:
firedrake_difFEM/difFEM_2d.py:345-372 looped per mesh by src/GNN.py:327-335) -- plumbing over
csrc/fem2d.cu: one CTA per mesh, topology shared by the batch.

Parity: green on the B200 against the fixtures minted from the reference (tests/test_fem2d_gpu.py).
`GNN.forward` routes `loss_type='pde_loss'` on 2-D meshes here (`GNN._pde_tail_2d`, mirroring src/GNN.py:327-335
including the `mapping_tensor_fine` reordering).  There is no CPU fallback."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib


class Fem2DTopology:
    """Device tables of one triangulation: cells, Dirichlet mask and the star table (cells around every node in
    ascending cell order, the order `torch.where(cell_node_map == n)` yields them, difFEM_2d.py:31)."""

    def __init__(self, cells, bc_nodes: Sequence[int], num_nodes: int, device):
        cells = np.asarray(cells, dtype=np.int64)
        T = cells.shape[0]
        deg = np.bincount(cells.reshape(-1), minlength=num_nodes)
        D = int(deg.max())
        star_cell = np.full((num_nodes, D), -1, dtype=np.int32)
        star_loc = np.zeros((num_nodes, D), dtype=np.int32)
        fill = np.zeros(num_nodes, dtype=np.int64)
        for t in range(T):
            for k in range(3):
                m = cells[t, k]
                star_cell[m, fill[m]], star_loc[m, fill[m]] = t, k
                fill[m] += 1
        is_bc = np.zeros(num_nodes, dtype=np.uint8)
        is_bc[np.asarray(bc_nodes, dtype=np.int64)] = 1
        dev = torch.device(device)
        self.T, self.N, self.D = T, int(num_nodes), D
        self.cells = torch.from_numpy(cells.astype(np.int32)).to(dev)
        self.is_bc = torch.from_numpy(is_bc).to(dev)
        self.star_cell = torch.from_numpy(star_cell).to(dev)
        self.star_loc = torch.from_numpy(star_loc).to(dev)
        self.device = dev


def pde_params_to_tensors(pde_params, B: int, device) -> tuple:
    """`data.pde_params['centers'][b]` is a list (one entry per Gaussian) of float32 arrays [2]
    (src/data.py:155-156; handed to torch_FEM_2D as tensors, src/GNN.py:319-320) -> fp64 [B, G, 2]."""
    def stack(key):
        try:
            arr = np.asarray(pde_params[key][:B], dtype=np.float32).reshape(B, -1, 2)
        except ValueError as e:
            raise ValueError("every mesh of a batch must carry the same number of 2-D Gaussians") from e
        return torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=True).double()
    return stack("centers"), stack("scales")


class FEM2DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, topo: Fem2DTopology, centers, scales, eval_x, eval_y, load_quad_points: int):
        if coords.device.type != "cuda":
            raise RuntimeError("fem2d needs CUDA tensors (there is no CPU fallback)")
        lib = _lib.load()
        B, N, _ = coords.shape
        coords = coords.detach().float().contiguous()
        centers = centers.to(coords.device, torch.float64).contiguous()
        scales = scales.to(coords.device, torch.float64).contiguous()
        eval_x, eval_y = eval_x.float().contiguous(), eval_y.float().contiguous()
        G, Q = centers.shape[1], eval_x.numel()
        coeffs = torch.empty((B, N), dtype=torch.float32, device=coords.device)
        sol = torch.empty((B, Q), dtype=torch.float32, device=coords.device)
        u64 = torch.empty((B, N), dtype=torch.float64, device=coords.device)
        iters = torch.zeros(B, dtype=torch.int32, device=coords.device)
        stream = torch.cuda.current_stream(coords.device).cuda_stream
        with torch.cuda.device(coords.device):
            _lib.check(lib.gad_fem2d_fwd(_lib.ptr(topo.cells), topo.T, _lib.ptr(topo.is_bc), N, _lib.ptr(topo.star_cell),
                                         _lib.ptr(topo.star_loc), topo.D, _lib.ptr(coords), _lib.ptr(centers), _lib.ptr(scales),
                                         G, B, int(load_quad_points), _lib.ptr(eval_x), _lib.ptr(eval_y), Q, _lib.ptr(coeffs),
                                         _lib.ptr(sol), _lib.ptr(u64), _lib.ptr(iters), stream), "gad_fem2d_fwd")
        ctx.save_for_backward(coords, centers, scales, eval_x, eval_y, u64)
        ctx.topo, ctx.K = topo, int(load_quad_points)
        ctx.mark_non_differentiable(coeffs, iters)
        return sol, coeffs, iters

    @staticmethod
    def backward(ctx, g_sol, _g_coeffs, _g_iters):
        coords, centers, scales, eval_x, eval_y, u64 = ctx.saved_tensors
        topo, lib = ctx.topo, _lib.load()
        B, N, _ = coords.shape
        g_sol = g_sol.float().contiguous()
        grad = torch.empty((B, N, 2), dtype=torch.float32, device=coords.device)
        stream = torch.cuda.current_stream(coords.device).cuda_stream
        with torch.cuda.device(coords.device):
            _lib.check(lib.gad_fem2d_bwd(_lib.ptr(topo.cells), topo.T, _lib.ptr(topo.is_bc), N, _lib.ptr(topo.star_cell),
                                         _lib.ptr(topo.star_loc), topo.D, _lib.ptr(coords), _lib.ptr(centers), _lib.ptr(scales),
                                         centers.shape[1], B, ctx.K, _lib.ptr(eval_x), _lib.ptr(eval_y), eval_x.numel(),
                                         _lib.ptr(u64), _lib.ptr(g_sol), _lib.ptr(grad), stream), "gad_fem2d_bwd")
        return grad, None, None, None, None, None, None


def fem2d_solve(coords, topo: Fem2DTopology, centers, scales, eval_x, eval_y, load_quad_points: int):
    """coords [B, N, 2] (differentiable) -> (sol [B, Q], coeffs [B, N], CG iterations [B])."""
    return FEM2DFunction.apply(coords, topo, centers, scales, eval_x, eval_y, load_quad_points)
