"""Global CNN feature extractor of the deformer input (scope row f3), on the sm_100a kernels of csrc/glob_cnn.cu.

Mirror of `GlobalFeatureExtractorCNN` (`src/feature_extractors.py:6-34`): same constructor, same sub-modules
(`convs`: `nn.Conv1d` / `nn.Conv2d`, kernel 3, stride 1, padding 1 -- so `state_dict` keys, shapes and the default
initialisation are the reference's), same result `[B, out_channels]`.  What differs is underneath: normalisation,
all conv + SELU layers and the average pool are ONE launch (one CTA per mesh, activation planes in shared memory),
the backward is one launch plus a fixed-order batch reduction, and the canonical-grid reordering the reference does
with `reshape_fd_tensor_to_grid` (`src/utils_data.py:125-141`) can be fused into the kernel's load through `gather`.

There is no CPU path: the module raises when its input is not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _CnnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, gather, scale, H, W, Cm, Co, *params):
        lib = _lib.load()
        L = len(params) // 2
        ws = [p.detach().float().contiguous() for p in params[0::2]]
        bs = [p.detach().float().contiguous() for p in params[1::2]]
        u = u.detach().float().contiguous()
        B = u.numel() // (H * W)
        out = torch.empty((B, Co), dtype=torch.float32, device=u.device)
        wsb = lib.gad_cnn_workspace_bytes(B, H, W, Cm, Co, L)
        work = torch.empty(wsb, dtype=torch.uint8, device=u.device)
        with torch.cuda.device(u.device):
            _lib.check(lib.gad_cnn_fwd(_lib.ptr(u), _lib.ptr(gather), _lib.ptr(scale), B, H, W, Cm, Co, L, _ptr_array(ws),
                                       _ptr_array(bs), _lib.ptr(out), _lib.ptr(work), wsb,
                                       torch.cuda.current_stream(u.device).cuda_stream), "gad_cnn_fwd")
        ctx.save_for_backward(u, scale, *ws, *bs)
        ctx.gather, ctx.shape = gather, (B, H, W, Cm, Co, L)
        ctx.param_shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        B, H, W, Cm, Co, L = ctx.shape
        saved = ctx.saved_tensors
        u, scale, ws, bs = saved[0], saved[1], saved[2:2 + L], saved[2 + L:2 + 2 * L]
        n = int(lib.gad_cnn_param_count(H, Cm, Co, L))
        g_flat = torch.empty(n, dtype=torch.float32, device=u.device)
        wsb = lib.gad_cnn_workspace_bytes(B, H, W, Cm, Co, L)
        work = torch.empty(wsb, dtype=torch.uint8, device=u.device)
        g_out = g_out.float().contiguous()
        with torch.cuda.device(u.device):
            _lib.check(lib.gad_cnn_bwd(_lib.ptr(u), _lib.ptr(ctx.gather), _lib.ptr(scale), B, H, W, Cm, Co, L, _ptr_array(ws),
                                       _ptr_array(bs), _lib.ptr(g_out), _lib.ptr(g_flat), _lib.ptr(work), wsb,
                                       torch.cuda.current_stream(u.device).cuda_stream), "gad_cnn_bwd")
        grads, o = [], 0
        for l in range(L):                      # flat order: w_0, b_0, w_1, b_1, ...
            for shp in (ctx.param_shapes[2 * l], ctx.param_shapes[2 * l + 1]):
                k = 1
                for d in shp:
                    k *= int(d)
                grads.append(g_flat[o:o + k].view(shp))
                o += k
        return (None, None, None, None, None, None, None, *grads)


class GlobalFeatureExtractorCNN(nn.Module):
    def __init__(self, in_channels, mid_channels, out_channels, dim=2, num_layers=4):
        super().__init__()
        if in_channels != 1:
            raise NotImplementedError("GlobalFeatureExtractorCNN: the deformer feeds one scalar field per call "
                                      "(in_channels = 1, src/GNN.py:174,177)")
        if dim not in (1, 2):
            raise ValueError("dim must be 1 or 2")
        conv = nn.Conv1d if dim == 1 else nn.Conv2d
        self.dim, self.mid_channels, self.out_channels = dim, int(mid_channels), int(out_channels)
        self.convs = nn.ModuleList([conv(in_channels, mid_channels, kernel_size=3, stride=1, padding=1)])
        for _ in range(num_layers - 2):
            self.convs.append(conv(mid_channels, mid_channels, kernel_size=3, stride=1, padding=1))
        self.convs.append(conv(mid_channels, out_channels, kernel_size=3, stride=1, padding=1))

    def forward(self, u: torch.Tensor, gather: Optional[torch.Tensor] = None) -> torch.Tensor:
        """u: `[B, 1, H, W]` / `[B, 1, W]` as the reference passes it -- or, with `gather` (int32 `[H*W]`, the node
        whose value sits in grid cell y*W + x), the raw nodal values `[B, H*W]` of square n x n meshes."""
        if u.device.type != "cuda":
            raise RuntimeError("GlobalFeatureExtractorCNN runs on the CUDA kernels only (there is no CPU fallback)")
        if gather is not None:
            B = u.shape[0]
            hw = u.numel() // B
            if self.dim == 2:
                n = int(round(hw ** 0.5))
                H, W = n, n
            else:
                H, W = 1, hw
            flat = u.reshape(B, hw)
            gather = gather.to(device=u.device, dtype=torch.int32).contiguous()
        else:
            if u.dim() != self.dim + 2 or u.shape[1] != 1:
                raise ValueError(f"expected [B, 1, {'H, W' if self.dim == 2 else 'W'}], got {tuple(u.shape)}")
            H, W = (1, u.shape[-1]) if self.dim == 1 else (u.shape[-2], u.shape[-1])
            flat = u.reshape(u.shape[0], -1)
        scale = torch.max(torch.abs(flat.detach())).reshape(1).float()        # u / torch.max(torch.abs(u)), :26
        params = []
        for c in self.convs:
            params += [c.weight, c.bias]
        return _CnnFunction.apply(flat, gather, scale, int(H), int(W), self.mid_channels, self.out_channels, *params)
