"""Device-resident mesh graph: the cached result of the K0 graph builder.

Host-side mirror of the graph prologue of `GNN.forward` (`src/GNN.py:206-223`).  The reference
re-does the mask filtering, the corner-loop concatenation and (through PyG) the gather/scatter
index handling on every call although the topology never changes between calls
(`src/utils_eval_Burgers.py:269,297` re-invoke the model on the same `data`); here it is done
once by `gad_graph_build` and cached.  The host only plans *tiles*: contiguous node ranges that
do not split a mesh, which the mesh-resident kernels process one CTA each.
"""
from __future__ import annotations

import collections
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

_FUSED_MAX_TILE_NODES = 3072   # csrc/fused_kernels.cu: MAX_FUSED_TILE_NODES
_DEFAULT_TILE_TARGET = 1024    # pack small meshes up to this many nodes per CTA


def _bwd_smem_bytes(nodes: int, edges: int, ce: int) -> int:
    """Shared memory of the backward mesh-resident kernel (csrc/fused_kernels.cu: bwd_layout)."""
    npt = 1 if nodes <= 256 else (2 if nodes <= 1024 else 4)
    threads = max(64, (((nodes + npt - 1) // npt) + 31) // 32 * 32)
    nacc = ce * ce + ce + 1
    a16 = lambda b: (b + 15) // 16 * 16
    return (3 * a16(nodes * ce * 4) + a16(nodes * 8) + 2 * a16((nodes + 1) * 4) + a16((ce * ce + ce) * 4)
            + a16(nacc * ((threads + 31) // 32) * 4) + 2 * a16(edges * 2))


def plan_tiles(mesh_sizes: Sequence[int], target_nodes: int = _DEFAULT_TILE_TARGET,
               max_nodes: int = _FUSED_MAX_TILE_NODES) -> Optional[np.ndarray]:
    """Greedy packing of consecutive meshes into tiles of at most `target_nodes` nodes (a mesh
    larger than the target gets a tile of its own).  Returns node offsets int32 [T+1], or None
    when some mesh exceeds `max_nodes` (the batch then runs on the streaming kernels)."""
    ptr = [0]
    cur = 0
    for n in mesh_sizes:
        n = int(n)
        if n > max_nodes:
            return None
        if cur > 0 and cur + n > target_nodes:
            ptr.append(ptr[-1] + cur)
            cur = 0
        cur += n
    if cur > 0:
        ptr.append(ptr[-1] + cur)
    return np.asarray(ptr, dtype=np.int32)


class MeshGraph:
    """CSR (by destination) / CSC (by source) of the filtered edge list + tile plan."""

    def __init__(self):
        self.N = 0
        self.E = 0
        self.device = None
        self.edge_index = None      # int64 [2, E]  filtered list, reference order (conv.stored_ei)
        self.rowptr = self.col = self.eid = None
        self.t_rowptr = self.t_dst = self.t_slot = None
        self.col_walk = self.t_dst_walk = None   # per-row ascending copies used by the deformer kernels
        self.max_in_deg = self.max_out_deg = 0
        self.tile_ptr = None        # int32 [T+1] on device, or None -> streaming kernels
        self.ell_in = self.ell_out = None   # uint16 [N, 8] ELL rows (mesh-resident ELL kernels), or None
        self.wide_in = self.wide_out = None  # int32 [N, 8] wide rows (streaming ELL kernels), built on demand
        self.wide_reach = -1                 # max |j - i| over the edges, known once the wide rows exist
        self._wide_tried = False
        self.cl_in = self.cl_out = None      # int32 [N, 8] cluster rows (cluster-resident training kernel)
        self.cl_C = self.cl_S = 0            # cluster size, slab size
        self.mesh_ptr = None                 # int32 [M + 1] on device
        self.mesh_sizes = None
        self._cl_tried = False
        self.ell_ce = 0             # channel width the ELL byte offsets were built for
        self.ell_deg = 0            # max(in-degree, out-degree)
        self.T = 0
        self.max_tile_nodes = 0
        self.max_tile_edges = 0
        self._keepalive = ()
        self.uniform = False        # shared-topology batch: equal tiles, ell_in / ell_out are ONE tile's rows
        self.sub = None             # ... and the general graph of that one tile

    # ------------------------------------------------------------------------------------
    @staticmethod
    def build(edge_index: torch.Tensor, num_nodes: int, masks: Sequence[Optional[torch.Tensor]] = (),
              extra_loops: Optional[torch.Tensor] = None, self_loops: bool = False,
              mesh_sizes: Optional[Sequence[int]] = None, device=None, ce: int = 4,
              tile_target: Optional[int] = None, use_ell: bool = True) -> "MeshGraph":
        lib = _lib.load()
        if device is None:
            device = edge_index.device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("MeshGraph.build needs a CUDA device: the deformer has no CPU path")
        ei = edge_index.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
        E0 = int(ei.shape[1])
        N = int(num_nodes)
        ms = [None, None, None]
        for k, m in enumerate(masks):
            if m is not None:
                ms[k] = m.to(device=device, non_blocking=True).to(torch.uint8).contiguous()
                assert ms[k].numel() == E0
        K = 0
        if extra_loops is not None and extra_loops.numel() > 0:
            extra_loops = extra_loops.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()
            K = int(extra_loops.numel())
        else:
            extra_loops = None
        Emax = E0 + K + (N if self_loops else 0)
        g = MeshGraph()
        g.N, g.device = N, device
        i32 = dict(dtype=torch.int32, device=device)
        filt = torch.empty((2, max(Emax, 1)), dtype=torch.int64, device=device)
        g.rowptr = torch.empty(N + 1, **i32)
        g.t_rowptr = torch.empty(N + 1, **i32)
        col = torch.empty(max(Emax, 1), **i32)
        eid = torch.empty(max(Emax, 1), **i32)
        t_dst = torch.empty(max(Emax, 1), **i32)
        t_slot = torch.empty(max(Emax, 1), **i32)
        info = torch.zeros(8, **i32)
        ws_bytes = lib.gad_graph_workspace_bytes(E0, K, N, int(self_loops))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            _lib.check(lib.gad_graph_build(
                _lib.ptr(ei), E0, _lib.ptr(ms[0]), _lib.ptr(ms[1]), _lib.ptr(ms[2]), _lib.ptr(extra_loops), K,
                int(self_loops), N, _lib.ptr(filt), _lib.ptr(g.rowptr), _lib.ptr(col), _lib.ptr(eid),
                _lib.ptr(g.t_rowptr), _lib.ptr(t_dst), _lib.ptr(t_slot), _lib.ptr(info), _lib.ptr(ws), ws_bytes,
                stream), "gad_graph_build")
        info_h = info.cpu()     # the one synchronisation of the build (E is data-dependent)
        if int(info_h[7]) != 0:
            raise ValueError(f"edge_index holds {int(info_h[7])} node ids outside [0, {N})")
        E = int(info_h[0])
        g.E = E
        g.max_in_deg, g.max_out_deg = int(info_h[1]), int(info_h[2])
        # trim views (column stride of `filt` stays Emax, so make the trimmed list contiguous)
        g.edge_index = filt[:, :E].contiguous() if E != filt.shape[1] else filt
        g.col, g.eid, g.t_dst, g.t_slot = col[:E], eid[:E], t_dst[:E], t_slot[:E]
        g._info = info
        # row-sorted copies: what the deformer kernels walk (see gad_graph_sort_rows)
        g.col_walk = torch.empty_like(g.col)
        g.t_dst_walk = torch.empty_like(g.t_dst)
        if E > 0:
            with torch.cuda.device(device):
                _lib.check(lib.gad_graph_sort_rows(_lib.ptr(g.rowptr), _lib.ptr(g.col), N, _lib.ptr(g.col_walk), stream),
                           "gad_graph_sort_rows")
                _lib.check(lib.gad_graph_sort_rows(_lib.ptr(g.t_rowptr), _lib.ptr(g.t_dst), N, _lib.ptr(g.t_dst_walk),
                                                   stream), "gad_graph_sort_rows")
        if mesh_sizes is not None:
            g.plan(mesh_sizes, ce=ce, tile_target=tile_target, use_ell=use_ell)
        return g

    # ------------------------------------------------------------------------------------
    @property
    def ell_tile_ptr(self):
        """What the gad_*_ell entry points take as `tile_ptr`: NULL for a shared-topology graph (equal tiles,
        one ELL table for all of them; include/gadapt.h), the tile offsets otherwise."""
        return None if self.uniform else self.tile_ptr

    @staticmethod
    def build_uniform(data, mesh_sizes: Sequence[int], dim: int, mesh_dims, fix_boundary: bool, self_loops: bool,
                      device, ce: int = 4, tile_target: Optional[int] = None) -> Optional["MeshGraph"]:
        """Shared-topology batch (every sample on the same mesh: the reference's `randg` datasets,
        src/data.py:143; PyG's collation only adds node offsets): run the graph prologue of `GNN.forward`
        (src/GNN.py:206-223) for the meshes of ONE tile and let every tile of the batch use that tile's ELL
        table.  Build cost and topology memory are O(mesh), not O(batch).

        Returns None when the batch is not of that form (unequal sizes, edge list not mesh-major, the last mesh's
        block differing from the first, degree > 7, ...): the caller then builds the general graph."""
        sizes = [int(n) for n in mesh_sizes]
        B = len(sizes)
        if B == 0 or any(n != sizes[0] for n in sizes) or ce not in (2, 4):
            return None
        N1 = sizes[0]
        ei = data.edge_index
        E0 = int(ei.shape[1])
        if E0 == 0 or E0 % B:
            return None
        E1 = E0 // B
        target = tile_target or _DEFAULT_TILE_TARGET
        if N1 > _FUSED_MAX_TILE_NODES:
            return None
        m = min(B, max(1, target // N1))
        names = ("to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask")
        masks_full = [getattr(data, n, None) for n in names] if fix_boundary else [None] * 3
        if fix_boundary and any(t is None for t in masks_full):
            return None
        # the promise, checked where the tensors live: first and last mesh carry the same block
        head = ei[:, :m * E1]
        if int(head.min()) < 0 or int(head.max()) >= m * N1:
            return None
        if B > 1:
            if not torch.equal(ei[:, (B - 1) * E1:] - (B - 1) * N1, ei[:, :E1]):
                return None
            if any(t is not None and not torch.equal(t[(B - 1) * E1:], t[:E1]) for t in masks_full):
                return None
        loops = None
        if fix_boundary:
            class _Head:
                corner_nodes = list(getattr(data, "corner_nodes", []) or [])[:m]
            loops = corner_loops(_Head, dim, mesh_dims, sizes[:m])
        sub = MeshGraph.build(head, m * N1, masks=[None if t is None else t[:m * E1] for t in masks_full],
                              extra_loops=loops, self_loops=self_loops, mesh_sizes=sizes[:m], device=device, ce=ce,
                              tile_target=tile_target, use_ell=True)
        if sub.tile_ptr is None or sub.T != 1 or sub.ell_in is None or sub.E % m:
            return None
        g = MeshGraph()
        g.uniform, g.sub = True, sub
        g.N, g.E, g.device = B * N1, sub.E // m * B, sub.device
        g.mesh_sizes = sizes
        g.T = (B + m - 1) // m
        tp = np.minimum(np.arange(g.T + 1, dtype=np.int64) * (m * N1), B * N1).astype(np.int32)
        g.tile_ptr = torch.from_numpy(tp).to(sub.device)
        g.max_tile_nodes, g.max_tile_edges = m * N1, sub.max_tile_edges
        g.max_in_deg, g.max_out_deg = sub.max_in_deg, sub.max_out_deg
        g.ell_in, g.ell_out, g.ell_ce, g.ell_deg = sub.ell_in, sub.ell_out, sub.ell_ce, sub.ell_deg
        g.sm_count = getattr(sub, "sm_count", 148)
        g._wide_tried = g._cl_tried = True        # no general arrays: the streaming / cluster fallbacks do not apply
        return g

    def plan(self, mesh_sizes: Sequence[int], ce: int = 4, tile_target: Optional[int] = None, use_ell: bool = True):
        """Choose tiles for the mesh-resident kernels; falls back to streaming (tile_ptr = None)
        when a mesh does not fit one CTA's shared memory or the batch is not a disjoint union."""
        lib = _lib.load()
        self.tile_ptr, self.T, self.max_tile_nodes, self.max_tile_edges = None, 0, 0, 0
        self.ell_in = self.ell_out = None
        if int(np.sum(mesh_sizes)) != self.N:
            raise ValueError("mesh_sizes do not add up to the node count")
        self.mesh_sizes = [int(n) for n in mesh_sizes]
        tp = plan_tiles(mesh_sizes, target_nodes=tile_target or _DEFAULT_TILE_TARGET)
        if tp is None:
            return
        tile_ptr = torch.from_numpy(tp).to(self.device)
        edges_at = self.rowptr[tile_ptr.long()].cpu().numpy().astype(np.int64)
        max_nodes = int(np.max(np.diff(tp)))
        max_edges = int(np.max(np.diff(edges_at))) if len(edges_at) > 1 else 0
        sm = C_int()
        smem = C_int()
        l2 = C_int()
        _lib.check(lib.gad_device_info(sm, smem, l2), "gad_device_info")
        if _bwd_smem_bytes(max_nodes, max(max_edges, 1), ce) > smem.value:
            return
        T = len(tp) - 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(lib.gad_graph_check_tiles(_lib.ptr(self.rowptr), _lib.ptr(self.col), self.N,
                                                 _lib.ptr(tile_ptr), T, _lib.ptr(self._info), stream),
                       "gad_graph_check_tiles")
        if int(self._info[3].item()) != 0:
            return   # edges cross mesh boundaries: not a disjoint union -> streaming kernels
        self.tile_ptr, self.T = tile_ptr, T
        self.max_tile_nodes, self.max_tile_edges = max_nodes, max(max_edges, 1)
        self.sm_count = sm.value
        if use_ell:
            self._build_ell(ce)

    def _build_ell(self, ce: int):
        """ELL rows for the mesh-resident ELL kernels (csrc/ell_kernels.cuh): possible when every
        node has at most 7 in- and out-edges and the tile fits 16-bit row offsets / shared memory."""
        import os
        lib = _lib.load()
        self.ell_in = self.ell_out = None
        self.ell_ce, self.ell_deg = 0, 0
        if os.environ.get("GAD_NO_ELL") or ce not in (2, 4) or self.E == 0:
            return
        deg = max(self.max_in_deg, self.max_out_deg)
        if not lib.gad_ell_supported(ce, self.max_tile_nodes, deg, 1):
            return
        ell_in = torch.empty((self.N, 8), dtype=torch.int16, device=self.device)
        ell_out = torch.empty((self.N, 8), dtype=torch.int16, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._info[4:5].zero_()
            _lib.check(lib.gad_graph_build_ell(_lib.ptr(self.rowptr), _lib.ptr(self.col_walk), self.N,
                                               _lib.ptr(self.tile_ptr), self.T, ce, deg, _lib.ptr(ell_in),
                                               _lib.ptr(self._info), stream), "gad_graph_build_ell")
            _lib.check(lib.gad_graph_build_ell(_lib.ptr(self.t_rowptr), _lib.ptr(self.t_dst_walk), self.N,
                                               _lib.ptr(self.tile_ptr), self.T, ce, deg, _lib.ptr(ell_out),
                                               _lib.ptr(self._info), stream), "gad_graph_build_ell")
        if int(self._info[4].item()) != 0:
            return
        self.ell_in, self.ell_out, self.ell_ce, self.ell_deg = ell_in, ell_out, ce, deg


def _ensure_wide(self, ce: int) -> bool:
    """Wide rows {j_0..j_6, valid} for the streaming ELL kernels (csrc/stream_ell.cu): degree <= 7,
    CE in {2, 4}.  Built once, on first use of the streaming path."""
    if self._wide_tried:
        return self.wide_in is not None
    self._wide_tried = True
    if ce not in (2, 4) or self.E == 0 or max(self.max_in_deg, self.max_out_deg) > 7:
        return False
    import os
    if os.environ.get("GAD_NO_WIDE"):
        return False
    lib = _lib.load()
    wi = torch.empty((self.N, 8), dtype=torch.int32, device=self.device)
    wo = torch.empty((self.N, 8), dtype=torch.int32, device=self.device)
    stream = torch.cuda.current_stream(self.device).cuda_stream
    with torch.cuda.device(self.device):
        self._info[4:5].zero_()
        _lib.check(lib.gad_graph_build_wide(_lib.ptr(self.rowptr), _lib.ptr(self.col_walk), self.N, _lib.ptr(wi),
                                            _lib.ptr(self._info), stream), "gad_graph_build_wide")
        _lib.check(lib.gad_graph_build_wide(_lib.ptr(self.t_rowptr), _lib.ptr(self.t_dst_walk), self.N, _lib.ptr(wo),
                                            _lib.ptr(self._info), stream), "gad_graph_build_wide")
    if int(self._info[4].item()) != 0:
        return False
    self.wide_in, self.wide_out = wi, wo
    self.wide_deg = max(self.max_in_deg, self.max_out_deg)
    # bandwidth of the node numbering, max |j - i| over the edges: the persistent streaming forward sizes its
    # shared-memory window with it (csrc/stream_ell.cu: k_wide_persist)
    ids = torch.arange(self.N, device=self.device, dtype=torch.int32).unsqueeze(1)
    live = (wi[:, 7:8] >> torch.arange(7, device=self.device, dtype=torch.int32)) & 1
    self.wide_reach = int(((wi[:, :7] - ids).abs() * live).max().item())
    return True


MeshGraph.ensure_wide = _ensure_wide


def _build_cluster_rows(self, ce: int, max_cluster: int, min_cluster: int = 0):
    """(rows_in, rows_out, C, S, mesh_ptr) for clusters of at most `max_cluster` (at least `min_cluster`)
    CTAs, or None."""
    import ctypes
    lib = _lib.load()
    Cc, Sc = ctypes.c_int(0), ctypes.c_int(0)
    if lib.gad_cluster_plan(ce, max(self.mesh_sizes), max_cluster, ctypes.byref(Cc), ctypes.byref(Sc)) != 0:
        return None
    if min_cluster > Cc.value:          # more, smaller slabs than the mesh needs
        Cc.value = min_cluster
        Sc.value = ((max(self.mesh_sizes) + min_cluster - 1) // min_cluster + 3) & ~3
    M = len(self.mesh_sizes)
    mp = np.concatenate([[0], np.cumsum(np.asarray(self.mesh_sizes, dtype=np.int64))]).astype(np.int32)
    mesh_ptr = torch.from_numpy(mp).to(self.device)
    ci = torch.empty((self.N, 8), dtype=torch.int32, device=self.device)
    co = torch.empty((self.N, 8), dtype=torch.int32, device=self.device)
    stream = torch.cuda.current_stream(self.device).cuda_stream
    with torch.cuda.device(self.device):
        self._info[4:5].zero_()
        for ptr, idx, out in ((self.rowptr, self.col_walk, ci), (self.t_rowptr, self.t_dst_walk, co)):
            _lib.check(lib.gad_graph_build_cluster(_lib.ptr(ptr), _lib.ptr(idx), _lib.ptr(mesh_ptr), M, max(self.mesh_sizes),
                                                   ce, Cc.value, _lib.ptr(out), _lib.ptr(self._info), stream),
                       "gad_graph_build_cluster")
    if int(self._info[4].item()) != 0:      # an edge leaves its mesh or a row is too long
        return None
    return ci, co, Cc.value, Sc.value, mesh_ptr


def _cluster_eligible(self, ce: int) -> bool:
    import os
    return not (os.environ.get("GAD_NO_CLUSTER") or ce not in (2, 4) or self.E == 0 or self.mesh_sizes is None
                or self.tile_ptr is not None or max(self.max_in_deg, self.max_out_deg) > 7)


def _cluster_size_for(num_meshes: int) -> int:
    """Cluster size that keeps the machine busy: the largest power of two <= 16 with at most ~two waves of
    CTAs (B200 holds 132 / 120 / 112 CTAs of 4- / 8- / 16-clusters at once).  Measured on 100x100 meshes,
    training step: 4 meshes 69 / 52 / 45 us on clusters of 4 / 8 / 16, 12 meshes 70 / 54 / 51 us, 64 meshes
    139 / 237 / 217 us."""
    c = 16
    while c > 2 and num_meshes * c > 256:
        c //= 2
    return c


def _ensure_cluster(self, ce: int) -> bool:
    """Cluster rows for the cluster-resident TRAINING kernel (csrc/cl_kernels.cu: k_cl_train): meshes
    that do not fit one CTA but fit a thread-block cluster (size by _cluster_size_for; degree <= 7, CE in {2, 4},
    every mesh a connected component of the batch).  Built once, on first use."""
    if self._cl_tried:
        return self.cl_in is not None
    self._cl_tried = True
    if not _cluster_eligible(self, ce):
        return False
    want = _cluster_size_for(len(self.mesh_sizes))
    r = _build_cluster_rows(self, ce, want, want) if want > 4 else None
    if r is None:
        r = _build_cluster_rows(self, ce, 0)             # the smallest cluster that holds the mesh, at most 4
    if r is None:
        return False
    self.cl_in, self.cl_out, self.cl_C, self.cl_S, self.mesh_ptr = r
    self.cl_deg = max(self.max_in_deg, self.max_out_deg)
    return True


def _ensure_cluster_fwd(self, ce: int) -> bool:
    """Cluster rows for the forward-only kernel (k_cl_fwd), which takes clusters of up to 16 CTAs
    (a 200x200 mesh); the training rows are reused when they exist."""
    if getattr(self, "_clf_tried", False):
        return self.clf_in is not None
    self._clf_tried = True
    self.clf_in = None
    if not _cluster_eligible(self, ce):
        return False
    # A forward call is latency-bound per stage (about 1.7 us + 0.65 us per 1000 nodes of a slab, measured on
    # 64 RK4 steps): with few meshes the LARGEST cluster whose clusters are all resident at once wins (a
    # single 100x100 mesh: 0.50 ms on 16 CTAs against 0.85 ms on 4); with many meshes, the training plan.
    M = len(self.mesh_sizes)
    want = _cluster_size_for(M)
    r = _build_cluster_rows(self, ce, want, want) if want > 4 else None
    if r is not None:
        self.clf_in, _, self.clf_C, _, self.clf_mesh_ptr = r
    elif _ensure_cluster(self, ce):
        self.clf_in, self.clf_C, self.clf_mesh_ptr = self.cl_in, self.cl_C, self.mesh_ptr
    else:
        r = _build_cluster_rows(self, ce, 16)
        if r is None:
            return False
        self.clf_in, _, self.clf_C, _, self.clf_mesh_ptr = r
    self.clf_deg = max(self.max_in_deg, self.max_out_deg)
    return True


# F-evaluation time of the cluster forward kernel against the nodes of one slab (us; 64 RK4 steps on one
# 64x64 / 100x100 / 200x200 mesh over 16 CTAs, scripts/fwd_sweep.py)
_CLUSTER_FEVAL_US = ((256, 1.125), (625, 2.03), (2500, 3.5))


def _stream_fwd_preferred(self, ce: int, n_fevals: int) -> bool:
    """Forward-only dispatch between the cluster kernel (one launch, M x C CTAs, a cluster barrier per
    F-evaluation) and the streaming wide-row chain (one dependent launch per F-evaluation over ALL SMs,
    csrc/stream_ell.cu).  With few meshes the clusters leave most of the machine idle: one 200x200 mesh on
    16 CTAs takes 3.5 us per F-evaluation against 1.85 us for the chain and 1.6 us for the persistent one-launch
    streaming kernel (cfg 4: 0.87 -> 0.46 -> 0.41 ms); for 64x64 meshes at any batch, and for many large meshes,
    the cluster kernel wins.  Measured model: chain = 1.7 us + 12.2 ns per 1000 nodes of the batch per
    F-evaluation, persistent kernel = 1.37 us + 5.5 ns per 1000 nodes, plus ~3 us once for the extra pack
    launch.  Call after ensure_cluster_fwd returned True."""
    import os
    pol = os.environ.get("GAD_FWD_POLICY")
    if pol == "cluster":
        return False
    M, C = len(self.mesh_sizes), int(self.clf_C)
    if pol != "stream":
        if M * C > 128:
            return False
        slab = -(-max(self.mesh_sizes) // C)
        pts = _CLUSTER_FEVAL_US
        k = 0 if slab <= pts[1][0] else 1
        (s0, t0), (s1, t1) = pts[k], pts[k + 1]
        cluster_us = t0 + (slab - s0) * (t1 - t0) / (s1 - s0)
        if not self.ensure_wide(ce):
            return False
        # the persistent one-launch forward (k_wide_persist) where the nodes fit the co-resident threads: measured
        # 1.45 us per F-evaluation at 14 400 nodes, 1.59 us at 40 000; beyond that the chain of dependent launches
        fits = self.N <= self.persist_nodes(ce)
        if fits and os.environ.get("GAD_WIDE_PERSIST", "1") != "0":
            stream_us = 1.37 + 5.5e-6 * self.N
        else:
            stream_us = 1.7 + 1.22e-5 * self.N
        if n_fevals * (cluster_us - stream_us) <= 3.0:
            return False
    return self.ensure_wide(ce)


def _persist_nodes(self, ce: int) -> int:
    """Largest node count the persistent streaming forward (k_wide_persist) takes for this graph's wide rows on its
    device: 148 SMs x 3 co-resident CTAs x 256 nodes on a B200 while the window fits shared memory; 0 = never."""
    cached = getattr(self, "_persist_nodes_cache", None)
    if cached is None or cached[0] != ce:
        with torch.cuda.device(self.device):
            cached = (ce, int(_lib.load().gad_wide_persist_nodes(ce, self.wide_deg, self.wide_reach)))
        self._persist_nodes_cache = cached
    return cached[1]


def _stream_train_preferred(self, ce: int) -> bool:
    """Training-step dispatch for FEW very large meshes: the one-launch cluster kernel takes 43 us on one
    100x100 mesh (625 nodes per CTA of a 16-cluster) and 70 us on one 200x200 mesh (2500 per CTA, 16 SMs
    busy), the chain of ~14 streaming launches 45 us and 51 us; with four meshes or more the cluster
    kernel always wins (scripts/widebench.py).  Call after ensure_cluster returned True."""
    import os
    pol = os.environ.get("GAD_TRAIN_POLICY")
    if pol == "cluster":
        return False
    if pol != "stream":
        M, C = len(self.mesh_sizes), int(self.cl_C)
        if M * C > 128:
            return False
        slab = -(-max(self.mesh_sizes) // C)
        cluster_us = 43.0 + max(0, slab - 625) * 0.01435
        stream_us = 45.0 + max(0, self.N - 10000) * 0.00019
        if stream_us >= cluster_us - 2.0:
            return False
    return self.ensure_wide(ce)


MeshGraph.ensure_cluster = _ensure_cluster
MeshGraph.ensure_cluster_fwd = _ensure_cluster_fwd
MeshGraph.stream_train_preferred = _stream_train_preferred
MeshGraph.stream_fwd_preferred = _stream_fwd_preferred
MeshGraph.persist_nodes = _persist_nodes


def edge_masks(edge_index: torch.Tensor, side_bits: torch.Tensor):
    """Device version of the three edge masks of `firedrake_mesh_to_PyG` (`src/data.py:465-494`):
    `side_bits` uint8 [N] has bit k set when the node lies on boundary marker k + 1.  Returns bool
    tensors (to_boundary_edge_mask, to_corner_nodes_mask, diff_boundary_edges_mask)."""
    lib = _lib.load()
    if edge_index.device.type != "cuda":
        raise RuntimeError("edge_masks needs CUDA tensors (there is no CPU fallback)")
    dev = edge_index.device
    ei = edge_index.to(torch.int64).contiguous()
    sb = side_bits.to(device=dev, dtype=torch.uint8).contiguous()
    E, N = int(ei.shape[1]), int(sb.numel())
    out = [torch.empty(E, dtype=torch.uint8, device=dev) for _ in range(3)]
    info = torch.zeros(8, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.gad_edge_masks(_lib.ptr(ei), E, _lib.ptr(sb), N, _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]),
                                      _lib.ptr(info), torch.cuda.current_stream(dev).cuda_stream), "gad_edge_masks")
    if int(info[7].item()) != 0:
        raise ValueError(f"edge_index holds {int(info[7])} node ids outside [0, {N})")
    return tuple(o.bool() for o in out)


def C_int():
    import ctypes
    return ctypes.c_int(0)


# ------------------------------------------------------------------------------------------
# prologue of GNN.forward: corner loops + cache
# ------------------------------------------------------------------------------------------
def corner_loops(data, dim: int, mesh_dims, mesh_sizes: Sequence[int]) -> Optional[torch.Tensor]:
    """Node ids that receive a self-loop (`src/GNN.py:209-218`): the two end points of every
    1-D mesh (canonical node order, :210) or the per-mesh `corner_nodes` plus node offsets."""
    offs = np.concatenate([[0], np.cumsum(np.asarray(mesh_sizes, dtype=np.int64))[:-1]])
    if dim == 1:
        n = int(mesh_dims[0])
        b = np.arange(len(mesh_sizes), dtype=np.int64)
        ids = np.stack([b * n, (b + 1) * n - 1], axis=1).reshape(-1)
    else:
        corners = [np.asarray(c, dtype=np.int64).reshape(-1) for c in data.corner_nodes]
        if len({len(c) for c in corners}) > 1:
            raise ValueError("every mesh must carry the same number of corner nodes (torch.stack, GNN.py:213)")
        ids = (np.stack(corners) + offs[:, None]).reshape(-1) if corners else np.zeros(0, dtype=np.int64)
    return torch.from_numpy(ids)


class GraphCache:
    """Small LRU keyed on the identity of the topology tensors.  Entries keep those tensors alive,
    so a data pointer cannot be recycled for different content while its entry exists."""

    def __init__(self, capacity: int = 8):
        self.capacity = capacity
        self._d = collections.OrderedDict()
        self.hits = 0
        self.misses = 0

    @staticmethod
    def key_of(data, flags) -> tuple:
        ei = data.edge_index
        parts = [ei.data_ptr(), tuple(ei.shape), ei._version, str(ei.device)]
        for name in ("to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask", "batch"):
            t = getattr(data, name, None)
            parts += [None if t is None else (t.data_ptr(), t._version)]
        return tuple(parts) + tuple(flags)

    TOPOLOGY_FIELDS = ("edge_index", "to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask", "batch")

    @staticmethod
    def content_key_of(data, flags, device, use_masks: bool = True):
        """(key, device copies): key = 128-bit device fingerprint (csrc/graph_build.cu: k_fingerprint)
        of edge_index / masks / batch + a host hash of the corner-node lists + `flags`.  Costs the
        host -> device copies the graph build needs anyway, one pass over them at HBM speed and one
        16-byte read-back."""
        import hashlib
        lib = _lib.load()
        dev = torch.device(device)
        out = torch.zeros(2, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        copies, shapes = {}, []
        with torch.cuda.device(dev):
            for k, name in enumerate(GraphCache.TOPOLOGY_FIELDS):
                t = getattr(data, name, None)
                if t is None or (not use_masks and name.endswith("_mask")):
                    shapes.append(None)
                    continue
                td = t.to(dev, non_blocking=True).contiguous()
                if td.data_ptr() % 8:
                    td = td.clone()
                copies[name] = td
                shapes.append((tuple(td.shape), str(td.dtype)))
                _lib.check(lib.gad_fingerprint(_lib.ptr(td), td.numel() * td.element_size(), 0x1234 + 7919 * k,
                                               _lib.ptr(out), stream), "gad_fingerprint")
        h = hashlib.blake2b(digest_size=16)
        for c in (getattr(data, "corner_nodes", None) or []):
            h.update(np.ascontiguousarray(np.asarray(c, dtype=np.int64)).tobytes())
            h.update(b"|")
        sizes = getattr(data, "mesh_sizes", None)
        if sizes is not None:
            h.update(np.asarray(list(sizes), dtype=np.int64).tobytes())
        fp = tuple(out.tolist())     # the one synchronisation
        return ("content", fp, tuple(shapes), h.hexdigest()) + tuple(flags), copies

    SHARED_SAMPLES = 64

    @staticmethod
    def sample_positions(E: int) -> torch.Tensor:
        """SHARED_SAMPLES evenly spaced positions in [0, E), first and last included.  Integer arithmetic: a float32
        `linspace(0, E - 1, n)` rounds its end point up past the array beyond 2^24 edges (cfg 5 has 1.2e8)."""
        E = int(E)
        pos = GraphCache._positions.get(E)
        if pos is None:
            n = min(E, GraphCache.SHARED_SAMPLES)
            pos = (torch.arange(n, dtype=torch.int64) * (E - 1)) // max(n - 1, 1)
            if len(GraphCache._positions) < 64:          # a handful of batch shapes per run
                GraphCache._positions[E] = pos
        return pos

    _positions: dict = {}

    @staticmethod
    def _sampled(t: torch.Tensor, idx: torch.Tensor, columns: bool):
        """The sampled entries of a topology tensor as a hashable value: bytes for host tensors (this runs on every
        module call with a fresh Batch), a tuple after one small copy for device tensors."""
        if t.device.type == "cpu":
            a, ix = t.detach().numpy(), idx.numpy()
            return (a[:, ix] if columns else a[ix]).tobytes()
        sel = t[:, idx.to(t.device)] if columns else t[idx.to(t.device)]
        return tuple(sel.reshape(-1).tolist())

    @staticmethod
    def shared_key_of(data, flags) -> tuple:
        """Key for datasets whose samples all live on ONE mesh (the reference's `randg` data: `dataset.mesh`,
        `x_comp_shared`, src/data.py:143), opted into with `opt['gad_shared_topology']`: batches of the same
        shape then have the same topology, so a fresh Batch is recognised from shapes, mesh sizes and
        SHARED_SAMPLES columns of `edge_index` / entries of the masks read on the host -- no host -> device
        copy of the topology (26 MB per step on cfg 2), no fingerprint pass.  The samples catch a batch that
        breaks the promise only with high probability, which is why this is an option and the content
        fingerprint (content_key_of) the default."""
        ei = data.edge_index
        E = int(ei.shape[1])
        parts = [tuple(ei.shape), int(data.x_comp.shape[0])]
        sizes = getattr(data, "mesh_sizes", None)
        if sizes is None:
            parts.append(None)
        else:       # equal-size batches (the shared-mesh case) compare as one C-level pass over the list
            n0 = sizes[0] if len(sizes) else 0
            same = sizes.count(n0) == len(sizes) if isinstance(sizes, list) else False
            parts.append((len(sizes), int(n0), int(n0)) if same else (len(sizes), int(min(sizes)), int(max(sizes))))
        if E > 0:
            idx = GraphCache.sample_positions(E)
            parts.append(GraphCache._sampled(ei, idx, True))
            for name in ("to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask"):
                t = getattr(data, name, None)
                parts.append(None if t is None else GraphCache._sampled(t, idx, False))
        return ("shared",) + tuple(parts) + tuple(flags)

    def alias(self, key, graph, keepalive):
        """Register another identity key for an existing graph (kept alive like any entry)."""
        self._d[key] = graph
        self._alias_keep = getattr(self, "_alias_keep", collections.OrderedDict())
        self._alias_keep[key] = keepalive
        while len(self._alias_keep) > 4 * self.capacity:
            self._alias_keep.popitem(last=False)
        self._trim()

    def _trim(self):
        while len(self._d) > 3 * self.capacity:
            k, _ = self._d.popitem(last=False)
            if hasattr(self, "_alias_keep"):
                self._alias_keep.pop(k, None)

    def get(self, key):
        g = self._d.get(key)
        if g is not None:
            self._d.move_to_end(key)
            self.hits += 1
        return g

    def put(self, key, graph, keepalive):
        self.misses += 1
        graph._keepalive = keepalive
        self._d[key] = graph
        self._trim()

    def clear(self):
        self._d.clear()
