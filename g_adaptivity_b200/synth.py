"""Synthetic meshes, features and PyG-like batches for the deformer hot path.

The reference builds its graphs from Firedrake meshes (`src/data.py:424-502`,
`firedrake_mesh_to_PyG`) and collates them with PyG's `Batch`.  Neither Firedrake
nor PyG exists in this image, so this module produces the same *objects* from
structured meshes:

* edges: both directions of every triangle (interval) side, de-duplicated through a
  CPython ``set`` of ``(int, int)`` tuples filled cell by cell and listed with
  ``list(set)`` -- the exact construction of `src/data.py:430-441`, so the edge
  order is the same arbitrary-but-deterministic permutation the reference sees;
* the three edge masks of `src/data.py:465-494`, `corner_nodes`, `boundary_nodes`;
* node features from the analytic multi-Gaussian Poisson data of
  `firedrake_difFEM/solve_poisson.py:78-79,145-147` (u and its Laplacian at the nodes);
* a `Batch` attribute bag with the fields `GNN.forward` reads (`src/GNN.py:191-217`).

Nothing here is on the timed path; it only manufactures inputs.
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# attribute bags standing in for torch_geometric.data.Data / Batch
# --------------------------------------------------------------------------------------
class Data:
    """Minimal stand-in for `torch_geometric.data.Data`: attributes + `.to()`."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device, non_blocking: bool = False):
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def clone(self):
        out = copy.copy(self)
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                setattr(out, k, v.clone())
        return out

    def pin_memory(self):
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor) and v.device.type == "cpu":
                setattr(self, k, v.pin_memory())
        return self

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                parts.append(f"{k}={list(v.shape)}")
        return f"{type(self).__name__}({', '.join(parts)})"


class Batch(Data):
    """Disjoint union of `Data` objects (PyG `Batch` semantics: node-level tensors are
    concatenated, `edge_index` is offset by the cumulative node count, non-tensor
    attributes become python lists, `batch[i]` is the graph id of node i)."""

    @property
    def num_graphs(self) -> int:
        return int(self._num_graphs)

    @staticmethod
    def from_data_list(data_list: Sequence[Data]) -> "Batch":
        node_keys = ("x_comp", "x_phys", "f_tensor", "uu_tensor", "u_true_tensor", "boundary_nodes")
        edge_keys = ("to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask")
        out = Batch()
        offs, ei, bvec = 0, [], []
        for b, d in enumerate(data_list):
            n = d.x_comp.shape[0]
            ei.append(d.edge_index + offs)
            bvec.append(torch.full((n,), b, dtype=torch.long))
            offs += n
        out.edge_index = torch.cat(ei, dim=1)
        out.batch = torch.cat(bvec)
        for k in node_keys + edge_keys:
            if all(hasattr(d, k) for d in data_list):
                setattr(out, k, torch.cat([getattr(d, k) for d in data_list], dim=0))
        out.corner_nodes = [d.corner_nodes for d in data_list]
        out.pde_params = {
            "centers": [d.pde_params["centers"] for d in data_list],
            "scales": [d.pde_params["scales"] for d in data_list],
        }
        out._num_graphs = len(data_list)
        out.mesh_sizes = [int(d.x_comp.shape[0]) for d in data_list]
        return out


class SyntheticMesh:
    """What the 2-D FEM tail asks of `dataset.mesh` (a Firedrake mesh in the reference): the cell-node map
    (`mesh.coordinates.cell_node_map().values`, difFEM_2d.py:363) and the Dirichlet nodes (the reference gets them
    from `DirichletBC(V, 0, "on_boundary").nodes`, :354-356; a stand-in carries them as `bc_nodes`)."""

    def __init__(self, cells: np.ndarray, bc_nodes: np.ndarray):
        self._cells = np.asarray(cells, dtype=np.int32)
        self.bc_nodes = np.asarray(bc_nodes, dtype=np.int32)
        self.coordinates = self

    def cell_node_map(self):
        class _Map:
            values = self._cells
        return _Map()


class SyntheticDataset:
    """What `GNN.__init__` / `GNN.forward` read from their `dataset` argument (`src/GNN.py:147-149,321,333`).
    With `eval_quad_points` (2-D) it also carries what `loss_type='pde_loss'` needs: `mesh` and
    `mapping_tensor_fine`, the canonical-grid -> fine-mesh-node map of `map_firedrake_to_cannonical_ordering_2d`
    (utils_data.py:53-77) for a fine mesh numbered row-major (node id = iy * Q + ix)."""

    def __init__(self, dim: int, mesh_dims: Sequence[int], eval_quad_points: Optional[int] = None):
        self.num_x_comp_features = dim
        self.mesh_dims = list(mesh_dims)
        self.mesh = None  # Firedrake mesh in the reference; only used by reg_skew / pde_loss
        self.x_comp_shared = None
        self.mapping_tensor = None
        self.mapping_tensor_fine = None
        if dim == 2:
            topo = MeshTopology(mesh_dims)
            self.mesh = SyntheticMesh(topo.cells, np.nonzero(topo.boundary_nodes)[0])
            self.x_comp_shared = torch.from_numpy(topo.coords.copy())
            # canonical-grid -> mesh-node map of the n x n mesh itself (map_firedrake_to_cannonical_ordering_2d,
            # utils_data.py:53-77; read by the global CNN features, src/GNN.py:245)
            n = int(mesh_dims[0])
            i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
            self.mapping_tensor = torch.from_numpy((j * n + i).reshape(-1).astype(np.int64))
            if eval_quad_points is not None:
                Q = int(eval_quad_points)
                # canonical index c = i * Q + j is the grid point (x_i, y_j) (torch.meshgrid 'ij', :60-62,71);
                # on the row-major fine mesh that point is node j * Q + i
                i, j = np.meshgrid(np.arange(Q), np.arange(Q), indexing="ij")
                self.mapping_tensor_fine = torch.from_numpy((j * Q + i).reshape(-1).astype(np.int64))


# --------------------------------------------------------------------------------------
# structured meshes -> reference graph objects
# --------------------------------------------------------------------------------------
def _edges_from_cells(cells: np.ndarray) -> np.ndarray:
    """`src/data.py:430-441`: set of directed pairs, listed in CPython set order."""
    edges_set = set()
    for cell in cells.tolist():
        k = len(cell)
        for i in range(k):
            for j in range(i + 1, k):
                edges_set.add((cell[i], cell[j]))
                edges_set.add((cell[j], cell[i]))
    return np.asarray(list(edges_set), dtype=np.int64).T.copy()  # [2, E0]


def unit_square_cells(n: int) -> np.ndarray:
    """Triangles of the n x n-node unit square, every grid square split along its
    anti-diagonal; node id = iy * n + ix."""
    ix, iy = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="xy")
    a = (iy * n + ix).ravel()
    b, c, d = a + 1, a + n, a + n + 1
    lower = np.stack([a, b, c], axis=1)
    upper = np.stack([b, d, c], axis=1)
    return np.stack([lower, upper], axis=1).reshape(-1, 3).astype(np.int64)


def _masks_from_sides(edge_index: np.ndarray, side_lists: List[np.ndarray], num_nodes: int):
    """Edge masks of `src/data.py:443-494` from the per-side boundary node lists."""
    on_boundary = np.zeros(num_nodes, dtype=bool)
    membership = np.zeros((num_nodes, len(side_lists)), dtype=bool)
    for s, nodes in enumerate(side_lists):
        on_boundary[nodes] = True
        membership[nodes, s] = True
    all_nodes = np.concatenate(side_lists) if side_lists else np.zeros(0, dtype=np.int64)
    uniq, counts = np.unique(all_nodes, return_counts=True)
    corner_nodes = uniq[counts > 1]
    is_corner = np.zeros(num_nodes, dtype=bool)
    is_corner[corner_nodes] = True
    src, dst = edge_index[0], edge_index[1]
    to_boundary = on_boundary[dst] & ~on_boundary[src]
    to_corner = is_corner[dst]
    # `node_boundary_map[src] != node_boundary_map[dst]`: lists of side ids in side order
    differ = (membership[src] != membership[dst]).any(axis=1)
    diff_boundary = on_boundary[src] & on_boundary[dst] & differ & ~is_corner[src] & ~is_corner[dst]
    return on_boundary, corner_nodes.astype(np.int64), to_boundary, to_corner, diff_boundary


class MeshTopology:
    """Everything about one mesh that does not depend on the PDE sample."""

    def __init__(self, mesh_dims: Sequence[int]):
        self.mesh_dims = list(mesh_dims)
        self.dim = len(mesh_dims)
        if self.dim == 2:
            n = int(mesh_dims[0])
            assert int(mesh_dims[1]) == n, "square meshes only (reference: UnitSquareMesh(n-1, n-1))"
            xs = np.linspace(0.0, 1.0, n, dtype=np.float32)
            X, Y = np.meshgrid(xs, xs, indexing="xy")
            self.coords = np.stack([X.ravel(), Y.ravel()], axis=1).astype(np.float32)  # [N, 2]
            self.cells = unit_square_cells(n)
            ids = np.arange(n * n).reshape(n, n)  # [iy, ix]
            # Firedrake UnitSquareMesh markers: 1: x=0, 2: x=1, 3: y=0, 4: y=1
            sides = [ids[:, 0].copy(), ids[:, -1].copy(), ids[0, :].copy(), ids[-1, :].copy()]
            self.num_nodes = n * n
        elif self.dim == 1:
            n = int(mesh_dims[0])
            self.coords = np.linspace(0.0, 1.0, n, dtype=np.float32)  # [N] (interval coords are 1-D)
            self.cells = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1).astype(np.int64)
            sides = [np.array([0]), np.array([n - 1])]
            self.num_nodes = n
        else:
            raise ValueError("mesh_dims must have 1 or 2 entries")
        self.edge_index = _edges_from_cells(self.cells)
        (self.boundary_nodes, self.corner_nodes, self.to_boundary_edge_mask,
         self.to_corner_nodes_mask, self.diff_boundary_edges_mask) = _masks_from_sides(
            self.edge_index, sides, self.num_nodes)


# --------------------------------------------------------------------------------------
# PDE features (analytic Gaussians)
# --------------------------------------------------------------------------------------
def gaussian_features(coords: np.ndarray, centers: np.ndarray, scales: np.ndarray, amplitude: float = 1.0):
    """u = sum_g A exp(-sum_d (x_d-c_d)^2/s_d^2) and the reference's forcing term.

    2D: f = -Laplace(u) (`solve_poisson.py:145-147`);  1D: f = +u'' (`solve_poisson.py:78-79`).
    coords [..., N, dim] float32, centers/scales [..., G, dim]; returns (u, f) float32 [..., N].
    """
    x = coords.astype(np.float64)
    if x.ndim == 1:
        x = x[:, None]
    c = centers.astype(np.float64)
    s = scales.astype(np.float64)
    diff = x[..., :, None, :] - c[..., None, :, :]           # [..., N, G, dim]
    s2 = (s * s)[..., None, :, :]
    g = amplitude * np.exp(-(diff * diff / s2).sum(-1))        # [..., N, G]
    lap = ((4.0 * diff * diff / (s2 * s2)) - 2.0 / s2).sum(-1) * g
    u = g.sum(-1)
    f = lap.sum(-1)
    if x.shape[-1] == 2:
        f = -f
    return u.astype(np.float32), f.astype(np.float32)


def sample_gaussians(rng: np.random.Generator, dim: int, num_gauss: int, burgers: bool = False,
                     scale: float = 0.1, limits: float = 3.0):
    """`src/data.py:147-158`: c ~ U(0,1), s ~ U(0.1,0.5) (Poisson) or the Burgers ranges."""
    cs, ss = [], []
    for _ in range(num_gauss):
        if burgers:
            s = rng.uniform(scale * 0.5, scale * 2.0, dim).astype("f")
            c = rng.uniform(scale * limits, 1 - scale * limits, dim).astype("f")
        else:
            c = rng.uniform(0, 1, dim).astype("f")
            s = rng.uniform(0.1, 0.5, dim).astype("f")
        cs.append(c)
        ss.append(s)
    return np.stack(cs), np.stack(ss)


def target_mesh(coords: np.ndarray, n: int) -> np.ndarray:
    """Deterministic stand-in for the MA/MMPDE target mesh `x_phys` (`src/data.py:205-212`):
    the computational mesh plus a smooth 0.3/n perturbation that vanishes on the boundary."""
    amp = 0.3 / n
    if coords.ndim == 1:
        x = coords.astype(np.float64)
        return (x + amp * np.sin(2 * np.pi * x)).astype(np.float32)
    x, y = coords[:, 0].astype(np.float64), coords[:, 1].astype(np.float64)
    bump = np.sin(np.pi * x) * np.sin(np.pi * y)
    tx = x + amp * np.sin(2 * np.pi * x) * bump
    ty = y + amp * np.sin(2 * np.pi * y) * bump
    return np.stack([tx, ty], axis=1).astype(np.float32)


# --------------------------------------------------------------------------------------
# batches
# --------------------------------------------------------------------------------------
def make_data(topo: MeshTopology, seed: int, num_gauss: int = 2, burgers: bool = False,
              amplitude: float = 1.0) -> Data:
    """One sample, as `MeshInMemoryDataset.process` would store it (`src/data.py:263-276`)."""
    rng = np.random.default_rng(seed)
    c, s = sample_gaussians(rng, topo.dim, num_gauss, burgers=burgers)
    u, f = gaussian_features(topo.coords, c, s, amplitude)
    return Data(
        x_comp=torch.from_numpy(topo.coords.copy()),
        x_phys=torch.from_numpy(target_mesh(topo.coords, topo.mesh_dims[0])),
        edge_index=torch.from_numpy(topo.edge_index.copy()),
        boundary_nodes=torch.from_numpy(topo.boundary_nodes.copy()),
        corner_nodes=topo.corner_nodes.copy(),
        to_boundary_edge_mask=torch.from_numpy(topo.to_boundary_edge_mask.copy()),
        to_corner_nodes_mask=torch.from_numpy(topo.to_corner_nodes_mask.copy()),
        diff_boundary_edges_mask=torch.from_numpy(topo.diff_boundary_edges_mask.copy()),
        f_tensor=torch.from_numpy(f), uu_tensor=torch.from_numpy(u.copy()),
        u_true_tensor=torch.from_numpy(u.copy()),
        pde_params={"centers": list(c), "scales": list(s)},
    )


def make_batch(mesh_dims: Sequence[int], num_meshes: int, seed: int = 0, num_gauss: Optional[int] = None,
               burgers: bool = False, first_mesh_id: int = 0, eval_quad_points: int = 101,
               with_u_true_fine: bool = False) -> Batch:
    """`num_meshes` samples on one shared topology, collated like PyG would.

    Vectorised (no per-mesh python graph work), bit-identical to
    `Batch.from_data_list([make_data(topo, seed + i) ...])`; sample i of the batch is seeded
    with `seed + first_mesh_id + i` so shards of one global batch can be generated per rank."""
    topo = MeshTopology(mesh_dims)
    dim, N, B = topo.dim, topo.num_nodes, int(num_meshes)
    if num_gauss is None:
        num_gauss = 1 if burgers else 2
    amplitude = 0.25 if burgers else 1.0
    cs = np.empty((B, num_gauss, dim), dtype=np.float32)
    ss = np.empty((B, num_gauss, dim), dtype=np.float32)
    for b in range(B):
        rng = np.random.default_rng(seed + first_mesh_id + b)
        cs[b], ss[b] = sample_gaussians(rng, dim, num_gauss, burgers=burgers)
    u = np.empty((B, N), dtype=np.float32)
    f = np.empty((B, N), dtype=np.float32)
    chunk = max(1, (1 << 22) // (N * num_gauss))
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        u[b0:b1], f[b0:b1] = gaussian_features(topo.coords, cs[b0:b1], ss[b0:b1], amplitude)
    offs = (np.arange(B, dtype=np.int64) * N)
    out = Batch()
    out.edge_index = torch.from_numpy(
        (topo.edge_index[:, None, :] + offs[None, :, None]).reshape(2, -1).copy())
    out.batch = torch.from_numpy(np.repeat(np.arange(B, dtype=np.int64), N))
    rep = (lambda a: torch.from_numpy(np.tile(a, (B,) + (1,) * (a.ndim - 1)).copy()))
    out.x_comp = rep(topo.coords)
    out.x_phys = rep(target_mesh(topo.coords, topo.mesh_dims[0]))
    out.boundary_nodes = rep(topo.boundary_nodes)
    out.to_boundary_edge_mask = rep(topo.to_boundary_edge_mask)
    out.to_corner_nodes_mask = rep(topo.to_corner_nodes_mask)
    out.diff_boundary_edges_mask = rep(topo.diff_boundary_edges_mask)
    out.f_tensor = torch.from_numpy(f.reshape(-1))
    out.uu_tensor = torch.from_numpy(u.reshape(-1).copy())
    out.u_true_tensor = torch.from_numpy(u.reshape(-1).copy())
    out.corner_nodes = [topo.corner_nodes.copy() for _ in range(B)]
    out.pde_params = {"centers": [list(c) for c in cs], "scales": [list(s) for s in ss]}
    if topo.dim == 1 and not burgers:
        # target of loss_type='pde_loss' (src/run_GNN.py:109-110): u_true at the evaluation points of every mesh
        q = np.linspace(0.0, 1.0, eval_quad_points, dtype=np.float32)
        fine = np.zeros((B, eval_quad_points), dtype=np.float32)
        for b in range(B):
            for c, s_ in zip(cs[b], ss[b]):
                fine[b] += np.exp(-(q - np.float32(np.asarray(c).reshape(-1)[0])) ** 2
                                  / np.float32(np.asarray(s_).reshape(-1)[0]) ** 2).astype(np.float32)
        out.u_true_fine_tensor = torch.from_numpy(fine.reshape(-1))
    if topo.dim == 2 and with_u_true_fine:
        # target of loss_type='pde_loss' in 2-D (src/run_GNN.py:109-110): u_true at the nodes of the fine mesh
        # (eval_quad_points x eval_quad_points, row-major numbering: node id = iy * Q + ix)
        q = np.linspace(0.0, 1.0, eval_quad_points, dtype=np.float32)
        Xf, Yf = np.meshgrid(q, q, indexing="xy")
        fine_xy = np.stack([Xf.ravel(), Yf.ravel()], axis=1).astype(np.float32)
        fine = np.empty((B, fine_xy.shape[0]), dtype=np.float32)
        for b0 in range(0, B, 64):
            b1 = min(B, b0 + 64)
            fine[b0:b1], _ = gaussian_features(fine_xy, cs[b0:b1], ss[b0:b1], amplitude)
        out.u_true_fine_tensor = torch.from_numpy(fine.reshape(-1))
    out._num_graphs = B
    out.mesh_sizes = [N] * B
    return out


def make_batch_device(mesh_dims: Sequence[int], num_meshes: int, device, seed: int = 0, block: int = 256,
                      burgers: bool = False) -> Batch:
    """A large batch assembled ON THE DEVICE for timing runs (BASELINE cfg 5: 8192 meshes of 50x50 = 20 M nodes,
    120 M edges): `block` distinct samples are generated on the host (`make_batch`) and tiled; the topology tensors
    are the one-mesh arrays plus per-mesh node offsets, exactly what PyG's collation yields, built with torch ops
    on `device`.  Sample values repeat every `block` meshes -- irrelevant for throughput, stated by the caller."""
    B = int(num_meshes)
    base = make_batch(mesh_dims, min(B, block), seed=seed, burgers=burgers)
    nb = base.num_graphs
    topo = MeshTopology(mesh_dims)
    N1, dev = topo.num_nodes, torch.device(device)
    reps = (B + nb - 1) // nb
    out = Batch()
    ei1 = torch.from_numpy(topo.edge_index).to(dev)                                   # [2, E0]
    offs = (torch.arange(B, device=dev, dtype=torch.int64) * N1).view(1, B, 1)
    out.edge_index = (ei1.view(2, 1, -1) + offs).reshape(2, -1).contiguous()
    out.batch = torch.arange(B, device=dev, dtype=torch.int64).repeat_interleave(N1)

    def tile_nodes(t):
        t = t.to(dev)
        per = t.shape[0] // nb
        return t.repeat((reps,) + (1,) * (t.dim() - 1))[:B * per].contiguous()

    for k in ("x_comp", "x_phys", "f_tensor", "uu_tensor", "u_true_tensor", "boundary_nodes", "to_boundary_edge_mask",
              "to_corner_nodes_mask", "diff_boundary_edges_mask"):
        setattr(out, k, tile_nodes(getattr(base, k)))
    out.corner_nodes = [topo.corner_nodes.copy() for _ in range(B)]
    out.pde_params = {"centers": (base.pde_params["centers"] * reps)[:B], "scales": (base.pde_params["scales"] * reps)[:B]}
    out._num_graphs = B
    out.mesh_sizes = [N1] * B
    return out


def make_mixed_batch(mesh_dims_list: Sequence[Sequence[int]], seed: int = 0) -> Batch:
    """Variable-size batch (the `randg_mix` input shape, `src/data_mixed.py:122-146`)."""
    topos = {}
    datas = []
    for i, md in enumerate(mesh_dims_list):
        key = tuple(md)
        if key not in topos:
            topos[key] = MeshTopology(md)
        datas.append(make_data(topos[key], seed + i))
    return Batch.from_data_list(datas)


def default_opt(mesh_dims: Sequence[int] = (15, 15), **overrides) -> dict:
    """The `opt` dict after `get_params()` + `tf_sweep_args` + `run_params()` for the GNN model
    (`src/params.py:8-161,199-303`), restricted to the keys the deformer reads,
    with `loss_type='mesh_loss'` so that `forward` returns `x_phys` (`src/GNN.py:303-304`)."""
    dim = len(mesh_dims)
    opt = dict(
        mesh_dims=list(mesh_dims), hidden_dim=8, num_layers=4, time_step=0.1, learn_step=False,
        share_conv=True, conv_type="GRAND_plus", residual=True, non_lin="identity", enc="identity",
        dec="identity", dropout=0.0, fix_boundary=True, self_loops=False, gnn_inc_feat_f=True,
        gnn_inc_feat_uu=True, gnn_inc_glob_feat_f=False, gnn_inc_glob_feat_uu=False,
        gnn_normalize=False, global_feat_dim=8, softmax_temp_type=None, softmax_temp=2.0,
        reg_skew=False, show_mesh_evol_plots=True, loss_type="mesh_loss", data_type="randg",
        device="cpu", eval_quad_points=101, load_quad_points=101, stiff_quad_points=3, loss_fn="l1", lr=0.001, decay=0.0, seed=42,
        pde_type="Poisson" if dim == 2 else "Poisson",
    )
    opt.update(overrides)
    return opt


def burgers_opt(mesh_dims: Sequence[int] = (21,), **overrides) -> dict:
    """`run_params` Burgers preset (`src/params.py:136-159`): GRAND conv, features [x, uu]."""
    opt = default_opt(mesh_dims, conv_type="GRAND", gnn_inc_feat_f=False, loss_type="modular",
                      pde_type="Burgers")
    opt.update(overrides)
    return opt
