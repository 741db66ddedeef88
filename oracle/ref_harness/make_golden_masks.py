"""Mint tests/golden_masks/*.pt by running the REFERENCE's own `firedrake_mesh_to_PyG`
(/root/reference/src/data.py:424-502, executed in place, unmodified) behind a stub Firedrake mesh.

    python -m oracle.ref_harness.make_golden_masks        # build container only

`src/data.py` imports torch_geometric, matplotlib, firedrake, utils_eval and two modules that do not even
exist in the reference tree (`params_poisson`, `firedrake_difFEM.difFEM_poisson_1d`) at module top.  All of
them are replaced by permissive empty modules; only what the function touches is given behaviour:

  * `torch_geometric.data.Data`            attribute bag;  `InMemoryDataset` = object (a base class at module top);
  * `firedrake.FunctionSpace(mesh,"CG",1)`  -> handle with `.boundary_nodes(marker)` = the side's node list;
  * `firedrake.DirichletBC(V, 0, where).nodes` -> all boundary nodes (`"on_boundary"`, sorted unique, as
    Firedrake returns them) or the side's node list;
  * the mesh: `coordinates.dat.data_ro`, `coordinates.cell_node_map().values` (int32, as Firedrake's maps are),
    `topology.exterior_facets.unique_markers`;
  * `utils_data` is the REAL reference module (for `convert_to_boundary_mask`), wandb / matplotlib stubbed.

Each fixture stores the mesh (cells, per-marker side lists, coordinates) and everything the function
returned: `edge_index` (order = CPython `list(set)` as the reference produces it), `boundary_nodes`,
`corner_nodes` and the three edge masks.  tests/test_masks_golden.py pins `synth.MeshTopology` /
`synth._masks_from_sides` (CPU) and `gad_edge_masks` (GPU) against them."""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _REPO)
REFERENCE_SRC = "/root/reference/src"
GOLDEN_DIR = os.path.join(_REPO, "tests", "golden_masks")


class _Permissive(types.ModuleType):
    """Module whose every missing attribute is None (`from x import a, b` succeeds)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return None


class _Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Space:
    def __init__(self, mesh):
        self.mesh = mesh

    def boundary_nodes(self, marker):
        return self.mesh.sides[marker]


class _BC:
    def __init__(self, V, value, where):
        m = V.mesh
        if where == "on_boundary":
            self.nodes = np.unique(np.concatenate([m.sides[k] for k in m.markers])).astype(np.int32)
        else:
            self.nodes = m.sides[where]


class StubMesh:
    """What firedrake_mesh_to_PyG reads from a Firedrake mesh."""

    def __init__(self, coords: np.ndarray, cells: np.ndarray, sides: dict):
        self._coords = coords
        self._cells = cells.astype(np.int32)
        self.sides = {int(k): np.asarray(v, dtype=np.int32) for k, v in sides.items()}
        self.markers = sorted(self.sides)
        self.coordinates = types.SimpleNamespace(
            dat=types.SimpleNamespace(data_ro=self._coords),
            cell_node_map=lambda: types.SimpleNamespace(values=self._cells))
        self.topology = types.SimpleNamespace(
            exterior_facets=types.SimpleNamespace(unique_markers=np.asarray(self.markers, dtype=np.int32)))


def load_reference_data_module():
    if not os.path.isfile(os.path.join(REFERENCE_SRC, "data.py")):
        raise RuntimeError("/root/reference is not present: the harness only runs in the build container")
    names = ("torch_geometric", "torch_geometric.data", "torch_geometric.utils", "matplotlib", "matplotlib.pyplot",
             "firedrake", "firedrake.pyplot", "utils_eval", "firedrake_difFEM", "firedrake_difFEM.solve_poisson",
             "firedrake_difFEM.difFEM_poisson_1d", "classical_meshing", "classical_meshing.ma_mesh_2d",
             "classical_meshing.ma_mesh_1d", "params_poisson", "wandb", "networkx")
    saved = {n: sys.modules.get(n) for n in names + ("utils_data", "data", "params")}
    for n in names:
        sys.modules[n] = _Permissive(n)
    for pkg in ("torch_geometric", "matplotlib", "firedrake", "firedrake_difFEM", "classical_meshing"):
        sys.modules[pkg].__path__ = []
    pyg = sys.modules["torch_geometric"]
    pyg.data = sys.modules["torch_geometric.data"]
    pyg.data.Data = _Data
    pyg.data.InMemoryDataset = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    fd = sys.modules["firedrake"]
    fd.FunctionSpace = lambda mesh, family, degree: _Space(mesh)
    fd.DirichletBC = _BC
    sys.path.insert(0, REFERENCE_SRC)
    try:
        for n in ("utils_data", "data", "params"):
            sys.modules.pop(n, None)
        mod = importlib.import_module("data")
        assert os.path.realpath(mod.__file__) == os.path.join(REFERENCE_SRC, "data.py"), mod.__file__
        ud = sys.modules["utils_data"]
        assert os.path.realpath(ud.__file__).startswith("/root/reference/"), ud.__file__
    finally:
        sys.path.remove(REFERENCE_SRC)
        for n, m in saved.items():           # leave no stubs behind for whoever imports next
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return mod


def structured_mesh(mesh_dims):
    from g_adaptivity_b200 import synth
    topo = synth.MeshTopology(mesh_dims)
    n = topo.num_nodes
    if topo.dim == 2:
        m = int(mesh_dims[0])
        ids = np.arange(n).reshape(m, m)
        sides = {1: ids[:, 0], 2: ids[:, -1], 3: ids[0, :], 4: ids[-1, :]}   # UnitSquareMesh markers
        coords = topo.coords
    else:
        sides = {1: np.array([0]), 2: np.array([n - 1])}
        coords = topo.coords.reshape(-1, 1)
    return coords, topo.cells, sides


def permuted(coords, cells, sides, seed):
    """The same mesh under a random node renumbering (Firedrake's DMPlex numbering is not row-major)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(coords.shape[0])          # new id of old node i
    new_coords = np.empty_like(coords)
    new_coords[perm] = coords
    return new_coords, perm[cells], {k: perm[v] for k, v in sides.items()}


CASES = [
    ("sq15", (15, 15), None), ("sq7", (7, 7), None), ("sq4", (4, 4), None), ("sq3", (3, 3), None), ("sq2", (2, 2), None),
    ("line200", (200,), None), ("line5", (5,), None), ("line2", (2,), None),
    ("sq6_perm", (6, 6), 11), ("line9_perm", (9,), 12),
]


def main():
    ref = load_reference_data_module()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, md, seed in CASES:
        coords, cells, sides = structured_mesh(md)
        if seed is not None:
            coords, cells, sides = permuted(coords, cells, sides, seed)
        mesh = StubMesh(coords, cells, sides)
        d = ref.firedrake_mesh_to_PyG(mesh)
        fx = {
            "source": "/root/reference/src/data.py:424-502 firedrake_mesh_to_PyG, executed in place behind a stub mesh",
            "mesh_dims": list(md), "permuted": seed is not None,
            "coords": torch.from_numpy(np.ascontiguousarray(coords)),
            "cells": torch.from_numpy(np.ascontiguousarray(cells).astype(np.int64)),
            "sides": {int(k): torch.from_numpy(np.asarray(v).astype(np.int64)) for k, v in sides.items()},
            "edge_index": d.edge_index.clone(),
            "boundary_nodes": d.boundary_nodes.clone(),
            "corner_nodes": torch.from_numpy(np.asarray(d.corner_nodes).astype(np.int64)),
            "to_boundary_edge_mask": d.to_boundary_edge_mask.clone(),
            "to_corner_nodes_mask": d.to_corner_nodes_mask.clone(),
            "diff_boundary_edges_mask": d.diff_boundary_edges_mask.clone(),
            "x_comp": d.x_comp.clone(),
        }
        torch.save(fx, os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: N={coords.shape[0]} E0={d.edge_index.shape[1]} corners={fx['corner_nodes'].tolist()} "
              f"tb={int(d.to_boundary_edge_mask.sum())} tc={int(d.to_corner_nodes_mask.sum())} "
              f"db={int(d.diff_boundary_edges_mask.sum())}")


if __name__ == "__main__":
    main()
