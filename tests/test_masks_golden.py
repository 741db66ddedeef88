"""Row a1 pinned: graph + edge masks of `firedrake_mesh_to_PyG` (src/data.py:424-502).

tests/golden_masks/*.pt were minted by the reference's OWN function executed in place behind a stub Firedrake
mesh (oracle/ref_harness/make_golden_masks.py).  Compared bit for bit:
  * CPU: `synth._edges_from_cells` (edge list in the reference's `list(set)` order), `synth._masks_from_sides`
    and, on the canonical structured meshes, `synth.MeshTopology` -- the generator of every other fixture;
  * GPU: `gad_edge_masks` (csrc/graph_build.cu) on the reference's edge list.
Includes renumbered meshes (Firedrake numbers nodes in DMPlex order, not row-major), the degenerate 2x2 /
2-node meshes and the 1-D case, where the reference's corner set is empty."""
import glob
import os

import numpy as np
import pytest
import torch

from g_adaptivity_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_masks")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.pt")))


def _load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=True)


def test_fixtures_present():
    assert len(NAMES) >= 10


@pytest.mark.parametrize("name", NAMES)
def test_synth_restatement_equals_reference_function(name):
    fx = _load(name)
    cells = fx["cells"].numpy()
    N = fx["coords"].shape[0]
    sides = [fx["sides"][k].numpy() for k in sorted(fx["sides"])]
    ei = synth._edges_from_cells(cells)
    assert np.array_equal(ei, fx["edge_index"].numpy())                       # same edges, same order
    on_b, corners, tb, tc, db = synth._masks_from_sides(ei, sides, N)
    assert np.array_equal(on_b, fx["boundary_nodes"].numpy())
    assert np.array_equal(corners, fx["corner_nodes"].numpy())
    assert np.array_equal(tb, fx["to_boundary_edge_mask"].numpy())
    assert np.array_equal(tc, fx["to_corner_nodes_mask"].numpy())
    assert np.array_equal(db, fx["diff_boundary_edges_mask"].numpy())
    if not fx["permuted"]:
        topo = synth.MeshTopology(fx["mesh_dims"])
        assert np.array_equal(topo.edge_index, fx["edge_index"].numpy())
        assert np.array_equal(topo.corner_nodes, fx["corner_nodes"].numpy())
        assert np.array_equal(topo.boundary_nodes, fx["boundary_nodes"].numpy())
        assert np.array_equal(topo.to_boundary_edge_mask, fx["to_boundary_edge_mask"].numpy())
        assert np.array_equal(topo.to_corner_nodes_mask, fx["to_corner_nodes_mask"].numpy())
        assert np.array_equal(topo.diff_boundary_edges_mask, fx["diff_boundary_edges_mask"].numpy())
        assert np.array_equal(topo.coords.reshape(N, -1), fx["x_comp"].numpy().reshape(N, -1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_edge_masks_equal_reference_function(name):
    from g_adaptivity_b200 import graph as G
    fx = _load(name)
    N = fx["coords"].shape[0]
    side_bits = np.zeros(N, dtype=np.uint8)
    for k in sorted(fx["sides"]):
        side_bits[fx["sides"][k].numpy()] |= np.uint8(1 << (k - 1))
    tb, tc, db = G.edge_masks(fx["edge_index"].cuda(), torch.from_numpy(side_bits))
    assert torch.equal(tb.cpu(), fx["to_boundary_edge_mask"])
    assert torch.equal(tc.cpu(), fx["to_corner_nodes_mask"])
    assert torch.equal(db.cpu(), fx["diff_boundary_edges_mask"])
