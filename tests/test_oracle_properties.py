"""CPU: oracle self-consistency (dense formulation) and the invariants the reference has by
construction."""
import numpy as np
import pytest
import torch

from g_adaptivity_b200 import synth
from oracle import gnn_oracle

import gad_testutil as util


def _model(opt, dim, mesh_dims, seed=42, dtype=torch.float32, wscale=1.0):
    torch.manual_seed(seed)
    m = gnn_oracle.GNNRef(synth.SyntheticDataset(dim, mesh_dims), opt)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lin_key" in n or "lin_query" in n:
                p.mul_(wscale)
    return m.to(dtype)


def _to(data, dtype):
    for k in ("x_comp", "f_tensor", "uu_tensor", "u_true_tensor", "x_phys"):
        setattr(data, k, getattr(data, k).to(dtype))
    return data


@pytest.mark.parametrize("mesh_dims,B", [((15, 15), 1), ((7, 7), 3), ((21,), 4)])
def test_sparse_equals_dense_fp64(mesh_dims, B):
    dim = len(mesh_dims)
    opt = synth.default_opt(mesh_dims)
    data = _to(synth.make_batch(mesh_dims, B, seed=3), torch.float64)
    model = _model(opt, dim, mesh_dims, dtype=torch.float64)
    out, states, ei = model(data, return_states=True)
    conv = model.conv_layers[0]
    x = states[0]
    for l in range(opt["num_layers"]):
        res, A = gnn_oracle.dense_layer(x, ei, conv.lin_query.weight, conv.lin_query.bias,
                                        conv.lin_key.weight, conv.lin_key.bias)
        x = x + opt["time_step"] * res
        assert (x - states[l + 1]).abs().max().item() < 1e-12
        # rows of alpha sum to one (GRAND_plus.py:333)
        assert (A.sum(1) - 1).abs().max().item() < 1e-12
    assert (x[:, :dim] - out).abs().max().item() < 1e-12


def test_fixed_points_boundary_sides_and_dead_channels():
    md = (9, 9)
    opt = synth.default_opt(md)
    data = synth.make_batch(md, 2, seed=5)
    model = _model(opt, 2, md)
    out, states, ei = model(data, return_states=True)
    n = md[0]
    x0 = data.x_comp
    # corners: single self-loop -> res = 0 (GNN.py:209-218)
    for b in range(2):
        for c in data.corner_nodes[b]:
            assert torch.equal(out[b * n * n + int(c)], x0[b * n * n + int(c)])
    # a node on side s stays on side s (it aggregates same-side nodes only, data.py:465-494)
    xs = x0.view(2, n, n, 2)
    os_ = out.view(2, n, n, 2)
    assert (os_[:, :, 0, 0] - 0.0).abs().max() < 1e-6 and (os_[:, :, -1, 0] - 1.0).abs().max() < 1e-6
    assert (os_[:, 0, :, 1] - 0.0).abs().max() < 1e-6 and (os_[:, -1, :, 1] - 1.0).abs().max() < 1e-6
    # channels >= in_dim stay exactly zero with the identity encoder (GNN.py:75-83)
    for s in states:
        assert torch.count_nonzero(s[:, 4:]) == 0
    # convexity: each layer output is inside the bounding box of {self} U in-neighbours (0<=tau<=1)
    for l in range(opt["num_layers"]):
        x, xn = states[l][:, :2], states[l + 1][:, :2]
        lo = x.clone()
        hi = x.clone()
        lo = lo.scatter_reduce(0, ei[1][:, None].expand(-1, 2), x[ei[0]], "amin", include_self=True)
        hi = hi.scatter_reduce(0, ei[1][:, None].expand(-1, 2), x[ei[0]], "amax", include_self=True)
        assert (xn >= lo - 1e-6).all() and (xn <= hi + 1e-6).all()


def test_key_bias_gradient_is_zero_and_skip_has_none():
    md = (8, 8)
    opt = synth.default_opt(md)
    data = _to(synth.make_batch(md, 2, seed=1), torch.float64)
    model = _model(opt, 2, md, dtype=torch.float64)
    out = model(data)
    gnn_oracle.mesh_loss(out, data.x_phys).backward()
    conv = model.conv_layers[0]
    assert conv.lin_key.bias.grad.abs().max().item() < 1e-14   # softmax shift invariance
    assert conv.lin_skip.weight.grad is None
    assert conv.lin_query.weight.grad.abs().max().item() > 0


def test_edge_permutation_and_disjoint_union():
    md = (7, 7)
    opt = synth.default_opt(md)
    model = _model(opt, 2, md)
    batch = synth.make_batch(md, 3, seed=11)
    out = model(batch)
    # disjoint union == per-mesh results
    topo = synth.MeshTopology(md)
    for b in range(3):
        single = synth.Batch.from_data_list([synth.make_data(topo, 11 + b)])
        ob = model(single)
        assert torch.allclose(ob, out[b * 49:(b + 1) * 49], rtol=0, atol=1e-6)
    # permuting the raw edge list changes only the summation order
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(batch.edge_index.shape[1], generator=g)
    pb = batch.clone()
    pb.edge_index = batch.edge_index[:, perm]
    for k in ("to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask"):
        setattr(pb, k, getattr(batch, k)[perm])
    assert util.rel_err(model(pb), out) < 1e-6


def test_make_batch_matches_per_sample_collation():
    md = (6, 6)
    topo = synth.MeshTopology(md)
    a = synth.make_batch(md, 3, seed=7)
    b = synth.Batch.from_data_list([synth.make_data(topo, 7 + i) for i in range(3)])
    for k in ("edge_index", "batch", "x_comp", "x_phys", "f_tensor", "uu_tensor", "to_boundary_edge_mask",
              "to_corner_nodes_mask", "diff_boundary_edges_mask"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("n", [5, 15, 30])
def test_edge_counts_match_survey_formulas(n):
    topo = synth.MeshTopology((n, n))
    assert topo.edge_index.shape[1] == 6 * n * n - 8 * n + 2
    opt = synth.default_opt((n, n))
    ei = gnn_oracle.filtered_edge_index(synth.make_batch((n, n), 1), opt, 2)
    assert ei.shape[1] == 6 * (n - 2) ** 2 + 8 * (n - 2) + 4
    # corner loops are appended last (GNN.py:217-218)
    assert torch.equal(ei[0, -4:], ei[1, -4:])
    t1 = synth.MeshTopology((n,))
    ei1 = gnn_oracle.filtered_edge_index(synth.make_batch((n,), 2), synth.default_opt((n,)), 1)
    assert ei1.shape[1] == 2 * 2 * (n - 1)


def test_csr_oracle_is_scatter_order():
    md = (6, 6)
    data = synth.make_batch(md, 2, seed=2)
    ei = gnn_oracle.filtered_edge_index(data, synth.default_opt(md), 2)
    N = data.x_comp.shape[0]
    rowptr, col, eid = gnn_oracle.csr_by_destination(ei, N)
    assert rowptr[-1].item() == ei.shape[1]
    # scatter_add_ on CPU accumulates rows in edge-list order == CSR slot order
    vals = torch.randn(ei.shape[1], dtype=torch.float32)
    ref = torch.zeros(N).scatter_add_(0, ei[1], vals)
    mine = torch.zeros(N)
    for i in range(N):
        acc = torch.zeros((), dtype=torch.float32)
        for s in range(rowptr[i], rowptr[i + 1]):
            acc = acc + vals[eid[s]]
        mine[i] = acc
    assert torch.equal(ref, mine)


def test_rk4_extension_reduces_to_reference_form():
    md = (6, 6)
    opt = synth.default_opt(md, ode_method="rk4", num_layers=3)
    data = _to(synth.make_batch(md, 1, seed=4), torch.float64)
    model = _model(opt, 2, md, dtype=torch.float64)
    out, states, ei = model(data, return_states=True)
    conv = model.conv_layers[0]
    F_ = lambda y: gnn_oracle.dense_layer(y, ei, conv.lin_query.weight, conv.lin_query.bias,
                                          conv.lin_key.weight, conv.lin_key.bias)[0]
    x, h = states[0], opt["time_step"]
    for _ in range(3):
        k1 = F_(x); k2 = F_(x + h * k1 / 2); k3 = F_(x + h * k2 / 2); k4 = F_(x + h * k3)
        x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert (x[:, :2] - out).abs().max().item() < 1e-12
