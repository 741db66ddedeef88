"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import g_adaptivity_b200 as gad
from g_adaptivity_b200 import _lib, build as gad_build, graph, synth
from g_adaptivity_b200 import functional as GF
from g_adaptivity_b200 import params as gparams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    gad_build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gadapt.h")).read()
    declared = set(re.findall(r"\b(gad_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in gadapt.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.gad_version() >= 100


def test_workspace_queries_need_no_gpu(lib):
    assert lib.gad_graph_workspace_bytes(1000, 4, 225, 0) > 0
    assert lib.gad_graph_workspace_bytes(1000, 4, 225, 1) > lib.gad_graph_workspace_bytes(1000, 4, 225, 0)
    a = lib.gad_deform_workspace_bytes(1000, 4, 0)
    b = lib.gad_deform_workspace_bytes(1000, 4, 1)
    assert b > a >= 2 * 1000 * 4 * 4


def test_argument_errors_are_reported_not_crashed(lib):
    rc = lib.gad_prepare_weights(None, None, None, 1, 8, 4, 1.0, None, None)
    assert rc == 1 and b"null" in lib.gad_last_error()
    rc = lib.gad_deform_fwd(None, None, 10, 10, None, 0, 0, 0, None, 2, 3, None, 1, None, 4, 0, None, None, None, 0, None)
    assert rc != 0


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    opt = synth.default_opt((6, 6), device="cuda")
    model = gad.GNN(synth.SyntheticDataset(2, (6, 6)), opt)
    with pytest.raises((RuntimeError, AssertionError)):
        model(synth.make_batch((6, 6), 1))
    opt = synth.default_opt((6, 6), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        gad.GNN(synth.SyntheticDataset(2, (6, 6)), opt)(synth.make_batch((6, 6), 1))
    with pytest.raises(RuntimeError):
        GF.pack_features(torch.zeros(4, 2), None, None, None, None, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "g_adaptivity_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_plan_tiles():
    tp = graph.plan_tiles([900] * 5, target_nodes=1024)
    assert tp.tolist() == [0, 900, 1800, 2700, 3600, 4500]
    tp = graph.plan_tiles([200] * 11, target_nodes=1024)
    assert tp.tolist() == [0, 1000, 2000, 2200]
    tp = graph.plan_tiles([25, 2500, 49, 36], target_nodes=1024)
    assert tp.tolist() == [0, 25, 2525, 2610]
    assert graph.plan_tiles([40000], target_nodes=1024) is None        # single 200x200 mesh -> streaming
    assert graph.plan_tiles([], target_nodes=1024).tolist() == [0]
    # the shared-memory model mirrors csrc/fused_kernels.cu (50x50, CE=4 fits one CTA on B200)
    assert graph._bwd_smem_bytes(2500, 14212, 4) < 227 * 1024
    assert graph._bwd_smem_bytes(3072, 18000, 4) > 227 * 1024


def test_live_channels():
    assert GF.live_channels(4, 8) == (4, 4)
    assert GF.live_channels(2, 8) == (2, 2)
    assert GF.live_channels(3, 8) == (3, 4)
    assert GF.live_channels(4, 2) == (2, 2)
    assert GF.live_channels(8, 8) == (8, 8)
    with pytest.raises(NotImplementedError):
        GF.live_channels(20, 16)


def test_model_surface_matches_reference_contract():
    opt = synth.default_opt((15, 15), device="cuda", learn_step=True)
    m = gad.GNN(synth.SyntheticDataset(2, (15, 15)), opt)
    keys = set(m.state_dict())
    assert "enc.weight" in keys and {f"steps.{i}" for i in range(4)} <= keys
    for i in range(4):
        for leaf in ("lin_key.weight", "lin_key.bias", "lin_query.weight", "lin_query.bias", "lin_skip.weight"):
            assert f"conv_layers.{i}.{leaf}" in keys
    assert m.conv_layers[0] is m.conv_layers[3]          # share_conv: one instance (GNN.py:131-137)
    assert not m.enc.weight.requires_grad
    assert opt["hidden_dims_list"] == [2, 1, 1]           # written back (GNN.py:161)
    assert torch.equal(m.enc.weight, torch.eye(8, 4))
    # a reference state_dict (golden fixture) loads strictly
    import gad_testutil as util
    fx = util.load_golden("learn_step_6x6")
    m.load_state_dict(fx["state_dict"], strict=True)
    m2 = gad.GNN(synth.SyntheticDataset(2, (6, 6)), synth.default_opt((6, 6), share_conv=False, num_layers=3))
    assert m2.conv_layers[0] is not m2.conv_layers[1]


@pytest.mark.parametrize("over,exc", [
    ({"conv_type": "GCN"}, NotImplementedError), ({"conv_type": "GAT_plus"}, NotImplementedError),
    ({"enc": "lin_layer"}, NotImplementedError), ({"dropout": 0.5}, NotImplementedError),
    ({"loss_type": "pde_loss", "data_type": "randg_mix"}, NotImplementedError),
    ({"reg_skew": True}, NotImplementedError), ({"softmax_temp_type": "learnable_v"}, NotImplementedError),
    ({"residual": False}, NotImplementedError), ({"ode_method": "dopri5"}, ValueError),
])
def test_unsupported_options_raise(over, exc):
    opt = synth.default_opt((6, 6), **over)
    with pytest.raises(exc):
        gad.GNN(synth.SyntheticDataset(2, (6, 6)), opt)


def test_global_feature_models_construct_like_the_reference():
    """gnn_inc_glob_feat_* (on by CLI default, params.py:259-260): the two CNNs exist under the reference's names,
    in_dims grows by global_feat_dim each, the state rows stay as wide as the node-varying inputs."""
    opt = synth.default_opt((6, 6), gnn_inc_glob_feat_f=True, gnn_inc_glob_feat_uu=True)
    m = gad.GNN(synth.SyntheticDataset(2, (6, 6)), opt)
    assert m.in_dims == [2, 1, 1, 8, 8] and opt["hidden_dims_list"] == [2, 1, 1, 8, 8]
    keys = set(m.state_dict().keys())
    for which in ("f", "uu"):
        for i in range(4):
            assert f"global_feature_extractor_cnn_{which}.convs.{i}.weight" in keys
    assert m.enc.weight.shape == (8, 20) and m.CE == 4 and m.n_glob_used == 4
    assert m.global_feature_extractor_cnn_f.convs[0].weight.shape == (8, 1, 3, 3)
    m1 = gad.GNN(synth.SyntheticDataset(1, (9,)), synth.default_opt((9,), gnn_inc_glob_feat_uu=True))
    assert m1.global_feature_extractor_cnn_uu.convs[1].weight.shape == (8, 8, 3) and m1.n_glob_used == 5


def test_params_mirror_defaults_and_presets():
    opt = gparams.get_params([])
    assert opt["hidden_dim"] == 8 and opt["num_layers"] == 4 and opt["time_step"] == 0.1
    assert opt["fix_boundary"] == "True" and opt["softmax_temp_type"] is None and opt["mesh_dims"] == [10, 10]
    opt = gparams.tf_sweep_args(opt)
    assert opt["fix_boundary"] is True and opt["self_loops"] is False and opt["show_mesh_evol_plots"] is True
    opt = gparams.run_params(opt)
    assert opt["conv_type"] == "GRAND_plus" and opt["mesh_dims"] == [11, 11] and opt["gnn_inc_feat_uu"] is True
    assert opt["gnn_inc_glob_feat_f"] is False and opt["share_conv"] is True and opt["enc"] == "identity"
    b = gparams.run_params(gparams.tf_sweep_args(gparams.get_params(["--pde_type", "Burgers"])))
    assert b["conv_type"] == "GRAND" and b["mesh_dims"] == [21] and b["gnn_inc_feat_f"] is False
    assert b["loss_type"] == "modular" and b["num_eval_time_steps"] == 20
    assert gparams.get_arg_list([15, 15]) == [15, 15] and gparams.get_arg_list(["[20, 20]"]) == [20, 20]
    o = gparams.get_params(["--mesh_dims", "30", "30", "--ode_method", "rk4"])
    assert o["ode_method"] == "rk4"


def test_corner_loops_follow_reference_order():
    data = synth.make_batch((5, 5), 3)
    ids = graph.corner_loops(data, 2, [5, 5], [25, 25, 25]).tolist()
    assert ids == [0, 4, 20, 24, 25, 29, 45, 49, 50, 54, 70, 74]
    d1 = synth.make_batch((7,), 2)
    assert graph.corner_loops(d1, 1, [7], [7, 7]).tolist() == [0, 6, 7, 13]


def test_shared_topology_sample_positions_stay_inside_huge_edge_lists():
    """cfg 5 on one GPU has 1.2e8 edges: float32 positions would round E - 1 up to E (found by the bench)."""
    for E in (1, 2, 63, 64, 65, 1000, 119_619_584, 2 ** 31 - 1):
        idx = graph.GraphCache.sample_positions(E)
        assert idx.dtype == torch.int64 and int(idx.min()) == 0 and int(idx.max()) == E - 1
        assert bool((idx[1:] >= idx[:-1]).all()) and idx.numel() == min(E, 64)


def test_device_batch_builder_equals_host_collation():
    """synth.make_batch_device (used by the bench for the 8192-mesh cfg 5 batch) tiles `block` host samples and
    builds the topology tensors with torch ops: identical to make_batch when block covers the batch."""
    a = synth.make_batch((6, 6), 9, seed=3)
    b = synth.make_batch_device((6, 6), 9, "cpu", seed=3, block=9)
    for k in a.keys():
        va, vb = getattr(a, k), getattr(b, k)
        if torch.is_tensor(va):
            assert torch.equal(va, vb), k
    c = synth.make_batch_device((6, 6), 9, "cpu", seed=3, block=4)
    assert torch.equal(c.edge_index, a.edge_index) and torch.equal(c.batch, a.batch)
    assert torch.equal(c.f_tensor[:4 * 36], a.f_tensor[:4 * 36]) and torch.equal(c.f_tensor[4 * 36:8 * 36], a.f_tensor[:4 * 36])
    assert c.mesh_sizes == a.mesh_sizes and len(c.corner_nodes) == 9


def test_host_path_relay_plan():
    """dp.plan_host_relays: slowest paths paired with fastest, fraction (R_fast - R_slow) / (R_fast + R_slow); the
    rates are the ones measured on an 8-GPU box (profiles/r02_h2d_probe_8gpu.json)."""
    from g_adaptivity_b200 import dp
    rates = [20.69, 20.93, 20.8, 20.93, 36.11, 36.45, 36.23, 36.21]
    plan = dp.plan_host_relays(rates, list(range(8)))
    assert sorted(plan) == [0, 1, 2, 3]                               # only the slow class relays
    assert sorted(v[0] for v in plan.values()) == [4, 5, 6, 7]        # each through a different fast GPU
    for q, (via, y) in plan.items():
        assert abs(y - (rates[via] - rates[q]) / (rates[via] + rates[q])) < 1e-12 and 0.25 < y < 0.29
        # both links finish together under the model
        assert abs((1 - y) / rates[q] - (1 + y) / rates[via]) < 1e-12
    assert dp.plan_host_relays([50.0, 51.0], [0, 1]) == {}            # equal paths: nothing to balance
    assert dp.plan_host_relays([20.0, 40.0, 30.0], [0, 1, 2]) == {0: (1, (40.0 - 20.0) / 60.0)}
