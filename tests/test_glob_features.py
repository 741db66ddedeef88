"""Row f3 end to end: global CNN features through `GNN.forward` (src/GNN.py:156-159,172-177,242-268).

tests/golden_glob/*.pt were minted by the reference's OWN `GNN` (with its real `feature_extractors.py` and
`utils_data.py`) on 1-D batches -- the only shape its global-feature branch runs on: for 2-D it reshapes to
[num_nodes, num_nodes] with the NODE count and raises (oracle/gnn_oracle.py: _append_global_features).  Checks:

* CPU: the oracle (`GNNRef`) reproduces the fixtures bit for bit (coordinates, loss, every gradient incl. the CNN's);
* GPU: `g_adaptivity_b200.GNN` -- same state_dict, CNN on csrc/glob_cnn.cu, the global channels folded into a
  per-mesh bias offset of the deformer kernels -- against the fixtures, and in 2-D (intended n x n grids) against
  the oracle evaluated in fp64."""
import copy
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import gad_testutil as util
from g_adaptivity_b200 import synth
from oracle import gnn_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_glob")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.pt")))


def _load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=True)


def _dataset(md):
    ds = synth.SyntheticDataset(len(md), md)
    if ds.x_comp_shared is None:
        ds.x_comp_shared = synth.make_batch(md, 1).x_comp
    return ds


def test_fixtures_present():
    assert len(NAMES) >= 3


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_with_global_features(name):
    fx = _load(name)
    opt, data, md = util.fixture_opt(fx), util.fixture_batch(fx), fx["mesh_dims_list"][0]
    ref = gnn_oracle.GNNRef(_dataset(md), copy.deepcopy(opt))
    assert sorted(ref.state_dict().keys()) == sorted(fx["state_dict"].keys())
    ref.load_state_dict(fx["state_dict"], strict=True)
    out = ref(data)
    assert torch.equal(out, fx["x_phys"])
    loss = gnn_oracle.mesh_loss(out, data.x_phys)
    assert float(loss.item()) == fx["loss"]
    loss.backward()
    for n, p in ref.named_parameters():
        if n in fx["grads"]:
            assert torch.equal(p.grad, fx["grads"][n]), n


def _check_against(model, ref_grads, scale_names_excl=("lin_key.bias",), tol=2e-4):
    got = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    scale = max(g.abs().max().item() for n, g in ref_grads.items() if not any(s in n for s in scale_names_excl))
    for n, g in ref_grads.items():
        if "lin_key.bias" in n or "lin_skip" in n:
            continue
        assert n in got, n
        err = (got[n].double() - g.double()).abs().max().item() / max(g.abs().max().item(), 1e-2 * scale)
        assert err <= tol, (n, err)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_module_matches_reference_fixture_with_global_features(name):
    from test_gpu_parity import cuda_model
    fx = _load(name)
    opt, data, md = util.fixture_opt(fx), util.fixture_batch(fx), fx["mesh_dims_list"][0]
    model = cuda_model(_dataset(md), opt, fx["state_dict"])
    assert model.n_glob_used > 0
    model.train()
    out = model(data)
    assert util.rel_err(out, fx["x_phys"]) <= 1e-5
    tgt = data.x_phys.cuda()
    loss = F.l1_loss(out, tgt if tgt.dim() == 2 else tgt.unsqueeze(-1))
    assert abs(loss.item() - fx["loss"]) <= 1e-5 * abs(fx["loss"])
    loss.backward()
    # the fixture's gradients are fp32 autograd: compare against the oracle in fp64 where they are noisy
    ref = gnn_oracle.GNNRef(_dataset(md), copy.deepcopy(opt))
    ref.load_state_dict(fx["state_dict"])
    ref = ref.double()
    d64 = data.clone()
    for k in d64.keys():
        v = getattr(d64, k)
        if torch.is_tensor(v) and v.dtype == torch.float32:
            setattr(d64, k, v.double())
    gnn_oracle.mesh_loss(ref(d64), d64.x_phys).backward()
    _check_against(model, {n: p.grad for n, p in ref.named_parameters() if p.grad is not None})
    with torch.no_grad():
        assert util.rel_err(model(data), fx["x_phys"]) <= 1e-5        # inference path (no autograd node)


@pytest.mark.gpu
@pytest.mark.parametrize("md,B,over", [
    ((12, 12), 5, {"gnn_inc_glob_feat_f": True, "gnn_inc_glob_feat_uu": True}),
    ((30, 30), 3, {"gnn_inc_glob_feat_f": True, "gnn_inc_glob_feat_uu": True, "hidden_dim": 16}),
    ((9, 9), 4, {"gnn_inc_glob_feat_uu": True, "share_conv": False, "num_layers": 3}),
    ((10, 10), 3, {"gnn_inc_glob_feat_f": True, "ode_method": "rk4", "num_layers": 2}),
    ((8, 8), 2, {"gnn_inc_glob_feat_f": True, "hidden_dim": 4}),          # nothing global survives the truncation
])
def test_cuda_module_matches_oracle_with_global_features_2d(md, B, over):
    """2-D meshes, the intended n x n grid per mesh, against the oracle in fp64: coordinates 1e-5, every gradient
    (Linear layers AND both CNNs) 2e-4."""
    from test_gpu_parity import cuda_model
    opt = synth.default_opt(md, **over)
    ds = _dataset(md)
    data = synth.make_batch(md, B, seed=9)
    torch.manual_seed(42)
    ref32 = gnn_oracle.GNNRef(ds, copy.deepcopy(opt))
    model = cuda_model(ds, opt, ref32.state_dict())
    model.train()
    out = model(data)
    ref = gnn_oracle.GNNRef(ds, copy.deepcopy(opt))
    ref.load_state_dict(ref32.state_dict())
    ref = ref.double()
    d64 = data.clone()
    for k in d64.keys():
        v = getattr(d64, k)
        if torch.is_tensor(v) and v.dtype == torch.float32:
            setattr(d64, k, v.double())
    ref_out = ref(d64)
    assert util.rel_err(out, ref_out) <= 1e-5
    F.l1_loss(out, data.x_phys.cuda()).backward()
    gnn_oracle.mesh_loss(ref_out, d64.x_phys).backward()
    grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    if model.n_glob_used == 0:
        assert all(p.grad is None or p.grad.abs().max().item() == 0.0
                   for n, p in model.named_parameters() if "global_feature" in n)
        grads = {n: g for n, g in grads.items() if "global_feature" not in n}
    _check_against(model, grads)
    assert model.n_glob_used == 0 or model.last_graph.T == B                   # one mesh per tile
