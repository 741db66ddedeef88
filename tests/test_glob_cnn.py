"""Row f3: the global CNN feature extractor (src/feature_extractors.py:6-34; grid reordering of
src/utils_data.py:125-141).

tests/golden_cnn/*.pt were minted by the reference's OWN class and helper executed in place
(oracle/ref_harness/make_golden_cnn.py).  CPU: the oracle restatement (oracle/cnn_oracle.py) reproduces them,
including the fused gather index of the grid reordering.  GPU: csrc/glob_cnn.cu through
g_adaptivity_b200.feature_extractors.GlobalFeatureExtractorCNN -- same constructor / state_dict -- features within
1e-5 and parameter gradients within 1e-4 of the reference's, from the raw nodal values (fused gather) and from the
reference-shaped [B, 1, H, W] input."""
import glob
import os

import pytest
import torch

from oracle import cnn_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_cnn")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.pt")))


def _load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=True)


def _params(fx):
    L = fx["channels"][2]
    ws = [fx["state_dict"][f"convs.{l}.weight"] for l in range(L)]
    bs = [fx["state_dict"][f"convs.{l}.bias"] for l in range(L)]
    return ws, bs


def test_fixtures_present():
    assert len(NAMES) >= 5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_restatement_reproduces_the_reference(name):
    fx = _load(name)
    B, n, dim = fx["B"], fx["n"], fx["dim"]
    grid = cnn_oracle.reshape_fd_tensor_to_grid(fx["u"], fx["mapping_tensor"], [n, n], B, dim)
    assert torch.equal(grid, fx["grid"])
    if dim == 2:
        idx = cnn_oracle.grid_gather_index(fx["mapping_tensor"], n)
        assert torch.equal(fx["u"].view(B, -1)[:, idx].view(B, n, n), fx["grid"])     # the fused form of the reordering
    ws, bs = _params(fx)
    ws = [w.clone().requires_grad_(True) for w in ws]
    bs = [b.clone().requires_grad_(True) for b in bs]
    feats = cnn_oracle.cnn_features(grid.unsqueeze(1), ws, bs)
    assert torch.equal(feats, fx["features"])
    (feats * fx["cotangent"]).sum().backward()
    for l, (w, b) in enumerate(zip(ws, bs)):
        assert torch.equal(w.grad, fx["grads"][f"convs.{l}.weight"])
        assert torch.equal(b.grad, fx["grads"][f"convs.{l}.bias"])


def test_module_mirrors_the_reference_constructor():
    from g_adaptivity_b200.feature_extractors import GlobalFeatureExtractorCNN
    for name in NAMES:
        fx = _load(name)
        mid, out, L = fx["channels"]
        m = GlobalFeatureExtractorCNN(1, mid, out, dim=fx["dim"], num_layers=L)
        assert sorted(m.state_dict().keys()) == sorted(fx["state_dict"].keys())
        m.load_state_dict(fx["state_dict"], strict=True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 1, 5))                       # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("fused_gather", [False, True])
@pytest.mark.parametrize("name", NAMES)
def test_cuda_cnn_matches_reference_fixture(name, fused_gather):
    from g_adaptivity_b200.feature_extractors import GlobalFeatureExtractorCNN
    fx = _load(name)
    mid, out, L = fx["channels"]
    B, n, dim = fx["B"], fx["n"], fx["dim"]
    m = GlobalFeatureExtractorCNN(1, mid, out, dim=dim, num_layers=L).cuda()
    m.load_state_dict(fx["state_dict"], strict=True)
    if fused_gather:
        idx = (cnn_oracle.grid_gather_index(fx["mapping_tensor"], n) if dim == 2 else torch.arange(n)).int().cuda()
        feats = m(fx["u"].cuda().view(B, -1), gather=idx)
    else:
        feats = m(fx["grid"].unsqueeze(1).cuda())
    sc = fx["features"].abs().max().item()
    assert (feats.detach().cpu() - fx["features"]).abs().max().item() <= 1e-5 * sc
    (feats * fx["cotangent"].cuda()).sum().backward()
    gs = max(g.abs().max().item() for g in fx["grads"].values())
    for k, p in m.named_parameters():
        err = (p.grad.cpu() - fx["grads"][k]).abs().max().item()
        assert err <= 1e-4 * max(fx["grads"][k].abs().max().item(), 1e-2 * gs), (k, err)
    # bit-reproducible run to run (no atomics anywhere)
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    f2 = m(fx["grid"].unsqueeze(1).cuda()) if not fused_gather else m(fx["u"].cuda().view(B, -1), gather=idx)
    (f2 * fx["cotangent"].cuda()).sum().backward()
    assert torch.equal(f2, feats)
    for k, p in m.named_parameters():
        assert torch.equal(p.grad, g1[k])
