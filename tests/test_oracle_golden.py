"""CPU: the oracle restatement reproduces what the reference's own source produced
(fixtures minted by oracle/ref_harness/make_golden.py), bit for bit where the op order is the same."""
import pytest
import torch
import torch.nn.functional as F

from oracle import gnn_oracle
import gad_testutil as util

NAMES = util.golden_names()


def test_goldens_present():
    assert len(NAMES) >= 12


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_run(name):
    fx = util.load_golden(name)
    opt = util.fixture_opt(fx)
    data = util.fixture_batch(fx)
    model = gnn_oracle.GNNRef(util.fixture_dataset(fx), opt)
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=True)
    model.train()
    out, states, ei = model(data, return_states=True)
    # graph prologue: integer work, bit-exact
    assert torch.equal(ei, fx["edge_index_filtered"])
    # forward: same torch ops in the same order -> identical bits
    assert torch.equal(out, fx["x_phys"]), util.rel_err(out, fx["x_phys"])
    alpha = model.conv_layers[-1].stored_alpha
    assert torch.equal(alpha, fx["alpha_last"])
    target = data.x_phys if data.x_phys.dim() == 2 else data.x_phys.unsqueeze(-1)
    loss = F.l1_loss(out, target)
    assert abs(loss.item() - fx["loss"]) <= 1e-7 * max(1.0, abs(fx["loss"]))
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert sorted(grads) == sorted(fx["grads"])
    for n, g in fx["grads"].items():
        scale = g.abs().max().item()
        assert (grads[n] - g).abs().max().item() <= 1e-6 * max(scale, 1e-6), n
    assert sorted(fx["params_without_grad"]) == sorted(
        n for n, p in model.named_parameters() if p.requires_grad and p.grad is None)


def test_state_dict_keys_match_reference():
    fx = util.load_golden("cfg1_15x15")
    keys = set(fx["state_dict"])
    assert "enc.weight" in keys
    for i in range(4):
        for leaf in ("lin_key.weight", "lin_key.bias", "lin_query.weight", "lin_query.bias", "lin_skip.weight"):
            assert f"conv_layers.{i}.{leaf}" in keys
    fx = util.load_golden("learn_step_6x6")
    assert {f"steps.{i}" for i in range(4)} <= set(fx["state_dict"])
