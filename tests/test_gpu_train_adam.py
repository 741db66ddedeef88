"""Single-GPU parity of the OPTIMIZER half of the benchmarked training step.

The benchmarked kernel (`k_ell_train`, tail = 2) ends with Adam + the refold of (M, u).  The reference
trains with `torch.optim.Adam(model.parameters(), lr=opt['lr'], weight_decay=opt['decay'])` and
`loss.backward(); optimizer.step()` (src/run_GNN.py:88,128-131).  Two checks, for every way the
trainer can issue a step (eager launches, one CUDA graph per step, one graph per pass over the ring):

* K steps on cfg-2-shaped data (30x30 meshes) against the oracle (`GNNRef`) trained by
  `torch.optim.Adam` on the same batches: losses and parameters;
* the Adam arithmetic in isolation: `torch.optim.Adam` is fed the gradient the kernel itself reduced
  at every step (`tr.gflat`), so gradient rounding does not enter and the parameters must agree to
  fp32 rounding (1e-6 of the largest weight) after K steps, bias corrections and weight decay included.
"""
import copy

import pytest
import torch

import gad_testutil as util
from g_adaptivity_b200 import synth
from g_adaptivity_b200.trainer import DeformerTrainer
from oracle import gnn_oracle
from test_gpu_parity import cuda_model, oracle_model

pytestmark = pytest.mark.gpu

MD, B, K, LR = (30, 30), 64, 10, 1e-2


def _flat_views(model):
    c = model.conv_layers[0]
    return [c.lin_query.weight, c.lin_query.bias, c.lin_key.weight, c.lin_key.bias]


def _oracle_run(ds, opt, batches, state, wd=0.0):
    ref = oracle_model(ds, opt, state)
    optim = torch.optim.Adam(ref.parameters(), lr=LR, weight_decay=wd)
    losses = []
    for k in range(K):
        data = batches[k % len(batches)]
        optim.zero_grad(set_to_none=True)
        loss = gnn_oracle.mesh_loss(ref(data), data.x_phys)
        loss.backward()
        optim.step()
        losses.append(float(loss.item()))
    return torch.cat([v.detach().flatten() for v in _flat_views(ref)]), losses


@pytest.fixture(scope="module")
def case():
    opt = synth.default_opt(MD, lr=LR)
    ds = synth.SyntheticDataset(2, MD)
    batches = [synth.make_batch(MD, B, seed=31, first_mesh_id=r * B) for r in range(2)]
    torch.manual_seed(42)
    ref0 = oracle_model(ds, opt)
    state = copy.deepcopy(ref0.state_dict())
    want, losses = _oracle_run(ds, opt, batches, state)
    return opt, ds, batches, state, want, losses


@pytest.mark.parametrize("mode", ["eager", "graph", "epoch"])
def test_k_steps_match_oracle_trained_by_torch_adam(case, mode):
    opt, ds, batches, state, want, ref_losses = case
    model = cuda_model(ds, opt, state, gad_store_alpha=False)
    tr = DeformerTrainer(model, use_cuda_graph=(mode != "eager"))
    assert tr.lr == LR
    sids = [tr.add_batch(b) for b in batches]
    assert tr._one_launch(tr.slots[0])          # the benchmarked kernel, tail = 2
    losses = []
    if mode == "epoch":
        for _ in range(K // 2):
            for l in tr.run_epoch(sids):
                pass
            tr.synchronize()
            losses += [float(tr.slots[s].loss.item()) for s in sids]
    else:
        for k in range(K):
            l = tr.step(sids[k % 2])
            tr.synchronize()
            losses.append(float(l.item()))
    assert int(tr.step_count.item()) == K
    got = tr.flat.detach().cpu()
    n = got.numel()
    live = torch.ones(n, dtype=torch.bool)
    live[n - _flat_views(model)[3].numel():] = False        # lin_key.bias: gradient analytically zero
    # the loss of step k depends on the parameters after k - 1 optimizer steps
    # (first step: same parameters on both sides, kernel-vs-oracle rounding only; later steps carry the
    # O(lr)-amplified rounding noise of Adam's normalised update, same bar as the parameters below)
    assert abs(losses[0] - ref_losses[0]) <= 1e-5 * abs(ref_losses[0])
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-3 * abs(b), (losses, ref_losses)
    assert losses[-1] < 0.7 * losses[0]                       # and it trains
    # Adam normalises each entry's gradient: rounding noise on near-zero entries is amplified to O(lr)
    err = (got - want)[live].abs().max().item() / want[live].abs().max().item()
    assert err <= 2e-3, err
    # d loss / d lin_key.bias is analytically zero (softmax shift invariance): the kernel writes exact zeros, so
    # Adam leaves the bias alone, while autograd's rounding noise there makes torch Adam random-walk it by O(lr)
    # per step -- without any effect on the model's output
    assert torch.equal(got[~live], state["conv_layers.0.lin_key.bias"].flatten())
    # the model's parameters are views of the trainer's flat vector
    assert torch.equal(model.state_dict()["conv_layers.0.lin_query.weight"].cpu().flatten(), got[:64])


@pytest.mark.parametrize("mode,wd,over", [
    ("eager", 0.0, {}),
    ("graph", 0.0, {}),
    ("epoch", 0.0, {}),
    ("graph", 1e-2, {}),                                      # weight decay (opt['decay'], run_GNN.py:88)
    ("graph", 0.0, {"learn_step": True, "share_conv": False}),
    ("eager", 0.0, {"gad_no_fused_train": True}),             # stand-alone gad_adam_step kernel
])
def test_adam_arithmetic_equals_torch_optim_adam_on_the_same_gradients(mode, wd, over):
    opt = synth.default_opt(MD, lr=LR, decay=wd, **over)
    ds = synth.SyntheticDataset(2, MD)
    batches = [synth.make_batch(MD, 16, seed=32, first_mesh_id=r * 16) for r in range(2)]
    torch.manual_seed(42)
    state = oracle_model(ds, opt).state_dict()
    model = cuda_model(ds, opt, state, gad_store_alpha=False)
    tr = DeformerTrainer(model, use_cuda_graph=(mode != "eager"))
    assert tr.wd == wd
    sids = [tr.add_batch(b) for b in batches]
    shadow = torch.nn.Parameter(tr.flat.detach().cpu().clone())
    optim = torch.optim.Adam([shadow], lr=LR, weight_decay=wd)
    scale = shadow.detach().abs().max().item()

    def follow():
        tr.synchronize()
        shadow.grad = tr.gflat.detach().cpu().clone()
        optim.step()
        err = (tr.flat.detach().cpu() - shadow.detach()).abs().max().item() / scale
        assert err <= 1e-6, err

    if mode == "epoch":
        # a pass over the ring is ONE graph: follow pass by pass with two-slot rings of length 1 and 2
        for k in range(K):
            tr.run_epoch([sids[k % 2]])
            follow()
    else:
        for k in range(K):
            tr.step(sids[k % 2])
            follow()
    assert int(tr.step_count.item()) == K
