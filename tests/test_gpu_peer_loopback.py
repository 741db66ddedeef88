"""The in-kernel gradient all-reduce (csrc/ell_kernels.cuh: peer_allreduce) at 2 / 4 / 8 ranks on ONE GPU.

`opt['gad_peer_loopback'] = (world, rank)` gives the trainer `world` receive buffers on this device
(dp.PeerExchange.loopback); the training kernel runs exactly the code it runs on `world` GPUs and the test
plays the other ranks by depositing their tagged words.  Checked:

* the reduced gradient is the RANK-ORDERED fp32 sum of the contributions, bit for bit, for this rank at
  any position in the order (the replicas must compute identical bits);
* both sequence parities, consecutive launches, eager and CUDA-graph replay;
* Adam of the same launch consumed the reduced gradient (torch.optim.Adam on it gives the parameters);
* the wait is BOUNDED: a peer that never sends makes the kernel give up after `gad_peer_timeout_ms`, raise
  the device error word and skip the Adam step; `check_peer` reports it; later launches do not wait again.
"""
import time

import pytest
import torch

from g_adaptivity_b200 import synth
from g_adaptivity_b200.trainer import DeformerTrainer
from test_gpu_parity import cuda_model, oracle_model

pytestmark = pytest.mark.gpu

MD, B = (12, 12), 6


def _trainer(world, rank, use_graph=False, timeout_ms=10000, **over):
    opt = synth.default_opt(MD, lr=1e-2, **over)
    ds = synth.SyntheticDataset(2, MD)
    torch.manual_seed(42)
    state = oracle_model(ds, opt).state_dict()
    model = cuda_model(ds, opt, state, gad_store_alpha=False, gad_peer_loopback=(world, rank),
                       gad_peer_timeout_ms=timeout_ms)
    tr = DeformerTrainer(model, use_cuda_graph=use_graph)
    sid = tr.add_batch(synth.make_batch(MD, B, seed=3))
    assert tr.fused_dp and tr.slots[sid].peer_route and tr._one_launch(tr.slots[sid])
    return tr, sid


def _local_gradient(tr, sid):
    """This rank's own contribution: the gradient of a step that stops before the exchange (tail = 1)."""
    with torch.cuda.stream(tr.stream):
        tr._issue(tr.slots[sid], tr.stream.cuda_stream, with_optimizer=False)
    tr.stream.synchronize()
    return tr.gflat.detach().clone()


def _ordered_sum(parts):
    acc = torch.zeros_like(parts[0])
    for p in parts:
        acc = acc + p          # fp32, rank order: ((0 + g_0) + g_1) + ...
    return acc


@pytest.mark.parametrize("world,rank", [(2, 0), (2, 1), (4, 2), (8, 0), (8, 5), (16, 15)])
def test_reduced_gradient_is_the_rank_ordered_sum_and_adam_consumes_it(world, rank):
    tr, sid = _trainer(world, rank)
    shadow = torch.nn.Parameter(tr.flat.detach().cpu().clone())
    optim = torch.optim.Adam([shadow], lr=1e-2)
    gen = torch.Generator().manual_seed(100 + world * 16 + rank)
    for seq in (1, 2, 3):                                   # both parities, and parity 1 reused
        g_own = _local_gradient(tr, sid)                    # depends on the current parameters
        parts = []
        for r in range(world):
            if r == rank:
                parts.append(g_own)
                continue
            g_r = (g_own.cpu() * (0.5 + torch.rand(g_own.numel(), generator=gen))).to(g_own.device)
            tr.peer.peer_words(seq, r, g_r)
            parts.append(g_r)
        want = _ordered_sum(parts)
        tr.step(sid)
        tr.synchronize()
        assert int(tr.peer.seq[0].item()) == seq and int(tr.peer.seq[1].item()) == 0
        assert torch.equal(tr.gflat, want), (tr.gflat - want).abs().max().item()
        # the words this rank sent to every peer: its own gradient, tagged with the launch's sequence number
        n = tr.flat.numel()
        off = ((seq & 1) * world + rank) * n
        for r in range(world):
            words = tr.peer.bufs[r][off:off + n]
            assert bool((words >> 32 == seq).all())
            assert torch.equal((words & 0xFFFFFFFF).to(torch.int32).view(torch.float32), g_own)
        shadow.grad = want.cpu().clone()
        optim.step()
        err = (tr.flat.detach().cpu() - shadow.detach()).abs().max().item() / shadow.detach().abs().max().item()
        assert err <= 1e-6, err
    assert int(tr.step_count.item()) == 3


def test_exchange_inside_a_cuda_graph_replay():
    world, rank = 4, 1
    tr, sid = _trainer(world, rank, use_graph=True)
    # capture() warms up with two real launches (exchanges 1 and 2) before it records the graph and restores
    # the optimizer state afterwards: the loopback peers must have "sent" for those two as well
    zeros = torch.zeros(tr.flat.numel())
    for seq in (1, 2):
        for r in range(world):
            if r != rank:
                tr.peer.peer_words(seq, r, zeros)
    tr.capture(sid)
    assert int(tr.peer.seq[1].item()) == 0
    seq0 = int(tr.peer.seq[0].item())
    g_own = _local_gradient(tr, sid)
    parts = [g_own if r == rank else (g_own * (1.0 + 0.25 * r)) for r in range(world)]
    for r in range(world):
        if r != rank:
            tr.peer.peer_words(seq0 + 1, r, parts[r])
    tr.step(sid)                       # graph replay
    tr.synchronize()
    assert torch.equal(tr.gflat, _ordered_sum(parts))


def test_missing_peer_times_out_sets_the_error_word_and_skips_adam():
    tr, sid = _trainer(2, 0, timeout_ms=50)
    before = tr.flat.detach().clone()
    t0 = time.perf_counter()
    tr.step(sid)                       # rank 1 never sends
    tr.stream.synchronize()
    waited = time.perf_counter() - t0
    assert 0.04 <= waited < 5.0, waited
    assert int(tr.peer.seq[1].item()) == 1
    assert torch.equal(tr.flat, before) and int(tr.step_count.item()) == 0      # no Adam step
    with pytest.raises(RuntimeError, match="timed out"):
        tr.check_peer()
    # the error is sticky and later launches do not wait for the timeout again
    t0 = time.perf_counter()
    for _ in range(20):
        tr.step(sid)
    tr.stream.synchronize()
    assert time.perf_counter() - t0 < 0.5
    assert torch.equal(tr.flat, before)
    with pytest.raises(RuntimeError, match="timed out"):
        tr.synchronize()
