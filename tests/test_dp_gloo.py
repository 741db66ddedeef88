"""World-size-2 data-parallel host logic on CPU (gloo): mesh sharding, cotangent scaling and the
gradient all-reduce give exactly the single-process gradient of the global batch.  The per-rank
compute is the CPU oracle (the CUDA product path needs a GPU); what is under test is the plumbing
of g_adaptivity_b200/dp.py that the trainer runs over NCCL."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from g_adaptivity_b200 import dp, synth
from oracle import gnn_oracle

MESH, NUM_MESHES = (8, 8), 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat_grad(model):
    return torch.cat([p.grad.reshape(-1) for n, p in model.named_parameters() if p.grad is not None])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        opt = synth.default_opt(MESH)
        ds = synth.SyntheticDataset(2, MESH)
        torch.manual_seed(42 + rank)                 # deliberately different initial weights per rank
        model = gnn_oracle.GNNRef(ds, copy.deepcopy(opt))
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        dp.broadcast_flat(flat, src=0)               # rank 0's parameters everywhere
        o = 0
        with torch.no_grad():
            for p in model.parameters():
                p.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
        first, last = dp.shard_range(NUM_MESHES, rank, world)
        data = synth.make_batch(MESH, last - first, seed=7, first_mesh_id=first)
        out = model(data)
        count_local = out.numel()
        count_global = NUM_MESHES * MESH[0] * MESH[1] * 2
        # sum of local |d| scaled so that SUM over ranks == mean over the global batch
        loss_local = (out - data.x_phys).abs().sum() * dp.local_grad_scale(count_local, count_global, world)
        loss_local.backward()
        g = _flat_grad(model)
        dp.allreduce_flat(g)
        torch.save({"grad": g, "flat": flat, "range": (first, last)}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_meshes():
    for n in (1, 5, 8, 8192):
        for world in (1, 2, 3, 8):
            got = [dp.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)
    assert dp.local_grad_scale(100, world=4) == pytest.approx(1.0 / 400)


@pytest.mark.timeout(180)
def test_two_rank_gloo_gradient_equals_global_batch(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert res[0]["range"] == (0, 3) and res[1]["range"] == (3, 6)
    assert torch.equal(res[0]["flat"], res[1]["flat"])              # broadcast worked
    assert torch.equal(res[0]["grad"], res[1]["grad"])              # every rank holds the same sum
    # single process, global batch, mean loss
    opt = synth.default_opt(MESH)
    ds = synth.SyntheticDataset(2, MESH)
    torch.manual_seed(42)
    model = gnn_oracle.GNNRef(ds, copy.deepcopy(opt))
    data = synth.make_batch(MESH, NUM_MESHES, seed=7, first_mesh_id=0)
    gnn_oracle.mesh_loss(model(data), data.x_phys).backward()
    want = _flat_grad(model)
    scale = want.abs().max().item()
    assert (res[0]["grad"] - want).abs().max().item() <= 2e-6 * scale
