"""Worker of tests/test_gpu_dp.py: one rank of a data-parallel training run (torchrun, NCCL).

Each rank trains K steps on its own shard of meshes, first with the gradient all-reduce fused into
the training kernel over peer memory (dp.PeerExchange), then again from the same initial state with
NCCL between the train kernel and Adam.  Rank 0 writes both parameter vectors, the per-rank
agreement and -- for the caller to compare with the single-process oracle -- nothing else."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    out_path, K, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from g_adaptivity_b200 import GNN, synth
    from g_adaptivity_b200.trainer import DeformerTrainer

    md = (15, 15)
    ds = synth.SyntheticDataset(2, md)
    res = {}
    for mode in ("peer", "nccl", "peer_graph"):
        opt = synth.default_opt(md, device=str(dev), gad_store_alpha=False, lr=1e-2,
                                gad_peer_allreduce=(mode != "nccl"))
        torch.manual_seed(42)
        model = GNN(ds, opt).to(dev)
        tr = DeformerTrainer(model, use_cuda_graph=(mode == "peer_graph"))
        # the same explicit initial parameters on every rank (and in the oracle of the test)
        flat0 = (torch.rand(tr.flat.numel(), generator=torch.Generator().manual_seed(123)) - 0.5) * 0.7
        tr.flat.copy_(flat0.to(dev))
        tr.broadcast_parameters()
        # global batch = world * B meshes, mesh ids contiguous per rank (dp.shard_range)
        sids = [tr.add_batch(synth.make_batch(md, B, seed=7, first_mesh_id=(r * world + rank) * B)) for r in range(2)]
        losses = []
        g_first = None
        if mode == "peer":
            # the all-reduced gradient of one step from the initial parameters (state restored afterwards)
            saved = [t.clone() for t in (tr.flat, tr.exp_avg, tr.exp_avg_sq, tr.step_count)]
            tr.step(sids[0])
            tr.synchronize()
            g_first = tr.gflat.detach().cpu().tolist()
            for t, v in zip((tr.flat, tr.exp_avg, tr.exp_avg_sq, tr.step_count), saved):
                t.copy_(v)
            tr.sync_weights()
        if mode == "peer_graph":
            for _ in range(K // 2):
                losses += [l.clone() for l in tr.run_epoch(sids)]
        else:
            for k in range(K):
                losses.append(tr.step(sids[k % 2]).clone())
        tr.synchronize()
        selfcheck = tr.dp_selfcheck(sids, steps=4) if mode == "peer" else None      # restores the optimizer state
        flat = tr.flat.detach().clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        res[mode] = {"fused_dp": bool(tr.fused_dp), "why": (tr.peer.why if tr.peer is not None else "off"),
                     "flat": flat.cpu().tolist(), "ranks_equal": all(bool(torch.equal(g, flat)) for g in gathered),
                     "steps": int(tr.step_count.item()), "loss_last": float(losses[-1].item()), "g_first": g_first,
                     "selfcheck": selfcheck}
        tr.close()
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump(res, fh)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)      # skip NCCL teardown (graphs captured collectives): nothing left to clean up


if __name__ == "__main__":
    main()
