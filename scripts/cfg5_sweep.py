#!/usr/bin/env python
"""BASELINE.json configs[4]: data-parallel training on 50x50 meshes, GLOBAL batch 8192 (strong
scaling: every rank owns 8192 / world meshes), gradient all-reduce inside the training kernel over
NVLink peer memory.  Run single-process or under torchrun; rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node N scripts/cfg5_sweep.py [--global-batch 8192]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

PEAK = 6539.9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=8192)
    ap.add_argument("--mesh", type=int, nargs=2, default=[50, 50])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--ring", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    from g_adaptivity_b200 import GNN, dp, synth
    from g_adaptivity_b200.trainer import DeformerTrainer
    md = tuple(a.mesh)
    first, last = dp.shard_range(a.global_batch, rank, world)
    B = last - first
    opt = synth.default_opt(md, device=str(dev), gad_store_alpha=False)
    ds = synth.SyntheticDataset(2, md)
    torch.manual_seed(42)
    model = GNN(ds, opt).to(dev)
    tr = DeformerTrainer(model)
    tr.broadcast_parameters()
    sids = [tr.add_batch(synth.make_batch(md, B, seed=1000, first_mesh_id=r * a.global_batch + first)) for r in range(a.ring)]
    tr.capture_epoch(sids)
    R = a.ring
    for _ in range(max(1, a.warmup // R)):
        tr.run_epoch(sids)
    tr.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, a.steps // R)
    e0.record(tr.stream)
    for _ in range(reps):
        tr.run_epoch(sids)
    e1.record(tr.stream)
    tr.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (reps * R)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    s = tr.slots[0]
    n_global = a.global_batch * md[0] * md[1]
    d = s.graph.E / s.N
    bpn = 4 * ((8 * 4 + 4 * d + 4) + (12 * 4 + 2 * (4 * d + 4))) + 4 * 4 + 8 * 2
    if rank == 0:
        print(json.dumps({
            "config": f"cfg5: {a.global_batch} x {md[0]}x{md[1]} meshes global batch, train step, strong scaling",
            "n_gpus": world, "meshes_per_gpu": B, "ms_per_step": round(ms, 4), "nodes_per_s": n_global / (ms * 1e-3),
            "per_gpu_roofline_frac": round((s.N * bpn / (ms * 1e-3) / 1e9) / PEAK, 4), "bytes_per_node": round(bpn, 1),
            "fused_dp": bool(tr.fused_dp), "tiles_per_gpu": s.graph.T, "loss": float(s.loss.item())}), flush=True)
    if world > 1:
        tr.close()
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
