"""Host->device copy ceiling of the box: every rank copies a pinned buffer to its GPU in a loop, first rank 0 alone,
then all ranks at once (torchrun).  Prints GB/s per rank and the aggregate -- the number the end-to-end training
loop (bench.py `e2e`, 16 B per node per step) is bounded by.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/h2d_probe.py"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist


def rate(dev, host, devbuf, seconds=1.0):
    st = torch.cuda.Stream(device=dev)
    n = 0
    with torch.cuda.stream(st):
        for _ in range(3):
            devbuf.copy_(host, non_blocking=True)
        st.synchronize()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(8):
                devbuf.copy_(host, non_blocking=True)
            st.synchronize()
            n += 8
        dt = time.perf_counter() - t0
    return n * host.numel() / dt / 1e9


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    for mb in (4, 64):        # 4 MB ~ one step's inputs of cfg 2 (3.7 MB); 64 MB = large-transfer ceiling
        host = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
        devbuf = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
        alone = None
        if world > 1:
            if rank == 0:
                alone = rate(dev, host, devbuf)
            dist.barrier()
        r = rate(dev, host, devbuf)
        if world > 1:
            t = torch.tensor([r], device=dev)
            allr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            rates = [round(float(x), 2) for x in allr]
        else:
            rates = [round(r, 2)]
        out[f"{mb}MB"] = {"rank0_alone_gb_s": None if alone is None else round(alone, 2), "all_ranks_gb_s": rates,
                          "aggregate_gb_s": round(sum(rates), 1)}
    if rank == 0:
        try:
            out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception as ex:  # noqa
            out["topo"] = str(ex)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
