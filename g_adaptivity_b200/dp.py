"""Data-parallel plumbing of the deformer training step.

A batch is a disjoint union of meshes, so the path shards by whole meshes with no data-path
collective: rank r of `world` owns the contiguous mesh range `shard_range(num_meshes, r, world)`.
The only exchange is one all-reduce (SUM) of the flat parameter gradient per step.  The reference
loss is a mean over ALL nodes of the global batch (`F.l1_loss`, run_GNN.py:80-84), so every rank
scales its local cotangent by `local_grad_scale` and the SUM all-reduce yields the global-batch
gradient exactly (for equal shard sizes; ragged shards weight by their node counts).

Pure host logic (works on CPU tensors with the gloo backend: tests/test_dp_gloo.py); the trainer
uses it with NCCL over NVLink.

`PeerExchange` sets up the faster route used on one NVLink / NVSwitch node: every rank allocates a
small receive buffer in the CUDA library, the ranks all-gather the CUDA IPC handles over
`torch.distributed` and map each other's buffers; the training kernel then performs the gradient
all-reduce itself, by peer stores and loads inside its tail (csrc/ell_kernels.cuh:
peer_allreduce), so a data-parallel step stays ONE launch.  NCCL remains the set-up plumbing and
the fallback (other nodes, IPC unavailable).
"""
from __future__ import annotations

import ctypes as C
import socket
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_meshes: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, last) mesh ids of `rank`; the first `num_meshes % world` ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(num_meshes), int(world))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def local_grad_scale(count_local: int, count_global: Optional[int] = None, world: int = 1) -> float:
    """Factor on d(sum of local per-entry losses)/d(out) such that a SUM all-reduce of the parameter
    gradients gives the gradient of the mean over the GLOBAL batch.  With equal shards
    (count_global = world * count_local) this is 1 / (count_local * world)."""
    if count_global is None:
        count_global = int(count_local) * int(world)
    return 1.0 / float(count_global)


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def allreduce_flat(gflat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the flat gradient (144 + L floats: latency-bound)."""
    if world_size(group) > 1:
        dist.all_reduce(gflat, op=dist.ReduceOp.SUM, group=group)
    return gflat


def allreduce_flat_ordered(gflat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce with a FIXED reduction order: all-gather the per-rank vectors, then add them
    in rank order in fp32 (0 + g_0 + g_1 + ...), exactly the order of the in-kernel peer exchange
    (csrc/ell_kernels.cuh: peer_allreduce), so both routes give the same bits on every rank."""
    world = world_size(group)
    if world > 1:
        parts = [torch.empty_like(gflat) for _ in range(world)]
        dist.all_gather(parts, gflat.contiguous(), group=group)
        acc = parts[0].clone()
        for r in range(1, world):
            acc += parts[r]
        gflat.copy_(acc)
    return gflat


def broadcast_flat(flat: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    if world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def probe_host_path(dev: torch.device, mbytes: int = 4, seconds: float = 0.15) -> float:
    """GB/s of pinned host -> device copies of `mbytes` MB on `dev`, in a loop for `seconds` (call it on every rank at
    the same time to see the paths under load)."""
    import time
    host = torch.empty(mbytes << 20, dtype=torch.uint8).pin_memory()
    buf = torch.empty(mbytes << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    n = 0
    with torch.cuda.stream(st):
        for _ in range(2):
            buf.copy_(host, non_blocking=True)
        st.synchronize()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                buf.copy_(host, non_blocking=True)
            st.synchronize()
            n += 4
        dt = time.perf_counter() - t0
    return n * host.numel() / dt / 1e9


def plan_host_relays(rates, devices, min_fraction: float = 0.12):
    """Pair the ranks with the slowest host -> device paths with those with the fastest and choose, per pair, the
    fraction y of the slow rank's bytes that should travel through the fast rank's GPU so that both finish together:
    (1 - y) / R_slow = (1 + y) / R_fast, i.e. y = (R_fast - R_slow) / (R_fast + R_slow).  `rates[q]` GB/s of rank q
    under load, `devices[q]` its device ordinal.  Returns {rank: (partner device, y)} for the ranks that relay."""
    order = sorted(range(len(rates)), key=lambda q: (rates[q], q))
    plan = {}
    for i in range(len(order) // 2):
        slow, fast = order[i], order[-1 - i]
        rs, rf = float(rates[slow]), float(rates[fast])
        if rs <= 0 or rf <= rs:
            continue
        y = (rf - rs) / (rf + rs)
        if y >= min_fraction and devices[slow] != devices[fast]:
            plan[slow] = (int(devices[fast]), y)
    return plan


def balance_host_paths(dev: torch.device, group=None, min_fraction: float = 0.12, rounds: int = 3):
    """One node, one process per GPU, every GPU visible to every process: measure all host -> device paths at once
    (median of `rounds` synchronised probes: a single short probe is noisy) and return (this rank's relay (partner
    device, fraction) or None, all rates).  Paths closer than a factor 1.27 (fraction < 0.12) are left alone.
    Collective."""
    import statistics
    import torch.distributed as dist
    world = world_size(group)
    if world < 2 or torch.cuda.device_count() < world:
        return None, None
    samples = []
    for _ in range(rounds):
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        samples.append(probe_host_path(dev, seconds=0.2))
    t = torch.tensor([statistics.median(samples), float(dev.index)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    rates = [float(o[0]) for o in out]
    devices = [int(o[1]) for o in out]
    plan = plan_host_relays(rates, devices, min_fraction)
    return plan.get(dist.get_rank(group)), rates


class PeerExchange:
    """Receive buffers of all ranks of `group`, mapped into this process (CUDA IPC), for the
    in-kernel gradient all-reduce.  `ok` is the same on every rank: the set-up ends with a MIN
    all-reduce of the local outcome, so either all ranks use the peer path or none does."""

    MAX_PEERS = 16     # GAD_MAX_PEERS (include/gadapt.h)

    def __init__(self, lib, n_params: int, device: torch.device, group=None):
        self.lib, self.dev, self.group = lib, device, group
        self.world = world_size(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.own = C.c_void_p()
        self.opened = []
        self.ptrs = None            # int64 [world] device array of receive-buffer pointers
        self.seq = None             # uint32 [2]: launch sequence (advanced by the kernel), error word
        self.ok = False
        self.why = ""
        if self.world <= 1:
            self.why = "single rank"
            return
        good, addrs = 1, [0] * self.world
        handle = C.create_string_buffer(64)
        try:
            if self.world > self.MAX_PEERS:
                raise RuntimeError(f"world {self.world} > {self.MAX_PEERS}")
            with torch.cuda.device(device):
                nbytes = lib.gad_peer_exchange_bytes(self.world, int(n_params))
                if lib.gad_peer_alloc(nbytes, C.byref(self.own), handle) != 0:
                    raise RuntimeError(lib.gad_last_error().decode())
        except Exception as e:   # noqa: BLE001 -- any local failure must still reach the collective below
            good, self.why = 0, f"alloc: {e}"
        mine = (socket.gethostname(), bytes(handle.raw), good)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if good and any(h[0] != mine[0] for h in everyone):
            good, self.why = 0, "ranks on different hosts"
        if good and not all(h[2] for h in everyone):
            good, self.why = 0, "a peer failed to allocate"
        if good:
            try:
                with torch.cuda.device(device):
                    for r, (_, raw, _) in enumerate(everyone):
                        if r == self.rank:
                            addrs[r] = self.own.value
                            continue
                        p = C.c_void_p()
                        if lib.gad_peer_open(C.create_string_buffer(raw, 64), C.byref(p)) != 0:
                            raise RuntimeError(lib.gad_last_error().decode())
                        self.opened.append(p)
                        addrs[r] = p.value
            except Exception as e:   # noqa: BLE001
                good, self.why = 0, f"open: {e}"
        flag = torch.tensor([good], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(flag.item())
        if self.ok:
            self.ptrs = torch.tensor(addrs, dtype=torch.int64, device=device)
            self.seq = torch.zeros(2, dtype=torch.int32, device=device)
        else:
            self.why = self.why or "a peer could not map the buffers"
            self.close()

    @classmethod
    def loopback(cls, lib, n_params: int, device: torch.device, world: int, rank: int) -> "PeerExchange":
        """Single-process stand-in used by the self-tests (tests/test_gpu_peer_loopback.py): `world` receive
        buffers, all on this device, owned by torch.  The training kernel behaves exactly as on `world` GPUs -- it
        stores its tagged gradient words into every buffer and waits for the other ranks' words in buffer `rank` --
        and the test plays the peers by writing their words (`peer_words`) into that buffer.  This is how the
        4- and 8-rank exchange, the fixed summation order and the bounded wait are checked on one GPU."""
        self = cls.__new__(cls)
        self.lib, self.dev, self.group = lib, device, None
        self.world, self.rank = int(world), int(rank)
        self.own, self.opened, self.why = C.c_void_p(), [], "loopback"
        if not (1 < self.world <= cls.MAX_PEERS and 0 <= self.rank < self.world):
            raise ValueError(f"loopback exchange: rank {rank} / world {world}")
        self.n_params = int(n_params)
        self.bufs = [torch.zeros(2 * self.world * self.n_params, dtype=torch.int64, device=device)
                     for _ in range(self.world)]
        self.ptrs = torch.tensor([b.data_ptr() for b in self.bufs], dtype=torch.int64, device=device)
        self.seq = torch.zeros(2, dtype=torch.int32, device=device)
        self.ok = True
        return self

    def peer_words(self, seq: int, src_rank: int, values: torch.Tensor):
        """Loopback only: deposit what rank `src_rank` would send in exchange number `seq` (fp32 `values`
        [n_params]) into this rank's receive buffer: words {value bits, seq} in slot [seq & 1][src_rank]."""
        n = self.n_params
        bits = values.detach().to(self.dev, torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        words = bits | (int(seq) << 32)
        off = ((int(seq) & 1) * self.world + int(src_rank)) * n
        self.bufs[self.rank][off:off + n].copy_(words)

    def close(self):
        """Unmap the peers' buffers and free the own one (all ranks must be past their last step)."""
        with torch.cuda.device(self.dev):
            for p in self.opened:
                self.lib.gad_peer_close(p)
            self.opened = []
            if self.own.value:
                self.lib.gad_peer_free(self.own)
                self.own = C.c_void_p()
        self.ok = False
