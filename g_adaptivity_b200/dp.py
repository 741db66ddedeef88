"""Data-parallel plumbing of the deformer training step.

A batch is a disjoint union of meshes, so the path shards by whole meshes with no data-path
collective: rank r of `world` owns the contiguous mesh range `shard_range(num_meshes, r, world)`.
The only exchange is one all-reduce (SUM) of the flat parameter gradient per step.  The reference
loss is a mean over ALL nodes of the global batch (`F.l1_loss`, run_GNN.py:80-84), so every rank
scales its local cotangent by `local_grad_scale` and the SUM all-reduce yields the global-batch
gradient exactly (for equal shard sizes; ragged shards weight by their node counts).

Pure host logic (works on CPU tensors with the gloo backend: tests/test_dp_gloo.py); the trainer
uses it with NCCL over NVLink.
"""
from __future__ import annotations

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


def broadcast_flat(flat: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    if world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat
