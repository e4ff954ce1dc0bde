"""Data-parallel execution over the GPUs of one box: one process per GPU, weights replicated, the batch sharded
contiguously, and exactly one collective -- an all-gather of the [B_local, num_classes] logits.

The encoder has no cross-sample operation (no BatchNorm; reference vit.py / simple_vit.py), so nothing else is
exchanged.  Backend: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` items owned by `rank`; the first (total % world) ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def all_gather_logits(local: torch.Tensor, total: Optional[int] = None,
                      group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Gather per-rank logits [B_r, C] into [sum_r B_r, C] on every rank, rank order = batch order.

    Equal shards (the benchmark case) use a single all_gather_into_tensor on the current stream; ragged shards
    (total % world != 0, sizes given by shard_sizes) pad to the largest shard and trim after the one collective."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    local = local.contiguous()
    if total is None or total % world == 0:
        out = torch.empty((local.shape[0] * world,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    sizes = shard_sizes(total, world)
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[: local.shape[0]] = local
    out = torch.empty((mx * world,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0)


@torch.no_grad()
def data_parallel_forward(model: torch.nn.Module, local_images: torch.Tensor, total: Optional[int] = None,
                          group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Forward this rank's shard and return the logits of the WHOLE batch on every rank."""
    return all_gather_logits(model(local_images), total=total, group=group)
