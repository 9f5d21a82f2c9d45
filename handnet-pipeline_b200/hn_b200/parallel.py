"""Data-parallel plumbing: frames shard by contiguous slices over the ranks, each rank runs the whole path on
its slice, and one fixed-size all-gather of per-frame records brings the results together.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `total` frames for `rank`; the first total % world ranks get one more."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_records(rec: torch.Tensor, frames_per_rank: int) -> torch.Tensor:
    """All-gather of [frames_per_rank, W] records -> [world * frames_per_rank, W] in rank order.
    Ranks with fewer frames pad their slice (callers trim with shard_range)."""
    world = dist.get_world_size()
    if rec.shape[0] < frames_per_rank:
        pad = torch.zeros((frames_per_rank - rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
        rec = torch.cat((rec, pad))
    out = torch.empty((world * frames_per_rank, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous())
    return out
