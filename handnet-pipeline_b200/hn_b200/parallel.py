"""Data-parallel plumbing: frames shard by contiguous slices over the ranks, each rank runs the whole path on
its slice, and one fixed-size all-gather of per-frame records brings the results together.
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).

    joints, crops, has_hand = parallel.run_sharded(net, rgb, depth)          # every rank, same arguments

`net` is a HandNet (or any object with ``submit_records(images, depth, post) -> ticket`` / ``result_records(ticket)``).
The all-gather is enqueued by the step's ``post`` hook on the stream that produced the records, so it needs no host
synchronisation and overlaps the next step's detector like the rest of the pose stage.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `total` frames for `rank`; the first total % world ranks get one more."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_records(rec: torch.Tensor, frames_per_rank: int, out: torch.Tensor = None) -> torch.Tensor:
    """All-gather of [frames_per_rank, W] records -> [world * frames_per_rank, W] in rank order.
    Ranks with fewer frames pad their slice (callers trim with shard_range)."""
    world = dist.get_world_size()
    if rec.shape[0] < frames_per_rank:
        pad = torch.zeros((frames_per_rank - rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
        rec = torch.cat((rec, pad))
    if out is None:
        out = torch.empty((world * frames_per_rank, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous())
    return out


def _world_rank() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def submit_sharded(net, rgb, depth):
    """Enqueue this rank's slice of the global batch (`rgb`: sequence of [3,H,W] frames or a [N,3,H,W] tensor, `depth`
    [N,C,H,W]; every rank passes the same global batch, or at least the same N) and the all-gather of the per-frame
    records.  Returns a ticket for ``result_sharded``; does not wait for the device."""
    world, rank = _world_rank()
    total = len(rgb)
    begin, end = shard_range(total, world, rank)
    hands = int(getattr(net, "max_hands", 1))            # record rows per frame (HandNet(max_hands=H): one per hand slot)
    per_rank = ((total + world - 1) // world) * hands
    holder = {}

    def post(rec: torch.Tensor):
        # runs on the stream that holds the finished records of this step
        holder["all"] = gather_records(rec, per_rank) if world > 1 else rec.clone()

    images = rgb[begin:end]
    if isinstance(images, torch.Tensor):
        images = list(images.unbind(0))
    ticket = net.submit_records(list(images), depth[begin:end], post)
    return (net, ticket, holder, total, world, per_rank, hands)


def result_sharded(handle):
    """Wait for a step enqueued by ``submit_sharded``: (joints [N,21,3], crops [N,4] int64, has_hand [N] bool) for the
    WHOLE batch, on every rank (device tensors; rank 0 is the consumer in the HandNet deployment).  With
    ``HandNet(max_hands=H)``, H > 1, the leading dimension is N*H: row n*H + h is hand slot h of frame n."""
    from .runtime import unpack_records
    net, ticket, holder, total, world, per_rank, hands = handle
    net.result_records(ticket)
    rec = holder["all"]
    if world > 1 and per_rank * world != total * hands:  # drop the padding rows of the short slices
        rows = []
        for r in range(world):
            b, e = shard_range(total, world, r)
            rows.extend(range(r * per_rank, r * per_rank + (e - b) * hands))
        rec = rec[torch.tensor(rows, device=rec.device)]
    return unpack_records(rec)


def run_sharded(net, rgb, depth):
    """Shard -> run -> all-gather, synchronously (see the module docstring)."""
    return result_sharded(submit_sharded(net, rgb, depth))
