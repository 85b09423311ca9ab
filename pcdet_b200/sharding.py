"""Frame sharding across ranks: one process per GPU, frames are independent units (no data-path collective).

Mirrors what the reference gets from `DistributedSampler` + one process per GPU
(pcdet/datasets/__init__.py:27, tools/scripts/dist_train.sh:7): rank r owns frames r, r+world, r+2*world, ...
Only the bookkeeping around the hot path uses torch.distributed (barrier, max-over-ranks timing, gathering the
per-frame keep lists); inference itself exchanges nothing.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def frames_of_rank(num_frames: int, rank: int, world: int) -> List[int]:
    """Strided ownership like torch's DistributedSampler without padding: every frame exactly once."""
    return list(range(rank, num_frames, world))


def batches_of_rank(num_frames: int, rank: int, world: int, frames_per_batch: int) -> List[List[int]]:
    own = frames_of_rank(num_frames, rank, world)
    return [own[i:i + frames_per_batch] for i in range(0, len(own), frames_per_batch)]


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU number is the maximum over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frame_results(local: Sequence, num_frames: int):
    """local: results for frames_of_rank(...) in that order.  Returns the list for all frames in frame order
    on every rank (all_gather_object; a few hundred kept indices per frame)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [None] * world
    dist.all_gather_object(parts, list(local))
    out = [None] * num_frames
    for r, part in enumerate(parts):
        for i, f in enumerate(frames_of_rank(num_frames, r, world)):
            out[f] = part[i]
    assert all(o is not None for o in out) or num_frames == 0
    return out
