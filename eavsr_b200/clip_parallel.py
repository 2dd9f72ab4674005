"""Clip-parallel execution: the only way the path shards.

The bidirectional second-order recurrence makes one clip serial in time and across the four
propagation branches (models/eavsrp_model.py:229-238,271-324), but clips are independent.  One
process per GPU; rank r owns clips {i : i mod world == r}; weights are replicated; there is NO
collective on the data path.  torch.distributed is used only to agree on timings (max over ranks)
and, optionally, to gather per-clip results/checksums on rank 0.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, List

import torch
import torch.distributed as dist

__all__ = ["world", "shard", "run_sharded", "max_over_ranks", "gather_results"]


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard(num_clips: int, rank: int, world_size: int) -> List[int]:
    """Round-robin clip ownership: rank r gets {i : i mod world_size == r}."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_clips, world_size))


def run_sharded(process_clip: Callable[[int], object], num_clips: int, rank: int, world_size: int) -> Dict[int, object]:
    """Run `process_clip(i)` for this rank's clips; returns {clip id: result}."""
    return {i: process_clip(i) for i in shard(num_clips, rank, world_size)}


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (e.g. CUDA-event milliseconds) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local: Dict[int, object]) -> Dict[int, object]:
    """All ranks' {clip id: result} merged (identical on every rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    parts: List[Dict[int, object]] = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged: Dict[int, object] = {}
    for p in parts:
        overlap = merged.keys() & p.keys()
        if overlap:
            raise RuntimeError(f"clips {sorted(overlap)} were processed by more than one rank")
        merged.update(p)
    return merged
