"""Multi-GPU plumbing of the hot path: clip sharding and the statistics all-reduce.

Clips (and event windows) are independent units: rank r of W simulates
clips r, r+W, r+2W, ... — the rule ``DistributedSampler`` applies in the reference
(train.py:54-56) — with no collective on the data path.  The only exchange is one
all-reduce(sum) of a small int64 statistics vector at the end (NCCL over
NVLink/NVSwitch on GPUs; the same code runs over gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def shard_indices(num_units: int, rank: int, world_size: int) -> List[int]:
    """Indices of the clips / windows owned by ``rank`` (round robin, like DistributedSampler without padding)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_units, world_size))


def shard_counts(num_units: int, world_size: int) -> List[int]:
    return [len(range(r, num_units, world_size)) for r in range(world_size)]


STAT_FIELDS = ("positive_events", "negative_events", "pixel_intervals", "clips")


def pack_stats(stats: Optional[torch.Tensor], pixel_intervals: int, clips: int, device=None,
               bin_abs_sums: Optional[torch.Tensor] = None, count_map: Optional[torch.Tensor] = None) -> torch.Tensor:
    """This rank's statistics as ONE int64 vector for the closing all-reduce: STAT_FIELDS, then the
    optional per-bin |count| sums ``[bins]`` (``consumer.bin_abs_sums``), then the optional per-pixel event-count map
    ``[H,W]`` flattened (``events.event_count_map``; scripts/testset_evcnt_maps.py:19-25).  ``stats`` is the kernels'
    ``[B,2]`` per-clip event totals.  ``unpack_stats`` splits the reduced vector again."""
    dev = device if device is not None else (stats.device if stats is not None else "cpu")
    v = torch.zeros(len(STAT_FIELDS), dtype=torch.int64, device=dev)
    if stats is not None and stats.numel():
        s = stats.to(torch.int64).reshape(-1, 2).sum(dim=0)
        v[0], v[1] = s[0], s[1]
    v[2], v[3] = int(pixel_intervals), int(clips)
    parts = [v]
    if bin_abs_sums is not None:
        parts.append(torch.round(bin_abs_sums.to(dev, torch.float64)).to(torch.int64).reshape(-1))
    if count_map is not None:
        parts.append(count_map.to(dev, torch.int64).reshape(-1))
    return torch.cat(parts) if len(parts) > 1 else v


def unpack_stats(vec: torch.Tensor, num_bins: int = 0, map_shape=None) -> dict:
    """Inverse of ``pack_stats`` on the (reduced) vector: the STAT_FIELDS as ints, ``bin_abs_sums`` and ``count_map`` tensors."""
    out = {k: int(x) for k, x in zip(STAT_FIELDS, vec[: len(STAT_FIELDS)].tolist())}
    o = len(STAT_FIELDS)
    if num_bins:
        out["bin_abs_sums"] = vec[o: o + num_bins]
        o += num_bins
    if map_shape is not None:
        out["count_map"] = vec[o: o + int(map_shape[0]) * int(map_shape[1])].reshape(tuple(map_shape))
    return out


def allreduce_stats(vec: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the statistics vector over all ranks (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def stats_dict(vec: torch.Tensor) -> dict:
    return {k: int(v) for k, v in zip(STAT_FIELDS, vec[: len(STAT_FIELDS)].tolist())}
