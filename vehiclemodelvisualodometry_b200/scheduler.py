"""Window scheduler: shards the windows of a batch of drives across the GPUs of one box.

Windows are independent, so the data path needs no collective: every rank
holds the (small) pose streams, searches a contiguous range of the global window list, and
the only exchange is one all-gather of the fixed-size 64-byte result records
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The search kernel writes its records
straight into this rank's slice of the gather buffer, so there is no staging copy.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

RECORD_BYTES = 64


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first ``n % world`` ranks take one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_capacity(n_items: int, world: int) -> int:
    return -(-int(n_items) // int(world))


def assign_drives(window_counts: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of whole drives to ranks (for seed modes
    that serialise the windows of a drive); returns the drive ids per rank."""
    order = sorted(range(len(window_counts)), key=lambda d: (-window_counts[d], d))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for d in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(d)
        load[r] += window_counts[d]
    for r in range(world):
        out[r].sort()
    return out


def gather_records(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather of per-rank record blocks -> [n_items, 64] on every rank.

    ``local`` is this rank's padded block [capacity, 64] (uint8) whose first
    ``hi - lo`` rows are valid.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    cap = shard_capacity(n_items, world)
    if local.shape != (cap, RECORD_BYTES):
        raise ValueError(f"local block must be [{cap}, {RECORD_BYTES}]")
    if world == 1:
        return local[:n_items]
    full = torch.empty((world * cap, RECORD_BYTES), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(full, local, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_items, world, r)
        parts.append(full[r * cap: r * cap + (hi - lo)])
    return torch.cat(parts, dim=0)


def search_sharded(search_fn: Callable[[Tuple[int, int], torch.Tensor], None], n_windows: int,
                   device, group=None) -> torch.Tensor:
    """Run ``search_fn((lo, hi), out_block)`` on this rank's window range and gather.

    ``search_fn`` must fill ``out_block[: hi - lo]`` (uint8 [hi-lo, 64]) with the records of
    windows lo..hi-1 of the global plan, e.g.
    ``lambda rng, out: grid_search(cfg, drives, plan, window_range=rng, out=out)``.
    """
    if dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_range(n_windows, world, rank)
    cap = shard_capacity(n_windows, world)
    block = torch.zeros((cap, RECORD_BYTES), dtype=torch.uint8, device=device)
    if hi > lo:
        search_fn((lo, hi), block[: hi - lo])
    return gather_records(block, n_windows, group)
