"""Window scheduler: shards the windows of a batch of drives across the GPUs of one box.

Windows are independent, so the data path needs no collective: every rank
holds the (small) pose streams, searches a contiguous range of the global window list, and
the only exchange is one all-gather of the fixed-size 64-byte result records
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The search kernel writes its records
straight into this rank's slice of the gather buffer, so there is no staging copy.

``PeerGather`` removes the collective as well: the gather buffers of the ranks of one box are
mapped into each other's address space (CUDA IPC) and the search kernel's epilogue stores every
record into all of them (``vmvo_set_result_mirrors``) -- the exchange rides on the kernel that
produces the data, as plain NVLink stores, with no second kernel competing for the SMs the
persistent search occupies.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

RECORD_BYTES = 64


class _DevicePointer:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class PeerGather:
    """Gather buffers of all ranks of one box, mapped into each other.

    Every rank owns ``buffer`` = uint8 [world * n_records, 64]; rank r's records belong in rows
    [r * n_records, (r + 1) * n_records) of EVERY rank's buffer.  Between ``enable()`` and
    ``disable()`` each record a search on this rank writes to ``local`` (its own slot) is also
    stored into that slot of all peers' buffers by the kernel itself.  The peers see them once the
    kernel has completed here -- order as after any kernel (event, stream sync, barrier).
    """

    def __init__(self, n_records: int, device, group=None):
        import ctypes as C

        from . import _lib

        if not dist.is_initialized():
            raise RuntimeError("PeerGather needs an initialised process group")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world - 1 > 16:
            raise ValueError("at most 17 ranks")
        self.n = int(n_records)
        self.ctx = _lib.context(torch.device(device).index)
        nbytes = self.world * self.n * RECORD_BYTES
        ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
        self.ctx.check(self.ctx.lib.vmvo_peer_buffer_create(self.ctx.handle, nbytes, C.byref(ptr), handle),
                       "vmvo_peer_buffer_create")
        self._ptr = ptr.value
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._peers: List[int] = []
        for q, h in enumerate(handles):
            if q == self.rank:
                continue
            buf, out = (C.c_uint8 * 64).from_buffer_copy(h), C.c_void_p()
            self.ctx.check(self.ctx.lib.vmvo_peer_buffer_open(self.ctx.handle, buf, C.byref(out)),
                           "vmvo_peer_buffer_open")
            self._peers.append(out.value)
        self._view = _DevicePointer(self._ptr, (self.world * self.n, RECORD_BYTES))
        self.buffer = torch.as_tensor(self._view, device=torch.device(device))
        self.buffer.zero_()
        self.local = self.buffer[self.rank * self.n:(self.rank + 1) * self.n]
        torch.cuda.synchronize(device)
        dist.barrier(group)               # every buffer exists and is zeroed before anyone stores into it

    def enable(self) -> None:
        import ctypes as C

        arr = (C.c_void_p * max(len(self._peers), 1))(*self._peers)
        self.ctx.check(self.ctx.lib.vmvo_set_result_mirrors(self.ctx.handle, len(self._peers), arr,
                                                            self.rank * self.n), "vmvo_set_result_mirrors")

    def disable(self) -> None:
        self.ctx.check(self.ctx.lib.vmvo_set_result_mirrors(self.ctx.handle, 0, None, 0),
                       "vmvo_set_result_mirrors")

    def close(self) -> None:
        self.disable()
        torch.cuda.synchronize()
        dist.barrier()
        for p in self._peers:
            self.ctx.lib.vmvo_peer_buffer_close(self.ctx.handle, p)
        self._peers = []
        self.buffer = self.local = None
        if self._ptr:
            self.ctx.lib.vmvo_peer_buffer_destroy(self.ctx.handle, self._ptr)
            self._ptr = None


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
