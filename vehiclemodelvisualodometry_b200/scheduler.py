"""Window scheduler: shards the windows of a batch of drives across the GPUs of one box.

Windows are independent (the reference's loop over them is
vmvo/scripts/optimize_trajectory_v2.py:48-146), so the data path needs no collective: every rank
holds the (small) pose streams and the plan of ALL windows, and the global window list is dealt to
the ranks block-cyclically -- window ``w`` belongs to rank ``(w // block) % world`` -- which spreads
the slow stretches of a drive (near-ties of a crawling vehicle) evenly instead of
leaving them to whichever rank owns that drive.  ``shard_range`` is the contiguous split (frames of
the write-back; whole-drive assignment for chained seeds is ``assign_drives``).

The only exchange is the fixed-size 64-byte result records, and ``PeerGather`` fuses it into the
kernels that produce and consume them: the gather buffers of the ranks of one box are mapped into
each other's address space (CUDA IPC), the search kernel's epilogue stores every record into all
of them (plain NVLink stores), and arrival is signalled by one flag word per (sender, receiver)
pair that the write-back kernel publishes and waits on (include/vmvo_b200.h, ``vmvo_exchange``) --
no collective, no extra launch, no host barrier inside a step.  ``gather_dealt`` is the collective
form of the same exchange (one all-reduce of the zero-initialised buffers; NCCL on GPUs without
peer access, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

RECORD_BYTES = 64
DEFAULT_BLOCK = 32
_FLAG_BYTES = 256          # arrival words [world <= 17] + step counter + timeout indicator


# ---- the deal (host mirror of the kernel's index arithmetic, csrc/vmvo_search.cu next_window) ----
def deal_owner(w, block: int, world: int):
    """Rank that searches global window ``w`` (array or int)."""
    return (np.asarray(w) // int(block)) % int(world)


def deal_indices(n_items: int, block: int, world: int, rank: int) -> np.ndarray:
    """Global indices of the windows dealt to ``rank``, in the order its queue hands them out."""
    w = np.arange(int(n_items), dtype=np.int64)
    return w[deal_owner(w, block, world) == rank]


def deal_count(n_items: int, block: int, world: int, rank: int) -> int:
    n, b = int(n_items), int(block)
    full, rest = divmod(n, b * world)
    return full * b + min(max(rest - rank * b, 0), b)


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


class _DevicePointer:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerGather:
    """One buffer set of the record exchange between the ranks of one box.

    Every rank owns ``buffer`` = uint8 [n_records, 64], one record per GLOBAL window, followed in
    the same allocation by its arrival words.  ``exchange`` is the ``vmvo_exchange`` to hand to
    ``grid_search(..., exchange=)`` and ``write_back(..., exchange=)``: the search stores each record
    it produces into every rank's buffer, the write-back publishes this rank's arrival word and
    waits for the peers' before it reads.  Use two sets alternately (``step % 2``) and no other
    synchronisation is needed between steps.

    The group's ranks must sit on one box with peer access (NVLink / NVSwitch); two ranks may even
    share one device (the CPU-free test of the exchange, tests/test_gpu_exchange.py).
    """

    def __init__(self, n_records: int, device, group=None, block: int = DEFAULT_BLOCK):
        from . import _lib

        if not dist.is_initialized():
            raise RuntimeError("PeerGather needs an initialised process group")
        self.group = group
        self.device = torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world - 1 > _lib.MAX_MIRRORS:
            raise ValueError(f"at most {_lib.MAX_MIRRORS + 1} ranks")
        if block < 1 or block & (block - 1):
            raise ValueError("block must be a power of two")
        self.n, self.block = int(n_records), int(block)
        self.ctx = _lib.context(self.device.index)
        rec_bytes = (self.n * RECORD_BYTES + 255) & ~255
        ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
        self.ctx.check(self.ctx.lib.vmvo_peer_buffer_create(self.ctx.handle, rec_bytes + _FLAG_BYTES,
                                                            C.byref(ptr), handle), "vmvo_peer_buffer_create")
        self._ptr = ptr.value
        self._flags_off = rec_bytes
        self._all = torch.as_tensor(_DevicePointer(self._ptr, (rec_bytes + _FLAG_BYTES,)), device=self.device)
        self._all.zero_()
        self.buffer = self._all[: self.n * RECORD_BYTES].view(self.n, RECORD_BYTES)
        # int32 view of the words behind the records: [0 .. world) arrival words, [32] step, [33] timeout
        self.words = self._all[rec_bytes:].view(torch.int32)
        torch.cuda.synchronize(self.device)
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
        ex = _lib.Exchange()
        ex.world, ex.rank, ex.block, ex.n_peers = self.world, self.rank, self.block, len(self._peers)
        for q, base in enumerate(self._peers):
            ex.peer_records[q] = base
            ex.peer_flags[q] = base + self._flags_off + 4 * self.rank
        ex.local_flags = self._ptr + self._flags_off
        ex.epoch = self._ptr + self._flags_off + 4 * 32
        self.exchange = ex
        dist.barrier(group)               # every buffer exists and is zeroed before anyone stores into it

    # -- the deal, for callers that want to know who searched what
    def my_windows(self) -> np.ndarray:
        return deal_indices(self.n, self.block, self.world, self.rank)

    def my_count(self) -> int:
        return deal_count(self.n, self.block, self.world, self.rank)

    # -- stand-alone arrival (the write-back kernel does both itself when handed ``exchange``)
    def publish(self) -> None:
        from . import _lib

        self.ctx.check(self.ctx.lib.vmvo_exchange_publish(self.ctx.handle, C.byref(self.exchange),
                                                          _lib.stream_ptr(self.device)), "vmvo_exchange_publish")

    def wait(self) -> None:
        from . import _lib

        self.ctx.check(self.ctx.lib.vmvo_exchange_wait(self.ctx.handle, C.byref(self.exchange),
                                                       _lib.stream_ptr(self.device)), "vmvo_exchange_wait")

    def step(self) -> int:
        """Steps this rank has started on this buffer set (synchronises)."""
        return int(self.words[32].item())

    def timed_out(self) -> int:
        """0, or 1 + the rank whose arrival word a wait gave up on (synchronises)."""
        return int(self.words[33].item())

    def close(self) -> None:
        if self._ptr is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)          # nobody is still storing into a buffer that is about to go
        for p in self._peers:
            self.ctx.lib.vmvo_peer_buffer_close(self.ctx.handle, p)
        self._peers = []
        self.buffer = self.words = self._all = None
        self.ctx.lib.vmvo_peer_buffer_destroy(self.ctx.handle, self._ptr)
        self._ptr = None


def gather_dealt(buffer: torch.Tensor, group=None) -> torch.Tensor:
    """The collective form of the exchange: ``buffer`` (uint8 [n, 64]) holds this rank's records at
    their global indices and ZEROS elsewhere; one all-reduce (sum of int32 words: every word is
    non-zero on at most one rank, and x + 0 is exact) leaves every record on every rank."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buffer.view(torch.int32), op=dist.ReduceOp.SUM, group=group)
    return buffer


def gather_records(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather of per-rank CONTIGUOUS record blocks -> [n_items, 64] on every rank.

    ``local`` is this rank's padded block [capacity, 64] (uint8) whose first
    ``hi - lo`` rows are valid (``shard_range``).
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
    """Run ``search_fn((lo, hi), out_block)`` on this rank's contiguous window range and gather.

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


def search_dealt(search_fn: Callable[[np.ndarray, torch.Tensor], None], n_windows: int, device,
                 group=None, block: int = DEFAULT_BLOCK) -> torch.Tensor:
    """The block-cyclic deal with the collective exchange: ``search_fn(my_windows, buffer)`` fills
    ``buffer[w]`` for every global window ``w`` in ``my_windows``; returns all records on every rank.
    (On a box with peer access use ``PeerGather`` and hand its ``exchange`` to the kernels instead.)"""
    if dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    buf = torch.zeros((n_windows, RECORD_BYTES), dtype=torch.uint8, device=device)
    mine = deal_indices(n_windows, block, world, rank)
    if len(mine):
        search_fn(mine, buf)
    return gather_dealt(buf, group)
