"""Batched tensor API of the window search: drives in HBM -> per-window results.

This is the additive API beneath the reference-shaped facades (``mpc.grid_run``,
``optimize.optimize_trajectory``): every drive of a batch is concatenated into one pose
stream per sensor -- float4 per frame (``stream_dtype=float32``, half the bytes) or double4
(``float64``, no input rounding: what the facades use, the reference being float64
throughout) -- the windows of all drives are planned once (``vmvo_plan_windows``), searched
by the fused kernel (``vmvo_grid_search_f32`` / ``_f64``) and written back
(``vmvo_write_back_f32`` / ``_f64``).  All device work goes through the C ABI of
include/vmvo_b200.h on the current torch stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .constants import MAX_ACCEL, MAX_STEER, MAX_STEER_RATE, STEERING_RATIO, WHEEL_BASE

_ENUMS = {
    "window_mode": {"frames": _lib.WINDOW_FRAMES, "time": _lib.WINDOW_TIME},
    "target_mode": {"time": _lib.TARGET_TIME, "traverse": _lib.TARGET_TRAVERSE},
    "seed_mode": {"data": _lib.SEED_DATA, "given": _lib.SEED_GIVEN, "chained": _lib.SEED_CHAINED},
    "primary": {"vo": _lib.PRIMARY_VO, "gps": _lib.PRIMARY_GPS},
}


@dataclass(frozen=True)
class SearchConfig:
    """The search spec; field meanings as in vmvo_search_cfg."""

    grid_v: int = 32
    grid_s: int = 32
    window_mode: str = "frames"
    window_frames: int = 30
    horizon_time: float = 3.0
    horizon_frames: int = 60
    target_mode: str = "time"
    target_offset: int = 1
    seed_mode: str = "data"
    primary: str = "vo"
    w_vo: float = 1.0
    w_gps: float = 0.0
    w_imu: float = 0.0
    k_steer: float = 0.0
    wheel_base: float = WHEEL_BASE
    steering_ratio: float = STEERING_RATIO
    max_steer: float = MAX_STEER
    max_accel: float = float(MAX_ACCEL)
    max_steer_rate: float = MAX_STEER_RATE
    max_window_poses: int = 0          # 0: derive (frames mode: W+1; time mode: 128)

    def horizon(self) -> int:
        return self.window_frames if self.window_mode == "frames" else self.horizon_frames

    def window_count(self, n_frames: int) -> int:
        return max(0, int(n_frames) - 2 * self.horizon())

    def pose_capacity(self) -> int:
        if self.max_window_poses:
            return int(self.max_window_poses)
        return self.window_frames + 1 if self.window_mode == "frames" else 128

    def to_c(self) -> _lib.SearchCfg:
        c = _lib.SearchCfg()
        for name in ("grid_v", "grid_s", "window_frames", "horizon_frames", "target_offset"):
            setattr(c, name, int(getattr(self, name)))
        for name, table in _ENUMS.items():
            try:
                setattr(c, name, table[getattr(self, name)])
            except KeyError:
                raise ValueError(f"{name}={getattr(self, name)!r}; expected one of {sorted(table)}")
        c.max_window_poses = self.pose_capacity()
        for name in ("horizon_time", "w_vo", "w_gps", "w_imu", "k_steer", "wheel_base",
                     "steering_ratio", "max_steer", "max_accel", "max_steer_rate"):
            setattr(c, name, float(getattr(self, name)))
        return c


def _as_dev(a, dtype, device) -> Optional[torch.Tensor]:
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device, non_blocking=True)


@dataclass
class DriveSet:
    """Drives resident in HBM: concatenated pose streams (all float32 or all float64) +
    float64 stamps."""

    time: torch.Tensor                     # float64 [F]
    vo: Optional[torch.Tensor]             # float32 / float64 [F, 4]
    gps: Optional[torch.Tensor]            # float32 / float64 [F, 4]
    imu: Optional[torch.Tensor]            # float32 / float64 [F]
    drive_offsets: List[int]               # host copy, len D+1
    d_drive_offsets: torch.Tensor          # int64 [D+1]
    dt: torch.Tensor                       # float64 [D]   step length per drive

    @property
    def n_drives(self) -> int:
        return len(self.drive_offsets) - 1

    @property
    def n_frames(self) -> int:
        return self.drive_offsets[-1]

    @property
    def device(self):
        return self.time.device

    @property
    def f64(self) -> bool:
        """True when the pose streams are float64 (the ``_f64`` entry points)."""
        kinds = {t.dtype for t in (self.vo, self.gps, self.imu) if t is not None}
        if len(kinds) > 1 or not kinds <= {torch.float32, torch.float64}:
            raise ValueError(f"pose streams must share one dtype (float32 or float64), got {kinds}")
        return kinds == {torch.float64}

    @staticmethod
    def from_arrays(time: Sequence, dt: Sequence[float], vo: Optional[Sequence] = None,
                    gps: Optional[Sequence] = None, imu: Optional[Sequence] = None,
                    device=None, stream_dtype=np.float32) -> "DriveSet":
        """``time[d]`` float64 [n_d]; ``vo[d]`` / ``gps[d]`` [n_d, 4]; ``imu[d]`` [n_d].

        ``stream_dtype``: float32 (inputs are rounded once) or float64 (as given).
        """
        sd = np.dtype(stream_dtype)
        if sd not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("stream_dtype must be float32 or float64")
        td = torch.float32 if sd == np.dtype(np.float32) else torch.float64
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        lens = [len(t) for t in time]
        offs = [0]
        for n in lens:
            offs.append(offs[-1] + n)

        def cat(parts, dtype, width):
            if parts is None:
                return None
            arrs = []
            for d, p in enumerate(parts):
                a = p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p)
                if a.shape[0] != lens[d] or (width and (a.ndim != 2 or a.shape[1] != width)):
                    raise ValueError(f"drive {d}: stream shape {a.shape} does not match {lens[d]} frames")
                arrs.append(a.astype(dtype, copy=False))
            return np.concatenate(arrs, axis=0) if arrs else np.zeros((0, width) if width else (0,), dtype)

        return DriveSet(
            time=_as_dev(cat(time, np.float64, 0), torch.float64, device),
            vo=_as_dev(cat(vo, sd, 4), td, device),
            gps=_as_dev(cat(gps, sd, 4), td, device),
            imu=_as_dev(cat(imu, sd, 0), td, device),
            drive_offsets=offs,
            d_drive_offsets=torch.tensor(offs, dtype=torch.int64, device=device),
            dt=_as_dev(np.asarray(dt, dtype=np.float64), torch.float64, device),
        )


@dataclass
class WindowPlan:
    window_offsets: List[int]              # host, len D+1
    d_window_offsets: torch.Tensor         # int64 [D+1]
    # the extents (None in a plan without them: frames mode, the search derives them itself)
    win_start: Optional[torch.Tensor]      # int64 [n_windows] absolute frame index
    win_len: Optional[torch.Tensor]        # int32 [n_windows]
    win_drive: Optional[torch.Tensor]      # int32 [n_windows]

    @property
    def n_windows(self) -> int:
        return self.window_offsets[-1]


def plan_windows(cfg: SearchConfig, drives: DriveSet, into: Optional[WindowPlan] = None,
                 extents: bool = True) -> WindowPlan:
    """a8 for every window of every drive (vmvo/schema.py:117-127, …v2.py:48).

    ``into``: a plan allocated by an earlier call for the same drive lengths; only the kernel
    is launched (no allocation, no host-to-device copy), e.g. under CUDA-graph capture.

    ``extents=False`` (frames mode only): just the window counts per drive -- no kernel, no
    extent arrays; ``grid_search`` then lets the search kernel derive each window's extent itself
    (window i of a drive = poses i .. i + window_frames), which saves the planning launch of a step.
    """
    ctx = _lib.context(drives.device.index)
    c = cfg.to_c()
    if not extents:
        if cfg.window_mode != "frames":
            raise ValueError("a plan without extents is for window_mode 'frames'")
        woffs = [0]
        for d in range(drives.n_drives):
            woffs.append(woffs[-1] + cfg.window_count(drives.drive_offsets[d + 1] - drives.drive_offsets[d]))
        return WindowPlan(window_offsets=woffs,
                          d_window_offsets=torch.tensor(woffs, dtype=torch.int64, device=drives.device),
                          win_start=None, win_len=None, win_drive=None)
    if into is not None:
        plan, n = into, into.n_windows
        if n and plan.win_start is not None:
            ctx.check(ctx.lib.vmvo_plan_windows(
                ctx.handle, C.byref(c), drives.n_drives, _lib.ptr(drives.d_drive_offsets),
                _lib.ptr(plan.d_window_offsets), n, _lib.ptr(drives.time), _lib.ptr(plan.win_start),
                _lib.ptr(plan.win_len), _lib.ptr(plan.win_drive), _lib.stream_ptr(drives.device)),
                "vmvo_plan_windows")
        return plan
    woffs = [0]
    for d in range(drives.n_drives):
        woffs.append(woffs[-1] + cfg.window_count(drives.drive_offsets[d + 1] - drives.drive_offsets[d]))
    n = woffs[-1]
    dev = drives.device
    plan = WindowPlan(
        window_offsets=woffs,
        d_window_offsets=torch.tensor(woffs, dtype=torch.int64, device=dev),
        win_start=torch.empty(n, dtype=torch.int64, device=dev),
        win_len=torch.empty(n, dtype=torch.int32, device=dev),
        win_drive=torch.empty(n, dtype=torch.int32, device=dev),
    )
    if n:
        ctx.check(ctx.lib.vmvo_plan_windows(
            ctx.handle, C.byref(c), drives.n_drives, _lib.ptr(drives.d_drive_offsets),
            _lib.ptr(plan.d_window_offsets), n, _lib.ptr(drives.time), _lib.ptr(plan.win_start),
            _lib.ptr(plan.win_len), _lib.ptr(plan.win_drive), _lib.stream_ptr(dev)), "vmvo_plan_windows")
    return plan


@dataclass
class SearchOutput:
    results: torch.Tensor                          # uint8 [n, 64]: vmvo_window_result records
    poses: Optional[torch.Tensor] = None           # float64 [n, stride, 3]
    steer: Optional[torch.Tensor] = None           # float64 [n, stride]
    vel: Optional[torch.Tensor] = None             # float64 [n, stride]

    def records(self) -> np.ndarray:
        """Host copy as a NumPy record array (synchronises)."""
        return self.results.cpu().numpy().view(_lib.RESULT_DTYPE).reshape(-1)


def grid_search(cfg: SearchConfig, drives: DriveSet, plan: WindowPlan,
                window_range: Optional[Tuple[int, int]] = None, seeds: Optional[torch.Tensor] = None,
                want_rollouts: bool = False, out: Optional[torch.Tensor] = None,
                exchange: Optional[_lib.Exchange] = None) -> SearchOutput:
    """The fused search over windows [lo, hi) of the plan (default: all).

    ``out``: optional uint8 [n, 64] tensor the records are written into (e.g. this rank's
    slice of the gather buffer, scheduler.py).

    ``exchange`` (``scheduler.PeerGather.exchange``): the multi-GPU form -- this rank searches the
    windows the block-cyclic deal gives it, ``out`` must be the gather buffer (one record per
    window of the WHOLE plan) and every record is also stored into the peers' buffers
    (``vmvo_grid_search_sharded``).
    """
    ctx = _lib.context(drives.device.index)
    c = cfg.to_c()
    if exchange is not None or plan.win_start is None:
        if window_range is not None or want_rollouts:
            raise ValueError("a sharded search / a search without planned extents covers the whole plan "
                             "and returns records only")
        n, dev = plan.n_windows, drives.device
        if out is None and exchange is None:
            out = torch.empty((n, 64), dtype=torch.uint8, device=dev)
        if out is None or out.shape != (n, 64) or out.dtype != torch.uint8 or not out.is_contiguous():
            raise ValueError("a sharded search writes into the gather buffer: out = uint8 [n_windows, 64]")
        d_seeds = None
        if seeds is not None:
            d_seeds = _as_dev(seeds, torch.float64, dev)
            if d_seeds.shape != (n, 2):
                raise ValueError("seeds must be [n_windows, 2]")
        if n:
            ctx.check(ctx.lib.vmvo_grid_search_sharded(
                ctx.handle, C.byref(c), drives.n_drives, _lib.ptr(drives.d_drive_offsets),
                _lib.ptr(plan.d_window_offsets), n, _lib.ptr(plan.win_start), _lib.ptr(plan.win_len),
                _lib.ptr(plan.win_drive), _lib.ptr(drives.dt), _lib.ptr(drives.vo), _lib.ptr(drives.gps),
                _lib.ptr(drives.imu), int(drives.f64), _lib.ptr(d_seeds), _lib.ptr(out),
                None if exchange is None else C.byref(exchange), _lib.stream_ptr(dev)),
                "vmvo_grid_search_sharded")
        return SearchOutput(results=out)
    lo, hi = (0, plan.n_windows) if window_range is None else window_range
    n = hi - lo
    dev = drives.device
    results = out if out is not None else torch.empty((n, 64), dtype=torch.uint8, device=dev)
    if results.shape != (n, 64) or results.dtype != torch.uint8 or not results.is_contiguous():
        raise ValueError("out must be a contiguous uint8 [n_windows, 64] tensor")
    so = SearchOutput(results=results)
    stride = cfg.pose_capacity() - 1
    if want_rollouts:
        so.poses = torch.full((n, stride, 3), float("nan"), dtype=torch.float64, device=dev)
        so.steer = torch.full((n, stride), float("nan"), dtype=torch.float64, device=dev)
        so.vel = torch.full((n, stride), float("nan"), dtype=torch.float64, device=dev)
    d_seeds = None
    if seeds is not None:
        d_seeds = _as_dev(seeds, torch.float64, dev)
        if d_seeds.shape != (plan.n_windows, 2):
            raise ValueError("seeds must be [n_windows, 2]")
        d_seeds = d_seeds[lo:hi].contiguous()
    if n and cfg.seed_mode == "chained":
        # runs = the drives whose windows fall in [lo, hi); the range must not cut a drive
        offs = [o for o in plan.window_offsets if lo <= o <= hi]
        if not offs or offs[0] != lo or offs[-1] != hi:
            raise ValueError("seed_mode 'chained' needs a window range aligned to whole drives")
        if lo == 0 and hi == plan.n_windows:
            runs = plan.d_window_offsets          # no allocation: usable under graph capture
        else:
            runs = torch.tensor([o - lo for o in offs], dtype=torch.int64, device=dev)
        ctx.check(ctx.lib.vmvo_grid_search_chained(
            ctx.handle, C.byref(c), n, _lib.ptr(plan.win_start[lo:hi]), _lib.ptr(plan.win_len[lo:hi]),
            _lib.ptr(plan.win_drive[lo:hi]), _lib.ptr(drives.dt), _lib.ptr(drives.vo),
            _lib.ptr(drives.gps), _lib.ptr(drives.imu), int(drives.f64), len(offs) - 1, _lib.ptr(runs),
            _lib.ptr(results), _lib.ptr(so.poses), _lib.ptr(so.steer), _lib.ptr(so.vel), stride,
            _lib.stream_ptr(dev)), "vmvo_grid_search_chained")
    elif n:
        name = "vmvo_grid_search_f64" if drives.f64 else "vmvo_grid_search_f32"
        ctx.check(getattr(ctx.lib, name)(
            ctx.handle, C.byref(c), n, _lib.ptr(plan.win_start[lo:hi]), _lib.ptr(plan.win_len[lo:hi]),
            _lib.ptr(plan.win_drive[lo:hi]), _lib.ptr(drives.dt), _lib.ptr(drives.vo),
            _lib.ptr(drives.gps), _lib.ptr(drives.imu), _lib.ptr(d_seeds), _lib.ptr(results),
            _lib.ptr(so.poses), _lib.ptr(so.steer), _lib.ptr(so.vel), stride,
            _lib.stream_ptr(dev)), name)
    return so


def grid_search_debug(cfg: SearchConfig, drives: DriveSet, plan: WindowPlan,
                      seeds: Optional[torch.Tensor] = None):
    """Test hook (``vmvo_grid_search_debug_f32``): records plus, for EVERY hypothesis, the FP32
    scan cost and the width of its error band, each float32 [n_windows, grid_v * grid_s]."""
    ctx = _lib.context(drives.device.index)
    c = cfg.to_c()
    n, dev = plan.n_windows, drives.device
    if drives.f64:
        raise ValueError("the debug export is built for float32 streams")
    results = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    cost = torch.full((n, cfg.grid_v * cfg.grid_s), float("nan"), dtype=torch.float32, device=dev)
    err = torch.full_like(cost, float("nan"))
    d_seeds = None if seeds is None else _as_dev(seeds, torch.float64, dev)
    ctx.check(ctx.lib.vmvo_grid_search_debug_f32(
        ctx.handle, C.byref(c), n, _lib.ptr(plan.win_start), _lib.ptr(plan.win_len),
        _lib.ptr(plan.win_drive), _lib.ptr(drives.dt), _lib.ptr(drives.vo), _lib.ptr(drives.gps),
        _lib.ptr(drives.imu), _lib.ptr(d_seeds), _lib.ptr(results), _lib.ptr(cost), _lib.ptr(err),
        _lib.stream_ptr(dev)), "vmvo_grid_search_debug_f32")
    return SearchOutput(results=results), cost, err


def write_back(cfg: SearchConfig, drives: DriveSet, plan: WindowPlan, results: torch.Tensor,
               blend_gps: bool = True, out: Optional[torch.Tensor] = None,
               frame_range: Optional[Tuple[int, int]] = None,
               exchange: Optional[_lib.Exchange] = None) -> torch.Tensor:
    """a12: float64 [4, F] = x, y, theta, velocity (optimize_trajectory_v2.py:32-33,122-137).

    ``frame_range`` = (lo, hi): only those frames of ``out`` are written (a rank's share of the
    trajectory); ``exchange``: the kernel publishes this rank's arrival word and waits for the
    peers' before reading ``results``, the gather buffer of the sharded search."""
    ctx = _lib.context(drives.device.index)
    c = cfg.to_c()
    if drives.vo is None:
        raise ValueError("write_back copies the VO stream: drives.vo is required")
    if results.shape != (plan.n_windows, 64):
        raise ValueError("results must hold one record per planned window")
    if out is None:
        out = torch.empty((4, drives.n_frames), dtype=torch.float64, device=drives.device)
    elif out.shape != (4, drives.n_frames) or out.dtype != torch.float64 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 [4, n_frames] tensor")
    gps = drives.gps if blend_gps else None
    if frame_range is not None or exchange is not None:
        lo, hi = (0, drives.n_frames) if frame_range is None else frame_range
        ctx.check(ctx.lib.vmvo_write_back_range(
            ctx.handle, C.byref(c), drives.n_drives, drives.n_frames, int(lo), int(hi),
            _lib.ptr(drives.d_drive_offsets), _lib.ptr(plan.d_window_offsets), _lib.ptr(drives.dt),
            _lib.ptr(drives.vo), _lib.ptr(gps), int(drives.f64), _lib.ptr(results), _lib.ptr(out[0]),
            _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]),
            None if exchange is None else C.byref(exchange), _lib.stream_ptr(drives.device)),
            "vmvo_write_back_range")
        return out
    name = "vmvo_write_back_f64" if drives.f64 else "vmvo_write_back_f32"
    ctx.check(getattr(ctx.lib, name)(
        ctx.handle, C.byref(c), drives.n_drives, drives.n_frames, _lib.ptr(drives.d_drive_offsets),
        _lib.ptr(plan.d_window_offsets), _lib.ptr(drives.dt), _lib.ptr(drives.vo), _lib.ptr(gps),
        _lib.ptr(results), _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(out[3]),
        _lib.stream_ptr(drives.device)), name)
    return out


def optimize_drives(cfg: SearchConfig, drives: DriveSet, plan: Optional[WindowPlan] = None,
                    seeds: Optional[torch.Tensor] = None) -> Tuple[SearchOutput, torch.Tensor, WindowPlan]:
    """plan -> search -> write-back for a batch of drives on the current device."""
    if plan is None:
        plan = plan_windows(cfg, drives)
    so = grid_search(cfg, drives, plan, seeds=seeds)
    traj = write_back(cfg, drives, plan, so.results)
    return so, traj, plan


class DrivePipeline:
    """plan -> search -> write-back for one resident batch of drives, captured once as a CUDA
    graph and replayed: on a small grid in frames mode four kernel launches (window preparation,
    search, second kernel, write-back: the extents are derived in the kernels), one more in time mode
    (the plan first), two small memsets, no per-pass host work (``kernels_per_pass``).

    The pose streams and stamps are read from ``drives`` at replay time, so new data of the
    same shape can be copied into ``drives.vo`` / ``.gps`` / ``.imu`` / ``.time`` between passes.
    ``records`` (uint8 [n_windows, 64]) may be a slice of a larger gather buffer.

    ``gather`` (a ``scheduler.PeerGather`` over the plan's windows): the multi-GPU form.  Every rank
    builds the same pipeline over the same drives; a pass searches this rank's share of the global
    window list, the records reach every rank from inside the search, and the write-back (of
    ``frame_range``, this rank's share of the frames; default all) consumes ALL ranks' records
    behind the arrival words -- still four launches and no collective.  Construction and every
    ``run`` are collective over the gather's group (each pass waits for the peers' records).
    """

    KERNELS_PER_PASS = 4       # (time mode; see kernels_per_pass)

    def __init__(self, cfg: SearchConfig, drives: DriveSet, blend_gps: bool = True,
                 records: Optional[torch.Tensor] = None, use_graph: bool = True,
                 split: bool = False, gather=None, frame_range: Optional[Tuple[int, int]] = None,
                 record_range: Optional[Tuple[int, int]] = None):
        if cfg.seed_mode == "given":
            raise ValueError("DrivePipeline derives seeds from the data (seed_mode data / chained)")
        self.cfg, self.drives, self.blend_gps = cfg, drives, blend_gps
        self.plan = plan_windows(cfg, drives, extents=cfg.window_mode != "frames")
        dev = drives.device
        n = self.plan.n_windows
        self.gather, self.frame_range = gather, frame_range
        # what a streaming caller reads back per pass (DriveStream): the frames this pipeline writes
        # and the records of ``record_range`` (default: all)
        self.record_range = record_range if record_range is not None else (0, n)
        if gather is not None:
            if records is not None or split:
                raise ValueError("with a gather the records live in its buffer and the pass is one graph")
            if gather.n != n:
                raise ValueError(f"the gather holds {gather.n} records, the plan has {n} windows")
            records = gather.buffer
        self.records = records if records is not None else torch.empty((n, 64), dtype=torch.uint8, device=dev)
        self.trajectory = torch.empty((4, drives.n_frames), dtype=torch.float64, device=dev)
        self.graphs = None
        ctx = _lib.context(dev.index)
        before = ctx.launch_count()
        self._search()                                # eager warm-up: sets kernel attributes
        self._write_back()
        torch.cuda.synchronize(dev)
        # kernels of one pass, as the library counted them: the preparation pass (small grids), the
        # plan (time mode), the search, the second kernel (where windows may be parked), the write-back
        self.kernels_per_pass = ctx.launch_count() - before
        if use_graph:
            # split: the search and the write-back are separate graphs so that a caller can put
            # the record gather between them (bench.py, N > 1)
            parts = [[self._search], [self._write_back]] if split else [[self._search, self._write_back]]
            self.graphs = []
            for fns in parts:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for fn in fns:
                        fn()
                self.graphs.append(g)
        self.split = split

    def _search(self):
        plan_windows(self.cfg, self.drives, into=self.plan)
        grid_search(self.cfg, self.drives, self.plan, out=self.records,
                    exchange=None if self.gather is None else self.gather.exchange)

    def _write_back(self):
        write_back(self.cfg, self.drives, self.plan, self.records, blend_gps=self.blend_gps,
                   out=self.trajectory, frame_range=self.frame_range,
                   exchange=None if self.gather is None else self.gather.exchange)

    def run_search(self) -> torch.Tensor:
        if self.graphs is None:
            self._search()
        elif self.split:
            self.graphs[0].replay()
        else:
            raise RuntimeError("run_search needs split=True (or use_graph=False)")
        return self.records

    def run_write_back(self) -> torch.Tensor:
        if self.graphs is None:
            self._write_back()
        elif self.split:
            self.graphs[1].replay()
        else:
            raise RuntimeError("run_write_back needs split=True (or use_graph=False)")
        return self.trajectory

    def run(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """One pass on the current stream; returns (records, trajectory) -- reused buffers."""
        if self.graphs is None:
            self._search()
            self._write_back()
        else:
            for g in self.graphs:
                g.replay()
        return self.records, self.trajectory

    def result_records(self) -> np.ndarray:
        return self.records.cpu().numpy().view(_lib.RESULT_DTYPE).reshape(-1)


class DriveStream:
    """Streams same-shaped batches of drives from HOST memory through plan -> search -> write-back
    and back, with the copies hidden behind the search: two resident drive buffers and three
    streams.  One ``step`` carries the host-to-device copy of the NEXT batch, the compute of the
    CURRENT one and the device-to-host copy of the PREVIOUS one; they start together and are
    joined before ``step`` returns to the caller's stream.

    ``inputs`` are dicts of pinned host tensors keyed like ``DriveSet`` (``vo``, ``gps``, ``imu``,
    ``time``), shaped like ``template``'s.  ``run`` wraps the stepping for an iterable of batches.

    ``pipe_factory(b, drives_b)`` may build the pipeline of buffer ``b`` itself -- e.g. with its
    records in a gather buffer and the result mirrors of ``scheduler.PeerGather`` in force while the
    graph is captured (the multi-GPU form; bench.py, N > 1).
    """

    def __init__(self, cfg: SearchConfig, template: DriveSet, blend_gps: bool = True, pipe_factory=None):
        dev = template.device
        self.dev = dev
        self.sets: List[DriveSet] = []
        self.pipes: List[DrivePipeline] = []
        self.host: List[Tuple[torch.Tensor, torch.Tensor]] = []
        for _ in range(2):
            d = DriveSet(time=template.time.clone(),
                         vo=None if template.vo is None else template.vo.clone(),
                         gps=None if template.gps is None else template.gps.clone(),
                         imu=None if template.imu is None else template.imu.clone(),
                         drive_offsets=list(template.drive_offsets),
                         d_drive_offsets=template.d_drive_offsets, dt=template.dt)
            pipe = (pipe_factory(len(self.pipes), d) if pipe_factory is not None
                    else DrivePipeline(cfg, d, blend_gps=blend_gps))
            self.sets.append(d)
            self.pipes.append(pipe)
            (rlo, rhi), (flo, fhi) = pipe.record_range, pipe.frame_range or (0, d.n_frames)
            self.host.append((torch.empty((rhi - rlo, 64), dtype=torch.uint8).pin_memory(),
                              torch.empty((4, fhi - flo), dtype=torch.float64).pin_memory()))
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.n = 0                      # steps computed so far
        self.loaded = False             # inputs of step self.n are resident

    def _h2d(self, b: int, inputs) -> None:
        d = self.sets[b]
        for name in ("vo", "gps", "imu", "time"):
            src = inputs.get(name)
            if src is not None:
                getattr(d, name).copy_(src, non_blocking=True)

    def _d2h(self, b: int) -> None:
        pipe = self.pipes[b]
        (rlo, rhi), fr = pipe.record_range, pipe.frame_range
        self.host[b][0].copy_(pipe.records[rlo:rhi], non_blocking=True)
        if fr is None:
            self.host[b][1].copy_(pipe.trajectory, non_blocking=True)
        else:                      # this rank's share of the frames: four contiguous row pieces
            for r in range(4):
                self.host[b][1][r].copy_(pipe.trajectory[r, fr[0]:fr[1]], non_blocking=True)

    def d2h_bytes(self) -> int:
        return int(self.host[0][0].numel() + self.host[0][1].numel() * 8)

    def prime(self, inputs) -> None:
        """Copy the first batch in (nothing to overlap it with yet)."""
        self._h2d(self.n & 1, inputs)
        self.loaded = True

    def step(self, next_inputs=None) -> Optional[int]:
        """Compute the resident batch while ``next_inputs`` go in and the previous results come
        out.  Returns the buffer index whose host copies (``host[b]``) now hold the PREVIOUS
        step's results, or None on the first step."""
        if not self.loaded:
            raise RuntimeError("prime() the stream with the first batch")
        b = self.n & 1
        main = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(main)
        self.s_in.wait_event(start)
        self.s_out.wait_event(start)
        if next_inputs is not None:
            with torch.cuda.stream(self.s_in):
                self._h2d(b ^ 1, next_inputs)
        if self.n > 0:
            with torch.cuda.stream(self.s_out):
                self._d2h(b ^ 1)
        self.pipes[b].run()
        main.wait_stream(self.s_in)
        main.wait_stream(self.s_out)
        self.loaded = next_inputs is not None
        prev = (b ^ 1) if self.n > 0 else None
        self.n += 1
        return prev

    def drain(self) -> int:
        """Read the last computed step's results back; returns its buffer index."""
        b = (self.n - 1) & 1
        self._d2h(b)
        return b

    def run(self, batches):
        """Yields ``(records, trajectory)`` NumPy copies per batch, in order."""
        it = iter(batches)
        try:
            first = next(it)
        except StopIteration:
            return
        self.prime(first)
        pending = next(it, None)
        while True:
            nxt = pending
            prev = self.step(nxt)
            if prev is not None:
                torch.cuda.current_stream(self.dev).synchronize()
                yield (self.host[prev][0].numpy().view(_lib.RESULT_DTYPE).reshape(-1).copy(),
                       self.host[prev][1].numpy().copy())
            if nxt is None:
                break
            pending = next(it, None)
        b = self.drain()
        torch.cuda.current_stream(self.dev).synchronize()
        yield (self.host[b][0].numpy().view(_lib.RESULT_DTYPE).reshape(-1).copy(),
               self.host[b][1].numpy().copy())


def executed_mufu_per_hypothesis_step(cfg: SearchConfig) -> float:
    """SFU operations the search kernel issues per hypothesis-step: 2 in the generic
    scan (sin + cos; tan is hoisted into a table), 1.25 in the packed / rotation scan, which
    grid_search_impl (csrc/vmvo_search.cu) selects for grids of >= 128 thread-items without an
    IMU term."""
    items = -(-int(cfg.grid_v) // 8) * int(cfg.grid_s)
    return 1.25 if (items >= 128 and cfg.w_imu == 0) else 2.0


def hypothesis_steps(cfg: SearchConfig, records: np.ndarray) -> int:
    """Sum over windows of G_v * G_s * N_w: the unit of the throughput metric."""
    return int(cfg.grid_v) * int(cfg.grid_s) * int(records["n_steps"].astype(np.int64).sum())
