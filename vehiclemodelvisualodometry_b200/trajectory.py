"""Trajectory pre-processing (the step before the optimizer), interface-compatible with
vmvo/utils/trajectory.py:13-99,177-335 of the reference: DataFrame in, ``Trajectory`` out.

``process_vo_trajectory`` / ``process_gps_trajectory`` / ``smoothen_traj`` keep the reference's
signatures; the arithmetic runs on the GPU (``vmvo_vo_prepare_f64`` / ``vmvo_gps_prepare_f64`` /
``vmvo_smooth_f64``), batched over drives in the tensor API below.  Plotting and drawing helpers of
the reference file are out of scope.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .schema import Trajectory


def _offsets(lengths: Sequence[int], dev) -> Tuple[List[int], torch.Tensor]:
    offs = [0]
    for n in lengths:
        offs.append(offs[-1] + int(n))
    return offs, torch.tensor(offs, dtype=torch.int64, device=dev)


def _cat(parts, dev, width=0) -> torch.Tensor:
    a = np.concatenate([np.asarray(p, dtype=np.float64).reshape((-1, width) if width else (-1,))
                        for p in parts], axis=0)
    return torch.as_tensor(np.ascontiguousarray(a)).to(dev)


def _device():
    _lib.context()          # raises without a GPU: there is no CPU fallback
    return torch.device("cuda", torch.cuda.current_device())


def smooth_batch(x: Sequence, y: Sequence, window: int):
    """Trailing moving average of every drive; returns per-drive (x, y) NumPy arrays."""
    dev = _device()
    ctx = _lib.context(dev.index)
    offs, d_off = _offsets([len(a) for a in x], dev)
    dx, dy = _cat(x, dev), _cat(y, dev)
    ox, oy = torch.empty_like(dx), torch.empty_like(dy)
    ctx.check(ctx.lib.vmvo_smooth_f64(ctx.handle, len(offs) - 1, offs[-1], _lib.ptr(d_off), _lib.ptr(dx),
                                      _lib.ptr(dy), int(window), _lib.ptr(ox), _lib.ptr(oy),
                                      _lib.stream_ptr(dev)), "vmvo_smooth_f64")
    ox, oy = ox.cpu().numpy(), oy.cpu().numpy()
    return [(ox[a:b], oy[a:b]) for a, b in zip(offs, offs[1:])]


def smoothen_traj(trajectory, window_size=3):
    """Moving-average smoothing of an [n, 2] trajectory (vmvo/utils/trajectory.py:68-99)."""
    traj = np.asarray(trajectory, dtype=np.float64)
    if len(traj) <= window_size:
        return trajectory
    (sx, sy), = smooth_batch([traj[:, 0]], [traj[:, 1]], window_size)
    return np.stack([sx, sy], axis=1)


def vo_prepare_device(d_off: torch.Tensor, n_drives: int, total: int, dx, dy, drot, dstamp,
                      scale: float = 0.25, window: int = 20, out: Optional[torch.Tensor] = None,
                      yaw_f32: bool = False):
    """Device-resident process_vo_trajectory: float64 CUDA tensors in, float64 [5, total] out
    (rows x, y, theta, velocity, time).  ``yaw_f32``: the rotation entries are float32 values
    (the cached trajectory) and the yaw is a float32 atan2, as in the reference."""
    dev = dx.device
    ctx = _lib.context(dev.index)
    if out is None:
        out = torch.empty((5, total), dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.vmvo_vo_prepare_f64(
        ctx.handle, n_drives, total, _lib.ptr(d_off), _lib.ptr(dx), _lib.ptr(dy), _lib.ptr(drot),
        _lib.ptr(dstamp), float(scale), int(window), int(bool(yaw_f32)), _lib.ptr(out[0]), _lib.ptr(out[1]),
        _lib.ptr(out[2]),
        _lib.ptr(out[3]), _lib.ptr(out[4]), _lib.stream_ptr(dev)), "vmvo_vo_prepare_f64")
    return out


def gps_prepare_device(d_off: torch.Tensor, n_drives: int, total: int, dlat, dlon, dspeed, dstamp,
                       window: int = 20, out: Optional[torch.Tensor] = None,
                       scratch: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None):
    """Device-resident process_gps_trajectory: float64 [5, total + n_drives] out (drive d starts at
    offset[d] + d) and int32 status [n_drives]."""
    dev = dlat.device
    ctx = _lib.context(dev.index)
    if out is None:
        out = torch.empty((5, total + n_drives), dtype=torch.float64, device=dev)
    if status is None:
        status = torch.zeros(n_drives, dtype=torch.int32, device=dev)
    if scratch is None:
        scratch = torch.empty(int(ctx.lib.vmvo_gps_prepare_scratch_bytes(total, n_drives)),
                              dtype=torch.uint8, device=dev)
    ctx.check(ctx.lib.vmvo_gps_prepare_f64(
        ctx.handle, n_drives, total, _lib.ptr(d_off), _lib.ptr(dlat), _lib.ptr(dlon), _lib.ptr(dspeed),
        _lib.ptr(dstamp), int(window), _lib.ptr(scratch), _lib.ptr(out[0]), _lib.ptr(out[1]),
        _lib.ptr(out[2]), _lib.ptr(out[3]), _lib.ptr(out[4]), _lib.ptr(status), _lib.stream_ptr(dev)),
        "vmvo_gps_prepare_f64")
    return out, status, scratch


def vo_prepare_batch(x, y, rot, stamp_ms, scale: float = 0.25, window: int = 20, yaw_f32: bool = False):
    """Batched process_vo_trajectory: lists of per-drive arrays -> list of dicts of columns."""
    dev = _device()
    offs, d_off = _offsets([len(a) for a in x], dev)
    dx, dy, dr, dt = _cat(x, dev), _cat(y, dev), _cat(rot, dev, 9), _cat(stamp_ms, dev)
    out = vo_prepare_device(d_off, len(offs) - 1, offs[-1], dx, dy, dr, dt, scale, window, yaw_f32=yaw_f32)
    o = out.cpu().numpy()
    names = ("x", "y", "theta", "velocity", "time")
    return [{k: o[c, a:b] for c, k in enumerate(names)} for a, b in zip(offs, offs[1:])]


def gps_prepare_batch(lat, lon, speed, stamp_ms, window: int = 20):
    """Batched process_gps_trajectory; a drive of n fixes yields n + 1 points and n headings."""
    dev = _device()
    lens = [len(a) for a in lat]
    offs, d_off = _offsets(lens, dev)
    D, F = len(lens), offs[-1]
    dla, dlo, dsp, dst = _cat(lat, dev), _cat(lon, dev), _cat(speed, dev), _cat(stamp_ms, dev)
    out, status, _ = gps_prepare_device(d_off, D, F, dla, dlo, dsp, dst, window)
    o, st = out.cpu().numpy(), status.cpu().numpy()
    res = []
    for d in range(D):
        a, n = offs[d] + d, lens[d]
        if st[d]:
            # the reference indexes velocity[n] when the log ends on a fresh fix
            raise IndexError(f"index {n} is out of bounds for axis 0 with size {n}")
        res.append({"x": o[0, a:a + n + 1], "y": o[1, a:a + n + 1], "theta": o[2, a:a + n],
                    "velocity": o[3, a:a + n + 1], "time": o[4, a:a + n + 1]})
    return res


def process_vo_trajectory(trajectory, scale: float = 0.25, smoothen_window: int = 20) -> Trajectory:
    """DataFrame with columns x, y, rot (3x3 per row), Timestamp [ms] -> Trajectory
    (vmvo/utils/trajectory.py:13-65)."""
    rots = [np.asarray(r) for r in trajectory["rot"].tolist()]
    # the cached trajectory holds float32 matrices (bdd_raw.py:163-164): np.arctan2 then works,
    # and returns, in float32 (trajectory.py:28)
    yaw_f32 = len(rots) > 0 and all(r.dtype == np.float32 for r in rots)
    rot = np.stack([r.astype(np.float64) for r in rots])
    (c,) = vo_prepare_batch([np.asarray(trajectory["x"])], [np.asarray(trajectory["y"])], [rot],
                            [np.asarray(trajectory["Timestamp"].tolist())], scale, smoothen_window,
                            yaw_f32=yaw_f32)
    return Trajectory(**{k: v for k, v in c.items()})


def process_gps_trajectory(trajectory, heading_num_frames: int = 25,
                           smoothen_window: int = 20) -> Trajectory:
    """DataFrame with columns heading, Latitude, Longitude, speed, Timestamp [ms] -> Trajectory
    (vmvo/utils/trajectory.py:177-335).  ``heading`` only feeds a value the reference never
    returns; it is read here for the same KeyError behaviour and otherwise unused."""
    _ = [trajectory["heading"][i] for i in range(heading_num_frames)]
    (c,) = gps_prepare_batch([np.asarray(trajectory["Latitude"])], [np.asarray(trajectory["Longitude"])],
                             [np.asarray(trajectory["speed"].tolist())],
                             [np.asarray(trajectory["Timestamp"].tolist())], smoothen_window)
    return Trajectory(**{k: v for k, v in c.items()})
