"""Sliding-window trajectory optimizer, signature-compatible with
vmvo/scripts/optimize_trajectory_v2.py:24-148.

``optimize_trajectory(vo_trajectory, gps_trajectory, model)`` keeps the reference's call
shape, objective and write-back semantics (quirks D4-D7, D10 of DESIGN.md); the per-window SLSQP
solve is replaced by the hypothesis-grid argmin, and windows, search, write-back and blends all
run on the GPU.  The default configuration is ``REFERENCE_CFG`` -- the reference's own objective:
the GPS sub-trajectory as target, GPS velocity seeds, arc-length decimation (mpc.py:35-40,
optimize_trajectory_v2.py:57-74).  ``VO_CFG`` (scoring against the raw VO window, the search spec
of BASELINE's north star) is an explicit opt-in through ``config=``.
"""
from __future__ import annotations

from dataclasses import replace
from typing import Optional

import numpy as np

from .bicycle_model import BicycleModel
from .schema import Trajectory
from .search import DriveSet, SearchConfig, optimize_drives

HORIZON_TIME = 3.0  # s, optimize_trajectory_v2.py:35

VO_CFG = SearchConfig(window_mode="time", target_mode="time", primary="vo",
                      w_vo=1.0, w_gps=0.0, w_imu=0.0)
DEFAULT_CFG = VO_CFG      # (round-1 name of the VO-scoring configuration; NOT the facade's default)
REFERENCE_CFG = SearchConfig(window_mode="time", target_mode="traverse", primary="gps",
                             w_vo=0.0, w_gps=1.0, w_imu=0.0)


def _stream(traj: Trajectory, n: int) -> np.ndarray:
    """[n, 4] float64 (x, y, theta, v); a short theta column (quirk D8) is padded."""
    cols = []
    for name in ("x", "y", "theta", "velocity"):
        c = np.asarray(getattr(traj, name), dtype=np.float64)[:n]
        if len(c) < n:
            c = np.concatenate([c, np.full(n - len(c), c[-1] if len(c) else 0.0)])
        cols.append(c)
    return np.stack(cols, axis=1)


def optimize_trajectory(vo_trajectory: Trajectory, gps_trajectory: Trajectory,
                        model: Optional[BicycleModel] = None,
                        config: Optional[SearchConfig] = None, imu_yaw=None,
                        return_details: bool = False, honour_model_limits: bool = False):
    """Returns the optimised ``Trajectory`` (a modified copy of ``vo_trajectory``).

    Like the reference, ``model`` is not consulted (quirk D10, optimize_trajectory_v2.py:25,44: a
    fresh ``BicycleModel()`` with the module constants is used) unless ``honour_model_limits``
    asks for its steering / acceleration / steering-rate limits to bound the grid."""
    cfg = config if config is not None else REFERENCE_CFG
    N = min(len(vo_trajectory), len(gps_trajectory))
    gps_time = np.asarray(gps_trajectory.time, dtype=np.float64)
    # optimize_trajectory_v2.py:35-42
    FPS = 1 / np.mean(np.diff(gps_time))
    horizon = int(HORIZON_TIME * FPS)
    dt = 1.0 / FPS
    if cfg.window_mode == "time":
        cfg = replace(cfg, horizon_time=HORIZON_TIME, horizon_frames=horizon)
    if model is not None and honour_model_limits:
        cfg = replace(cfg, max_steer=float(model.max_steer), max_accel=float(model.max_accel),
                      max_steer_rate=float(model.max_steer_rate))

    out = Trajectory(**dict(vo_trajectory))
    n_windows = cfg.window_count(N)
    if n_windows <= 0:
        return (out, None) if return_details else out

    vo = _stream(vo_trajectory, N)
    gps = _stream(gps_trajectory, N)
    imu = None if imu_yaw is None else [np.asarray(imu_yaw, dtype=np.float64)[:N]]
    # float64 streams: nothing is rounded on the way in (the reference is float64 throughout)
    drives = DriveSet.from_arrays([gps_time[:N]], [dt], vo=[vo], gps=[gps], imu=imu,
                                  stream_dtype=np.float64)
    so, traj, plan = optimize_drives(cfg, drives)
    rec = so.records()
    if np.any(rec["status"] & 8):
        raise AssertionError("No frames found")                  # schema.py:122
    if np.any(rec["status"] & 4):
        raise ValueError("a window holds more poses than max_window_poses; raise it in the config")
    if np.any(rec["n_steps"] == 0):
        # the reference dies on steering_angles[-1] of an empty solve (quirk D7, …v2.py:146)
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    vo_time = np.asarray(vo_trajectory.time, dtype=np.float64)[:n_windows]
    ok = np.isclose(vo_time, vo_time, atol=0.1)
    if not np.all(ok):
        i = int(np.argmin(ok))
        raise AssertionError(f"Time mismatch: {vo_time[i]} != {vo_time[i]}")

    res = traj.cpu().numpy()
    # x, y: every frame covered by some window; theta, velocity: the window starts
    diff = np.zeros(N + 1, dtype=np.int64)
    idx = np.arange(n_windows)
    np.add.at(diff, idx, 1)
    np.add.at(diff, np.minimum(idx + rec["n_steps"], N), -1)
    covered = np.cumsum(diff[:N]) > 0
    x = np.asarray(out.x, dtype=np.float64)
    y = np.asarray(out.y, dtype=np.float64)
    th = np.asarray(out.theta, dtype=np.float64)
    v = np.asarray(out.velocity, dtype=np.float64)
    x[:N][covered] = res[0][covered]
    y[:N][covered] = res[1][covered]
    th[:n_windows] = res[2][:n_windows]
    v[:n_windows] = res[3][:n_windows]
    out.x, out.y, out.theta, out.velocity = x.tolist(), y.tolist(), th.tolist(), v.tolist()
    return (out, (so, plan, rec)) if return_details else out
