"""Synthetic drives shaped like the Bengaluru Driving Dataset sequences.

The dataset is not available offline; SURVEY.md section 8(d) records the shape facts the
reference does state: 20 Hz log with the GPS fix refreshed at 10 Hz
(vmvo/utils/trajectory.py:220-223), millisecond epoch stamps (…:64,228), VO smoothed with a
trailing 20-tap average (…:15-16,60-61, 68-99), VO speed from raw position differences
(…:36-43).  Ground truth is an urban stop-and-go profile integrated with the same kinematic
bicycle recurrence as vmvo/bicycle_model.py:66-75.

Host-side input synthesis only (NumPy); nothing here is on the measured path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .constants import STEERING_RATIO, WHEEL_BASE

EPOCH0 = 1658384707.877      # dataset id 1658384707877 is a millisecond epoch stamp


@dataclass
class DriveBatch:
    """n_drives drives of n_frames frames each (row-major, drive-major)."""

    time: np.ndarray    # float64 [D, n]
    vo: np.ndarray      # float32 [D, n, 4]  x, y, theta, v
    gps: np.ndarray     # float32 [D, n, 4]
    imu: np.ndarray     # float32 [D, n]     yaw
    gt: np.ndarray      # float64 [D, n, 4]
    dt: float

    def drive(self, d: int):
        return self.time[d], self.vo[d], self.gps[d], self.imu[d]


def _trailing_mean(a: np.ndarray, w: int) -> np.ndarray:
    """Mean of the last min(i+1, w) samples along axis 1 (the reference's smoothen_traj)."""
    c = np.cumsum(a, axis=1)
    out = c.copy()
    out[:, w:] = c[:, w:] - c[:, :-w]
    cnt = np.minimum(np.arange(1, a.shape[1] + 1), w).reshape((1, -1) + (1,) * (a.ndim - 2))
    return out / cnt


def synthetic_drives(n_drives: int, n_frames: int, seed: int = 1658384707877 % (2 ** 32),
                     dt: float = 0.05) -> DriveBatch:
    rng = np.random.default_rng(seed)
    D, n = n_drives, n_frames

    # piecewise targets, re-drawn every 4-20 s per drive
    v_t = np.zeros(D)
    s_t = np.zeros(D)
    hold_v = np.zeros(D, dtype=np.int64)
    hold_s = np.zeros(D, dtype=np.int64)
    v = rng.uniform(0.0, 12.0, D)
    s = np.zeros(D)
    v_gt = np.empty((D, n))
    s_gt = np.empty((D, n))
    for k in range(n):
        new_v = hold_v <= 0
        if new_v.any():
            m = int(new_v.sum())
            stop = rng.random(m) < 0.15
            v_t[new_v] = np.where(stop, 0.0, rng.uniform(3.0, 15.0, m))
            hold_v[new_v] = rng.integers(int(4 / dt), int(20 / dt), m)
        new_s = hold_s <= 0
        if new_s.any():
            m = int(new_s.sum())
            straight = rng.random(m) < 0.5
            s_t[new_s] = np.where(straight, rng.normal(0.0, 5.0, m), rng.uniform(-200.0, 200.0, m))
            hold_s[new_s] = rng.integers(int(2 / dt), int(10 / dt), m)
        hold_v -= 1
        hold_s -= 1
        v = np.clip(v + np.clip(v_t - v, -3.0 * dt, 3.0 * dt), 0.0, 15.0)
        s = s + np.clip(s_t - s, -50.0 * dt, 50.0 * dt)
        v_gt[:, k] = v
        s_gt[:, k] = s

    delta = np.radians(s_gt) / STEERING_RATIO
    th = np.cumsum(v_gt / WHEEL_BASE * np.tan(delta) * dt, axis=1)
    th += rng.uniform(-np.pi, np.pi, (D, 1))
    x = np.cumsum(v_gt * np.cos(th) * dt, axis=1)
    y = np.cumsum(v_gt * np.sin(th) * dt, axis=1)
    gt = np.stack([x, y, th, v_gt], axis=2)

    time = EPOCH0 + np.arange(n)[None, :] * dt + np.zeros((D, 1))

    # VO: drift + small white noise + sparse jerks, trailing 20-tap average, speed from raw diffs
    drift = np.cumsum(rng.normal(0.0, 0.02, (D, n, 2)), axis=1)
    white = rng.normal(0.0, 0.02, (D, n, 2))
    jerk = (rng.random((D, n, 1)) < 0.002) * rng.normal(0.0, 0.5, (D, n, 2))
    raw = gt[:, :, :2] + drift + white + jerk
    vo_xy = _trailing_mean(raw, 20)
    d = np.diff(raw, axis=1)
    vo_v = np.concatenate([np.zeros((D, 1)), np.hypot(d[..., 0], d[..., 1]) / dt], axis=1)
    vo_th = th + np.cumsum(rng.normal(0.0, 2e-4, (D, n)), axis=1) + rng.normal(0.0, 0.01, (D, n))
    vo = np.concatenate([vo_xy, vo_th[..., None], vo_v[..., None]], axis=2)

    # GPS: 10 Hz fix held to 20 Hz, sigma 1.5 m, same smoothing, tangent heading
    fix = gt[:, ::2, :2] + rng.normal(0.0, 1.5, (D, (n + 1) // 2, 2))
    held = np.repeat(fix, 2, axis=1)[:, :n]
    gps_xy = _trailing_mean(held, 20)
    g = np.diff(gps_xy, axis=1)
    gps_th = np.arctan2(g[..., 1], g[..., 0])
    gps_th = np.concatenate([gps_th, gps_th[:, -1:]], axis=1)
    gps_v = np.abs(v_gt + rng.normal(0.0, 0.3, (D, n)))
    gps = np.concatenate([gps_xy, gps_th[..., None], gps_v[..., None]], axis=2)

    imu = th + np.cumsum(rng.normal(0.0, 1e-4, (D, n)), axis=1) + rng.normal(0.0, 0.002, (D, n))

    return DriveBatch(time=time, vo=vo.astype(np.float32), gps=gps.astype(np.float32),
                      imu=imu.astype(np.float32), gt=gt, dt=dt)


def off_float32_grid(a: np.ndarray) -> np.ndarray:
    """float64 copy of ``a`` nudged off the float32 grid by a deterministic pattern of pure IEEE
    operations (no libm), relative size ~1e-9: inputs for the float64 stream entry points."""
    a = np.asarray(a, dtype=np.float64)
    k = (np.arange(a.size, dtype=np.float64).reshape(a.shape) % 7.0) - 3.0
    return a + (np.abs(a) + 1.0) * (k / 3.0) * 1.0e-9
