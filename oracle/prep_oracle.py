"""CPU oracle for the trajectory pre-processing step before the window search
 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Float64 restatement, array in / array out, of (paths relative to /root/reference):

  smoothen_traj             vmvo/utils/trajectory.py:68-99    trailing moving average
  process_vo_trajectory     vmvo/utils/trajectory.py:13-65    yaw from R, speed, smoothing, scale
  geodetic_to_euclidean     vmvo/utils/trajectory.py:120-174  WGS84 -> ECEF delta of two fixes
  process_gps_trajectory    vmvo/utils/trajectory.py:177-335  ECEF path, speed, 10->20 Hz de-dup
                                                               interpolation, smoothing, tangent yaw

Pinned to the unmodified reference by oracle/make_golden_prep.py (pandas DataFrames shaped like
the dataset's CSV / VO cache) and frozen in tests/golden/prep_kats.json.  Quirks reproduced
(SURVEY Appendix D8): VO speed divides by a millisecond difference; GPS speed uses the PRODUCT
(dx^2 * dy^2)^0.5; the GPS path has n+1 points (a leading duplicate of the origin) and theta has
one element fewer than x; `heading` only feeds a variable that is never returned.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

WGS84_A = 6378137
WGS84_E = 8.1819190842622e-2
TWO_PI = 2 * np.pi


def smoothen(xy: np.ndarray, window: int) -> np.ndarray:
    """vmvo/utils/trajectory.py:68-99: mean of the last min(i+1, window) points, summed left to
    right from 0 like Python's ``sum``; returned unchanged when n <= window."""
    xy = np.asarray(xy, dtype=np.float64)
    n = len(xy)
    if n <= window:
        return xy.copy()
    out = np.empty((n, 2), dtype=np.float64)
    for i in range(n):
        lo = max(0, i - window + 1)
        sx = 0.0
        sy = 0.0
        for p in xy[lo:i + 1]:
            sx = sx + p[0]
            sy = sy + p[1]
        cnt = i + 1 - lo
        out[i, 0] = sx / cnt
        out[i, 1] = sy / cnt
    return out


def process_vo(x, y, rot, stamp_ms, scale=0.25, window=20) -> Dict[str, np.ndarray]:
    """vmvo/utils/trajectory.py:13-65.  ``rot`` [n, 3, 3]; ``stamp_ms`` the Timestamp column."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    rot = np.asarray(rot)
    if rot.dtype != np.float32:     # float32 matrices (the cached trajectory, bdd_raw.py:163-164)
        rot = rot.astype(np.float64)  # give a float32 arctan2, like the reference's scalar calls
    t = np.asarray(stamp_ms, dtype=np.float64)
    n = len(x)
    theta = np.arctan2(rot[:, 1, 0], rot[:, 0, 0]).astype(np.float64)
    vel = np.zeros(n, dtype=np.float64)
    for i in range(n - 1):
        dist = np.sqrt((x[i] - x[i + 1]) ** 2 + (y[i] - y[i + 1]) ** 2)
        vel[i + 1] = dist / (t[i + 1] - t[i])          # quirk: a millisecond difference
    sm = smoothen(np.stack([x, y], axis=1), window)
    return {"x": sm[:, 0] * scale, "y": sm[:, 1] * scale, "theta": theta, "velocity": vel,
            "time": t / 1000.0}


def ecef(lat_deg: float, lon_deg: float) -> Tuple[float, float, float]:
    """One point of vmvo/utils/trajectory.py:127-165, same operation order."""
    a, e = WGS84_A, WGS84_E
    lat, lon = math.radians(lat_deg), math.radians(lon_deg)
    x = a / math.sqrt(1 - e ** 2 * math.sin(lat) ** 2) * math.cos(lat) * math.cos(lon)
    y = a / math.sqrt(1 - e ** 2 * math.sin(lat) ** 2) * math.cos(lat) * math.sin(lon)
    z = a * (1 - e ** 2) / math.sqrt(1 - e ** 2 * math.sin(lat) ** 2) * math.sin(lat)
    return x, y, z


def geodetic_to_euclidean(p1, p2):
    x1, y1, z1 = ecef(p1[0], p1[1])
    x2, y2, z2 = ecef(p2[0], p2[1])
    return x2 - x1, y2 - y1, z2 - z1


def process_gps(lat, lon, speed, stamp_ms, window=20) -> Dict[str, np.ndarray]:
    """vmvo/utils/trajectory.py:177-335 (the ``heading`` column only feeds dead code there)."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    n = len(lat)
    x = np.zeros(n + 1)
    y = np.zeros(n + 1)
    la1, lo1 = lat[0], lon[0]
    for i in range(n):
        dx, dy, _ = geodetic_to_euclidean((la1, lo1), (lat[i], lon[i]))
        x[i + 1] = dx + x[i]
        y[i + 1] = dy + y[i]
        la1, lo1 = lat[i], lon[i]
    velocity = np.asarray(speed, dtype=np.float64)
    time = np.asarray(stamp_ms, dtype=np.float64) / 1000.0
    est = np.zeros_like(velocity)
    est[0] = velocity[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(1, n):
            est[i] = ((x[i] - x[i - 1]) ** 2 * (y[i] - y[i - 1]) ** 2) ** 0.5 / (time[i] - time[i - 1])
    velocity = est

    m = n + 1
    x_new, y_new, v_new, t_new = [x[0]], [y[0]], [velocity[0]], [time[0]]
    last = 0
    for i in range(1, m):
        if x[last] != x[i] or y[last] != y[i]:
            for j in range(last + 1, i + 1):
                alpha = (j - last) / (i - last)
                x_new.append(x[last] * (1 - alpha) + x[i] * alpha)
                y_new.append(y[last] * (1 - alpha) + y[i] * alpha)
                v_new.append(velocity[last] * (1 - alpha) + velocity[i] * alpha)   # IndexError at i == n
                t_new.append(time[last] * (1 - alpha) + time[i] * alpha)
            last = i
    for j in range(last + 1, m):
        alpha = (j - last) / (m - last)
        x_new.append(x[last] * (1 - alpha) + x[-1] * alpha)
        y_new.append(y[last] * (1 - alpha) + y[-1] * alpha)
        v_new.append(velocity[last] * (1 - alpha) + velocity[-1] * alpha)
        t_new.append(time[last] * (1 - alpha) + time[-1] * alpha)
    assert len(x_new) == m, f"Length mismatch {len(x_new)} != {m}"

    sm = smoothen(np.stack([x_new, y_new], axis=1), window)
    xs, ys = sm[:, 0], sm[:, 1]
    theta = np.empty(m - 1)
    for i in range(m - 1):
        angle = np.arctan2(xs[i + 1] - xs[i], ys[i + 1] - ys[i])
        theta[i] = (angle + np.pi) % TWO_PI
    return {"x": -xs, "y": ys.copy(), "theta": theta, "velocity": np.asarray(v_new), "time": np.asarray(t_new)}
