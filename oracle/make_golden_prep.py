"""Pins oracle/prep_oracle.py to the UNMODIFIED reference pre-processing functions and freezes
golden vectors (tests/golden/prep_kats.json).  Build container only (needs /root/reference and
pandas).  TEST INFRASTRUCTURE.

    python oracle/make_golden_prep.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import prep_oracle as P  # noqa: E402
from oracle import ref_bridge  # noqa: E402
from oracle.make_golden import hexf  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "prep_kats.json")


def vo_frame(n, seed):
    """A VO cache shaped like <id>_traj.csv after loading (bdd_raw.py:157-168)."""
    rng = np.random.default_rng(seed)
    yaw = np.cumsum(rng.normal(0, 0.01, n))
    rot = np.zeros((n, 3, 3))
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1], rot[:, 2, 2] = np.cos(yaw), -np.sin(yaw), np.sin(yaw), np.cos(yaw), 1
    x = np.cumsum(rng.normal(0.5, 0.2, n))
    y = np.cumsum(rng.normal(0.1, 0.2, n))
    stamp = 1658384707877 + np.cumsum(rng.integers(45, 56, n))
    return x, y, rot, stamp.astype(np.int64)


def gps_frame(n, seed, repeat_last=True):
    """A BDD CSV: 20 Hz log, fix refreshed at 10 Hz (vmvo/utils/trajectory.py:220-223)."""
    rng = np.random.default_rng(seed)
    k = (n + 1) // 2
    lat = 12.97 + np.cumsum(rng.normal(2e-6, 1e-6, k))
    lon = 77.59 + np.cumsum(rng.normal(3e-6, 1e-6, k))
    lat, lon = np.repeat(lat, 2)[:n], np.repeat(lon, 2)[:n]
    if repeat_last and n % 2 == 1:        # an odd count would end on a fresh fix
        lat[-1], lon[-1] = lat[-2], lon[-2]
    if not repeat_last:
        lat[-1] += 1e-6
    heading = rng.uniform(0, 360, n)
    speed = np.abs(rng.normal(8, 2, n))
    stamp = 1658384707877 + 50 * np.arange(n)
    return lat, lon, heading, speed, stamp.astype(np.int64)


def main():
    assert ref_bridge.available()
    ref_bridge.load()
    import vmvo.utils.trajectory as T   # needs the matplotlib shim installed by ref_bridge.load()

    out = {"generator": "oracle/make_golden_prep.py", "vo": [], "gps": [], "smooth": []}

    rng = np.random.default_rng(0)
    for n, w in ((5, 20), (20, 20), (21, 20), (64, 3), (300, 20)):
        xy = np.cumsum(rng.normal(0, 1, (n, 2)), axis=0)
        ref = np.asarray(T.smoothen_traj(xy, window_size=w), dtype=np.float64)
        mine = P.smoothen(xy, w)
        assert np.array_equal(ref, mine), ("smoothen", n, w)
        out["smooth"].append({"n": n, "window": w, "xy": hexf(xy), "out": hexf(mine)})

    for n, seed in ((40, 1), (257, 2)):
        x, y, rot, stamp = vo_frame(n, seed)
        df = pd.DataFrame({"x": x, "y": y, "rot": list(rot), "Timestamp": stamp})
        ref = T.process_vo_trajectory(df)
        mine = P.process_vo(x, y, rot, stamp)
        for k in ("x", "y", "theta", "velocity", "time"):
            assert np.array_equal(np.asarray(getattr(ref, k)), mine[k]), ("vo", k)
        out["vo"].append({"n": n, "seed": seed, **{k: hexf(v) for k, v in mine.items()}})

    for n, seed in ((60, 3), (301, 4)):
        lat, lon, heading, speed, stamp = gps_frame(n, seed)
        df = pd.DataFrame({"heading": heading, "Latitude": lat, "Longitude": lon, "speed": speed,
                           "Timestamp": stamp})
        ref = T.process_gps_trajectory(df)
        mine = P.process_gps(lat, lon, speed, stamp)
        for k in ("x", "y", "theta", "velocity", "time"):
            assert np.array_equal(np.asarray(getattr(ref, k)), mine[k], equal_nan=True), ("gps", k)
        assert len(ref.x) == n + 1 and len(ref.theta) == n
        out["gps"].append({"n": n, "seed": seed, **{k: hexf(v) for k, v in mine.items()}})

    # a log that ends on a fresh fix: the reference indexes past the end of `velocity`
    lat, lon, heading, speed, stamp = gps_frame(30, 5, repeat_last=False)
    df = pd.DataFrame({"heading": heading, "Latitude": lat, "Longitude": lon, "speed": speed, "Timestamp": stamp})
    for fn in (lambda: T.process_gps_trajectory(df), lambda: P.process_gps(lat, lon, speed, stamp)):
        try:
            fn()
            raise SystemExit("expected IndexError")
        except IndexError:
            pass
    out["gps_fresh_last_fix"] = "IndexError"

    with open(GOLDEN, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
