"""Trajectory pre-processing.  CPU: the oracle against vectors frozen from the
unmodified reference (tests/golden/prep_kats.json).  GPU: the kernels against the oracle and the
golden vectors, through the reference-shaped DataFrame facade and the batched API."""
import json
import os

import numpy as np
import pytest

from oracle import prep_oracle as P
from tests.helpers import unhex

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prep_kats.json")))
COLS = ("x", "y", "theta", "velocity", "time")


def _vo_frame(n, seed):
    from oracle.make_golden_prep import vo_frame
    return vo_frame(n, seed)


def _gps_frame(n, seed, repeat_last=True):
    from oracle.make_golden_prep import gps_frame
    return gps_frame(n, seed, repeat_last)


# ---- CPU: oracle vs golden ----------------------------------------------------------------------
def test_oracle_smoothing_golden():
    for c in G["smooth"]:
        xy = unhex(c["xy"], (c["n"], 2))
        np.testing.assert_array_equal(P.smoothen(xy, c["window"]), unhex(c["out"], (c["n"], 2)))


def test_oracle_vo_golden():
    for c in G["vo"]:
        x, y, rot, stamp = _vo_frame(c["n"], c["seed"])
        got = P.process_vo(x, y, rot, stamp)
        for k in COLS:
            np.testing.assert_allclose(got[k], unhex(c[k]), rtol=4e-16, atol=1e-300)
        assert got["velocity"][0] == 0.0 and len(got["x"]) == c["n"]


def test_oracle_gps_golden():
    for c in G["gps"]:
        lat, lon, heading, speed, stamp = _gps_frame(c["n"], c["seed"])
        got = P.process_gps(lat, lon, speed, stamp)
        assert len(got["x"]) == c["n"] + 1 and len(got["theta"]) == c["n"]     # quirk D8
        for k in COLS:
            np.testing.assert_allclose(got[k], unhex(c[k]), rtol=1e-12, atol=1e-9)
    lat, lon, heading, speed, stamp = _gps_frame(30, 5, repeat_last=False)
    with pytest.raises(IndexError):
        P.process_gps(lat, lon, speed, stamp)


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_smoothen_traj(cuda_device):
    from vehiclemodelvisualodometry_b200.trajectory import smooth_batch, smoothen_traj
    for c in G["smooth"]:
        xy = unhex(c["xy"], (c["n"], 2))
        np.testing.assert_array_equal(np.asarray(smoothen_traj(xy, window_size=c["window"])),
                                      unhex(c["out"], (c["n"], 2)))          # IEEE adds: bit-exact
    rng = np.random.default_rng(3)
    xs = [np.cumsum(rng.normal(0, 1, n)) for n in (7, 100, 21, 1)]
    ys = [np.cumsum(rng.normal(0, 1, n)) for n in (7, 100, 21, 1)]
    for (sx, sy), x, y in zip(smooth_batch(xs, ys, 20), xs, ys):
        ref = P.smoothen(np.stack([x, y], axis=1), 20)
        np.testing.assert_array_equal(sx, ref[:, 0])
        np.testing.assert_array_equal(sy, ref[:, 1])


@pytest.mark.gpu
def test_gpu_process_vo_trajectory(cuda_device):
    import pandas as pd
    from vehiclemodelvisualodometry_b200.schema import Trajectory
    from vehiclemodelvisualodometry_b200.trajectory import process_vo_trajectory, vo_prepare_batch
    for c in G["vo"]:
        x, y, rot, stamp = _vo_frame(c["n"], c["seed"])
        df = pd.DataFrame({"x": x, "y": y, "rot": list(rot), "Timestamp": stamp})
        tr = process_vo_trajectory(df)
        assert isinstance(tr, Trajectory) and len(tr) == c["n"]
        np.testing.assert_array_equal(tr.x, unhex(c["x"]))                    # IEEE only: bit-exact
        np.testing.assert_array_equal(tr.y, unhex(c["y"]))
        np.testing.assert_array_equal(tr.velocity, unhex(c["velocity"]))
        np.testing.assert_array_equal(tr.time, unhex(c["time"]))
        np.testing.assert_allclose(tr.theta, unhex(c["theta"]), rtol=0, atol=1e-15)   # atan2
    frames = [_vo_frame(n, s) for n, s in ((33, 9), (5, 10), (400, 11))]
    outs = vo_prepare_batch(*[[f[k] for f in frames] for k in range(4)])
    for f, o in zip(frames, outs):
        ref = P.process_vo(*f)
        for k in COLS:
            np.testing.assert_allclose(o[k], ref[k], rtol=0, atol=1e-15)


@pytest.mark.gpu
def test_gpu_process_gps_trajectory(cuda_device):
    import pandas as pd
    from vehiclemodelvisualodometry_b200.trajectory import gps_prepare_batch, process_gps_trajectory
    for c in G["gps"]:
        lat, lon, heading, speed, stamp = _gps_frame(c["n"], c["seed"])
        df = pd.DataFrame({"heading": heading, "Latitude": lat, "Longitude": lon, "speed": speed,
                           "Timestamp": stamp})
        tr = process_gps_trajectory(df)
        assert len(tr.x) == c["n"] + 1 and len(tr.theta) == c["n"] and len(tr.velocity) == c["n"] + 1
        # ECEF goes through sin/cos of CUDA's libm (<= 2 ulp from glibc's; one ulp of an ECEF
        # coordinate is 9.3e-10 m).  Measured on the fixtures (tools/gps_deviation.py): positions
        # within 2.0e-10 m, speed within 1.8e-9 relative, heading within 2.1e-10 rad; asserted with a
        # 50x margin.  The path is a running sum, so the position bound grows like sqrt(n) ulps.
        np.testing.assert_allclose(tr.x, unhex(c["x"]), rtol=0, atol=1e-8)
        np.testing.assert_allclose(tr.y, unhex(c["y"]), rtol=0, atol=1e-8)
        np.testing.assert_array_equal(tr.time, unhex(c["time"]))
        want_v = unhex(c["velocity"])
        np.testing.assert_allclose(tr.velocity, want_v, rtol=1e-7, atol=1e-12)   # product of tiny deltas
        dth = np.abs(np.asarray(tr.theta) - unhex(c["theta"]))
        moving = np.hypot(np.diff(unhex(c["x"])), np.diff(unhex(c["y"]))) > 1e-4
        assert np.all(np.minimum(dth, 2 * np.pi - dth)[moving] < 1e-8)
    # the reference dies with IndexError when the log ends on a fresh fix
    lat, lon, heading, speed, stamp = _gps_frame(30, 5, repeat_last=False)
    df = pd.DataFrame({"heading": heading, "Latitude": lat, "Longitude": lon, "speed": speed, "Timestamp": stamp})
    with pytest.raises(IndexError):
        process_gps_trajectory(df)
    # batched, ragged drives
    frames = [_gps_frame(n, s) for n, s in ((41, 20), (200, 21), (8, 22))]
    outs = gps_prepare_batch([f[0] for f in frames], [f[1] for f in frames], [f[3] for f in frames],
                             [f[4] for f in frames])
    for f, o in zip(frames, outs):
        ref = P.process_gps(f[0], f[1], f[3], f[4])
        np.testing.assert_allclose(o["x"], ref["x"], rtol=0, atol=1e-8)
        np.testing.assert_allclose(o["y"], ref["y"], rtol=0, atol=1e-8)
        np.testing.assert_array_equal(o["time"], ref["time"])


@pytest.mark.gpu
def test_gpu_full_chain_like_reference_main(cuda_device):
    """The call sequence of the reference's main() (optimize_trajectory_v2.py:168-183):
    process_vo_trajectory + process_gps_trajectory + optimize_trajectory, DataFrames in,
    Trajectory out -- against the same chain through the oracles."""
    import pandas as pd
    from oracle import vmvo_oracle as O
    from tests.helpers import spec_of
    from vehiclemodelvisualodometry_b200 import BicycleModel, SearchConfig, optimize_trajectory
    from vehiclemodelvisualodometry_b200.optimize import DEFAULT_CFG
    from vehiclemodelvisualodometry_b200.trajectory import process_gps_trajectory, process_vo_trajectory

    n = 260
    x, y, rot, _ = _vo_frame(n, 31)
    lat, lon, heading, speed, stamp = _gps_frame(n, 32)
    vo_df = pd.DataFrame({"x": x, "y": y, "rot": list(rot), "Timestamp": stamp})
    gps_df = pd.DataFrame({"heading": heading, "Latitude": lat, "Longitude": lon, "speed": speed,
                           "Timestamp": stamp})
    vo_t, gps_t = process_vo_trajectory(vo_df), process_gps_trajectory(gps_df)
    assert len(vo_t) == n and len(gps_t) == n + 1
    cfg = SearchConfig(**{**DEFAULT_CFG.__dict__, "grid_v": 8, "grid_s": 8})
    out = optimize_trajectory(vo_t, gps_t, BicycleModel(), config=cfg)
    assert len(out) == n

    # the same chain on the CPU: oracle prep (float64), streams rounded to float32 like the
    # search ABI does, oracle driver
    pv, pg = P.process_vo(x, y, rot, stamp), P.process_gps(lat, lon, speed, stamp)
    N = min(n, n + 1)
    def stream(p):
        th = np.concatenate([p["theta"], p["theta"][-1:]])[:len(p["x"])] if len(p["theta"]) < len(p["x"]) else p["theta"]
        return np.stack([p["x"], p["y"], th, p["velocity"]], axis=1)[:N].astype(np.float32)
    # the facade derives dt and the horizon from the GPS stamps (optimize_trajectory_v2.py:35-42)
    dt, horizon, _ = O.reference_dt(np.asarray(gps_t.time))
    spec = O.replace(spec_of(cfg), horizon_frames=horizon, horizon_time=3.0)
    ref = O.optimize_drive(spec, np.asarray(gps_t.time)[:N], dt, stream(pv), stream(pg))
    np.testing.assert_allclose(out.x[:N], ref.x, rtol=0, atol=2e-6)
    np.testing.assert_allclose(out.y[:N], ref.y, rtol=0, atol=2e-6)
    np.testing.assert_allclose(out.velocity[:N], ref.velocity, rtol=1e-3, atol=1e-9)
