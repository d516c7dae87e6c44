"""Shared helpers of the parity tests: config mapping and oracle runs."""
from __future__ import annotations

import dataclasses
import json
import os

import numpy as np

from oracle import vmvo_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def unhex(lst, shape=None):
    a = np.array([float.fromhex(v) for v in lst], dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def spec_of(cfg) -> O.SearchSpec:
    """SearchConfig (product) -> SearchSpec (oracle): same field names by construction."""
    names = {f.name for f in dataclasses.fields(O.SearchSpec)}
    return O.SearchSpec(**{k: v for k, v in dataclasses.asdict(cfg).items() if k in names})


def oracle_windows(cfg, time, dt, vo, gps=None, imu=None, seeds=None):
    """Per-window oracle results for one drive."""
    spec = spec_of(cfg)
    starts, lens = O.window_extents(spec, np.asarray(time))
    out = []
    for w, (s, l) in enumerate(zip(starts, lens)):
        given = None if seeds is None else (seeds[w, 0], seeds[w, 1])
        wt = O.build_window(spec, int(s), int(l), dt, vo, gps, imu, given)
        out.append(O.solve_window(spec, wt, dt))
    return out


def assert_records_match(rec, ref_windows, cost_rtol=1e-9, pose_atol=1e-9):
    """Bit-exact indices / step counts / status; float64 fields to a few ulp-scale units."""
    assert len(rec) == len(ref_windows)
    want_idx = np.array([r.best_idx for r in ref_windows])
    want_n = np.array([r.n_steps for r in ref_windows])
    want_status = np.array([r.status for r in ref_windows])
    np.testing.assert_array_equal(rec["n_steps"], want_n)
    np.testing.assert_array_equal(rec["status"], want_status)
    bad = np.nonzero(rec["best_idx"] != want_idx)[0]
    assert len(bad) == 0, (f"{len(bad)} argmin mismatches, first at window {bad[:5]}: "
                           f"got {rec['best_idx'][bad[:5]]} want {want_idx[bad[:5]]}")
    for w, r in enumerate(ref_windows):
        if r.n_steps == 0:
            continue
        np.testing.assert_allclose(rec["v_seed"][w], r.v_seed, rtol=1e-14, atol=0, equal_nan=True)
        np.testing.assert_allclose(rec["s_seed"][w], r.s_seed, rtol=1e-12, atol=1e-12, equal_nan=True)
        if r.status & O.WIN_NONFINITE:
            assert np.isnan(rec["best_cost"][w])
            continue
        np.testing.assert_allclose(rec["best_cost"][w], r.best_cost, rtol=cost_rtol, atol=1e-18)
        np.testing.assert_allclose([rec["x1"][w], rec["y1"][w], rec["theta1"][w]], r.poses[0],
                                   rtol=0, atol=pose_atol)
