"""Randomised equivalence of the kernel variants (run on a B200): random grids (powers of two and not),
windows, cost terms, target modes and stream widths; the default search (pruning votes, lean / preparation
kernels) against the generic exhaustive one (tuning prune = 0, lean = 0, prep = 0) -- every record field
but n_rescored must be bit-identical.  Also repeats the default search to catch run-to-run differences.

    python tools/random_equivalence.py [cases] [seed]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vehiclemodelvisualodometry_b200 import DriveSet, SearchConfig, _lib, grid_search, plan_windows  # noqa: E402
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = _lib.context(0)
FIELDS = ("best_idx", "n_steps", "status", "best_cost", "x1", "y1", "theta1", "v_seed", "s_seed")
bad = 0
for c in range(cases):
    gv = int(rng.choice([1, 3, 8, 16, 24, 32, 40, 64, 128]))
    gs = int(rng.choice([1, 5, 8, 16, 20, 32, 64, 128]))
    if gv * gs > 128 * 64:
        gs = 32
    W = int(rng.integers(4, 66))
    terms = rng.choice(["vo", "vo_gps", "vo_imu", "vo_gps_imu", "gps"])
    kw = dict(grid_v=gv, grid_s=gs, window_frames=W)
    if terms == "vo_gps":
        kw.update(w_vo=1.0, w_gps=float(rng.uniform(0.05, 2.0)))
    elif terms == "vo_imu":
        kw.update(w_imu=float(rng.uniform(1.0, 50.0)))
    elif terms == "vo_gps_imu":
        kw.update(w_vo=1.0, w_gps=float(rng.uniform(0.05, 2.0)), w_imu=float(rng.uniform(1.0, 50.0)))
    elif terms == "gps":
        kw.update(w_vo=0.0, w_gps=1.0, primary="gps")
    if rng.random() < 0.25:
        kw.update(target_mode="traverse")
    if rng.random() < 0.2:
        kw.update(target_offset=0)
    cfg = SearchConfig(**kw)
    n_drives = int(rng.integers(1, 4))
    frames = 2 * W + int(rng.integers(20, 260))
    b = synthetic_drives(n_drives, frames, seed=int(rng.integers(0, 2 ** 31)))
    sd = np.float64 if rng.random() < 0.3 else np.float32
    dr = DriveSet.from_arrays(list(b.time), [b.dt] * n_drives, vo=list(b.vo), gps=list(b.gps), imu=list(b.imu),
                              stream_dtype=sd)
    plan = plan_windows(cfg, dr)
    fast = grid_search(cfg, dr, plan).records()
    again = grid_search(cfg, dr, plan).records()
    with ctx.tuning(prune=0, lean=0, prep=0):
        full = grid_search(cfg, dr, plan).records()
    for name, other in (("repeat", again), ("generic exhaustive", full)):
        diff = [f for f in FIELDS if np.ascontiguousarray(fast[f]).tobytes() != np.ascontiguousarray(other[f]).tobytes()]
        if diff:
            bad += 1
            w = int(np.nonzero(fast["best_idx"] != other["best_idx"])[0][:1].sum())
            print(f"case {c}: {kw} drives {n_drives} frames {frames} {sd.__name__}: {name} differs in {diff} "
                  f"(first idx mismatch at window {w})", flush=True)
print(f"{cases} random cases, {bad} with differences")
sys.exit(1 if bad else 0)
