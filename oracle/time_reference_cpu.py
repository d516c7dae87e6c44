"""Times the UNMODIFIED Python reference on the path -- build container only.

TEST / MEASUREMENT INFRASTRUCTURE (needs /root/reference; never runs on the GPU box).

    python oracle/time_reference_cpu.py            # writes profiles/reference_cpu_r02.json

Two legs, both on a synthetic drive shaped like dataset 1658384707877 (BASELINE configs[0]:
6 000 frames at 20 Hz, seed 1658384707877 mod 2^32; SURVEY.md 8d "Config 1"), one worker
process per host core (``multiprocessing.Pool``):

1. **reference as-is**: ``vmvo.scripts.optimize_trajectory_v2.optimize_trajectory`` -- SciPy
   SLSQP per window (vmvo/utils/mpc.py:112-119), default 3.0 s horizon -- on short segments of
   the drive (each segment of 2*horizon + k frames has k windows; the reference has no way to
   run a sub-range of windows).  ``BicycleModel.run`` calls are counted by a wrapper that only
   increments an integer.  Reported: windows/s and model steps/s over all cores.
2. **the hypothesis grid through the reference's own model**: every hypothesis of a 32x32 grid
   over 30-step windows (BASELINE configs[1]'s shape) rolled with ``BicycleModel.run``
   (vmvo/bicycle_model.py:40-78) and scored with the literal cost closure of
   vmvo/utils/mpc.py:68-80 -- the same hypotheses the GPU searches -- on a sample of windows.
   Reported: hypothesis-steps/s per core and over all cores.

bench.py carries the numbers of the committed JSON as ``cpu_baseline.reference_python``.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

SEED = 1658384707877 % (2 ** 32)
N_FRAMES = 6000


def _drive():
    from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

    b = synthetic_drives(1, N_FRAMES, seed=SEED)
    t, vo, gps, _ = b.drive(0)
    return b.dt, t, vo.astype(np.float64), gps.astype(np.float64)


def _slsqp_segment(args):
    """The reference's optimize_trajectory on frames [start, start + 2*horizon + k)."""
    start, k = args
    from oracle import ref_bridge

    ref = ref_bridge.load()
    v2 = ref_bridge.load_v2()
    dt, t, vo, gps = _drive()
    horizon = int(3.0 * int(1.0 / dt))                    # optimize_trajectory_v2.py:35-42
    sl = slice(start, start + 2 * horizon + k)
    T = ref.schema.Trajectory
    vo_t = T(x=vo[sl, 0], y=vo[sl, 1], theta=vo[sl, 2], velocity=vo[sl, 3], time=t[sl])
    gps_t = T(x=gps[sl, 0], y=gps[sl, 1], theta=gps[sl, 2], velocity=gps[sl, 3], time=t[sl])
    calls = [0]
    run0 = ref.bicycle_model.BicycleModel.run

    def counted(self, *a, **kw):
        calls[0] += 1
        return run0(self, *a, **kw)

    ref.bicycle_model.BicycleModel.run = counted
    t0 = time.perf_counter()
    try:
        with ref_bridge.quiet():
            v2.optimize_trajectory(vo_t, gps_t, ref.bicycle_model.BicycleModel())
        done = k
    except (IndexError, AssertionError):                  # quirk D7: a stationary window
        done = 0
    sec = time.perf_counter() - t0
    ref.bicycle_model.BicycleModel.run = run0
    return done, calls[0], sec


def _grid_windows(args):
    """32x32 hypotheses x 30 steps through BicycleModel.run for the given window starts."""
    starts, gv, gs, W = args
    from oracle import ref_bridge
    from oracle import vmvo_oracle as O
    from oracle.make_golden import ref_grid_window

    ref = ref_bridge.load()
    dt, t, vo, gps = _drive()
    spec = O.SearchSpec(grid_v=gv, grid_s=gs, window_frames=W)
    steps = 0
    t0 = time.perf_counter()
    for s in starts:
        wt = O.build_window(spec, int(s), W + 1, dt, vo.astype(np.float32), gps.astype(np.float32), None)
        if wt.n_steps == 0:
            continue
        cost = ref_grid_window(ref, spec, wt, dt)
        assert int(np.argmin(cost.reshape(-1))) >= 0
        steps += gv * gs * wt.n_steps
    return steps, time.perf_counter() - t0


def main():
    cores = max(1, len(os.sched_getaffinity(0)))
    out = {"generator": "oracle/time_reference_cpu.py (reference imported unmodified from /root/reference)",
           "where": "build container (the reference does not exist on the GPU box)",
           "host_cores": cores,
           "drive": f"synthetic, {N_FRAMES} frames @ 20 Hz, seed {SEED} (BASELINE configs[0] shape)"}
    try:
        import platform

        out["cpu"] = platform.processor() or open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t")
    except Exception:
        pass

    # leg 1: SLSQP as-is; one segment per core, k windows each
    k = int(os.environ.get("VMVO_REF_WINDOWS_PER_CORE", "8"))
    seg_starts = np.linspace(200, N_FRAMES - 200 - 2 * 60 - k, cores).astype(int)
    t0 = time.perf_counter()
    with mp.Pool(cores) as pool:
        res = pool.map(_slsqp_segment, [(int(s), k) for s in seg_starts])
    wall = time.perf_counter() - t0
    wins = sum(r[0] for r in res)
    calls = sum(r[1] for r in res)
    out["slsqp_as_is"] = {
        "what": "vmvo.scripts.optimize_trajectory_v2.optimize_trajectory (SciPy SLSQP, 3.0 s horizon)",
        "windows": wins, "segments": len(res), "model_steps": calls, "wall_s": wall,
        "windows_per_s": wins / wall, "model_steps_per_s": calls / wall,
        "model_steps_per_s_per_core": calls / sum(r[2] for r in res),
        "cores": cores}
    print(json.dumps(out["slsqp_as_is"]), flush=True)

    # leg 2: the 32x32 x 30 grid through BicycleModel.run; 4 windows per core
    gv, gs, W = 32, 32, 30
    per = int(os.environ.get("VMVO_REF_GRID_WINDOWS_PER_CORE", "4"))
    starts = np.linspace(0, N_FRAMES - 2 * W - 1, cores * per).astype(int)
    t0 = time.perf_counter()
    with mp.Pool(cores) as pool:
        res = pool.map(_grid_windows, [(starts[c::cores].tolist(), gv, gs, W) for c in range(cores)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    out["grid_through_bicycle_model_run"] = {
        "what": "every hypothesis of the 32x32 grid, 30-step windows, rolled with BicycleModel.run and "
                "scored with the cost closure of vmvo/utils/mpc.py:68-80",
        "windows": int(len(starts)), "hypothesis_steps": steps, "wall_s": wall,
        "hypothesis_steps_per_s": steps / wall,
        "hypothesis_steps_per_s_per_core": steps / sum(r[1] for r in res), "cores": cores}
    print(json.dumps(out["grid_through_bicycle_model_run"]), flush=True)

    path = os.path.join(ROOT, "profiles", "reference_cpu_r02.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
