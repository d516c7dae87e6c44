"""Times the search kernel alone on config 2 (32x32, W=30, 9940 windows) and config 3 (256x256,
W=60, 4096 windows): CUDA events around the C-ABI call, L2 flushed between launches.

    python tools/time_search.py [cfg2] [cfg3] [cfg5]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vehiclemodelvisualodometry_b200 import DriveSet, SearchConfig, grid_search, plan_windows  # noqa: E402
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives  # noqa: E402

# key=value arguments go to the library's tuning hook (e.g. defer_warps=8); "eager" times eager
# launches through ctypes instead of a replayed CUDA graph of the search's launches
tune = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
eager = "eager" in sys.argv[1:]
which = [a for a in sys.argv[1:] if "=" not in a and a != "eager"] or ["cfg2", "cfg3"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
from vehiclemodelvisualodometry_b200 import _lib  # noqa: E402
for k_, v_ in tune.items():
    _lib.context(0).set_tuning(k_, int(v_))


def timeit(name, cfg, drives, reps):
    plan = plan_windows(cfg, drives, extents=False)
    out = torch.empty((plan.n_windows, 64), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        so = grid_search(cfg, drives, plan, out=out)
    torch.cuda.synchronize()
    g = None
    if not eager:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            grid_search(cfg, drives, plan, out=out)
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if g is None:
            grid_search(cfg, drives, plan, out=out)
        else:
            g.replay()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    rec = so.records()
    hs = cfg.grid_v * cfg.grid_s * int(rec["n_steps"].astype(np.int64).sum())
    med = float(np.median(ms))
    print(f"{name} {tune or ''}: {med:.4f} ms (min {min(ms):.4f})  {hs / med / 1e9:.3f} T hyp-steps/s  windows {len(rec)}  "
          f"rescored mean {rec['n_rescored'].mean():.2f} p99 {np.percentile(rec['n_rescored'], 99):.0f} "
          f"max {rec['n_rescored'].max()}", flush=True)


if "cfg2" in which:
    n, cfg = bench.make_cfg("config2_single_drive_10k_32x32_w30")
    b = synthetic_drives(1, n, seed=bench.BASE_SEED)
    t, vo, _, _ = b.drive(0)
    timeit("cfg2 32x32 w30", cfg, DriveSet.from_arrays([t], [b.dt], vo=[vo]), 20)
if "cfg3" in which:
    n, cfg = bench.make_cfg("config3_dense_256x256_w60")
    b = synthetic_drives(1, n, seed=bench.BASE_SEED + 3)
    t, vo, _, _ = b.drive(0)
    timeit("cfg3 256x256 w60", cfg, DriveSet.from_arrays([t], [b.dt], vo=[vo]), 5)
if "cfg5" in which:   # VO + GPS + IMU fused cost, 128x128, W = 60 (a slice of BASELINE configs[4])
    cfg = SearchConfig(grid_v=128, grid_s=128, window_frames=60, w_vo=1.0, w_gps=0.5, w_imu=40.0)
    b = synthetic_drives(4, 2000, seed=bench.BASE_SEED + 5)
    dr = DriveSet.from_arrays(list(b.time), [b.dt] * 4, vo=list(b.vo), gps=list(b.gps), imu=list(b.imu))
    timeit("cfg5 128x128 w60 vo+gps+imu", cfg, dr, 5)
if "cfg4" in which:   # batch of drives, 128x128, W = 60, VO only
    cfg = SearchConfig(grid_v=128, grid_s=128, window_frames=60)
    b = synthetic_drives(4, 2000, seed=bench.BASE_SEED + 4)
    dr = DriveSet.from_arrays(list(b.time), [b.dt] * 4, vo=list(b.vo))
    timeit("cfg4 128x128 w60 vo", cfg, dr, 5)
