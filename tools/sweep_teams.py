"""Times the search kernel for each team size on one workload (tuning hook team_warps)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vehiclemodelvisualodometry_b200 import DriveSet, _lib, grid_search, plan_windows  # noqa: E402
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "config2_single_drive_10k_32x32_w30"
n_frames, cfg = bench.make_cfg(workload)
if len(sys.argv) > 2:
    n_frames = int(sys.argv[2])
batch = synthetic_drives(1, n_frames, seed=bench.BASE_SEED)
t, vo, _, _ = batch.drive(0)
drives = DriveSet.from_arrays([t], [batch.dt], vo=[vo])
plan = plan_windows(cfg, drives)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tw in (1, 2, 4, 8):
    _lib.context(0).set_tuning("team_warps", tw)
    for _ in range(3):
        so = grid_search(cfg, drives, plan)
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        so = grid_search(cfg, drives, plan)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    rec = so.records()
    hs = cfg.grid_v * cfg.grid_s * int(rec["n_steps"].astype(np.int64).sum())
    print(f"{workload} team_warps={tw}: {np.median(ms):.4f} ms  {hs / np.median(ms) / 1e9:.1f} G hyp-steps/s  "
          f"rescored mean {rec['n_rescored'].mean():.2f} max {rec['n_rescored'].max()}")
