"""Runs the fused search a few times on one bench workload (target for ncu / sanitizer runs).

    python tools/profile_search.py [workload] [launches] [frames] [seed offset]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vehiclemodelvisualodometry_b200 import DriveSet, grid_search, plan_windows  # noqa: E402
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "config2_single_drive_10k_32x32_w30"
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_frames, cfg = bench.make_cfg(workload)
if len(sys.argv) > 3:
    n_frames = int(sys.argv[3])
seed_off = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # drive base + seed_off (bench.py: rank)
batch = synthetic_drives(1, n_frames, seed=bench.BASE_SEED + seed_off)
t, vo, _, _ = batch.drive(0)
drives = DriveSet.from_arrays([t], [batch.dt], vo=[vo])
plan = plan_windows(cfg, drives)
for _ in range(launches):
    so = grid_search(cfg, drives, plan)
torch.cuda.synchronize()
rec = so.records()
print(workload, "windows", plan.n_windows, "rescored/window", rec["n_rescored"].mean(),
      "max", rec["n_rescored"].max(), "status!=0:", int((rec["status"] != 0).sum()))
print("rescored percentiles", np.percentile(rec["n_rescored"], [50, 90, 99, 100]))
