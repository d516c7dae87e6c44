"""Times the search kernel on BASELINE configs[2] (256x256, W = 60, 4096 windows)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vehiclemodelvisualodometry_b200 import DriveSet, grid_search, plan_windows  # noqa: E402
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives  # noqa: E402

n, cfg = bench.make_cfg("config3_dense_256x256_w60")
b = synthetic_drives(1, n, seed=bench.BASE_SEED + 3)
t, vo, _, _ = b.drive(0)
dr = DriveSet.from_arrays([t], [b.dt], vo=[vo])
pl = plan_windows(cfg, dr)
for _ in range(2):
    so = grid_search(cfg, dr, pl)
torch.cuda.synchronize()
ms = []
for _ in range(3):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    so = grid_search(cfg, dr, pl)
    e.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(e))
rec = so.records()
hs = cfg.grid_v * cfg.grid_s * int(rec["n_steps"].astype(np.int64).sum())
print(f"dense: {min(ms):.3f} ms  {hs / min(ms) / 1e9:.1f} T hyp-steps/s  rescored mean {rec['n_rescored'].mean():.1f}")
