"""One GPU: how long the search of one rank's SHARE of a pooled window list takes (no peers, no
arrival words), against the same number of windows of one drive -- is the deal itself free?

    python tools/deal_probe.py [world ...]       # default: 2 8
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vehiclemodelvisualodometry_b200 import _lib, grid_search, plan_windows, write_back  # noqa: E402

dev = torch.device("cuda", 0)
timer = bench.Timer(dev, 1)
workload = "config2_single_drive_10k_32x32_w30"


def graph_ms(fn, k=30):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timer(g.replay, k, 3, collective=False) / k


def share(world, rank, block):
    ex = _lib.Exchange()
    ex.world, ex.rank, ex.block, ex.n_peers = world, rank, block, 0
    return ex


for world in [int(a) for a in sys.argv[1:]] or [2, 8]:
    n_frames, cfg, drives, _, _ = bench.pooled_drives(workload, world, dev)
    plan = plan_windows(cfg, drives)
    n = plan.n_windows
    out = torch.zeros((n, 64), dtype=torch.uint8, device=dev)
    per = n // world
    print(f"-- pool of {world} drives, {n} windows", flush=True)
    print(f"   all windows on one GPU: {graph_ms(lambda: grid_search(cfg, drives, plan, out=out)) * 1e3:8.1f} us")
    for d in range(min(world, 4)):
        ms = graph_ms(lambda: grid_search(cfg, drives, plan, window_range=(d * per, (d + 1) * per),
                                          out=out[d * per:(d + 1) * per]))
        print(f"   drive {d} alone (contiguous share {d}): {ms * 1e3:8.1f} us")
    for block in (1, 8, 32, 128, 512):
        t = []
        for r in range(min(world, 4)):
            ex = share(world, r, block)
            t.append(graph_ms(lambda: grid_search(cfg, drives, plan, out=out, exchange=ex)) * 1e3)
        print(f"   dealt share, block {block:4d}: ranks 0..{len(t) - 1}: " + " ".join(f"{x:7.1f}" for x in t) + " us")
    traj = torch.empty((4, drives.n_frames), dtype=torch.float64, device=dev)
    grid_search(cfg, drives, plan, out=out)
    print(f"   plan of all windows: {graph_ms(lambda: plan_windows(cfg, drives, into=plan)) * 1e3:6.1f} us;  "
          f"write-back of one rank's frames: "
          f"{graph_ms(lambda: write_back(cfg, drives, plan, out, blend_gps=False, out=traj, frame_range=(0, drives.n_frames // world))) * 1e3:6.1f} us;  "
          f"of all frames: {graph_ms(lambda: write_back(cfg, drives, plan, out, blend_gps=False, out=traj)) * 1e3:6.1f} us", flush=True)
