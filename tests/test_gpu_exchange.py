"""The multi-GPU path on ONE device: two processes share cuda:0, deal the global window list
block-cyclically, store their records into each other's gather buffer over CUDA IPC from inside the
search kernel, and signal arrival with the flag words the write-back kernel publishes and waits on
(scheduler.PeerGather, include/vmvo_b200.h vmvo_exchange).  Every record and every written-back
frame must equal what one process computes alone.

(The loop being sharded is vmvo/scripts/optimize_trajectory_v2.py:48-146 of the reference: windows
are independent, so who searches which one must not matter.)
"""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
rank, world = int(sys.argv[1]), 2
torch.cuda.set_device(0)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
from vehiclemodelvisualodometry_b200 import (DrivePipeline, DriveSet, SearchConfig, _lib, grid_search,
                                             plan_windows, write_back)
from vehiclemodelvisualodometry_b200.scheduler import PeerGather, deal_indices, shard_range
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

dev = torch.device("cuda", 0)
mode = {mode!r}
cfg = (SearchConfig(grid_v=16, grid_s=16, window_frames=20) if mode != "chained" else
       SearchConfig(grid_v=8, grid_s=16, window_frames=20, seed_mode="chained"))
batch = synthetic_drives(3, 300, seed=21)
drives = DriveSet.from_arrays(list(batch.time), [batch.dt] * 3, vo=list(batch.vo), gps=list(batch.gps), device=dev)
plan = plan_windows(cfg, drives)
n = plan.n_windows
assert n == 3 * 260

def same_records(got, want, what):
    a = got.cpu().numpy().reshape(-1, 64).copy()
    b = want.cpu().numpy().reshape(-1, 64).copy()
    a[:, 12:16] = 0            # n_rescored: a diagnostic that may vary from run to run
    b[:, 12:16] = 0
    bad = np.nonzero((a != b).any(axis=1))[0]
    assert len(bad) == 0, (what, rank, bad[:8])

# what one process computes alone
alone = grid_search(cfg, drives, plan)
traj_alone = write_back(cfg, drives, plan, alone.results)
torch.cuda.synchronize()

# 1. stand-alone arrival: search -> publish -> wait; every record of both ranks is here afterwards
g = PeerGather(n, dev, block=4)
mine = deal_indices(n, 4, world, rank)
assert len(mine) == g.my_count() and 0 < len(mine) < n
if mode != "chained":
    grid_search(cfg, drives, plan, out=g.buffer, exchange=g.exchange)
    g.publish()
    g.wait()
    torch.cuda.synchronize()
    same_records(g.buffer, alone.results, "publish/wait")
    assert g.step() == 1 and g.timed_out() == 0
g.close()

# 1b. windows that are not searched (more poses than max_window_poses: status TOO_LONG) and windows with
#     a non-finite pose leave their records on both ranks as well
if mode != "chained":
    import dataclasses
    vo_nan = [v.copy() for v in batch.vo]
    vo_nan[1][40:43, 0] = np.nan
    drives_nan = DriveSet.from_arrays(list(batch.time), [batch.dt] * 3, vo=vo_nan, gps=list(batch.gps), device=dev)
    cfg_long = SearchConfig(grid_v=8, grid_s=8, window_mode="time", horizon_time=1.0, horizon_frames=20,
                            max_window_poses=16)
    for cfg_u, dr_u in ((cfg_long, drives), (cfg, drives_nan)):
        plan_u = plan_windows(cfg_u, dr_u)
        alone_u = grid_search(cfg_u, dr_u, plan_u)
        st = alone_u.records()["status"]
        assert (st != 0).any()
        g = PeerGather(plan_u.n_windows, dev, block=4)
        grid_search(cfg_u, dr_u, plan_u, out=g.buffer, exchange=g.exchange)
        g.publish()
        g.wait()
        torch.cuda.synchronize()
        same_records(g.buffer, alone_u.results, "unsearched / non-finite windows")
        g.close()

# 2. the pipeline: two buffer sets used alternately, CUDA graphs, no host synchronisation between
#    steps; the write-back of this rank's frames consumes both ranks' records
sets = [PeerGather(n, dev, block={block}) for _ in range(2)]
lo, hi = shard_range(drives.n_frames, world, rank)
pipes = [DrivePipeline(cfg, drives, gather=s, frame_range=(lo, hi)) for s in sets]
for p in pipes:
    p.trajectory.fill_(float("nan"))
steps = 7
for s in range(steps):
    pipes[s & 1].run()
torch.cuda.synchronize()
for b, (s, p) in enumerate(zip(sets, pipes)):
    same_records(s.buffer, alone.results, "pipeline set %d" % b)
    assert s.timed_out() == 0
    assert s.step() == 1 + (steps + 1 - b) // 2, (s.step(), b)     # the eager warm-up pass counts
    t = p.trajectory.cpu().numpy()
    want = traj_alone.cpu().numpy()
    assert np.array_equal(t[:, lo:hi], want[:, lo:hi], equal_nan=True), "write-back share differs"
    outside = np.ones(drives.n_frames, bool)
    outside[lo:hi] = False
    assert np.all(np.isnan(t[:, outside])), "write-back touched frames outside its share"
del pipes
for s in sets:
    s.close()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def _run_pair(tmp_path, mode, block):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port, mode=mode, block=block))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=420)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,block", [("data", 8), ("data", 64), ("chained", 1)])
def test_two_ranks_on_one_device_exchange_records(cuda_device, tmp_path, mode, block):
    """data seeds: windows dealt in blocks of 8 / 64; chained seeds: whole drives dealt."""
    _run_pair(tmp_path, mode, block)
