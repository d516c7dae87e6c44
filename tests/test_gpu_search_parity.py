"""GPU parity of the fused window search against the oracle (through the C ABI).

Bar (BASELINE.json north_star): selected hypothesis indices bit-exact; rollout poses within
1e-4 m / 1e-5 rad -- the float64 re-score puts the measured deviation near 1e-13, asserted
here at 1e-9.
"""
import dataclasses
import zlib

import numpy as np
import pytest
import torch

from oracle import vmvo_oracle as O
from tests.helpers import assert_records_match, oracle_windows, spec_of
from vehiclemodelvisualodometry_b200 import (DriveSet, SearchConfig, grid_search, optimize_drives,
                                             plan_windows)
from vehiclemodelvisualodometry_b200.synthetic import off_float32_grid, synthetic_drives

pytestmark = pytest.mark.gpu

CASES = {
    "vo_4x4": SearchConfig(grid_v=4, grid_s=4, window_frames=10),
    "vo_5x7": SearchConfig(grid_v=5, grid_s=7, window_frames=12),
    "vo_1x9": SearchConfig(grid_v=1, grid_s=9, window_frames=16),
    "vo_9x1": SearchConfig(grid_v=9, grid_s=1, window_frames=16),
    "vo_16x16": SearchConfig(grid_v=16, grid_s=16, window_frames=20),
    # neither grid_s nor the VD table width (24 columns) is a power of two: the kernel divides
    "vo_24x12": SearchConfig(grid_v=24, grid_s=12, window_frames=18),
    "vo_32x32_w30": SearchConfig(grid_v=32, grid_s=32, window_frames=30),
    "vo_offset0": SearchConfig(grid_v=12, grid_s=12, window_frames=20, target_offset=0),
    "gps_traverse": SearchConfig(grid_v=8, grid_s=24, window_frames=40, target_mode="traverse",
                                 primary="gps", w_vo=0.0, w_gps=1.0),
    "vo_gps": SearchConfig(grid_v=16, grid_s=12, window_frames=24, w_vo=1.0, w_gps=0.25),
    "vo_gps_imu": SearchConfig(grid_v=16, grid_s=16, window_frames=24, w_vo=1.0, w_gps=0.5, w_imu=40.0),
    "vo_imu": SearchConfig(grid_v=8, grid_s=32, window_frames=30, w_imu=10.0),
    "vo_ksteer": SearchConfig(grid_v=8, grid_s=16, window_frames=20, k_steer=5e-6),
    "time_windows": SearchConfig(grid_v=8, grid_s=8, window_mode="time", horizon_time=1.5,
                                 horizon_frames=29),
    "w70_three_rounds": SearchConfig(grid_v=8, grid_s=8, window_frames=70),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_search_matches_oracle(cuda_device, name):
    cfg = CASES[name]
    n = 2 * cfg.horizon() + 25
    batch = synthetic_drives(2, n, seed=zlib.crc32(name.encode()) % 1000)
    drives = DriveSet.from_arrays(list(batch.time), [batch.dt] * 2, vo=list(batch.vo), gps=list(batch.gps),
                                  imu=list(batch.imu))
    plan = plan_windows(cfg, drives)
    assert plan.n_windows == 2 * 25
    out = grid_search(cfg, drives, plan, want_rollouts=True)
    rec = out.records()
    ref = []
    for d in range(2):
        ref += oracle_windows(cfg, batch.time[d], batch.dt, batch.vo[d], batch.gps[d], batch.imu[d])
    assert_records_match(rec, ref)
    poses, steer, vel = out.poses.cpu().numpy(), out.steer.cpu().numpy(), out.vel.cpu().numpy()
    for w, r in enumerate(ref):
        N = r.n_steps
        np.testing.assert_allclose(poses[w, :N], r.poses, rtol=0, atol=1e-9)
        # IEEE ops on the seeds; the seed itself carries atan's last-bit difference
        np.testing.assert_allclose(steer[w, :N], r.steer, rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(vel[w, :N], r.vel, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("name", ["vo_16x16", "gps_traverse", "vo_gps_imu", "time_windows"])
def test_float64_streams_match_oracle(cuda_device, name):
    """vmvo_grid_search_f64: double4 pose streams, inputs float32 cannot represent."""
    cfg = CASES[name]
    n = 2 * cfg.horizon() + 25
    batch = synthetic_drives(2, n, seed=zlib.crc32(name.encode()) % 1000 + 1)
    vo, gps, imu = off_float32_grid(batch.vo), off_float32_grid(batch.gps), off_float32_grid(batch.imu)
    drives = DriveSet.from_arrays(list(batch.time), [batch.dt] * 2, vo=list(vo), gps=list(gps),
                                  imu=list(imu), stream_dtype=np.float64)
    assert drives.f64 and drives.vo.dtype == torch.float64
    plan = plan_windows(cfg, drives)
    out = grid_search(cfg, drives, plan, want_rollouts=True)
    rec = out.records()
    ref = []
    for d in range(2):
        ref += oracle_windows(cfg, batch.time[d], batch.dt, vo[d], gps[d], imu[d])
    assert_records_match(rec, ref)
    poses = out.poses.cpu().numpy()
    for w, r in enumerate(ref):
        np.testing.assert_allclose(poses[w, :r.n_steps], r.poses, rtol=0, atol=1e-9)
    # the float32 entry point on the same data sees rounded inputs: its costs differ
    d32 = DriveSet.from_arrays(list(batch.time), [batch.dt] * 2, vo=list(vo), gps=list(gps), imu=list(imu))
    rec32 = grid_search(cfg, d32, plan).records()
    assert not np.array_equal(rec32["best_cost"], rec["best_cost"])


def test_mixed_stream_dtypes_are_rejected(cuda_device):
    cfg = CASES["vo_gps"]
    batch = synthetic_drives(1, 80, seed=3)
    drives = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]], gps=[batch.gps[0]])
    drives.gps = drives.gps.double()
    with pytest.raises(ValueError, match="one dtype"):
        grid_search(cfg, drives, plan_windows(cfg, drives))


@pytest.mark.parametrize("fast", ["0", "1"])
def test_both_scan_paths_match_oracle(cuda_device, tuning, fast):
    """Both scans (the packed / rotation one is the default on large grids only) on a 32x32 VO
    search."""
    tuning("fast_scan", int(fast))
    cfg = SearchConfig(grid_v=32, grid_s=32, window_frames=30)
    batch = synthetic_drives(1, 110, seed=71)
    drives = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]])
    rec = grid_search(cfg, drives, plan_windows(cfg, drives)).records()
    assert_records_match(rec, oracle_windows(cfg, batch.time[0], batch.dt, batch.vo[0]))


def test_negative_speed_seed_uses_generic_scan(cuda_device, tuning):
    tuning("fast_scan", 1)
    # V_w < 0 (a caller-supplied seed): hypotheses start clamped and move later, so the
    # affine-heading scan does not apply; results must still match
    cfg = SearchConfig(grid_v=16, grid_s=16, window_frames=20, seed_mode="given")
    batch = synthetic_drives(1, 70, seed=72)
    seeds = np.tile([[-2.0, 15.0]], (30, 1))
    seeds[::2, 0] = 3.0
    drives = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]])
    rec = grid_search(cfg, drives, plan_windows(cfg, drives), seeds=torch.as_tensor(seeds)).records()
    assert_records_match(rec, oracle_windows(cfg, batch.time[0], batch.dt, batch.vo[0], seeds=seeds))


def test_given_seeds(cuda_device):
    cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=15, seed_mode="given")
    batch = synthetic_drives(1, 60, seed=4)
    rng = np.random.default_rng(0)
    seeds = np.stack([rng.uniform(0, 15, 30), rng.uniform(-460, 460, 30)], axis=1)
    drives = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]])
    plan = plan_windows(cfg, drives)
    rec = grid_search(cfg, drives, plan, seeds=torch.as_tensor(seeds)).records()
    ref = oracle_windows(cfg, batch.time[0], batch.dt, batch.vo[0], seeds=seeds)
    assert_records_match(rec, ref)


def test_window_range_and_out_buffer(cuda_device):
    cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=10)
    batch = synthetic_drives(1, 80, seed=2)
    drives = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]])
    plan = plan_windows(cfg, drives)
    full = grid_search(cfg, drives, plan).records()
    buf = torch.zeros((plan.n_windows, 64), dtype=torch.uint8, device=cuda_device)
    grid_search(cfg, drives, plan, window_range=(0, 23), out=buf[:23])
    grid_search(cfg, drives, plan, window_range=(23, plan.n_windows), out=buf[23:])
    from vehiclemodelvisualodometry_b200 import _lib
    got = buf.cpu().numpy().view(_lib.RESULT_DTYPE).reshape(-1)
    for f in ("best_idx", "n_steps", "status", "best_cost", "x1", "y1", "theta1"):
        np.testing.assert_array_equal(got[f], full[f])


@pytest.mark.parametrize("stream", ["f32", "f64"])
@pytest.mark.parametrize("target_mode", ["time", "traverse"])
def test_chained_seed_mode(cuda_device, target_mode, stream):
    """seed_mode chained (optimize_trajectory_v2.py:46,72,146): S_w of window w+1 is the last
    steering angle of window w's optimum; drives are independent runs."""
    cfg = SearchConfig(grid_v=8, grid_s=16, window_frames=20, seed_mode="chained", target_mode=target_mode)
    lengths = [75, 40, 120]           # the middle drive has no windows at all
    batch = synthetic_drives(3, 120, seed=17)
    time = [batch.time[d][:n] for d, n in enumerate(lengths)]
    vo = [batch.vo[d][:n] for d, n in enumerate(lengths)]
    if stream == "f64":
        vo = [off_float32_grid(v) for v in vo]
    drives = DriveSet.from_arrays(time, [batch.dt] * 3, vo=vo,
                                  stream_dtype=np.float64 if stream == "f64" else np.float32)
    so, traj, plan = optimize_drives(cfg, drives)
    rec = so.records()
    assert plan.window_offsets == [0, 35, 35, 115]
    spec = spec_of(cfg)
    ref_windows = []
    for d in range(3):
        ref = O.optimize_drive(spec, time[d], batch.dt, vo[d])
        ref_windows += ref.windows
        lo, hi = drives.drive_offsets[d], drives.drive_offsets[d + 1]
        np.testing.assert_allclose(traj[0, lo:hi].cpu().numpy(), ref.x, rtol=0, atol=1e-9)
        np.testing.assert_allclose(traj[1, lo:hi].cpu().numpy(), ref.y, rtol=0, atol=1e-9)
    assert_records_match(rec, ref_windows)
    # the chain itself is pure IEEE arithmetic on the selected index: bit-exact
    np.testing.assert_array_equal(rec["s_seed"], [r.s_seed for r in ref_windows])
    assert rec["s_seed"][0] == 0.0 and rec["s_seed"][35] == 0.0 and np.any(rec["s_seed"] != 0.0)
    with pytest.raises(ValueError, match="whole drives"):
        grid_search(cfg, drives, plan, window_range=(0, 20))


def test_search_without_planned_extents_matches_planned(cuda_device):
    """Frames mode: the search kernel derives each window's extent itself (no planning launch);
    same records as with the plan of vmvo_plan_windows, ragged drives and a drive without windows
    included; data and chained seeds."""
    batch = synthetic_drives(4, 100, seed=31)
    lens = [80, 30, 24, 100]
    t = [batch.time[d][:n] for d, n in enumerate(lens)]
    vo = [batch.vo[d][:n] for d, n in enumerate(lens)]
    for seed_mode in ("data", "chained"):
        cfg = SearchConfig(grid_v=8, grid_s=16, window_frames=12, seed_mode=seed_mode)
        drives = DriveSet.from_arrays(t, [batch.dt] * 4, vo=vo)
        plan = plan_windows(cfg, drives)
        bare = plan_windows(cfg, drives, extents=False)
        assert bare.win_start is None and bare.window_offsets == plan.window_offsets == [0, 56, 62, 62, 138]
        a = grid_search(cfg, drives, plan).records()
        b = grid_search(cfg, drives, bare).records()
        for f in ("best_idx", "n_steps", "status", "best_cost", "v_seed", "s_seed", "x1", "y1", "theta1"):
            np.testing.assert_array_equal(a[f], b[f])
        if seed_mode == "data":
            for d in (0, 1, 3):
                lo, hi = plan.window_offsets[d], plan.window_offsets[d + 1]
                assert_records_match(b[lo:hi], oracle_windows(cfg, t[d], batch.dt, vo[d]))
    with pytest.raises(ValueError, match="frames"):
        plan_windows(SearchConfig(window_mode="time"), drives, extents=False)


PRUNE_CASES = {
    # two-pass kernel, packed scan, chunks dealt middle-out (config 2's shape)
    "vo_32x32_w30": (SearchConfig(grid_v=32, grid_s=32, window_frames=30), 400, {}),
    # the same through the generic scan
    "vo_32x32_generic": (SearchConfig(grid_v=32, grid_s=32, window_frames=30), 300, {"fast_scan": 0}),
    # many-pass kernel (four-warp teams, middle-out passes), packed scan
    "vo_128x128_w60": (SearchConfig(grid_v=128, grid_s=128, window_frames=60), 130, {}),
    # the dense grid of BASELINE configs[2]: eight-warp teams, one acceleration chunk per pass
    "vo_256x256_w60": (SearchConfig(grid_v=256, grid_s=256, window_frames=60), 128, {}),
    # the same grid as the first many-pass case with eight-warp teams
    "vo_128x128_w60_tw8": (SearchConfig(grid_v=128, grid_s=128, window_frames=60), 126, {"team_warps": 8}),
    # two position terms + yaw term: generic scan, many passes
    "vo_gps_imu_64x64": (SearchConfig(grid_v=64, grid_s=64, window_frames=40, w_vo=1.0, w_gps=0.5,
                                      w_imu=40.0), 120, {}),
    # a grid whose item count is not a multiple of the warp (partial vote masks)
    "vo_40x20": (SearchConfig(grid_v=40, grid_s=20, window_frames=25), 150, {}),
    # votes after every step
    "vo_32x32_every1": (SearchConfig(grid_v=32, grid_s=32, window_frames=30), 200, {"prune_every": 1}),
}


@pytest.mark.parametrize("name", sorted(PRUNE_CASES))
def test_pruned_scan_gives_the_exhaustive_records(cuda_device, tuning, name):
    """A warp stops scanning once every hypothesis it holds has passed the candidate threshold of the
    bound its team held at the start of the pass: index, float64 cost and first pose of every window
    are those of the exhaustive scan BIT FOR BIT (and those of the oracle)."""
    cfg, frames, tune = PRUNE_CASES[name]
    for k, v in tune.items():
        tuning(k, v)
    batch = synthetic_drives(1, frames, seed=zlib.crc32(name.encode()) % 1000)
    t, vo, gps, imu = batch.drive(0)
    kw = {"vo": [vo]}
    if cfg.w_gps:
        kw["gps"] = [gps]
    if cfg.w_imu:
        kw["imu"] = [imu]
    drives = DriveSet.from_arrays([t], [batch.dt], **kw)
    plan = plan_windows(cfg, drives)
    pruned = grid_search(cfg, drives, plan).records()
    tuning("prune", 0)
    full = grid_search(cfg, drives, plan).records()
    for f in ("best_idx", "n_steps", "status"):
        np.testing.assert_array_equal(pruned[f], full[f])
    for f in ("best_cost", "x1", "y1", "theta1", "v_seed", "s_seed"):
        np.testing.assert_array_equal(pruned[f].view(np.uint64), full[f].view(np.uint64))
    ref = oracle_windows(cfg, t, batch.dt, vo, gps if cfg.w_gps else None, imu if cfg.w_imu else None)
    assert_records_match(pruned, ref)


@pytest.mark.parametrize("name", ["vo_32x32_w30", "vo_128x128_w60", "vo_gps_imu_64x64"])
def test_lean_kernels_give_the_generic_kernels_records(cuda_device, tuning, name):
    """The specialised kernels (MODE 1: switches of the non-default features compiled out; MODE 2: also
    always fed preparation records) against the generic one (MODE 0), and MODE 1 against MODE 2 on
    the small grid: the same records bit for bit."""
    cfg, frames, _ = PRUNE_CASES[name]
    batch = synthetic_drives(1, frames, seed=zlib.crc32(name.encode()) % 1000 + 1)
    t, vo, gps, imu = batch.drive(0)
    kw = {"vo": [vo]}
    if cfg.w_gps:
        kw["gps"] = [gps]
    if cfg.w_imu:
        kw["imu"] = [imu]
    drives = DriveSet.from_arrays([t], [batch.dt], **kw)
    plan = plan_windows(cfg, drives)
    got = {}
    for label, tune in (("default", {}), ("generic", {"lean": 0}), ("lean_no_prep", {"prep": 0})):
        for k, v in tune.items():
            tuning(k, v)
        got[label] = grid_search(cfg, drives, plan).records()
        for k in tune:
            tuning(k, -1)
    for label in ("generic", "lean_no_prep"):
        for f in ("best_idx", "n_steps", "status"):
            np.testing.assert_array_equal(got[label][f], got["default"][f])
        for f in ("best_cost", "x1", "y1", "theta1", "v_seed", "s_seed"):
            np.testing.assert_array_equal(got[label][f].view(np.uint64), got["default"][f].view(np.uint64))


def test_random_configurations_default_equals_generic_exhaustive(cuda_device):
    """tools/random_equivalence.py: 40 random grids / windows / cost terms / target modes / stream widths,
    the default search (pruning votes, lean and preparation kernels) against the generic exhaustive one,
    and against its own repeat: bit-identical records."""
    import subprocess
    import sys as _sys
    import os as _os

    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, _os.path.join(root, "tools", "random_equivalence.py"), "40", "11"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
