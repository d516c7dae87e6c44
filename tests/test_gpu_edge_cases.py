"""Edge cases of the window search: structural ties, saturation, mirror symmetry, non-finite
and empty windows, capacity overflow of the candidate list."""
import numpy as np
import pytest
import torch

from oracle import vmvo_oracle as O
from tests.helpers import assert_records_match, oracle_windows
from vehiclemodelvisualodometry_b200 import DriveSet, SearchConfig, grid_search, plan_windows
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

pytestmark = pytest.mark.gpu


def _run(cfg, time, dt, vo, gps=None, imu=None, seeds=None):
    drives = DriveSet.from_arrays([time], [dt], vo=[vo], gps=None if gps is None else [gps],
                                  imu=None if imu is None else [imu])
    plan = plan_windows(cfg, drives)
    out = grid_search(cfg, drives, plan, seeds=None if seeds is None else torch.as_tensor(seeds))
    return out.records()


def _straight(n, v, dt=0.05, heading=0.0):
    t = 100.0 + np.arange(n) * dt
    s = v * dt * np.arange(n)
    vo = np.stack([s * np.cos(heading), s * np.sin(heading), np.full(n, heading), np.full(n, v)], axis=1)
    return t, vo.astype(np.float32)


def test_stationary_vehicle_all_tied_lowest_index_wins(cuda_device):
    """V_w = 0: every a_i <= 0 row never moves, all its costs tie exactly -> index 0."""
    cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=10)
    n = 30
    t = 5.0 + np.arange(n) * 0.05
    vo = np.zeros((n, 4), dtype=np.float32)
    rec = _run(cfg, t, 0.05, vo)
    ref = oracle_windows(cfg, t, 0.05, vo)
    assert_records_match(rec, ref)
    assert np.all(rec["best_idx"] == 0)
    # motionless rows are structural duplicates: only their lowest index is re-scored
    assert np.all(rec["n_rescored"] >= 1) and np.all(rec["n_rescored"] <= 4)


def test_candidate_list_overflow_is_flushed(cuda_device, tuning):
    """More candidates than list entries: the list is re-scored and refilled until drained."""
    cfg = SearchConfig(grid_v=32, grid_s=32, window_frames=30)
    batch = synthetic_drives(1, 160, seed=5)
    full = _run(cfg, batch.time[0], batch.dt, batch.vo[0])
    assert full["n_rescored"].max() >= 2
    tuning("cand_cap", 1)
    rec = _run(cfg, batch.time[0], batch.dt, batch.vo[0])
    ref = oracle_windows(cfg, batch.time[0], batch.dt, batch.vo[0])
    assert_records_match(rec, ref)
    np.testing.assert_array_equal(rec["best_idx"], full["best_idx"])
    np.testing.assert_array_equal(rec["best_cost"], full["best_cost"])


def test_motionless_rows_with_steering_penalty(cuda_device):
    """With K > 0 a motionless row's cost still depends on j: only i-duplicates are dropped."""
    cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=10, k_steer=1e-5, seed_mode="given")
    n = 30
    t = 5.0 + np.arange(n) * 0.05
    vo = np.zeros((n, 4), dtype=np.float32)
    seeds = np.tile([[0.0, -37.0]], (10, 1))
    rec = _run(cfg, t, 0.05, vo, seeds=seeds)
    ref = oracle_windows(cfg, t, 0.05, vo, seeds=seeds)
    assert_records_match(rec, ref)
    # the steering rate that brings S back towards zero fastest wins, in the first motionless row
    assert np.all(rec["best_idx"] // 8 == 0) and np.all(rec["best_idx"] % 8 == 7)


def test_nearly_stationary(cuda_device):
    """Tiny V_w: decelerating rows die after one step; costs differ far below FP32 resolution."""
    cfg = SearchConfig(grid_v=16, grid_s=16, window_frames=12)
    rng = np.random.default_rng(1)
    n = 40
    t = 5.0 + np.arange(n) * 0.05
    vo = np.zeros((n, 4), dtype=np.float32)
    vo[:, 0] = np.cumsum(rng.normal(0, 1e-3, n))
    vo[:, 1] = np.cumsum(rng.normal(0, 1e-3, n))
    vo[:, 2] = rng.normal(0, 0.01, n)
    vo[:, 3] = np.abs(rng.normal(0, 0.02, n))
    rec = _run(cfg, t, 0.05, vo)
    ref = oracle_windows(cfg, t, 0.05, vo)
    assert_records_match(rec, ref)


def test_mirror_symmetric_hypotheses_tie_exactly(cuda_device):
    """Straight target along x with S_w = 0: rate j and G_s-1-j mirror each other and their
    float64 costs are identical; np.argmin keeps the lower index."""
    cfg = SearchConfig(grid_v=6, grid_s=8, window_frames=20, seed_mode="given")
    t, vo = _straight(60, 8.0)
    # seed speed off the true speed so the best steering rate is not the centre one
    seeds = np.tile([[8.0, 0.0]], (20, 1))
    rec = _run(cfg, t, 0.05, vo, seeds=seeds)
    ref = oracle_windows(cfg, t, 0.05, vo, seeds=seeds)
    assert_records_match(rec, ref)
    j = rec["best_idx"] % 8
    assert np.all(j < 4)                # of each mirrored pair the lower index won


def test_steering_seed_saturated(cuda_device):
    """S_w at +max_steer: every r_j >= 0 clamps to the same sequence -> exact ties."""
    cfg = SearchConfig(grid_v=4, grid_s=9, window_frames=15, seed_mode="given")
    batch = synthetic_drives(1, 50, seed=12)
    seeds = np.tile([[6.0, 460.0]], (20, 1))
    rec = _run(cfg, batch.time[0], batch.dt, batch.vo[0], seeds=seeds)
    ref = oracle_windows(cfg, batch.time[0], batch.dt, batch.vo[0], seeds=seeds)
    assert_records_match(rec, ref)


def test_nonfinite_window_flagged(cuda_device):
    cfg = SearchConfig(grid_v=4, grid_s=4, window_frames=10)
    batch = synthetic_drives(1, 45, seed=6)
    vo = batch.vo[0].copy()
    vo[17, 0] = np.nan
    vo[30, 1] = np.inf
    rec = _run(cfg, batch.time[0], batch.dt, vo)
    ref = oracle_windows(cfg, batch.time[0], batch.dt, vo)
    assert_records_match(rec, ref)
    assert np.any(rec["status"] & 2) and not np.all(rec["status"] & 2)
    assert np.all(rec["best_idx"][(rec["status"] & 2) != 0] == 0)


def test_traverse_empty_window_when_stationary(cuda_device):
    """Quirk D7: a stationary window decimates to one point -> N = 0, no result."""
    cfg = SearchConfig(grid_v=4, grid_s=4, window_frames=10, target_mode="traverse")
    n = 40
    t = 5.0 + np.arange(n) * 0.05
    vo = np.zeros((n, 4), dtype=np.float32)
    vo[:, 3] = 1.0                      # claims 1 m/s but never moves
    rec = _run(cfg, t, 0.05, vo)
    ref = oracle_windows(cfg, t, 0.05, vo)
    assert_records_match(rec, ref)
    assert np.all(rec["status"] & 1) and np.all(rec["n_steps"] == 0) and np.all(rec["best_idx"] == -1)


def test_window_longer_than_capacity_is_flagged(cuda_device):
    cfg = SearchConfig(grid_v=4, grid_s=4, window_mode="time", horizon_time=2.0, horizon_frames=40,
                       max_window_poses=16)
    batch = synthetic_drives(1, 100, seed=6)
    rec = _run(cfg, batch.time[0], batch.dt, batch.vo[0])
    assert np.all(rec["status"] == 4) and np.all(rec["best_idx"] == -1)


def test_window_without_frames_is_flagged_and_raises_like_reference(cuda_device):
    """A window whose extent holds no pose (unsorted stamps can produce one): status EMPTY | NO_FRAMES,
    not TOO_LONG, and the facade raises the reference's assert (vmvo/schema.py:122)."""
    from vehiclemodelvisualodometry_b200 import WindowPlan, _lib

    cfg = SearchConfig(grid_v=4, grid_s=4, window_frames=10)
    batch = synthetic_drives(1, 40, seed=6)
    drives = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]])
    plan = plan_windows(cfg, drives)
    assert plan.n_windows == 20
    lens = plan.win_len.clone()
    lens[[3, 11]] = 0
    lens[7] = -5
    hand = WindowPlan(window_offsets=plan.window_offsets, d_window_offsets=plan.d_window_offsets,
                      win_start=plan.win_start, win_len=lens, win_drive=plan.win_drive)
    rec = grid_search(cfg, drives, hand).records()
    ok = grid_search(cfg, drives, plan).records()
    none = np.array([3, 7, 11])
    assert np.all(rec["status"][none] == (_lib.WIN_EMPTY | _lib.WIN_NO_FRAMES))
    assert np.all(rec["best_idx"][none] == -1) and np.all(rec["n_steps"][none] == 0)
    rest = np.setdiff1d(np.arange(20), none)
    for f in ("best_idx", "n_steps", "status", "best_cost"):
        np.testing.assert_array_equal(rec[f][rest], ok[f][rest])


def test_searches_of_one_ctx_overlap_on_two_streams(cuda_device):
    """Every search launch owns its scratch (queue head, deferred-window slots) until it has
    completed: two searches of ONE ctx issued on two streams, both with many parked windows, must
    give the records each gives alone -- launch after launch."""
    cfg = SearchConfig(grid_v=16, grid_s=32, window_frames=30)
    rng = np.random.default_rng(17)
    sets = []
    for speed in (0.9, 1.2):            # crawling vehicles: near-ties, long candidate lists
        t, vo = _straight(600, speed)
        vo[:, :2] += rng.normal(0, 0.05, (600, 2)).astype(np.float32)
        vo[:, 3] += rng.normal(0, 0.05, 600).astype(np.float32)
        d = DriveSet.from_arrays([t], [0.05], vo=[vo])
        sets.append((d, plan_windows(cfg, d)))
    alone = [grid_search(cfg, d, p).records() for d, p in sets]
    assert all(a["n_rescored"].max() >= 16 for a in alone)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    for rep in range(12):
        for q, (d, p) in enumerate(sets):
            with torch.cuda.stream(streams[q]):
                outs[q].append(grid_search(cfg, d, p))
    torch.cuda.synchronize()
    for q in range(2):
        for o in outs[q]:
            r = o.records()
            for f in ("best_idx", "n_steps", "status", "best_cost", "x1", "y1", "theta1"):
                np.testing.assert_array_equal(r[f], alone[q][f])


def test_entry_points_leave_the_current_device_alone(cuda_device):
    """An operator called for another device runs there and restores the caller's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=10)
    batch = synthetic_drives(1, 60, seed=2)
    torch.cuda.set_device(0)
    d1 = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]], device=torch.device("cuda", 1))
    with torch.cuda.device(1):
        want = grid_search(cfg, d1, plan_windows(cfg, d1)).records()
    assert torch.cuda.current_device() == 0
    ref = oracle_windows(cfg, batch.time[0], batch.dt, batch.vo[0])
    assert_records_match(want, ref)


def test_high_speed_large_heading(cuda_device):
    """30 m/s seed, full lock available: headings of several radians, largest FP32 error band."""
    cfg = SearchConfig(grid_v=16, grid_s=32, window_frames=60)
    rng = np.random.default_rng(3)
    n = 130
    dt = 0.05
    t = 1.0 + np.arange(n) * dt
    th = np.cumsum(np.full(n, 0.02))
    x = np.cumsum(30 * np.cos(th) * dt) + rng.normal(0, 0.05, n)
    y = np.cumsum(30 * np.sin(th) * dt) + rng.normal(0, 0.05, n)
    vo = np.stack([x, y, th, np.full(n, 30.0)], axis=1).astype(np.float32)
    rec = _run(cfg, t, dt, vo)
    ref = oracle_windows(cfg, t, dt, vo)
    assert_records_match(rec, ref)


@pytest.mark.parametrize("defer_min", ["0", "1", "3"])
def test_deferred_windows_give_the_same_records(cuda_device, tuning, defer_min):
    """Windows with long candidate lists park them for the second kernel (tuning hook defer_min entries or
    more; 0 = never, 1 = every window): index, cost and first pose must not depend on which kernel
    did the float64 re-scores.  Slow stretches (near-ties over the steering rates) included."""
    cfg = SearchConfig(grid_v=16, grid_s=32, window_frames=30)
    # a vehicle creeping at 0.9 m/s with VO noise: the hardest-braking rows stop after one or two
    # steps and all 32 steering rates of such a row nearly tie; then a normal stretch
    rng = np.random.default_rng(7)
    t, slow = _straight(200, 0.9)
    slow[:, :2] += rng.normal(0, 0.05, (200, 2)).astype(np.float32)
    slow[:, 3] += rng.normal(0, 0.05, 200).astype(np.float32)
    batch = synthetic_drives(1, 200, seed=7)
    fast = batch.vo[0].copy()
    fast[:, :2] += slow[-1, :2] - fast[0, :2]
    vo = np.concatenate([slow, fast]).astype(np.float32)
    time = np.concatenate([t, t[-1] + 0.05 + np.arange(200) * 0.05])
    tuning("defer_min", int(defer_min))
    rec = _run(cfg, time, 0.05, vo)
    tuning("defer_min", 0)
    base = _run(cfg, time, 0.05, vo)
    assert base["n_rescored"].max() >= 16
    for f in ("best_idx", "n_steps", "status", "best_cost", "x1", "y1", "theta1", "v_seed", "s_seed"):
        np.testing.assert_array_equal(rec[f], base[f])
    assert_records_match(rec, oracle_windows(cfg, time, 0.05, vo))


@pytest.mark.parametrize("case", ["w30", "w60_two_rounds", "vo_gps_imu", "ksteer", "many_pass_kernel"])
def test_every_window_deferred_matches_in_kernel_rescore(cuda_device, tuning, case):
    """defer_min = 1 sends EVERY window's list through vmvo_deferred_rescore_kernel; index, cost and
    first pose must equal the in-kernel re-score bit for bit, whichever of its three paths a list
    entry takes (warp_cost64, the 16-lane and the 8-lane packed form), with one or two rounds of 32
    steps and with every cost term.  A stop-and-go drive: its optima include rows that stop within
    8 steps, within 16, and rows that never stop."""
    from oracle import c_oracle

    cfg = {
        "w30": SearchConfig(grid_v=32, grid_s=16, window_frames=30),
        "w60_two_rounds": SearchConfig(grid_v=32, grid_s=8, window_frames=60),
        "vo_gps_imu": SearchConfig(grid_v=32, grid_s=8, window_frames=30, w_vo=1.0, w_gps=0.3, w_imu=20.0),
        "ksteer": SearchConfig(grid_v=32, grid_s=8, window_frames=40, k_steer=2e-6),
        # four passes per window: the kernel variant whose float64 sums are Hillis-Steele; a forced
        # deferral must re-score with the same form
        "many_pass_kernel": SearchConfig(grid_v=128, grid_s=64, window_frames=20),
    }[case]
    batch = synthetic_drives(1, 1500 if case != "many_pass_kernel" else 700, seed=11)
    t, vo, gps, imu = batch.drive(0)
    tuning("defer_min", 1)
    rec = _run(cfg, t, batch.dt, vo, gps, imu)
    tuning("defer_min", 0)
    base = _run(cfg, t, batch.dt, vo, gps, imu)
    a = cfg.max_accel * (2 * (base["best_idx"] // cfg.grid_s) - (cfg.grid_v - 1)) / (cfg.grid_v - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        stop = np.where(a < 0, np.ceil(base["v_seed"] / (-a * batch.dt)), np.inf)
    if case != "many_pass_kernel":     # (that case is about the form of the sums, not the three paths)
        assert (stop <= 8).sum() > 50 and ((stop > 8) & (stop <= 16)).sum() > 10 and (stop > 16).sum() > 500
    for f in ("best_idx", "n_steps", "status", "best_cost", "x1", "y1", "theta1", "v_seed", "s_seed"):
        np.testing.assert_array_equal(rec[f], base[f])
    n_win = len(rec)
    starts = np.arange(n_win, dtype=np.int64)
    lens = np.full(n_win, cfg.window_frames + 1, np.int32)
    ref, _ = c_oracle.search(cfg.to_c(), starts, lens, np.zeros(n_win, np.int32), [batch.dt], vo, gps, imu)
    np.testing.assert_array_equal(rec["best_idx"], ref["best_idx"])
    np.testing.assert_allclose(rec["best_cost"], ref["best_cost"], rtol=1e-9, atol=1e-18)
