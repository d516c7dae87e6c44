"""BASELINE.json's full-size configurations on the GPU, checked against the C restatement of the
oracle (oracle/vmvo_oracle.c; itself pinned to the NumPy oracle, which is pinned to the
reference) and through size-independent properties: determinism, shard invariance, totals."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from vehiclemodelvisualodometry_b200 import DriveSet, SearchConfig, _lib, grid_search, plan_windows
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

pytestmark = pytest.mark.gpu


def _check(cfg, batch, n_drives, sample=None, with_gps=False, with_imu=False):
    drives = DriveSet.from_arrays(list(batch.time[:n_drives]), [batch.dt] * n_drives,
                                  vo=list(batch.vo[:n_drives]),
                                  gps=list(batch.gps[:n_drives]) if with_gps else None,
                                  imu=list(batch.imu[:n_drives]) if with_imu else None)
    plan = plan_windows(cfg, drives)
    rec = grid_search(cfg, drives, plan).records()
    ws, wl, wd = (t.cpu().numpy() for t in (plan.win_start, plan.win_len, plan.win_drive))
    sel = np.arange(plan.n_windows) if sample is None else \
        np.unique(np.linspace(0, plan.n_windows - 1, sample).astype(np.int64))
    n = batch.vo.shape[1]
    vo = batch.vo[:n_drives].reshape(n_drives * n, 4)
    gps = batch.gps[:n_drives].reshape(n_drives * n, 4) if with_gps else None
    imu = batch.imu[:n_drives].reshape(n_drives * n) if with_imu else None
    ref, _ = c_oracle.search(cfg.to_c(), ws[sel], wl[sel], wd[sel], [batch.dt] * n_drives, vo, gps, imu)
    bad = np.nonzero(rec["best_idx"][sel] != ref["best_idx"])[0]
    detail = [(int(sel[b]), int(rec["best_idx"][sel[b]]), int(ref["best_idx"][b]), float(rec["best_cost"][sel[b]]),
               float(ref["best_cost"][b]), int(rec["n_rescored"][sel[b]]), int(rec["status"][sel[b]]),
               float(rec["v_seed"][sel[b]]), float(ref["v_seed"][b])) for b in bad[:5]]
    assert len(bad) == 0, (f"{len(bad)} of {len(sel)} argmin mismatches; (window, got, want, got cost, want cost, "
                           f"rescored, status, got v_seed, want v_seed): {detail}")
    np.testing.assert_array_equal(rec["n_steps"][sel], ref["n_steps"])
    np.testing.assert_array_equal(rec["status"][sel], ref["status"])
    np.testing.assert_allclose(rec["best_cost"][sel], ref["best_cost"], rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(rec["x1"][sel], ref["x1"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rec["y1"][sel], ref["y1"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rec["theta1"][sel], ref["theta1"], rtol=0, atol=1e-9)
    return drives, plan, rec


def test_config2_single_drive_10k_frames_every_window(cuda_device):
    """configs[1]: 10 000 frames, 32x32, W = 30 -- all 9 940 windows against the CPU oracle."""
    cfg = SearchConfig(grid_v=32, grid_s=32, window_frames=30)
    batch = synthetic_drives(1, 10000, seed=1658384707877 % 2 ** 32)
    drives, plan, rec = _check(cfg, batch, 1)
    assert plan.n_windows == 9940 and int(rec["n_steps"].sum()) * 1024 == 305356800
    # determinism: a second run gives identical bytes; shards give the same records
    again = grid_search(cfg, drives, plan).results
    first = grid_search(cfg, drives, plan).results
    # (bytes 12..15 are n_rescored, a diagnostic: which warp re-scores which candidate follows the
    # order of an atomic counter, and a warp prunes with the float64 costs it has seen itself, so
    # the COUNT of re-scores may differ between runs; everything else is bit-reproducible)
    assert torch.equal(again[:, :12], first[:, :12]) and torch.equal(again[:, 16:], first[:, 16:])
    buf = torch.zeros_like(first)
    for lo, hi in ((0, 3313), (3313, 3314), (3314, 9940)):
        grid_search(cfg, drives, plan, window_range=(lo, hi), out=buf[lo:hi])
    a = first.cpu().numpy().view(_lib.RESULT_DTYPE).reshape(-1)
    b = buf.cpu().numpy().view(_lib.RESULT_DTYPE).reshape(-1)
    for f in ("best_idx", "n_steps", "status", "best_cost", "x1", "y1", "theta1", "v_seed", "s_seed"):
        np.testing.assert_array_equal(a[f], b[f])


def test_config3_dense_grid_sample(cuda_device):
    """configs[2]: 256x256 grid, W = 60 -- 48 windows against the CPU oracle."""
    cfg = SearchConfig(grid_v=256, grid_s=256, window_frames=60)
    batch = synthetic_drives(1, 700, seed=33)
    _check(cfg, batch, 1, sample=48)


def test_config4_batch_of_drives(cuda_device):
    """configs[3] in miniature: 48 drives, 32x32, W = 60, every 7th window checked."""
    cfg = SearchConfig(grid_v=32, grid_s=32, window_frames=60)
    batch = synthetic_drives(48, 400, seed=44)
    drives, plan, rec = _check(cfg, batch, 48, sample=1900)
    assert plan.n_windows == 48 * 280
    assert np.all(np.diff(plan.win_drive.cpu().numpy()) >= 0)


def test_config5_fused_vo_gps_imu(cuda_device):
    """configs[4] in miniature: VO + GPS + IMU cost, 128x128, W = 60 (IMU term: parity unpinned
    against the reference -- it has no IMU code -- but pinned to the oracle's spec)."""
    cfg = SearchConfig(grid_v=128, grid_s=128, window_frames=60, w_vo=1.0, w_gps=0.05, w_imu=30.0)
    batch = synthetic_drives(4, 200, seed=55)
    _check(cfg, batch, 4, sample=96, with_gps=True, with_imu=True)


@pytest.mark.parametrize("shape", ["config2", "config3"])
def test_full_size_pruned_search_equals_exhaustive_search(cuda_device, tuning, shape):
    """A size-independent property at BASELINE's full sizes: the pruning votes and the specialised kernels
    change no byte of any record (all 9 940 windows of configs[1]; 4 096 windows of configs[2])."""
    if shape == "config2":
        cfg, frames = SearchConfig(grid_v=32, grid_s=32, window_frames=30), 10000
    else:
        cfg, frames = SearchConfig(grid_v=256, grid_s=256, window_frames=60), 4216
    batch = synthetic_drives(1, frames, seed=77)
    drives = DriveSet.from_arrays([batch.time[0]], [batch.dt], vo=[batch.vo[0]])
    plan = plan_windows(cfg, drives)
    fast = grid_search(cfg, drives, plan).results.clone()
    tuning("prune", 0)
    tuning("lean", 0)
    full = grid_search(cfg, drives, plan).results.clone()
    # (bytes 12..15: n_rescored, a run-to-run diagnostic -- see above)
    assert torch.equal(fast[:, :12], full[:, :12]) and torch.equal(fast[:, 16:], full[:, 16:])
