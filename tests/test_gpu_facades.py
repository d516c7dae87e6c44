"""The reference-shaped Python surface on the GPU: grid_run (mpc_run's signature) and
optimize_trajectory (optimize_trajectory_v2.optimize_trajectory's signature)."""
import numpy as np
import pytest

from oracle import vmvo_oracle as O
from tests.helpers import load_golden, spec_of, unhex
from vehiclemodelvisualodometry_b200 import (DEFAULT_CFG, REFERENCE_CFG, BicycleModel, SearchConfig,
                                             Trajectory, grid_run, mpc_run, optimize_trajectory)
from vehiclemodelvisualodometry_b200.synthetic import off_float32_grid, synthetic_drives

pytestmark = pytest.mark.gpu
G = load_golden()


def _traj(time, s):
    s = s.astype(np.float64)
    return Trajectory(x=s[:, 0], y=s[:, 1], theta=s[:, 2], velocity=s[:, 3], time=time)


def test_grid_run_has_mpc_run_contract(cuda_device):
    """Same arguments and return as mpc_run (vmvo/utils/mpc.py:14-20,121-122)."""
    assert mpc_run is grid_run
    batch = synthetic_drives(1, 120, seed=31)
    time, gps = batch.time[0], batch.gps[0]
    tr = _traj(time, gps)
    sub = tr.sub_trajectory_from_time(time[10], time[10] + 3.0)
    v = (sub.velocity[0] + sub.velocity[-1]) / 2
    cfg = SearchConfig(grid_v=8, grid_s=16, target_mode="traverse", seed_mode="given")
    u, rec, out = grid_run(sub, BicycleModel(), v, 25.0, 0.05, config=cfg, return_info=True)
    # oracle on the same local window: float64 end to end, nothing is rounded on the way in
    xy = np.stack([sub.x, sub.y], axis=1).astype(np.float64)
    assert not np.array_equal(xy, xy.astype(np.float32))
    interp = O.traverse_trajectory(xy, v * 0.05)
    spec = O.SearchSpec(grid_v=8, grid_s=16, target_mode="traverse", seed_mode="given")
    wt = O.WindowTargets(n_steps=len(interp) - 1, status=0, v_seed=v, s_seed=25.0, vo_xy=interp)
    ref = O.solve_window(spec, wt, 0.05)
    assert isinstance(u, np.ndarray) and u.shape == (ref.n_steps,)
    assert int(rec["best_idx"]) == ref.best_idx
    np.testing.assert_array_equal(u, ref.steer)
    np.testing.assert_allclose(rec["best_cost"], ref.best_cost, rtol=1e-9)


def test_grid_run_empty_window_returns_empty(cuda_device):
    """mpc.py:42-43: a target that decimates to one point gives np.zeros(0)."""
    n = 20
    tr = Trajectory(x=[0.0] * n, y=[0.0] * n, theta=[0.0] * n, velocity=[1.0] * n,
                    time=list(np.arange(n) * 0.05))
    u = grid_run(tr, BicycleModel(), 1.0, 0.0, 0.05)
    assert isinstance(u, np.ndarray) and u.shape == (0,)


@pytest.mark.parametrize("which", ["default", "reference"])
def test_optimize_trajectory_matches_oracle(cuda_device, which):
    n = 230
    batch = synthetic_drives(1, n, seed=41)
    # Trajectory columns are List[float] (float64): use values float32 cannot hold
    time, vo, gps = batch.time[0], off_float32_grid(batch.vo[0]), off_float32_grid(batch.gps[0])
    base = DEFAULT_CFG if which == "default" else REFERENCE_CFG
    cfg = SearchConfig(**{**base.__dict__, "grid_v": 8, "grid_s": 12})
    vo_t, gps_t = _traj(time, vo), _traj(time[:-3], gps[:-3])    # unequal lengths: N = min
    before = vo_t.model_dump()
    out = optimize_trajectory(vo_t, gps_t, BicycleModel(), config=cfg)
    assert vo_t.model_dump() == before                            # inputs are not mutated
    assert isinstance(out, Trajectory) and len(out) == n and out.time == vo_t.time
    N = n - 3
    dt, horizon, _ = O.reference_dt(time[:N])
    spec = O.replace(spec_of(cfg), horizon_frames=horizon, horizon_time=3.0)
    ref = O.optimize_drive(spec, time[:N], dt, vo[:N], gps[:N])
    np.testing.assert_allclose(out.x[:N], ref.x, rtol=0, atol=1e-9)
    np.testing.assert_allclose(out.y[:N], ref.y, rtol=0, atol=1e-9)
    np.testing.assert_array_equal(out.theta[:N], ref.theta)
    np.testing.assert_array_equal(out.velocity[:N], ref.velocity)
    np.testing.assert_array_equal(out.x[N:], vo_t.x[N:])


def test_optimize_trajectory_reproduces_frozen_reference_output(cuda_device):
    """optimize_trajectory (this package) against the frozen output of the reference's own
    optimize_trajectory loop (tests/golden, key driver_f64: float64 inputs off the float32 grid,
    grid solver patched in for mpc_run)."""
    g = G["driver_f64"]
    batch = synthetic_drives(1, g["n"], seed=g["seed"])
    time = batch.time[0]
    vo_t, gps_t = _traj(time, off_float32_grid(batch.vo[0])), _traj(time, off_float32_grid(batch.gps[0]))
    cfg = SearchConfig(**{**REFERENCE_CFG.__dict__, "grid_v": g["grid"][0], "grid_s": g["grid"][1]})
    out, (so, plan, rec) = optimize_trajectory(vo_t, gps_t, BicycleModel(), config=cfg, return_details=True)
    np.testing.assert_array_equal(rec["best_idx"], g["best_idx"])
    np.testing.assert_allclose(out.x, unhex(g["x"]), rtol=0, atol=1e-9)
    np.testing.assert_allclose(out.y, unhex(g["y"]), rtol=0, atol=1e-9)
    np.testing.assert_array_equal(out.theta, unhex(g["theta"]))
    np.testing.assert_array_equal(out.velocity, unhex(g["velocity"]))


def test_optimize_trajectory_stationary_raises_like_reference(cuda_device):
    """Quirk D7: the reference dies with IndexError on an empty solve."""
    n = 200
    t = 10.0 + np.arange(n) * 0.05
    z = np.zeros((n, 4), dtype=np.float32)
    z[:, 3] = 1.0
    tr = _traj(t, z)
    with pytest.raises(IndexError):
        optimize_trajectory(tr, tr, BicycleModel(), config=REFERENCE_CFG)
