"""CPU: the oracle against the vectors frozen from the unmodified reference
(tests/golden/reference_kats.json, written by oracle/make_golden.py), and -- when the reference
tree is present (build container only) -- against the reference itself."""
import numpy as np
import pytest

from oracle import ref_bridge
from oracle import vmvo_oracle as O
from tests.helpers import load_golden, unhex
from vehiclemodelvisualodometry_b200.synthetic import off_float32_grid, synthetic_drives

G = load_golden()
# the frozen floats come from this container's libm; other hosts may differ in the last bit
ULP = dict(rtol=4e-16, atol=1e-15)


def test_b0_reference_unittest():
    """vmvo/bicycle_model.py:110-117: zero velocity leaves x, y, theta, velocity unchanged."""
    x, y, th = O.bicycle_step(0.0, 0.0, 0.0, 30.0, 0.0, 0.1)
    assert (x, y, th) == (0.0, 0.0, 0.0)
    np.testing.assert_array_equal(unhex(G["model"]["B0"]), [0, 0, 0, 0])


def test_b1_b2_single_steps():
    x, y, th = O.bicycle_step(0.0, 0.0, 0.0, 30.0, 5.0, 0.1)
    np.testing.assert_allclose([x, y, th], unhex(G["model"]["B1"]), **ULP)
    x, y, th = O.bicycle_step(x, y, th, -460.0, 5.5, 0.1)
    np.testing.assert_allclose([x, y, th], unhex(G["model"]["B2"]), **ULP)


def test_b3_sequence():
    b3 = G["model"]["B3"]
    got = O.rollout(unhex(b3["steer"]), unhex(b3["vel"]), b3["dt"], (0, 0, 0, b3["v0"]))
    np.testing.assert_allclose(got, unhex(b3["poses"], (30, 3)), rtol=1e-14, atol=1e-14)
    # SURVEY.md Appendix B, last state
    np.testing.assert_allclose(got[-1], [14.940522636271067, 10.280875911767795, 1.672877377123603],
                               rtol=1e-13)


def test_b4_asserts():
    assert G["model"]["B4"] == ["Steering angle is out of bounds", "Acceleration is out of bounds"]
    with pytest.raises(AssertionError, match="Steering angle is out of bounds"):
        O.rollout([461.0], [0.0], 0.1)
    with pytest.raises(AssertionError, match="Acceleration is out of bounds"):
        O.rollout([0.0], [30.0], 0.05, (0, 0, 0, 10.0))


def test_b5_sub_trajectory_from_time():
    b5 = G["schema"]["B5"]
    t = np.array([0, .1, .2, .3, .4])
    s, e = O.window_extent_time(t, 0.1, 0.3)
    assert [s, e] == b5["extent"]
    lx, ly, lth = O.local_frame(np.array([0, 1, 2, 3, 4.])[s:e], np.array([0, 0, 1, 1, 2.])[s:e],
                                np.array([.5, .5, .6, .7, .8])[s:e])
    np.testing.assert_allclose(lx, unhex(b5["x"]), rtol=0, atol=1e-15)
    np.testing.assert_allclose(ly, unhex(b5["y"]), rtol=0, atol=1e-15)
    np.testing.assert_array_equal(lth, unhex(b5["theta"]))
    with pytest.raises(AssertionError, match="No frames found"):
        O.window_extent_time(t, 5.0, 6.0)
    np.testing.assert_array_equal(O.rollout_times(12.5, 0.05, 7), unhex(G["schema"]["B5_times"]))


def test_b6_traverse():
    b6 = G["traverse"]["B6"]
    xy = unhex(b6["xy"], (-1, 2))
    np.testing.assert_array_equal(O.traverse_indices(xy, b6["D"]), b6["keep"])
    np.testing.assert_array_equal(O.traverse_trajectory(xy, b6["D"]), unhex(b6["out"], (-1, 2)))
    for case in G["traverse"]["random"]:
        xy = unhex(case["xy"], (-1, 2))
        np.testing.assert_array_equal(O.traverse_indices(xy, float.fromhex(case["D"])), case["keep"])


def test_cost_closure():
    for case in G["cost"]:
        got = O.sequence_cost(unhex(case["u"]), float.fromhex(case["v"]), case["dt"],
                              unhex(case["target"], (-1, 2)), float.fromhex(case["K"]))
        np.testing.assert_allclose(got, float.fromhex(case["cost"]), rtol=1e-14)


def test_grid_cases_pinned_through_reference_model():
    batch = synthetic_drives(2, 400, seed=3)
    for case in G["grid"]:
        gv, gs = case["grid"]
        prim = case["primary"]
        spec = O.SearchSpec(grid_v=gv, grid_s=gs, window_frames=case["W"], target_mode=case["target_mode"],
                            primary=prim, w_vo=1.0 if prim == "vo" else 0.0,
                            w_gps=0.0 if prim == "vo" else 1.0)
        time, vo, gps, imu = batch.drive(case["drive"])
        wt = O.build_window(spec, case["start"], case["W"] + 1, batch.dt, vo, gps, None)
        res = O.solve_window(spec, wt, batch.dt)
        assert res.best_idx == case["best_idx"] and res.n_steps == case["n_steps"]
        np.testing.assert_allclose(res.best_cost, float.fromhex(case["best_cost"]), rtol=1e-13)
        np.testing.assert_allclose(res.poses[0], unhex(case["first_pose"]), rtol=0, atol=1e-14)
        np.testing.assert_allclose(wt.s_seed, float.fromhex(case["s_seed"]), rtol=1e-14, atol=1e-13)


@pytest.mark.parametrize("key", ["driver", "driver_f64"])
def test_driver_loop_golden(key):
    g = G[key]
    batch = synthetic_drives(1, g["n"], seed=g["seed"])
    time, vo, gps, imu = batch.drive(0)
    if key == "driver_f64":      # inputs float32 cannot represent
        vo, gps = off_float32_grid(vo), off_float32_grid(gps)
    dt, horizon, _ = O.reference_dt(time)
    assert horizon == g["horizon"] and dt == float.fromhex(g["dt"])
    spec = O.SearchSpec(grid_v=g["grid"][0], grid_s=g["grid"][1], window_mode="time", target_mode="traverse",
                        primary="gps", w_vo=0.0, w_gps=1.0, horizon_frames=horizon)
    res = O.optimize_drive(spec, time, dt, vo, gps)
    assert [r.best_idx for r in res.windows] == g["best_idx"]
    np.testing.assert_allclose(res.x, unhex(g["x"]), rtol=0, atol=1e-12)
    np.testing.assert_allclose(res.y, unhex(g["y"]), rtol=0, atol=1e-12)
    np.testing.assert_array_equal(res.theta, unhex(g["theta"]))
    np.testing.assert_array_equal(res.velocity, unhex(g["velocity"]))


def test_grid_axis_is_antisymmetric_and_spans_the_band():
    for g in (1, 2, 5, 32, 255, 256):
        a = O.grid_axis(10.0, g)
        np.testing.assert_array_equal(a, -a[::-1])
        assert g == 1 or (a[0] == -10.0 and a[-1] == 10.0)


def test_hypotheses_respect_the_reference_bounds():
    """Per-step |dv/dt| <= A and |ds/dt| <= S' (bicycle_model.py:48-62, mpc.py:96-104)."""
    spec = O.SearchSpec(grid_v=9, grid_s=9)
    V, S = O.hypothesis_controls(spec, 4.0, 300.0, 40, 0.05)
    dv = np.diff(np.concatenate([np.full((1, 9), 4.0), V]), axis=0) / 0.05
    ds = np.diff(np.concatenate([np.full((1, 9), 300.0), S]), axis=0) / 0.05
    assert np.all(np.abs(dv) <= 10 * (1 + 1e-9)) and np.all(V >= 0)
    assert np.all(np.abs(ds) <= 100 * (1 + 1e-9)) and np.all(np.abs(S) <= 460.0)


def test_nonfinite_and_empty_windows():
    spec = O.SearchSpec(grid_v=4, grid_s=4, window_frames=10)
    batch = synthetic_drives(1, 40, seed=1)
    vo = batch.vo[0].copy()
    vo[5, 0] = np.nan
    wt = O.build_window(spec, 0, 11, 0.05, vo, None, None)
    res = O.solve_window(spec, wt, 0.05)
    assert res.status & O.WIN_NONFINITE and res.best_idx == 0 and np.isnan(res.best_cost)
    spec_t = O.replace(spec, target_mode="traverse")
    z = np.zeros((20, 4), dtype=np.float32)
    z[:, 3] = 1.0
    res = O.solve_window(spec_t, O.build_window(spec_t, 0, 11, 0.05, z, None, None), 0.05)
    assert res.status & O.WIN_EMPTY and res.n_steps == 0 and res.best_idx == -1


@pytest.mark.skipif(not ref_bridge.available(), reason="reference tree only exists in the build container")
def test_restatement_against_live_reference():
    ref = ref_bridge.load()
    rng = np.random.default_rng(0)
    St = ref.schema.State
    for _ in range(5):
        N = int(rng.integers(1, 50))
        v0 = float(rng.uniform(0, 20))
        steer = rng.uniform(-460, 460, N)
        vel = np.clip(v0 + np.cumsum(rng.uniform(-0.4, 0.4, N)), 0, None)
        m = ref.bicycle_model.BicycleModel(state=St(x=1.0, y=-2.0, theta=0.3, velocity=v0, steering_angle=0))
        st = m.run_sequence(steer, vel, 0.05)
        got = O.rollout(steer, vel, 0.05, (1.0, -2.0, 0.3, v0))
        np.testing.assert_array_equal(np.array([[s.x, s.y, s.theta] for s in st]), got)
        p = np.cumsum(rng.normal(0, 0.3, (N + 1, 2)), axis=0)
        D = float(rng.uniform(0.05, 1))
        np.testing.assert_array_equal(ref.mpc.traverse_trajectory(p, D), O.traverse_trajectory(p, D))


def test_c_restatement_equals_numpy_oracle():
    """oracle/vmvo_oracle.c (the fast checker / CPU baseline) against oracle/vmvo_oracle.py."""
    from oracle import c_oracle
    from tests.helpers import oracle_windows, spec_of
    from vehiclemodelvisualodometry_b200 import SearchConfig

    cases = [
        SearchConfig(grid_v=16, grid_s=16, window_frames=20),
        SearchConfig(grid_v=8, grid_s=12, window_frames=24, w_vo=1.0, w_gps=0.5, w_imu=40.0),
        SearchConfig(grid_v=8, grid_s=24, window_frames=40, target_mode="traverse", primary="gps",
                     w_vo=0.0, w_gps=1.0),
        SearchConfig(grid_v=5, grid_s=7, window_frames=20, k_steer=5e-6, target_offset=0),
    ]
    for cfg in cases:
        n = 2 * cfg.horizon() + 30
        b = synthetic_drives(1, n, seed=5)
        st, ln = O.window_extents(spec_of(cfg), b.time[0])
        rec, steps = c_oracle.search(cfg.to_c(), st, ln, np.zeros(len(st), np.int32), [b.dt], b.vo[0],
                                     b.gps[0], b.imu[0], n_threads=2)
        ref = oracle_windows(cfg, b.time[0], b.dt, b.vo[0], b.gps[0], b.imu[0])
        np.testing.assert_array_equal(rec["best_idx"], [r.best_idx for r in ref])
        np.testing.assert_array_equal(rec["n_steps"], [r.n_steps for r in ref])
        np.testing.assert_allclose(rec["best_cost"], [r.best_cost for r in ref], rtol=1e-13)
        np.testing.assert_allclose(rec["x1"], [r.poses[0, 0] for r in ref], rtol=0, atol=1e-14)
        assert steps == cfg.grid_v * cfg.grid_s * sum(r.n_steps for r in ref)
