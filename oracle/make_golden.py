"""Pins the oracle to the UNMODIFIED reference and freezes golden vectors.

TEST INFRASTRUCTURE -- run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # checks + writes tests/golden/*.json

1. Every restated function of oracle/vmvo_oracle.py is compared with the imported
   reference function it restates (bit-exact, except the np.dot-based local frame: 1e-12).
2. The derived grid search is pinned through the reference primitives: for small grids
   every hypothesis is rolled out with the real ``BicycleModel.run`` and scored with a
   literal transcription of the ``cost`` closure (vmvo/utils/mpc.py:68-80); the argmin and
   costs must equal the vectorised oracle's bit for bit.
3. The reference's own sliding-window driver (optimize_trajectory_v2.py:24-148) is run with
   ``mpc_run`` monkey-patched to the oracle's grid solver, which pins the window rule,
   write-back order and blends (a12) against the reference loop itself.
The outputs are frozen as JSON (floats as hex strings, so nothing is lost).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_bridge  # noqa: E402
from oracle import vmvo_oracle as O  # noqa: E402
from vehiclemodelvisualodometry_b200.synthetic import off_float32_grid, synthetic_drives  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def hexf(a):
    a = np.asarray(a, dtype=np.float64)
    return [float(v).hex() for v in a.reshape(-1)]


def ref_model(ref, v0=0.0, relax=False):
    St = ref.schema.State
    kw = {}
    if relax:  # band-edge hypotheses: (v' - v)/dt can round just above MAX_ACCEL
        kw["max_accel"] = ref.constants.MAX_ACCEL * (1 + 1e-9)
    return ref.bicycle_model.BicycleModel(
        state=St(x=0.0, y=0.0, theta=0.0, velocity=v0, steering_angle=0.0), **kw)


def kat_model(ref):
    """KATs B0-B4 of SURVEY.md Appendix B."""
    out = {}
    m = ref_model(ref)
    s = m.run(30.0, 0.0, 0.1)
    out["B0"] = hexf([s.x, s.y, s.theta, s.velocity])
    m = ref_model(ref, 5.0)
    s = m.run(30.0, 5.0, 0.1)
    out["B1"] = hexf([s.x, s.y, s.theta])
    s = m.run(-460.0, 5.5, 0.1)
    out["B2"] = hexf([s.x, s.y, s.theta])
    m = ref_model(ref, 10.0)
    steers = [10.0 * k for k in range(1, 31)]
    vels = [10 + 0.25 * k for k in range(1, 31)]
    states = m.run_sequence(steers, vels, 0.05)
    ref_poses = np.array([[s.x, s.y, s.theta] for s in states])
    mine = O.rollout(steers, vels, 0.05, (0, 0, 0, 10.0))
    assert np.array_equal(ref_poses, mine), "rollout restatement differs from BicycleModel.run_sequence"
    out["B3"] = {"steer": hexf(steers), "vel": hexf(vels), "dt": 0.05, "v0": 10.0, "poses": hexf(ref_poses)}
    msgs = []
    for args, v0 in (((461.0, 0.0, 0.1), 0.0), ((0.0, 30.0, 0.05), 10.0)):
        try:
            ref_model(ref, v0).run(*args)
            msgs.append(None)
        except AssertionError as e:
            msgs.append(str(e))
    assert msgs == ["Steering angle is out of bounds", "Acceleration is out of bounds"], msgs
    out["B4"] = msgs
    for args, v0 in (((461.0, 0.0, 0.1), 0.0), ((0.0, 30.0, 0.05), 10.0)):
        try:
            O.rollout([args[0]], [args[1]], args[2], (0, 0, 0, v0))
            raise SystemExit("oracle rollout did not raise")
        except AssertionError as e:
            assert str(e) in msgs
    return out


def kat_schema(ref):
    T = ref.schema.Trajectory
    tr = T(x=[0, 1, 2, 3, 4], y=[0, 0, 1, 1, 2], theta=[.5, .5, .6, .7, .8], velocity=[1] * 5,
           time=[0, .1, .2, .3, .4])
    sub = tr.sub_trajectory_from_time(0.1, 0.3)
    s, e = O.window_extent_time(np.array(tr.time), 0.1, 0.3)
    assert (s, e) == (1, 4)
    lx, ly, lth = O.local_frame(tr.x[s:e], tr.y[s:e], tr.theta[s:e])
    assert np.allclose(lx, sub.x, rtol=0, atol=1e-12) and np.allclose(ly, sub.y, rtol=0, atol=1e-12)
    assert np.array_equal(lth, sub.theta)
    out = {"B5": {"x": hexf(sub.x), "y": hexf(sub.y), "theta": hexf(sub.theta), "time": hexf(sub.time),
                  "extent": [s, e]}}
    # a larger randomised check of the local frame and the extents
    rng = np.random.default_rng(5)
    n = 200
    tr = T(x=np.cumsum(rng.normal(0, 1, n)), y=np.cumsum(rng.normal(0, 1, n)),
           theta=rng.uniform(-3, 3, n), velocity=rng.uniform(0, 10, n),
           time=np.cumsum(rng.uniform(0.01, 0.1, n)))
    t = np.array(tr.time)
    for _ in range(50):
        a = rng.uniform(t[0], t[-1])
        b = a + rng.uniform(0, 3)
        sub = tr.sub_trajectory_from_time(a, b) if b >= a else None
        s, e = O.window_extent_time(t, a, b)
        lx, ly, lth = O.local_frame(np.array(tr.x)[s:e], np.array(tr.y)[s:e], np.array(tr.theta)[s:e])
        assert len(sub) == e - s
        assert np.allclose(lx, sub.x, rtol=0, atol=1e-11) and np.allclose(ly, sub.y, rtol=0, atol=1e-11)
        assert np.array_equal(lth, sub.theta)
    # states_list_to_trajectory stamps (a13)
    St = ref.schema.State
    states = [St(x=i, y=0, theta=0, velocity=1, steering_angle=0) for i in range(7)]
    tt = ref.schema.states_list_to_trajectory(states, 12.5, 0.05)
    assert np.array_equal(tt.time, O.rollout_times(12.5, 0.05, 7))
    out["B5_times"] = hexf(tt.time)
    return out


def kat_traverse(ref):
    xy = np.array([[0, 0], [.3, 0], [.6, 0], [.9, 0], [1.2, 0], [1.5, 0]], dtype=np.float64)
    r = ref.mpc.traverse_trajectory(xy, 0.5)
    assert np.array_equal(r, O.traverse_trajectory(xy, 0.5))
    out = {"B6": {"xy": hexf(xy), "D": 0.5, "keep": O.traverse_indices(xy, 0.5).tolist(), "out": hexf(r)}}
    rng = np.random.default_rng(7)
    cases = []
    for c in range(20):
        n = int(rng.integers(2, 80))
        p = np.cumsum(rng.normal(0, 0.3, (n, 2)), axis=0)
        D = float(rng.uniform(0.05, 1.0))
        r = ref.mpc.traverse_trajectory(p, D)
        keep = O.traverse_indices(p, D)
        assert np.array_equal(r, p[keep])
        if c < 4:
            cases.append({"xy": hexf(p), "D": float(D).hex(), "keep": keep.tolist()})
    out["random"] = cases
    return out


def ref_cost_closure(ref, u, v, dt, interp, K=0.0):
    """Literal transcription of ``cost`` (vmvo/utils/mpc.py:56-80) driven by the real model."""
    model = ref_model(ref, v, relax=True)
    St = ref.schema.State
    N = len(interp) - 1
    x = np.array([interp[0, 0], interp[0, 1], 0.0, v])
    cost_val = 0.0
    for i in range(N):
        model.set_state(St(x=x[0], y=x[1], theta=x[2], velocity=x[3], steering_angle=u[i]))
        ns = model.run(u[i], x[3], dt)
        x = np.array([ns.x, ns.y, ns.theta, ns.velocity])
        cost_val += (x[0] - interp[i, 0]) ** 2 + (x[1] - interp[i, 1]) ** 2 + K * u[i] ** 2
    return cost_val


def kat_cost(ref):
    rng = np.random.default_rng(11)
    out = []
    for _ in range(6):
        N = int(rng.integers(3, 40))
        tgt = np.cumsum(np.abs(rng.normal(0.4, 0.1, (N + 1, 2))), axis=0)
        tgt -= tgt[0]
        u = rng.uniform(-460, 460, N)
        v = float(rng.uniform(0, 20))
        K = float(rng.choice([0.0, 5e-6]))
        c_ref = ref_cost_closure(ref, u, v, 0.05, tgt, K)
        c_mine = O.sequence_cost(u, v, 0.05, tgt, K)
        assert c_ref == c_mine, (c_ref, c_mine)
        out.append({"u": hexf(u), "v": v.hex(), "dt": 0.05, "K": K.hex(), "target": hexf(tgt),
                    "cost": float(c_ref).hex()})
    return out


def ref_grid_window(ref, spec, wt, dt):
    """Every hypothesis rolled with the real BicycleModel.run, scored like the closure."""
    V, S = O.hypothesis_controls(spec, wt.v_seed, wt.s_seed, wt.n_steps, dt)
    T = wt.vo_xy if spec.w_vo else wt.gps_xy
    St = ref.schema.State
    cost = np.empty((spec.grid_v, spec.grid_s))
    for i in range(spec.grid_v):
        for j in range(spec.grid_s):
            m = ref_model(ref, wt.v_seed, relax=True)
            m.set_state(St(x=0.0, y=0.0, theta=0.0, velocity=wt.v_seed, steering_angle=0.0))
            c = 0.0
            for k in range(1, wt.n_steps + 1):
                s = m.run(S[k - 1, j], V[k - 1, i], dt)
                t = k - spec.target_offset
                c += (s.x - T[t, 0]) ** 2 + (s.y - T[t, 1]) ** 2 + 0.0 * S[k - 1, j] ** 2
            cost[i, j] = c
    return cost


def kat_grid(ref):
    """Derived spec pinned through the reference primitives (small grids)."""
    batch = synthetic_drives(2, 400, seed=3)
    out = []
    for case, (gv, gs, W, mode, prim) in enumerate(
            [(5, 7, 12, "time", "vo"), (8, 8, 20, "traverse", "gps"), (4, 4, 10, "time", "vo")]):
        spec = O.SearchSpec(grid_v=gv, grid_s=gs, window_frames=W, target_mode=mode, primary=prim,
                            w_vo=1.0 if prim == "vo" else 0.0, w_gps=0.0 if prim == "vo" else 1.0)
        time, vo, gps, imu = batch.drive(case % 2)
        for start in (0, 57, 203):
            wt = O.build_window(spec, start, W + 1, batch.dt, vo, gps, None)
            if wt.n_steps == 0:
                continue
            cost_ref = ref_grid_window(ref, spec, wt, batch.dt)
            res, cost = O.solve_window(spec, wt, batch.dt, want_costs=True)
            assert np.array_equal(cost_ref, cost), "vectorised grid costs differ from the reference model"
            assert res.best_idx == int(np.argmin(cost_ref.reshape(-1)))
            # the re-rollout of the optimum equals BicycleModel.run_sequence (…v2.py:76-91)
            m = ref_model(ref, wt.v_seed, relax=True)
            st = m.run_sequence(res.steer, res.vel, batch.dt)
            assert np.array_equal(np.array([[s.x, s.y, s.theta] for s in st]), res.poses)
            out.append({"grid": [gv, gs], "W": W, "target_mode": mode, "primary": prim, "drive": case % 2,
                        "start": start, "seed": 3, "best_idx": res.best_idx,
                        "best_cost": float(res.best_cost).hex(), "n_steps": res.n_steps,
                        "v_seed": float(wt.v_seed).hex(), "s_seed": float(wt.s_seed).hex(),
                        "first_pose": hexf(res.poses[0])})
    return out


def kat_driver(ref, f64_inputs=False):
    """The reference loop with mpc_run replaced by the oracle's grid solver (pins a12).

    ``f64_inputs``: pose values that float32 cannot represent (synthetic.off_float32_grid), the
    case the float64 stream entry points (vmvo_grid_search_f64 / vmvo_write_back_f64) exist for.
    """
    v2 = ref_bridge.load_v2()
    T = ref.schema.Trajectory
    n = 260
    batch = synthetic_drives(1, n, seed=9)
    time, vo, gps, imu = batch.drive(0)
    if f64_inputs:
        vo, gps = off_float32_grid(vo), off_float32_grid(gps)
        assert not np.array_equal(vo, vo.astype(np.float32).astype(np.float64))
    vo64, gps64 = vo.astype(np.float64), gps.astype(np.float64)
    # G_v = 1: the reference loop re-rolls the returned steering at CONSTANT speed
    # (optimize_trajectory_v2.py:84-91), which only a zero-acceleration hypothesis reproduces
    spec = O.SearchSpec(grid_v=1, grid_s=15, window_mode="time", target_mode="traverse", primary="gps",
                        w_vo=0.0, w_gps=1.0)
    dt, horizon, fps = O.reference_dt(time)
    spec = O.replace(spec, horizon_frames=horizon)

    def grid_mpc(trajectory, bicycle_model, velocity, starting_steering_angle, time_step):
        # the window arrives already in its local frame; build targets exactly like mpc_run
        xy = trajectory.to_numpy()[:, [0, 1]]
        interp = ref.mpc.traverse_trajectory(xy, velocity * time_step)
        wt = O.WindowTargets(n_steps=len(interp) - 1, status=0, v_seed=float(velocity),
                             s_seed=float(grid_mpc.s_seed(trajectory, velocity, time_step)),
                             gps_xy=interp)
        res = O.solve_window(spec, wt, time_step)
        return res.steer

    def s_seed(trajectory, velocity, time_step):
        return O.seed_from_window(spec, np.asarray(trajectory.theta), np.asarray(trajectory.velocity),
                                  time_step)[1]

    grid_mpc.s_seed = s_seed
    v2.mpc_run = grid_mpc
    vo_t = T(x=vo64[:, 0], y=vo64[:, 1], theta=vo64[:, 2], velocity=vo64[:, 3], time=time)
    gps_t = T(x=gps64[:, 0], y=gps64[:, 1], theta=gps64[:, 2], velocity=gps64[:, 3], time=time)
    with ref_bridge.quiet():
        ref_out = v2.optimize_trajectory(vo_t, gps_t, ref.bicycle_model.BicycleModel())
    mine = O.optimize_drive(spec, time, dt, vo, gps)
    # np.dot in the reference's local frame may differ from the written-out form by 1 ulp
    for name, col in (("x", mine.x), ("y", mine.y)):
        assert np.allclose(getattr(ref_out, name), col, rtol=0, atol=1e-9), name
    assert np.array_equal(ref_out.theta, mine.theta)
    assert np.array_equal(ref_out.velocity, mine.velocity)
    return {"n": n, "seed": 9, "grid": [1, 15], "horizon": horizon, "dt": float(dt).hex(),
            "best_idx": [r.best_idx for r in mine.windows],
            "x": hexf(mine.x), "y": hexf(mine.y), "theta": hexf(mine.theta), "velocity": hexf(mine.velocity)}


def main():
    assert ref_bridge.available(), "reference tree not found"
    ref = ref_bridge.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    golden = {
        "generator": "oracle/make_golden.py (reference imported unmodified from /root/reference)",
        "numpy": np.__version__,
        "model": kat_model(ref),
        "schema": kat_schema(ref),
        "traverse": kat_traverse(ref),
        "cost": kat_cost(ref),
        "grid": kat_grid(ref),
        "driver": kat_driver(ref),
        "driver_f64": kat_driver(ref, f64_inputs=True),
    }
    path = os.path.join(GOLDEN_DIR, "reference_kats.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
