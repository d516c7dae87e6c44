"""GPU parity of the operators either side of the search: model rollout (a1/a2), window
extraction (a7/a8), decimation (a9), sequence cost (a10), write-back + blends (a12), against
the golden vectors frozen from the reference and against the oracle."""
import numpy as np
import pytest
import torch

from oracle import vmvo_oracle as O
from tests.helpers import load_golden, spec_of, unhex
from vehiclemodelvisualodometry_b200 import (BicycleModel, DriveSet, SearchConfig, State, Trajectory,
                                             optimize_drives, rollout_batch, sequence_cost,
                                             traverse_trajectory)
from vehiclemodelvisualodometry_b200 import _lib
from vehiclemodelvisualodometry_b200.synthetic import off_float32_grid, synthetic_drives

pytestmark = pytest.mark.gpu
G = load_golden()


# ---- a1 / a2 -----------------------------------------------------------------------------------
def test_reference_unittest_zero_velocity_leaves_state(cuda_device):
    """The reference's only test (vmvo/bicycle_model.py:110-117)."""
    m = BicycleModel()
    old = State(**m.state.model_dump())
    s = m.run(steering_angle=30.0, velocity=0.0, dt=0.1)
    assert (s.x, s.y, s.theta, s.velocity) == (old.x, old.y, old.theta, old.velocity)
    assert s.steering_angle == 30.0 and m.state is s


def test_kat_single_steps(cuda_device):
    m = BicycleModel(state=State(x=0, y=0, theta=0, velocity=5.0, steering_angle=0))
    s = m.run(30.0, 5.0, 0.1)
    np.testing.assert_allclose([s.x, s.y, s.theta], unhex(G["model"]["B1"]), rtol=0, atol=1e-14)
    s = m.run(-460.0, 5.5, 0.1)
    np.testing.assert_allclose([s.x, s.y, s.theta], unhex(G["model"]["B2"]), rtol=0, atol=1e-14)


def test_kat_sequence_f64_and_f32(cuda_device):
    b3 = G["model"]["B3"]
    steer, vel = unhex(b3["steer"]), unhex(b3["vel"])
    want = unhex(b3["poses"], (30, 3))
    m = BicycleModel(state=State(x=0, y=0, theta=0, velocity=b3["v0"], steering_angle=0))
    states = m.run_sequence(steer, vel, b3["dt"])
    got = np.array([[s.x, s.y, s.theta] for s in states])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    assert m.state is states[-1] and states[-1].velocity == vel[-1]
    # float32 rollout: the north-star tolerance, 1e-4 m and 1e-5 rad
    s32 = torch.tensor(steer, dtype=torch.float32, device=cuda_device)[None]
    v32 = torch.tensor(vel, dtype=torch.float32, device=cuda_device)[None]
    st0 = torch.tensor([[0, 0, 0, b3["v0"]]], dtype=torch.float32, device=cuda_device)
    p32, fail = rollout_batch(s32, v32, b3["dt"], st0)
    p32 = p32[0].cpu().numpy().astype(np.float64)
    assert fail[0, 0].item() == 0
    assert np.max(np.abs(p32[:, :2] - want[:, :2])) <= 1e-4
    assert np.max(np.abs(p32[:, 2] - want[:, 2])) <= 1e-5


def test_bounds_raise_reference_messages(cuda_device):
    with pytest.raises(AssertionError, match="Steering angle is out of bounds"):
        BicycleModel().run(461.0, 0.0, 0.1)
    m = BicycleModel(state=State(x=0, y=0, theta=0, velocity=10.0, steering_angle=0))
    with pytest.raises(AssertionError, match="Acceleration is out of bounds"):
        m.run(0.0, 30.0, 0.05)
    # a violation in the middle of a sequence: earlier steps were applied, as in the reference
    m = BicycleModel(state=State(x=0, y=0, theta=0, velocity=1.0, steering_angle=0))
    with pytest.raises(AssertionError, match="Steering angle"):
        m.run_sequence([1.0, 2.0, 500.0, 3.0], [1.0, 1.0, 1.0, 1.0], 0.1)
    ref = O.rollout([1.0, 2.0], [1.0, 1.0], 0.1, (0, 0, 0, 1.0))
    np.testing.assert_allclose([m.state.x, m.state.y, m.state.theta], ref[-1], rtol=0, atol=1e-14)


def test_rollout_batch_random(cuda_device):
    rng = np.random.default_rng(8)
    B, N = 257, 75          # three 32-step rounds, ragged last round
    steer = rng.uniform(-460, 460, (B, N))
    vel = np.abs(np.cumsum(rng.uniform(-0.4, 0.4, (B, N)), axis=1) + 8)
    st0 = np.concatenate([rng.normal(0, 5, (B, 3)), vel[:, :1]], axis=1)
    poses, fail = rollout_batch(torch.tensor(steer, device=cuda_device), torch.tensor(vel, device=cuda_device),
                                0.05, torch.tensor(st0, device=cuda_device))
    poses = poses.cpu().numpy()
    assert not fail[:, 0].any()
    for b in range(0, B, 17):
        ref = O.rollout(steer[b], vel[b], 0.05, st0[b])
        np.testing.assert_allclose(poses[b], ref, rtol=0, atol=1e-10)


def test_host_rollout_is_the_tensor_rollout(cuda_device):
    """vmvo_rollout_host_f64 (what BicycleModel.run / run_sequence call: host buffers, pinned staging,
    one launch) against vmvo_rollout_f64 through tensors: the same kernel, the same bits; sequences
    that outgrow the staging buffer; the (kind, step) of a violated bound."""
    rng = np.random.default_rng(81)
    for n in (1, 2, 31, 32, 33, 100, 9000, 5):
        steer = rng.uniform(-460, 460, n)
        vel = np.abs(np.cumsum(rng.uniform(-0.4, 0.4, n)) + 8)
        st0 = np.array([1.5, -2.0, 0.3, vel[0]])
        got, kind, step = _lib.rollout_host_f64(steer, vel, 0.05, st0, 460.0, 10.0)
        want, fail = rollout_batch(torch.tensor(steer[None], device=cuda_device),
                                   torch.tensor(vel[None], device=cuda_device), 0.05,
                                   torch.tensor(st0[None], device=cuda_device))
        assert np.array_equal(got, want[0].cpu().numpy())
        assert (kind, step) == (0, -1) and fail[0].tolist() == [0, -1]
        if n >= 31:
            np.testing.assert_allclose(got, O.rollout(steer, vel, 0.05, st0), rtol=0, atol=1e-9)
    steer = np.array([1.0, 2.0, 500.0, 3.0])
    got, kind, step = _lib.rollout_host_f64(steer, np.ones(4), 0.1, [0, 0, 0, 1.0], 460.0, 10.0)
    assert (kind, step) == (_lib.FAIL_STEER, 2)
    got, kind, step = _lib.rollout_host_f64(np.zeros(3), [1.0, 1.2, 9.0], 0.1, [0, 0, 0, 1.0], 460.0, 10.0)
    assert (kind, step) == (_lib.FAIL_ACCEL, 2)
    assert _lib.rollout_host_f64(np.zeros(0), np.zeros(0), 0.1, [0, 0, 0, 1.0], 460.0, 10.0)[1:] == (0, -1)


# ---- a7 / a8 / a9 / a13 ----------------------------------------------------------------------
def test_sub_trajectory_from_time_kat(cuda_device):
    tr = Trajectory(x=[0, 1, 2, 3, 4], y=[0, 0, 1, 1, 2], theta=[.5, .5, .6, .7, .8], velocity=[1] * 5,
                    time=[0, .1, .2, .3, .4])
    sub = tr.sub_trajectory_from_time(0.1, 0.3)
    b5 = G["schema"]["B5"]
    np.testing.assert_allclose(sub.x, unhex(b5["x"]), rtol=0, atol=1e-14)
    np.testing.assert_allclose(sub.y, unhex(b5["y"]), rtol=0, atol=1e-14)
    np.testing.assert_array_equal(sub.theta, unhex(b5["theta"]))
    np.testing.assert_array_equal(sub.time, unhex(b5["time"]))
    assert isinstance(sub.x, list) and isinstance(sub.x[0], float)
    with pytest.raises(AssertionError, match="No frames found"):
        tr.sub_trajectory_from_time(5.0, 6.0)


def test_time_extent_matches_searchsorted(cuda_device):
    rng = np.random.default_rng(2)
    t = np.cumsum(rng.choice([0.0, 0.05, 0.05, 0.1], 500))      # repeated stamps included
    for _ in range(40):
        a = rng.choice(t) if rng.random() < 0.5 else rng.uniform(t[0] - 1, t[-1] + 1)
        b = a + rng.uniform(0, 4)
        got = _lib.time_extent_f64(t, a, b)
        assert got == (int(np.searchsorted(t, a, "left")), int(np.searchsorted(t, b, "right")))


def test_traverse_kats(cuda_device):
    b6 = G["traverse"]["B6"]
    xy = unhex(b6["xy"], (-1, 2))
    np.testing.assert_array_equal(traverse_trajectory(xy, b6["D"]), unhex(b6["out"], (-1, 2)))
    for case in G["traverse"]["random"]:
        xy = unhex(case["xy"], (-1, 2))
        keep = _lib.traverse_f64(xy, float.fromhex(case["D"]))
        np.testing.assert_array_equal(keep, case["keep"])


# ---- a10 ----------------------------------------------------------------------------------------
def test_sequence_cost_kats(cuda_device):
    for case in G["cost"]:
        u = unhex(case["u"])
        tgt = unhex(case["target"], (-1, 2))
        got = sequence_cost(u, float.fromhex(case["v"]), case["dt"], tgt, float.fromhex(case["K"]))
        np.testing.assert_allclose(got[0], float.fromhex(case["cost"]), rtol=1e-12)


# ---- a12 ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["frames", "time", "traverse_gps"])
def test_write_back_matches_oracle_driver(cuda_device, mode):
    n = 210
    batch = synthetic_drives(2, n, seed=21)
    if mode == "frames":
        cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=25)
    elif mode == "time":
        cfg = SearchConfig(grid_v=8, grid_s=8, window_mode="time", horizon_time=1.2, horizon_frames=23)
    else:
        cfg = SearchConfig(grid_v=6, grid_s=10, window_mode="time", horizon_time=3.0, horizon_frames=59,
                           target_mode="traverse", primary="gps", w_vo=0.0, w_gps=1.0)
    drives = DriveSet.from_arrays(list(batch.time), [batch.dt] * 2, vo=list(batch.vo), gps=list(batch.gps))
    so, traj, plan = optimize_drives(cfg, drives)
    traj = traj.cpu().numpy()
    rec = so.records()
    spec = spec_of(cfg)
    for d in range(2):
        ref = O.optimize_drive(spec, batch.time[d], batch.dt, batch.vo[d], batch.gps[d])
        lo, hi = plan.window_offsets[d], plan.window_offsets[d + 1]
        np.testing.assert_array_equal(rec["best_idx"][lo:hi], [r.best_idx for r in ref.windows])
        sl = slice(d * n, (d + 1) * n)
        np.testing.assert_allclose(traj[0, sl], ref.x, rtol=0, atol=1e-9)
        np.testing.assert_allclose(traj[1, sl], ref.y, rtol=0, atol=1e-9)
        np.testing.assert_array_equal(traj[2, sl], ref.theta)      # fmod / add / halve: bit-exact
        np.testing.assert_array_equal(traj[3, sl], ref.velocity)


@pytest.mark.parametrize("key", ["driver", "driver_f64"])
def test_write_back_golden_reference_loop(cuda_device, key):
    """The frozen output of the REFERENCE driver loop (optimize_trajectory_v2.py:24-148 with the
    grid solver patched in for mpc_run), reproduced by plan -> search -> write-back; with
    float32 streams, and with float64 streams on inputs float32 cannot represent."""
    g = G[key]
    batch = synthetic_drives(1, g["n"], seed=g["seed"])
    vo, gps, sd = batch.vo[0], batch.gps[0], np.float32
    if key == "driver_f64":
        vo, gps, sd = off_float32_grid(vo), off_float32_grid(gps), np.float64
    cfg = SearchConfig(grid_v=g["grid"][0], grid_s=g["grid"][1], window_mode="time", horizon_time=3.0,
                       horizon_frames=g["horizon"], target_mode="traverse", primary="gps", w_vo=0.0,
                       w_gps=1.0)
    drives = DriveSet.from_arrays([batch.time[0]], [float.fromhex(g["dt"])], vo=[vo], gps=[gps],
                                  stream_dtype=sd)
    assert drives.f64 == (key == "driver_f64")
    so, traj, plan = optimize_drives(cfg, drives)
    traj = traj.cpu().numpy()
    np.testing.assert_array_equal(so.records()["best_idx"], g["best_idx"])
    np.testing.assert_allclose(traj[0], unhex(g["x"]), rtol=0, atol=1e-9)
    np.testing.assert_allclose(traj[1], unhex(g["y"]), rtol=0, atol=1e-9)
    np.testing.assert_array_equal(traj[2], unhex(g["theta"]))
    np.testing.assert_array_equal(traj[3], unhex(g["velocity"]))


@pytest.mark.parametrize("split", [False, True])
def test_drive_pipeline_graph_replay(cuda_device, split):
    """DrivePipeline (CUDA-graph replay of plan + search + write-back) equals the eager calls and
    follows new data copied into the resident streams."""
    from vehiclemodelvisualodometry_b200 import DrivePipeline

    cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=15)
    a, b = synthetic_drives(2, 120, seed=61), synthetic_drives(2, 120, seed=62)
    drives = DriveSet.from_arrays(list(a.time), [a.dt] * 2, vo=list(a.vo), gps=list(a.gps))
    pipe = DrivePipeline(cfg, drives, split=split)
    for batch in (a, b, a):
        drives.vo.copy_(torch.from_numpy(batch.vo.reshape(-1, 4)))
        drives.gps.copy_(torch.from_numpy(batch.gps.reshape(-1, 4)))
        if split:
            pipe.run_search()
            pipe.run_write_back()
        else:
            pipe.run()
        got_rec, got_traj = pipe.records.clone(), pipe.trajectory.clone()
        so, traj, _ = optimize_drives(cfg, drives)
        assert torch.equal(got_rec.view(torch.int32)[:, :3], so.results.view(torch.int32)[:, :3])
        assert torch.equal(got_rec[:, 16:], so.results[:, 16:])     # costs, seeds, first pose
        assert torch.equal(got_traj, traj)


def test_tabulated_tangent_is_within_one_ulp(cuda_device):
    """tan_steer (the TL table's tangent): every float of [2^-14, 0.62] -- 1.1e8 values -- against
    float64 tan, <= 1 ulp (the error band's eps_TL assumes 2u = 1 ulp); tanf's range beyond."""
    import ctypes as C

    from vehiclemodelvisualodometry_b200 import _lib

    ctx = _lib.context(cuda_device.index)
    lo = int(np.float32(2.0 ** -14).view(np.uint32))
    hi = int(np.float32(0.62).view(np.uint32))
    worst = 0.0
    step = 1 << 24
    for a in range(lo, hi + 1, step):
        b = min(a + step, hi + 1)
        bits = torch.arange(a, b, dtype=torch.int64, device=cuda_device).to(torch.int32)
        x = bits.view(torch.float32)
        out = torch.empty_like(x)
        ctx.check(ctx.lib.vmvo_tan_steer_f32(ctx.handle, x.numel(), _lib.ptr(x), _lib.ptr(out),
                                             _lib.stream_ptr(cuda_device)), "vmvo_tan_steer_f32")
        ref = torch.tan(x.double())
        ulp = torch.abs(ref.float()) * 2.0 ** -23          # >= the spacing at ref / 2
        err = (torch.abs(out.double() - ref) / ulp.double()).max().item()
        worst = max(worst, err)
    assert worst <= 1.0, worst
    # odd symmetry, zero, and the library range above the polynomial's
    x = torch.tensor([0.0, -0.3, 0.3, 0.7, -1.2, 1e-30], dtype=torch.float32, device=cuda_device)
    out = torch.empty_like(x)
    ctx.check(ctx.lib.vmvo_tan_steer_f32(ctx.handle, x.numel(), _lib.ptr(x), _lib.ptr(out),
                                         _lib.stream_ptr(cuda_device)), "vmvo_tan_steer_f32")
    o = out.cpu().numpy()
    assert o[0] == 0.0 and o[1] == -o[2] and o[5] == np.float32(1e-30)
    np.testing.assert_allclose(o, np.tan(x.cpu().numpy().astype(np.float64)), rtol=3e-7)


def test_drive_stream_overlaps_copies_and_keeps_order(cuda_device):
    """DriveStream: batches streamed from pinned host memory (H2D of the next, compute of the
    current, D2H of the previous on three streams) give, in order, what optimize_drives gives."""
    from vehiclemodelvisualodometry_b200 import DriveStream

    cfg = SearchConfig(grid_v=8, grid_s=8, window_frames=15)
    batches = [synthetic_drives(2, 110, seed=80 + i) for i in range(4)]
    a = batches[0]
    template = DriveSet.from_arrays(list(a.time), [a.dt] * 2, vo=list(a.vo), gps=list(a.gps))
    stream = DriveStream(cfg, template)

    def pinned(b):
        return {"vo": torch.from_numpy(np.ascontiguousarray(b.vo.reshape(-1, 4))).pin_memory(),
                "gps": torch.from_numpy(np.ascontiguousarray(b.gps.reshape(-1, 4))).pin_memory(),
                "time": torch.from_numpy(np.ascontiguousarray(b.time.reshape(-1))).pin_memory()}

    got = list(stream.run(pinned(b) for b in batches))
    assert len(got) == len(batches)
    for b, (rec, traj) in zip(batches, got):
        d = DriveSet.from_arrays(list(b.time), [b.dt] * 2, vo=list(b.vo), gps=list(b.gps))
        so, want, _ = optimize_drives(cfg, d)
        r = so.records()
        for f in ("best_idx", "n_steps", "status", "best_cost", "x1", "y1", "theta1"):
            np.testing.assert_array_equal(rec[f], r[f])
        np.testing.assert_array_equal(traj, want.cpu().numpy())
    assert list(DriveStream(cfg, template).run([])) == []
