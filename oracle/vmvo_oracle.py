"""CPU oracle for the VMVO window search -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  The product (``vehiclemodelvisualodometry_b200``) never does; it
fails loudly when the CUDA library is missing.

What is restated here, in float64 NumPy, and where it comes from in the reference
(paths relative to /root/reference):

  bicycle_step / rollout        vmvo/bicycle_model.py:40-92      (a1, a2)
  sub_trajectory[_from_time]    vmvo/schema.py:59-127            (a7, a8)
  traverse_trajectory           vmvo/utils/mpc.py:125-141        (a9)
  sequence_cost                 vmvo/utils/mpc.py:56-85          (a10)
  write-back + blends           vmvo/scripts/optimize_trajectory_v2.py:32-148 (a12)
  rollout timestamps            vmvo/schema.py:130-147           (a13)

The reference has no hypothesis grid (it calls SciPy SLSQP, vmvo/utils/mpc.py:112-119);
the grid, seeds and multi-sensor cost follow the derived spec in DESIGN.md section 2
(SURVEY.md Appendix C).  That part has no reference counterpart and is pinned only
through the reference primitives it is built from (see oracle/make_golden.py).

Parity pinning: every function above is checked against the imported, unmodified
reference in oracle/make_golden.py (run in the build container) and against the
frozen vectors in tests/golden/.  The IMU cost term is spec-only: parity unpinned.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# vmvo/constants.py:3-7
WHEEL_BASE = 2.83972
STEERING_RATIO = 13.27
MAX_STEER = 460.0
MAX_ACCEL = 10
MAX_STEER_RATE = 100.0

TWO_PI = 2 * np.pi

# status bits of a window result (mirrors include/vmvo_b200.h)
WIN_OK = 0
WIN_EMPTY = 1        # fewer than two targets: N == 0 (reference: np.zeros(0), mpc.py:42-43)
WIN_NONFINITE = 2    # NaN/Inf in the window's inputs; argmin degenerates to index 0


@dataclass(frozen=True)
class SearchSpec:
    """The derived search spec."""

    grid_v: int = 32
    grid_s: int = 32
    window_mode: str = "frames"      # "frames": W+1 poses per window; "time": reference rule
    window_frames: int = 30          # W (steps) in "frames" mode
    horizon_time: float = 3.0        # seconds, "time" mode (optimize_trajectory_v2.py:35)
    horizon_frames: int = 60         # int(horizon_time * FPS), "time" mode (…v2.py:42)
    target_mode: str = "time"        # "time" | "traverse"
    target_offset: int = 1           # 1 = reference (state k vs target k-1, mpc.py:70-78)
    seed_mode: str = "data"          # "data" | "given" | "chained"
    primary: str = "vo"              # which stream defines window frame, seeds and dt
    w_vo: float = 1.0
    w_gps: float = 0.0
    w_imu: float = 0.0
    k_steer: float = 0.0             # K of mpc.py:31
    wheel_base: float = WHEEL_BASE
    steering_ratio: float = STEERING_RATIO
    max_steer: float = MAX_STEER
    max_accel: float = float(MAX_ACCEL)
    max_steer_rate: float = MAX_STEER_RATE

    def horizon(self) -> int:
        return self.window_frames if self.window_mode == "frames" else self.horizon_frames


# --------------------------------------------------------------------------------------
# a1 / a2: the kinematic bicycle model
# --------------------------------------------------------------------------------------

def bicycle_step(x, y, theta, steer_deg, v, dt, wheel_base=WHEEL_BASE, ratio=STEERING_RATIO):
    """One step of vmvo/bicycle_model.py:66-75 (scalars or arrays, float64).

    Left-to-right evaluation as in the reference: ((v / L) * tan d) * dt, (v * cos th') * dt;
    x and y use the UPDATED heading (quirk D1).
    """
    delta = np.radians(steer_deg) / ratio
    theta_n = theta + (v / wheel_base * np.tan(delta) * dt)
    x_n = x + (v * np.cos(theta_n) * dt)
    y_n = y + (v * np.sin(theta_n) * dt)
    return x_n, y_n, theta_n


def check_feasible(steer_deg, v, v_prev, dt, max_steer=MAX_STEER, max_accel=MAX_ACCEL):
    """The two asserts of vmvo/bicycle_model.py:48-62, same messages."""
    assert abs(steer_deg) <= max_steer, "Steering angle is out of bounds"
    estimated_accel = (v - v_prev) / dt
    assert abs(estimated_accel) <= max_accel, "Acceleration is out of bounds"


def rollout(steers, vels, dt, state0=(0.0, 0.0, 0.0, 0.0), check=True,
            max_steer=MAX_STEER, max_accel=MAX_ACCEL):
    """BicycleModel.run_sequence (bicycle_model.py:80-92): states AFTER each step.

    Returns float64 array [N, 3] = (x, y, theta).  state0 = (x, y, theta, velocity).
    """
    steers = np.asarray(steers, dtype=np.float64)
    vels = np.asarray(vels, dtype=np.float64)
    assert len(steers) == len(vels)
    x, y, th, v_prev = (np.float64(s) for s in state0)
    out = np.empty((len(steers), 3), dtype=np.float64)
    for k in range(len(steers)):
        if check:
            check_feasible(steers[k], vels[k], v_prev, dt, max_steer, max_accel)
        x, y, th = bicycle_step(x, y, th, steers[k], vels[k], dt)
        v_prev = vels[k]
        out[k] = (x, y, th)
    return out


# --------------------------------------------------------------------------------------
# a7 / a8: window extraction and the local-frame transform
# --------------------------------------------------------------------------------------

def local_frame(x, y, theta):
    """vmvo/schema.py:64-107: translate to the first point, rotate by -theta_0.

    The reference evaluates ``(p - p0) @ [[c, -s], [s, c]]`` with np.dot, i.e.
    x' = dx*c + dy*s, y' = -dx*s + dy*c; written out here so the order of the two
    products is fixed (np.dot may fuse them -- agreement is to 1 ulp, see make_golden).
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    th0 = theta[0]
    c, s = np.cos(th0), np.sin(th0)
    dx = x - x[0]
    dy = y - y[0]
    return dx * c + dy * s, -dx * s + dy * c, theta - th0


def window_extent_time(time, t0, t1):
    """vmvo/schema.py:119-122: [searchsorted_left(t0), searchsorted_right(t1))."""
    start = int(np.searchsorted(time, t0, side="left"))
    end = int(np.searchsorted(time, t1, side="right"))
    assert end > start, "No frames found"
    return start, end


def window_extents(spec: SearchSpec, time: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Start index and pose count of every window of one drive.

    Window count follows optimize_trajectory_v2.py:48 (``range(N - horizon * 2)``).
    "time" mode: poses with time in [t_i, t_i + horizon_time] (…v2.py:49-59);
    "frames" mode: poses i .. i+W.
    """
    n = len(time)
    h = spec.horizon()
    nw = max(0, n - 2 * h)
    starts = np.arange(nw, dtype=np.int64)
    if spec.window_mode == "frames":
        lens = np.minimum(spec.window_frames + 1, n - starts).astype(np.int32)
    else:
        t = np.asarray(time, dtype=np.float64)
        lens = np.empty(nw, dtype=np.int32)
        for i in range(nw):
            # the reference searches the whole array: with repeated stamps s can be < i
            s, e = window_extent_time(t, t[i], t[i] + spec.horizon_time)
            starts[i] = s
            lens[i] = e - s
    return starts, lens


# --------------------------------------------------------------------------------------
# a9: arc-length decimation of the target
# --------------------------------------------------------------------------------------

def traverse_indices(xy: np.ndarray, D: float) -> np.ndarray:
    """Indices kept by vmvo/utils/mpc.py:125-141 (the reference returns xy[indices])."""
    keep = [0]
    dist = 0.0
    for i in range(1, xy.shape[0]):
        seg = ((xy[i, 0] - xy[i - 1, 0]) ** 2 + (xy[i, 1] - xy[i - 1, 1]) ** 2) ** 0.5
        if dist + seg > D:
            keep.append(i - 1)
            dist = seg
        else:
            dist += seg
    return np.asarray(keep, dtype=np.int64)


def traverse_trajectory(xy: np.ndarray, D: float) -> np.ndarray:
    xy = np.asarray(xy, dtype=np.float64)
    return xy[traverse_indices(xy, D)]


# --------------------------------------------------------------------------------------
# a10: cost of one control sequence (the closure inside mpc_run)
# --------------------------------------------------------------------------------------

def sequence_cost(u, v, dt, target_xy, K=0.0):
    """vmvo/utils/mpc.py:68-85 for one steering sequence at constant speed ``v``.

    Start state (target[0], theta=0); state after step i+1 is compared with target i
    (quirk D3); N = len(target) - 1.
    """
    target_xy = np.asarray(target_xy, dtype=np.float64)
    N = len(target_xy) - 1
    x, y, th = np.float64(target_xy[0, 0]), np.float64(target_xy[0, 1]), np.float64(0.0)
    cost = 0.0
    for i in range(N):
        x, y, th = bicycle_step(x, y, th, u[i], v, dt)
        cost += (x - target_xy[i, 0]) ** 2 + (y - target_xy[i, 1]) ** 2 + K * u[i] ** 2
    return cost


# --------------------------------------------------------------------------------------
# derived spec: hypothesis grid, seeds, window targets, argmin
# --------------------------------------------------------------------------------------

def grid_axis(limit: float, g: int) -> np.ndarray:
    """Rates spanning [-limit, +limit]: limit * (2i - (g-1)) / (g-1); exactly antisymmetric."""
    if g == 1:
        return np.zeros(1)
    num = (2 * np.arange(g) - (g - 1)).astype(np.float64)
    return (limit * num) / np.float64(g - 1)


def hypothesis_controls(spec: SearchSpec, v_seed: float, s_seed: float, n_steps: int, dt: float):
    """V[k, i] and S[k, j] for k = 1..N (row k-1): constant-rate profiles, clamped."""
    a = grid_axis(spec.max_accel, spec.grid_v)
    r = grid_axis(spec.max_steer_rate, spec.grid_s)
    t = np.arange(1, n_steps + 1, dtype=np.float64) * dt
    V = np.maximum(0.0, v_seed + a[None, :] * t[:, None])
    S = np.minimum(spec.max_steer, np.maximum(-spec.max_steer, s_seed + r[None, :] * t[:, None]))
    return V, S


def _ieee_remainder(d):
    """IEEE remainder of d by 2*pi (exact operation): result in [-pi, pi]."""
    d = np.asarray(d, dtype=np.float64)
    out = np.vectorize(lambda z: math.remainder(z, TWO_PI) if math.isfinite(z) else float("nan"),
                       otypes=[np.float64])(d)
    return out if out.ndim else np.float64(out)


def seed_from_window(spec: SearchSpec, th_local: np.ndarray, vel: np.ndarray, dt: float):
    """Data seeds: V_w from optimize_trajectory_v2.py:61-63; S_w by the
    inverse model on the heading change over the window's first frame interval."""
    v_seed = (np.float64(vel[0]) + np.float64(vel[-1])) / 2
    s_seed = np.float64(0.0)
    if len(th_local) >= 2 and v_seed * dt > 1e-6:
        dth = _ieee_remainder(th_local[1] - th_local[0])
        ang = np.arctan(spec.wheel_base * dth / (v_seed * dt))
        s_seed = (ang * (180.0 / np.pi)) * spec.steering_ratio
        s_seed = np.minimum(spec.max_steer, np.maximum(-spec.max_steer, s_seed))
    return np.float64(v_seed), np.float64(s_seed)


@dataclass
class WindowTargets:
    n_steps: int
    status: int
    v_seed: float
    s_seed: float
    vo_xy: Optional[np.ndarray] = None     # [n_targets, 2] local frame
    gps_xy: Optional[np.ndarray] = None
    imu_th: Optional[np.ndarray] = None    # [n_targets] yaw relative to the window start
    keep: Optional[np.ndarray] = None      # source index of every target


def build_window(spec: SearchSpec, start: int, length: int, dt: float,
                 vo: Optional[np.ndarray], gps: Optional[np.ndarray], imu: Optional[np.ndarray],
                 seeds: Optional[Tuple[float, float]] = None) -> WindowTargets:
    """Targets, step count and seeds of one window.

    ``vo`` / ``gps`` are the drive's pose streams [n, 4] = (x, y, theta, v) (float32 values
    are widened to float64 here, so both sides see identical inputs); ``imu`` is yaw [n].
    """
    sl = slice(start, start + length)
    prim = vo if spec.primary == "vo" else gps
    assert prim is not None
    p = np.asarray(prim[sl], dtype=np.float64)
    px, py, pth = local_frame(p[:, 0], p[:, 1], p[:, 2])
    if seeds is None:
        v_seed, s_seed = seed_from_window(spec, pth, p[:, 3], dt)
    else:
        v_seed, s_seed = np.float64(seeds[0]), np.float64(seeds[1])
    if spec.target_mode == "traverse":
        keep = traverse_indices(np.stack([px, py], axis=1), v_seed * dt)
    else:
        keep = np.arange(length, dtype=np.int64)
    wt = WindowTargets(n_steps=len(keep) - 1, status=WIN_OK, v_seed=float(v_seed),
                       s_seed=float(s_seed), keep=keep)
    finite = np.isfinite(v_seed) and np.isfinite(s_seed)
    for name, stream, w in (("vo", vo, spec.w_vo), ("gps", gps, spec.w_gps)):
        if w == 0.0:
            continue
        assert stream is not None, f"w_{name} != 0 but no {name} stream"
        if stream is prim:
            lx, ly = px, py
        else:
            q = np.asarray(stream[sl], dtype=np.float64)
            lx, ly, _ = local_frame(q[:, 0], q[:, 1], q[:, 2])
        xy = np.stack([lx[keep], ly[keep]], axis=1)
        finite = finite and bool(np.all(np.isfinite(xy)))
        setattr(wt, f"{name}_xy", xy)
    if spec.w_imu != 0.0:
        assert imu is not None
        yaw = np.asarray(imu[sl], dtype=np.float64)
        wt.imu_th = (yaw - yaw[0])[keep]
        finite = finite and bool(np.all(np.isfinite(wt.imu_th)))
    if wt.n_steps <= 0:
        wt.n_steps = 0
        wt.status |= WIN_EMPTY
    if not finite:
        wt.status |= WIN_NONFINITE
    return wt


def grid_costs(spec: SearchSpec, wt: WindowTargets, dt: float) -> np.ndarray:
    """Cost of every hypothesis, float64 [G_v, G_s], by rolling the model forward.

    Per step the term is  w_vo*|p - T_vo|^2 + w_gps*|p - T_gps|^2 + w_imu*wrap(th - th_imu)^2
    + K*S^2 (zero-weight terms skipped), accumulated in step order like mpc.py:70-78 (the CUDA
    kernels associate the same sums differently -- warp scans, butterflies -- and agree to ~1e-15).
    """
    N = wt.n_steps
    V, S = hypothesis_controls(spec, wt.v_seed, wt.s_seed, N, dt)
    tan_d = np.tan(np.radians(S) / spec.steering_ratio)        # [N, G_s]
    shape = (spec.grid_v, spec.grid_s)
    th = np.zeros(shape)
    x = np.zeros(shape)
    y = np.zeros(shape)
    cost = np.zeros(shape)
    off = spec.target_offset
    with np.errstate(invalid="ignore", over="ignore"):
        for k in range(1, N + 1):
            v = V[k - 1][:, None]
            th = th + (v / spec.wheel_base * tan_d[k - 1][None, :] * dt)
            x = x + (v * np.cos(th) * dt)
            y = y + (v * np.sin(th) * dt)
            term = None
            for w, T in ((spec.w_vo, wt.vo_xy), (spec.w_gps, wt.gps_xy)):
                if w == 0.0:
                    continue
                e = (x - T[k - off, 0]) ** 2 + (y - T[k - off, 1]) ** 2
                e = e if w == 1.0 else w * e
                term = e if term is None else term + e
            if spec.w_imu != 0.0:
                d = _ieee_remainder(th - wt.imu_th[k - off])
                e = spec.w_imu * d ** 2
                term = e if term is None else term + e
            if spec.k_steer != 0.0:
                e = spec.k_steer * S[k - 1][None, :] ** 2
                term = e if term is None else term + e
            if term is not None:
                cost = cost + term
    return cost


def argmin_first(cost: np.ndarray) -> int:
    """np.argmin semantics: lowest flat index among minimal costs (NaN wins, first NaN)."""
    return int(np.argmin(cost.reshape(-1)))


@dataclass
class WindowResult:
    best_idx: int
    best_cost: float
    n_steps: int
    status: int
    v_seed: float
    s_seed: float
    steer: np.ndarray        # [N] steering-wheel angle, degrees (what mpc_run returns)
    vel: np.ndarray          # [N]
    poses: np.ndarray        # [N, 3] rollout of the best hypothesis, local frame


def solve_window(spec: SearchSpec, wt: WindowTargets, dt: float, want_costs=False):
    N = wt.n_steps
    if (wt.status & WIN_EMPTY) or N == 0:
        res = WindowResult(-1, float("nan"), 0, wt.status, wt.v_seed, wt.s_seed,
                           np.zeros(0), np.zeros(0), np.zeros((0, 3)))
        return (res, None) if want_costs else res
    if wt.status & WIN_NONFINITE:
        # every hypothesis' cost is NaN or Inf alike; np.argmin then returns index 0
        best = 0
        cost = None
        best_cost = float("nan")
    else:
        cost = grid_costs(spec, wt, dt)
        best = argmin_first(cost)
        best_cost = float(cost.reshape(-1)[best])
    i, j = divmod(best, spec.grid_s)
    V, S = hypothesis_controls(spec, wt.v_seed, wt.s_seed, N, dt)
    steer, vel = S[:, j].copy(), V[:, i].copy()
    if wt.status & WIN_NONFINITE:
        poses = np.full((N, 3), np.nan)
    else:
        poses = rollout(steer, vel, dt, (0.0, 0.0, 0.0, wt.v_seed), check=False)
    res = WindowResult(best, best_cost, N, wt.status, wt.v_seed, wt.s_seed, steer, vel, poses)
    return (res, cost) if want_costs else res


# --------------------------------------------------------------------------------------
# a12 / a13: the sliding-window driver with write-back and blends
# --------------------------------------------------------------------------------------

def blend_theta(vo_th, gps_th):
    """optimize_trajectory_v2.py:126-133 (wrapped-angle midpoint)."""
    d = (vo_th - gps_th) % TWO_PI
    if d > np.pi:
        d -= TWO_PI
    return (vo_th - d / 2) % TWO_PI


@dataclass
class DriveResult:
    windows: List[WindowResult]
    x: np.ndarray
    y: np.ndarray
    theta: np.ndarray
    velocity: np.ndarray
    time: np.ndarray


def optimize_drive(spec: SearchSpec, time: np.ndarray, dt: float,
                   vo: Optional[np.ndarray], gps: Optional[np.ndarray] = None,
                   imu: Optional[np.ndarray] = None,
                   seeds: Optional[np.ndarray] = None) -> DriveResult:
    """The loop of optimize_trajectory_v2.py:48-146 with the grid argmin in place of SLSQP.

    Output columns start as a copy of VO (…v2.py:33); window i overwrites x,y[i:i+size] with
    its LOCAL-frame rollout (quirk D4, …v2.py:122-123), later windows win; theta[i] and
    velocity[i] are the VO/GPS blends (…v2.py:126-137) when a GPS stream is present.
    ``seeds`` [n_windows, 2] = (V_w, S_w) for seed_mode "given".
    """
    time = np.asarray(time, dtype=np.float64)
    base = vo if vo is not None else gps
    base = np.asarray(base, dtype=np.float64)
    ox, oy, oth, ov = (base[:, c].copy() for c in range(4))
    starts, lens = window_extents(spec, time)
    results: List[WindowResult] = []
    s_chain = 0.0
    for w, (st, ln) in enumerate(zip(starts, lens)):
        i = w                                            # the reference's loop index
        given = None
        if spec.seed_mode == "given":
            given = (seeds[w, 0], seeds[w, 1])
        wt = build_window(spec, int(st), int(ln), dt, vo, gps, imu, given)
        if spec.seed_mode == "chained":
            wt.s_seed = float(s_chain)
        res = solve_window(spec, wt, dt)
        results.append(res)
        if res.n_steps > 0:
            size = res.n_steps
            hi = min(len(ox), i + size)
            # Python slice assignment on lists can grow the list; the reference never
            # hits that because i + size <= N - horizon.  Clip defensively.
            ox[i:hi] = res.poses[: hi - i, 0]
            oy[i:hi] = res.poses[: hi - i, 1]
            s_chain = res.steer[-1]                      # …v2.py:146
        if vo is not None and gps is not None:
            oth[i] = blend_theta(np.float64(vo[i, 2]), np.float64(gps[i, 2]))
            ov[i] = (np.float64(vo[i, 3]) + np.float64(gps[i, 3])) / 2
    return DriveResult(results, ox, oy, oth, ov, time.copy())


def rollout_times(start_time: float, dt: float, n: int) -> np.ndarray:
    """vmvo/schema.py:135: the state after step 1 is stamped start_time (quirk D5)."""
    return np.asarray([start_time + i * dt for i in range(n)], dtype=np.float64)


def reference_dt(time: np.ndarray) -> Tuple[float, int, float]:
    """optimize_trajectory_v2.py:35-42: (dt, horizon, FPS) for horizon_time = 3.0."""
    FPS = 1 / np.mean(np.diff(np.asarray(time, dtype=np.float64)))
    return float(1.0 / FPS), int(3.0 * FPS), float(FPS)
