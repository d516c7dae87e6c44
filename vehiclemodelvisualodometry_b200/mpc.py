"""Window solver, signature-compatible with vmvo/utils/mpc.py:14-141.

``grid_run`` takes ``mpc_run``'s arguments and returns the same thing -- the steering
sequence [deg] of the selected control -- but selects it by the exhaustive hypothesis-grid
argmin of the fused CUDA search instead of SciPy SLSQP.  ``mpc_run`` is an alias.
"""
from __future__ import annotations

from dataclasses import replace
from typing import Optional

import numpy as np
import torch

from . import _lib
from .bicycle_model import BicycleModel
from .schema import Trajectory
from .search import DriveSet, SearchConfig, WindowPlan, grid_search

# reference-faithful single-window configuration: the window handed in IS the target
REFERENCE_WINDOW_CFG = SearchConfig(target_mode="traverse", seed_mode="given")


def traverse_trajectory(traj: np.ndarray, D: float) -> np.ndarray:
    """Arc-length decimation of a polyline (vmvo/utils/mpc.py:125-141), on the GPU."""
    traj = np.asarray(traj, dtype=np.float64)
    keep = _lib.traverse_f64(traj[:, :2], D)
    return traj[keep]


def sequence_cost(u: np.ndarray, velocity: float, dt: float, target_xy: np.ndarray,
                  K: float = 0.0) -> np.ndarray:
    """The cost closure of mpc_run (mpc.py:56-85) for a batch of steering sequences [B, N]."""
    u = np.atleast_2d(np.asarray(u, dtype=np.float64))
    target_xy = np.ascontiguousarray(target_xy, dtype=np.float64)
    n_seq, n_steps = u.shape
    if target_xy.shape[0] < n_steps + 1:
        raise ValueError("target needs n_steps + 1 points")
    ctx = _lib.context()
    du = torch.as_tensor(np.ascontiguousarray(u)).cuda()
    dt_xy = torch.as_tensor(target_xy).cuda()
    cost = torch.empty(n_seq, dtype=torch.float64, device=du.device)
    ctx.check(ctx.lib.vmvo_sequence_cost_f64(ctx.handle, n_seq, n_steps, _lib.ptr(du), float(velocity),
                                             float(dt), _lib.ptr(dt_xy), float(K), _lib.ptr(cost),
                                             _lib.stream_ptr()), "vmvo_sequence_cost_f64")
    return cost.cpu().numpy()


def grid_run(trajectory: Trajectory, bicycle_model: BicycleModel, velocity: float,
             starting_steering_angle: float, time_step: float,
             config: Optional[SearchConfig] = None, return_info: bool = False):
    """Steering sequence [N] of the best hypothesis for one local-frame window.

    ``trajectory`` is the window already in its local frame (what
    ``Trajectory.sub_trajectory_from_time`` returns); ``velocity`` seeds V_w and
    ``starting_steering_angle`` seeds S_w.  N = len(traverse_trajectory(xy, v*dt)) - 1 as in
    the reference; an empty decimation returns ``np.zeros(0)`` (mpc.py:42-43).
    """
    cfg = config if config is not None else REFERENCE_WINDOW_CFG
    n = len(trajectory)
    cfg = replace(cfg, window_mode="frames", window_frames=max(n - 1, 1), seed_mode="given",
                  primary="vo", w_vo=1.0 if cfg.w_vo == 0 and cfg.w_gps == 0 else cfg.w_vo,
                  max_steer=float(bicycle_model.max_steer), max_accel=float(bicycle_model.max_accel),
                  max_steer_rate=float(bicycle_model.max_steer_rate),
                  max_window_poses=max(n, 2))
    if cfg.w_gps or cfg.w_imu:
        raise ValueError("grid_run scores against the single window it is given")
    poses = np.stack([np.asarray(trajectory.x, dtype=np.float64),
                      np.asarray(trajectory.y, dtype=np.float64),
                      np.asarray(trajectory.theta, dtype=np.float64)[:n]
                      if len(trajectory.theta) >= n else np.zeros(n),
                      np.asarray(trajectory.velocity, dtype=np.float64)[:n]], axis=1)
    drives = DriveSet.from_arrays([np.asarray(trajectory.time, dtype=np.float64)[:n]], [time_step],
                                  vo=[poses], stream_dtype=np.float64)
    dev = drives.device
    plan = WindowPlan(window_offsets=[0, 1],
                      d_window_offsets=torch.tensor([0, 1], dtype=torch.int64, device=dev),
                      win_start=torch.zeros(1, dtype=torch.int64, device=dev),
                      win_len=torch.tensor([n], dtype=torch.int32, device=dev),
                      win_drive=torch.zeros(1, dtype=torch.int32, device=dev))
    seeds = torch.tensor([[float(velocity), float(starting_steering_angle)]], dtype=torch.float64)
    out = grid_search(cfg, drives, plan, seeds=seeds, want_rollouts=True)
    rec = out.records()[0]
    N = int(rec["n_steps"])
    steer = out.steer[0, :N].cpu().numpy()
    if return_info:
        return steer, rec, out
    return steer


mpc_run = grid_run
