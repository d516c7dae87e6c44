"""Kinematic bicycle model, interface-compatible with vmvo/bicycle_model.py:16-100.

``run`` / ``run_sequence`` keep the reference's signatures, state mutation and assertion
messages; the integration itself is the batched CUDA rollout (``vmvo_rollout_f64`` /
``vmvo_rollout_f32``).  ``rollout_batch`` is the additive tensor API.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .constants import MAX_ACCEL, MAX_STEER, MAX_STEER_RATE, STEERING_RATIO, WHEEL_BASE
from .schema import State

STARTING_STATE = State(x=0.0, y=0.0, theta=0.0, velocity=0.0, steering_angle=0.0)

_FAIL_MESSAGES = {
    _lib.FAIL_STEER: "Steering angle is out of bounds",      # bicycle_model.py:48-50
    _lib.FAIL_ACCEL: "Acceleration is out of bounds",        # bicycle_model.py:59-62
}


def rollout_batch(steer: torch.Tensor, vel: torch.Tensor, dt: float, state0: torch.Tensor,
                  max_steer: float = MAX_STEER, max_accel: float = MAX_ACCEL
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Roll ``n_seq`` control sequences forward on the GPU.

    steer, vel: [n_seq, n_steps] (degrees at the steering wheel, m/s), float64 or float32 CUDA
    tensors; state0: [n_seq, 4] = (x, y, theta, velocity).  Returns (poses [n_seq, n_steps, 3],
    fail [n_seq, 2] int32 = (VMVO_FAIL_* kind, first offending step or -1)).
    """
    if steer.dtype not in (torch.float64, torch.float32):
        raise TypeError("steer must be float32 or float64")
    dtype = steer.dtype
    if not steer.is_cuda:
        raise RuntimeError("rollout_batch needs CUDA tensors (there is no CPU fallback)")
    steer = steer.contiguous()
    vel = vel.to(dtype).contiguous()
    state0 = state0.to(dtype).contiguous()
    n_seq, n_steps = steer.shape
    if vel.shape != steer.shape or state0.shape != (n_seq, 4):
        raise ValueError("shape mismatch: steer/vel [n_seq, n_steps], state0 [n_seq, 4]")
    ctx = _lib.context(steer.device.index)
    out = torch.empty((n_seq, n_steps, 3), dtype=dtype, device=steer.device)
    fail = torch.empty((n_seq, 2), dtype=torch.int32, device=steer.device)
    fn = ctx.lib.vmvo_rollout_f64 if dtype == torch.float64 else ctx.lib.vmvo_rollout_f32
    ctx.check(fn(ctx.handle, n_seq, n_steps, _lib.ptr(steer), _lib.ptr(vel), float(dt),
                 _lib.ptr(state0), float(max_steer), float(max_accel), _lib.ptr(out), _lib.ptr(fail),
                 _lib.stream_ptr(steer.device)), "vmvo_rollout")
    return out, fail


class BicycleModel:
    """Bicycle model for vehicle motion control (reference signature and behaviour)."""

    def __init__(
        self,
        state: State = STARTING_STATE,
        wheel_base: float = WHEEL_BASE,
        steering_ratio: float = STEERING_RATIO,
        max_steer: float = MAX_STEER,
        max_steer_rate: float = MAX_STEER_RATE,
        max_accel: float = MAX_ACCEL,
    ) -> None:
        self.state = state
        # stored but, as in the reference (quirk D2, bicycle_model.py:66-68), the dynamics
        # use the module constants
        self.wheel_base = wheel_base
        self.steering_ratio = steering_ratio
        self.max_steer = max_steer
        self.max_accel = max_accel
        self.max_steer_rate = max_steer_rate

    def run(self, steering_angle: float, velocity: float, dt: float) -> State:
        """One timestep; mutates and returns the state."""
        return self.run_sequence([steering_angle], [velocity], dt)[0]

    def run_sequence(self, steering_angles, velocities, dt: float) -> List[State]:
        """States after each step (the initial state is not included)."""
        assert len(steering_angles) == len(velocities)
        n = len(steering_angles)
        if n == 0:
            return []
        # the scalar caller's path (vmvo_rollout_host_f64): host buffers in and out, one launch and one
        # stream synchronisation per call; raises when there is no GPU (no CPU fallback)
        st = self.state
        poses, kind, step = _lib.rollout_host_f64(steering_angles, velocities, float(dt),
                                                  (st.x, st.y, st.theta, st.velocity),
                                                  self.max_steer, self.max_accel)
        good = n if kind == 0 else step
        states: List[State] = []
        sa = np.asarray(steering_angles, dtype=np.float64)
        va = np.asarray(velocities, dtype=np.float64)
        for k in range(good):
            states.append(State(x=poses[k, 0], y=poses[k, 1], theta=poses[k, 2], velocity=va[k],
                                steering_angle=sa[k]))
        if states:
            self.set_state(states[-1])      # steps before a violation have been applied
        if kind != 0:
            raise AssertionError(_FAIL_MESSAGES[kind])
        return states

    def set_state(self, state: State) -> None:
        self.state = state

    def reset(self) -> None:
        self.state = STARTING_STATE
