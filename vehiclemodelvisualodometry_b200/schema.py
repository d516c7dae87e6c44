"""State / Trajectory types of the path, interface-compatible with vmvo/schema.py:10-147.

Same class names, field names, units and method signatures as the reference so that
``vmvo.scripts`` callers drop in unchanged.  ``sub_trajectory`` (window slice + local-frame
transform, schema.py:59-115) and ``sub_trajectory_from_time`` (schema.py:117-127) run on
the GPU through the C ABI (``vmvo_extract_window_f64``); there is no CPU fallback.
"""
from __future__ import annotations

from typing import List

import numpy as np
from pydantic import BaseModel


class State(BaseModel):
    """Vehicle state: x, y [m], theta [rad], velocity [m/s], steering_angle [deg, wheel]."""

    x: float
    y: float
    theta: float
    velocity: float
    steering_angle: float


class Trajectory(BaseModel):
    """Five equally indexed columns; ``len`` is the length of ``x`` (schema.py:30-31)."""

    x: List[float]
    y: List[float]
    theta: List[float]
    velocity: List[float]
    time: List[float]

    def __len__(self):
        return len(self.x)

    def __getitem__(self, key):
        return (self.x[key], self.y[key], self.theta[key], self.velocity[key], self.time[key])

    def __repr__(self):
        return f"Trajectory(len={len(self)})"

    __str__ = __repr__

    def to_numpy(self):
        """[len, 5] float64, columns x, y, theta, velocity, time (schema.py:47-57)."""
        return np.array([self.x, self.y, self.theta, self.velocity, self.time]).T

    def sub_trajectory(self, start: int, end: int, theta_window: int = 10):
        """Poses [start, end) in the frame of the first one: origin there, heading 0."""
        from . import _lib

        start, end = int(start), int(end)
        x = np.asarray(self.x[start:end], dtype=np.float64)
        y = np.asarray(self.y[start:end], dtype=np.float64)
        th = np.asarray(self.theta[start:end], dtype=np.float64)
        lx, ly, lth = _lib.extract_window_f64(x, y, th)
        return Trajectory(x=lx, y=ly, theta=lth, velocity=self.velocity[start:end],
                          time=self.time[start:end])

    def sub_trajectory_from_time(self, start_time: float, end_time: float):
        """Poses with time in [start_time, end_time] (left / right bisection)."""
        from . import _lib

        start_index, end_index = _lib.time_extent_f64(
            np.asarray(self.time, dtype=np.float64), float(start_time), float(end_time))
        assert end_index > start_index, "No frames found"
        return self.sub_trajectory(start=start_index, end=end_index)


def states_list_to_trajectory(states: List[State], start_time: float, dt: float) -> Trajectory:
    """``time[k] = start_time + k*dt`` from k = 0 for the state after step 1 (schema.py:130-147)."""
    return Trajectory(
        x=[s.x for s in states],
        y=[s.y for s in states],
        theta=[s.theta for s in states],
        velocity=[s.velocity for s in states],
        time=[start_time + i * dt for i in range(len(states))],
    )
