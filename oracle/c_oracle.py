"""ctypes wrapper of the C restatement (oracle/vmvo_oracle.c) -- TEST INFRASTRUCTURE.

Used by tests (large-size parity) and by bench.py's CPU-baseline legs only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from vehiclemodelvisualodometry_b200._lib import RESULT_DTYPE, SearchCfg

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libvmvo_oracle.so")
_lib = None


def build() -> str:
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.vmvo_oracle_search.restype = C.c_int64
        lib.vmvo_oracle_search.argtypes = [C.POINTER(SearchCfg), C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
        lib.vmvo_oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def search(cfg: SearchCfg, win_start, win_len, win_drive, dt_per_drive, vo, gps=None, imu=None,
           seeds=None, n_threads: int = 0):
    """Returns (records, hypothesis_steps); arrays are host NumPy, streams float32 [F, 4]."""
    lib = load()
    win_start = np.ascontiguousarray(win_start, dtype=np.int64)
    win_len = np.ascontiguousarray(win_len, dtype=np.int32)
    win_drive = np.ascontiguousarray(win_drive, dtype=np.int32)
    dt_per_drive = np.ascontiguousarray(dt_per_drive, dtype=np.float64)
    vo = None if vo is None else np.ascontiguousarray(vo, dtype=np.float32)
    gps = None if gps is None else np.ascontiguousarray(gps, dtype=np.float32)
    imu = None if imu is None else np.ascontiguousarray(imu, dtype=np.float32)
    seeds = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.float64)
    out = np.zeros(len(win_start), dtype=RESULT_DTYPE)
    steps = lib.vmvo_oracle_search(C.byref(cfg), len(win_start), _p(win_start), _p(win_len), _p(win_drive),
                                   _p(dt_per_drive), _p(vo), _p(gps), _p(imu), _p(seeds), _p(out),
                                   int(n_threads))
    return out, int(steps)


def max_threads() -> int:
    return int(load().vmvo_oracle_max_threads())
