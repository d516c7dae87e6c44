"""ctypes binding of libvmvo_b200.so (include/vmvo_b200.h).

The library is the only compute path of this package: when it is missing, or when no
CUDA device is present, every operator raises -- there is no CPU fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
from typing import Dict, Optional, Tuple

import numpy as np

from . import build as _build

_c_i32, _c_i64, _c_f64, _c_f32, _c_vp = C.c_int32, C.c_int64, C.c_double, C.c_float, C.c_void_p

VMVO_OK = 0
WIN_EMPTY, WIN_NONFINITE, WIN_TOO_LONG, WIN_NO_FRAMES = 1, 2, 4, 8
FAIL_NONE, FAIL_STEER, FAIL_ACCEL = 0, 1, 2
CSV_BAD_NUMBER, CSV_TOO_MANY_FIELDS, CSV_BAD_ROT, CSV_UNSORTED = 1, 2, 4, 8
CSV_SLOT_ROT, CSV_MAX_COLS = 1000, 64
WINDOW_FRAMES, WINDOW_TIME = 0, 1
TARGET_TIME, TARGET_TRAVERSE = 0, 1
SEED_DATA, SEED_GIVEN, SEED_CHAINED = 0, 1, 2
PRIMARY_VO, PRIMARY_GPS = 0, 1


class SearchCfg(C.Structure):
    """struct vmvo_search_cfg."""

    _fields_ = [
        ("grid_v", _c_i32), ("grid_s", _c_i32), ("window_mode", _c_i32), ("window_frames", _c_i32),
        ("horizon_frames", _c_i32), ("target_mode", _c_i32), ("target_offset", _c_i32),
        ("seed_mode", _c_i32), ("primary", _c_i32), ("max_window_poses", _c_i32),
        ("horizon_time", _c_f64), ("w_vo", _c_f64), ("w_gps", _c_f64), ("w_imu", _c_f64),
        ("k_steer", _c_f64), ("wheel_base", _c_f64), ("steering_ratio", _c_f64),
        ("max_steer", _c_f64), ("max_accel", _c_f64), ("max_steer_rate", _c_f64),
    ]


MAX_MIRRORS = 16


class Exchange(C.Structure):
    """struct vmvo_exchange: the window deal and the record exchange of one sharded call."""

    _fields_ = [
        ("world", _c_i32), ("rank", _c_i32), ("block", _c_i32), ("n_peers", _c_i32),
        ("peer_records", _c_vp * MAX_MIRRORS), ("peer_flags", _c_vp * MAX_MIRRORS),
        ("local_flags", _c_vp), ("epoch", _c_vp),
    ]


# struct vmvo_window_result as a NumPy record (64 bytes)
RESULT_DTYPE = np.dtype([
    ("best_idx", np.int32), ("n_steps", np.int32), ("status", np.int32), ("n_rescored", np.int32),
    ("best_cost", np.float64), ("v_seed", np.float64), ("s_seed", np.float64),
    ("x1", np.float64), ("y1", np.float64), ("theta1", np.float64),
])
assert RESULT_DTYPE.itemsize == 64

_SIGNATURES = {
    "vmvo_abi_version": (C.c_int, []),
    "vmvo_ctx_create": (C.c_int, [C.c_int, C.POINTER(_c_vp)]),
    "vmvo_ctx_destroy": (C.c_int, [_c_vp]),
    "vmvo_last_error": (C.c_char_p, [_c_vp]),
    "vmvo_search_cfg_default": (None, [C.POINTER(SearchCfg)]),
    "vmvo_window_count": (_c_i64, [C.POINTER(SearchCfg), _c_i64]),
    "vmvo_plan_windows": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i32, _c_vp, _c_vp, _c_i64, _c_vp,
                                    _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_grid_search_f32": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i64, _c_vp, _c_vp, _c_vp, _c_vp,
                                       _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32,
                                       _c_vp]),
    "vmvo_grid_search_f64": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i64, _c_vp, _c_vp, _c_vp, _c_vp,
                                       _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32,
                                       _c_vp]),
    "vmvo_grid_search_chained": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i64, _c_vp, _c_vp, _c_vp,
                                           _c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_i64, _c_vp, _c_vp,
                                           _c_vp, _c_vp, _c_vp, _c_i32, _c_vp]),
    "vmvo_grid_search_debug_f32": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i64, _c_vp, _c_vp, _c_vp,
                                             _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                             _c_vp]),
    "vmvo_write_back_f32": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i32, _c_i64, _c_vp, _c_vp, _c_vp,
                                      _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_write_back_f64": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i32, _c_i64, _c_vp, _c_vp, _c_vp,
                                      _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_rollout_f64": (C.c_int, [_c_vp, _c_i64, _c_i32, _c_vp, _c_vp, _c_f64, _c_vp, _c_f64, _c_f64,
                                   _c_vp, _c_vp, _c_vp]),
    "vmvo_rollout_f32": (C.c_int, [_c_vp, _c_i64, _c_i32, _c_vp, _c_vp, _c_f32, _c_vp, _c_f32, _c_f32,
                                   _c_vp, _c_vp, _c_vp]),
    "vmvo_rollout_host_f64": (C.c_int, [_c_vp, _c_i32, _c_vp, _c_vp, _c_f64, _c_vp, _c_f64, _c_f64, _c_vp, _c_vp]),
    "vmvo_sequence_cost_f64": (C.c_int, [_c_vp, _c_i64, _c_i32, _c_vp, _c_f64, _c_f64, _c_vp, _c_f64,
                                         _c_vp, _c_vp]),
    "vmvo_extract_window_f64": (C.c_int, [_c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_time_extent_f64": (C.c_int, [_c_vp, _c_i64, _c_vp, _c_f64, _c_f64, _c_vp, _c_vp]),
    "vmvo_traverse_f64": (C.c_int, [_c_vp, _c_i32, _c_vp, _c_f64, _c_vp, _c_vp, _c_vp]),
    "vmvo_smooth_f64": (C.c_int, [_c_vp, _c_i32, _c_i64, _c_vp, _c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp]),
    "vmvo_vo_prepare_f64": (C.c_int, [_c_vp, _c_i32, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_f64,
                                      _c_i32, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_gps_prepare_scratch_bytes": (_c_i64, [_c_i64, _c_i32]),
    "vmvo_gps_prepare_f64": (C.c_int, [_c_vp, _c_i32, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32,
                                       _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_csv_scratch_bytes": (_c_i64, [_c_i32, _c_vp]),
    "vmvo_csv_count_rows": (C.c_int, [_c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_csv_index_rows": (C.c_int, [_c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                      _c_vp]),
    "vmvo_csv_parse_f64": (C.c_int, [_c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp,
                                     _c_i32, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp]),
    "vmvo_grid_search_sharded": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i32, _c_vp, _c_vp, _c_i64, _c_vp,
                                           _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_vp, _c_vp,
                                           C.POINTER(Exchange), _c_vp]),
    "vmvo_exchange_publish": (C.c_int, [_c_vp, C.POINTER(Exchange), _c_vp]),
    "vmvo_exchange_wait": (C.c_int, [_c_vp, C.POINTER(Exchange), _c_vp]),
    "vmvo_write_back_range": (C.c_int, [_c_vp, C.POINTER(SearchCfg), _c_i32, _c_i64, _c_i64, _c_i64, _c_vp,
                                        _c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp,
                                        _c_vp, C.POINTER(Exchange), _c_vp]),
    "vmvo_debug_set_tuning": (C.c_int, [_c_vp, C.c_char_p, _c_i32]),
    "vmvo_peer_buffer_create": (C.c_int, [_c_vp, _c_i64, C.POINTER(_c_vp), _c_vp]),
    "vmvo_peer_buffer_destroy": (C.c_int, [_c_vp, _c_vp]),
    "vmvo_peer_buffer_open": (C.c_int, [_c_vp, _c_vp, C.POINTER(_c_vp)]),
    "vmvo_peer_buffer_close": (C.c_int, [_c_vp, _c_vp]),
    "vmvo_tan_steer_f32": (C.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp]),
    "vmvo_peak_probe": (C.c_int, [_c_vp, _c_i32, _c_i32, _c_i32, _c_i32, _c_vp, _c_vp]),
    "vmvo_launch_count": (_c_i64, [_c_vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()
_contexts: Dict[int, "Context"] = {}


class VmvoError(RuntimeError):
    """A non-zero vmvo_status from the library."""


def library_path() -> str:
    # VMVO_B200_LIBRARY: development override used to A/B kernel variants
    return os.environ.get("VMVO_B200_LIBRARY") or _build.LIB_PATH


def load() -> C.CDLL:
    """dlopen the in-tree library; raise loudly when it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.isfile(path):
            raise RuntimeError(
                f"{path} is missing: the CUDA library has not been built (run "
                "`python -c 'import __graft_entry__ as g; g.build()'` or "
                "`python -m vehiclemodelvisualodometry_b200.build`). There is no CPU fallback.")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.vmvo_abi_version() != 2:
            raise RuntimeError("libvmvo_b200.so ABI version mismatch; rebuild")
        _lib = lib
        return lib


class Context:
    """One vmvo_ctx per (process, device)."""

    def __init__(self, device: int):
        lib = load()
        handle = _c_vp()
        rc = lib.vmvo_ctx_create(int(device), C.byref(handle))
        if rc != VMVO_OK or not handle:
            raise VmvoError(
                f"vmvo_ctx_create(device={device}) failed with status {rc}: a B200 (sm_100) "
                "device is required; there is no CPU fallback")
        self.lib = lib
        self.handle = handle
        self.device = int(device)

    def check(self, rc: int, what: str):
        if rc != VMVO_OK:
            msg = self.lib.vmvo_last_error(self.handle)
            raise VmvoError(f"{what}: status {rc}: {msg.decode() if msg else ''}")

    def launch_count(self) -> int:
        return int(self.lib.vmvo_launch_count(self.handle))

    def set_tuning(self, key: str, value: int) -> None:
        """Test / tuning hook (``vmvo_debug_set_tuning``); a negative value restores the default."""
        self.check(self.lib.vmvo_debug_set_tuning(self.handle, key.encode(), int(value)),
                   "vmvo_debug_set_tuning")

    @contextlib.contextmanager
    def tuning(self, **overrides):
        """``with ctx.tuning(defer_min=1): ...`` -- overrides in force inside the block only."""
        for k, v in overrides.items():
            self.set_tuning(k, v)
        try:
            yield self
        finally:
            for k in overrides:
                self.set_tuning(k, -1)

    def close(self):
        if self.handle:
            self.lib.vmvo_ctx_destroy(self.handle)
            self.handle = _c_vp()


def context(device: Optional[int] = None) -> Context:
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device visible: the VMVO operators run on a B200 only "
                           "(there is no CPU fallback)")
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def default_cfg() -> SearchCfg:
    cfg = SearchCfg()
    load().vmvo_search_cfg_default(C.byref(cfg))
    return cfg


def window_count(cfg: SearchCfg, n_frames: int) -> int:
    return int(load().vmvo_window_count(C.byref(cfg), int(n_frames)))


def ptr(t) -> Optional[int]:
    """data_ptr of a CUDA tensor, or None."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


# ---- small NumPy-in / NumPy-out operators used by the Trajectory / BicycleModel facades --------

def rollout_host_f64(steer, vel, dt: float, state0, max_steer: float, max_accel: float):
    """One control sequence through ``vmvo_rollout_host_f64``: NumPy in, (poses [n, 3], fail kind,
    fail step) out; one launch and one stream synchronisation, no tensors."""
    ctx = context()
    steer = np.ascontiguousarray(steer, dtype=np.float64)
    vel = np.ascontiguousarray(vel, dtype=np.float64)
    s0 = np.ascontiguousarray(state0, dtype=np.float64)
    n = steer.shape[0]
    out = np.empty((n, 3), dtype=np.float64)
    fail = np.zeros(2, dtype=np.int32)
    ctx.check(ctx.lib.vmvo_rollout_host_f64(ctx.handle, n, steer.ctypes.data, vel.ctypes.data, float(dt),
                                            s0.ctypes.data, float(max_steer), float(max_accel),
                                            out.ctypes.data, fail.ctypes.data), "vmvo_rollout_host_f64")
    return out, int(fail[0]), int(fail[1])


def _dev(a, dtype):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).cuda()


def extract_window_f64(x, y, theta) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Trajectory.sub_trajectory's transform on the GPU (vmvo/schema.py:59-115)."""
    import torch

    ctx = context()
    n = len(x)
    if n < 1:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    dx, dy, dth = _dev(x, np.float64), _dev(y, np.float64), _dev(theta, np.float64)
    out = torch.empty((3, n), dtype=torch.float64, device=dx.device)
    ctx.check(ctx.lib.vmvo_extract_window_f64(ctx.handle, n, ptr(dx), ptr(dy), ptr(dth), ptr(out[0]),
                                              ptr(out[1]), ptr(out[2]), stream_ptr()),
              "vmvo_extract_window_f64")
    o = out.cpu().numpy()
    return o[0], o[1], o[2]


def time_extent_f64(time, t0: float, t1: float) -> Tuple[int, int]:
    import torch

    ctx = context()
    dt = _dev(time, np.float64)
    ext = torch.empty(2, dtype=torch.int64, device=dt.device)
    ctx.check(ctx.lib.vmvo_time_extent_f64(ctx.handle, dt.numel(), ptr(dt), float(t0), float(t1),
                                           ptr(ext), stream_ptr()), "vmvo_time_extent_f64")
    s, e = ext.cpu().tolist()
    return int(s), int(e)


def traverse_f64(xy, D: float) -> np.ndarray:
    """Indices kept by traverse_trajectory (vmvo/utils/mpc.py:125-141)."""
    import torch

    ctx = context()
    dxy = _dev(xy, np.float64)
    n = dxy.shape[0]
    keep = torch.empty(max(n, 1), dtype=torch.int32, device=dxy.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=dxy.device)
    ctx.check(ctx.lib.vmvo_traverse_f64(ctx.handle, n, ptr(dxy), float(D), ptr(keep), ptr(cnt),
                                        stream_ptr()), "vmvo_traverse_f64")
    c = int(cnt.item())
    return keep[:c].cpu().numpy().astype(np.int64)
