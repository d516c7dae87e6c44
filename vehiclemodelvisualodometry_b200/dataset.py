"""The reference's on-disk formats, read on the GPU.

``<id>.csv`` is the Android log the reference loads with ``pd.read_csv`` and sorts by Timestamp
(vmvo/datasets/bdd/bdd_raw.py:53-55); ``<id>_traj.csv`` is the cached VO trajectory with its
stringified 3x3 ``rot`` column (bdd_raw.py:150-168, written at :331-332).  The bytes go to HBM
once; row indexing, field splitting and number conversion run in CUDA
(``vmvo_csv_count_rows`` / ``vmvo_csv_index_rows`` / ``vmvo_csv_parse_f64``), converting numbers
exactly as the reference's readers do (pandas' default converter; NumPy's float32 cast for
``rot``).  Only the header line of each file is looked at on the host, to map column names.

``load_android_drive`` returns the two DataFrames ``AndroidDatasetIterator`` exposes as
``.trajectory`` and ``.csv_dat`` (what ``process_vo_trajectory`` / ``process_gps_trajectory``
take); ``load_android_drives_device`` keeps everything on the device for the batched chain
parse -> pre-process -> search (``optimize_android_drives``).  Video decoding and the VO
front-end that WRITES the cache are out of scope.
"""
from __future__ import annotations

import csv
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

LOG_COLUMNS = ("Timestamp", "Latitude", "Longitude", "heading", "speed")   # trajectory.py:191-228
CACHE_COLUMNS = ("x", "y", "z")                                            # bdd_raw.py:241
Blob = Union[bytes, bytearray, memoryview, str, os.PathLike]


def _blob(src: Blob) -> bytes:
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    with open(src, "rb") as f:
        return f.read()


def _header(blob: bytes) -> List[str]:
    """Column names: the first non-blank line (headers hold no quoted newlines)."""
    start = 0
    while start < len(blob) and blob[start] in b"\r\n":
        start += 1
    end = blob.find(b"\n", start)
    line = blob[start:len(blob) if end < 0 else end].rstrip(b"\r")
    if not line:
        return []
    return next(csv.reader([line.decode("utf-8")]))


@dataclass
class ParsedCsv:
    """Numeric columns of a batch of CSV files, resident on the device."""

    names: Tuple[str, ...]
    columns: torch.Tensor                  # float64 [len(names), n_data]
    rot: Optional[torch.Tensor]            # float64 [n_data, 9] (float32 values) or None
    row_offsets: List[int]                 # data rows of file f: [row_offsets[f], row_offsets[f + 1])
    status: np.ndarray                     # int32 [n_files]: _lib.CSV_* bits

    def column(self, name: str) -> torch.Tensor:
        return self.columns[self.names.index(name)]

    @property
    def n_files(self) -> int:
        return len(self.row_offsets) - 1


@dataclass
class CsvStage:
    """The bytes of a batch of files in HBM (every file on a 16-byte boundary) + their headers."""

    d_bytes: torch.Tensor                  # uint8
    offs: np.ndarray                       # int64 [n] host
    lens: np.ndarray                       # int64 [n] host
    d_off: torch.Tensor
    d_len: torch.Tensor
    headers: List[List[str]]

    @property
    def n_files(self) -> int:
        return len(self.headers)

    @property
    def n_bytes(self) -> int:
        return int(self.lens.sum())


def stage_csv_files(sources: Sequence[Blob], device=None) -> CsvStage:
    """Host -> HBM: one pinned staging buffer, one copy."""
    import pandas.errors as pe

    ctx = _lib.context(None if device is None else torch.device(device).index)
    dev = torch.device("cuda", ctx.device)
    blobs = [_blob(s) for s in sources]
    n = len(blobs)
    if n < 1:
        raise ValueError("no files")
    headers = [_header(b) for b in blobs]
    for h in headers:
        if not h:
            raise pe.EmptyDataError("No columns to parse from file")
    lens = np.array([len(b) for b in blobs], dtype=np.int64)
    offs = np.zeros(n, dtype=np.int64)
    pos = 0
    for f in range(n):
        offs[f] = pos
        pos += (int(lens[f]) + 15) // 16 * 16
    stage = torch.zeros(max(pos, 16), dtype=torch.uint8).pin_memory()
    view = stage.numpy()
    for f, b in enumerate(blobs):
        view[offs[f]:offs[f] + lens[f]] = np.frombuffer(b, dtype=np.uint8)
    return CsvStage(d_bytes=stage.to(dev, non_blocking=True), offs=offs, lens=lens,
                    d_off=torch.from_numpy(offs).to(dev), d_len=torch.from_numpy(lens).to(dev),
                    headers=headers)


def parse_staged(stage: CsvStage, wanted: Optional[Sequence[str]] = None, rot_column: Optional[str] = None,
                 sorted_column: Optional[str] = None, check: bool = True) -> ParsedCsv:
    """Row index + field split + number conversion of staged files, all on the device; the only
    host round trip is the per-file row count that sizes the outputs."""
    import pandas.errors as pe

    dev = stage.d_bytes.device
    ctx = _lib.context(dev.index)
    n, headers = stage.n_files, stage.headers
    if wanted is None:
        wanted = [c for c in headers[0] if c != rot_column]
    wanted = tuple(wanted)
    if len(wanted) > 32:
        raise ValueError("at most 32 numeric columns per call")
    colmap = np.full((n, _lib.CSV_MAX_COLS), -1, dtype=np.int32)
    for f, h in enumerate(headers):
        if len(h) > _lib.CSV_MAX_COLS:
            raise ValueError(f"file {f}: more than {_lib.CSV_MAX_COLS} columns")
        for name in wanted + ((rot_column,) if rot_column else ()):
            if name not in h:
                raise KeyError(name)
        for c, name in enumerate(h):
            if name in wanted:
                colmap[f, c] = wanted.index(name)
            elif rot_column is not None and name == rot_column:
                colmap[f, c] = _lib.CSV_SLOT_ROT
    n_fields = np.array([len(h) for h in headers], dtype=np.int32)
    d_bytes, d_off, d_len = stage.d_bytes, stage.d_off, stage.d_len
    h_off, h_len = stage.offs.ctypes.data, stage.lens.ctypes.data
    sp = _lib.stream_ptr(dev)
    scratch = torch.empty(int(ctx.lib.vmvo_csv_scratch_bytes(n, h_len)), dtype=torch.uint8, device=dev)
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    ctx.check(ctx.lib.vmvo_csv_count_rows(ctx.handle, _lib.ptr(d_bytes), n, h_off, h_len, _lib.ptr(d_off),
                                          _lib.ptr(d_len), _lib.ptr(scratch), _lib.ptr(counts), sp),
              "vmvo_csv_count_rows")
    rows = counts.cpu().numpy()                       # the one synchronisation: sizes the outputs
    row_off = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    total = int(row_off[-1])
    d_row_off = torch.from_numpy(row_off).to(dev)
    row_starts = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    ctx.check(ctx.lib.vmvo_csv_index_rows(ctx.handle, _lib.ptr(d_bytes), n, h_off, h_len, _lib.ptr(d_off),
                                          _lib.ptr(d_len), _lib.ptr(scratch), _lib.ptr(d_row_off),
                                          _lib.ptr(row_starts), sp), "vmvo_csv_index_rows")
    n_data = total - n
    cols = torch.empty((len(wanted), max(n_data, 0)), dtype=torch.float64, device=dev)
    rot = torch.empty((max(n_data, 0), 9), dtype=torch.float64, device=dev) if rot_column else None
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    sorted_slot = wanted.index(sorted_column) if sorted_column is not None else -1
    d_colmap, d_n_fields = torch.from_numpy(colmap).to(dev), torch.from_numpy(n_fields).to(dev)
    ctx.check(ctx.lib.vmvo_csv_parse_f64(
        ctx.handle, _lib.ptr(d_bytes), n, _lib.ptr(d_off), _lib.ptr(d_len), _lib.ptr(d_row_off),
        _lib.ptr(row_starts), total, _lib.ptr(d_colmap), _lib.ptr(d_n_fields), len(wanted), sorted_slot,
        _lib.ptr(cols), _lib.ptr(rot), _lib.ptr(status), sp), "vmvo_csv_parse_f64")
    st = status.cpu().numpy()
    if check:
        for f in range(n):
            if st[f] & _lib.CSV_TOO_MANY_FIELDS:
                raise pe.ParserError(f"Error tokenizing data: file {f} has a row with more than "
                                     f"{n_fields[f]} fields")
            if st[f] & _lib.CSV_BAD_NUMBER:
                raise ValueError(f"file {f}: a field of a numeric column is not a number")
            if st[f] & _lib.CSV_BAD_ROT:
                raise ValueError(f"file {f}: a rot field does not hold nine numbers")
    data_off = [int(row_off[f] - f) for f in range(n + 1)]
    return ParsedCsv(names=wanted, columns=cols, rot=rot, row_offsets=data_off, status=st)


def parse_csv_files(sources: Sequence[Blob], wanted: Optional[Sequence[str]] = None,
                    rot_column: Optional[str] = None, sorted_column: Optional[str] = None,
                    device=None, check: bool = True) -> ParsedCsv:
    """Parse the named numeric columns (default: every column of the first file except
    ``rot_column``) of a batch of CSV files on the GPU."""
    return parse_staged(stage_csv_files(sources, device), wanted, rot_column, sorted_column, check)


def read_csv(source: Blob, columns: Optional[Sequence[str]] = None, rot_column: Optional[str] = None):
    """One file -> ``pandas.DataFrame`` of float64 columns (and, for ``rot_column``, a column of
    float32 3x3 arrays as ``parse_rot`` makes them, bdd_raw.py:157-167)."""
    import pandas as pd

    p = parse_csv_files([source], columns, rot_column)
    host = p.columns.cpu().numpy()
    data: Dict[str, object] = {name: host[i] for i, name in enumerate(p.names)}
    if p.rot is not None:
        r = p.rot.cpu().numpy().astype(np.float32).reshape(-1, 3, 3)
        data[rot_column] = list(r)
    return pd.DataFrame(data)


def _drive_paths(folder: Union[str, os.PathLike]) -> Tuple[str, str]:
    folder = os.fspath(folder).rstrip("/")
    ident = folder.split("/")[-1]                                # bdd_raw.py:44-45
    log = os.path.join(folder, ident + ".csv")
    return log, log.replace(".csv", "_traj.csv")                # bdd_raw.py:144-147


def _unsorted_error(path):
    return NotImplementedError(
        f"{path}: the log is not in Timestamp order.  The reference sorts it (bdd_raw.py:55) but then "
        "mixes label-based and positional access (trajectory.py:191-207), so its own result is only "
        "meaningful for a log already in time order")


def load_android_drive(folder: Union[str, os.PathLike]):
    """``(trajectory, csv_dat)`` of one drive, like ``AndroidDatasetIterator(folder,
    compute_trajectory=True)`` exposes them (bdd_raw.py:53-55, 150-168)."""
    import pandas as pd

    log_path, cache_path = _drive_paths(folder)
    log_blob = _blob(log_path)
    log = parse_csv_files([log_blob], None, None, sorted_column=_header(log_blob)[0])
    if log.status[0] & _lib.CSV_UNSORTED:
        raise _unsorted_error(log_path)
    host = log.columns.cpu().numpy()
    csv_dat = pd.DataFrame({name: host[i] for i, name in enumerate(log.names)})
    if "Timestamp" in csv_dat and np.all(np.isfinite(csv_dat["Timestamp"])):
        csv_dat["Timestamp"] = csv_dat["Timestamp"].astype(np.int64)   # millisecond epoch stamps
    traj = read_csv(cache_path, CACHE_COLUMNS, "rot")
    traj["Timestamp"] = csv_dat["Timestamp"]                     # bdd_raw.py:168 (aligned on the index)
    return traj, csv_dat


@dataclass
class AndroidDrives:
    """A batch of drives on the device, columns ready for the pre-processing kernels."""

    offsets: List[int]                     # frames of drive d: [offsets[d], offsets[d + 1])
    d_offsets: torch.Tensor                # int64 [D + 1]
    log: ParsedCsv                         # Timestamp, Latitude, Longitude, heading, speed
    cache: ParsedCsv                       # x, y, z and rot

    @property
    def n_drives(self) -> int:
        return len(self.offsets) - 1


def load_android_drives_device(folders: Sequence[Union[str, os.PathLike]], device=None) -> AndroidDrives:
    """Both files of every drive parsed in two batched passes; nothing returns to the host."""
    paths = [_drive_paths(f) for f in folders]
    log = parse_csv_files([p[0] for p in paths], LOG_COLUMNS, None, sorted_column="Timestamp",
                          device=device)
    for f, p in enumerate(paths):
        if log.status[f] & _lib.CSV_UNSORTED:
            raise _unsorted_error(p[0])
    cache = parse_csv_files([p[1] for p in paths], CACHE_COLUMNS, "rot", device=device)
    if cache.row_offsets != log.row_offsets:
        raise ValueError("every drive needs one cached VO pose per log row "
                         f"(log rows {np.diff(log.row_offsets)}, poses {np.diff(cache.row_offsets)})")
    return AndroidDrives(offsets=list(log.row_offsets),
                         d_offsets=torch.tensor(log.row_offsets, dtype=torch.int64, device=log.columns.device),
                         log=log, cache=cache)


def prepare_android_drives(drives: AndroidDrives, scale: float = 0.25, smoothen_window: int = 20):
    """process_vo_trajectory + process_gps_trajectory of the whole batch on the device, packed as
    float64 pose streams: returns ``(DriveSet, fps)`` -- time = the GPS stamps, FPS per drive as
    optimize_trajectory_v2.py:37-42 derives it -- with N = min(len(vo), len(gps)) = the log's row
    count per drive."""
    from .search import DriveSet
    from .trajectory import gps_prepare_device, vo_prepare_device

    D, F = drives.n_drives, drives.offsets[-1]
    lg, ch = drives.log, drives.cache
    stamp = lg.column("Timestamp").contiguous()
    vo = vo_prepare_device(drives.d_offsets, D, F, ch.column("x").contiguous(), ch.column("y").contiguous(),
                           ch.rot, stamp, scale, smoothen_window, yaw_f32=True)
    gps, status, _ = gps_prepare_device(drives.d_offsets, D, F, lg.column("Latitude").contiguous(),
                                        lg.column("Longitude").contiguous(), lg.column("speed").contiguous(),
                                        stamp, smoothen_window)
    st = status.cpu().numpy()
    for d in range(D):
        if st[d]:
            n = drives.offsets[d + 1] - drives.offsets[d]
            raise IndexError(f"index {n} is out of bounds for axis 0 with size {n}")
    dev = vo.device
    # drive d owns rows [off[d] + d, off[d + 1] + d + 1) of the GPS output (n + 1 points); the
    # optimizer uses the first n of them
    idx = torch.arange(F, device=dev) + torch.repeat_interleave(
        torch.arange(D, device=dev), torch.tensor(np.diff(drives.offsets), device=dev))
    g = gps[:, idx]
    vo_stream = torch.stack([vo[0], vo[1], vo[2], vo[3]], dim=1).contiguous()
    gps_stream = torch.stack([g[0], g[1], g[2], g[3]], dim=1).contiguous()
    time = g[4].contiguous()
    # dt per drive: 1 / FPS with FPS = 1 / mean(diff(gps time)) (optimize_trajectory_v2.py:37-42)
    # FPS from ALL n + 1 GPS stamps of a drive, like optimize_trajectory_v2.py:37 (the window
    # stamps are the first n of them)
    fps = []
    t_all = gps[4].cpu().numpy()
    for d in range(D):
        a, b = drives.offsets[d] + d, drives.offsets[d + 1] + d + 1
        fps.append(1 / np.mean(np.diff(t_all[a:b])))
    ds = DriveSet(time=time, vo=vo_stream, gps=gps_stream, imu=None, drive_offsets=list(drives.offsets),
                  d_drive_offsets=drives.d_offsets,
                  dt=torch.tensor([1.0 / f for f in fps], dtype=torch.float64, device=dev))
    return ds, fps


def optimize_android_drives(folders: Sequence[Union[str, os.PathLike]], config=None, scale: float = 0.25,
                            smoothen_window: int = 20):
    """The reference's ``main`` (optimize_trajectory_v2.py:168-183) for a batch of drives, device
    resident from the file bytes on: parse -> pre-process -> window search -> write-back.
    Returns ``(SearchOutput, trajectory float64 [4, F], WindowPlan, DriveSet)``."""
    from dataclasses import replace

    from .optimize import HORIZON_TIME, REFERENCE_CFG
    from .search import optimize_drives

    cfg = config if config is not None else REFERENCE_CFG
    ds, fps = prepare_android_drives(load_android_drives_device(folders), scale, smoothen_window)
    horizons = {int(HORIZON_TIME * f) for f in fps}              # optimize_trajectory_v2.py:35-42
    if len(horizons) != 1:
        raise ValueError(f"drives logged at different rates (horizons {sorted(horizons)}): "
                         "optimize them in separate batches")
    if cfg.window_mode == "time":
        cfg = replace(cfg, horizon_time=HORIZON_TIME, horizon_frames=horizons.pop())
    so, traj, plan = optimize_drives(cfg, ds)
    return so, traj, plan, ds
