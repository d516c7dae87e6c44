"""Writes the fixture drives under tests/golden/csv/ and pins oracle/csv_oracle.py to the
UNMODIFIED reference reader: ``AndroidDatasetIterator`` (vmvo/datasets/bdd/bdd_raw.py:19-171) is
constructed on the fixture folders and its ``.csv_dat`` / ``.trajectory`` frames, then
``process_vo_trajectory`` / ``process_gps_trajectory`` of them, are frozen in
tests/golden/csv_kats.json.  Build container only (needs /root/reference).  TEST INFRASTRUCTURE.

    python oracle/make_golden_csv.py

Environment shims (none of them touches the reference's code):
  * cv2.VideoCapture -> a stub reporting 30 fps and a long video (the reader asserts on the .mp4,
    bdd_raw.py:48-63; video decoding is out of scope);
  * pandas 3 removed the positional fallback of ``Series[int]`` that bdd_raw.py:95 relies on
    (``self.csv_dat.loc[key][0]``): restored for non-integer indexes, as pandas < 3 behaved;
  * a camera calibration YAML with the keys bdd_raw.py:109-133 reads.
The cache file is written the way the reference writes it (bdd_raw.py:241, 305-332:
``pd.DataFrame({"x", "y", "z", "rot"}).to_csv(path, index=False)`` with rot = float64 3x3 arrays).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import csv_oracle as C  # noqa: E402
from oracle import ref_bridge  # noqa: E402
from oracle.make_golden import hexf  # noqa: E402
from oracle.make_golden_prep import gps_frame, vo_frame  # noqa: E402

FIXTURES = os.path.join(ROOT, "tests", "golden", "csv")
GOLDEN = os.path.join(ROOT, "tests", "golden", "csv_kats.json")
DRIVES = (("1658384707877", 64, 11), ("1654493684259", 41, 12))


def write_drive(folder: str, ident: str, n: int, seed: int) -> None:
    os.makedirs(folder, exist_ok=True)
    x, y, rot, _ = vo_frame(n, seed)
    lat, lon, heading, speed, stamp = gps_frame(n, seed + 100)
    rng = np.random.default_rng(seed)
    log = pd.DataFrame({"Timestamp": stamp, "Latitude": lat, "Longitude": lon, "heading": heading,
                        "speed": speed, "accuracy": rng.uniform(2, 9, n).round(1),
                        "altitude": rng.normal(900, 3, n)})
    log.to_csv(os.path.join(folder, ident + ".csv"), index=False)
    traj = {"x": list(x), "y": list(y), "z": list(rng.normal(0, 0.01, n)), "rot": [r for r in rot]}
    pd.DataFrame(traj).to_csv(os.path.join(folder, ident + "_traj.csv"), index=False)


def install_shims():
    import cv2

    class Capture:
        def __init__(self, path):
            self.path = path

        def get(self, prop):
            return 30.0 if prop == cv2.CAP_PROP_FPS else 10 ** 7

    cv2.VideoCapture = Capture
    orig = pd.Series.__getitem__

    def getitem(self, key):
        try:
            return orig(self, key)
        except KeyError:
            if isinstance(key, (int, np.integer)) and not pd.api.types.is_integer_dtype(self.index.dtype):
                return self.iloc[key]
            raise

    pd.Series.__getitem__ = getitem


def main():
    assert ref_bridge.available()
    ref_bridge.load()
    install_shims()
    import vmvo.utils.trajectory as T
    from vmvo.datasets.bdd.bdd_raw import AndroidDatasetIterator

    out = {"generator": "oracle/make_golden_csv.py (AndroidDatasetIterator of the unmodified reference)",
           "pandas": pd.__version__, "numpy": np.__version__, "drives": []}
    with tempfile.TemporaryDirectory() as tmp:
        calib = os.path.join(tmp, "calib.yaml")
        with open(calib, "w") as f:
            f.write("Camera.k1: 0.0\nCamera.k2: 0.0\nCamera.p1: 0.0\nCamera.p2: 0.0\n"
                    "Camera.fx: 500.0\nCamera.fy: 500.0\nCamera.cx: 320.0\nCamera.cy: 240.0\n")
        for ident, n, seed in DRIVES:
            folder = os.path.join(FIXTURES, ident)
            write_drive(folder, ident, n, seed)
            mp4 = os.path.join(folder, ident + ".mp4")
            open(mp4, "wb").close()                     # bdd_raw.py:48 only asserts that it exists
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    ds = AndroidDatasetIterator(folder_path=folder, settings_doc=calib, compute_trajectory=True)
                    vo = T.process_vo_trajectory(ds.trajectory)
                    gps = T.process_gps_trajectory(ds.csv_dat)
            finally:
                os.unlink(mp4)
            # the oracle restatement against the frames the reference built
            log = C.read_csv(open(os.path.join(folder, ident + ".csv"), "rb").read())
            cache = C.read_csv(open(os.path.join(folder, ident + "_traj.csv"), "rb").read(), ("x", "y", "z"), "rot")
            assert list(ds.csv_dat.columns) == list(log)
            for k in log:
                assert np.array_equal(ds.csv_dat[k].to_numpy().astype(np.float64), log[k]), k
            for k in ("x", "y", "z"):
                assert np.array_equal(ds.trajectory[k].to_numpy(), cache[k]), k
            ref_rot = np.stack(ds.trajectory["rot"].tolist())
            assert ref_rot.dtype == np.float32 and np.array_equal(ref_rot, np.stack(cache["rot"]))
            assert np.array_equal(ds.trajectory["Timestamp"].to_numpy(), ds.csv_dat["Timestamp"].to_numpy())
            assert isinstance(vo.theta[0], float)
            out["drives"].append({
                "id": ident, "n": n,
                "log": {k: hexf(v) for k, v in log.items()},
                "cache": {k: hexf(cache[k]) for k in ("x", "y", "z")},
                "rot": hexf(ref_rot.astype(np.float64)),
                "vo": {k: hexf(getattr(vo, k)) for k in ("x", "y", "theta", "velocity", "time")},
                "gps": {k: hexf(getattr(gps, k)) for k in ("x", "y", "theta", "velocity", "time")},
            })
    with open(GOLDEN, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes;", FIXTURES)


if __name__ == "__main__":
    main()
