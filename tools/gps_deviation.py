"""Measures how far the GPU GPS pre-processing is from the reference-pinned golden vectors
(tests/golden/prep_kats.json): positions, speed, heading -- the bounds tests/test_prep.py asserts."""
import json
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import unhex  # noqa: E402
from tests.test_prep import G, _gps_frame  # noqa: E402
from vehiclemodelvisualodometry_b200.trajectory import process_gps_trajectory  # noqa: E402

for c in G["gps"]:
    lat, lon, heading, speed, stamp = _gps_frame(c["n"], c["seed"])
    df = pd.DataFrame({"heading": heading, "Latitude": lat, "Longitude": lon, "speed": speed, "Timestamp": stamp})
    tr = process_gps_trajectory(df)
    dx = np.max(np.abs(np.asarray(tr.x) - unhex(c["x"])))
    dy = np.max(np.abs(np.asarray(tr.y) - unhex(c["y"])))
    wv = unhex(c["velocity"])
    dv = np.abs(np.asarray(tr.velocity) - wv)
    rel = np.max(dv / np.maximum(np.abs(wv), 1e-300))
    dth = np.abs(np.asarray(tr.theta) - unhex(c["theta"]))
    dth = np.minimum(dth, 2 * np.pi - dth)
    step = np.hypot(np.diff(unhex(c["x"])), np.diff(unhex(c["y"])))
    print(f"n={c['n']}: max|dx| {dx:.3e} m  max|dy| {dy:.3e} m  max|dv| {dv.max():.3e} (rel {rel:.3e}, "
          f"v up to {np.abs(wv).max():.3e})  heading: max {dth.max():.3e} rad; where the smoothed step > 1e-4 m: "
          f"{dth[step > 1e-4].max() if (step > 1e-4).any() else 0:.3e}; max dth*step {np.max(dth * step):.3e} m")
