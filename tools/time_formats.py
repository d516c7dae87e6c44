"""Times the CSV reader on a batch of synthetic drives (device-resident bytes -> columns) and the
reference's own reader (pandas.read_csv + parse_rot) on the host.

    python tools/time_formats.py [n_drives] [rows]
"""
import io
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.csv_oracle import parse_rot  # noqa: E402
from oracle.make_golden_prep import gps_frame, vo_frame  # noqa: E402
from vehiclemodelvisualodometry_b200.dataset import CACHE_COLUMNS, LOG_COLUMNS, parse_staged, stage_csv_files  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
x, y, rot, _ = vo_frame(n, 1)
lat, lon, heading, speed, stamp = gps_frame(n, 2)
b1, b2 = io.StringIO(), io.StringIO()
pd.DataFrame({"Timestamp": stamp, "Latitude": lat, "Longitude": lon, "heading": heading, "speed": speed}).to_csv(b1, index=False)
pd.DataFrame({"x": list(x), "y": list(y), "z": list(x * 0), "rot": [r for r in rot]}).to_csv(b2, index=False)
log, cache = b1.getvalue().encode(), b2.getvalue().encode()
print(f"log {len(log)} B ({len(log) / n:.0f} B/row), cache {len(cache)} B ({len(cache) / n:.0f} B/row), {D} drives")

for name, blob, cols, rc, sc in (("log", log, LOG_COLUMNS, None, "Timestamp"), ("cache", cache, CACHE_COLUMNS, "rot", None)):
    st = stage_csv_files([blob] * D)
    for _ in range(2):
        p = parse_staged(st, cols, rc, sc)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        p = parse_staged(st, cols, rc, sc)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = min(ms)
    print(f"GPU {name}: {t:.3f} ms  {st.n_bytes / t / 1e6:.1f} GB/s  {D * n / t / 1e3:.1f} M rows/s")
    t0 = time.perf_counter()
    df = pd.read_csv(io.BytesIO(blob))
    if rc:
        df[rc] = df[rc].apply(parse_rot)
    dt = time.perf_counter() - t0
    print(f"CPU {name} (pandas, 1 drive, 1 thread): {dt * 1e3:.1f} ms  {len(blob) / dt / 1e6:.1f} MB/s")
