"""The reference's on-disk formats: <id>.csv (the Android log) and <id>_traj.csv
(the cached VO trajectory with its stringified rot column).

CPU: oracle/csv_oracle.py against the vectors frozen from the reference's own AndroidDatasetIterator
(tests/golden/csv_kats.json, fixtures under tests/golden/csv/) and against pandas.read_csv, the
third-party reader the reference calls.  GPU: the CUDA parser (through the C ABI) against the same
vectors, against pandas on adversarial files, and the device-resident chain file bytes -> optimized
trajectory against the host facades."""
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import csv_oracle as C
from tests.helpers import unhex

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "csv")
G = json.load(open(os.path.join(HERE, "golden", "csv_kats.json")))
COLS = ("x", "y", "theta", "velocity", "time")


def _paths(d):
    folder = os.path.join(FIX, d["id"])
    return folder, os.path.join(folder, d["id"] + ".csv"), os.path.join(folder, d["id"] + "_traj.csv")


def _number_strings(rng, n):
    """Decimal strings that exercise the converter: 17-digit reprs, exponents, integers, signs."""
    out = []
    for i in range(n):
        k = i % 8
        v = float(rng.normal(0, 1) * 10.0 ** int(rng.integers(-12, 13)))
        if k == 0:
            out.append(repr(v))
        elif k == 1:
            out.append("%.3f" % v)
        elif k == 2:
            out.append("%.17e" % v)
        elif k == 3:
            out.append(str(int(rng.integers(-10 ** 15, 10 ** 15))))
        elif k == 4:
            out.append("+" + "%.9g" % abs(v))
        elif k == 5:
            out.append("%.25f" % v)                       # more than 17 digits
        elif k == 6:
            out.append(repr(float(rng.normal(0, 1) * 10.0 ** int(rng.integers(-300, 300)))))
        else:
            out.append("%dE%d" % (int(rng.integers(1, 10 ** 6)), int(rng.integers(-340, 320))))
    return out


def _adversarial_csv(seed=0, n=400):
    """Quoted numbers, NA words, empty fields, short rows, CRLF, blank lines, no final newline."""
    rng = np.random.default_rng(seed)
    nums = _number_strings(rng, 4 * n)
    lines = ["a,b,c,d"]
    for i in range(n):
        f = nums[4 * i:4 * i + 4]
        if i % 7 == 1:
            f[1] = ""
        if i % 11 == 2:
            f[2] = ["NaN", "NA", "null", "nan", "N/A", "#N/A", "None", "<NA>"][i % 8]
        if i % 13 == 3:
            f[0] = '"%s"' % f[0]
        if i % 17 == 4:
            f[3] = ["inf", "-inf", "+Infinity", "INF"][i % 4]
        if i % 23 == 6 and i % 17 != 4:
            f[3] = "  %s  " % f[3]                          # (pandas does not trim around "inf")
        if i % 19 == 5:
            f = f[:2]                                      # short row: padded with NaN
        lines.append(",".join(f))
        if i % 29 == 7:
            lines.append("")                               # blank line: skipped
    text = "\r\n".join(lines[:n // 2]) + "\r\n" + "\n".join(lines[n // 2:])
    return text.encode()


# ---- CPU: the oracle ------------------------------------------------------------------------------
def test_oracle_reproduces_the_reference_reader_on_the_fixtures():
    assert G["drives"]
    for d in G["drives"]:
        _, log_path, cache_path = _paths(d)
        log = C.read_csv(open(log_path, "rb").read())
        cache = C.read_csv(open(cache_path, "rb").read(), ("x", "y", "z"), "rot")
        assert list(log) == list(d["log"])
        for k, v in d["log"].items():
            np.testing.assert_array_equal(log[k], unhex(v))
        for k, v in d["cache"].items():
            np.testing.assert_array_equal(cache[k], unhex(v))
        rot = np.stack(cache["rot"])
        assert rot.dtype == np.float32 and rot.shape == (d["n"], 3, 3)
        np.testing.assert_array_equal(rot.astype(np.float64).reshape(-1), unhex(d["rot"]))


def test_oracle_converter_equals_pandas_default():
    """precise_xstrtod restated == pandas.read_csv, including where pandas is NOT correctly rounded."""
    rng = np.random.default_rng(1)
    s = _number_strings(rng, 20000)
    got = pd.read_csv(io.StringIO("v\n" + "\n".join(s) + "\n"))["v"].to_numpy(dtype=np.float64)
    mine = np.array([C.to_double(t)[1] for t in s])
    np.testing.assert_array_equal(mine, got)
    exact = np.array([float(t) for t in s])
    assert np.any(mine != exact), "the sample should include inputs pandas rounds differently from float()"


def test_oracle_file_equals_pandas_on_adversarial_input():
    data = _adversarial_csv()
    want = pd.read_csv(io.BytesIO(data))
    got = C.read_csv(data)
    assert list(got) == list(want.columns) and len(want) == 400
    for k in want.columns:
        np.testing.assert_array_equal(got[k], want[k].to_numpy(dtype=np.float64))


def test_oracle_rows_and_fields():
    data = b'a,b,rot\r\n1,2,"[[1 2]\n [3 4]]"\n\n3,"x,""y""",z\n4,5,'
    rows = C.split_rows(data)
    assert rows == [b"a,b,rot", b'1,2,"[[1 2]\n [3 4]]"', b'3,"x,""y""",z', b"4,5,"]
    assert C.split_fields(rows[1]) == ["1", "2", "[[1 2]\n [3 4]]"]
    assert C.split_fields(rows[2]) == ["3", 'x,"y"', "z"]
    assert C.split_fields(rows[3]) == ["4", "5", ""]
    with pytest.raises(ValueError, match="Expected 2 fields"):
        C.read_csv(b"a,b\n1,2,3\n")
    with pytest.raises(ValueError, match="not a number"):
        C.read_csv(b"a,b\n1,x\n")


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_fixture_drives_match_the_reference_reader(cuda_device):
    """load_android_drive == AndroidDatasetIterator(...).trajectory / .csv_dat (frozen), and the two
    pre-processing facades on those frames == the reference's own outputs (frozen)."""
    from vehiclemodelvisualodometry_b200 import load_android_drive
    from vehiclemodelvisualodometry_b200.trajectory import process_gps_trajectory, process_vo_trajectory

    for d in G["drives"]:
        folder, _, _ = _paths(d)
        traj, csv_dat = load_android_drive(folder)
        assert list(csv_dat.columns) == list(d["log"]) and len(csv_dat) == d["n"]
        for k, v in d["log"].items():
            np.testing.assert_array_equal(csv_dat[k].to_numpy(dtype=np.float64), unhex(v))    # bit-exact
        assert csv_dat["Timestamp"].dtype == np.int64
        assert list(traj.columns) == ["x", "y", "z", "rot", "Timestamp"]
        for k, v in d["cache"].items():
            np.testing.assert_array_equal(traj[k].to_numpy(), unhex(v))
        rot = np.stack(traj["rot"].tolist())
        assert rot.dtype == np.float32
        np.testing.assert_array_equal(rot.astype(np.float64).reshape(-1), unhex(d["rot"]))
        np.testing.assert_array_equal(traj["Timestamp"].to_numpy(), csv_dat["Timestamp"].to_numpy())
        vo, gps = process_vo_trajectory(traj), process_gps_trajectory(csv_dat)
        for k in ("x", "y", "velocity", "time"):
            np.testing.assert_array_equal(getattr(vo, k), unhex(d["vo"][k]))
        # the yaw is a float32 atan2 on both sides (np.arctan2 of float32 scalars; atan2f): 2 ulp
        np.testing.assert_allclose(vo.theta, unhex(d["vo"]["theta"]), rtol=2.4e-7, atol=1e-9)
        assert np.array_equal(np.float32(vo.theta), np.asarray(vo.theta))   # float32 values, widened
        np.testing.assert_allclose(gps.x, unhex(d["gps"]["x"]), rtol=0, atol=1e-8)   # measured <= 2e-10 m (tools/gps_deviation.py)
        np.testing.assert_allclose(gps.y, unhex(d["gps"]["y"]), rtol=0, atol=1e-8)
        np.testing.assert_array_equal(gps.time, unhex(d["gps"]["time"]))
        assert len(gps.theta) == d["n"] and len(gps.x) == d["n"] + 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gpu_parser_equals_pandas_on_adversarial_input(cuda_device, seed):
    from vehiclemodelvisualodometry_b200 import read_csv

    data = _adversarial_csv(seed)
    want = pd.read_csv(io.BytesIO(data))
    got = read_csv(data)
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    for k in want.columns:
        np.testing.assert_array_equal(got[k].to_numpy(), want[k].to_numpy(dtype=np.float64))


@pytest.mark.gpu
def test_gpu_converter_equals_pandas_bit_for_bit(cuda_device):
    from vehiclemodelvisualodometry_b200 import read_csv

    rng = np.random.default_rng(5)
    s = _number_strings(rng, 200000)
    data = ("v\n" + "\n".join(s) + "\n").encode()
    want = pd.read_csv(io.BytesIO(data))["v"].to_numpy(dtype=np.float64)
    got = read_csv(data)["v"].to_numpy()
    np.testing.assert_array_equal(got, want)
    assert np.any(got != np.array([float(t) for t in s]))     # pandas' own last-digit behaviour kept


@pytest.mark.gpu
def test_gpu_rot_tokens_equal_numpy_float32_cast(cuda_device):
    """np.array(tokens).astype(np.float32) (bdd_raw.py:163-164) for every shape of token numpy's
    array printer produces, plus exponents and digit counts beyond the exact fast path."""
    from vehiclemodelvisualodometry_b200 import read_csv

    rng = np.random.default_rng(9)
    toks = []
    for i in range(9 * 3000):
        v = float(rng.normal(0, 1) * 10.0 ** int(rng.integers(-44, 39)))
        k = i % 6
        toks.append(["%.8e" % v, "%.8f" % (v % 1), repr(v), "%d." % int(rng.integers(-9, 9)), "%.16e" % v,
                     "%.3e" % v][k])
    toks[5], toks[17], toks[40], toks[77] = "nan", "inf", "-inf", "-0."
    # a float32 tie that only a correctly rounded double resolves (1 + 2^-24 + 1e-25)
    toks[100] = "1.0000000596046447753906251"
    rows = [toks[9 * r:9 * r + 9] for r in range(len(toks) // 9)]
    lines = ["x,rot"]
    for r, t in enumerate(rows):
        lines.append('%d,"[[%s %s %s]\n [%s  %s %s]\n [ %s %s %s]]"' % ((r,) + tuple(t)))
    got = read_csv(("\n".join(lines) + "\n").encode(), ["x"], "rot")
    want = np.array(toks).astype(np.float32).reshape(-1, 3, 3)
    have = np.stack(got["rot"].tolist())
    assert have.dtype == np.float32
    np.testing.assert_array_equal(have, want)
    np.testing.assert_array_equal(got["x"].to_numpy(), np.arange(len(rows), dtype=np.float64))


@pytest.mark.gpu
def test_gpu_many_files_columns_in_any_order_and_block_straddling_quotes(cuda_device):
    """A batch of cache files much longer than one 4 KiB block (quoted three-line fields straddle
    block boundaries), every file with its own column order."""
    from oracle.make_golden_prep import vo_frame
    from vehiclemodelvisualodometry_b200 import parse_csv_files

    blobs, want = [], []
    orders = (["x", "y", "z", "rot"], ["rot", "z", "y", "x"], ["y", "rot", "x", "z"], ["z", "x", "rot", "y"])
    for f, n in enumerate((700, 3, 1500, 64)):
        x, y, rot, _ = vo_frame(n, 40 + f)
        df = pd.DataFrame({"x": list(x), "y": list(y), "z": list(x * 0.5), "rot": [r for r in rot]})[orders[f]]
        buf = io.StringIO()
        df.to_csv(buf, index=False)
        blobs.append(buf.getvalue().encode())
        want.append(C.read_csv(blobs[-1], ("x", "y", "z"), "rot"))
    assert max(len(b) for b in blobs) > 40 * 4096
    p = parse_csv_files(blobs, ("x", "y", "z"), "rot")
    assert p.row_offsets == [0, 700, 703, 2203, 2267] and not p.status.any()
    cols, rot = p.columns.cpu().numpy(), p.rot.cpu().numpy()
    for f, w in enumerate(want):
        a, b = p.row_offsets[f], p.row_offsets[f + 1]
        for i, k in enumerate(("x", "y", "z")):
            np.testing.assert_array_equal(cols[i, a:b], w[k])
        np.testing.assert_array_equal(rot[a:b].reshape(-1, 3, 3), np.stack(w["rot"]).astype(np.float64))


@pytest.mark.gpu
def test_gpu_reader_errors(cuda_device):
    import pandas.errors as pe

    from vehiclemodelvisualodometry_b200 import load_android_drive, parse_csv_files, read_csv

    with pytest.raises(pe.ParserError):
        read_csv(b"a,b\n1,2\n3,4,5\n")
    with pytest.raises(ValueError, match="not a number"):
        read_csv(b"a,b\n1,2\n3,four\n")
    with pytest.raises(ValueError, match="nine numbers"):
        read_csv(b'x,rot\n1,"[[1 2 3]\n [4 5 6]]"\n', ["x"], "rot")
    with pytest.raises(KeyError):
        read_csv(b"a,b\n1,2\n", ["a", "speed"])
    with pytest.raises(pe.EmptyDataError):
        read_csv(b"\n\n")
    empty = read_csv(b"a,b\n")                                   # header only: no rows
    assert list(empty.columns) == ["a", "b"] and len(empty) == 0
    p = parse_csv_files([b"t,v\n1,5\n3,6\n2,7\n", b"t,v\n1,1\n2,2\n"], ("t", "v"), sorted_column="t")
    assert list(p.status) == [8, 0]


@pytest.mark.gpu
def test_gpu_unsorted_log_is_refused(cuda_device, tmp_path):
    from vehiclemodelvisualodometry_b200 import load_android_drive

    d = tmp_path / "1650000000000"
    d.mkdir()
    (d / "1650000000000.csv").write_text("Timestamp,Latitude,Longitude,heading,speed\n2,1,1,0,1\n1,1,1,0,1\n")
    (d / "1650000000000_traj.csv").write_text('x,y,z,rot\n0,0,0,"[[1. 0. 0.]\n [0. 1. 0.]\n [0. 0. 1.]]"\n')
    with pytest.raises(NotImplementedError, match="Timestamp order"):
        load_android_drive(d)


@pytest.mark.gpu
def test_gpu_device_chain_equals_host_facades(cuda_device, tmp_path):
    """File bytes -> parse -> pre-process -> search -> write-back without leaving the device
    (optimize_android_drives) against the reference-shaped facades called one drive at a time."""
    from oracle.make_golden_csv import write_drive
    from vehiclemodelvisualodometry_b200 import (BicycleModel, SearchConfig, load_android_drive,
                                                 optimize_android_drives, optimize_trajectory)
    from vehiclemodelvisualodometry_b200.optimize import DEFAULT_CFG
    from vehiclemodelvisualodometry_b200.trajectory import process_gps_trajectory, process_vo_trajectory

    folders = []
    for ident, n, seed in (("1650000000001", 200, 51), ("1650000000002", 170, 52)):
        write_drive(str(tmp_path / ident), ident, n, seed)
        folders.append(str(tmp_path / ident))
    cfg = SearchConfig(**{**DEFAULT_CFG.__dict__, "grid_v": 8, "grid_s": 8})
    so, traj, plan, ds = optimize_android_drives(folders, cfg)
    traj = traj.cpu().numpy()
    rec = so.records()
    assert ds.f64 and ds.drive_offsets == [0, 200, 370]
    for d, folder in enumerate(folders):
        t, c = load_android_drive(folder)
        out, (so1, plan1, rec1) = optimize_trajectory(process_vo_trajectory(t), process_gps_trajectory(c),
                                                      BicycleModel(), config=cfg, return_details=True)
        a, b = ds.drive_offsets[d], ds.drive_offsets[d + 1]
        w0, w1 = plan.window_offsets[d], plan.window_offsets[d + 1]
        np.testing.assert_array_equal(rec["best_idx"][w0:w1], rec1["best_idx"])
        np.testing.assert_array_equal(rec["best_cost"][w0:w1], rec1["best_cost"])
        nw = w1 - w0
        np.testing.assert_array_equal(traj[2, a:a + nw], np.asarray(out.theta)[:nw])
        np.testing.assert_array_equal(traj[3, a:a + nw], np.asarray(out.velocity)[:nw])
        covered = np.asarray(out.x) != np.asarray(process_vo_trajectory(t).x)
        np.testing.assert_array_equal(traj[0, a:b][covered], np.asarray(out.x)[covered])


# ---- randomised structure: quoted text columns with commas / quotes / newlines around the numbers ----
def _random_csv(rng, n_rows):
    """A file with two numeric columns to read and two free-text columns to skip."""
    words = ["", "abc", "a,b", 'say ""hi""', "line1\nline2", "x\r\ny", " ", "1.5", "NaN", "[[1 2]\n [3 4]]", ",,,", '""']
    nums = _number_strings(rng, 2 * n_rows)
    eol = "\r\n" if rng.random() < 0.3 else "\n"
    lines = ["u,note,v,tag"]
    for i in range(n_rows):
        a, b = nums[2 * i], nums[2 * i + 1]
        if rng.random() < 0.1:
            a = ""
        if rng.random() < 0.1:
            b = ["NA", "nan", "inf", "-inf"][int(rng.integers(4))]
        note, tag = words[int(rng.integers(len(words)))], words[int(rng.integers(len(words)))]
        q = lambda s: '"%s"' % s if (s == "" and rng.random() < 0.5) or any(c in s for c in ',"\n\r') else s  # noqa: E731
        row = [a, q(note), b, q(tag)]
        if rng.random() < 0.05:
            row = row[:3]                       # short row
        lines.append(",".join(row))
        if rng.random() < 0.05:
            lines.append("")                    # blank line
    text = eol.join(lines)
    if rng.random() < 0.7:
        text += eol
    return text.encode()


def test_oracle_equals_pandas_on_random_structure():
    rng = np.random.default_rng(11)
    for _ in range(40):
        data = _random_csv(rng, int(rng.integers(0, 60)))
        want = pd.read_csv(io.BytesIO(data), usecols=["u", "v"], dtype={"u": np.float64, "v": np.float64})
        got = C.read_csv(data, ("u", "v"))
        np.testing.assert_array_equal(got["u"], want["u"].to_numpy())
        np.testing.assert_array_equal(got["v"], want["v"].to_numpy())


@pytest.mark.gpu
def test_gpu_batch_of_random_files_equals_pandas(cuda_device):
    """300 files of random structure in ONE batched parse (sizes from a header alone to several
    4 KiB blocks), each against pandas."""
    from vehiclemodelvisualodometry_b200 import parse_csv_files

    rng = np.random.default_rng(12)
    blobs = [_random_csv(rng, int(rng.integers(0, 400)) if k % 7 else 0) for k in range(300)]
    p = parse_csv_files(blobs, ("u", "v"))
    cols = p.columns.cpu().numpy()
    assert not p.status.any()
    for f, data in enumerate(blobs):
        want = pd.read_csv(io.BytesIO(data), usecols=["u", "v"], dtype={"u": np.float64, "v": np.float64})
        a, b = p.row_offsets[f], p.row_offsets[f + 1]
        assert b - a == len(want), (f, b - a, len(want))
        np.testing.assert_array_equal(cols[0, a:b], want["u"].to_numpy())
        np.testing.assert_array_equal(cols[1, a:b], want["v"].to_numpy())
