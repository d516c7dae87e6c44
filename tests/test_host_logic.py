"""CPU: host-side logic -- config mapping, the C-ABI library's exports, the window scheduler,
and the world-size-2 gather over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import vmvo_oracle as O
from tests.helpers import spec_of
from vehiclemodelvisualodometry_b200 import SearchConfig, _lib, scheduler
from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_library):
    header = open(os.path.join(ROOT, "include", "vmvo_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void|const char\*)\s+(vmvo_[a-z0-9_]+)\(", header, re.M))
    assert len(declared) >= 17
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(built_library)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().vmvo_abi_version() == 2


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.SearchCfg) == 10 * 4 + 10 * 8
    assert _lib.RESULT_DTYPE.itemsize == 64
    assert _lib.RESULT_DTYPE.fields["best_cost"][1] == 16 and _lib.RESULT_DTYPE.fields["theta1"][1] == 56


def test_cfg_defaults_and_window_count(built_library):
    c = _lib.default_cfg()
    assert (c.grid_v, c.grid_s, c.window_frames, c.horizon_time) == (32, 32, 30, 3.0)
    assert (c.wheel_base, c.steering_ratio, c.max_steer, c.max_accel, c.max_steer_rate) == (
        O.WHEEL_BASE, O.STEERING_RATIO, O.MAX_STEER, float(O.MAX_ACCEL), O.MAX_STEER_RATE)
    assert _lib.window_count(c, 10000) == 9940          # BASELINE config 2
    assert _lib.window_count(c, 60) == 0 and _lib.window_count(c, 5) == 0
    cfg = SearchConfig(window_mode="time", horizon_frames=59)
    assert cfg.window_count(260) == 142 and _lib.window_count(cfg.to_c(), 260) == 142


def test_search_config_maps_to_oracle_spec_and_c_struct():
    cfg = SearchConfig(grid_v=7, grid_s=9, target_mode="traverse", primary="gps", w_vo=0.0, w_gps=2.0,
                       k_steer=1e-6, seed_mode="given")
    spec = spec_of(cfg)
    assert (spec.grid_v, spec.target_mode, spec.primary, spec.w_gps) == (7, "traverse", "gps", 2.0)
    c = cfg.to_c()
    assert (c.grid_v, c.grid_s, c.target_mode, c.primary, c.seed_mode) == (7, 9, 1, 1, 1)
    assert c.max_window_poses == 31 and c.w_gps == 2.0 and c.k_steer == 1e-6
    with pytest.raises(ValueError):
        SearchConfig(target_mode="nope").to_c()


def test_no_gpu_means_loud_failure(built_library):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vehiclemodelvisualodometry_b200 import BicycleModel
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BicycleModel().run(1.0, 0.0, 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.context()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 9940, 3010560):
        for world in (1, 2, 3, 4, 8):
            spans = [scheduler.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= scheduler.shard_capacity(n, world)


def test_assign_drives_balances():
    counts = [5880] * 13 + [100, 19940, 7]
    plan = scheduler.assign_drives(counts, 4)
    assert sorted(d for r in plan for d in r) == list(range(len(counts)))
    loads = [sum(counts[d] for d in r) for r in plan]
    assert max(loads) - min(loads) <= max(counts)


def test_synthetic_drive_shape():
    b = synthetic_drives(3, 500, seed=5)
    assert b.vo.shape == (3, 500, 4) and b.vo.dtype == np.float32 and b.time.dtype == np.float64
    assert np.all(np.diff(b.time, axis=1) > 0) and abs(np.mean(np.diff(b.time[0])) - 0.05) < 1e-6
    assert np.all(b.gt[..., 3] >= 0) and np.all(b.gt[..., 3] <= 15)
    b2 = synthetic_drives(3, 500, seed=5)
    np.testing.assert_array_equal(b.vo, b2.vo)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from vehiclemodelvisualodometry_b200 import scheduler, _lib
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
n = 37
want = np.zeros(n, dtype=_lib.RESULT_DTYPE)
want["best_idx"] = np.arange(n) * 3
want["best_cost"] = np.arange(n) * 0.5
def fake_search(rng, out):      # stands in for grid_search(..., window_range=rng, out=out)
    lo, hi = rng
    out.copy_(torch.from_numpy(want[lo:hi].view(np.uint8).reshape(hi - lo, 64)))
got = scheduler.search_sharded(fake_search, n, torch.device("cpu"))
rec = got.numpy().view(_lib.RESULT_DTYPE).reshape(-1)
assert rec.shape == (n,) and np.array_equal(rec["best_idx"], want["best_idx"])
assert np.array_equal(rec["best_cost"], want["best_cost"])
# the block-cyclic deal with the collective form of the exchange (gather_dealt)
def fake_dealt(mine, buf):      # stands in for grid_search(..., out=buf, exchange=...)
    buf[torch.from_numpy(mine)] = torch.from_numpy(want[mine].view(np.uint8).reshape(len(mine), 64))
for block in (1, 4, 16, 64):
    got = scheduler.search_dealt(fake_dealt, n, torch.device("cpu"), block=block)
    rec = got.numpy().view(_lib.RESULT_DTYPE).reshape(-1)
    assert np.array_equal(rec["best_idx"], want["best_idx"]) and np.array_equal(rec["best_cost"], want["best_cost"])
dist.destroy_process_group()
print("rank", sys.argv[1], "ok")
"""


def test_sharded_search_gathers_over_gloo_world2(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_schema_types_behave_like_the_reference():
    """Trajectory.__len__ / __getitem__ / to_numpy and states_list_to_trajectory of the PRODUCT
    (vmvo/schema.py:30-57,130-147): against the stamps frozen from the reference (golden B5_times)
    and, where the reference tree is present, against the reference's own classes."""
    from oracle import ref_bridge
    from tests.helpers import load_golden, unhex
    from vehiclemodelvisualodometry_b200 import State, Trajectory, states_list_to_trajectory

    g = load_golden()
    states = [State(x=i, y=-0.5 * i, theta=0.1 * i, velocity=1 + i, steering_angle=3 * i) for i in range(7)]
    tt = states_list_to_trajectory(states, 12.5, 0.05)
    assert tt.time == unhex(g["schema"]["B5_times"]).tolist()          # stamps start AT start_time (quirk D5)
    assert tt.x == [float(i) for i in range(7)] and tt.velocity == [1.0 + i for i in range(7)]
    assert isinstance(tt.x, list) and len(tt) == 7
    assert states_list_to_trajectory([], 3.0, 0.1).time == []
    a = tt.to_numpy()
    assert a.shape == (7, 5) and a.dtype == np.float64
    for c, name in enumerate(("x", "y", "theta", "velocity", "time")):
        assert a[:, c].tolist() == getattr(tt, name)
    assert tt[2] == (2.0, -1.0, 0.2, 3.0, tt.time[2])
    assert tt[-1][0] == 6.0
    sl = tt[1:4]
    assert sl[0] == [1.0, 2.0, 3.0] and len(sl) == 5 and sl[4] == tt.time[1:4]
    assert repr(tt) == str(tt) == "Trajectory(len=7)"
    short = Trajectory(x=[0, 1, 2], y=[0, 0, 0], theta=[0, 0], velocity=[1, 1, 1], time=[0, 1, 2])
    assert len(short) == 3                                             # len is the length of x (GPS theta is one short)
    with pytest.raises(ValueError):
        short.to_numpy()                                               # ragged columns: NumPy refuses, as in the reference
    if ref_bridge.available():
        ref = ref_bridge.load()
        rstates = [ref.schema.State(**s.model_dump()) for s in states]
        rt = ref.schema.states_list_to_trajectory(rstates, 12.5, 0.05)
        assert dict(rt) == dict(tt)
        assert np.array_equal(rt.to_numpy(), a) and rt[2] == tt[2] and rt[1:4] == tt[1:4]
        assert len(rt) == len(tt) and repr(rt) == repr(tt)
        rshort = ref.schema.Trajectory(**dict(short))
        assert len(rshort) == 3
        with pytest.raises(ValueError):
            rshort.to_numpy()


def test_block_cyclic_deal_matches_the_kernel_queue():
    """scheduler.deal_* against the index arithmetic of the search kernel's queue
    (csrc/vmvo_search.cu next_window): item r of rank q is window ((r >> s) * world + q) << s | low
    bits, items run while the window exists; every window is dealt exactly once."""
    from vehiclemodelvisualodometry_b200 import scheduler

    for n in (0, 1, 31, 32, 33, 1000, 9940, 79520):
        for world in (1, 2, 3, 4, 8):
            for block in (1, 2, 32, 64):
                sh = block.bit_length() - 1
                seen = np.zeros(n, dtype=np.int64)
                for rank in range(world):
                    nb = -(-n // block)
                    mine = (nb - rank + world - 1) // world if nb > rank else 0
                    got = []
                    for r in range(mine * block):             # the host's n_local
                        b = r >> sh
                        w = ((b * world + rank) << sh) + (r - (b << sh))
                        if w >= n:
                            break
                        got.append(w)
                    want = scheduler.deal_indices(n, block, world, rank)
                    assert np.array_equal(got, want), (n, world, block, rank)
                    assert len(want) == scheduler.deal_count(n, block, world, rank)
                    assert np.all(scheduler.deal_owner(want, block, world) == rank)
                    seen[want] += 1
                assert np.all(seen == 1)
    assert [scheduler.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def _sklansky(values, steps):
    """The association of warp_scan_add (csrc/vmvo_device.cuh), restated lane by lane in float64:
    in step s the lanes whose bit s is set add the value of the last lane of the aligned block of
    2^s lanes to their left."""
    v = np.array(values, dtype=np.float64)
    for s in range(steps):
        prev = v.copy()
        for lane in range(len(v)):
            if (lane >> s) & 1:
                v[lane] = prev[lane] + prev[((lane >> s) << s) - 1]
    return v


def test_sklansky_scan_properties_the_packed_rescore_relies_on():
    """(1) it is an inclusive prefix sum; (2) inputs that are zero from lane m on leave the SAME
    bits in every lane >= m - 1; (3) a group of 8 (16) lanes performs exactly the whole-warp scan's
    operations on lanes 0..7 (0..15); (4) negated inputs give exactly negated outputs."""
    rng = np.random.default_rng(5)
    for trial in range(200):
        a = rng.normal(0, 1, 32) * 10.0 ** rng.integers(-6, 3)
        full = _sklansky(a, 5)
        np.testing.assert_allclose(full, np.cumsum(a), rtol=0, atol=1e-14 * np.abs(a).sum())
        assert np.array_equal(_sklansky(-a, 5), -full)
        for m in (1, 2, 3, 5, 8, 11, 16):
            z = a.copy()
            z[m:] = 0.0
            out = _sklansky(z, 5)
            assert np.all(out[m - 1:] == out[m - 1]), (trial, m)
            for lg in (3, 4):
                g = 1 << lg
                if m <= g:
                    assert np.array_equal(_sklansky(z[:g], lg), out[:g]), (trial, m, lg)
    # ... which the Hillis-Steele form does not have: (a0 + a1) + a2 at lane 2, but a later lane
    # may see a0 + (a1 + a2)
    def hillis_steele(values):
        v = np.array(values, dtype=np.float64)
        o = 1
        while o < len(v):
            prev = v.copy()
            for lane in range(o, len(v)):
                v[lane] = prev[lane] + prev[lane - o]
            o <<= 1
        return v
    differs = 0
    for trial in range(200):
        z = np.zeros(32)
        z[:3] = rng.normal(0, 1, 3)
        out = hillis_steele(z)
        differs += int(not np.all(out[2:] == out[2]))
    assert differs > 0


def test_pass_and_chunk_orders_are_permutations():
    """The orders the search kernel walks its passes in (vmvo_search_kernels.cuh: many-pass kernels start
    at an estimated pass and alternate outward while either side lasts; two-pass kernels deal the
    acceleration chunks middle-out) restated: every pass / chunk exactly once, nearest first.  Any order
    gives the same records (the pruning votes are exact); this pins that none is skipped."""
    def pass_order(n_pass, p_est):
        p_lo, p_hi, up, out = p_est - 1, p_est, True, []
        for _ in range(n_pass):
            take_hi = (up and p_hi < n_pass) or p_lo < 0
            if take_hi:
                out.append(p_hi)
                p_hi += 1
            else:
                out.append(p_lo)
                p_lo -= 1
            up = not up
        return out

    for n_pass in (1, 2, 3, 7, 16, 32, 33):
        for p_est in range(n_pass):
            order = pass_order(n_pass, p_est)
            assert sorted(order) == list(range(n_pass)), (n_pass, p_est, order)
            assert order[0] == p_est
            dist = [abs(q - p_est) for q in order]
            # never further out than one step beyond what the other side has reached
            assert all(dist[i + 1] >= dist[i] - 1 for i in range(n_pass - 1))

    def chunk_of(ic, n_ic):
        mid = (n_ic - 1) >> 1
        return mid + ((ic + 1) >> 1) if ic & 1 else mid - (ic >> 1)

    for n_ic in (1, 2, 3, 4, 5, 8):
        assert sorted(chunk_of(c, n_ic) for c in range(n_ic)) == list(range(n_ic))
    assert [chunk_of(c, 4) for c in range(4)] == [1, 2, 0, 3]      # 32 accelerations: the middle 16 first
