#!/usr/bin/env python
"""Throughput of the VMVO window search: bicycle-model hypothesis-steps per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A step is one pass of the hot path over one batch of synthetic drives: window planning, the fused
grid search (with the float64 re-scores of parked windows) and the write-back.

N = 1: the workload is BASELINE.json configs[1] -- one 10 000-frame drive, 32x32 hypothesis grid,
30-step windows (9 940 windows, 3.05e8 hypothesis-steps) -- or, with ``--workload
config3_dense_256x256_w60``, configs[2] (4 096 windows of a 256x256 grid over 60 steps, 1.6e10).

N > 1 (weak scaling): N drives of that shape (seeds base + 0..N-1) are POOLED: every rank holds all
pose streams (they are tiny), the global window list is dealt to the ranks block-cyclically by the
window scheduler (scheduler.py), every rank searches its share, the 64-byte records reach every rank
from inside the search kernel (peer stores over NVLink), and each rank's write-back -- of its share of
the frames -- consumes ALL ranks' records behind the arrival words the kernels exchange
(vmvo_exchange, include/vmvo_b200.h).  So the exchange is on the critical path of the timed step and
nothing else is: no collective, no host barrier.  ``value`` = hypothesis-steps of all N drives / the
slowest rank's time.

Rank 0 prints ONE JSON line (the driver's contract).  Keys beyond the contract: "roofline" (FP32
issue / FMA pipe, with the SFU and HBM figures beside it), "cpu_baseline", "windows_per_s",
"e2e_serial", "config3" (BASELINE configs[2] as a first-class measured workload with its own
roofline / e2e / cpu_baseline; N = 1), "config4" / "config5" (BASELINE configs[3], [4]: batches of
drives dealt over the N ranks, strong scaling against a one-GPU run of the same batch inside the
same process group).

--impl reference times the CPU implementation of the same path on the host cores: the reference
itself is pure Python and does not exist on the GPU box, so this is the oracle port
(oracle/vmvo_oracle.c, OpenMP over windows) on the same config; the numbers of the unmodified Python
reference, measured in the build container by oracle/time_reference_cpu.py, ride along as
``cpu_baseline.reference_python``.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bicycle_model_hypothesis_steps_per_sec"
UNIT = "hypothesis-steps/s"
BASE_SEED = 1658384707877 % (2 ** 32)

# algorithmic work per hypothesis-step (SURVEY.md 8d / DESIGN.md 5): tan, cos, sin = 5 MUFU
# ops (tan = sin + cos + rcp) and 25 FP32 flops
MUFU_PER_HSTEP = 5
FLOP_PER_HSTEP = 25

WORKLOADS = {
    # name: (frames per drive, grid_v, grid_s, window steps)
    "config2_single_drive_10k_32x32_w30": (10000, 32, 32, 30),
    "config3_dense_256x256_w60": (4216, 256, 256, 60),
}
DEAL_BLOCK = 32          # windows per block of the block-cyclic deal (N > 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2_single_drive_10k_32x32_w30", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true",
                    help="skip config 3/4/5, the next-row kernels, the probes and the CPU baseline")
    return ap.parse_args()


def make_cfg(workload):
    from vehiclemodelvisualodometry_b200 import SearchConfig

    n, gv, gs, w = WORKLOADS[workload]
    return n, SearchConfig(grid_v=gv, grid_s=gs, window_frames=w)


def config_of(workload, n_frames, cfg, world=1):
    """The ``config`` object: the same keys in both arms."""
    return {"workload": workload, "frames_per_drive": n_frames, "drives": world,
            "grid": [cfg.grid_v, cfg.grid_s], "window_steps": cfg.window_frames,
            "windows": world * cfg.window_count(n_frames), "base_seed": BASE_SEED,
            "cache": "256 MiB L2 flush between timed steps"}


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return {}


# ---- clocks ----------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax, power = [], set(), None, []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                    power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=smax, samples=len(sm),
                       power_w_max=max(power) if power else None)
        out["reasons"] = sorted(reasons)
        return out


# ---- timing ----------------------------------------------------------------------------------
class Timer:
    """CUDA events on the launching stream around each step, an L2 flush (256 MiB written) before
    each, a barrier + synchronize on both sides, the MAX over ranks of the summed step times."""

    def __init__(self, dev, world):
        import torch

        self.dev, self.world = dev, world
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # 2x the 126 MB L2

    def sync(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def __call__(self, fn, k, w, drain=None, collective=True):
        import torch
        import torch.distributed as dist

        for _ in range(w):
            self.flush.zero_()
            fn()
        if drain:
            drain()
        self.sync() if collective else torch.cuda.synchronize(self.dev)
        evs = []
        for _ in range(k):
            self.flush.zero_()  # evict the pose stream and the records from L2 between steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        if drain:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            drain()          # e.g. the last step's device-to-host copy completes inside the timed total
            b.record()
            evs.append((a, b))
        self.sync() if collective else torch.cuda.synchronize(self.dev)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if self.world > 1 and collective:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def records_equal(a, b):
    """Byte equality of two record buffers except n_rescored (a run-to-run diagnostic)."""
    import torch

    x, y = a.reshape(-1, 64).clone(), b.reshape(-1, 64).clone()
    x[:, 12:16] = 0
    y[:, 12:16] = 0
    return bool(torch.equal(x, y))


# ---- one workload through the pipeline ---------------------------------------------------------
def pooled_drives(workload, world, dev, seed0=BASE_SEED):
    from vehiclemodelvisualodometry_b200 import DriveSet
    from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

    n_frames, cfg = make_cfg(workload)
    t, vo, dt = [], [], []
    for d in range(world):
        b = synthetic_drives(1, n_frames, seed=seed0 + d)
        th, voh, _, _ = b.drive(0)
        t.append(th)
        vo.append(voh)
        dt.append(b.dt)
    return n_frames, cfg, DriveSet.from_arrays(t, dt, vo=vo, device=dev), np.concatenate(t), np.concatenate(vo)


def measure_workload(workload, dev, world, rank, timer, steps, warmup, seed0=BASE_SEED, exhaustive=False):
    """Device-resident step time, search-kernel time, end-to-end (pipelined and serial) for one
    workload; at world > 1 through the window scheduler and the fused exchange."""
    import torch

    from vehiclemodelvisualodometry_b200 import DrivePipeline, DriveStream, _lib, grid_search
    from vehiclemodelvisualodometry_b200.scheduler import PeerGather, shard_range

    n_frames, cfg, drives, time_h, vo_h = pooled_drives(workload, world, dev, seed0)
    F = drives.n_frames
    n_win = world * cfg.window_count(n_frames)
    frame_range = shard_range(F, world, rank) if world > 1 else None
    record_range = shard_range(n_win, world, rank) if world > 1 else None
    sets = []

    def make_pipe(d, b_unused=None):
        if world == 1:
            return DrivePipeline(cfg, d, blend_gps=False)
        g = PeerGather(n_win, dev, block=DEAL_BLOCK)
        sets.append(g)
        return DrivePipeline(cfg, d, blend_gps=False, gather=g, frame_range=frame_range,
                             record_range=record_range)

    # two buffer sets used alternately at N > 1 (what makes the flag exchange race-free)
    pipes = [make_pipe(drives) for _ in range(2 if world > 1 else 1)]
    state = {"n": 0}

    def step():
        p = pipes[state["n"] % len(pipes)]
        state["n"] += 1
        return p.run()

    total_ms = timer(step, steps, warmup)
    ms_per_step = total_ms / steps
    step()
    torch.cuda.synchronize(dev)
    last = pipes[(state["n"] - 1) % len(pipes)]
    rec = last.result_records()
    hsteps = int(cfg.grid_v) * int(cfg.grid_s) * int(rec["n_steps"].astype(np.int64).sum())
    out = {"workload": workload, "cfg": cfg, "n_frames": n_frames, "n_win": n_win, "hsteps": hsteps,
           "ms_per_step": ms_per_step, "value": hsteps / (ms_per_step * 1e-3),
           "rescored_per_window": float(rec["n_rescored"].mean()),
           "launches_per_step": last.kernels_per_pass}

    # the search alone on one GPU over the same windows: kernel time for the roofline, and (N > 1)
    # what every rank's gathered buffer must hold
    alone = torch.empty((n_win, 64), dtype=torch.uint8, device=dev)
    grid_search(cfg, drives, last.plan, out=alone)
    torch.cuda.synchronize(dev)
    only = torch.cuda.CUDAGraph()            # the search launches alone (memset, search, deferred
    with torch.cuda.graph(only):             # re-scores), replayed: no host time between the events
        grid_search(cfg, drives, last.plan, out=alone)
    kk = max(3, min(steps, 50))
    kern_ms = timer(only.replay, kk, 2, collective=False) / kk
    out["kernel_ms_all_windows_one_gpu"] = kern_ms
    if world == 1 and exhaustive:
        # the same step and the same search launches with the pruning votes off (every scan runs to its
        # last step; the library's tuning hook, read when the graphs are captured): identical records
        ctx = _lib.context(dev.index)
        with ctx.tuning(prune=0):
            full_pipe = DrivePipeline(cfg, drives, blend_gps=False)
            full_alone = torch.empty((n_win, 64), dtype=torch.uint8, device=dev)
            grid_search(cfg, drives, last.plan, out=full_alone)
            torch.cuda.synchronize(dev)
            full_only = torch.cuda.CUDAGraph()
            with torch.cuda.graph(full_only):
                grid_search(cfg, drives, last.plan, out=full_alone)
        full_ms = timer(full_pipe.run, steps, warmup, collective=False) / steps
        full_kern = timer(full_only.replay, kk, 2, collective=False) / kk
        torch.cuda.synchronize(dev)
        if not records_equal(alone, full_alone):
            raise SystemExit("the pruned search and the exhaustive search disagree")
        out["exhaustive"] = {"ms_per_step": full_ms, "value": hsteps / (full_ms * 1e-3), "unit": UNIT,
                             "kernel_ms_per_launch": full_kern, "records_identical_to_pruned": True,
                             "note": "every hypothesis rolled to its last step (pruning votes off): the "
                                     "same records bit for bit; `value` above is the default, pruned search"}
        del full_pipe, full_only
    if world > 1:
        # what part of the gap to N x (one GPU) is the workload and what part the exchange: the search
        # launches of (i) the N = 1 workload (drive base + 0 alone) and (ii) this rank's deal of the pool
        # with no record stores and no arrival words, both graph-replayed like `kern_ms`
        from vehiclemodelvisualodometry_b200 import DriveSet, plan_windows

        d0 = DriveSet.from_arrays([time_h[:n_frames]], [float(drives.dt[0].item())], vo=[vo_h[:n_frames]], device=dev)
        p0 = plan_windows(cfg, d0, extents=False)
        o0 = torch.empty((p0.n_windows, 64), dtype=torch.uint8, device=dev)
        ex0 = _lib.Exchange.from_buffer_copy(sets[0].exchange)
        ex0.n_peers = 0                                     # the deal only
        ex0.epoch = None                                    # (and the exchange's step counter stays put)
        share = torch.empty((n_win, 64), dtype=torch.uint8, device=dev)
        graphs = []
        for fn in (lambda: grid_search(cfg, d0, p0, out=o0),
                   lambda: grid_search(cfg, drives, last.plan, out=share, exchange=ex0)):
            fn()
            torch.cuda.synchronize(dev)
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                fn()
            graphs.append(gph)
        n1_ms = timer(graphs[0].replay, kk, 2) / kk
        share_ms = timer(graphs[1].replay, kk, 2) / kk
        out["scaling_breakdown"] = {
            "n1_workload_search_ms": n1_ms, "rank_share_of_pool_search_ms": share_ms,
            "step_ms": ms_per_step,
            "note": "max over ranks; search launches only.  The pool holds drives base + 0 .. N - 1 and the "
                    "N = 1 workload is drive base + 0 alone: (share - n1) is the workload's part of the gap "
                    "to N x (one GPU), (step - share) the exchange, the write-back and the slowest rank"}
        del graphs
        torch.cuda.synchronize(dev)
        if not records_equal(alone, last.records):
            raise SystemExit(f"rank {rank}: the gathered records differ from a one-GPU search of the pool")
        bad = [g.timed_out() for g in sets]
        if any(bad):
            raise SystemExit(f"rank {rank}: an arrival wait timed out ({bad})")
        out["exchange"] = ("fused: block-cyclic deal (block %d), records stored into every peer's gather "
                           "buffer by the search kernel (CUDA IPC, NVLink), arrival words published and "
                           "awaited by the write-back kernel; verified against a one-GPU search of all "
                           "%d windows on every rank" % (DEAL_BLOCK, n_win))
    else:
        out["exchange"] = "none (one GPU)"

    # end to end through the public API with HOST buffers: H2D of the pose streams and stamps,
    # plan + search + write-back, D2H of this rank's records and frames
    vo_pin = torch.from_numpy(np.ascontiguousarray(vo_h)).pin_memory()
    t_pin = torch.from_numpy(np.ascontiguousarray(time_h)).pin_memory()
    inputs = {"vo": vo_pin, "time": t_pin}
    stream = DriveStream(cfg, drives, blend_gps=False, pipe_factory=lambda b, d: make_pipe(d))
    stream.prime(inputs)
    torch.cuda.synchronize(dev)
    e2e_ms = timer(lambda: stream.step(inputs), steps, warmup,
                   drain=lambda: stream.drain() if stream.n > 0 else None) / steps
    torch.cuda.synchronize(dev)
    rlo, rhi = record_range or (0, n_win)
    chk = stream.host[(stream.n - 1) & 1][0].numpy().view(_lib.RESULT_DTYPE).reshape(-1)
    assert np.array_equal(chk["best_idx"], rec["best_idx"][rlo:rhi]), "e2e records differ"
    h2d = int(vo_pin.numel() * 4 + t_pin.numel() * 8)
    d2h = stream.d2h_bytes()
    out["e2e"] = {"value": hsteps / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                  "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "schedule": "streaming API (DriveStream), pipelined across steps on three streams: each "
                              "timed interval = H2D(step s+1) || plan + search + write-back(step s)"
                              + (" with the fused exchange" if world > 1 else "") +
                              " || D2H(step s-1), joined before the interval ends"}

    # the same, one call at a time on one stream: H2D, compute, D2H strictly in sequence (the two
    # buffer sets still alternate: at N > 1 that is what keeps the exchange race-free)
    ser = {"n": 0}

    def serial_step():
        b = ser["n"] & 1
        ser["n"] += 1
        stream._h2d(b, inputs)
        stream.pipes[b].run()
        stream._d2h(b)

    ser_ms = timer(serial_step, steps, warmup) / steps
    out["e2e_serial"] = {"value": hsteps / (ser_ms * 1e-3), "unit": UNIT, "ms_per_step": ser_ms,
                         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "schedule": "one call at a time, one stream: H2D -> plan + search + write-back -> D2H"}
    out["_keep"] = (pipes, stream, sets, only)       # closed by the caller after the last collective
    return out


def close_sets(res):
    pipes, stream, sets, only = res.pop("_keep")
    del pipes, stream, only
    for g in sets:
        g.close()


# ---- roofline --------------------------------------------------------------------------------
def probe_peak(ctx, dev, kind, ops_per_iter):
    """Measured issue peak of one pipe: ops/s over the whole chip."""
    import torch

    from vehiclemodelvisualodometry_b200 import _lib

    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, threads, iters = sms * 8, 256, 4096
    sink = torch.empty(blocks * threads, dtype=torch.float32, device=dev)
    best = None
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.check(ctx.lib.vmvo_peak_probe(ctx.handle, kind, blocks, threads, iters, _lib.ptr(sink),
                                          _lib.stream_ptr(dev)), "vmvo_peak_probe")
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if it > 0:
            best = ms if best is None else min(best, ms)
    return blocks * threads * iters * ops_per_iter / (best * 1e-3)


_PEAKS = {}


def pipe_peaks(ctx, dev):
    """MUFU / FFMA / DFMA lane-operations per second, measured live once per process."""
    if not _PEAKS:
        _PEAKS.update(mufu=probe_peak(ctx, dev, 0, 16), ffma=probe_peak(ctx, dev, 1, 16),
                      dfma=probe_peak(ctx, dev, 2, 8))
    return _PEAKS


def roofline(ctx, dev, res, clocks, extras=True):
    """What binds the search kernel: FP32 issue (the FMA pipe).  ``frac`` = FP32-pipe lane-operations
    the kernel EXECUTES per launch (from the committed ncu instruction counts of this workload,
    profiles/kernel_counts_r02.json; a packed FFMA2 / FADD2 / FMUL2 counts as two) divided by the
    kernel's launch time and by the FFMA lane-operation rate measured live.  Beside it: the hardware
    counters of the same launch under ncu (FMA-pipe active cycles, XU issue share), the algorithmic
    FP32 figure (25 flops per hypothesis-step), and both SFU figures."""
    import torch

    from vehiclemodelvisualodometry_b200.search import executed_mufu_per_hypothesis_step

    cfg, hsteps, kern_ms = res["cfg"], res["hsteps"], res["kernel_ms_all_windows_one_gpu"]
    peaks = load_json("MEASURED_PEAKS.json")
    counts = load_json("profiles", "kernel_counts_r02.json").get(res["workload"], {})
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    nominal_fp32 = sms * 128 * sm_max * 1e6          # FP32 lane-operations per second
    sec = kern_ms * 1e-3
    out = {"bound": "fp32", "kernel": "vmvo_window_search_kernel (+ vmvo_deferred_rescore_kernel)",
           "unit": "G FP32 lane-op/s", "kernel_ms_per_launch": kern_ms,
           "algorithmic_flop_per_hypothesis_step": FLOP_PER_HSTEP,
           "algorithmic_mufu_per_hypothesis_step": MUFU_PER_HSTEP,
           "executed_mufu_per_hypothesis_step": executed_mufu_per_hypothesis_step(cfg)}
    if extras:
        pk = pipe_peaks(ctx, dev)
        ffma, mufu = pk["ffma"], pk["mufu"]
        out["peak_source"] = ("measured live: vmvo_peak_probe FFMA issue rate over the whole chip "
                              "(MEASURED_PEAKS.json holds only HBM and bf16 figures)")
        out["fp32_tflops_measured"] = 2 * ffma / 1e12
        out["fp64_tflops_measured"] = 2 * pk["dfma"] / 1e12
        out["mufu_gops_measured"] = mufu / 1e9
    else:
        ffma, mufu = nominal_fp32, sms * 16 * sm_max * 1e6
        out["peak_source"] = "nominal: SMs x 128 FP32 lanes (16 MUFU lanes) x max SM clock"
    out["peak"] = ffma / 1e9
    out["peak_nominal"] = nominal_fp32 / 1e9
    out["frac_algorithmic_fp32"] = hsteps * FLOP_PER_HSTEP / sec / (2 * ffma)
    out["sfu_frac_algorithmic"] = hsteps * MUFU_PER_HSTEP / sec / mufu
    out["sfu_frac_executed"] = hsteps * out["executed_mufu_per_hypothesis_step"] / sec / mufu
    if counts:
        if counts.get("mufu_per_hypothesis_step"):      # as counted under ncu (pruning included)
            out["executed_mufu_per_hypothesis_step"] = counts["mufu_per_hypothesis_step"]
            out["sfu_frac_executed"] = hsteps * counts["mufu_per_hypothesis_step"] / sec / mufu
        lane_ops = counts["fp32_lane_ops_per_hypothesis_step"] * hsteps
        out["executed_fp32_lane_ops_per_hypothesis_step"] = counts["fp32_lane_ops_per_hypothesis_step"]
        # (per hypothesis-step of the GRID: the pruning votes stop scans early, so the kernel executes
        # this share of them -- and the algorithmic fractions below may exceed what a pipe can do)
        out["executed_share_of_hypothesis_steps"] = counts.get("executed_share_of_hypothesis_steps")
        out["achieved"] = lane_ops / sec / 1e9
        out["frac"] = lane_ops / sec / ffma
        # the hardware's own view of the same launch (captured under ncu, profiles/): share of cycles the
        # FMA pipe is busy, share of issue slots the XU (MUFU) pipe takes, share of cycles an instruction issues
        out["ncu_pipe_fma_cycles_active_pct"] = counts.get("hw_pipe_fma_cycles_active_pct")
        out["ncu_pipe_xu_inst_pct"] = counts.get("hw_pipe_xu_inst_pct")
        out["ncu_issue_active_pct"] = counts.get("hw_issue_active_pct")
        out["counts_source"] = counts.get("source")
    else:
        out["achieved"] = hsteps * FLOP_PER_HSTEP / 2 / sec / 1e9
        out["frac"] = out["frac_algorithmic_fp32"]
        out["counts_source"] = "no ncu counts committed for this workload: frac is the algorithmic figure"
    # HBM side, for the record: one read of the pose stream + one 64-byte record per window
    algo_bytes = res["n_frames"] * 16 * max(1, res["n_win"] // max(1, cfg.window_count(res["n_frames"]))) \
        + res["n_win"] * (64 + 16)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    out["hbm"] = {"algorithmic_bytes_per_launch": algo_bytes, "achieved_gbs": algo_bytes / sec / 1e9,
                  "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                  "frac": algo_bytes / sec / 1e9 / hbm_peak}
    out["traffic"] = counts.get("dram_bytes_per_launch") if counts else \
        load_json("profiles", "traffic.json").get(res["workload"])
    return out


# ---- BASELINE configs[3] / [4]: batches of drives dealt over the ranks ---------------------------
def batch_config(name, cfg, n_drives, n_frames, dev, world, rank, timer, sensors, note):
    """One pass over a batch of drives: plan + search of this rank's share of the global window list
    with the fused exchange + write-back of this rank's frames.  Strong scaling: the batch is fixed,
    and rank 0 also runs the whole batch alone (the one-GPU time of the same batch)."""
    import torch
    import torch.distributed as dist

    from vehiclemodelvisualodometry_b200 import DrivePipeline, DriveSet, grid_search
    from vehiclemodelvisualodometry_b200.scheduler import PeerGather, shard_range
    from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

    b = synthetic_drives(n_drives, n_frames, seed=BASE_SEED + 1000 + n_drives)
    kw = {"vo": list(b.vo)}
    if "gps" in sensors:
        kw["gps"] = list(b.gps)
    if "imu" in sensors:
        kw["imu"] = list(b.imu)
    drives = DriveSet.from_arrays(list(b.time), [b.dt] * n_drives, device=dev, **kw)
    del b
    n_win = n_drives * cfg.window_count(n_frames)
    gather = PeerGather(n_win, dev, block=DEAL_BLOCK) if world > 1 else None
    pipe = DrivePipeline(cfg, drives, blend_gps="gps" in sensors, gather=gather, use_graph=False,
                         frame_range=shard_range(drives.n_frames, world, rank) if world > 1 else None)
    # (the median of three separately timed passes after one warm-up: a single slow pass -- seen once at
    # N = 8, 52 ms instead of 14 -- would otherwise double a two-pass mean)
    timer(pipe.run, 1, 1)
    ms = float(np.median([timer(pipe.run, 1, 0) for _ in range(3)]))
    torch.cuda.synchronize(dev)
    rec = pipe.result_records()
    hsteps = int(cfg.grid_v) * int(cfg.grid_s) * int(rec["n_steps"].astype(np.int64).sum())
    out = {"workload": name, "drives": n_drives, "frames_per_drive": n_frames,
           "grid": [cfg.grid_v, cfg.grid_s], "window_steps": cfg.window_frames, "sensors": sensors,
           "windows": n_win, "hypothesis_steps": hsteps, "ms_per_step": ms,
           "value": hsteps / (ms * 1e-3), "unit": UNIT, "windows_per_s": n_win / (ms * 1e-3),
           "rescored_per_window": float(rec["n_rescored"].mean()), "note": note,
           "timing": "median of three timed passes (max over ranks each)"}
    if world > 1:
        # the same batch on ONE GPU (rank 0 alone; the others wait), and the gathered records checked
        alone = torch.empty((n_win, 64), dtype=torch.uint8, device=dev)
        one_ms = None
        if rank == 0:
            one_ms = timer(lambda: grid_search(cfg, drives, pipe.plan, out=alone), 1, 1, collective=False)
        else:
            grid_search(cfg, drives, pipe.plan, out=alone)
        torch.cuda.synchronize(dev)
        if not records_equal(alone, pipe.records):
            raise SystemExit(f"rank {rank}: {name}: gathered records differ from a one-GPU search")
        if gather.timed_out():
            raise SystemExit(f"rank {rank}: {name}: an arrival wait timed out")
        t = torch.tensor([one_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        one_ms = float(t.item())
        # (the one-GPU figure is the search alone; the N-GPU step also carries plan + write-back)
        out["one_gpu_search_ms"] = one_ms
        out["scaling_efficiency_vs_one_gpu_search"] = one_ms / (world * ms)
        out["records_verified"] = True
        del pipe
        gather.close()
    return out


def config45(dev, world, rank, timer):
    from vehiclemodelvisualodometry_b200 import SearchConfig

    out = {}
    # BASELINE configs[3]: 512 drives x 6 000 frames, W = 60, 32x32 and 128x128
    c32 = SearchConfig(grid_v=32, grid_s=32, window_frames=60)
    out["config4_512_drives_32x32_w60"] = batch_config(
        "config4_512_drives_6000f_32x32_w60", c32, 512, 6000, dev, world, rank, timer, ["vo"],
        "BASELINE configs[3] in full: 512 drives x 6 000 frames (3.0e6 windows)")
    c128 = SearchConfig(grid_v=128, grid_s=128, window_frames=60)
    out["config4_64_drives_128x128_w60"] = batch_config(
        "config4_64_of_512_drives_6000f_128x128_w60", c128, 64, 6000, dev, world, rank, timer, ["vo"],
        "BASELINE configs[3] at 128x128, time-boxed: 64 of the 512 drives (1/8 of the batch; windows are "
        "independent, so the full batch is 8 passes like this one)")
    # BASELINE configs[4]: 1 024 drives x 20 000 frames, 128x128, W = 60, VO + GPS + IMU -- time-boxed
    c5 = SearchConfig(grid_v=128, grid_s=128, window_frames=60, w_vo=1.0, w_gps=0.5, w_imu=40.0)
    r = batch_config("config5_16_of_1024_drives_20000f_128x128_w60_vo_gps_imu", c5, 16, 20000, dev, world,
                     rank, timer, ["vo", "gps", "imu"],
                     "BASELINE configs[4], time-boxed: 16 of the 1 024 drives (1/64 of the batch)")
    r["full_batch_extrapolated_s"] = r["ms_per_step"] * 64 / 1e3
    out["config5_fused_vo_gps_imu"] = r
    return out


# ---- next rows ---------------------------------------------------------------------
def prep_rows(ctx, dev, timer):
    """VO and GPS pre-processing of 64 drives x 10 000 frames, device resident.
    HBM-bound streaming work; algorithmic bytes per frame: VO 96 in (x, y, 3x3 R, stamp) + 40 out,
    GPS 32 in (lat, lon, speed, stamp) + 40 out."""
    import torch

    from vehiclemodelvisualodometry_b200.trajectory import gps_prepare_device, vo_prepare_device

    D, n = 64, 10000
    F = D * n
    g = torch.Generator(device="cpu").manual_seed(7)
    off = torch.arange(D + 1, dtype=torch.int64, device=dev) * n
    x = torch.cumsum(torch.randn(F, generator=g, dtype=torch.float64), 0).to(dev)
    y = torch.cumsum(torch.randn(F, generator=g, dtype=torch.float64), 0).to(dev)
    rot = torch.randn(F, 9, generator=g, dtype=torch.float64).to(dev)
    stamp = (1658384707877 + 50 * torch.arange(F, dtype=torch.float64)).to(dev)
    lat = (12.97 + 2e-6 * torch.arange(F, dtype=torch.float64)).to(dev)
    lat = torch.repeat_interleave(lat[::2], 2)[:F].contiguous()      # 10 Hz fix on a 20 Hz log
    lon = (77.59 + lat - 12.97).contiguous()
    speed = torch.rand(F, generator=g, dtype=torch.float64).to(dev)
    vo_out = torch.empty((5, F), dtype=torch.float64, device=dev)
    gps_out, status, scratch = gps_prepare_device(off, D, F, lat, lon, speed, stamp)
    k = 5
    vo_ms = timer(lambda: vo_prepare_device(off, D, F, x, y, rot, stamp, out=vo_out), k, 1) / k
    gps_ms = timer(lambda: gps_prepare_device(off, D, F, lat, lon, speed, stamp, out=gps_out,
                                              scratch=scratch, status=status), k, 1) / k
    hbm = load_json("MEASURED_PEAKS.json").get("hbm_gbs", 6650.0)
    vo_gbs, gps_gbs = F * 136 / (vo_ms * 1e-3) / 1e9, F * 72 / (gps_ms * 1e-3) / 1e9
    return {"workload": "64 drives x 10000 frames", "vo_ms": vo_ms, "gps_ms": gps_ms,
            "vo_frames_per_s": F / (vo_ms * 1e-3), "gps_frames_per_s": F / (gps_ms * 1e-3),
            "vo_hbm": {"achieved_gbs": vo_gbs, "peak_gbs": hbm, "frac": vo_gbs / hbm},
            "gps_hbm": {"achieved_gbs": gps_gbs, "peak_gbs": hbm, "frac": gps_gbs / hbm},
            "note": "GPS includes the per-drive sequential path sum and de-duplication scan"}


def formats_rows(ctx, dev, timer):
    """the reference's two CSV files of 64 drives x 10 000 rows, bytes resident in HBM
    -> numeric columns (and the 3x3 rot matrices) resident in HBM, through parse_staged (row index,
    field split, number conversion; includes the one host round trip that sizes the outputs).
    HBM-bound in principle: algorithmic bytes = the file bytes once + the columns written.  Beside it
    the reference's own reader (pandas.read_csv + parse_rot, bdd_raw.py:53,155-167) on one host core."""
    import io

    import pandas as pd

    from oracle.csv_oracle import parse_rot          # the CPU leg only: the reference's rot un-stringifier
    from vehiclemodelvisualodometry_b200.dataset import (CACHE_COLUMNS, LOG_COLUMNS, parse_staged,
                                                         stage_csv_files)

    D, n = 64, 10000
    rng = np.random.default_rng(7)
    yaw = np.cumsum(rng.normal(0, 0.01, n))
    rot = np.zeros((n, 3, 3))
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1], rot[:, 2, 2] = (np.cos(yaw), -np.sin(yaw),
                                                                            np.sin(yaw), np.cos(yaw), 1)
    x, y = np.cumsum(rng.normal(0.5, 0.2, n)), np.cumsum(rng.normal(0.1, 0.2, n))
    lat = np.repeat(12.97 + np.cumsum(rng.normal(2e-6, 1e-6, n // 2)), 2)
    lon = np.repeat(77.59 + np.cumsum(rng.normal(3e-6, 1e-6, n // 2)), 2)
    heading, speed = rng.uniform(0, 360, n), np.abs(rng.normal(8, 2, n))
    stamp = 1658384707877 + 50 * np.arange(n)
    b1, b2 = io.StringIO(), io.StringIO()
    pd.DataFrame({"Timestamp": stamp, "Latitude": lat, "Longitude": lon, "heading": heading,
                  "speed": speed}).to_csv(b1, index=False)
    pd.DataFrame({"x": list(x), "y": list(y), "z": list(x * 0), "rot": [r for r in rot]}).to_csv(b2, index=False)
    hbm = load_json("MEASURED_PEAKS.json").get("hbm_gbs", 6650.0)
    out = {"workload": f"{D} drives x {n} rows per file"}
    for name, text, cols, rc, sc in (("log", b1.getvalue(), LOG_COLUMNS, None, "Timestamp"),
                                     ("cache", b2.getvalue(), CACHE_COLUMNS, "rot", None)):
        blob = text.encode()
        st = stage_csv_files([blob] * D, dev)
        k = 5
        ms = timer(lambda: parse_staged(st, cols, rc, sc), k, 2) / k
        written = D * n * 8 * (len(cols) + (9 if rc else 0))
        gbs = (st.n_bytes + written) / (ms * 1e-3) / 1e9
        t0 = time.perf_counter()
        df = pd.read_csv(io.BytesIO(blob))
        if rc:
            df[rc] = df[rc].apply(parse_rot)
        cpu_s = time.perf_counter() - t0
        out[name] = {"file_bytes": st.n_bytes, "ms": ms, "rows_per_s": D * n / (ms * 1e-3),
                     "file_gbs": st.n_bytes / (ms * 1e-3) / 1e9,
                     "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm},
                     "cpu_reference_reader": {"mb_per_s": len(blob) / cpu_s / 1e6, "cores": 1,
                                              "sample": "one drive, pandas.read_csv"
                                                        + (" + parse_rot" if rc else "")}}
    return out


def facade_latency(dev):
    """BicycleModel.run / run_sequence through the reference-shaped facade (host buffers, one launch
    and one synchronisation per call) beside the reference's own 15 us per step (BASELINE.md 2;
    vmvo/bicycle_model.py:40-78)."""
    from vehiclemodelvisualodometry_b200 import BicycleModel, State

    m = BicycleModel(state=State(x=0.0, y=0.0, theta=0.0, velocity=5.0, steering_angle=0.0))
    for _ in range(20):
        m.run(10.0, 5.0, 0.05)
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        m.run(10.0, 5.0, 0.05)
    one = (time.perf_counter() - t0) / n * 1e6
    m.set_state(State(x=0.0, y=0.0, theta=0.0, velocity=5.0, steering_angle=0.0))
    t0 = time.perf_counter()
    for _ in range(20):
        m.set_state(State(x=0.0, y=0.0, theta=0.0, velocity=5.0, steering_angle=0.0))
        m.run_sequence([10.0] * 60, [5.0] * 60, 0.05)
    seq = (time.perf_counter() - t0) / 20 / 60 * 1e6
    return {"bicycle_model_run_us_per_call": one, "run_sequence_60_us_per_step": seq,
            "reference_python_us_per_call": 15.0,
            "note": "the scalar facade pays one H2D + launch + D2H per call; scalar callers are not what the "
                    "GPU path is for (the batched rollout_batch / grid search are)"}


# ---- the B200 arm ---------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from vehiclemodelvisualodometry_b200 import _lib
    from vehiclemodelvisualodometry_b200 import build as vbuild

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0 and vbuild.is_stale():
        vbuild.build_library()
    if world > 1:
        dist.barrier()
    ctx = _lib.context(local_rank)
    timer = Timer(dev, world)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = measure_workload(args.workload, dev, world, rank, timer, args.steps, args.warmup,
                           exhaustive=not args.no_extras)
    clocks = sampler.stop() if rank == 0 else None
    cfg, n_frames = res["cfg"], res["n_frames"]

    extras = {}
    if not args.no_extras:
        if world == 1:
            other = [w for w in sorted(WORKLOADS) if w != args.workload]
            for w in other:                      # the other single-GPU config as a first-class object
                r2 = measure_workload(w, dev, 1, 0, timer, 3 if "config3" in w else 20, 2, exhaustive=True)
                key = "config3" if "config3" in w else "config2"
                obj = {k: r2[k] for k in ("workload", "n_win", "hsteps", "ms_per_step", "value",
                                          "rescored_per_window", "e2e", "e2e_serial", "exhaustive")}
                obj.update(unit=UNIT, windows_per_s=r2["n_win"] / (r2["ms_per_step"] * 1e-3),
                           config=config_of(w, r2["n_frames"], r2["cfg"]),
                           roofline=roofline(ctx, dev, r2, clocks),
                           cpu_baseline=cpu_baseline(w, budget_s=8.0))
                close_sets(r2)
                extras[key] = obj
            extras["prep"] = prep_rows(ctx, dev, timer)
            extras["formats"] = formats_rows(ctx, dev, timer)
            extras["facade"] = facade_latency(dev)
        extras.update(config45(dev, world, rank, timer))

    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 scan + f64 re-score",
        "data": "synthetic (urban stop-and-go drives shaped like BDD sequences, seeds base + drive)",
        "config": config_of(args.workload, n_frames, cfg, world),
        "windows_per_s": res["n_win"] / (res["ms_per_step"] * 1e-3),
        "exchange": res["exchange"],
        "gpu_launches": int(res["launches_per_step"] * args.steps),
        "e2e": res["e2e"], "e2e_serial": res["e2e_serial"],
        "clocks": clocks,
        "rescored_per_window": res["rescored_per_window"],
        **({"scaling_breakdown": res["scaling_breakdown"]} if "scaling_breakdown" in res else {}),
        "search": "exact branch and bound: a scanning warp stops once every hypothesis it holds has passed "
                  "the candidate threshold of the bound its team held when the pass began; `value` counts "
                  "the grid's hypothesis-steps (G_v x G_s x N per window, SURVEY 8d), records are those of "
                  "the exhaustive scan bit for bit (`exhaustive`: the same step with the votes off)",
    }
    if "exhaustive" in res:
        line["exhaustive"] = res["exhaustive"]
    if rank == 0:
        line["roofline"] = roofline(ctx, dev, res, clocks, extras=not args.no_extras)
        line.update(extras)
        if world == 1 and not args.no_extras:
            line["cpu_baseline"] = cpu_baseline(args.workload, budget_s=12.0)
            try:
                pk = pipe_peaks(ctx, dev)
                line["peaks_probe"] = {"mufu_gops": pk["mufu"] / 1e9, "ffma_glaneops": pk["ffma"] / 1e9,
                                       "dfma_glaneops": pk["dfma"] / 1e9}
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    close_sets(res)
    if world > 1:
        dist.destroy_process_group()


# ---- CPU legs -------------------------------------------------------------------------------------
def cpu_pass(workload, seed, max_windows=None, threads=0):
    """One pass of the oracle port over (a sample of) the workload; returns (hyp-steps, seconds)."""
    from oracle import c_oracle
    from oracle import vmvo_oracle as O
    from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

    n_frames, cfg = make_cfg(workload)
    batch = synthetic_drives(1, n_frames, seed=seed)
    t, vo, _, _ = batch.drive(0)
    spec = O.SearchSpec(grid_v=cfg.grid_v, grid_s=cfg.grid_s, window_frames=cfg.window_frames)
    starts, lens = O.window_extents(spec, t)
    if max_windows is not None and len(starts) > max_windows:
        sel = np.linspace(0, len(starts) - 1, max_windows).astype(np.int64)
        starts, lens = starts[sel], lens[sel]
    t0 = time.perf_counter()
    rec, steps = c_oracle.search(cfg.to_c(), starts, lens, np.zeros(len(starts), np.int32), [batch.dt],
                                 vo, n_threads=threads)
    return steps, time.perf_counter() - t0, len(starts)


def host_threads():
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1; ignore that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def reference_python():
    """The unmodified Python reference, timed in the BUILD CONTAINER by oracle/time_reference_cpu.py
    (it cannot travel to the GPU box); the committed numbers ride along, labelled."""
    d = load_json("profiles", "reference_cpu_r02.json")
    if not d:
        return None
    g, s = d.get("grid_through_bicycle_model_run", {}), d.get("slsqp_as_is", {})
    return {"where": d.get("where"), "cores": d.get("host_cores"), "drive": d.get("drive"),
            "grid_through_BicycleModel_run_hypothesis_steps_per_s": g.get("hypothesis_steps_per_s"),
            "grid_through_BicycleModel_run_hypothesis_steps_per_s_per_core": g.get("hypothesis_steps_per_s_per_core"),
            "slsqp_as_is_windows_per_s": s.get("windows_per_s"),
            "slsqp_as_is_model_steps_per_s": s.get("model_steps_per_s"),
            "source": "profiles/reference_cpu_r02.json (oracle/time_reference_cpu.py)"}


def cpu_baseline(workload, budget_s=12.0):
    cores = host_threads()
    n_frames, cfg = make_cfg(workload)
    total_win = n_frames - 2 * cfg.window_frames
    per_win = cfg.grid_v * cfg.grid_s * cfg.window_frames
    s0, t0, _ = cpu_pass(workload, BASE_SEED, max_windows=64, threads=cores)   # warms library and threads
    # a pass sized to ~1/3 of the budget, at least 64 windows, at most the whole workload
    want = int(max(64, min(total_win, (s0 / t0) * (budget_s / 3) / per_win)))
    steps = sec = 0.0
    passes = nwin = 0
    t_end = time.perf_counter() + budget_s
    while passes < 2 or time.perf_counter() < t_end:
        s1, t1, nwin = cpu_pass(workload, BASE_SEED, max_windows=want, threads=cores)
        steps += s1
        sec += t1
        passes += 1
        if passes >= 200:
            break
    out = {"value": steps / sec, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{passes} passes over {nwin} of {total_win} windows of {workload}, evenly spaced; "
                     f"oracle/vmvo_oracle.c (float64, OpenMP over windows), {sec:.1f} s of search time"}
    rp = reference_python()
    if rp:
        out["reference_python"] = rp
    return out


def run_reference(args):
    """The CPU implementation of the path on all host cores (oracle port; rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    n_frames, cfg = make_cfg(args.workload)
    total_win = n_frames - 2 * cfg.window_frames
    per_win = cfg.grid_v * cfg.grid_s * cfg.window_frames
    # bounded sample per step: size it from a calibration pass so the whole run takes ~1 min
    steps0, sec0, _ = cpu_pass(args.workload, BASE_SEED, max_windows=128, threads=cores)
    budget = 60.0 / max(1, args.steps + args.warmup)
    want = int(max(64, min(total_win, (steps0 / sec0) * budget / per_win)))
    for _ in range(args.warmup):
        cpu_pass(args.workload, BASE_SEED, max_windows=want, threads=cores)
    tot_steps, tot_sec, nwin = 0, 0.0, 0
    for _ in range(args.steps):
        s, sec, nwin = cpu_pass(args.workload, BASE_SEED, max_windows=want, threads=cores)
        tot_steps += s
        tot_sec += sec
    value = tot_steps / tot_sec
    sample = (f"{nwin} of {total_win} windows per step, evenly spaced; oracle/vmvo_oracle.c "
              f"(float64, OpenMP over windows)")
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    rp = reference_python()
    if rp:
        cb["reference_python"] = rp
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (same generator and seeds as the b200 arm)",
        "config": config_of(args.workload, n_frames, cfg, max(1, args.gpus)),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "this arm is the repo's own C/OpenMP restatement of the path (kind = port) on all host "
                "threads, NOT the Python reference: the reference is pure Python (~15 us per model step), "
                "does not exist on the GPU box, and its own numbers (build container) are under "
                "cpu_baseline.reference_python",
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
