#!/usr/bin/env python
"""Throughput of the VMVO window search: bicycle-model hypothesis-steps per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one pass of the hot path over one batch of synthetic drives: window planning,
the fused grid search and the write-back (plus, at N > 1, the all-gather of the 64-byte
window records over NCCL).  At N = 1 the workload is BASELINE.json configs[1]: one
10 000-frame drive, 32x32 hypothesis grid, 30-step windows (9 940 windows, 3.05e8
hypothesis-steps).  At N > 1 every rank searches its own drive of that shape (weak scaling).

Rank 0 prints ONE JSON line (see the driver's contract in the task statement).  Keys beyond
the contract: "roofline" (SFU-issue bound, with the HBM and FP32 figures beside it),
"cpu_baseline" (the C/OpenMP oracle port on all host cores), "windows_per_s", and
"dense_grid" (BASELINE configs[2], 256x256 x 60 steps: the compute-bound stress the
roofline fraction is normally quoted on).

--impl reference times the CPU implementation of the same path on the host cores: the
reference itself is pure Python and does not travel to the GPU box, so this is the oracle
port (oracle/vmvo_oracle.c, OpenMP over windows) on the same config.
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
# (what the kernel executes per hypothesis-step depends on the scan it selects:
# search.executed_mufu_per_hypothesis_step)

WORKLOADS = {
    # name: (frames per drive, grid_v, grid_s, window steps)
    "config2_single_drive_10k_32x32_w30": (10000, 32, 32, 30),
    "config3_dense_256x256_w60": (4216, 256, 256, 60),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2_single_drive_10k_32x32_w30", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="skip dense-grid, probes and CPU baseline")
    return ap.parse_args()


def make_cfg(workload):
    from vehiclemodelvisualodometry_b200 import SearchConfig

    n, gv, gs, w = WORKLOADS[workload]
    return n, SearchConfig(grid_v=gv, grid_s=gs, window_frames=w)


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


# ---- the B200 arm --------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from vehiclemodelvisualodometry_b200 import (DrivePipeline, DriveSet, _lib, grid_search, plan_windows)
    from vehiclemodelvisualodometry_b200 import build as vbuild
    from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

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

    n_frames, cfg = make_cfg(args.workload)
    batch = synthetic_drives(1, n_frames, seed=BASE_SEED + rank)
    time_h, vo_h, gps_h, imu_h = batch.drive(0)
    drives = DriveSet.from_arrays([time_h], [batch.dt], vo=[vo_h], device=dev)
    plan = plan_windows(cfg, drives)
    n_win = plan.n_windows
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 2x the 126 MB L2
    # The only exchange of the path is the per-window records.  Preferred: fused into the search --
    # every rank's gather buffer is mapped into its peers (CUDA IPC) and the kernel's epilogue
    # stores each record into all of them over NVLink (scheduler.PeerGather); no second kernel has
    # to squeeze in beside the persistent search, which owns every SM.  Fallback: an NCCL
    # all-gather of the record buffer, overlapped with the next pass' search.  Two buffer sets
    # either way, so that step s + 1 never overwrites records of step s in flight.
    n_buf = 2 if world > 1 else 1
    peer = None
    gather_kind = "none (one GPU)"
    if world > 1 and not os.environ.get("VMVO_BENCH_GATHER", "").lower().startswith("nccl"):
        try:
            from vehiclemodelvisualodometry_b200.scheduler import PeerGather

            peer = [PeerGather(n_win, dev) for _ in range(n_buf)]
            gather_kind = "fused: records stored into every peer's buffer by the search kernel (CUDA IPC, NVLink)"
        except Exception as exc:       # no peer access on this box: use the collective
            peer = None
            print(f"[bench] peer buffers unavailable ({exc}); using the NCCL all-gather", file=sys.stderr)
    ok = torch.tensor([1 if (peer is not None or world == 1) else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks take the same path
        if int(ok.item()) == 0 and peer is not None:
            for pg in peer:
                pg.close()
            peer = None
    if peer is not None:
        gathered = [pg.buffer for pg in peer]
        local = [pg.local for pg in peer]
        pipes = []
        for b in range(n_buf):
            peer[b].enable()           # the graph captures the mirrors in force
            pipes.append(DrivePipeline(cfg, drives, blend_gps=False, records=local[b]))
            peer[b].disable()
    else:
        if world > 1:
            gather_kind = "NCCL all_gather_into_tensor, overlapped with the next search"
        gathered = [torch.empty((world * n_win, 64), dtype=torch.uint8, device=dev) for _ in range(n_buf)]
        local = [g[rank * n_win:(rank + 1) * n_win] for g in gathered]
        # the public batched API: plan + search + write-back captured once as CUDA graphs
        pipes = [DrivePipeline(cfg, drives, blend_gps=False, records=local[b], split=world > 1)
                 for b in range(n_buf)]
    pipe = pipes[0]
    state = {"n": 0, "work": None}

    def step():
        if world == 1:
            return pipe.run()
        b = state["n"] & 1
        state["n"] += 1
        if peer is not None:
            return pipes[b].run()          # the records reach every rank from inside the search
        pipes[b].run_search()
        if state["work"] is not None:      # the previous pass's gather had this search to hide behind
            state["work"].wait()
        # the only exchange of the path: per-window records to every rank, on NCCL's stream,
        # while the write-back (which needs the local records only) and the next search proceed
        state["work"] = dist.all_gather_into_tensor(gathered[b], local[b], async_op=True)
        traj = pipes[b].run_write_back()
        return local[b], traj

    def drain():
        if state["work"] is not None:
            state["work"].wait()
            state["work"] = None

    def timed(fn, k, w, drain=drain):
        for _ in range(w):
            flush.zero_()
            fn()
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(k):
            flush.zero_()  # evict the pose stream and the records from L2 between steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        drain()          # the last gather completes inside the timed total
        b.record()
        evs.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(step, args.steps, args.warmup)
    # graph replays do not pass through the library's launch counter: count the captured kernels
    launches = DrivePipeline.KERNELS_PER_PASS * args.steps
    step()
    drain()
    torch.cuda.synchronize()
    rec = pipe.result_records()
    if world > 1:
        # what arrived: every rank's slot of this rank's gather buffer against that rank's own records
        torch.cuda.synchronize()
        dist.barrier()
        last = (state["n"] - 1) & 1
        check = torch.empty((world * n_win, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(check, local[last].contiguous())
        torch.cuda.synchronize()
        if not torch.equal(check, gathered[last]):
            raise SystemExit(f"rank {rank}: gathered records differ from the ranks' own ({gather_kind})")
    hsteps = int(cfg.grid_v) * int(cfg.grid_s) * int(rec["n_steps"].astype(np.int64).sum())
    ms_per_step = total_ms / args.steps
    value = hsteps * world / (ms_per_step * 1e-3)

    # search kernel alone (the dominant kernel): average launch duration for the roofline
    kern_ms = timed(lambda: grid_search(cfg, drives, plan, out=local[0]), args.steps, 2) / args.steps

    # end to end through the public API with HOST buffers: H2D of the pose stream and stamps,
    # plan + search + write-back, D2H of the records and the written-back trajectory
    vo_pin = torch.from_numpy(np.ascontiguousarray(vo_h)).pin_memory()
    t_pin = torch.from_numpy(np.ascontiguousarray(time_h)).pin_memory()
    rec_pin = torch.empty((n_win, 64), dtype=torch.uint8).pin_memory()
    traj_pin = torch.empty((4, n_frames), dtype=torch.float64).pin_memory()

    def e2e_step():
        drives.vo.copy_(vo_pin, non_blocking=True)
        drives.time.copy_(t_pin, non_blocking=True)
        records, traj = step()
        rec_pin.copy_(records, non_blocking=True)
        traj_pin.copy_(traj, non_blocking=True)

    e2e_note = "serial per step: H2D, plan + search + write-back (+ gather), D2H on one stream"
    stream = None
    if world == 1 or peer is not None:
        # software-pipelined across steps, as a caller streaming drives through the API would run
        # it: two resident drive buffers; the timed interval of step s carries the H2D of step
        # s + 1's inputs, the compute of step s and the D2H of step s - 1's results, on three
        # streams that all start at the interval's first event and are joined before its last one
        # (so every interval pays for one full H2D and one full D2H; the L2 flush between
        # intervals runs with all streams idle).  The last step's D2H is timed by the drain.
        from vehiclemodelvisualodometry_b200 import DriveStream

        factory = None
        if peer is not None:
            # N > 1: the stream's two pipelines write their records into the two gather buffers, with
            # the result mirrors in force while their graphs are captured (the fused gather)
            def factory(b, d):
                peer[b].enable()
                try:
                    return DrivePipeline(cfg, d, blend_gps=False, records=local[b])
                finally:
                    peer[b].disable()

        stream = DriveStream(cfg, drives, blend_gps=False, pipe_factory=factory)   # the public streaming API
        inputs = {"vo": vo_pin, "time": t_pin}
        stream.prime(inputs)                                     # inputs of the very first step
        torch.cuda.synchronize(dev)

        def e2e_step():  # noqa: F811
            stream.step(inputs)

        def e2e_drain():
            if stream.n > 0:
                stream.drain()
            drain()

        e2e_ms = timed(e2e_step, args.steps, args.warmup, drain=e2e_drain) / args.steps
        e2e_note = ("pipelined across steps on three streams: each timed interval = H2D(step s+1) || "
                    "plan + search + write-back(step s)" + (" with the fused gather" if peer is not None else "") +
                    " || D2H(step s-1), joined before the interval ends")
        torch.cuda.synchronize(dev)
        chk = stream.host[(stream.n - 1) & 1][0].numpy().view(_lib.RESULT_DTYPE).reshape(-1)
        assert np.array_equal(chk["best_idx"], pipe.result_records()["best_idx"]), "e2e records differ"
    else:
        e2e_ms = timed(e2e_step, args.steps, args.warmup) / args.steps
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras["dense_grid"] = dense_grid(ctx, dev, timed, args)
        extras["prep"] = prep_rows(ctx, dev, timed)
        extras["formats"] = formats_rows(ctx, dev, timed)
    clocks = sampler.stop() if rank == 0 else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 scan + f64 re-score",
        "data": "synthetic (urban stop-and-go drives shaped like BDD sequences, seeds base+rank)",
        "config": {"workload": args.workload, "frames_per_drive": n_frames, "drives_per_gpu": 1,
                   "grid": [cfg.grid_v, cfg.grid_s], "window_steps": cfg.window_frames,
                   "windows_per_gpu": n_win, "hypothesis_steps_per_gpu": hsteps,
                   "cache": "256 MiB L2 flush between timed steps", "base_seed": BASE_SEED},
        "windows_per_s": n_win * world / (ms_per_step * 1e-3),
        "gather": gather_kind,
        "gpu_launches": int(launches),
        "e2e": {"value": hsteps * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(vo_pin.numel() * 4 + t_pin.numel() * 8),
                "d2h_bytes_per_step": int(rec_pin.numel() + traj_pin.numel() * 8),
                "schedule": e2e_note},
        "clocks": clocks,
        "rescored_per_window": float(rec["n_rescored"].mean()),
    }

    if rank == 0:
        line["roofline"] = roofline(ctx, dev, hsteps, kern_ms, n_frames, n_win, clocks, args, cfg)
        line.update(extras)
        if world == 1 and not args.no_extras:
            line["cpu_baseline"] = cpu_baseline(args.workload, budget_s=12.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if peer is not None:
            del pipes, pipe, stream
            for pg in peer:
                pg.close()
        dist.destroy_process_group()


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


def roofline(ctx, dev, hsteps, kern_ms, n_frames, n_win, clocks, args, cfg):
    import torch

    from vehiclemodelvisualodometry_b200.search import executed_mufu_per_hypothesis_step

    EXEC_MUFU_PER_HSTEP = executed_mufu_per_hypothesis_step(cfg)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    nominal_mufu = sms * 16 * sm_max * 1e6
    out = {"bound": "sfu", "kernel": "vmvo_window_search_kernel", "unit": "GMUFU-op/s",
           "kernel_ms_per_launch": kern_ms,
           "algorithmic_mufu_per_hypothesis_step": MUFU_PER_HSTEP,
           "executed_mufu_per_hypothesis_step": EXEC_MUFU_PER_HSTEP}
    achieved = hsteps * MUFU_PER_HSTEP / (kern_ms * 1e-3)
    out["achieved"] = achieved / 1e9
    out["peak_nominal"] = nominal_mufu / 1e9
    if not args.no_extras:
        mufu = probe_peak(ctx, dev, 0, 16)
        ffma = probe_peak(ctx, dev, 1, 16)
        dfma = probe_peak(ctx, dev, 2, 8)
        out["peak"] = mufu / 1e9
        out["peak_source"] = "measured live: vmvo_peak_probe MUFU.SIN/COS issue rate, whole chip"
        out["fp32_tflops_measured"] = 2 * ffma / 1e12
        out["fp64_tflops_measured"] = 2 * dfma / 1e12
        out["fp32_frac_algorithmic"] = hsteps * FLOP_PER_HSTEP / (kern_ms * 1e-3) / (2 * ffma)
    else:
        out["peak"] = nominal_mufu / 1e9
        out["peak_source"] = "nominal: SMs x 16 MUFU lanes x max SM clock"
    out["frac"] = out["achieved"] / out["peak"]
    out["frac_executed"] = out["frac"] * EXEC_MUFU_PER_HSTEP / MUFU_PER_HSTEP
    # HBM side, for the record: one read of the pose stream + one 64-byte record per window
    algo_bytes = n_frames * 16 + n_win * (64 + 16)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    out["hbm"] = {"algorithmic_bytes_per_launch": algo_bytes,
                  "achieved_gbs": algo_bytes / (kern_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                  "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                  "frac": algo_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak}
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass
    out["traffic"] = traffic
    return out


def prep_rows(ctx, dev, timed):
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
    vo_ms = timed(lambda: vo_prepare_device(off, D, F, x, y, rot, stamp, out=vo_out), k, 1) / k
    gps_ms = timed(lambda: gps_prepare_device(off, D, F, lat, lon, speed, stamp, out=gps_out,
                                              scratch=scratch, status=status), k, 1) / k
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    vo_gbs, gps_gbs = F * 136 / (vo_ms * 1e-3) / 1e9, F * 72 / (gps_ms * 1e-3) / 1e9
    return {"workload": "64 drives x 10000 frames", "vo_ms": vo_ms, "gps_ms": gps_ms,
            "vo_frames_per_s": F / (vo_ms * 1e-3), "gps_frames_per_s": F / (gps_ms * 1e-3),
            "vo_hbm": {"achieved_gbs": vo_gbs, "peak_gbs": hbm, "frac": vo_gbs / hbm},
            "gps_hbm": {"achieved_gbs": gps_gbs, "peak_gbs": hbm, "frac": gps_gbs / hbm},
            "note": "GPS includes the per-drive sequential path sum and de-duplication scan (one warp per drive)"}


def formats_rows(ctx, dev, timed):
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
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    out = {"workload": f"{D} drives x {n} rows per file"}
    for name, text, cols, rc, sc in (("log", b1.getvalue(), LOG_COLUMNS, None, "Timestamp"),
                                     ("cache", b2.getvalue(), CACHE_COLUMNS, "rot", None)):
        blob = text.encode()
        st = stage_csv_files([blob] * D, dev)
        k = 5
        ms = timed(lambda: parse_staged(st, cols, rc, sc), k, 2) / k
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


def dense_grid(ctx, dev, timed, args):
    """BASELINE configs[2]: 256x256 grid, 60-step windows, 4096 windows (1.6e10 hyp-steps)."""
    import torch

    from vehiclemodelvisualodometry_b200 import DriveSet, grid_search, plan_windows
    from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

    name = "config3_dense_256x256_w60"
    n_frames, cfg = make_cfg(name)
    batch = synthetic_drives(1, n_frames, seed=BASE_SEED + 3)
    t, vo, _, _ = batch.drive(0)
    drives = DriveSet.from_arrays([t], [batch.dt], vo=[vo], device=dev)
    plan = plan_windows(cfg, drives)
    out = torch.empty((plan.n_windows, 64), dtype=torch.uint8, device=dev)
    k = 3
    ms = timed(lambda: grid_search(cfg, drives, plan, out=out), k, 1) / k
    from vehiclemodelvisualodometry_b200 import _lib

    rec = out.cpu().numpy().view(_lib.RESULT_DTYPE).reshape(-1)
    hsteps = cfg.grid_v * cfg.grid_s * int(rec["n_steps"].astype(np.int64).sum())
    return {"workload": name, "windows": plan.n_windows, "hypothesis_steps": hsteps, "kernel_ms": ms,
            "value": hsteps / (ms * 1e-3), "unit": UNIT,
            "achieved_gmufu": hsteps * MUFU_PER_HSTEP / (ms * 1e-3) / 1e9,
            "rescored_per_window": float(rec["n_rescored"].mean())}


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


def cpu_baseline(workload, budget_s=12.0):
    cores = host_threads()
    cpu_pass(workload, BASE_SEED, max_windows=256, threads=cores)  # warm the library and the threads
    n_frames, cfg = make_cfg(workload)
    total_win = n_frames - 2 * cfg.window_frames
    steps = sec = 0.0
    passes = nwin = 0
    t_end = time.perf_counter() + budget_s
    while passes < 3 or time.perf_counter() < t_end:
        s1, t1, nwin = cpu_pass(workload, BASE_SEED, threads=cores)
        steps += s1
        sec += t1
        passes += 1
        if passes >= 200:
            break
    return {"value": steps / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{passes} passes over all {nwin} of {total_win} windows of {workload}; "
                      f"oracle/vmvo_oracle.c (float64, OpenMP over windows), {sec:.1f} s of search time"}


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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (same generator and seed as the b200 arm)",
        "config": {"workload": args.workload, "frames_per_drive": n_frames,
                   "grid": [cfg.grid_v, cfg.grid_s], "window_steps": cfg.window_frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python (15 us per model step, ~6.7e4 steps/s/core, SURVEY.md F8) "
                "and cannot travel to the GPU box; this arm is its C restatement on all host threads",
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
