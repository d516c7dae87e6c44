"""Instruction counts of the search kernel from an .ncu-rep (captured with --set full
--import-source on) -> one entry of profiles/kernel_counts_r02.json, the file bench.py's roofline
reads its EXECUTED figures from.

    python tools/ncu_counts.py <rep> <workload name> <hypothesis-steps of the captured launch>

Per hypothesis-step: FP32 lane-operations executed on the FMA pipe (FFMA / FMUL / FADD count one per
lane, the packed FFMA2 / FADD2 / FMUL2 two), the FMA-pipe issue time those instructions (and the
IMADs that share the pipe) take at the rates measured by tools/pipe_probe.cu, MUFU operations; and
the hardware's own pipe-activity counters of the capture.
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload, hsteps = sys.argv[1], sys.argv[2], int(sys.argv[3])

# clk per warp-instruction per SM sub-partition at 1965 MHz (tools/pipe_probe.cu, profiles/README.md)
FMA_PIPE_CLK = {"FFMA": 1.11, "FMUL": 1.11, "FADD": 1.11, "FFMA2": 2.86, "FADD2": 2.26, "FMUL2": 2.60,
                "IMAD": 2.06}
LANE_OPS = {"FFMA": 1, "FMUL": 1, "FADD": 1, "FFMA2": 2, "FADD2": 2, "FMUL2": 2}

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2]
get = lambda k: float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] else None
unit = lambda k: rows[1][hdr.index(k)] if k in hdr else ""

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = next(i for i, x in enumerate(srows) if "Address" in x and "Source" in x)
hd = srows[h]
ie, isrc = hd.index("Instructions Executed"), hd.index("Source")
ops = Counter()
for x in srows[h + 1:]:
    if len(x) > ie and x[ie]:
        tok = x[isrc].split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        ops[op] += int(x[ie])
total = sum(ops.values())

dur = get("gpu__time_duration.sum")
dur_us = dur * (1e3 if unit("gpu__time_duration.sum") == "ms" else 1.0 if unit("gpu__time_duration.sum") == "us" else 1e-3)
dram = 0.0
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    v = get(k)
    if v:
        dram += v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit(k), 1.0)
entry = {
    "source": f"ncu --set full of {os.path.basename(rep)} (profiles/{os.path.basename(rep).replace('.ncu-rep', '.txt')})",
    "kernel": r[hdr.index("Kernel Name")],
    "hypothesis_steps": hsteps,
    "duration_us_under_ncu": dur_us,
    "warp_instructions": total,
    "warp_instructions_per_hypothesis_step_x32": total * 32 / hsteps,
    "opcode_counts": {k: v for k, v in ops.most_common(24)},
    "fp32_lane_ops_per_hypothesis_step": sum(ops[o] * n for o, n in LANE_OPS.items()) * 32 / hsteps,
    "fma_pipe_clk_per_hypothesis_step": sum(ops[o] * c for o, c in FMA_PIPE_CLK.items()) * 32 / hsteps,
    "mufu_per_hypothesis_step": ops["MUFU"] * 32 / hsteps,
    # share of the grid's hypothesis-steps the scans executed (the rest: pruned).  The packed scan runs
    # 20 FFMA2 per 8 hypothesis-steps of a thread and FFMA2 occurs nowhere else in the kernel
    "executed_share_of_hypothesis_steps": ops["FFMA2"] / 20 * 8 * 32 / hsteps,
    "hw_pipe_fma_cycles_active_pct": get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "hw_pipe_xu_inst_pct": get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "hw_issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "dram_bytes_per_launch": dram,
}
path = os.path.join(ROOT, "profiles", "kernel_counts_r02.json")
try:
    allc = json.load(open(path))
except Exception:
    allc = {}
allc[workload] = entry
json.dump(allc, open(path, "w"), indent=1)
print(json.dumps(entry, indent=1))
