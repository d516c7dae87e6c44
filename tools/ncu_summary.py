"""Summarise an .ncu-rep (raw + source pages) into a short text report for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg",
]
# pipe activity as the hardware counts it (cycles a pipe is busy, not instructions sent to it: a
# packed FFMA2 keeps the FMA pipe busy longer than a scalar FFMA), and the issue rate
EXTRA = ["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
         "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
         "sm__inst_issued.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {r[i]:>18s} {units[i]}")
    for k in EXTRA:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {r[i]:>18s} {units[i]}")
    stall = [(h, r[i]) for i, h in enumerate(hdr)
             if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    stall = sorted(((float(v or 0), h) for h, v in stall), reverse=True)[:8]
    print("-- top stall reasons (warps stalled per issue-active cycle)")
    for v, h in stall:
        print(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:8.3f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if "Address" in r and "Source" in r:
        h = i
        break
if h is not None:
    hd = rows[h]
    ie, isrc, ismp = hd.index("Instructions Executed"), hd.index("Source"), hd.index("# Samples")
    data = [(r[isrc], int(r[ie] or 0), int(r[ismp] or 0)) for r in rows[h + 1:] if len(r) > ie]
    tot = sum(d[1] for d in data)
    ts = sum(d[2] for d in data)
    ops = Counter()
    for s, e, _ in data:
        tok = s.split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        ops[op] += e
    print(f"-- SASS: {len(data)} instructions, {tot} warp-instructions executed, {ts} samples")
    print("   opcode mix:", ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(12)))
    mx = max(d[1] for d in data)
    hot = [d for d in data if d[1] >= 0.5 * mx]
    print(f"   hottest region: {len(hot)} instructions executed >= {0.5 * mx:.0f} times each = "
          f"{100 * sum(d[1] for d in hot) / tot:.1f}% of instructions, "
          f"{100 * sum(d[2] for d in hot) / max(ts, 1):.1f}% of samples; "
          f"MUFU in it: {sum(1 for d in hot if d[0].split()[0].startswith('MUFU') or (d[0].startswith('@') and 'MUFU' in d[0]))}")
