"""Per-source-line attribution of an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]

Prints, for the source lines with the most stall samples, the warp-instructions executed and the
samples (share of the kernel); inlined helpers are attributed to their own file:line.
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        ie, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > ie and r[0].isdigit():
        num = lambda v: int(v) if v.lstrip("-").isdigit() else 0
        lines.append((cur_file, int(r[0]), r[1].strip(), num(r[ie]), num(r[ismp])))
ti, ts = sum(l[3] for l in lines), sum(l[4] for l in lines)
print(f"total warp-instructions {ti}, samples {ts}")
for f, n, src, e, s in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{100 * s / ts:5.1f}% smp {100 * e / ti:5.1f}% inst  {f}:{n:<5d} {src[:100]}")

# ---- per-phase totals for vmvo_search_kernels.cuh (line ranges found from the phase markers of the file) ----
import os
import re

src_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "vehiclemodelvisualodometry_b200", "csrc", "vmvo_search_kernels.cuh")
if os.path.isfile(src_path):
    text = open(src_path).read().splitlines()
    marks = []
    pats = [("band / make_band", r"^struct BandWin"), ("float64 re-score (warp_cost64)", r"^__device__ double warp_cost64"),
            ("scan loop: generic", r"^__device__ __forceinline__ bool scan_item\("),
            ("scan loop: packed / rotation", r"^__device__ __forceinline__ float2 pk\("),
            ("kernel prologue / queue / TMA wait", r"^vmvo_window_search_kernel\("),
            ("A1 local frames", r"// ---- phase A1"), ("A2 seeds / decimation", r"// ---- phase A2"),
            ("A3 targets", r"// ---- phase A3"), ("A4 TL table", r"// ---- phase A4"),
            ("B setup + candidate re-score driver", r"// ---- phase B"),
            ("B per pass: VD table", r"const int n_pass = "), ("B per pass: scan call + band", r"const int q = pass \* T \+ tid;"),
            ("B per pass: min / candidates", r"// upper bound on the minimum"),
            ("D results", r"// ---- phase D"), ("host", r"^static int launch_search")]
    for name, pat in pats:
        for i, l in enumerate(text):
            if re.search(pat, l):
                marks.append((i + 1, name))
                break
    marks.sort()
    tot = {}
    for f, n, src, e, s in lines:
        if f in ("vmvo_search.cu", "vmvo_search_kernels.cuh"):
            name = "before"
            for ln, nm in marks:
                if n >= ln:
                    name = nm
        else:
            name = "inlined helpers (" + f + ")"
        a = tot.setdefault(name, [0, 0])
        a[0] += e
        a[1] += s
    print("-- by phase (note: file line numbers must match the profiled build)")
    for name, (e, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"   {100 * s / ts:5.1f}% smp {100 * e / ti:5.1f}% inst  {name}")
