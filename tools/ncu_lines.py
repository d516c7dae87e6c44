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
