"""Lists the innermost loops of a kernel's SASS that hold MUFU.SIN (the scan loops): address range,
instruction count, opcode histogram, local-memory traffic.  Usage:
  python tools/sass_loops.py <lib.so> <substring of the mangled kernel name> [--dump]"""
import collections
import re
import subprocess
import sys


def main():
    lib, key = sys.argv[1], sys.argv[2]
    dump = "--dump" in sys.argv
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, funcs = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            funcs[fn] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and fn:
            funcs[fn].append((int(m.group(1), 16), m.group(2).strip()))
    for fn, ins in funcs.items():
        if key not in fn:
            continue
        print("==", fn, len(ins), "instructions")
        addr = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\S+,\s+)*0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr:
                loops.append((addr[int(m.group(1), 16)], i))
        inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
        for lo, hi in inner:
            body = ins[lo:hi + 1]
            if not any("MUFU.SIN" in t for _, t in body):
                continue
            hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
            loc = sum(1 for _, t in body if t.split()[0] in ("STL", "LDL") or " STL" in t or " LDL" in t)
            print(f"  loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {len(body)} instructions, local-memory ops {loc}")
            print("    ", dict(hist.most_common()))
            if dump:
                for a, t in body:
                    print(f"    /*{a:05x}*/ {t}")


main()
