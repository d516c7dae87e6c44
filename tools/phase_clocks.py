"""Per-phase clock64() accounting of the search kernel (how long a team spends in each phase of a
window), on a separately built, instrumented copy of the library -- the shipped one is untouched.

    python tools/phase_clocks.py build     # here: copies csrc/ to tools/_build/src_clk, inserts the
                                           # markers, compiles tools/_build/libvmvo_clk.so
    python tools/phase_clocks.py run       # on the GPU box: config 2, prints clk per window and phase

The markers are placed by text anchors in vmvo_search_kernels.cuh (the script fails loudly when one moves).
Lane 0 of the first two warps of every team adds the clocks since its previous marker to a
shared-memory accumulator; the totals are added to a device array at kernel end and read back
through vmvo_exp_clk (exported by the instrumented build only).
"""
import ctypes as C
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "_build", "src_clk")
LIB = os.path.join(ROOT, "tools", "_build", "libvmvo_clk.so")
NAMES = ["0 plan entry / TMA wait", "1 A1 local frame", "2 A2 seeds", "3 A3 targets + A4 TL table",
         "4 table statistics, header", "5 VD table", "6 scan", "7 band, minimum, U", "8 candidate list",
         "9 deferral / float64 re-score", "10 record", "11 -"]


def build():
    shutil.rmtree(SRC, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, "vehiclemodelvisualodometry_b200", "csrc"), SRC)
    inc = os.path.join(ROOT, "include", "vmvo_b200.h")
    for f in glob.glob(SRC + "/*"):
        t = open(f).read()
        if '"../../include/vmvo_b200.h"' in t:
            open(f, "w").write(t.replace('"../../include/vmvo_b200.h"', '"%s"' % inc))
    p = os.path.join(SRC, "vmvo_search_kernels.cuh")
    s = open(p).read()

    def after(anchor, text):
        nonlocal s
        assert s.count(anchor) == 1, anchor
        s = s.replace(anchor, anchor + text)

    def before(anchor, text):
        nonlocal s
        assert s.count(anchor) == 1, anchor
        s = s.replace(anchor, text + anchor)

    after("    mbar_wait(&hd->mbar[cur], (unsigned)((it >> 1) & 1));\n", "    CLK(0);\n")
    before("    // ---- phase A2 (warp 0)", "    CLK(1);\n")
    before("    const int n_targets = hd->wi.n_targets;\n    const int N = hd->wi.n_steps;", "    CLK(2);\n")
    after("    const bool bad = !(dmax < CUDART_INF_F);\n", "    CLK(3);\n")
    before("      for (int pidx = 0; pidx < n_pass; ++pidx) {", "      CLK(4);\n")
    before("        if (pidx == 0 && use_skip) {   // hd->bw and hd->ts are visible now", "        CLK(5);\n")
    before("        float mj = CUDART_INF_F;       // the item's smallest cost", "        CLK(6);\n")
    after("        U = fminf(U, bm);\n", "        CLK(7);\n")
    before("          const int overflow = team.any(pend != 0);\n", "          CLK(8);\n")
    after("          if (overflow || (last && !deferred)) process_list();\n", "          CLK(9);\n")
    before("    team.sync();\n  }\n}\n\n// ---- window preparation as a pass", "    CLK(10);\n")
    s = s.replace("namespace vmvo {\n", "namespace vmvo {\nstatic __device__ unsigned long long g_clk[2][12];\n"
                  "#define CLK(i) do { if (lane == 0 && warp < 2) { long long t_ = clock64(); "
                  "s_clk[team.id][warp][i] += t_ - s_prev[team.id][warp]; s_prev[team.id][warp] = t_; } } while (0)\n", 1)
    before("  const bool fetcher = tid == T - 32;\n",
           "  __shared__ long long s_clk[8][2][12], s_prev[8][2];\n"
           "  if (lane == 0 && warp < 2) { for (int i = 0; i < 12; ++i) s_clk[team.id][warp][i] = 0; "
           "s_prev[team.id][warp] = clock64(); }\n")
    after("    CLK(10);\n    team.sync();\n  }\n",
          "  if (lane == 0 && warp < 2) for (int i = 0; i < 12; ++i) "
          "atomicAdd(&g_clk[warp][i], (unsigned long long)s_clk[team.id][warp][i]);\n")
    open(p, "w").write(s)
    # the counters are per translation unit: read those of the MODE 2 kernels (config 2 runs them)
    p = os.path.join(SRC, "vmvo_search_prep.cu")
    s = open(p).read()
    s += ('\nextern "C" int vmvo_exp_clk(unsigned long long* out) {\n  unsigned long long z[24] = {0};\n'
          "  cudaDeviceSynchronize();\n  cudaMemcpyFromSymbol(out, vmvo::g_clk, sizeof(z));\n"
          "  cudaMemcpyToSymbol(vmvo::g_clk, z, sizeof(z));\n  return 0;\n}\n")
    open(p, "w").write(s)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
           "-Xcompiler", "-fPIC", "--threads", "4", "-o", LIB] + [os.path.join(SRC, f) for f in
                                                                 ("vmvo_search.cu", "vmvo_search_lean.cu", "vmvo_search_prep.cu", "vmvo_aux.cu", "vmvo_prep.cu", "vmvo_csv.cu")]
    subprocess.run(cmd, check=True)
    print("built", LIB)


def run():
    os.environ["VMVO_B200_LIBRARY"] = LIB
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch

    import bench
    from vehiclemodelvisualodometry_b200 import DriveSet, _lib, grid_search, plan_windows
    from vehiclemodelvisualodometry_b200.synthetic import synthetic_drives

    n, cfg = bench.make_cfg("config2_single_drive_10k_32x32_w30")
    b = synthetic_drives(1, n, seed=bench.BASE_SEED)
    t, vo, _, _ = b.drive(0)
    dr = DriveSet.from_arrays([t], [b.dt], vo=[vo])
    plan = plan_windows(cfg, dr)
    lib = _lib.context(0).lib
    for a in sys.argv[2:]:            # key=value: the library's tuning hook
        k, v = a.split("=")
        _lib.context(0).set_tuning(k, int(v))
    for _ in range(3):
        grid_search(cfg, dr, plan)
    out = (C.c_ulonglong * 24)()
    lib.vmvo_exp_clk(out)
    reps = 10
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        grid_search(cfg, dr, plan)
    e.record()
    torch.cuda.synchronize()
    print("ms per launch (instrumented build)", a.elapsed_time(e) / reps)
    lib.vmvo_exp_clk(out)
    v = np.array(list(out), dtype=np.float64).reshape(2, 12) / (reps * plan.n_windows)
    for i, nm in enumerate(NAMES):
        print(f"{nm:32s} warp 0 {v[0, i]:8.0f} clk   warp 1 {v[1, i]:8.0f} clk")
    print(f"{'per window':32s} warp 0 {v[0].sum():8.0f} clk   warp 1 {v[1].sum():8.0f} clk")


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1] if len(sys.argv) > 1 else "run"]()
