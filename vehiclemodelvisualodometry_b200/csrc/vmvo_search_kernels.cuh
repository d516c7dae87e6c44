// Fused window search: window staging (TMA bulk copy) -> local frame (a7) -> decimation (a9)
// -> seeds -> hypothesis grid generated on chip -> batched bicycle integration (a1/a2)
// -> cost (a10) -> block argmin -> exact float64 re-score of every hypothesis whose FP32 cost
// lies within a rigorous error band of the minimum -> per-window result record.
//
// Replaces mpc_run (vmvo/utils/mpc.py:14-122) and the per-window preparation of
// optimize_trajectory (vmvo/scripts/optimize_trajectory_v2.py:49-96) of the reference.
//
// Work decomposition: a TEAM of 1, 2, 4 or 8 warps works on one window
// at a time and pulls windows from a global queue (persistent grid, 8 warps per CTA, teams
// synchronise on their own named barrier -- a one-warp team needs no barrier at all, so small
// grids run one window per warp).  Per window the team builds two small tables in shared
// memory -- TL[k][j] = tan(delta_k(j))/L and VD[k][i] = V_k(i)*dt -- so the scan's inner loop
// is, per hypothesis-step: 1 FFMA (heading), sin+cos on the SFU, 4 FP32 ops for the position
// error recurrence and 2 FFMA for the cost.  A thread owns one steering rate j and C = 8
// consecutive accelerations per pass.
//
// Small teams park windows with long candidate lists (near-ties of a slow vehicle) in a slot;
// vmvo_deferred_rescore_kernel, launched right behind the search, re-scores them with the whole GPU,
// four hypotheses per warp where they stop within a few steps -- same arithmetic, same records.
// (device code and launch templates; included by vmvo_search.cu, vmvo_search_lean.cu and
// vmvo_search_prep.cu, one kernel MODE each)
#pragma once
#include "vmvo_device.cuh"
#include "vmvo_internal.h"

#include <math_constants.h>
#include <stdlib.h>

#ifndef VMVO_MINB
#define VMVO_MINB 2     // CTAs of eight warps per SM the search kernel is built for (128 registers)
#endif

namespace vmvo {

constexpr int kCandPerWarp = 128;  // candidate list entries per team warp (flushed when full)
constexpr int kMaxWarps = 16;      // most warps per CTA (and per team) of any kernel variant
constexpr int kHeaderBytes = 1280;

struct SearchParams {
  int gv, gs;
  int n_ic;              // ceil(gv / C)
  int n_items;           // n_ic * gs  (one item = one thread's C hypotheses)
  int target_mode, target_offset, seed_mode, primary;
  int maxp;              // pose capacity per window
  int use_vo, use_gps;   // position terms with non-zero weight
  int load_vo, load_gps; // streams staged into shared memory
  int vd_cols;           // columns of the VD table (accelerations covered by one pass)
  int team_warps;        // warps per team (1, 2, 4 or 8)
  int cand_cap;          // candidate list entries per team (<= kCandPerWarp * team_warps)
  int allow_fast;        // use the packed / rotation scan where its preconditions hold
  int prune_every;       // steps between two pruning votes of a scanning warp (0: every scan runs to the end)
  // launch constants the kernel would otherwise divide for, window after window: passes per window,
  // threads that share a rate in the TL fill / a column in the VD fill, and log2 of gs, vd_cols and
  // the team's thread count when they are powers of two (-1: divide)
  int n_pass, tl_slices, vd_kpar, gs_sh, vd_sh, t_sh;
  int scan_hs;           // the launch's float64 prefix sums are Hillis-Steele (the many-pass kernel)
  double w_vo, w_gps, w_imu, k_steer;
  double L, ratio, max_steer, max_accel, max_rate;
  double delta_max, kappa;     // kappa = 2*delta_max / sin(2*delta_max): tan's condition number
  // for the FP32 tables only (the float64 re-score divides as the spec does): the grid spacings
  // max_accel / (gv - 1) and max_rate / (gs - 1) (0 on a one-point axis), and 1 / L
  double acc_step, rate_step;
  float inv_L;
  const long long* win_start;   // the window plan (vmvo_plan_windows), or all three NULL: frames mode,
  const int* win_len;           // extents computed by the fetcher from the offsets below
  const int* win_drive;
  const long long* drive_off;   // [n_drives + 1] frames
  const long long* win_off;     // [n_drives + 1] windows
  int n_drives, window_frames;
  const double* dt_drive;
  const void* vo;        // pose streams: float4 or double4 per frame (kernel template SF)
  const void* gps;
  const void* imu;       // yaw per frame, float or double
  const double* seeds;
  vmvo_window_result* results;
  double* out_poses;
  double* out_steer;
  double* out_vel;
  int out_stride;
  long long n_windows;
  unsigned long long* work_counter;
  // seed_mode chained: the queue hands out RUNS (the windows of one drive, in order) and the
  // steering seed of a window is the last steering angle of the previous window's optimum
  const long long* run_offsets;   // [n_runs + 1] window index ranges, or NULL
  long long n_runs;
  // the window deal (vmvo_exchange): queue item r of this rank is global window
  // ((r >> sh_block_sh) * sh_world + sh_rank) << sh_block_sh | (r & (block - 1)); sh_world <= 1 or
  // sh_block_sh < 0: item r is window r.  In chained mode whole runs are dealt: r * sh_world + sh_rank.
  int sh_world, sh_rank, sh_block_sh;
  long long n_local;     // queue items of this rank (an upper bound: the last block may overhang)
  // result mirrors: the gather fused into the epilogue (peers' gather buffers, indexed like results)
  int n_mirrors;
  vmvo_window_result* mirrors[VMVO_MAX_MIRRORS];
  unsigned* epoch;       // step counter of the exchange, advanced once per launch (or NULL)
  // deferred windows: candidate lists of at least defer_min entries are parked here (one slot
  // per window) and re-scored by vmvo_deferred_rescore_kernel with the whole GPU
  unsigned char* defer_buf;
  unsigned* defer_count;
  int defer_slots, defer_slot_bytes, defer_min;
  // the second kernel runs BESIDE the end of the search (programmatic dependent launch): a parked
  // window is published through defer_ready[slot] and every team that runs out of windows counts in
  // windows_done, so the second kernel knows when the set of slots is final without waiting for the
  // grid to drain
  unsigned* defer_ready;               // [defer_slots], cleared with the counters before the launch
  unsigned long long* windows_done;    // (counts finished TEAMS)
  long long n_todo;                    // teams of the launch
  // window preparation as a pass of its own (vmvo_window_prep_kernel): one record per queue item,
  // staged into shared memory by the fetcher instead of the raw poses; NULL: phases A1-A3 run in
  // the search kernel
  unsigned char* prep;
  int prep_stride;
  float* dbg_cost;       // optional [n_windows][gv*gs] FP32 scan costs   (tests only)
  float* dbg_err;        // optional [n_windows][gv*gs] error-band widths (tests only)
};

// per-window scalars shared by the CTA
struct WinInfo {
  double v_seed, s_seed, dt;
  int n_targets, n_steps, status;
  // structural duplicates: rows i < n_dead never move (V_k = 0 for every k) and
  // share one cost for every j; with the seed at a steering bound, rates j in [sat_lo, sat_hi]
  // all clamp to the same sequence.  Only the lowest index of a class can be the argmin.
  int n_dead, sat_lo, sat_hi;
};

// threads of a team that share one steering rate in the TL fill (each takes every kpar-th step)
__host__ __device__ inline int tl_kpar(int T, int gs) { return T >= gs ? T / gs : 1; }

// ---- shared-memory carve-up (same function on host and device) ---------------------------
struct SmemLayout {
  int off_raw, off_loc, off_loci, off_tgt, off_df, off_dab, off_fi, off_keep, off_tl, off_js,
      off_ts, off_tsp, off_vd, off_cand, total;
  // P poses per window, n_streams pose streams staged, optional terms only when configured
  int raw_buf;   // bytes of one staging buffer
  __host__ __device__ SmemLayout(int P, int gs, int vd_cols, int team_warps, int n_streams,
                                 bool dual, bool imu, bool traverse, int pose_bytes, int prep_bytes = 0) {
    int o = kHeaderBytes;                       // barriers, window ids, reductions
    // raw[2 buffers][streams][P] float4/double4 -- or, with a preparation pass, two window records
    raw_buf = n_streams * P * pose_bytes;
    if (prep_bytes > raw_buf) raw_buf = prep_bytes;
    off_raw = o;  o += 2 * raw_buf;
    off_loc = o;  o += n_streams * 3 * P * 8;   // double loc[streams][3][P]  (lx, ly, lth)
    off_loci = o; o += imu ? P * 8 : 0;         // double imu yaw relative to the window start
    off_tgt = o;  o += (2 + (dual ? 2 : 0) + (imu ? 1 : 0)) * P * 8;  // tAx, tAy, [tBx, tBy], [tI]
    off_df = o;   o += P * 8;                   // float2 D[k]   = T_A[k-off] - T_A[k-1-off]
    off_dab = o;  o += dual ? P * 8 : 0;        // float2 DAB[k] = T_A[k-off] - T_B[k-off]
    off_fi = o;   o += imu ? P * 4 : 0;         // float  imu target per step
    off_keep = o; o += traverse ? P * 4 : 0;    // int keep
    o = (o + 15) & ~15;
    off_tl = o;   o += P * gs * 4;              // float TL[k][j]
    off_js = o;   o += ((gs + 3) & ~3) * 4;     // float K * sum_k S_k(j)^2 (steering penalty)
    off_ts = o;   o += 3 * ((gs + 3) & ~3) * 4; // float max_k |TL|, sum_k |TL|, sum_k k |TL|  per j
    // the same three, per (step slice, j), as the TL fill leaves them: T / gs threads share a rate
    off_tsp = o;  o += 3 * tl_kpar(team_warps * 32, gs) * ((gs + 3) & ~3) * 4;
    o = (o + 15) & ~15;
    off_vd = o;   o += P * vd_cols * 4;         // float VD[k][m]
    o = (o + 15) & ~15;
    off_cand = o; o += kCandPerWarp * team_warps * 8;  // uint2 (hypothesis, lower-bound bits)
    total = (o + 127) & ~127;
  }
};

// acceleration chunks (of C) one pass of T items can touch: items [pass*T, pass*T + T) of the
// (chunk, j) item space, gs items per chunk
__host__ __device__ inline int chunks_per_pass(int T, int gs) {
  if (T % gs == 0) return T / gs;
  if (gs % T == 0) return 1;
  return (T + gs - 1) / gs + 1;
}

// header of a deferred window's slot; the float64 targets and the candidate list follow it
struct DeferHdr {
  long long w;
  int n_steps, status, count, n_rescored;
  float U;
  int best_h;                 // best of the re-scores already done in the search kernel, or -1
  double best_cost, bpose[3];
  WinInfo wi;
};
constexpr int kDeferHdrBytes = 128;
static_assert(sizeof(DeferHdr) <= kDeferHdrBytes, "deferred-window header too large");

// ---- window preparation records (vmvo_window_prep_kernel -> the search kernel) -------------------
// What phases A1-A3 produce for one window: seeds, step count, status, duplicate classes, the band's
// maxima, then the FP32 target increments of the scan and the float64 targets of the re-score, laid
// out like the search kernel's own arrays so that it can work on the staged record in place.
struct PrepHdr {
  double v_seed, s_seed, dt;
  int n_targets, n_steps, status;     // status: EMPTY | NO_FRAMES / TOO_LONG for a window without a search
  int n_dead, sat_lo, sat_hi, len;
  float dmax, dabmax, imax;           // dmax = +inf: a non-finite input somewhere in the window
};
constexpr int kPrepHdrBytes = 128;
static_assert(sizeof(PrepHdr) <= kPrepHdrBytes, "preparation header too large");
struct PrepLayout {
  int off_df, off_dab, off_fi, off_tgt, total;
  __host__ __device__ PrepLayout(int P, bool dual, bool imu) {
    int o = kPrepHdrBytes;
    off_df = o;  o += P * 8;                    // float2 D[k]
    off_dab = o; o += dual ? P * 8 : 0;         // float2 DAB[k]
    off_fi = o;  o += imu ? P * 4 : 0;          // float imu target per step
    o = (o + 15) & ~15;
    off_tgt = o; o += (2 + (dual ? 2 : 0) + (imu ? 1 : 0)) * P * 8;
    total = (o + 127) & ~127;
  }
};

// window-level inputs of the FP32 error band
struct BandWin {       // window-level inputs, identical in every thread
  float n, s2, s4;     // N, sum k^2, bound on sum ((k^2+k)/2)^2
  float dmax;          // max |target increment|
  float dabmax;        // max |T_A - T_B|
  float imax;          // max |imu target|
  float eps_tl;        // relative error of TL
  float wpos, wimu;    // weight sums
  float c2;            // J-proportional coefficient (already includes the safety factor)
};
struct SmemHeader {
  uint64_t mbar[2];
  long long wid[2];
  vmvo_window_result rec;   // the window's record, assembled by one thread, stored by a warp
  WinInfo wi;
  BandWin bw;
  float red[3 * kMaxWarps];
  double bcost[kMaxWarps];
  double bpose[kMaxWarps][3];
  int bh[kMaxWarps];
  int nres[kMaxWarps];
  float ts[3 * 8];           // per warp (<= 8 per team): max over its rates of max|TL|, sum|TL|, sum k|TL|
  int count;
  int winner;
  int slot;          // deferred-window slot handed out for this window (or -1)
  int first[2];      // chained mode: the window is the first of its run
  // plan entry of the window in each staging buffer, read once by the fetcher (the other threads
  // would each wait for two dependent L2 round trips at the top of the window)
  int wlen[2];
  long long wstart[2];
  double wdt[2];
  long long witem[2];   // queue item of the window in each staging buffer (its preparation record)
};
static_assert(sizeof(SmemHeader) <= kHeaderBytes, "header too large");

// x / d for x >= 0, by a shift when the host found d to be a power of two (sh = log2 d, else -1)
__device__ __forceinline__ int div_sh(int x, int d, int sh) { return sh >= 0 ? x >> sh : x / d; }
static inline int log2_exact(int d) {
  for (int b = 0; b < 31; ++b)
    if (d == (1 << b)) return b;
  return -1;
}

// ---- team barriers: named barrier 1 + team index; a one-warp team only needs __syncwarp ------
struct Team {
  int warps, threads, id;
  __device__ __forceinline__ void sync() const {
    if (warps == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id + 1), "r"(threads) : "memory");
  }
  __device__ __forceinline__ int any(int pred) const {
    if (warps == 1) return __any_sync(FULL, pred);
    int r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 q, %3, 0;\n"
        "bar.red.or.pred p, %1, %2, q;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(r)
        : "r"(id + 1), "r"(threads), "r"(pred)
        : "memory");
    return r;
  }
};

// ---- FP32 error band ----------------------------------------------
// |J_fp32 - J_fp64| <= c0 + c1*sqrt(J) + c2*J for every hypothesis of an item, from the item's
// own step length, heading excursion and tan magnitude.
// Square root for the band: one MUFU (sqrt.approx, relative error <= 2^-22) scaled up by 2^-20 so
// that the result is never below the true root -- every use in the band wants an upper bound (a wider
// band, a larger threshold), and the IEEE sqrtf costs ten instructions and a dependent chain per use,
// three uses per thread and pass.  NaN and +inf pass through.
__device__ __forceinline__ float sqrt_up(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 1.000001f;
}

struct Band {
  float c0, c1, c2;
  __device__ __forceinline__ float err(float J) const { return fmaf(c1, sqrt_up(J), fmaf(c2, J, c0)); }
  // Largest cost that can still be a candidate under the bound U on the minimum: g(J) = J - err(J)
  // is convex with g(0) <= 0 <= U, so {J >= 0 : g(J) <= U} = [0, T] with sqrt(T) the larger root of
  // (1 - c2) s^2 - c1 s - (c0 + U) = 0.  One sqrt per thread instead of one per hypothesis; T is
  // rounded up (relative 2^-16) so that the test J <= T never drops a hypothesis g(J) <= U keeps.
  __device__ __forceinline__ float threshold(float U) const {
    if (!(c2 < 0.5f)) return CUDART_INF_F;
    const float a = 1.f - c2;
    // (2a is in (1, 2]: the approximate division is good to 2 ulp, far inside the rounding-up)
    const float s = __fdividef(c1 + sqrt_up(fmaf(c1, c1, 4.f * a * (c0 + U))), 2.f * a);
    return s * s * 1.0000153f;
  }
};

// fast = the packed / rotation scan (scan_item_fast): headings are the affine form A_k + a_i*B_k
// (error <= (eps_TL + 5u + k*u) * Theta', Theta' = sum (V_w*dt + |a|max*t*dt)*|TL|), half of the
// sin/cos pairs come from one rotation of a MUFU pair by a MUFU angle.
template <bool IMU>
__device__ __forceinline__ Band make_band(const BandWin& w, float vmax, float theta_tv, float tlmax,
                                          bool fast) {
  const float u = 5.9604644775390625e-8f;           // 2^-24
  const float SF = 2.0f;                            // safety factor on the whole bound
  const float dth = vmax * tlmax;                   // largest heading step
  // heading error grows <= g per step
  const float g = fast ? u * theta_tv : (w.eps_tl + u) * dth + u * theta_tv;
  const float mufu = 4.76837158203125e-7f + 4.f * u * theta_tv;  // MUFU sin/cos, |x| <= theta_tv
  const float eps_trig = fast ? 2.f * mufu + 4.f * u : mufu;
  const float head = fast ? vmax * (w.eps_tl + 5.f * u) * theta_tv : 0.f;
  const float q1 = vmax * (eps_trig + u) + 2.f * u * w.dmax + 2.f * u * w.dabmax + head;
  const float q2 = vmax * g;
  const float e2pos = 3.f * (q1 * q1 * w.s2 + q2 * q2 * w.s4);
  float e2 = w.wpos * e2pos;
  if (IMU) {     // (compile-time: a search without the yaw term does not pay for its band)
    const float ei = u * (w.imax + 3.1415927f) + (theta_tv * 0.15915494f + 1.f) * 1.75e-7f;
    e2 += w.wimu * (2.f * (g * g * w.s2 + w.n * ei * ei));
  }
  Band b;
  b.c0 = SF * 2.f * e2;
  b.c1 = SF * 2.f * sqrt_up(2.f * e2);
  b.c2 = w.c2;
  return b;
}

// ---- float64 cost of one hypothesis, one warp, one step per lane ---------------------------
// the cost term of step k (state after the step against target k - target_offset; mpc.py:70-78)
template <bool DUAL, bool IMU>
__device__ __forceinline__ double step_term(const SearchParams& p, const double* tgt, int P, int k,
                                            const Pose<double>& pz, double s, double wA, double wB) {
  const double* tAx = tgt;
  const double* tAy = tgt + P;
  const double* tBx = tgt + 2 * P;
  const double* tBy = tgt + 3 * P;
  const double* tI = tgt + (DUAL ? 4 : 2) * P;
  const int t = k - p.target_offset;
  double ex = dsub(pz.x, tAx[t]), ey = dsub(pz.y, tAy[t]);
  double e = dadd(dmul(ex, ex), dmul(ey, ey));
  double term = (wA == 1.0) ? e : dmul(wA, e);
  if (DUAL) {
    ex = dsub(pz.x, tBx[t]);
    ey = dsub(pz.y, tBy[t]);
    e = dadd(dmul(ex, ex), dmul(ey, ey));
    term = dadd(term, (wB == 1.0) ? e : dmul(wB, e));
  }
  if (IMU) {
    double d = remainder(dsub(pz.th, tI[t]), kTwoPi);
    term = dadd(term, dmul(p.w_imu, dmul(d, d)));
  }
  if (p.k_steer != 0.0) term = dadd(term, dmul(p.k_steer, dmul(s, s)));
  return term;
}

// (HS: the form of the prefix sums, warp_scan_add)
template <bool DUAL, bool IMU, bool HS = false>
__device__ double warp_cost64(const SearchParams& p, const WinInfo& wi, const double* tgt, int P,
                              int h, int lane, double wA, double wB, Pose<double>* first) {
  const int i = div_sh(h, p.gs, p.gs_sh), j = h - i * p.gs;
  GridCtl g{wi.v_seed, wi.s_seed, wi.dt, grid_rate(p.max_accel, i, p.gv),
            grid_rate(p.max_rate, j, p.gs), p.max_steer};
  Pose<double> carry{0.0, 0.0, 0.0};
  double J = 0.0;
  const int N = wi.n_steps;
  for (int base = 0; base < N; base += 32) {
    const int k = base + lane + 1;
    const bool active = k <= N;
    double v = 0.0, s = 0.0;
    if (active) g.at(k, &v, &s);
    Pose<double> pz = warp_model_round<double, 5, HS>(v, s, active, wi.dt, p.L, p.ratio, carry, lane);
    if (base == 0) {
      first->x = __shfl_sync(FULL, pz.x, 0);
      first->y = __shfl_sync(FULL, pz.y, 0);
      first->th = __shfl_sync(FULL, pz.th, 0);
    }
    const double term = active ? step_term<DUAL, IMU>(p, tgt, P, k, pz, s, wA, wB) : 0.0;
    J = dadd(J, warp_sum(term));
  }
  return J;
}

// ---- the same cost for several hypotheses that all stop within G = 2^LG steps, one warp -----------
// (the near-ties of a slow vehicle: a braking row, every steering rate).  Lane group c (G lanes; 32 / G
// groups: four hypotheses that stop within 8 steps, or two within 16) rolls hypothesis c through its
// first G steps -- one pass through tan / sincos for all of them -- with the scans confined to the
// group; from step G on the pose no longer changes.  Every value is bit-identical to warp_cost64's:
// the group scan IS the warp scan on lanes 0..G-1 (warp_scan_add), the lanes beyond the last moving
// step hold copies of its sums, a later round adds +0 to the carry, and the terms go through the same
// step_term and the same butterfly.  h_grp: the hypothesis of this lane's group, or -1 (computed like
// hypothesis 0 and ignored by the caller: without branches the butterflies of a round interleave).
// Precondition (checked by the caller): V_k = 0 for every k > G of every hypothesis given.
template <bool DUAL, bool IMU, int LG>
__device__ __forceinline__ void warp_cost64_pack(const SearchParams& p, const WinInfo& wi,
                                                 const double* tgt, int P, int h_grp, int lane,
                                                 double wA, double wB, double (&cost)[32 >> LG],
                                                 Pose<double> (&first)[32 >> LG]) {
  constexpr int G = 1 << LG, NC = 32 >> LG;
  const int N = wi.n_steps;
  const int kk = (lane & (G - 1)) + 1;
  const bool act = h_grp >= 0 && kk <= N;
  const int hh = h_grp >= 0 ? h_grp : 0;
  const int i = div_sh(hh, p.gs, p.gs_sh), j = hh - i * p.gs;
  GridCtl g{wi.v_seed, wi.s_seed, wi.dt, grid_rate(p.max_accel, i, p.gv),
            grid_rate(p.max_rate, j, p.gs), p.max_steer};
  double v = 0.0, s = 0.0;
  if (act) g.at(kk, &v, &s);
  Pose<double> carry{0.0, 0.0, 0.0};
  const Pose<double> pz = warp_model_round<double, LG>(v, s, act, wi.dt, p.L, p.ratio, carry, lane);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    first[c].x = __shfl_sync(FULL, pz.x, G * c);
    first[c].y = __shfl_sync(FULL, pz.y, G * c);
    first[c].th = __shfl_sync(FULL, pz.th, G * c);
    cost[c] = 0.0;
  }
  for (int base = 0; base < N; base += 32) {
    const int k = base + lane + 1;
    const bool active = k <= N;
    double term[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int src = G * c + (k < G ? k : G) - 1;
      Pose<double> q;
      q.x = __shfl_sync(FULL, pz.x, src);
      q.y = __shfl_sync(FULL, pz.y, src);
      q.th = __shfl_sync(FULL, pz.th, src);
      if (base > 0) {   // warp_cost64's later rounds: carry + (a scan of zeros)
        q.x = dadd(q.x, 0.0);
        q.y = dadd(q.y, 0.0);
        q.th = dadd(q.th, 0.0);
      }
      double sk = 0.0;
      if (p.k_steer != 0.0) {   // kernel-uniform
        const int hc = __shfl_sync(FULL, hh, G * c);
        const int jc = hc - div_sh(hc, p.gs, p.gs_sh) * p.gs;
        const GridCtl gc{wi.v_seed, wi.s_seed, wi.dt, 0.0, grid_rate(p.max_rate, jc, p.gs), p.max_steer};
        double v_unused;
        gc.at(k, &v_unused, &sk);
      }
      term[c] = active ? step_term<DUAL, IMU>(p, tgt, P, active ? k : 1, q, sk, wA, wB) : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < NC; ++c) term[c] = dadd(term[c], __shfl_xor_sync(FULL, term[c], o));
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) cost[c] = dadd(cost[c], term[c]);
  }
}

// ---- FP32 scan of one item: steering rate j, accelerations m0 .. m0+C-1 of the VD table ------
template <int C>
struct ScanOut {
  float J[C];
};

// Pruning (both scans).  The partial sums of a cost only grow (every term is a square, rounding is
// monotone), so a hypothesis whose partial cost has passed the candidate threshold T(U) of its item
// under the bound U the team held when the pass began can neither become a candidate of this pass
// (T(U') <= T(U) for the pass's own U' <= U) nor lower U (its J + err(J) > T(U) >= U): when that holds
// for every hypothesis of every lane the warp stops scanning, and the caller treats its items like
// items outside the grid.  Candidates, bounds and records are those of the exhaustive scan, bit for
// bit.  `Tq` is the threshold on the FIRST position term's accumulator (w_A * J_A is a lower bound of
// the cost; +inf: never), `every` the steps between two votes, `lanes` the lanes that scan.
// Returns true when the warp stopped early.
template <int C, bool DUAL, bool IMU>
__device__ __forceinline__ bool scan_item(int N, int gs, int vd_cols, int j, int m0,
                                          const float* __restrict__ TL,
                                          const float* __restrict__ VD,
                                          const float2* __restrict__ Df,
                                          const float2* __restrict__ Dab,
                                          const float* __restrict__ fI, float wA, float wB,
                                          float wI, float kJS, float Tq, int every, unsigned lanes,
                                          ScanOut<C>& out) {
  float th[C], ex[C], ey[C], JA[C], JB[C], JI[C];
#pragma unroll
  for (int c = 0; c < C; ++c) th[c] = ex[c] = ey[c] = JA[c] = JB[c] = JI[c] = 0.f;
  // running pointers: the table strides are run-time values, and an index multiply per load
  // would sit on the FMA pipe the loop is bound by
  const float* tl = TL + j;
  const float* vd = VD + m0;
  const float2* df = Df;
  const int tl_step = gs * 4, vd_step = vd_cols * 4;     // byte strides, computed once (not an IMAD per step)
  bool pruned = false;
  int k = 1;
#pragma unroll 1
  for (int kend = every;; kend += every) {
  const int ke = kend < N ? kend : N;
#pragma unroll 1
  for (; k <= ke; ++k, tl = reinterpret_cast<const float*>(reinterpret_cast<const char*>(tl) + tl_step),
           vd = reinterpret_cast<const float*>(reinterpret_cast<const char*>(vd) + vd_step)) {
    const float tlk = *tl;
    float v[C];
#pragma unroll
    for (int c4 = 0; c4 < C; c4 += 4) {
      const float4 q = *reinterpret_cast<const float4*>(vd + c4);
      v[c4] = q.x; v[c4 + 1] = q.y; v[c4 + 2] = q.z; v[c4 + 3] = q.w;
    }
    const float2 d = *++df;
    float2 dab = make_float2(0.f, 0.f);
    float ti = 0.f;
    if (DUAL) dab = Dab[k];
    if (IMU) ti = fI[k];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      th[c] = fmaf(v[c], tlk, th[c]);
      float sn, cs;
      __sincosf(th[c], &sn, &cs);
      ex[c] = fmaf(v[c], cs, ex[c] - d.x);   // error recurrence: e_k = e_{k-1} - dT_k + v*cos
      ey[c] = fmaf(v[c], sn, ey[c] - d.y);
      JA[c] = fmaf(ex[c], ex[c], JA[c]);
      JA[c] = fmaf(ey[c], ey[c], JA[c]);
      if (DUAL) {
        const float bx = ex[c] + dab.x, by = ey[c] + dab.y;
        JB[c] = fmaf(bx, bx, JB[c]);
        JB[c] = fmaf(by, by, JB[c]);
      }
      if (IMU) {
        float e = th[c] - ti;
        e = fmaf(-rintf(e * 0.15915494309189535f), 6.283185307179586f, e);
        JI[c] = fmaf(e, e, JI[c]);
      }
    }
  }
  if (k > N) break;
  float mn = JA[0];
#pragma unroll
  for (int c = 1; c < C; ++c) mn = fminf(mn, JA[c]);
  if (__all_sync(lanes, mn > Tq)) {
    pruned = true;
    break;
  }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float t = fmaf(wA, JA[c], kJS);
    if (DUAL) t = fmaf(wB, JB[c], t);
    if (IMU) t = fmaf(wI, JI[c], t);
    out.J[c] = t;
  }
  return pruned;
}

// ---- packed FP32x2 + rotation scan (C = 8, no IMU term, V_w >= 0) --------------------------------
// Blackwell's FFMA2 / FADD2 / FMUL2 process two hypotheses per instruction.  While a hypothesis
// is still moving its heading is affine in its acceleration, theta_k(i) = A_k(j) + a_i * B_k(j)
// (after V clamps to zero the step length is zero and the heading no longer matters), so the
// sin/cos of accelerations {0,1,4,5} of the chunk come from the SFU and those of {2,3,6,7} from one
// rotation by 2*da*B_k: 10 MUFU and ~64 FP32-pipe instructions per 8 hypothesis-steps.
__device__ __forceinline__ float2 pk(float a, float b) { return make_float2(a, b); }

__device__ __forceinline__ void rot_pair(float2 c2, float2 s2, float cr, float sr, float2& co,
                                         float2& so) {
  co = __ffma2_rn(c2, pk(cr, cr), __fmul2_rn(s2, pk(-sr, -sr)));
  so = __ffma2_rn(s2, pk(cr, cr), __fmul2_rn(c2, pk(sr, sr)));
}

template <bool DUAL>
__device__ __forceinline__ bool scan_item_fast(int N, int gs, int vd_cols, int j, int m0,
                                               const float* __restrict__ TL,
                                               const float* __restrict__ VD,
                                               const float2* __restrict__ Df,
                                               const float2* __restrict__ Dab, float wA, float wB,
                                               float kJS, float vwdt, float dt2, float a0, float da,
                                               float Tq, int every, unsigned lanes, ScanOut<8>& out) {
  // x and y cost terms accumulate separately (JA / JA2): one chain of dependent FFMA2 per
  // accumulator measured 7 % faster on the dense grid than a single merged one
  float2 ex[4], ey[4], JA[4], JA2[4], JB[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) ex[q] = ey[q] = JA[q] = JA2[q] = JB[q] = pk(0.f, 0.f);
  float A = 0.f, B = 0.f;
  const float* tl = TL + j;
  const float* vd = VD + m0;
  const float2* df = Df;
  const int tl_step = gs * 4, vd_step = vd_cols * 4;     // byte strides, computed once (not an IMAD per step)
  // sin / cos of the eight headings of step k (advances A, B)
  auto trig = [&](int k, float2 (&c2)[4], float2 (&s2)[4]) {
    const float tlk = *tl;
    tl = reinterpret_cast<const float*>(reinterpret_cast<const char*>(tl) + tl_step);
    const float kdt2 = (float)k * dt2;
    A = fmaf(vwdt, tlk, A);
    B = fmaf(kdt2, tlk, B);
    float s0, c0, s1, c1, s4, c4, s5, c5, sr, cr;
    __sincosf(fmaf(a0, B, A), &s0, &c0);
    __sincosf(fmaf(a0 + da, B, A), &s1, &c1);
    __sincosf(fmaf(fmaf(4.f, da, a0), B, A), &s4, &c4);
    __sincosf(fmaf(fmaf(5.f, da, a0), B, A), &s5, &c5);
    __sincosf((da + da) * B, &sr, &cr);
    c2[0] = pk(c0, c1);
    s2[0] = pk(s0, s1);
    c2[2] = pk(c4, c5);
    s2[2] = pk(s4, s5);
    rot_pair(c2[0], s2[0], cr, sr, c2[1], s2[1]);
    rot_pair(c2[2], s2[2], cr, sr, c2[3], s2[3]);
  };
  // positions and costs of step k from its sin / cos
  auto step = [&](int k, const float2 (&c2)[4], const float2 (&s2)[4]) {
    const float4 va = *reinterpret_cast<const float4*>(vd);
    const float4 vb = *reinterpret_cast<const float4*>(vd + 4);
    vd = reinterpret_cast<const float*>(reinterpret_cast<const char*>(vd) + vd_step);
    const float2 d = *++df;
    float2 dab = pk(0.f, 0.f);
    if (DUAL) dab = Dab[k];
    const float2 v2[4] = {pk(va.x, va.y), pk(va.z, va.w), pk(vb.x, vb.y), pk(vb.z, vb.w)};
    const float2 ndx = pk(-d.x, -d.x), ndy = pk(-d.y, -d.y);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ex[q] = __ffma2_rn(v2[q], c2[q], __fadd2_rn(ex[q], ndx));
      ey[q] = __ffma2_rn(v2[q], s2[q], __fadd2_rn(ey[q], ndy));
      // (scalar FFMAs for these eight accumulations -- 16 FFMA instead of 8 FFMA2, less FMA-pipe time,
      // eight more issue slots -- measured the same on the dense grid and 1.5 % slower on 32x32)
      JA[q] = __ffma2_rn(ex[q], ex[q], JA[q]);
      JA2[q] = __ffma2_rn(ey[q], ey[q], JA2[q]);
      if (DUAL) {
        const float2 bx = __fadd2_rn(ex[q], pk(dab.x, dab.x)), by = __fadd2_rn(ey[q], pk(dab.y, dab.y));
        JB[q] = __ffma2_rn(bx, bx, JB[q]);
        JB[q] = __ffma2_rn(by, by, JB[q]);
      }
    }
  };
  // groups of `every` steps with a pruning vote behind each (see scan_item); x and y terms accumulate
  // apart here, so the vote adds them first
  bool pruned = false;
  int k = 1;
#pragma unroll 1
  for (int kend = every;; kend += every) {
    const int ke = kend < N ? kend : N;
#pragma unroll 1
    for (; k <= ke; ++k) {
      float2 c2[4], s2[4];
      trig(k, c2, s2);
      step(k, c2, s2);
    }
    if (k > N) break;
    float2 m2 = __fadd2_rn(JA[0], JA2[0]);
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      const float2 t2 = __fadd2_rn(JA[q], JA2[q]);
      m2.x = fminf(m2.x, t2.x);
      m2.y = fminf(m2.y, t2.y);
    }
    if (__all_sync(lanes, fminf(m2.x, m2.y) > Tq)) {
      pruned = true;
      break;
    }
  }
  // (running the SFU one step ahead of the FMA pipe -- a two-stage software pipeline -- needs 16
  // more registers, spills, and measured 13 % slower)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    JA[q] = __fadd2_rn(JA[q], JA2[q]);
    float t0 = fmaf(wA, JA[q].x, kJS), t1 = fmaf(wA, JA[q].y, kJS);
    if (DUAL) {
      t0 = fmaf(wB, JB[q].x, t0);
      t1 = fmaf(wB, JB[q].y, t1);
    }
    out.J[2 * q] = t0;
    out.J[2 * q + 1] = t1;
  }
  return pruned;
}

// A window that is not searched leaves this record (one thread stores it to every destination).
static __device__ __forceinline__ void store_unsearched(vmvo_window_result* dst, int st, int n_steps, double vs,
                                              double ss) {
  vmvo_window_result r;
  r.best_idx = -1;
  r.n_steps = n_steps;
  r.status = st;
  r.n_rescored = 0;
  r.best_cost = CUDART_NAN;
  r.v_seed = vs;
  r.s_seed = ss;
  r.x1 = r.y1 = r.theta1 = CUDART_NAN;
  *dst = r;
}

// ---- the kernel -----------------------------------------------------------------------------
// WARPS warps per CTA, MINB CTAs per SM.  Only <8, 2> (128 registers per thread, 16 warps/SM) is
// instantiated: at 80 or 64 registers (24 / 32 warps/SM) both scan loops spill inside the loop and
// every such variant measured slower (profiles/README.md).
template <typename SF> struct PoseOf;
template <> struct PoseOf<float> { using type = float4; };
template <> struct PoseOf<double> { using type = double4; };

// SKIP: the grid needs more than two passes per window -- passes run middle-out and warps skip the
// band / candidate work of passes that cannot matter (a separate instantiation: on two-pass grids
// the extra code measured 2 % slower for nothing)
//
// MODE: 0 = every feature behind its run-time switch; 1 = LEAN: what the product's default path never
// uses is compiled out -- chained seeds, the steering penalty, rollout / control outputs, the debug
// export, grid sizes that are not powers of two, window preparation records; 2 = lean AND the windows
// always come as preparation records (phases A1-A3 compiled out; two-pass kernel only).  A 32x32
// window is ~3 000 straight-line instructions per warp, executed once each: a quarter of their stall
// samples waited for instruction fetch, and the specialised kernel (10.7k -> 7.5k SASS instructions)
// measured 10 % faster on config 2; the dense grids gain 2-3 %.
template <int C, int WARPS, int MINB, bool DUAL, bool IMU, typename SF, bool SKIP, int MODE>
__global__ void __launch_bounds__(32 * WARPS, MINB)
vmvo_window_search_kernel(const SearchParams p) {
  using Pose4 = typename PoseOf<SF>::type;
  constexpr int kC = C;
  constexpr bool kLean = MODE != 0;
  // x / d for the table sizes: a shift in the lean kernels (the host only picks them for powers of two)
  auto dv = [](int x, int d, int sh) { return kLean ? x >> sh : div_sh(x, d, sh); };
  extern __shared__ __align__(1024) unsigned char smem_cta[];
  const int P = p.maxp;
  const int n_streams = p.load_vo + p.load_gps;
  // phases A1-A3 were run by vmvo_window_prep_kernel (MODE 2: always; MODE 1: never)
  const bool use_prep = MODE == 2 ? true : MODE == 1 ? false : p.prep != nullptr;
  const PrepLayout prl(P, DUAL, IMU);
  const SmemLayout lay(P, p.gs, p.vd_cols, p.team_warps, n_streams, DUAL, IMU,
                       p.target_mode == VMVO_TARGET_TRAVERSE, (int)sizeof(Pose4), use_prep ? prl.total : 0);
  // stream s (0 = VO, 1 = GPS) lives in slot s when both are staged, else in slot 0
  const int slot_vo = 0, slot_gps = p.load_vo ? 1 : 0;
  const Team team{p.team_warps, p.team_warps * 32, (int)threadIdx.x >> p.t_sh};
  unsigned char* smem = smem_cta + (size_t)team.id * lay.total;
  SmemHeader* hd = reinterpret_cast<SmemHeader*>(smem);
  Pose4* raw = reinterpret_cast<Pose4*>(smem + lay.off_raw);
  const Pose4* g_vo = reinterpret_cast<const Pose4*>(p.vo);
  const Pose4* g_gps = reinterpret_cast<const Pose4*>(p.gps);
  const SF* g_imu = reinterpret_cast<const SF*>(p.imu);
  double* loc = reinterpret_cast<double*>(smem + lay.off_loc);
  double* loci = reinterpret_cast<double*>(smem + lay.off_loci);
  double* tgt = reinterpret_cast<double*>(smem + lay.off_tgt);
  float2* Df = reinterpret_cast<float2*>(smem + lay.off_df);
  float2* Dab = reinterpret_cast<float2*>(smem + lay.off_dab);
  float* fI = reinterpret_cast<float*>(smem + lay.off_fi);
  int* keep = reinterpret_cast<int*>(smem + lay.off_keep);
  float* TL = reinterpret_cast<float*>(smem + lay.off_tl);
  float* JS = reinterpret_cast<float*>(smem + lay.off_js);
  float* TS = reinterpret_cast<float*>(smem + lay.off_ts);
  float* TSP = reinterpret_cast<float*>(smem + lay.off_tsp);
  const int gs4 = (p.gs + 3) & ~3;
  float* VD = reinterpret_cast<float*>(smem + lay.off_vd);
  uint2* cand = reinterpret_cast<uint2*>(smem + lay.off_cand);

  const int tid = threadIdx.x - team.id * team.threads;   // thread index within the team
  const int lane = tid & 31, warp = tid >> 5;              // warp index within the team
  const int T = team.threads, NW = team.warps;
  const int cand_cap = p.cand_cap;

  // A = first position stream with a weight, B = the second one (DUAL only)
  const int sA = p.use_vo ? 0 : 1;
  const double wA64 = p.use_vo ? p.w_vo : p.w_gps;
  const double wB64 = p.w_gps;
  const float wA = (float)wA64, wB = (float)wB64, wI = (float)p.w_imu;
  const bool ksteer = kLean ? false : p.k_steer != 0.0;
  const bool dbg = kLean ? false : p.dbg_cost != nullptr;
  const double kd = kDegToRad / p.ratio;   // steering-wheel degrees -> road-wheel radians

  long long q_item = 0;      // fetcher thread: the queue item behind the window just popped
  auto issue_load = [&](long long w, int buf) {  // the fetcher thread only
    if (use_prep) {          // the window's preparation record instead of its poses
      mbar_arrive_expect_tx(&hd->mbar[buf], (unsigned)prl.total);
      bulk_g2s(smem + lay.off_raw + buf * lay.raw_buf, p.prep + (size_t)q_item * p.prep_stride,
               (unsigned)prl.total, &hd->mbar[buf]);
      return;
    }
    long long start;
    int len, drv;
    if (p.win_start) {
      start = p.win_start[w];
      len = p.win_len[w];
      drv = p.win_drive[w];
    } else {
      // frames mode without a plan: window i of drive d covers poses i .. min(i + W, n_d - 1)
      // (optimize_trajectory_v2.py:48-56 with a frame horizon; what plan_windows_kernel writes)
      int lo = 0, hi = p.n_drives;           // largest d with win_off[d] <= w
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.win_off[mid] <= w) lo = mid; else hi = mid;
      }
      drv = lo;
      const long long i = w - p.win_off[drv], f0 = p.drive_off[drv], n = p.drive_off[drv + 1] - f0;
      long long e = i + p.window_frames + 1;
      e = e < n ? e : n;
      start = f0 + i;
      len = (int)(e - i);
    }
    hd->wstart[buf] = start;
    hd->wlen[buf] = len;
    hd->wdt[buf] = p.dt_drive[drv];
    len = len < P ? len : P;
    len = len > 0 ? len : 0;
    const unsigned bytes = (unsigned)len * (unsigned)sizeof(Pose4);
    const unsigned total = bytes * (unsigned)(p.load_vo + p.load_gps);
    mbar_arrive_expect_tx(&hd->mbar[buf], total);
    if (bytes) {
      if (p.load_vo)
        bulk_g2s(raw + (buf * n_streams + slot_vo) * P, g_vo + start, bytes, &hd->mbar[buf]);
      if (p.load_gps)
        bulk_g2s(raw + (buf * n_streams + slot_gps) * P, g_gps + start, bytes, &hd->mbar[buf]);
    }
  };

  // One record, to this GPU's buffer and to every mirror (peer GPUs' gather buffers: plain stores
  // over NVLink, visible to the peers when the kernel has completed).  One thread assembles it in
  // shared memory (store_record); warp 0 stores it (flush_record): four lanes per destination, so
  // that each copy leaves as ONE 64-byte write -- with seven peers a single thread would issue 32
  // separate 16-byte stores per window, 28 of them small NVLink packets.
  // Without mirrors (one GPU) the assembling thread stores the record itself.
  auto store_record = [&](long long w, const vmvo_window_result& r) {
    if (p.n_mirrors == 0) p.results[w] = r;
    else hd->rec = r;
  };
  auto flush_record = [&](long long w) {   // every lane of the team's warp 0
    if (p.n_mirrors == 0) return;
    __syncwarp();
    const uint4 part = reinterpret_cast<const uint4*>(&hd->rec)[lane & 3];
    for (int t = lane >> 2; t <= p.n_mirrors; t += 8) {
      vmvo_window_result* base = t == 0 ? p.results : p.mirrors[t - 1];
      reinterpret_cast<uint4*>(base + w)[lane & 3] = part;
    }
    // (no fence here: a fence behind stores to peer memory would wait for their acknowledgement
    // over NVLink, window after window)
    __syncwarp();
  };

  const bool chained = kLean ? false : p.run_offsets != nullptr;
  long long run_end = 0;     // fetcher thread: end of the run being walked (chained mode)
  double s_chain = 0.0;      // all threads: steering seed handed from window to window
  // fetcher thread: the window after `w` -- the next one of the same run, else a fresh queue item
  auto next_window = [&](long long w, int slot) -> long long {
    if (chained && w >= 0 && w + 1 < run_end) {
      hd->first[slot] = 0;
      return w + 1;
    }
    hd->first[slot] = 1;
    const bool dealt = p.sh_world > 1;
    for (;;) {
      long long r = (long long)atomicAdd(p.work_counter, 1ULL);
      q_item = r;
      if (!chained) {
        if (!dealt || p.sh_block_sh < 0) return r < p.n_local ? r : p.n_windows;
        if (r >= p.n_local) return p.n_windows;
        const long long b = r >> p.sh_block_sh;
        const long long w = (((b * p.sh_world + p.sh_rank) << p.sh_block_sh)) + (r - (b << p.sh_block_sh));
        return w < p.n_windows ? w : p.n_windows;    // (w grows with r: nothing valid follows)
      }
      if (dealt) r = r * p.sh_world + p.sh_rank;
      if (r >= p.n_runs) return p.n_windows;
      run_end = p.run_offsets[r + 1];
      if (p.run_offsets[r] < run_end) return p.run_offsets[r];
    }
  };

  // the queue pop and the TMA issue belong to lane 0 of the team's LAST warp: phases A1 / A2 keep
  // warp 0 busy, so in a multi-warp team the atomic and its dependent loads overlap with them
  const bool fetcher = tid == T - 32;
  // a sharded search advances the exchange's step counter (nothing reads it while a search runs)
  if (p.epoch && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.epoch, 1u);
  // the CTAs of the second kernel may take over SMs as teams run out of windows
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (fetcher) {
    mbar_init(&hd->mbar[0], 1);
    mbar_init(&hd->mbar[1], 1);
    mbar_fence_init();
    hd->count = 0;
    // with a preparation pass this kernel is its programmatic dependent: everything above ran beside
    // the pass's last wave; the records are complete from here on
    if (use_prep) asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long w = next_window(-1, 0);
    hd->wid[0] = w;
    if (w < p.n_windows) issue_load(w, 0);
  }
  team.sync();

  for (int it = 0;; ++it) {
    const int cur = it & 1;
    const long long w = hd->wid[cur];
    if (w >= p.n_windows) {
      // this team parks no more windows: the second kernel watches the count of finished teams to
      // learn when the set of parked windows is final (its ready words are behind fences already)
      if (tid == 0 && p.defer_ready) {
        __threadfence();
        atomicAdd(p.windows_done, 1ULL);
      }
      break;
    }
    if (fetcher) {  // prefetch the next window's poses while this one is searched
      const long long wn = next_window(w, cur ^ 1);
      hd->wid[cur ^ 1] = wn;
      if (wn < p.n_windows) issue_load(wn, cur ^ 1);
    }
    if (chained && hd->first[cur]) s_chain = 0.0;   // optimize_trajectory_v2.py:46
    const long long start = hd->wstart[cur];
    mbar_wait(&hd->mbar[cur], (unsigned)((it >> 1) & 1));
    // with a preparation pass the staged bytes are the window's record: header, FP32 increments,
    // float64 targets -- used in place
    const PrepHdr* ph = reinterpret_cast<const PrepHdr*>(smem + lay.off_raw + cur * lay.raw_buf);
    const int len = use_prep ? ph->len : hd->wlen[cur];
    const double dt = use_prep ? ph->dt : hd->wdt[cur];
    if (use_prep) {
      unsigned char* rec = smem + lay.off_raw + cur * lay.raw_buf;
      Df = reinterpret_cast<float2*>(rec + prl.off_df);
      Dab = reinterpret_cast<float2*>(rec + prl.off_dab);
      fI = reinterpret_cast<float*>(rec + prl.off_fi);
      tgt = reinterpret_cast<double*>(rec + prl.off_tgt);
    }

    // the record of a window that is not searched (too long / empty): everything but the status,
    // the step count and the seeds stays "none" (one thread, every destination: store_unsearched)
    auto write_unsearched = [&](int st, int n_steps, double vs, double ss) {
      for (int t = 0; t <= p.n_mirrors; ++t)
        store_unsearched((t == 0 ? p.results : p.mirrors[t - 1]) + w, st, n_steps, vs, ss);
    };

    if (len > P || len < 1) {  // uniform branch
      // no pose at all (an empty time extent, NaN stamps): "No frames found", schema.py:122
      if (tid == 0)
        write_unsearched(len < 1 ? (VMVO_WIN_EMPTY | VMVO_WIN_NO_FRAMES) : VMVO_WIN_TOO_LONG, 0,
                         CUDART_NAN, CUDART_NAN);
      team.sync();
      continue;
    }

    const int slot_prim = p.primary == VMVO_PRIMARY_VO ? slot_vo : slot_gps;
    const Pose4* rp = raw + (cur * n_streams + slot_prim) * P;
    if (use_prep) {
      if (tid == 0) {
        hd->wi.v_seed = ph->v_seed;
        hd->wi.s_seed = ph->s_seed;
        hd->wi.dt = ph->dt;
        hd->wi.n_targets = ph->n_targets;
        hd->wi.n_steps = ph->n_steps;
        hd->wi.n_dead = ph->n_dead;
        hd->wi.sat_lo = ph->sat_lo;
        hd->wi.sat_hi = ph->sat_hi;
      }
    } else {
    if (warp == (NW > 1 ? 1 : 0)) {
      // rows that never move: a_i <= 0 and V_w + a_i*t_1 <= 0 (a_i grows with i: a prefix).  On a
      // second warp when the team has one: warp 0 has the seed's division and atan to wait for.
      const double v_seed = p.seed_mode == VMVO_SEED_GIVEN
                                ? p.seeds[2 * w]
                                : dmul(dadd((double)rp[0].w, (double)rp[len - 1].w), 0.5);
      int n_dead = 0;
      for (int i0 = 0; i0 < p.gv; i0 += 32) {
        const int i = i0 + lane;
        bool dead = false;
        if (i < p.gv) {
          const double a = grid_rate(p.max_accel, i, p.gv);
          dead = a <= 0.0 && !(dadd(v_seed, dmul(a, dmul(1.0, dt))) > 0.0) && v_seed == v_seed;
        }
        const unsigned b = __ballot_sync(FULL, dead);
        n_dead += __popc(b);
        if (b != FULL) break;
      }
      if (lane == 0) hd->wi.n_dead = n_dead;
    }

    // ---- phase A1: local frames (a7), one pose per thread ---------------------------------
    if (tid < ((len + 31) & ~31)) {  // only the warps that own poses pay for sincos
      for (int s = 0; s < 2; ++s) {
        if (!(s == 0 ? p.load_vo : p.load_gps)) continue;
        const int slot = s == 0 ? slot_vo : slot_gps;
        const Pose4* rs = raw + (cur * n_streams + slot) * P;
        const Pose4 p0 = rs[0];
        const double th0 = (double)p0.z;
        double sn, cs;
        sincos(th0, &sn, &cs);
        double* lx = loc + (slot * 3 + 0) * P;
        double* ly = loc + (slot * 3 + 1) * P;
        double* lt = loc + (slot * 3 + 2) * P;
        for (int m = tid; m < len; m += T) {
          const Pose4 q = rs[m];
          const double dx = dsub((double)q.x, (double)p0.x);
          const double dy = dsub((double)q.y, (double)p0.y);
          lx[m] = dadd(dmul(dx, cs), dmul(dy, sn));
          ly[m] = dadd(dmul(-dx, sn), dmul(dy, cs));
          lt[m] = dsub((double)q.z, th0);
        }
      }
    }
    if (IMU) {
      const double y0 = (double)g_imu[start];
      for (int m = tid; m < len; m += T) loci[m] = dsub((double)g_imu[start + m], y0);
    }
    team.sync();

    // ---- phase A2 (warp 0): seeds, decimation (a9) ----------------------------------------
    const double* plx = loc + (slot_prim * 3 + 0) * P;
    const double* ply = loc + (slot_prim * 3 + 1) * P;
    const double* plt = loc + (slot_prim * 3 + 2) * P;
    if (warp == 0) {
      double v_seed, s_seed;
      if (p.seed_mode == VMVO_SEED_GIVEN) {
        v_seed = p.seeds[2 * w];
        s_seed = p.seeds[2 * w + 1];
      } else {
        v_seed = dmul(dadd((double)rp[0].w, (double)rp[len - 1].w), 0.5);   // (/ 2, exactly)
        s_seed = 0.0;
        if (len >= 2 && dmul(v_seed, dt) > 1e-6) {
          // IEEE remainder by 2 pi; below pi in magnitude it is the argument itself, exactly
          const double dth0 = dsub(plt[1], plt[0]);
          const double dth = fabs(dth0) < kPi ? dth0 : remainder(dth0, kTwoPi);
          const double ang = atan(ddiv(dmul(p.L, dth), dmul(v_seed, dt)));
          s_seed = dmul(dmul(ang, kRadToDeg), p.ratio);
          s_seed = s_seed < -p.max_steer ? -p.max_steer : s_seed;
          s_seed = s_seed > p.max_steer ? p.max_steer : s_seed;
        }
      }
      if (chained) s_seed = s_chain;
      int n_targets = len;
      if (p.target_mode == VMVO_TARGET_TRAVERSE) {
        if (lane == 0) {  // sequential by definition (distance accumulator with reset)
          const double D = dmul(v_seed, dt);
          int cnt = 1;
          keep[0] = 0;
          double dist = 0.0;
          for (int i = 1; i < len; ++i) {
            const double ddx = dsub(plx[i], plx[i - 1]), ddy = dsub(ply[i], ply[i - 1]);
            const double seg = sqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy)));
            if (dadd(dist, seg) > D) {
              keep[cnt++] = i - 1;
              dist = seg;
            } else {
              dist = dadd(dist, seg);
            }
          }
          n_targets = cnt;
        }
        n_targets = __shfl_sync(FULL, n_targets, 0);
      }
      // steering rates that clamp to the seed's bound for every step
      int sat_lo = 0, sat_hi = -1;
      if (s_seed == p.max_steer || s_seed == -p.max_steer) {
        const bool hi = s_seed > 0;
        int first = p.gs, last = -1;
        for (int j0 = 0; j0 < p.gs; j0 += 32) {
          const int j = j0 + lane;
          bool in = false;
          if (j < p.gs) {
            const double r = grid_rate(p.max_rate, j, p.gs);
            in = hi ? (r >= 0.0) : (r <= 0.0);
          }
          const unsigned b = __ballot_sync(FULL, in);
          if (b) {
            first = min(first, j0 + __ffs(b) - 1);
            last = max(last, j0 + 31 - __clz(b));
          }
        }
        sat_lo = first;
        sat_hi = last;
      }
      if (lane == 0) {
        hd->wi.v_seed = v_seed;
        hd->wi.s_seed = s_seed;
        hd->wi.dt = dt;
        hd->wi.n_targets = n_targets;
        hd->wi.n_steps = n_targets > 1 ? n_targets - 1 : 0;
        hd->wi.sat_lo = sat_lo;
        hd->wi.sat_hi = sat_hi;
      }
    }
    }
    team.sync();

    const int n_targets = hd->wi.n_targets;
    const int N = hd->wi.n_steps;
    const double v_seed = hd->wi.v_seed, s_seed = hd->wi.s_seed;
    const int off = p.target_offset;
    const bool traverse = p.target_mode == VMVO_TARGET_TRAVERSE;

    // ---- phase A3: targets in float64, FP32 increments for the scan, finiteness -------------
    bool finite = isfinite(v_seed) && isfinite(s_seed) && isfinite(dt);
    float dmax = 0.f, dabmax = 0.f, imax = 0.f;
    if (use_prep) {      // (the record's maxima already carry a non-finite input as dmax = +inf)
      finite = true;
      dmax = ph->dmax;
      dabmax = ph->dabmax;
      imax = ph->imax;
    } else {
      const int slot_a = sA == 0 ? slot_vo : slot_gps;
      const double* aX = loc + (slot_a * 3 + 0) * P;
      const double* aY = loc + (slot_a * 3 + 1) * P;
      const double* bX = loc + (slot_gps * 3 + 0) * P;   // B is GPS (DUAL only)
      const double* bY = loc + (slot_gps * 3 + 1) * P;
      for (int q = tid; q < n_targets; q += T) {
        const int m = traverse ? keep[q] : q;
        const double ax = aX[m], ay = aY[m];
        tgt[q] = ax;
        tgt[P + q] = ay;
        finite = finite && isfinite(ax) && isfinite(ay);
        if (DUAL) {
          const double bx = bX[m], by = bY[m];
          tgt[2 * P + q] = bx;
          tgt[3 * P + q] = by;
          finite = finite && isfinite(bx) && isfinite(by);
        }
        if (IMU) {
          const double yi = loci[m];
          tgt[(DUAL ? 4 : 2) * P + q] = yi;
          finite = finite && isfinite(yi);
        }
        // step k = q + off compares against target q; its increment needs target q - 1
        const int k = q + off;
        if (k >= 1 && k <= N) {
          const int mp = (q >= 1) ? (traverse ? keep[q - 1] : q - 1) : -1;
          const double px = (k >= 2) ? aX[mp] : 0.0, py = (k >= 2) ? aY[mp] : 0.0;
          const float dx = (float)dsub(ax, px), dy = (float)dsub(ay, py);
          Df[k] = make_float2(dx, dy);
          dmax = fmaxf(dmax, fmaxf(fabsf(dx), fabsf(dy)));
          if (DUAL) {
            const float ex = (float)dsub(ax, bX[m]), ey = (float)dsub(ay, bY[m]);
            Dab[k] = make_float2(ex, ey);
            dabmax = fmaxf(dabmax, fmaxf(fabsf(ex), fabsf(ey)));
          }
          if (IMU) {
            const float yi = (float)loci[m];
            fI[k] = yi;
            imax = fmaxf(imax, fabsf(yi));
          }
        }
      }
    }
    // ---- phase A4: TL[k][j] = tan(delta_k(j)) / L and the steering penalty per j -------------
    if (N > 0) {
      // one steering rate per thread (one division), steps strided over the threads that share it
      // Four steps at a time: the entries are independent chains (three float64 operations, the
      // conversion, a ten-term Horner polynomial), and a team of two warps is latency-bound here.
      // The per-rate statistics of the band (max |TL|, sum |TL|, sum k |TL|) are taken on the way;
      // the threads that share a rate leave their parts in TSP, added up after the barrier.
      const float invL = p.inv_L;
      const int kpar = p.tl_slices;
      const bool poly = p.delta_max <= 0.6199;     // every clamped angle is inside the polynomial's range
      const bool noclamp = fabs(s_seed) + p.max_rate * ((double)N * fabs(dt)) * 1.000001 <= p.max_steer;
      for (int c = tid; c < p.gs * kpar; c += T) {
        const int ks = dv(c, p.gs, p.gs_sh), j = c - ks * p.gs;
        // r_j by multiplication: its last float64 bit is far below the float rounding of the angle
        const double rdt = p.rate_step * (double)(2 * j - (p.gs - 1)) * dt;
        const double c0 = s_seed * kd, ck = rdt * kd, kstep = (double)kpar;
        const float kstepf = (float)kpar;
        float mx = 0.f, s0 = 0.f, s1 = 0.f;
        // the common case -- no rate reaches the steering limit inside the window, every angle inside
        // the polynomial's range -- four steps at a time (independent chains); otherwise one entry
        // at a time through the clamps and tan_steer (one copy of that code: it is rarely run, and
        // what the kernel's code weighs is paid for in instruction fetches by every window)
        auto group = [&](int k0, const bool full) {     // full: all four steps inside the window
          float x[4], tl[4];
          const double kq = (double)k0;
          const float kf = (float)k0;
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = (float)fma(ck, fma((double)u, kstep, kq), c0);   // k = k0 + u * kpar
#pragma unroll
          for (int u = 0; u < 4; ++u) tl[u] = tan_poly(x[u]) * invL;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = k0 + u * kpar;
            if (full || k <= N) {
              TL[(k - 1) * p.gs + j] = tl[u];
              const float a = fabsf(tl[u]);
              mx = fmaxf(mx, a);
              s0 += a;
              s1 = fmaf(fmaf((float)u, kstepf, kf), a, s1);
            }
          }
        };
        if (noclamp && poly) {     // window-uniform
          for (int k0 = 1 + ks; k0 <= N; k0 += 4 * kpar) {     // (the lambda is specialised per call)
            if (k0 + 3 * kpar <= N) group(k0, true);
            else group(k0, false);
          }
        } else {
#pragma unroll 1
          for (int k = 1 + ks; k <= N; k += kpar) {
            float x;
            if (noclamp) {
              x = (float)fma(ck, (double)k, c0);
            } else {
              double sd = dadd(s_seed, dmul(rdt, (double)k));
              sd = sd < -p.max_steer ? -p.max_steer : sd;
              sd = sd > p.max_steer ? p.max_steer : sd;
              x = (float)(sd * kd);
            }
            const float tl = (poly ? tan_poly(x) : tan_steer(x)) * invL;
            TL[(k - 1) * p.gs + j] = tl;
            const float a = fabsf(tl);
            mx = fmaxf(mx, a);
            s0 += a;
            s1 = fmaf((float)k, a, s1);
          }
        }
        float* tp = TSP + 3 * gs4 * ks;
        tp[j] = mx;
        tp[gs4 + j] = s0;
        tp[2 * gs4 + j] = s1;
      }
      if (ksteer) {
        for (int j = tid; j < p.gs; j += T) {
          const double r = grid_rate(p.max_rate, j, p.gs);
          double acc = 0.0;
          for (int k = 1; k <= N; ++k) {
            double s = dadd(s_seed, dmul(r, dmul((double)k, dt)));
            s = s < -p.max_steer ? -p.max_steer : s;
            s = s > p.max_steer ? p.max_steer : s;
            acc += s * s;
          }
          JS[j] = (float)(p.k_steer * acc);
        }
      }
    }
    {
      // block max of the band inputs
      // (non-negative floats order like their bit patterns; a non-finite input anywhere in the
      // window travels as dmax = +inf, which no finite increment can reach)
      if (!finite) dmax = CUDART_INF_F;
      dmax = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(dmax)));
      if (DUAL) dabmax = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(dabmax)));
      if (IMU) imax = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(imax)));
      if (lane == 0) {
        hd->red[warp] = dmax;
        hd->red[kMaxWarps + warp] = dabmax;
        hd->red[2 * kMaxWarps + warp] = imax;
      }
    }
    team.sync();   // red[] complete
    dmax = dabmax = imax = 0.f;
    for (int q = 0; q < NW; ++q) {
      dmax = fmaxf(dmax, hd->red[q]);
      if (DUAL) dabmax = fmaxf(dabmax, hd->red[kMaxWarps + q]);
      if (IMU) imax = fmaxf(imax, hd->red[2 * kMaxWarps + q]);
    }
    const bool bad = !(dmax < CUDART_INF_F);
    const int status = (N <= 0 ? VMVO_WIN_EMPTY : 0) | (bad ? VMVO_WIN_NONFINITE : 0);
    // team-uniform per-window state stays in the shared-memory header (hd->wi, hd->bw, the warps'
    // running best): registers are for the scan
    const WinInfo& wi = hd->wi;

    if (status & VMVO_WIN_EMPTY) {
      if (tid == 0) write_unsearched(status, N, v_seed, s_seed);
      team.sync();
      continue;
    }

    // per steering rate: max |TL|, sum |TL|, sum k |TL| over the steps (the items' band inputs)
    float tmx = 0.f, ts0 = 0.f, ts1 = 0.f;
    const int tl_slices = p.tl_slices;
    for (int j = tid; j < p.gs; j += T) {
      float mx = 0.f, s0 = 0.f, s1 = 0.f;
      const float* t = TSP + j;
      for (int q = 0; q < tl_slices; ++q, t += 3 * gs4) {
        mx = fmaxf(mx, t[0]);
        s0 += t[gs4];
        s1 += t[2 * gs4];
      }
      TS[j] = mx;
      TS[gs4 + j] = s0;
      TS[2 * gs4 + j] = s1;
      tmx = fmaxf(tmx, mx);
      ts0 = fmaxf(ts0, s0);
      ts1 = fmaxf(ts1, s1);
    }
    if constexpr (SKIP) {   // (non-negative floats are ordered like their bit patterns)
      tmx = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(tmx)));
      ts0 = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(ts0)));
      ts1 = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(ts1)));
      if (lane == 0) {
        hd->ts[warp] = tmx;
        hd->ts[8 + warp] = ts0;
        hd->ts[16 + warp] = ts1;
      }
    }

    // this warp's running best (float64 cost, index, first pose) and its re-score count
    if (lane == 0) {
      const bool nonfinite = (status & VMVO_WIN_NONFINITE) != 0;
      // every hypothesis of a non-finite window costs NaN or Inf alike: np.argmin returns index 0
      hd->bh[warp] = nonfinite ? 0 : -1;
      hd->bcost[warp] = nonfinite ? CUDART_NAN : CUDART_INF;
      hd->bpose[warp][0] = hd->bpose[warp][1] = hd->bpose[warp][2] = CUDART_NAN;
      hd->nres[warp] = 0;
    }
    if (tid == 0) {
      hd->wi.status = status;
      BandWin bw;
      const float u = 5.9604644775390625e-8f;
      const float n = (float)N;
      bw.n = n;
      bw.s2 = n * (n + 1.f) * (2.f * n + 1.f) * (1.f / 6.f);
      const float n2 = n + 2.f;
      bw.s4 = 0.05f * n2 * n2 * n2 * n2 * n2;
      bw.dmax = dmax;
      bw.dabmax = dabmax;
      bw.imax = imax;
      // relative error of a TL entry: tan_steer <= 1 ulp = 2u below 0.62 rad (tanf: 4 ulp = 8u
      // above), the float rounding of the angle amplified by tan's condition number kappa, the
      // product with 1/L and its own rounding
      bw.eps_tl = ((p.delta_max <= (double)kTanPolyMax ? 4.f : 10.f) + (float)p.kappa) * u;
      bw.wpos = wA + (DUAL ? wB : 0.f);
      bw.wimu = IMU ? wI : 0.f;
      // J-proportional part.  The error recurrence rounds twice per step, each <= u*|e_m| (+ u*vmax,
      // carried by q1); by Cauchy-Schwarz sum_{m<=k} |e_m| <= sqrt(k) * sqrt(J_A), so the pose error
      // from these roundings after k steps is <= 2u*sqrt(k)*sqrt(J_A), and over both axes and the
      // three-term split sum_k (.)^2 <= S1*u^2*J_A with S1 = 6 * 4 * sum k = 12 N (N+1).  With two
      // position terms the recurrence state is A's error: J_A <= J / w_A, and
      // w_A*J_A + w_B*sqrt(J_A*J_B) <= (1 + w_B/(2 w_A)) * J.  Last: rounding of the cost sums.
      const float S1 = 12.f * n * (n + 1.f);
      const float fac = DUAL ? 1.f + wB / (2.f * wA) : 1.f;
      bw.c2 = 2.0f * (fac * (2.f * u * sqrtf(S1) + u * u * S1) + 2.f * (n + 2.f) * u + 16.f * u);
      hd->bw = bw;
    }
    __syncwarp();

    bool deferred = false;     // the float64 re-scores of this window go to the second kernel
    if (!(status & VMVO_WIN_NONFINITE)) {
      // ---- phase B: FP32 scan of the whole grid, candidates within the error band -------
      // (hd->bw and hd->wi.status become visible with the barrier that follows the first VD fill)
      float U = CUDART_INF_F;          // upper bound on the true minimum cost
      float Uw = CUDART_INF_F;         // this warp's tightened copy (after float64 re-scores)

      auto process_list = [&]() {
        const int count = hd->count < cand_cap ? hd->count : cand_cap;
        for (int e = warp; e < count; e += NW) {
          const uint2 ce = cand[e];
          if (__uint_as_float(ce.y) > fminf(U, Uw)) continue;  // warp-uniform; NaN stays in
          const int h = (int)ce.x;
          Pose<double> first;
          const double c64 = warp_cost64<DUAL, IMU, SKIP>(p, wi, tgt, P, h, lane, wA64, wB64, &first);
          const int best_h = hd->bh[warp];
          const double best_cost = hd->bcost[warp];
          __syncwarp();
          if (lane == 0) {
            ++hd->nres[warp];
            if (best_h < 0 || c64 < best_cost || (c64 == best_cost && h < best_h)) {
              hd->bh[warp] = h;
              hd->bcost[warp] = c64;
              hd->bpose[warp][0] = first.x;
              hd->bpose[warp][1] = first.y;
              hd->bpose[warp][2] = first.th;
            }
          }
          __syncwarp();
          // a float64 cost is itself an upper bound on the minimum (rounded up to float)
          Uw = fminf(Uw, __double2float_ru(c64));
        }
        team.sync();
        if (tid == 0) hd->count = 0;
        team.sync();
      };

      const bool fast_w = (C == 8 && !IMU) && p.allow_fast && v_seed >= 0.0;   // window-uniform
      const int n_pass = p.n_pass;
      const bool vd_full = p.vd_cols >= p.n_ic * kC;
      // With more than two passes: (i) the passes run middle-out over the accelerations -- the
      // optimum usually sits near a = 0, so U is tight after the first pass and few candidates are
      // listed; (ii) a warp whose smallest cost of a pass lies beyond the reach of the window's
      // WIDEST band around U skips its band and candidate work (it still meets the barriers).
      constexpr bool use_skip = SKIP;
      Band loose{0.f, 0.f, 0.f};
      // ... starting from the pass that holds the acceleration the window itself suggests: the chord to
      // the target of the last step against V_w * t (any order gives the same records; this one has U
      // tight after the first pass also when the vehicle speeds up or brakes, which is what the
      // pruning votes of all later passes work with)
      int p_lo = 0, p_hi = 0;
      bool p_up = true;
      if (use_skip) {
        int p_est = (n_pass - 1) >> 1;
        const int tl_ = N - off;
        if (tl_ >= 0 && tl_ < n_targets && p.gv > 1) {
          const double tx = tgt[tl_], ty = tgt[P + tl_], tt = (double)N * dt;
          const double chord = copysign(sqrt(tx * tx + ty * ty), tx);
          const double a_est = 2.0 * (chord - v_seed * tt) / (tt * tt);
          if (fabs(a_est) <= 1e6) {       // (NaN / Inf: keep the middle)
            double fi = (a_est + p.max_accel) * (0.5 / p.max_accel) * (double)(p.gv - 1);
            fi = fi < 0.0 ? 0.0 : fi;
            fi = fi > (double)(p.gv - 1) ? (double)(p.gv - 1) : fi;
            const long long item = (long long)((int)fi / kC) * p.gs + (p.gs >> 1);
            p_est = (int)(item / T);
            p_est = p_est < n_pass ? p_est : n_pass - 1;
          }
        }
        p_lo = p_est - 1;
        p_hi = p_est;
      }
      for (int pidx = 0; pidx < n_pass; ++pidx) {
        int pass = pidx;
        if (use_skip) {      // p_est, then alternately above and below it while either side lasts
          const bool take_hi = (p_up && p_hi < n_pass) || p_lo < 0;
          pass = take_hi ? p_hi++ : p_lo--;
          p_up = !p_up;
        }
        // VD[k][m] = V_k(ic0*C + m) * dt for the accelerations this pass touches
        // (a table that covers every acceleration is filled once per window: vd_full)
        const int ic0 = vd_full ? 0 : dv(pass * T, p.gs, p.gs_sh);
        if (!vd_full || pidx == 0) {
          const int kpar = p.vd_kpar;
          // a_i by multiplication (the float64 re-score uses the spec's division; here the last
          // bit is far below FP32 resolution)
          const double inv_a = p.acc_step;
          for (int c = tid; c < p.vd_cols * kpar; c += T) {
            const int kv = dv(c, p.vd_cols, p.vd_sh), m = c - kv * p.vd_cols;
            int i = ic0 * kC + m;
            i = i < p.gv ? i : p.gv - 1;
            const double adt = inv_a * (double)(2 * i - (p.gv - 1)) * dt;
#pragma unroll 4
            for (int k = 1 + kv; k <= N; k += kpar) {
              const double vv = fma(adt, (double)k, v_seed);
              VD[(k - 1) * p.vd_cols + m] = (float)((vv > 0.0 ? vv : 0.0) * dt);
            }
          }
          team.sync();
        }
        if (pidx == 0 && use_skip) {   // hd->bw and hd->ts are visible now: the widest band
          float tl_all = 0.f, s0_all = 0.f, s1_all = 0.f;
          for (int qq = 0; qq < NW; ++qq) {
            tl_all = fmaxf(tl_all, hd->ts[qq]);
            s0_all = fmaxf(s0_all, hd->ts[8 + qq]);
            s1_all = fmaxf(s1_all, hd->ts[16 + qq]);
          }
          const float dtf = (float)dt, vpos = fmaxf((float)v_seed, 0.f), amax = (float)p.max_accel;
          const float vmax_all = 1.000002f * (vpos + amax * (float)N * dtf) * dtf;
          const float tv_all = 1.000002f * fmaf(vpos * dtf, s0_all, 1.000002f * amax * dtf * dtf * s1_all);
          loose = make_band<IMU>(hd->bw, vmax_all, tv_all, tl_all, fast_w);
        }
        const int q = pass * T + tid;
        ScanOut<C> so;
        int ic = 0, j = 0;
        unsigned valid = 0;
        Band band{0.f, 0.f, 0.f};
        const bool fast = fast_w;
        const bool in_grid = q < p.n_items;
        if (in_grid) {
          ic = dv(q, p.gs, p.gs_sh);
          j = q - ic * p.gs;
          // two passes over a whole-window VD table: the acceleration chunks are dealt middle-out, so
          // that the first pass holds the accelerations around zero and the second one -- the
          // hardest braking and the hardest acceleration -- is scanned under the first one's bound
          if (!use_skip && vd_full) {
            const int mid = (p.n_ic - 1) >> 1;
            ic = (ic & 1) ? mid + ((ic + 1) >> 1) : mid - (ic >> 1);
          }
#pragma unroll
          for (int c = 0; c < kC; ++c)
            if (ic * kC + c < p.gv) valid |= 1u << c;
        }
        // the item's band inputs, from the per-rate table statistics instead of from the loop:
        // largest step of its fastest hypothesis (V_k is monotone in k, VD non-decreasing in i)
        // and a bound on the total heading variation, sum_k step_k * |TL_k| with
        // step_k <= max(V_w, 0) dt + a k dt^2  (a = largest |a_i| of the chunk's valid
        // hypotheses for the affine headings of the packed scan, max(a_last, 0) otherwise)
        auto item_band = [&]() {
          const int i0 = ic * kC;
          const int il = i0 + kC - 1 < p.gv ? i0 + kC - 1 : p.gv - 1;
          const float inv = (float)p.acc_step;
          const float a_first = inv * (float)(2 * i0 - (p.gv - 1));
          const float a_last = inv * (float)(2 * il - (p.gv - 1));
          const float acoef = fast ? fmaxf(fabsf(a_first), fabsf(a_last)) : fmaxf(a_last, 0.f);
          const float dtf = (float)dt;
          const float* vdc = VD + (ic - ic0) * kC + (kC - 1);
          const float vmax = fmaxf(vdc[0], vdc[(N - 1) * p.vd_cols]);
          const float tv = 1.000002f * fmaf(fmaxf((float)v_seed, 0.f) * dtf, TS[gs4 + j],
                                            1.000002f * acoef * dtf * dtf * TS[2 * gs4 + j]);
          return make_band<IMU>(hd->bw, vmax, tv, TS[j], fast);
        };
        // Pruning (see scan_item): from the second pass on the team holds a bound U, and a warp whose
        // hypotheses have all passed T(U) stops scanning.  The vote compares the first position
        // term's accumulator with T(U) / w_A, rounded up (without a steering penalty the cost is
        // >= fl(w_A * J_A), which then exceeds T(U)).
        float Tq = CUDART_INF_F;
        int every = N > 0 ? N : 1;
        if (p.prune_every > 0 && U < CUDART_INF_F && !ksteer && in_grid) {     // (U: team-uniform)
          // (many-pass kernels without a yaw term: under the window's WIDEST band, which every item's
          // band is below -- a slightly later stop, but no band evaluation per item and pass in front
          // of the scan: -2 % on 128x128; with the yaw term the widest band costs more than it saves)
          Tq = __fdividef((use_skip && !IMU ? loose : item_band()).threshold(U), wA) * 1.000002f;
          Tq = Tq == Tq ? Tq : CUDART_INF_F;
        }
        if (p.prune_every > 0 && U < CUDART_INF_F && !ksteer) every = p.prune_every;
        const unsigned lanes = __ballot_sync(FULL, in_grid);
        // (opaque to the compiler: it would otherwise recompute the vote period and the scan's float
        // constants -- float64 products and conversions -- behind every vote instead of keeping them)
        asm volatile("" : "+r"(every), "+f"(Tq));
        bool pruned = false;
        if (in_grid) {
          if constexpr (C == 8 && !IMU) {
            if (fast) {
              // a_i by multiplication with 1/(G-1): last-bit differences from the spec's
              // division are far below the FP32 resolution this scan works at
              const double inv = p.acc_step;
              const int i0 = ic * kC;
              const double a0d = inv * (double)(2 * i0 - (p.gv - 1));
              const double dtd = dt * dt;
              float f_vwdt = (float)(v_seed * dt), f_dt2 = (float)dtd, f_a0 = (float)a0d,
                    f_da = (float)(2.0 * inv);
              asm volatile("" : "+f"(f_vwdt), "+f"(f_dt2), "+f"(f_a0), "+f"(f_da));
              pruned = scan_item_fast<DUAL>(N, p.gs, p.vd_cols, j, (ic - ic0) * kC, TL, VD, Df, Dab, wA, wB,
                                            ksteer ? JS[j] : 0.f, f_vwdt, f_dt2, f_a0, f_da, Tq, every,
                                            lanes, so);
            }
          }
          if (!fast)
            pruned = scan_item<C, DUAL, IMU>(N, p.gs, p.vd_cols, j, (ic - ic0) * kC, TL, VD, Df, Dab, fI, wA,
                                             wB, wI, ksteer ? JS[j] : 0.f, Tq, every, lanes, so);
        }
        if (pruned) valid = 0;         // nothing of this warp can matter: m = inf, no candidates
        float mj = CUDART_INF_F;       // the item's smallest cost; fminf drops NaN
        bool has_nan = false;
#pragma unroll
        for (int c = 0; c < kC; ++c)
          if ((valid >> c) & 1u) {
            mj = fminf(mj, so.J[c]);
            has_nan |= !(so.J[c] == so.J[c]);
          }
        // J - err(J) <= U needs J <= T(U) under the item's band; T under the widest band is larger
        // still.  A NaN cost is always a candidate.
        if constexpr (SKIP) {
          const bool work = dbg || __any_sync(FULL, has_nan) ||
                            !(warp_min_f32_nonneg(fmaxf(mj, 0.f)) > loose.threshold(U));
          if (!work) valid = 0;        // nothing of this warp can matter: m = inf, no candidates
        }
        if (valid) {
          band = item_band();
          if (dbg) {
#pragma unroll
            for (int c = 0; c < kC; ++c)
              if ((valid >> c) & 1u) {
                const long long o = w * (long long)p.gv * p.gs + (long long)(ic * kC + c) * p.gs + j;
                p.dbg_cost[o] = so.J[c];
                p.dbg_err[o] = band.err(so.J[c]);
              }
          }
        }
        // upper bound on the minimum: J + err(J) grows with J, so only the item's smallest cost
        // needs the band evaluated
        float m = valid ? mj : CUDART_INF_F;
        if (m < CUDART_INF_F) m += band.err(m);
        m = warp_min_f32_nonneg(fmaxf(m, 0.f));
        if (lane == 0) hd->red[warp] = m;
        team.sync();
        float bm = lane < NW ? hd->red[lane] : CUDART_INF_F;
        bm = warp_min_f32_nonneg(bm);
        U = fminf(U, bm);
        unsigned pend = 0;
        const float Jcut = band.threshold(U);   // J - err(J) <= U  <=>  J <= Jcut; NaN stays in
#pragma unroll
        for (int c = 0; c < kC; ++c)
          if (((valid >> c) & 1u) && !(so.J[c] > Jcut)) pend |= 1u << c;
        // drop structural duplicates: only the lowest index of a class can win (np.argmin)
        if (pend) {      // (few threads hold a candidate at all)
          if (j > wi.sat_lo && j <= wi.sat_hi) pend = 0;
#pragma unroll
          for (int c = 0; c < kC; ++c) {
            const int i = ic * kC + c;
            // (with a steering penalty the cost of a motionless row still depends on j)
            if (i < wi.n_dead && (i > 0 || (j > 0 && !ksteer))) pend &= ~(1u << c);
          }
        }
        for (;;) {
#pragma unroll
          for (int c = 0; c < kC; ++c) {
            if (pend == 0) break;
            if ((pend >> c) & 1u) {
              const int slot = atomicAdd(&hd->count, 1);
              if (slot < cand_cap) {
                cand[slot] = make_uint2((unsigned)((ic * kC + c) * p.gs + j),
                                        __float_as_uint(so.J[c] - band.err(so.J[c])));
                pend &= ~(1u << c);
              }
            }
          }
          const int overflow = team.any(pend != 0);
          // (the window's last list is handled here as well, so that the float64 re-score exists once
          // in the kernel's code: ~1 200 instructions less for the instruction cache)
          const bool last = !overflow && pidx == n_pass - 1;
          if (last) {
            // A long list means near-ties (a slow vehicle: every steering rate of the hardest-braking
            // rows costs almost the same).  Re-scoring it here would keep this team busy for tens of
            // microseconds; instead the targets and the list are parked in a slot and the second kernel
            // shares the float64 work of all such windows over the whole GPU.
            {
              const int count = hd->count < cand_cap ? hd->count : cand_cap;
              if (p.defer_buf && count >= p.defer_min) {        // team-uniform
                if (tid == 0) {
                  const unsigned sl = atomicAdd(p.defer_count, 1u);
                  hd->slot = sl < (unsigned)p.defer_slots ? (int)sl : -1;
                }
                team.sync();
                if (hd->slot >= 0) {
                  deferred = true;
                  unsigned char* slot = p.defer_buf + (size_t)hd->slot * p.defer_slot_bytes;
                  const int n_arr = 2 + (DUAL ? 2 : 0) + (IMU ? 1 : 0);
                  double* g_tgt = reinterpret_cast<double*>(slot + kDeferHdrBytes);
                  uint2* g_cand = reinterpret_cast<uint2*>(slot + kDeferHdrBytes + (size_t)n_arr * P * 8);
                  for (int q = tid; q < n_arr * P; q += T) g_tgt[q] = tgt[q];
                  for (int q = tid; q < count; q += T) g_cand[q] = cand[q];
                  if (tid == 0) {
                    DeferHdr dh;
                    dh.w = w;
                    dh.n_steps = N;
                    dh.status = status;
                    dh.count = count;
                    dh.U = U;
                    dh.best_h = -1;
                    dh.best_cost = CUDART_INF;
                    dh.bpose[0] = dh.bpose[1] = dh.bpose[2] = CUDART_NAN;
                    int total = 0;
                    for (int q = 0; q < NW; ++q) {     // what earlier list flushes of this window found
                      total += hd->nres[q];
                      if (hd->bh[q] < 0) continue;
                      if (dh.best_h < 0 || hd->bcost[q] < dh.best_cost ||
                          (hd->bcost[q] == dh.best_cost && hd->bh[q] < dh.best_h)) {
                        dh.best_h = hd->bh[q];
                        dh.best_cost = hd->bcost[q];
                        dh.bpose[0] = hd->bpose[q][0];
                        dh.bpose[1] = hd->bpose[q][1];
                        dh.bpose[2] = hd->bpose[q][2];
                      }
                    }
                    dh.n_rescored = total;
                    dh.wi = hd->wi;
                    *reinterpret_cast<DeferHdr*>(slot) = dh;
                    hd->count = 0;
                  }
                }
              }
            }
          }
          if (overflow || (last && !deferred)) process_list();
          if (!overflow) break;
        }
      }
    }

    // ---- phase D: winner across warps, result record, optional rollout outputs ----------
    team.sync();
    if (tid == 0 && deferred) {     // the slot is complete (every thread's stores precede the barrier)
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(p.defer_ready + hd->slot) = 1u;
    }
    if (tid == 0 && !deferred) {
      int bwi = -1;
      int total = 0;
      for (int q = 0; q < NW; ++q) {
        total += hd->nres[q];
        if (hd->bh[q] < 0) continue;
        if (bwi < 0 || hd->bcost[q] < hd->bcost[bwi] ||
            (hd->bcost[q] == hd->bcost[bwi] && hd->bh[q] < hd->bh[bwi]))
          bwi = q;
      }
      if (status & VMVO_WIN_NONFINITE) bwi = 0;
      vmvo_window_result r;
      r.best_idx = bwi >= 0 ? hd->bh[bwi] : -1;
      r.n_steps = N;
      r.status = status;
      r.n_rescored = total;
      r.best_cost = bwi >= 0 ? hd->bcost[bwi] : CUDART_NAN;
      r.v_seed = hd->wi.v_seed;
      r.s_seed = hd->wi.s_seed;
      r.x1 = bwi >= 0 ? hd->bpose[bwi][0] : CUDART_NAN;
      r.y1 = bwi >= 0 ? hd->bpose[bwi][1] : CUDART_NAN;
      r.theta1 = bwi >= 0 ? hd->bpose[bwi][2] : CUDART_NAN;
      hd->winner = r.best_idx;
      store_record(w, r);
    }
    if (warp == 0 && !deferred) flush_record(w);
    if (chained) {  // last steering angle of the optimum (optimize_trajectory_v2.py:146)
      team.sync();
      const int h = hd->winner;
      if (h >= 0) {
        const int j = h % p.gs;
        GridCtl g{wi.v_seed, wi.s_seed, wi.dt, 0.0, grid_rate(p.max_rate, j, p.gs), p.max_steer};
        double v_unused, s_last;
        g.at(N, &v_unused, &s_last);
        s_chain = s_last;
      }
    }
#ifndef VMVO_EXP_NO_OUT
    if (!kLean && (p.out_poses || p.out_steer || p.out_vel)) {
      team.sync();
      const int h = hd->winner;
      if (warp == 0 && h >= 0) {
        const int i = h / p.gs, j = h - i * p.gs;
        GridCtl g{wi.v_seed, wi.s_seed, wi.dt, grid_rate(p.max_accel, i, p.gv),
                  grid_rate(p.max_rate, j, p.gs), p.max_steer};
        Pose<double> carry{0.0, 0.0, 0.0};
        const bool nonfinite = (status & VMVO_WIN_NONFINITE) != 0;
        for (int base = 0; base < N; base += 32) {
          const int k = base + lane + 1;
          const bool active = k <= N;
          double v = 0.0, s = 0.0;
          if (active) g.at(k, &v, &s);
          Pose<double> pz = warp_model_round<double, 5, SKIP>(nonfinite ? 0.0 : v, s, active && !nonfinite,
                                                     wi.dt, p.L, p.ratio, carry, lane);
          if (active && k <= p.out_stride) {
            const long long o = w * (long long)p.out_stride + (k - 1);
            if (p.out_poses) {
              p.out_poses[o * 3 + 0] = nonfinite ? CUDART_NAN : pz.x;
              p.out_poses[o * 3 + 1] = nonfinite ? CUDART_NAN : pz.y;
              p.out_poses[o * 3 + 2] = nonfinite ? CUDART_NAN : pz.th;
            }
            if (p.out_steer) p.out_steer[o] = s;
            if (p.out_vel) p.out_vel[o] = v;
          }
        }
      }
    }
#endif
    team.sync();
  }
}

// ---- window preparation as a pass of its own ----------------------------------------------------------
// Phases A1-A3 of the search -- local frames (a7), seeds, decimation (a9), float64 targets, FP32 target
// increments, the band's maxima, the duplicate classes -- are a few hundred instructions of serial
// float64 work per window (a sincos, a division, an atan), which a two-warp team of the search kernel
// executes at 16 warps per SM with its partner warp waiting.  Here ONE WARP prepares one window, all
// windows of the launch side by side at full occupancy; the search kernel then stages the finished
// record (PrepLayout) instead of the raw poses and starts at the tables.  The operations are the search
// kernel's own, in the same order: the records hold the same bits either way (every parity test runs
// through both paths).  Item r of the queue is window deal(r), exactly as in the search kernel.
constexpr int kPrepWarps = 8;

// GS lanes per window: a whole warp, or -- for windows of at most 32 poses, the small-grid case the
// pass exists for -- half a warp, two windows per warp side by side (the pass is a latency chain per
// window, so what it costs is waves of resident windows: 11.3 us -> measured below with 16 lanes).
// The two halves run as independent groups: every vote, shuffle and reduction carries the group's
// lane mask.
template <bool DUAL, bool IMU, typename SF, int GS>
__global__ void __launch_bounds__(32 * kPrepWarps)
vmvo_window_prep_kernel(const SearchParams p) {
  using Pose4 = typename PoseOf<SF>::type;
  extern __shared__ __align__(16) unsigned char smem_prep[];
  const int P = p.maxp;
  constexpr int kGroups = 32 / GS;                                  // windows per warp
  const int lane = threadIdx.x & (GS - 1);                          // lane within the group
  const int gshift = (threadIdx.x & 31) & ~(GS - 1);                // first lane of the group in its warp
  const unsigned gmask = GS == 32 ? FULL : (((1u << GS) - 1u) << gshift);
  const int warp = (threadIdx.x >> 5) * kGroups + (gshift / GS);    // group index within the CTA
  const int n_streams = p.load_vo + p.load_gps;
  const int slot_vo = 0, slot_gps = p.load_vo ? 1 : 0;
  const bool traverse = p.target_mode == VMVO_TARGET_TRAVERSE;
  const int warp_bytes = (n_streams * 3 * P + (IMU ? P : 0)) * 8 + (traverse ? ((P * 4 + 15) & ~15) : 0);
  unsigned char* ws = smem_prep + (size_t)warp * warp_bytes;
  double* loc = reinterpret_cast<double*>(ws);
  double* loci = loc + n_streams * 3 * P;
  int* keep = reinterpret_cast<int*>(ws + (n_streams * 3 * P + (IMU ? P : 0)) * 8);
  const Pose4* g_vo = reinterpret_cast<const Pose4*>(p.vo);
  const Pose4* g_gps = reinterpret_cast<const Pose4*>(p.gps);
  const SF* g_imu = reinterpret_cast<const SF*>(p.imu);
  const PrepLayout prl(P, DUAL, IMU);
  const int sA = p.use_vo ? 0 : 1;
  const int off = p.target_offset;
  const unsigned long long kNaN = 0x7ff8000000000000ULL;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // (the search waits before its first record)

  for (long long r = (long long)blockIdx.x * (kPrepWarps * kGroups) + warp; r < p.n_local;
       r += (long long)gridDim.x * (kPrepWarps * kGroups)) {
    long long w = r;
    if (p.sh_world > 1 && p.sh_block_sh >= 0) {
      const long long b = r >> p.sh_block_sh;
      w = (((b * p.sh_world + p.sh_rank) << p.sh_block_sh)) + (r - (b << p.sh_block_sh));
    }
    if (w >= p.n_windows) continue;        // (the deal's padding: the search never asks for it)
    unsigned char* rec = p.prep + (size_t)r * p.prep_stride;
    PrepHdr* hdr = reinterpret_cast<PrepHdr*>(rec);
    float2* Df = reinterpret_cast<float2*>(rec + prl.off_df);
    float2* Dab = reinterpret_cast<float2*>(rec + prl.off_dab);
    float* fI = reinterpret_cast<float*>(rec + prl.off_fi);
    double* tgt = reinterpret_cast<double*>(rec + prl.off_tgt);
    // the plan entry (as the search kernel's fetcher derives it)
    long long start;
    int len, drv;
    if (p.win_start) {
      start = p.win_start[w];
      len = p.win_len[w];
      drv = p.win_drive[w];
    } else {
      int lo = 0, hi = p.n_drives;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.win_off[mid] <= w) lo = mid; else hi = mid;
      }
      drv = lo;
      const long long i = w - p.win_off[drv], f0 = p.drive_off[drv], n = p.drive_off[drv + 1] - f0;
      long long e = i + p.window_frames + 1;
      e = e < n ? e : n;
      start = f0 + i;
      len = (int)(e - i);
    }
    const double dt = p.dt_drive[drv];
    if (len > P || len < 1) {              // not searched: the search kernel only looks at len
      if (lane == 0) {
        PrepHdr h;
        h.v_seed = h.s_seed = __longlong_as_double((long long)kNaN);
        h.dt = dt;
        h.n_targets = h.n_steps = 0;
        h.status = len < 1 ? (VMVO_WIN_EMPTY | VMVO_WIN_NO_FRAMES) : VMVO_WIN_TOO_LONG;
        h.n_dead = 0; h.sat_lo = 0; h.sat_hi = -1; h.len = len;
        h.dmax = h.dabmax = h.imax = 0.f;
        *hdr = h;
      }
      continue;
    }
    const int slot_prim = p.primary == VMVO_PRIMARY_VO ? slot_vo : slot_gps;
    const Pose4* rp = (p.primary == VMVO_PRIMARY_VO ? g_vo : g_gps) + start;

    // rows that never move
    int n_dead = 0;
    {
      const double v_seed = p.seed_mode == VMVO_SEED_GIVEN
                                ? p.seeds[2 * w]
                                : dmul(dadd((double)rp[0].w, (double)rp[len - 1].w), 0.5);
      for (int i0 = 0; i0 < p.gv; i0 += GS) {
        const int i = i0 + lane;
        bool dead = false;
        if (i < p.gv) {
          const double a = grid_rate(p.max_accel, i, p.gv);
          dead = a <= 0.0 && !(dadd(v_seed, dmul(a, dmul(1.0, dt))) > 0.0) && v_seed == v_seed;
        }
        const unsigned b = __ballot_sync(gmask, dead);
        n_dead += __popc(b);
        if (b != gmask) break;
      }
    }

    // ---- phase A1: local frames (a7) -----------------------------------------------------
    for (int s = 0; s < 2; ++s) {
      if (!(s == 0 ? p.load_vo : p.load_gps)) continue;
      const int slot = s == 0 ? slot_vo : slot_gps;
      const Pose4* rs = (s == 0 ? g_vo : g_gps) + start;
      const Pose4 p0 = rs[0];
      const double th0 = (double)p0.z;
      double sn, cs;
      sincos(th0, &sn, &cs);
      double* lx = loc + (slot * 3 + 0) * P;
      double* ly = loc + (slot * 3 + 1) * P;
      double* lt = loc + (slot * 3 + 2) * P;
      for (int m = lane; m < len; m += GS) {
        const Pose4 q = rs[m];
        const double dx = dsub((double)q.x, (double)p0.x);
        const double dy = dsub((double)q.y, (double)p0.y);
        lx[m] = dadd(dmul(dx, cs), dmul(dy, sn));
        ly[m] = dadd(dmul(-dx, sn), dmul(dy, cs));
        lt[m] = dsub((double)q.z, th0);
      }
    }
    if (IMU) {
      const double y0 = (double)g_imu[start];
      for (int m = lane; m < len; m += GS) loci[m] = dsub((double)g_imu[start + m], y0);
    }
    __syncwarp(gmask);

    // ---- phase A2: seeds, decimation (a9) -----------------------------------------------
    const double* plx = loc + (slot_prim * 3 + 0) * P;
    const double* ply = loc + (slot_prim * 3 + 1) * P;
    const double* plt = loc + (slot_prim * 3 + 2) * P;
    double v_seed, s_seed;
    if (p.seed_mode == VMVO_SEED_GIVEN) {
      v_seed = p.seeds[2 * w];
      s_seed = p.seeds[2 * w + 1];
    } else {
      v_seed = dmul(dadd((double)rp[0].w, (double)rp[len - 1].w), 0.5);   // (/ 2, exactly)
      s_seed = 0.0;
      if (len >= 2 && dmul(v_seed, dt) > 1e-6) {
        const double dth0 = dsub(plt[1], plt[0]);
        const double dth = fabs(dth0) < kPi ? dth0 : remainder(dth0, kTwoPi);
        const double ang = atan(ddiv(dmul(p.L, dth), dmul(v_seed, dt)));
        s_seed = dmul(dmul(ang, kRadToDeg), p.ratio);
        s_seed = s_seed < -p.max_steer ? -p.max_steer : s_seed;
        s_seed = s_seed > p.max_steer ? p.max_steer : s_seed;
      }
    }
    int n_targets = len;
    if (traverse) {
      if (lane == 0) {  // sequential by definition (distance accumulator with reset)
        const double D = dmul(v_seed, dt);
        int cnt = 1;
        keep[0] = 0;
        double dist = 0.0;
        for (int i = 1; i < len; ++i) {
          const double ddx = dsub(plx[i], plx[i - 1]), ddy = dsub(ply[i], ply[i - 1]);
          const double seg = sqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy)));
          if (dadd(dist, seg) > D) {
            keep[cnt++] = i - 1;
            dist = seg;
          } else {
            dist = dadd(dist, seg);
          }
        }
        n_targets = cnt;
      }
      n_targets = __shfl_sync(gmask, n_targets, gshift);
      __syncwarp(gmask);
    }
    int sat_lo = 0, sat_hi = -1;
    if (s_seed == p.max_steer || s_seed == -p.max_steer) {
      const bool hi = s_seed > 0;
      int first = p.gs, last = -1;
      for (int j0 = 0; j0 < p.gs; j0 += GS) {
        const int j = j0 + lane;
        bool in = false;
        if (j < p.gs) {
          const double rr = grid_rate(p.max_rate, j, p.gs);
          in = hi ? (rr >= 0.0) : (rr <= 0.0);
        }
        const unsigned b = __ballot_sync(gmask, in) >> gshift;      // (bit l = lane l of the group)
        if (b) {
          first = min(first, j0 + __ffs(b) - 1);
          last = max(last, j0 + 31 - __clz(b));
        }
      }
      sat_lo = first;
      sat_hi = last;
    }
    const int N = n_targets > 1 ? n_targets - 1 : 0;

    // ---- phase A3: targets in float64, FP32 increments for the scan, finiteness -------------
    bool finite = isfinite(v_seed) && isfinite(s_seed) && isfinite(dt);
    float dmax = 0.f, dabmax = 0.f, imax = 0.f;
    {
      const int slot_a = sA == 0 ? slot_vo : slot_gps;
      const double* aX = loc + (slot_a * 3 + 0) * P;
      const double* aY = loc + (slot_a * 3 + 1) * P;
      const double* bX = loc + (slot_gps * 3 + 0) * P;   // B is GPS (DUAL only)
      const double* bY = loc + (slot_gps * 3 + 1) * P;
      for (int q = lane; q < n_targets; q += GS) {
        const int m = traverse ? keep[q] : q;
        const double ax = aX[m], ay = aY[m];
        tgt[q] = ax;
        tgt[P + q] = ay;
        finite = finite && isfinite(ax) && isfinite(ay);
        if (DUAL) {
          const double bx = bX[m], by = bY[m];
          tgt[2 * P + q] = bx;
          tgt[3 * P + q] = by;
          finite = finite && isfinite(bx) && isfinite(by);
        }
        if (IMU) {
          const double yi = loci[m];
          tgt[(DUAL ? 4 : 2) * P + q] = yi;
          finite = finite && isfinite(yi);
        }
        const int k = q + off;
        if (k >= 1 && k <= N) {
          const int mp = (q >= 1) ? (traverse ? keep[q - 1] : q - 1) : -1;
          const double px = (k >= 2) ? aX[mp] : 0.0, py = (k >= 2) ? aY[mp] : 0.0;
          const float dx = (float)dsub(ax, px), dy = (float)dsub(ay, py);
          Df[k] = make_float2(dx, dy);
          dmax = fmaxf(dmax, fmaxf(fabsf(dx), fabsf(dy)));
          if (DUAL) {
            const float ex = (float)dsub(ax, bX[m]), ey = (float)dsub(ay, bY[m]);
            Dab[k] = make_float2(ex, ey);
            dabmax = fmaxf(dabmax, fmaxf(fabsf(ex), fabsf(ey)));
          }
          if (IMU) {
            const float yi = (float)loci[m];
            fI[k] = yi;
            imax = fmaxf(imax, fabsf(yi));
          }
        }
      }
    }
    if (!__all_sync(gmask, finite)) dmax = CUDART_INF_F;
    dmax = __uint_as_float(__reduce_max_sync(gmask, __float_as_uint(dmax)));
    if (DUAL) dabmax = __uint_as_float(__reduce_max_sync(gmask, __float_as_uint(dabmax)));
    if (IMU) imax = __uint_as_float(__reduce_max_sync(gmask, __float_as_uint(imax)));
    if (lane == 0) {
      PrepHdr h;
      h.v_seed = v_seed; h.s_seed = s_seed; h.dt = dt;
      h.n_targets = n_targets; h.n_steps = N; h.status = 0;
      h.n_dead = n_dead; h.sat_lo = sat_lo; h.sat_hi = sat_hi; h.len = len;
      h.dmax = dmax; h.dabmax = dabmax; h.imax = imax;
      *hdr = h;
    }
    __syncwarp(gmask);
  }
}

// ---- second kernel: the float64 re-scores of the deferred windows -------------------------------
// One CTA of eight warps per slot.  The slot (header, float64 targets, list: a few KB) is copied to
// shared memory first, and the warps share the list four entries at a time, each through the same
// arithmetic as in the search kernel: warp_cost64, or warp_cost64_pack when all four stop within 8
// or 16 steps -- which is what the near-ties of a slow vehicle do (same values, same operations,
// same cost: test_deferred_windows_give_the_same_records).  Measured alternatives: 4 or 2 warps per
// slot (6 / 12 slots per SM) and a persistent grid are slower.
//
// The kernel is launched with programmatic stream serialization right behind the search, whose CTAs
// (one team each) all signal launch_dependents when they start: its CTAs take over an SM's registers
// as soon as teams there have run out of windows, i.e. they work through the parked windows during
// the END of the search, when a growing share of the SMs would otherwise idle (the last window of a
// team ends up to one window time after the queue runs dry).  No CTA relies on the search having
// completed: CTA b handles slots b, b + grid, ... and waits for each slot's ready word; a slot index
// is known to stay empty once every team of the search has run out of windows (a counter the teams
// bump on their way out) and the allocation count is below it.


__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <bool DUAL, bool IMU, int kDeferWarps>
__global__ void __launch_bounds__(32 * kDeferWarps, 24 / kDeferWarps)
vmvo_deferred_rescore_kernel(const SearchParams p) {
  extern __shared__ __align__(16) unsigned char s_slot[];
  __shared__ double s_cost[kDeferWarps], s_pose[kDeferWarps][3];
  __shared__ int s_h[kDeferWarps], s_n[kDeferWarps];
  __shared__ int s_go;
  __shared__ __align__(16) vmvo_window_result s_rec;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = p.maxp;
  const int n_arr = 2 + (DUAL ? 2 : 0) + (IMU ? 1 : 0);
  const double wA = p.use_vo ? p.w_vo : p.w_gps, wB = p.w_gps;
  for (unsigned s = blockIdx.x; s < (unsigned)p.defer_slots; s += gridDim.x) {
    if (threadIdx.x == 0) {     // slot s: published, or known to stay empty
      int go = 0;
      unsigned long long t0 = 0;
      for (unsigned spin = 0;; ++spin) {
        if (ld_acquire_gpu_u32(p.defer_ready + s)) { go = 1; break; }
        if (ld_acquire_gpu_u64(p.windows_done) >= (unsigned long long)p.n_todo) {
          go = ld_acquire_gpu_u32(p.defer_ready + s) != 0;     // every list is final now
          break;
        }
        if ((spin & 1023) == 1023) {     // a search that died must not leave this kernel spinning
          unsigned long long t;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
          if (t0 == 0) t0 = t;
          else if (t - t0 > 20ull * 1000 * 1000 * 1000) break;
        }
        __nanosleep(200);
      }
      s_go = go;
    }
    __syncthreads();
    if (!s_go) break;       // (slots are handed out in order: none beyond this one either)
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.defer_buf + (size_t)s * p.defer_slot_bytes);
      uint4* dst = reinterpret_cast<uint4*>(s_slot);
      for (int q = threadIdx.x; q < p.defer_slot_bytes / 16; q += 32 * kDeferWarps) dst[q] = src[q];
    }
    __syncthreads();
    const unsigned char* slot = s_slot;
    const DeferHdr* dh = reinterpret_cast<const DeferHdr*>(slot);
    const double* tgt = reinterpret_cast<const double*>(slot + kDeferHdrBytes);
    const uint2* cand = reinterpret_cast<const uint2*>(slot + kDeferHdrBytes + (size_t)n_arr * P * 8);
    const int count = dh->count;
    const float U = dh->U;
    float Uw = CUDART_INF_F;
    int best_h = -1, n_res = 0;
    double best_cost = CUDART_INF;
    Pose<double> best_first{CUDART_NAN, CUDART_NAN, CUDART_NAN};
    if (warp == 0 && dh->best_h >= 0) {   // what the search kernel had already re-scored
      best_h = dh->best_h;
      best_cost = dh->best_cost;
      best_first = Pose<double>{dh->bpose[0], dh->bpose[1], dh->bpose[2]};
    }
    auto take = [&](int h, double c64, const Pose<double>& first) {
      ++n_res;
      if (best_h < 0 || c64 < best_cost || (c64 == best_cost && h < best_h)) {
        best_h = h;
        best_cost = c64;
        best_first = first;
      }
      Uw = fminf(Uw, __double2float_ru(c64));
    };
    // a warp takes four consecutive list entries at a time, one per group of eight lanes
    const int N = dh->n_steps;
    for (int e0 = 4 * warp; e0 < count; e0 += 4 * kDeferWarps) {
      const int e = e0 + (lane >> 3);
      int h_grp = -1;
      if (e < count) {
        const uint2 ce = cand[e];
        if (!(__uint_as_float(ce.y) > fminf(U, Uw))) h_grp = (int)ce.x;   // NaN stays in
      }
      // packable: none of the four still moves after step 8 (or 16); each group checks its own,
      // eight steps per turn, with the very control formula the rollout uses
      bool late8 = false, late16 = false;
      if (h_grp >= 0) {
        const int i = div_sh(h_grp, p.gs, p.gs_sh);
        const GridCtl g{dh->wi.v_seed, dh->wi.s_seed, dh->wi.dt, grid_rate(p.max_accel, i, p.gv), 0.0,
                        p.max_steer};
        for (int k = 9 + (lane & 7); k <= N; k += 8) {
          double v, s_unused;
          g.at(k, &v, &s_unused);
          late8 |= v != 0.0;
          late16 |= v != 0.0 && k > 16;
        }
      }
      if (!__any_sync(FULL, h_grp >= 0)) continue;
      if (p.scan_hs) {     // (a forced deferral of a many-pass launch: its sums are Hillis-Steele)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int hc = __shfl_sync(FULL, h_grp, 8 * c);
          if (hc < 0) continue;              // warp-uniform
          Pose<double> first;
          const double c64 = warp_cost64<DUAL, IMU, true>(p, dh->wi, tgt, P, hc, lane, wA, wB, &first);
          take(hc, c64, first);
        }
      } else if (!__any_sync(FULL, late8)) {
        double c64[4];
        Pose<double> first[4];
        warp_cost64_pack<DUAL, IMU, 3>(p, dh->wi, tgt, P, h_grp, lane, wA, wB, c64, first);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int hc = __shfl_sync(FULL, h_grp, 8 * c);
          if (hc >= 0) take(hc, c64[c], first[c]);
        }
      } else if (!__any_sync(FULL, late16)) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {      // entries (0, 1), then (2, 3): one per 16 lanes
          const int h16 = __shfl_sync(FULL, h_grp, 8 * (2 * half + (lane >> 4)));
          if (!__any_sync(FULL, h16 >= 0)) continue;
          double c64[2];
          Pose<double> first[2];
          warp_cost64_pack<DUAL, IMU, 4>(p, dh->wi, tgt, P, h16, lane, wA, wB, c64, first);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int hc = __shfl_sync(FULL, h16, 16 * c);
            if (hc >= 0) take(hc, c64[c], first[c]);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int hc = __shfl_sync(FULL, h_grp, 8 * c);
          if (hc < 0) continue;              // warp-uniform
          Pose<double> first;
          const double c64 = warp_cost64<DUAL, IMU>(p, dh->wi, tgt, P, hc, lane, wA, wB, &first);
          take(hc, c64, first);
        }
      }
    }
    if (lane == 0) {
      s_h[warp] = best_h;
      s_cost[warp] = best_cost;
      s_pose[warp][0] = best_first.x;
      s_pose[warp][1] = best_first.y;
      s_pose[warp][2] = best_first.th;
      s_n[warp] = n_res;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int bwi = -1, total = dh->n_rescored;
      for (int q = 0; q < kDeferWarps; ++q) {
        total += s_n[q];
        if (s_h[q] < 0) continue;
        if (bwi < 0 || s_cost[q] < s_cost[bwi] || (s_cost[q] == s_cost[bwi] && s_h[q] < s_h[bwi])) bwi = q;
      }
      vmvo_window_result r;
      r.best_idx = bwi >= 0 ? s_h[bwi] : -1;
      r.n_steps = dh->n_steps;
      r.status = dh->status;
      r.n_rescored = total;
      r.best_cost = bwi >= 0 ? s_cost[bwi] : CUDART_NAN;
      r.v_seed = dh->wi.v_seed;
      r.s_seed = dh->wi.s_seed;
      r.x1 = bwi >= 0 ? s_pose[bwi][0] : CUDART_NAN;
      r.y1 = bwi >= 0 ? s_pose[bwi][1] : CUDART_NAN;
      r.theta1 = bwi >= 0 ? s_pose[bwi][2] : CUDART_NAN;
      s_rec = r;
    }
    __syncthreads();
    {   // four threads per destination: every copy of the record leaves as one 64-byte write
      const uint4 part = reinterpret_cast<const uint4*>(&s_rec)[threadIdx.x & 3];
      for (int t = threadIdx.x >> 2; t <= p.n_mirrors; t += (32 * kDeferWarps) >> 2) {
        vmvo_window_result* base = t == 0 ? p.results : p.mirrors[t - 1];
        reinterpret_cast<uint4*>(base + dh->w)[threadIdx.x & 3] = part;
      }
    }
    __syncthreads();
  }
  // what follows this kernel in the stream follows the search as well
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <int C, int WARPS, int MINB, bool DUAL, bool IMU, typename SF, bool SKIP, int MODE>
static int launch_search_v(vmvo_ctx* ctx, const SearchParams& p_in, cudaStream_t st) {
  SearchParams p = p_in;
  p.scan_hs = SKIP ? 1 : 0;      // the many-pass kernel re-scores with Hillis-Steele sums
  auto kern = vmvo_window_search_kernel<C, WARPS, MINB, DUAL, IMU, SF, SKIP, MODE>;
  // one team per CTA (team_warps <= WARPS warps): a team that runs out of windows gives its
  // registers and shared memory back at once, which is what lets the second kernel move in
  // WARPS warps per CTA, i.e. several teams: CTAs of one two-warp team measured 1.35x slower (a CTA's
  // warps are dealt to the four SM sub-partitions by warp index, so two-warp CTAs leave two of them idle)
  const int teams = ctx->tune.cta_teams > 0 ? ctx->tune.cta_teams : WARPS / p.team_warps;
  const int cta_threads = 32 * p.team_warps * teams;
  const PrepLayout prl(p.maxp, DUAL, IMU);
  const SmemLayout lay(p.maxp, p.gs, p.vd_cols, p.team_warps, p.load_vo + p.load_gps, DUAL, IMU,
                       p.target_mode == VMVO_TARGET_TRAVERSE, (int)(4 * sizeof(SF)), p.prep ? prl.total : 0);
  const int smem = lay.total * teams;
  if (smem > 200 * 1024)
    return fail(ctx, VMVO_ERR_UNSUPPORTED,
                "window tables need %d bytes of shared memory (max_window_poses %d x grid_s %d): "
                "reduce max_window_poses or grid_s", smem, p.maxp, p.gs);
  VMVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  // the whole carve-out for shared memory: eight small CTAs per SM need ~130 KB between them
  VMVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  VMVO_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, cta_threads, smem));
  if (per_sm < 1) return fail(ctx, VMVO_ERR_CUDA, "search kernel does not fit on an SM");
  // (the kernel is built for 128 registers per thread: 16 warps per SM)
  const int cap = (WARPS * MINB) / (p.team_warps * teams);
  if (per_sm > cap) per_sm = cap;
  if (ctx->tune.max_ctas_per_sm >= 1 && ctx->tune.max_ctas_per_sm < per_sm) per_sm = ctx->tune.max_ctas_per_sm;
  long long grid = (long long)ctx->sm_count * per_sm;
  long long need = (p.n_local + teams - 1) / teams;
  if (need < 1) need = 1;      // (an empty share still advances the exchange's step counter)
  if (grid > need) grid = need;
  p.n_todo = grid * teams;
  if (p.prep) {     // phases A1-A3 of every window of the launch, one warp per window
    const bool half = p.maxp <= 32 && ctx->tune.prep != 32;       // two windows per warp
    auto pk = half ? vmvo_window_prep_kernel<DUAL, IMU, SF, 16> : vmvo_window_prep_kernel<DUAL, IMU, SF, 32>;
    const int per_cta = kPrepWarps * (half ? 2 : 1);
    const int n_streams = p.load_vo + p.load_gps;
    const bool trav = p.target_mode == VMVO_TARGET_TRAVERSE;
    const int warp_bytes = (n_streams * 3 * p.maxp + (IMU ? p.maxp : 0)) * 8 + (trav ? ((p.maxp * 4 + 15) & ~15) : 0);
    const int psmem = warp_bytes * per_cta;
    VMVO_CUDA(ctx, cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem));
    long long pg = (p.n_local + per_cta - 1) / per_cta;
    const long long pcap = (long long)ctx->sm_count * 16;
    if (pg > pcap) pg = pcap;
    if (pg < 1) pg = 1;
    pk<<<(unsigned)pg, 32 * kPrepWarps, psmem, st>>>(p);
    int prc = check_launch(ctx, "vmvo_window_prep_kernel");
    if (prc) return prc;
  }
  if (p.prep && ctx->tune.pdl != 0) {
    cudaLaunchConfig_t sc = {};
    sc.gridDim = dim3((unsigned)grid);
    sc.blockDim = dim3(cta_threads);
    sc.dynamicSmemBytes = (size_t)smem;
    sc.stream = st;
    cudaLaunchAttribute sa[1];
    sa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    sa[0].val.programmaticStreamSerializationAllowed = 1;
    sc.attrs = sa;
    sc.numAttrs = 1;
    VMVO_CUDA(ctx, cudaLaunchKernelEx(&sc, kern, p));
  } else {
    kern<<<(unsigned)grid, cta_threads, smem, st>>>(p);
  }
  int rc = check_launch(ctx, "vmvo_window_search_kernel");
  if (rc || !p.defer_buf) return rc;
  long long g2 = p.defer_slots < (long long)ctx->sm_count * 16 ? p.defer_slots : (long long)ctx->sm_count * 16;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(g2 > 0 ? g2 : 1));
  lc.blockDim = dim3(32 * 8);
  lc.dynamicSmemBytes = (size_t)p.defer_slot_bytes;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = ctx->tune.pdl == 0 ? 0 : 1;
  if (ctx->tune.defer_warps != 2) {
    VMVO_CUDA(ctx, cudaLaunchKernelEx(&lc, vmvo_deferred_rescore_kernel<DUAL, IMU, 8>, p));
  } else {
    lc.blockDim = dim3(32 * 2);
    VMVO_CUDA(ctx, cudaLaunchKernelEx(&lc, vmvo_deferred_rescore_kernel<DUAL, IMU, 2>, p));
  }
  return check_launch(ctx, "vmvo_deferred_rescore_kernel");
}

template <int C, int WARPS, int MINB, bool DUAL, bool IMU, typename SF, int MODE>
static int launch_search(vmvo_ctx* ctx, const SearchParams& p, cudaStream_t st) {
  const int n_pass = (p.n_items + p.team_warps * 32 - 1) / (p.team_warps * 32);
  if constexpr (MODE == 2) {     // (preparation records go with teams of one or two warps: two passes)
    if (n_pass > 2) return fail(ctx, VMVO_ERR_UNSUPPORTED, "internal: preparation records with %d passes", n_pass);
    return launch_search_v<C, WARPS, MINB, DUAL, IMU, SF, false, MODE>(ctx, p, st);
  } else {
    return n_pass > 2 ? launch_search_v<C, WARPS, MINB, DUAL, IMU, SF, true, MODE>(ctx, p, st)
                      : launch_search_v<C, WARPS, MINB, DUAL, IMU, SF, false, MODE>(ctx, p, st);
  }
}

// every (cost terms, stream type) instantiation of one MODE: each MODE is compiled in a translation
// unit of its own (vmvo_search.cu: 0, vmvo_search_lean.cu: 1, vmvo_search_prep.cu: 2), side by side
template <int MODE>
static int launch_search_mode(vmvo_ctx* ctx, const SearchParams& p, cudaStream_t st, bool dual, bool imu,
                              bool f64) {
#define VMVO_LAUNCH(SF)                                                                       \
  (dual ? (imu ? launch_search<8, 8, VMVO_MINB, true, true, SF, MODE>(ctx, p, st)             \
               : launch_search<8, 8, VMVO_MINB, true, false, SF, MODE>(ctx, p, st))           \
        : (imu ? launch_search<8, 8, VMVO_MINB, false, true, SF, MODE>(ctx, p, st)            \
               : launch_search<8, 8, VMVO_MINB, false, false, SF, MODE>(ctx, p, st)))
  return f64 ? VMVO_LAUNCH(double) : VMVO_LAUNCH(float);
#undef VMVO_LAUNCH
}

int launch_search_mode0(vmvo_ctx* ctx, const SearchParams& p, cudaStream_t st, bool dual, bool imu, bool f64);
int launch_search_mode1(vmvo_ctx* ctx, const SearchParams& p, cudaStream_t st, bool dual, bool imu, bool f64);
int launch_search_mode2(vmvo_ctx* ctx, const SearchParams& p, cudaStream_t st, bool dual, bool imu, bool f64);

}  // namespace vmvo
