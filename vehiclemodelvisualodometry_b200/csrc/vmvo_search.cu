// Fused window search: window staging (TMA bulk copy) -> local frame (a7) -> decimation (a9)
// -> seeds -> hypothesis grid generated in registers -> batched bicycle integration (a1/a2)
// -> cost (a10) -> block argmin -> exact float64 re-score of every hypothesis whose FP32 cost
// lies within a rigorous error band of the minimum -> per-window result record.
//
// Replaces mpc_run (vmvo/utils/mpc.py:14-122) and the per-window preparation of
// optimize_trajectory (vmvo/scripts/optimize_trajectory_v2.py:49-96) of the reference.
//
// Work decomposition: one CTA works on one window at a time and pulls
// windows from a global queue (persistent grid).  Inside a window a thread owns one steering
// rate r_j and C consecutive accelerations a_i: tan(delta_k(j)) is computed once per step and
// shared by the C hypotheses; each hypothesis costs 2 MUFU (sin, cos) + ~10 FP32 per step.
#include "vmvo_device.cuh"
#include "vmvo_internal.h"

#include <math_constants.h>

namespace vmvo {

constexpr int kCandCap = 1024;   // candidate list entries per CTA (flushed when full)
constexpr int kMaxWarps = 8;     // CTA size <= 256 threads

struct SearchParams {
  int gv, gs;
  int n_ic;              // ceil(gv / C)
  int n_items;           // n_ic * gs  (one item = one thread's C hypotheses)
  int target_mode, target_offset, seed_mode, primary;
  int maxp;              // pose capacity per window
  int use_vo, use_gps;   // position terms with non-zero weight
  int load_vo, load_gps; // streams staged into shared memory
  double w_vo, w_gps, w_imu, k_steer;
  double L, ratio, max_steer, max_accel, max_rate;
  double delta_max, tan_max;
  const long long* win_start;
  const int* win_len;
  const int* win_drive;
  const double* dt_drive;
  const float4* vo;
  const float4* gps;
  const float* imu;
  const double* seeds;
  vmvo_window_result* results;
  double* out_poses;
  double* out_steer;
  double* out_vel;
  int out_stride;
  long long n_windows;
  unsigned long long* work_counter;
};

// per-window scalars shared by the CTA
struct WinInfo {
  double v_seed, s_seed, dt;
  long long start;
  int len, n_targets, n_steps, status;
};

// ---- shared-memory carve-up (same function on host and device) ---------------------------
struct SmemLayout {
  int off_raw, off_loc, off_loci, off_tgt, off_fa, off_fb, off_fi, off_keep, off_cand, total;
  __host__ __device__ explicit SmemLayout(int P) {
    int o = 1024;                       // fixed header: barriers, window ids, reductions
    off_raw = o;  o += 2 * 2 * P * 16;  // float4 raw[2 buffers][2 streams][P]
    off_loc = o;  o += 2 * 3 * P * 8;   // double loc[2 streams][3][P]  (lx, ly, lth)
    off_loci = o; o += P * 8;           // double imu yaw relative to the window start
    off_tgt = o;  o += 5 * P * 8;       // double tAx, tAy, tBx, tBy, tI
    off_fa = o;   o += P * 8;           // float2 fA
    off_fb = o;   o += P * 8;           // float2 fB
    off_fi = o;   o += P * 4;           // float fI
    off_keep = o; o += P * 4;           // int keep
    o = (o + 15) & ~15;
    off_cand = o; o += kCandCap * 8;    // uint2 (hypothesis, float cost bits)
    total = o;
  }
};

struct SmemHeader {
  uint64_t mbar[2];
  long long wid[2];
  WinInfo wi;
  float red[32];
  double bcost[kMaxWarps];
  double bpose[kMaxWarps][3];
  int bh[kMaxWarps];
  int nres[kMaxWarps];
  int count;
  int winner;
};
static_assert(sizeof(SmemHeader) <= 1024, "header too large");

// err(J) = c0 + c1*sqrt(J) + c2*J bounds |J_fp32 - J_fp64|
struct Band {
  float c0, c1, c2;
  __device__ __forceinline__ float err(float J) const { return fmaf(c1, sqrtf(J), fmaf(c2, J, c0)); }
  __device__ __forceinline__ float upper(float J) const { return J + err(J); }
  __device__ __forceinline__ float lower(float J) const { return J - err(J); }
};

__device__ Band make_band(const SearchParams& p, int N, double dt, double v_seed, double tmax,
                          double imax) {
  const double u = 5.9604644775390625e-8;  // 2^-24
  const double n = (double)N;
  double Vmax = (v_seed > 0 ? v_seed : 0) + p.max_accel * n * dt;
  double vdtm = Vmax * dt;
  double eps_d = 4 * u * p.delta_max;
  double eps_T = (1 + p.tan_max * p.tan_max) * eps_d * 1.05 + 8 * u * p.tan_max;
  double eps_v = 6 * u * vdtm;
  double dth_max = vdtm * p.tan_max / p.L;
  double th_max = n * dth_max;
  double g = vdtm * eps_T / p.L + eps_v * p.tan_max / p.L + 2 * u * dth_max + 2 * u * th_max;
  double eps_trig = 4.76837158203125e-7 /* 2^-21 */ + 4 * u * th_max;
  double Xmax = n * vdtm;
  double a = vdtm * eps_trig + eps_v + 2 * u * Xmax;  // pose error grows by <= a + b*k per step
  double b = vdtm * g;
  double e = u * tmax;
  double s2 = n * (n + 1) * (2 * n + 1) / 6;                    // sum k^2
  double s4 = n * n * n * n * n / 5 + n * n * n * n / 2 + n * n * n / 3;  // >= sum k^4
  double E2pos = 3 * (a * a * s2 + b * b * s4 + n * e * e);
  double ei = 8 * u * kPi * (1 + th_max / kTwoPi) + u * imax;
  double E2imu = 2 * (g * g * s2 + n * ei * ei);
  double E2w = (p.w_vo + p.w_gps) * E2pos + p.w_imu * E2imu;
  double c2 = 2.8284271247461903 * u * sqrt(n) + 4 * u * u * n + 2 * (n + 2) * u + 16 * u;
  const double SF = 2.0;  // safety factor on the whole bound
  Band bd;
  bd.c0 = __double2float_ru(SF * 4 * E2w);
  bd.c1 = __double2float_ru(SF * 2.8284271247461903 * sqrt(E2w));
  bd.c2 = __double2float_ru(SF * c2);
  return bd;
}

// ---- float64 cost of one hypothesis, one warp, one step per lane ---------------------------
template <bool DUAL, bool IMU>
__device__ double warp_cost64(const SearchParams& p, const WinInfo& wi, const double* tgt, int P,
                              int h, int lane, double wA, double wB, Pose<double>* first) {
  const int i = h / p.gs, j = h - i * p.gs;
  GridCtl g{wi.v_seed, wi.s_seed, wi.dt, grid_rate(p.max_accel, i, p.gv),
            grid_rate(p.max_rate, j, p.gs), p.max_steer};
  const double* tAx = tgt;
  const double* tAy = tgt + P;
  const double* tBx = tgt + 2 * P;
  const double* tBy = tgt + 3 * P;
  const double* tI = tgt + 4 * P;
  Pose<double> carry{0.0, 0.0, 0.0};
  double J = 0.0;
  const int N = wi.n_steps;
  for (int base = 0; base < N; base += 32) {
    const int k = base + lane + 1;
    const bool active = k <= N;
    double v = 0.0, s = 0.0;
    if (active) g.at(k, &v, &s);
    Pose<double> pz = warp_model_round<double>(v, s, active, wi.dt, p.L, p.ratio, carry, lane);
    if (base == 0) {
      first->x = __shfl_sync(FULL, pz.x, 0);
      first->y = __shfl_sync(FULL, pz.y, 0);
      first->th = __shfl_sync(FULL, pz.th, 0);
    }
    double term = 0.0;
    if (active) {
      const int t = k - p.target_offset;
      double ex = dsub(pz.x, tAx[t]), ey = dsub(pz.y, tAy[t]);
      double e = dadd(dmul(ex, ex), dmul(ey, ey));
      term = (wA == 1.0) ? e : dmul(wA, e);
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
    }
    J = dadd(J, warp_sum(term));
  }
  return J;
}

// ---- FP32 scan of one item: steering rate j, accelerations ic*C .. ic*C+C-1 -----------------
template <int C, bool DUAL, bool IMU>
__device__ __forceinline__ void scan_item(const SearchParams& p, const WinInfo& wi, int ic, int j,
                                          const float2* __restrict__ fA,
                                          const float2* __restrict__ fB,
                                          const float* __restrict__ fI, float wA, float wB,
                                          float (&Jt)[C]) {
  const double kd = kDegToRad / p.ratio;            // steering-wheel degrees -> road-wheel rad
  const float rdd = (float)(grid_rate(p.max_rate, j, p.gs) * wi.dt * kd);
  const float dw = (float)(wi.s_seed * kd);
  const float dmaxf = (float)p.delta_max;
  const float vwdt = (float)(wi.v_seed * wi.dt);
  const float invL = (float)(1.0 / p.L);
  const float rad2deg = (float)(1.0 / kd);
  const float wI = (float)p.w_imu;
  const bool ksteer = p.k_steer != 0.0;
  float adt2[C], th[C], x[C], y[C], JA[C], JB[C], JI[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    int i = ic * C + c;
    i = i < p.gv ? i : p.gv - 1;
    adt2[c] = (float)(grid_rate(p.max_accel, i, p.gv) * wi.dt * wi.dt);
    th[c] = x[c] = y[c] = JA[c] = JB[c] = JI[c] = 0.f;
  }
  float JS = 0.f;
  const int N = wi.n_steps;
  const int off = p.target_offset;
  for (int k = 1; k <= N; ++k) {
    const float kf = (float)k;
    float d = fmaf(rdd, kf, dw);
    d = fminf(fmaxf(d, -dmaxf), dmaxf);
    const float TL = tanf(d) * invL;
    const float2 ta = fA[k - off];
    float2 tb = make_float2(0.f, 0.f);
    float ti = 0.f;
    if (DUAL) tb = fB[k - off];
    if (IMU) ti = fI[k - off];
    if (ksteer) {
      float sd = d * rad2deg;
      JS = fmaf(sd, sd, JS);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float vdt = fmaxf(fmaf(adt2[c], kf, vwdt), 0.f);
      th[c] = fmaf(vdt, TL, th[c]);
      float sn, cs;
      __sincosf(th[c], &sn, &cs);
      x[c] = fmaf(vdt, cs, x[c]);
      y[c] = fmaf(vdt, sn, y[c]);
      float ex = x[c] - ta.x, ey = y[c] - ta.y;
      JA[c] = fmaf(ex, ex, JA[c]);
      JA[c] = fmaf(ey, ey, JA[c]);
      if (DUAL) {
        ex = x[c] - tb.x;
        ey = y[c] - tb.y;
        JB[c] = fmaf(ex, ex, JB[c]);
        JB[c] = fmaf(ey, ey, JB[c]);
      }
      if (IMU) {
        float e = th[c] - ti;
        e = fmaf(-rintf(e * 0.15915494309189535f), 6.283185307179586f, e);
        JI[c] = fmaf(e, e, JI[c]);
      }
    }
  }
  const float kS = ksteer ? (float)p.k_steer * JS : 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float t = wA * JA[c] + kS;
    if (DUAL) t = fmaf(wB, JB[c], t);
    if (IMU) t = fmaf(wI, JI[c], t);
    Jt[c] = t;
  }
}

// ---- the kernel -----------------------------------------------------------------------------
template <int C, bool DUAL, bool IMU>
__global__ void __launch_bounds__(256)
vmvo_window_search_kernel(const SearchParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int P = p.maxp;
  const SmemLayout lay(P);
  SmemHeader* hd = reinterpret_cast<SmemHeader*>(smem);
  float4* raw = reinterpret_cast<float4*>(smem + lay.off_raw);
  double* loc = reinterpret_cast<double*>(smem + lay.off_loc);
  double* loci = reinterpret_cast<double*>(smem + lay.off_loci);
  double* tgt = reinterpret_cast<double*>(smem + lay.off_tgt);
  float2* fA = reinterpret_cast<float2*>(smem + lay.off_fa);
  float2* fB = reinterpret_cast<float2*>(smem + lay.off_fb);
  float* fI = reinterpret_cast<float*>(smem + lay.off_fi);
  int* keep = reinterpret_cast<int*>(smem + lay.off_keep);
  uint2* cand = reinterpret_cast<uint2*>(smem + lay.off_cand);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockDim.x, NW = T >> 5;

  // A = first position stream with a weight, B = the second one (DUAL only)
  const int sA = p.use_vo ? 0 : 1;
  const double wA64 = p.use_vo ? p.w_vo : p.w_gps;
  const double wB64 = p.w_gps;
  const float wA = (float)wA64, wB = (float)wB64;

  auto issue_load = [&](long long w, int buf) {  // thread 0 only
    const long long start = p.win_start[w];
    int len = p.win_len[w];
    len = len < P ? len : P;
    const unsigned bytes = (unsigned)len * 16u;
    const unsigned total = bytes * (unsigned)(p.load_vo + p.load_gps);
    mbar_arrive_expect_tx(&hd->mbar[buf], total);
    if (p.load_vo) bulk_g2s(raw + (buf * 2 + 0) * P, p.vo + start, bytes, &hd->mbar[buf]);
    if (p.load_gps) bulk_g2s(raw + (buf * 2 + 1) * P, p.gps + start, bytes, &hd->mbar[buf]);
  };

  if (tid == 0) {
    mbar_init(&hd->mbar[0], 1);
    mbar_init(&hd->mbar[1], 1);
    mbar_fence_init();
    hd->count = 0;
    long long w = (long long)atomicAdd(p.work_counter, 1ULL);
    hd->wid[0] = w;
    if (w < p.n_windows) issue_load(w, 0);
  }
  __syncthreads();

  for (int it = 0;; ++it) {
    const int cur = it & 1;
    const long long w = hd->wid[cur];
    if (w >= p.n_windows) break;
    if (tid == 0) {  // prefetch the next window's poses while this one is searched
      long long wn = (long long)atomicAdd(p.work_counter, 1ULL);
      hd->wid[cur ^ 1] = wn;
      if (wn < p.n_windows) issue_load(wn, cur ^ 1);
    }
    const long long start = p.win_start[w];
    const int len = p.win_len[w];
    const double dt = p.dt_drive[p.win_drive[w]];
    mbar_wait(&hd->mbar[cur], (unsigned)((it >> 1) & 1));

    vmvo_window_result res;
    res.best_idx = -1;
    res.n_steps = 0;
    res.status = 0;
    res.n_rescored = 0;
    res.best_cost = CUDART_NAN;
    res.v_seed = res.s_seed = CUDART_NAN;
    res.x1 = res.y1 = res.theta1 = CUDART_NAN;

    if (len > P || len < 1) {  // uniform branch
      res.status = VMVO_WIN_TOO_LONG;
      if (tid == 0) p.results[w] = res;
      __syncthreads();
      continue;
    }

    // ---- phase A: local frames (a7), seeds, decimation (a9), targets ----------------------
    for (int s = 0; s < 2; ++s) {
      if (!(s == 0 ? p.load_vo : p.load_gps)) continue;
      const float4* rs = raw + (cur * 2 + s) * P;
      const float4 p0 = rs[0];
      const double th0 = (double)p0.z;
      double sn, cs;
      sincos(th0, &sn, &cs);
      double* lx = loc + (s * 3 + 0) * P;
      double* ly = loc + (s * 3 + 1) * P;
      double* lt = loc + (s * 3 + 2) * P;
      for (int m = tid; m < len; m += T) {
        const float4 q = rs[m];
        const double dx = dsub((double)q.x, (double)p0.x);
        const double dy = dsub((double)q.y, (double)p0.y);
        lx[m] = dadd(dmul(dx, cs), dmul(dy, sn));
        ly[m] = dadd(dmul(-dx, sn), dmul(dy, cs));
        lt[m] = dsub((double)q.z, th0);
      }
    }
    if (IMU) {
      const double y0 = (double)p.imu[start];
      for (int m = tid; m < len; m += T) loci[m] = dsub((double)p.imu[start + m], y0);
    }
    __syncthreads();

    const float4* rp = raw + (cur * 2 + p.primary) * P;
    const double* plx = loc + (p.primary * 3 + 0) * P;
    const double* ply = loc + (p.primary * 3 + 1) * P;
    const double* plt = loc + (p.primary * 3 + 2) * P;
    double v_seed, s_seed;
    if (p.seed_mode == VMVO_SEED_GIVEN) {
      v_seed = p.seeds[2 * w];
      s_seed = p.seeds[2 * w + 1];
    } else {
      v_seed = ddiv(dadd((double)rp[0].w, (double)rp[len - 1].w), 2.0);
      s_seed = 0.0;
      if (len >= 2 && dmul(v_seed, dt) > 1e-6) {
        const double dth = remainder(dsub(plt[1], plt[0]), kTwoPi);
        const double ang = atan(ddiv(dmul(p.L, dth), dmul(v_seed, dt)));
        s_seed = dmul(dmul(ang, kRadToDeg), p.ratio);
        s_seed = s_seed < -p.max_steer ? -p.max_steer : s_seed;
        s_seed = s_seed > p.max_steer ? p.max_steer : s_seed;
      }
    }

    if (p.target_mode == VMVO_TARGET_TRAVERSE) {
      if (tid == 0) {  // sequential by definition (distance accumulator with reset)
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
        hd->wi.n_targets = cnt;
      }
    } else {
      for (int m = tid; m < len; m += T) keep[m] = m;
      if (tid == 0) hd->wi.n_targets = len;
    }
    __syncthreads();

    const int n_targets = hd->wi.n_targets;
    const int N = n_targets - 1;
    bool finite = isfinite(v_seed) && isfinite(s_seed) && isfinite(dt);
    float tmax = 0.f, imax = 0.f;
    {
      const double* aX = loc + (sA * 3 + 0) * P;
      const double* aY = loc + (sA * 3 + 1) * P;
      const double* bX = loc + (1 * 3 + 0) * P;
      const double* bY = loc + (1 * 3 + 1) * P;
      for (int q = tid; q < n_targets; q += T) {
        const int m = keep[q];
        const double ax = aX[m], ay = aY[m];
        tgt[q] = ax;
        tgt[P + q] = ay;
        fA[q] = make_float2((float)ax, (float)ay);
        finite = finite && isfinite(ax) && isfinite(ay);
        tmax = fmaxf(tmax, fmaxf(fabsf((float)ax), fabsf((float)ay)));
        if (DUAL) {
          const double bx = bX[m], by = bY[m];
          tgt[2 * P + q] = bx;
          tgt[3 * P + q] = by;
          fB[q] = make_float2((float)bx, (float)by);
          finite = finite && isfinite(bx) && isfinite(by);
          tmax = fmaxf(tmax, fmaxf(fabsf((float)bx), fabsf((float)by)));
        }
        if (IMU) {
          const double yi = loci[m];
          tgt[4 * P + q] = yi;
          fI[q] = (float)yi;
          finite = finite && isfinite(yi);
          imax = fmaxf(imax, fabsf((float)yi));
        }
      }
    }
    {
      // block max of |target| (for the band) and the non-finite flag
      float wm = tmax, wi_ = imax;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wm = fmaxf(wm, __shfl_xor_sync(FULL, wm, o));
        wi_ = fmaxf(wi_, __shfl_xor_sync(FULL, wi_, o));
      }
      if (lane == 0) {
        hd->red[warp] = wm;
        hd->red[16 + warp] = wi_;
      }
    }
    const int bad = __syncthreads_or(finite ? 0 : 1);
    {
      float wm = 0.f, wi_ = 0.f;
      for (int q = 0; q < NW; ++q) {
        wm = fmaxf(wm, hd->red[q]);
        wi_ = fmaxf(wi_, hd->red[16 + q]);
      }
      tmax = wm;
      imax = wi_;
    }
    if (tid == 0) {
      hd->wi.v_seed = v_seed;
      hd->wi.s_seed = s_seed;
      hd->wi.dt = dt;
      hd->wi.start = start;
      hd->wi.len = len;
      hd->wi.n_steps = N > 0 ? N : 0;
      hd->wi.status = (N <= 0 ? VMVO_WIN_EMPTY : 0) | (bad ? VMVO_WIN_NONFINITE : 0);
    }
    __syncthreads();  // wi visible; red[] free for reuse
    const WinInfo wi = hd->wi;
    res.n_steps = wi.n_steps;
    res.status = wi.status;
    res.v_seed = v_seed;
    res.s_seed = s_seed;

    if (wi.status & VMVO_WIN_EMPTY) {
      if (tid == 0) p.results[w] = res;
      __syncthreads();
      continue;
    }

    int best_h = -1;
    double best_cost = CUDART_INF;
    Pose<double> best_first{CUDART_NAN, CUDART_NAN, CUDART_NAN};
    int n_rescored = 0;

    if (wi.status & VMVO_WIN_NONFINITE) {
      // every hypothesis costs NaN or Inf alike: np.argmin returns index 0
      best_h = 0;
      best_cost = CUDART_NAN;
    } else {
      // ---- phase B: FP32 scan of the whole grid, candidates within the error band -------
      const Band band = make_band(p, N, dt, v_seed, (double)tmax, (double)imax);
      float U = CUDART_INF_F;          // upper bound on the true minimum cost
      float Uw = CUDART_INF_F;         // this warp's tightened copy (after float64 re-scores)

      auto process_list = [&]() {
        const int count = hd->count < kCandCap ? hd->count : kCandCap;
        for (int e = warp; e < count; e += NW) {
          const uint2 ce = cand[e];
          const float j32 = __uint_as_float(ce.y);
          if (band.lower(j32) > fminf(U, Uw)) continue;  // warp-uniform
          const int h = (int)ce.x;
          Pose<double> first;
          const double c64 = warp_cost64<DUAL, IMU>(p, wi, tgt, P, h, lane, wA64, wB64, &first);
          ++n_rescored;
          if (best_h < 0 || c64 < best_cost || (c64 == best_cost && h < best_h)) {
            best_h = h;
            best_cost = c64;
            best_first = first;
          }
          // a float64 cost is itself an upper bound on the minimum (rounded up to float)
          Uw = fminf(Uw, __double2float_ru(c64));
        }
        __syncthreads();
        if (tid == 0) hd->count = 0;
        __syncthreads();
      };

      const int n_pass = (p.n_items + T - 1) / T;
      for (int pass = 0; pass < n_pass; ++pass) {
        const int q = pass * T + tid;
        float Jt[C];
        int ic = 0, j = 0;
        unsigned valid = 0;
        if (q < p.n_items) {
          ic = q / p.gs;
          j = q - ic * p.gs;
          scan_item<C, DUAL, IMU>(p, wi, ic, j, fA, fB, fI, wA, wB, Jt);
#pragma unroll
          for (int c = 0; c < C; ++c)
            if (ic * C + c < p.gv) valid |= 1u << c;
        }
        float m = CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < C; ++c)
          if ((valid >> c) & 1u) m = fminf(m, Jt[c]);  // fminf drops NaN; NaN stays a candidate below
        m = warp_min_f32_nonneg(m);
        if (lane == 0) hd->red[warp] = m;
        __syncthreads();
        float bm = lane < NW ? hd->red[lane] : CUDART_INF_F;
        bm = warp_min_f32_nonneg(bm);
        U = fminf(U, band.upper(bm));
        unsigned pend = 0;
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (((valid >> c) & 1u) && !(band.lower(Jt[c]) > U)) pend |= 1u << c;
        for (;;) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if ((pend >> c) & 1u) {
              const int slot = atomicAdd(&hd->count, 1);
              if (slot < kCandCap) {
                cand[slot] = make_uint2((unsigned)((ic * C + c) * p.gs + j), __float_as_uint(Jt[c]));
                pend &= ~(1u << c);
              }
            }
          }
          const int overflow = __syncthreads_or(pend != 0);
          if (!overflow) break;
          process_list();
        }
      }
      process_list();
    }

    // ---- phase D: winner across warps, result record, optional rollout outputs ----------
    if (lane == 0) {
      hd->bcost[warp] = best_cost;
      hd->bh[warp] = best_h;
      hd->bpose[warp][0] = best_first.x;
      hd->bpose[warp][1] = best_first.y;
      hd->bpose[warp][2] = best_first.th;
      hd->nres[warp] = n_rescored;
    }
    __syncthreads();
    if (tid == 0) {
      int bw = -1;
      int total = 0;
      for (int q = 0; q < NW; ++q) {
        total += hd->nres[q];
        if (hd->bh[q] < 0) continue;
        if (bw < 0 || hd->bcost[q] < hd->bcost[bw] ||
            (hd->bcost[q] == hd->bcost[bw] && hd->bh[q] < hd->bh[bw]))
          bw = q;
      }
      if (wi.status & VMVO_WIN_NONFINITE) bw = 0;
      res.n_rescored = total;
      if (bw >= 0) {
        res.best_idx = hd->bh[bw];
        res.best_cost = hd->bcost[bw];
        res.x1 = hd->bpose[bw][0];
        res.y1 = hd->bpose[bw][1];
        res.theta1 = hd->bpose[bw][2];
      }
      hd->winner = res.best_idx;
      p.results[w] = res;
    }
    if (p.out_poses || p.out_steer || p.out_vel) {
      __syncthreads();
      const int h = hd->winner;
      if (warp == 0 && h >= 0) {
        const int i = h / p.gs, j = h - i * p.gs;
        GridCtl g{wi.v_seed, wi.s_seed, wi.dt, grid_rate(p.max_accel, i, p.gv),
                  grid_rate(p.max_rate, j, p.gs), p.max_steer};
        Pose<double> carry{0.0, 0.0, 0.0};
        const bool nonfinite = (wi.status & VMVO_WIN_NONFINITE) != 0;
        for (int base = 0; base < N; base += 32) {
          const int k = base + lane + 1;
          const bool active = k <= N;
          double v = 0.0, s = 0.0;
          if (active) g.at(k, &v, &s);
          Pose<double> pz = warp_model_round<double>(nonfinite ? 0.0 : v, s, active && !nonfinite,
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
    __syncthreads();
  }
}

template <int C, bool DUAL, bool IMU>
static int launch_search(vmvo_ctx* ctx, const SearchParams& p, int threads, cudaStream_t st) {
  auto kern = vmvo_window_search_kernel<C, DUAL, IMU>;
  const SmemLayout lay(p.maxp);
  VMVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total));
  int per_sm = 0;
  VMVO_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, lay.total));
  if (per_sm < 1) return fail(ctx, VMVO_ERR_CUDA, "search kernel does not fit on an SM");
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > p.n_windows) grid = p.n_windows;
  kern<<<(unsigned)grid, threads, lay.total, st>>>(p);
  return check_launch(ctx, "vmvo_window_search_kernel");
}

}  // namespace vmvo

using namespace vmvo;

extern "C" int vmvo_grid_search_f32(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                                    const int64_t* d_win_start, const int32_t* d_win_len,
                                    const int32_t* d_win_drive, const double* d_dt_per_drive,
                                    const float* d_vo, const float* d_gps, const float* d_imu,
                                    const double* d_seeds, vmvo_window_result* d_results,
                                    double* d_out_poses, double* d_out_steer, double* d_out_vel,
                                    int32_t out_stride, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  int rc = validate_cfg(ctx, cfg);
  if (rc) return rc;
  if (n_windows < 0) return fail(ctx, VMVO_ERR_BAD_ARG, "n_windows < 0");
  if (n_windows == 0) return VMVO_OK;
  if (!d_win_start || !d_win_len || !d_win_drive || !d_dt_per_drive || !d_results)
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL window plan / dt / result pointer");
  if (cfg->seed_mode == VMVO_SEED_CHAINED)
    return fail(ctx, VMVO_ERR_UNSUPPORTED,
                "seed_mode chained serialises the windows of a drive; use vmvo_grid_search_chained_f32");
  if (cfg->seed_mode == VMVO_SEED_GIVEN && !d_seeds)
    return fail(ctx, VMVO_ERR_BAD_ARG, "seed_mode given needs d_seeds");
  const bool use_vo = cfg->w_vo != 0, use_gps = cfg->w_gps != 0, use_imu = cfg->w_imu != 0;
  if (!use_vo && !use_gps)
    return fail(ctx, VMVO_ERR_UNSUPPORTED, "at least one of w_vo, w_gps must be non-zero");
  const bool load_vo = use_vo || cfg->primary == VMVO_PRIMARY_VO;
  const bool load_gps = use_gps || cfg->primary == VMVO_PRIMARY_GPS;
  if (load_vo && !d_vo) return fail(ctx, VMVO_ERR_BAD_ARG, "VO stream needed but d_vo is NULL");
  if (load_gps && !d_gps) return fail(ctx, VMVO_ERR_BAD_ARG, "GPS stream needed but d_gps is NULL");
  if (use_imu && !d_imu) return fail(ctx, VMVO_ERR_BAD_ARG, "w_imu != 0 but d_imu is NULL");
  if (((uintptr_t)d_vo | (uintptr_t)d_gps) & 15)
    return fail(ctx, VMVO_ERR_BAD_ARG, "pose streams must be 16-byte aligned");
  if ((d_out_poses || d_out_steer || d_out_vel) && out_stride < 1)
    return fail(ctx, VMVO_ERR_BAD_ARG, "out_stride < 1");

  const int C = cfg->grid_v >= 8 ? 8 : 4;
  SearchParams p;
  p.gv = cfg->grid_v;
  p.gs = cfg->grid_s;
  p.n_ic = (p.gv + C - 1) / C;
  p.n_items = p.n_ic * p.gs;
  p.target_mode = cfg->target_mode;
  p.target_offset = cfg->target_offset;
  p.seed_mode = cfg->seed_mode;
  p.primary = cfg->primary;
  p.maxp = cfg->max_window_poses;
  p.use_vo = use_vo;
  p.use_gps = use_gps;
  p.load_vo = load_vo;
  p.load_gps = load_gps;
  p.w_vo = cfg->w_vo;
  p.w_gps = cfg->w_gps;
  p.w_imu = cfg->w_imu;
  p.k_steer = cfg->k_steer;
  p.L = cfg->wheel_base;
  p.ratio = cfg->steering_ratio;
  p.max_steer = cfg->max_steer;
  p.max_accel = cfg->max_accel;
  p.max_rate = cfg->max_steer_rate;
  p.delta_max = cfg->max_steer * kDegToRad / cfg->steering_ratio;
  p.tan_max = tan(p.delta_max);
  p.win_start = (const long long*)d_win_start;
  p.win_len = d_win_len;
  p.win_drive = d_win_drive;
  p.dt_drive = d_dt_per_drive;
  p.vo = (const float4*)d_vo;
  p.gps = (const float4*)d_gps;
  p.imu = d_imu;
  p.seeds = d_seeds;
  p.results = d_results;
  p.out_poses = d_out_poses;
  p.out_steer = d_out_steer;
  p.out_vel = d_out_vel;
  p.out_stride = out_stride;
  p.n_windows = n_windows;

  cudaStream_t st = (cudaStream_t)stream;
  VMVO_CUDA(ctx, cudaSetDevice(ctx->device));
  // one of 64 queue heads per launch so launches on different streams do not share one
  unsigned long long* counter = ctx->d_work_counter + (ctx->launches & 63);
  VMVO_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
  p.work_counter = counter;

  int threads = ((p.n_items + 31) / 32) * 32;
  threads = threads < 64 ? 64 : (threads > 256 ? 256 : threads);
  const bool dual = use_vo && use_gps;
#define VMVO_LAUNCH(CC)                                                                   \
  (dual ? (use_imu ? launch_search<CC, true, true>(ctx, p, threads, st)                   \
                   : launch_search<CC, true, false>(ctx, p, threads, st))                 \
        : (use_imu ? launch_search<CC, false, true>(ctx, p, threads, st)                  \
                   : launch_search<CC, false, false>(ctx, p, threads, st)))
  return C == 8 ? VMVO_LAUNCH(8) : VMVO_LAUNCH(4);
#undef VMVO_LAUNCH
}
