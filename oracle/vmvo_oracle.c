/*
 * C restatement of the VMVO window search -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Same arithmetic, in the same order, as oracle/vmvo_oracle.py (which is pinned to the
 * unmodified reference by oracle/make_golden.py): float64 throughout, one hypothesis at a
 * time, sequential in the step index like the reference's own loop.
 *
 *   bicycle step        vmvo/bicycle_model.py:66-75
 *   local frame         vmvo/schema.py:64-107
 *   decimation          vmvo/utils/mpc.py:125-141
 *   cost                vmvo/utils/mpc.py:68-80  (+ DESIGN.md 2.3 for GPS / IMU / K terms)
 *   seeds               vmvo/scripts/optimize_trajectory_v2.py:61-63 (+ DESIGN.md 2.2)
 *
 * Only tests/ and bench.py's CPU-baseline legs may load the resulting library
 * (oracle/_build/libvmvo_oracle.so).  Windows are spread over OpenMP threads; it doubles
 * as the "all host cores" CPU baseline.  Built with -ffp-contract=off: no fused
 * multiply-adds, like NumPy.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/vmvo_b200.h" /* struct layouts only: vmvo_search_cfg, vmvo_window_result */

#define PI 3.14159265358979323846
#define TWO_PI (2 * PI)

static double grid_rate(double limit, int idx, int g) {
  if (g <= 1) return 0.0;
  return (limit * (double)(2 * idx - (g - 1))) / (double)(g - 1);
}

static void local_frame(const float* s, int len, double* lx, double* ly, double* lth) {
  const double x0 = s[0], y0 = s[1], th0 = s[2];
  const double c = cos(th0), sn = sin(th0);
  for (int m = 0; m < len; ++m) {
    const double dx = (double)s[4 * m] - x0, dy = (double)s[4 * m + 1] - y0;
    lx[m] = dx * c + dy * sn;
    ly[m] = -dx * sn + dy * c;
    lth[m] = (double)s[4 * m + 2] - th0;
  }
}

static int all_finite(const double* a, int n) {
  for (int i = 0; i < n; ++i)
    if (!isfinite(a[i])) return 0;
  return 1;
}

static void solve_window(const vmvo_search_cfg* cfg, int64_t start, int len, double dt,
                         const float* vo, const float* gps, const float* imu, const double* seed,
                         vmvo_window_result* out, double* scratch) {
  const int P = len;
  double* plx = scratch;            /* primary local frame */
  double* ply = plx + P;
  double* plt = ply + P;
  double* slx = plt + P;            /* secondary local frame */
  double* sly = slx + P;
  double* slt = sly + P;
  double* tax = slt + P;            /* targets */
  double* tay = tax + P;
  double* tbx = tay + P;
  double* tby = tbx + P;
  double* tim = tby + P;
  double* V = tim + P;              /* [N][gv] */
  const int gv = cfg->grid_v, gs = cfg->grid_s;
  double* S = V + (size_t)P * gv;   /* [N][gs] */
  double* TD = S + (size_t)P * gs;  /* tan(delta) [N][gs] */
  int* keep = (int*)(TD + (size_t)P * gs);

  out->best_idx = -1;
  out->n_steps = 0;
  out->status = 0;
  out->n_rescored = 0;
  out->best_cost = NAN;
  out->x1 = out->y1 = out->theta1 = NAN;

  const float* prim = (cfg->primary == VMVO_PRIMARY_VO ? vo : gps) + 4 * start;
  const float* sec = (cfg->primary == VMVO_PRIMARY_VO ? gps : vo);
  local_frame(prim, len, plx, ply, plt);

  double v_seed, s_seed;
  if (cfg->seed_mode == VMVO_SEED_GIVEN) {
    v_seed = seed[0];
    s_seed = seed[1];
  } else {
    v_seed = ((double)prim[3] + (double)prim[4 * (len - 1) + 3]) / 2;
    s_seed = 0.0;
    if (len >= 2 && v_seed * dt > 1e-6) {
      const double dth = remainder(plt[1] - plt[0], TWO_PI);
      const double ang = atan(cfg->wheel_base * dth / (v_seed * dt));
      s_seed = (ang * (180.0 / PI)) * cfg->steering_ratio;
      s_seed = fmin(cfg->max_steer, fmax(-cfg->max_steer, s_seed));
    }
  }
  out->v_seed = v_seed;
  out->s_seed = s_seed;

  int nt;
  if (cfg->target_mode == VMVO_TARGET_TRAVERSE) {
    const double D = v_seed * dt;
    nt = 0;
    keep[nt++] = 0;
    double dist = 0.0;
    for (int i = 1; i < len; ++i) {
      const double ddx = plx[i] - plx[i - 1], ddy = ply[i] - ply[i - 1];
      const double seg = sqrt(ddx * ddx + ddy * ddy);
      if (dist + seg > D) {
        keep[nt++] = i - 1;
        dist = seg;
      } else {
        dist += seg;
      }
    }
  } else {
    nt = len;
    for (int m = 0; m < len; ++m) keep[m] = m;
  }
  const int N = nt - 1;
  const int use_vo = cfg->w_vo != 0, use_gps = cfg->w_gps != 0, use_imu = cfg->w_imu != 0;
  const int dual = use_vo && use_gps;
  int finite = isfinite(v_seed) && isfinite(s_seed);
  /* A = first weighted position stream, B = the second (vo then gps) */
  const int a_is_vo = use_vo;
  const double wA = use_vo ? cfg->w_vo : cfg->w_gps, wB = cfg->w_gps;
  const int a_is_prim = (a_is_vo && cfg->primary == VMVO_PRIMARY_VO) || (!a_is_vo && cfg->primary == VMVO_PRIMARY_GPS);
  if ((!a_is_prim || dual) && sec) local_frame(sec + 4 * start, len, slx, sly, slt);
  for (int q = 0; q < nt; ++q) {
    const int m = keep[q];
    tax[q] = a_is_prim ? plx[m] : slx[m];
    tay[q] = a_is_prim ? ply[m] : sly[m];
    if (dual) { /* B is gps */
      tbx[q] = cfg->primary == VMVO_PRIMARY_GPS ? plx[m] : slx[m];
      tby[q] = cfg->primary == VMVO_PRIMARY_GPS ? ply[m] : sly[m];
    }
    if (use_imu) tim[q] = (double)imu[start + m] - (double)imu[start];
  }
  finite = finite && all_finite(tax, nt) && all_finite(tay, nt);
  if (dual) finite = finite && all_finite(tbx, nt) && all_finite(tby, nt);
  if (use_imu) finite = finite && all_finite(tim, nt);

  if (N <= 0) {
    out->status |= VMVO_WIN_EMPTY;
    if (!finite) out->status |= VMVO_WIN_NONFINITE;
    return;
  }
  out->n_steps = N;
  for (int k = 1; k <= N; ++k) {
    const double t = (double)k * dt;
    for (int i = 0; i < gv; ++i) {
      const double vv = v_seed + grid_rate(cfg->max_accel, i, gv) * t;
      V[(size_t)(k - 1) * gv + i] = vv > 0.0 ? vv : (vv == vv ? 0.0 : vv);
    }
    for (int j = 0; j < gs; ++j) {
      double ss = s_seed + grid_rate(cfg->max_steer_rate, j, gs) * t;
      ss = fmin(cfg->max_steer, fmax(-cfg->max_steer, ss));
      if (s_seed != s_seed) ss = s_seed;
      S[(size_t)(k - 1) * gs + j] = ss;
      TD[(size_t)(k - 1) * gs + j] = tan((ss * (PI / 180.0)) / cfg->steering_ratio);
    }
  }
  int best = 0;
  double best_cost = NAN;
  if (!finite) {
    out->status |= VMVO_WIN_NONFINITE;
  } else {
    const int off = cfg->target_offset;
    const double L = cfg->wheel_base;
    best = -1;
    for (int i = 0; i < gv; ++i) {
      for (int j = 0; j < gs; ++j) {
        double th = 0.0, x = 0.0, y = 0.0, cost = 0.0;
        for (int k = 1; k <= N; ++k) {
          const double v = V[(size_t)(k - 1) * gv + i];
          th = th + (v / L * TD[(size_t)(k - 1) * gs + j] * dt);
          x = x + (v * cos(th) * dt);
          y = y + (v * sin(th) * dt);
          const int t = k - off;
          double ex = x - tax[t], ey = y - tay[t];
          double e = ex * ex + ey * ey;
          double term = wA == 1.0 ? e : wA * e;
          if (dual) {
            ex = x - tbx[t];
            ey = y - tby[t];
            e = ex * ex + ey * ey;
            term = term + (wB == 1.0 ? e : wB * e);
          }
          if (use_imu) {
            const double d = remainder(th - tim[t], TWO_PI);
            term = term + cfg->w_imu * (d * d);
          }
          if (cfg->k_steer != 0.0) {
            const double s = S[(size_t)(k - 1) * gs + j];
            term = term + cfg->k_steer * (s * s);
          }
          cost = cost + term;
        }
        if (best < 0 || cost < best_cost) { /* strict: the lowest index wins ties */
          best = i * gs + j;
          best_cost = cost;
        }
      }
    }
  }
  out->best_idx = best;
  out->best_cost = best_cost;
  if (finite) { /* first pose of the best rollout */
    const int i = best / gs, j = best % gs;
    const double v = V[i];
    const double th = 0.0 + (v / cfg->wheel_base * TD[j] * dt);
    out->theta1 = th;
    out->x1 = 0.0 + (v * cos(th) * dt);
    out->y1 = 0.0 + (v * sin(th) * dt);
  }
}

/* Every window [win_start[w], +win_len[w]) of one or more concatenated drives.
 * dt_per_drive / win_drive as in vmvo_grid_search_f32.  Returns the number of
 * hypothesis-steps evaluated (the unit of the throughput metric). */
int64_t vmvo_oracle_search(const vmvo_search_cfg* cfg, int64_t n_windows, const int64_t* win_start,
                           const int32_t* win_len, const int32_t* win_drive,
                           const double* dt_per_drive, const float* vo, const float* gps,
                           const float* imu, const double* seeds, vmvo_window_result* results,
                           int n_threads) {
  int max_len = 2;
  for (int64_t w = 0; w < n_windows; ++w)
    if (win_len[w] > max_len) max_len = win_len[w];
  const size_t per = (size_t)max_len * (11 + cfg->grid_v + 2 * (size_t)cfg->grid_s) * sizeof(double) +
                     (size_t)max_len * sizeof(int) + 64;
  int64_t total = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : total)
  {
    double* scratch = (double*)malloc(per);
#pragma omp for schedule(dynamic, 4)
    for (int64_t w = 0; w < n_windows; ++w) {
      solve_window(cfg, win_start[w], win_len[w], dt_per_drive[win_drive[w]], vo, gps, imu,
                   seeds ? seeds + 2 * w : NULL, &results[w], scratch);
      total += (int64_t)results[w].n_steps * cfg->grid_v * cfg->grid_s;
    }
    free(scratch);
  }
  return total;
}

int vmvo_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
