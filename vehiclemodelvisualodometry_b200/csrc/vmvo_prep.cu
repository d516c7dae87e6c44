// Trajectory pre-processing, the step before the window search:
// process_vo_trajectory (vmvo/utils/trajectory.py:13-65), smoothen_traj (:68-99),
// geodetic_to_euclidean (:120-174) and process_gps_trajectory (:177-335) of the reference,
// batched over drives.  Float64 throughout, products and sums in the reference's order with
// explicit round-to-nearest intrinsics (no FMA contraction); only sin/cos/atan2 come from CUDA's
// libm (<= 2 ulp from glibc's).  Memory-bound streaming kernels plus two short sequential scans
// per drive (cumulative path, de-duplication state machine) that are serial by definition.
#include "vmvo_device.cuh"
#include "vmvo_internal.h"

namespace vmvo {

__device__ __forceinline__ int seg_of(const long long* offsets, int n_seg, long long idx) {
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (offsets[mid] <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

// mean of the last min(i+1, w) points, summed left to right from zero (Python's sum);
// a drive with n <= w points is returned unchanged (vmvo/utils/trajectory.py:83-84)
__device__ __forceinline__ void trailing_mean(const double* x, const double* y, long long i,
                                              long long n, int w, double* ox, double* oy) {
  if (n <= w) { *ox = x[i]; *oy = y[i]; return; }
  const long long lo = i - w + 1 > 0 ? i - w + 1 : 0;
  double sx = 0.0, sy = 0.0;
  for (long long q = lo; q <= i; ++q) { sx = dadd(sx, x[q]); sy = dadd(sy, y[q]); }
  const double cnt = (double)(i + 1 - lo);
  *ox = ddiv(sx, cnt);
  *oy = ddiv(sy, cnt);
}

__global__ void smooth_kernel(int n_drives, const long long* off, long long total, const double* x,
                              const double* y, int w, double scale_x, double scale_y, double* ox,
                              double* oy) {
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total;
       f += (long long)gridDim.x * blockDim.x) {
    const int d = seg_of(off, n_drives, f);
    const long long f0 = off[d], n = off[d + 1] - f0;
    double sx, sy;
    trailing_mean(x + f0, y + f0, f - f0, n, w, &sx, &sy);
    ox[f] = dmul(sx, scale_x);
    oy[f] = dmul(sy, scale_y);
  }
}

// process_vo_trajectory: yaw from the rotation, speed from RAW positions over a millisecond
// difference (quirk kept), stamps in seconds
__global__ void vo_prepare_kernel(int n_drives, const long long* off, long long total, const double* x,
                                  const double* y, const double* rot, const double* stamp, int yaw_f32,
                                  double* theta, double* vel, double* time) {
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total;
       f += (long long)gridDim.x * blockDim.x) {
    const int d = seg_of(off, n_drives, f);
    const long long f0 = off[d];
    // np.arctan2(rot[1, 0], rot[0, 0]) (trajectory.py:28): float32 scalars when the rotation came
    // from the cached trajectory (bdd_raw.py:163-164), and then the result is a float32 too
    theta[f] = yaw_f32 ? (double)atan2f((float)rot[f * 9 + 3], (float)rot[f * 9 + 0])
                       : atan2(rot[f * 9 + 3], rot[f * 9 + 0]);
    double v = 0.0;
    if (f > f0) {
      const double ddx = dsub(x[f - 1], x[f]), ddy = dsub(y[f - 1], y[f]);
      v = ddiv(sqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy))), dsub(stamp[f], stamp[f - 1]));
    }
    vel[f] = v;
    time[f] = ddiv(stamp[f], 1000.0);
  }
}

// one ECEF point, operation order of vmvo/utils/trajectory.py:127-144
__device__ __forceinline__ void ecef_xy(double lat_deg, double lon_deg, double* X, double* Y) {
  const double a = 6378137.0, e = 8.1819190842622e-2;
  const double lat = dmul(lat_deg, kDegToRad), lon = dmul(lon_deg, kDegToRad);
  double sl, cl, so, co;
  sincos(lat, &sl, &cl);
  sincos(lon, &so, &co);
  const double den = sqrt(dsub(1.0, dmul(dmul(e, e), dmul(sl, sl))));
  const double r = dmul(ddiv(a, den), cl);
  *X = dmul(r, co);
  *Y = dmul(r, so);
}

__global__ void gps_delta_kernel(int n_drives, const long long* off, long long total, const double* lat,
                                 const double* lon, const double* stamp, double* dxy /* [2][total] */,
                                 double* time) {
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < total;
       f += (long long)gridDim.x * blockDim.x) {
    const int d = seg_of(off, n_drives, f);
    const long long prev = f > off[d] ? f - 1 : f;   // the first fix is differenced with itself
    double x1, y1, x2, y2;
    ecef_xy(lat[prev], lon[prev], &x1, &y1);
    ecef_xy(lat[f], lon[f], &x2, &y2);
    dxy[f] = dsub(x2, x1);
    dxy[total + f] = dsub(y2, y1);
    time[f] = ddiv(stamp[f], 1000.0);
  }
}

// per drive: the cumulative path (n+1 points, a leading duplicate of the origin) -- a dependent
// chain of additions, sequential by definition -- and the de-duplication state machine
// (vmvo/utils/trajectory.py:206-216, 243-300).  One warp per drive: 32 deltas per axis are loaded
// coalesced and staged in shared memory; lane 0 (x) and lane 1 (y) pull their 32 deltas into
// registers and run the chain -- 32 dependent DADDs, nothing else on the critical path -- leaving
// the 32 points in shared memory for the warp (a chain step costs one DADD latency; walking the
// chain redundantly in every lane with a shuffle per step measured 88 clk per step).  The state
// machine needs no chain: every point between `last` and i equals point `last` bit for bit, so
// "x[last] != x[i]" is the local test "point i differs from point i-1"; segment ends are then found
// with ballot / ffs / clz, and only runs of repeated fixes longer than a batch need the strided fill.
constexpr int kGpsScanWarps = 4;
__global__ void __launch_bounds__(32 * kGpsScanWarps)
gps_scan_kernel(int n_drives, const long long* off, long long total, const double* dxy,
                double* X, double* Y, int* seg_lo, int* seg_hi, int* status) {
  __shared__ __align__(16) double s_delta[kGpsScanWarps][2][32];
  __shared__ __align__(16) double s_point[kGpsScanWarps][2][32];
  const int lane = threadIdx.x & 31, wb = threadIdx.x >> 5;
  const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (d >= n_drives) return;
  const long long f0 = off[d], n = off[d + 1] - f0;
  const long long o0 = f0 + d;          // outputs hold n + 1 points per drive
  if (lane == 0) {
    X[o0] = 0.0;
    Y[o0] = 0.0;
    seg_lo[o0] = 0;
    seg_hi[o0] = 0;
  }
  double carry = 0.0;                   // lanes 0 / 1: last x / y produced
  double x_last = 0.0, y_last = 0.0;    // every lane: point `base`
  long long last = 0;                   // latest segment end
  int st = 0;
  // the deltas of the next batch are fetched while this batch's chain runs (a global load costs
  // more than the 32 additions it feeds)
  double nx = lane < n ? dxy[f0 + lane] : 0.0, ny = lane < n ? dxy[total + f0 + lane] : 0.0;
  for (long long base = 0; base < n; base += 32) {
    const long long i0 = base + lane;
    const bool valid = i0 < n;
    s_delta[wb][0][lane] = nx;          // padding deltas are zero
    s_delta[wb][1][lane] = ny;
    {
      const long long i1 = i0 + 32;
      nx = i1 < n ? dxy[f0 + i1] : 0.0;
      ny = i1 < n ? dxy[total + f0 + i1] : 0.0;
    }
    __syncwarp();
    if (lane < 2) {                     // x[i+1] = dx_i + x[i], one axis per lane
      double v[32];
      const double2* src = reinterpret_cast<const double2*>(s_delta[wb][lane]);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double2 t = src[q];
        v[2 * q] = t.x;
        v[2 * q + 1] = t.y;
      }
      double2* dst = reinterpret_cast<double2*>(s_point[wb][lane]);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double a = dadd(v[2 * q], carry);
        carry = dadd(v[2 * q + 1], a);
        dst[q] = make_double2(a, carry);
      }
    }
    __syncwarp();
    const double mx = s_point[wb][0][lane], my = s_point[wb][1][lane];
    // lane l holds point i = base + l + 1; its predecessor is lane l-1's point (or point `base`)
    double px = __shfl_up_sync(FULL, mx, 1), py = __shfl_up_sync(FULL, my, 1);
    if (lane == 0) { px = x_last; py = y_last; }
    x_last = __shfl_sync(FULL, mx, 31);   // (padding deltas are zero: lane 31 holds the last valid point)
    y_last = __shfl_sync(FULL, my, 31);
    const long long i = base + lane + 1;
    const unsigned bits = __ballot_sync(FULL, valid && (mx != px || my != py));
    if (valid) {
      X[o0 + i] = mx;
      Y[o0 + i] = my;
      const unsigned ahead = bits >> lane, behind = bits & ((1u << lane) - 1u);
      seg_lo[o0 + i] = behind ? (int)(base + (31 - __clz(behind)) + 1) : (int)last;
      seg_hi[o0 + i] = ahead ? (int)(i + __ffs(ahead) - 1) : -1;   // -1: open (tail unless closed later)
    }
    if (bits) {
      const long long first_end = base + __ffs(bits);              // point index of the first end
      for (long long j = last + 1 + lane; j <= base; j += 32) seg_hi[o0 + j] = (int)first_end;
      last = base + (31 - __clz(bits)) + 1;
      if (last == n) st = 1;   // the reference indexes velocity[n] here: IndexError
    }
  }
  if (lane == 0) status[d] = st;
}

__global__ void gps_interp_kernel(int n_drives, const long long* off, long long total_out,
                                  const double* X, const double* Y, const double* speed,
                                  const double* time, const int* seg_lo, const int* seg_hi,
                                  const int* status, double* xn, double* yn, double* vn, double* tn) {
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total_out;
       o += (long long)gridDim.x * blockDim.x) {
    // output index o belongs to drive d with o0 = off[d] + d
    int lo_d = 0, hi_d = n_drives;
    while (hi_d - lo_d > 1) {
      int mid = (lo_d + hi_d) >> 1;
      if (off[mid] + mid <= o) lo_d = mid; else hi_d = mid;
    }
    const int d = lo_d;
    const long long f0 = off[d], n = off[d + 1] - f0, o0 = f0 + d, m = n + 1;
    const long long j = o - o0;
    // estimated speed (vmvo/utils/trajectory.py:229-236): the PRODUCT of the squared deltas
    auto est = [&](long long i) -> double {
      if (i == 0) return speed[f0];
      const double ddx = dsub(X[o0 + i], X[o0 + i - 1]), ddy = dsub(Y[o0 + i], Y[o0 + i - 1]);
      return ddiv(sqrt(dmul(dmul(ddx, ddx), dmul(ddy, ddy))), dsub(time[f0 + i], time[f0 + i - 1]));
    };
    if (status[d]) { xn[o] = yn[o] = vn[o] = tn[o] = __longlong_as_double(0x7ff8000000000000LL); continue; }
    if (j == 0) {
      xn[o] = X[o0]; yn[o] = Y[o0]; vn[o] = est(0); tn[o] = time[f0];
      continue;
    }
    const long long a = seg_lo[o];
    const int hi = seg_hi[o];
    double alpha, xe, ye, ve, te;
    if (hi >= 0) {
      alpha = ddiv((double)(j - a), (double)(hi - a));
      xe = X[o0 + hi]; ye = Y[o0 + hi]; ve = est(hi); te = time[f0 + hi];
    } else {   // tail: ends at x[-1], velocity[-1], time[-1]
      alpha = ddiv((double)(j - a), (double)(m - a));
      xe = X[o0 + m - 1]; ye = Y[o0 + m - 1]; ve = est(n - 1); te = time[f0 + n - 1];
    }
    const double om = dsub(1.0, alpha);
    xn[o] = dadd(dmul(X[o0 + a], om), dmul(xe, alpha));
    yn[o] = dadd(dmul(Y[o0 + a], om), dmul(ye, alpha));
    vn[o] = dadd(dmul(est(a), om), dmul(ve, alpha));
    tn[o] = dadd(dmul(time[f0 + a], om), dmul(te, alpha));
  }
}

__global__ void gps_finish_kernel(int n_drives, const long long* off, long long total_out,
                                  const double* xn, const double* yn, int w, double* ox, double* oy,
                                  double* oth) {
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total_out;
       o += (long long)gridDim.x * blockDim.x) {
    int lo_d = 0, hi_d = n_drives;
    while (hi_d - lo_d > 1) {
      int mid = (lo_d + hi_d) >> 1;
      if (off[mid] + mid <= o) lo_d = mid; else hi_d = mid;
    }
    const int d = lo_d;
    const long long o0 = off[d] + d, m = off[d + 1] - off[d] + 1, j = o - o0;
    double sx, sy;
    trailing_mean(xn + o0, yn + o0, j, m, w, &sx, &sy);
    ox[o] = -sx;
    oy[o] = sy;
    if (j + 1 < m) {   // tangent heading; theta has one element fewer than x (quirk D8)
      double nx, ny;
      trailing_mean(xn + o0, yn + o0, j + 1, m, w, &nx, &ny);
      const double ang = atan2(dsub(nx, sx), dsub(ny, sy));
      oth[o] = pymod_pos(dadd(ang, kPi), kTwoPi);
    } else {
      oth[o] = __longlong_as_double(0x7ff8000000000000LL);
    }
  }
}

static inline unsigned blocks_for(long long n, int threads, int cap) {
  long long g = (n + threads - 1) / threads;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace vmvo

using namespace vmvo;

extern "C" int vmvo_smooth_f64(vmvo_ctx* ctx, int32_t n_drives, int64_t total_frames,
                               const int64_t* d_offsets, const double* d_x, const double* d_y,
                               int32_t window, double* d_out_x, double* d_out_y, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_drives < 1 || total_frames < 0 || window < 1 || !d_offsets)
    return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  if (total_frames == 0) return VMVO_OK;
  if (!d_x || !d_y || !d_out_x || !d_out_y) return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  VMVO_ON_DEVICE(ctx);
  smooth_kernel<<<blocks_for(total_frames, 256, ctx->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
      n_drives, (const long long*)d_offsets, total_frames, d_x, d_y, window, 1.0, 1.0, d_out_x, d_out_y);
  return check_launch(ctx, "smooth_kernel");
}

extern "C" int vmvo_vo_prepare_f64(vmvo_ctx* ctx, int32_t n_drives, int64_t total_frames,
                                   const int64_t* d_offsets, const double* d_x, const double* d_y,
                                   const double* d_rot, const double* d_stamp_ms, double scale,
                                   int32_t window, int32_t yaw_f32, double* d_out_x, double* d_out_y,
                                   double* d_out_theta, double* d_out_vel, double* d_out_time,
                                   void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_drives < 1 || total_frames < 0 || window < 1 || !d_offsets)
    return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  if (total_frames == 0) return VMVO_OK;
  if (!d_x || !d_y || !d_rot || !d_stamp_ms || !d_out_x || !d_out_y || !d_out_theta || !d_out_vel ||
      !d_out_time)
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  VMVO_ON_DEVICE(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = blocks_for(total_frames, 256, ctx->sm_count * 8);
  vo_prepare_kernel<<<g, 256, 0, st>>>(n_drives, (const long long*)d_offsets, total_frames, d_x, d_y,
                                      d_rot, d_stamp_ms, yaw_f32, d_out_theta, d_out_vel, d_out_time);
  int rc = check_launch(ctx, "vo_prepare_kernel");
  if (rc) return rc;
  smooth_kernel<<<g, 256, 0, st>>>(n_drives, (const long long*)d_offsets, total_frames, d_x, d_y,
                                  window, scale, scale, d_out_x, d_out_y);
  return check_launch(ctx, "smooth_kernel");
}

extern "C" int64_t vmvo_gps_prepare_scratch_bytes(int64_t total_frames, int32_t n_drives) {
  const int64_t m = total_frames + n_drives;
  return 8 * (3 * total_frames + 4 * m) + 4 * (2 * m) + 64;
}

extern "C" int vmvo_gps_prepare_f64(vmvo_ctx* ctx, int32_t n_drives, int64_t total_frames,
                                    const int64_t* d_offsets, const double* d_lat, const double* d_lon,
                                    const double* d_speed, const double* d_stamp_ms, int32_t window,
                                    void* d_scratch, double* d_out_x, double* d_out_y,
                                    double* d_out_theta, double* d_out_vel, double* d_out_time,
                                    int32_t* d_status, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_drives < 1 || total_frames < 0 || window < 1 || !d_offsets)
    return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  if (total_frames == 0) return VMVO_OK;
  if (!d_lat || !d_lon || !d_speed || !d_stamp_ms || !d_scratch || !d_out_x || !d_out_y ||
      !d_out_theta || !d_out_vel || !d_out_time || !d_status)
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  if ((uintptr_t)d_scratch & 7) return fail(ctx, VMVO_ERR_BAD_ARG, "scratch must be 8-byte aligned");
  VMVO_ON_DEVICE(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  const long long F = total_frames, M = total_frames + n_drives;
  double* dxy = (double*)d_scratch;          // [2][F]
  double* X = dxy + 2 * F;                   // [M]
  double* Y = X + M;
  double* xn = Y + M;
  double* yn = xn + M;
  double* time = yn + M;                     // [F] stamps in seconds
  int* seg_lo = (int*)(time + F);
  int* seg_hi = seg_lo + M;
  const long long* off = (const long long*)d_offsets;
  const int cap = ctx->sm_count * 8;
  gps_delta_kernel<<<blocks_for(F, 256, cap), 256, 0, st>>>(n_drives, off, F, d_lat, d_lon, d_stamp_ms,
                                                           dxy, time);
  int rc = check_launch(ctx, "gps_delta_kernel");
  if (rc) return rc;
  gps_scan_kernel<<<blocks_for((long long)n_drives * 32, 128, 1 << 30), 128, 0, st>>>(
      n_drives, off, F, dxy, X, Y, seg_lo, seg_hi, d_status);
  rc = check_launch(ctx, "gps_scan_kernel");
  if (rc) return rc;
  gps_interp_kernel<<<blocks_for(M, 256, cap), 256, 0, st>>>(n_drives, off, M, X, Y, d_speed, time,
                                                            seg_lo, seg_hi, d_status, xn, yn, d_out_vel,
                                                            d_out_time);
  rc = check_launch(ctx, "gps_interp_kernel");
  if (rc) return rc;
  gps_finish_kernel<<<blocks_for(M, 256, cap), 256, 0, st>>>(n_drives, off, M, xn, yn, window, d_out_x,
                                                            d_out_y, d_out_theta);
  return check_launch(ctx, "gps_finish_kernel");
}
