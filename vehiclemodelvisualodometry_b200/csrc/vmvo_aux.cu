// Kernels either side of the fused search: window planning (a8), write-back + blends (a12),
// batched model rollout (a1/a2), sequence cost (a10), stand-alone window operators (a7-a9),
// and the issue-rate probes bench.py uses for the roofline denominators.
#include "vmvo_device.cuh"
#include "vmvo_internal.h"

#include <math_constants.h>
#include <string.h>

namespace vmvo {

// ---- a8: window extents ---------------------------------------------------------------------
// np.searchsorted(time, t, 'left' / 'right') over one drive (vmvo/schema.py:119-120).
__device__ __forceinline__ long long lower_bound_f64(const double* t, long long n, double key) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (t[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ long long upper_bound_f64(const double* t, long long n, double key) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (key < t[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}
__device__ __forceinline__ int find_segment(const long long* offsets, int n_seg, long long idx) {
  int lo = 0, hi = n_seg;  // largest d with offsets[d] <= idx
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (offsets[mid] <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void plan_windows_kernel(int window_mode, int window_frames, double horizon_time,
                                    int n_drives, const long long* drive_off,
                                    const long long* win_off, long long n_windows,
                                    const double* time, long long* win_start, int* win_len,
                                    int* win_drive) {
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < n_windows;
       w += (long long)gridDim.x * blockDim.x) {
    const int d = find_segment(win_off, n_drives, w);
    const long long i = w - win_off[d];
    const long long f0 = drive_off[d];
    const long long n = drive_off[d + 1] - f0;
    long long s, e;
    if (window_mode == VMVO_WINDOW_FRAMES) {
      s = i;
      e = i + window_frames + 1;
      e = e < n ? e : n;
    } else {
      const double* t = time + f0;
      const double t0 = t[i];
      s = lower_bound_f64(t, n, t0);
      e = upper_bound_f64(t, n, dadd(t0, horizon_time));
    }
    win_start[w] = f0 + s;
    long long len = e - s;
    win_len[w] = (int)(len > 0x7fffffff ? 0x7fffffff : len);
    win_drive[w] = d;
  }
}

// ---- the record exchange: arrival words between the ranks of a box (vmvo_exchange) ---------------
// What a kernel needs of a vmvo_exchange.  publish: st.release.sys of the step number into the word
// this rank owns in every peer's flag array -- issued by a kernel that follows, in stream order,
// every kernel that stored records of the step, so the release is cumulative over those stores.
// wait: ld.acquire.sys on this rank's own flag array until every peer's word has reached the step.
struct ExchangeDev {
  int world, rank, n_peers;
  unsigned* peer_flags[VMVO_MAX_MIRRORS];
  const unsigned* local_flags;
  unsigned* epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kWaitTimeoutNs = 10ull * 1000 * 1000 * 1000;

__device__ __forceinline__ void exchange_publish(const ExchangeDev& ex, int t) {
  if (t < ex.n_peers) {
    const unsigned e = *reinterpret_cast<volatile unsigned*>(ex.epoch);
    __threadfence_system();
    st_release_sys(ex.peer_flags[t], e);
  }
}
// thread t < world waits for rank t's word (its own: nothing to wait for)
__device__ __forceinline__ void exchange_wait(const ExchangeDev& ex, int t) {
  if (t < ex.world && t != ex.rank) {
    const unsigned e = *reinterpret_cast<volatile unsigned*>(ex.epoch);
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(ex.local_flags + t) - e) < 0) {
      if (global_ns() - t0 > kWaitTimeoutNs) {      // a dead peer must not hang this GPU
        atomicExch(ex.epoch + 1, 1u + (unsigned)t);
        break;
      }
      __nanosleep(32);
    }
  }
}

__global__ void exchange_publish_kernel(const ExchangeDev ex) { exchange_publish(ex, threadIdx.x); }
__global__ void exchange_wait_kernel(const ExchangeDev ex) { exchange_wait(ex, threadIdx.x); }

// ---- a12: write-back and blends -----------------------------------------------------------------
// One thread per frame.  Frame m of a drive takes x,y from the LAST window that covers it:
// the largest i <= min(m, n_w - 1) with m - i < N_i (optimize_trajectory_v2.py:122-123 executed
// for i = 0 .. n_w-1 in order).  Offset m - i == 0 is the first rollout pose kept in the
// record; deeper offsets (the tail after the last window, or frames behind an empty window)
// re-run the winning hypothesis with the same warp routine the search used.
template <typename Pose4>
__global__ void write_back_kernel(int gv, int gs, double L, double ratio, double max_steer,
                                  double max_accel, double max_rate, int max_steps, int n_drives,
                                  const long long* drive_off, const long long* win_off,
                                  const double* dt_drive, const Pose4* vo, const Pose4* gps,
                                  const vmvo_window_result* results, long long frame_lo,
                                  long long frame_hi, double* out_x, double* out_y, double* out_th,
                                  double* out_v, const ExchangeDev ex) {
  // the consumer end of the record exchange: block 0 tells every peer that this rank's records of
  // the step are complete (the searches precede this kernel in the stream), then every block waits
  // for the peers' words before it reads a record
  if (ex.n_peers > 0) {
    if (blockIdx.x == 0) exchange_publish(ex, threadIdx.x);
    exchange_wait(ex, threadIdx.x);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const long long f_base = frame_lo + (blockIdx.x * (long long)blockDim.x + threadIdx.x) - lane;
  if (f_base >= frame_hi) return;  // whole warp out of range
  const long long f = f_base + lane;
  const bool in_range = f < frame_hi;
  int d = 0;
  long long m = 0, w0 = 0, nw = 0;
  double x = 0, y = 0, th = 0, v = 0;
  long long cover = -1;  // window (absolute index) that owns this frame's x,y
  int depth = 0;
  if (in_range) {
    d = find_segment(drive_off, n_drives, f);
    m = f - drive_off[d];
    w0 = win_off[d];
    nw = win_off[d + 1] - w0;
    const Pose4 q = vo[f];
    x = (double)q.x; y = (double)q.y; th = (double)q.z; v = (double)q.w;
    if (m < nw && gps != nullptr) {
      const Pose4 g = gps[f];
      double dd = pymod_pos(dsub(th, (double)g.z), kTwoPi);
      if (dd > kPi) dd = dsub(dd, kTwoPi);
      th = pymod_pos(dsub(th, ddiv(dd, 2.0)), kTwoPi);
      v = ddiv(dadd(v, (double)g.w), 2.0);
    }
    long long i = m < nw - 1 ? m : nw - 1;
    const long long i_min = m - max_steps + 1;
    for (; i >= 0 && i >= i_min; --i) {
      if ((long long)results[w0 + i].n_steps > m - i) { cover = w0 + i; depth = (int)(m - i); break; }
    }
    if (cover >= 0 && depth == 0) {
      x = results[cover].x1;
      y = results[cover].y1;
      cover = -1;
    }
  }
  // deeper offsets: serve every lane that waits for the same window with one re-rollout
  unsigned need = __ballot_sync(FULL, cover >= 0);
  while (need) {
    const int leader = __ffs(need) - 1;
    const long long cw = __shfl_sync(FULL, cover, leader);
    const int cd = __shfl_sync(FULL, d, leader);
    const vmvo_window_result r = results[cw];
    const int hi = r.best_idx / gs, hj = r.best_idx - hi * gs;
    GridCtl g{r.v_seed, r.s_seed, dt_drive[cd], grid_rate(max_accel, hi, gv),
              grid_rate(max_rate, hj, gs), max_steer};
    const bool nonfinite = (r.status & VMVO_WIN_NONFINITE) != 0;
    Pose<double> carry{0.0, 0.0, 0.0};
    const bool mine = cover == cw;
    for (int base = 0; base < r.n_steps; base += 32) {
      const int k = base + lane + 1;
      const bool active = k <= r.n_steps && !nonfinite;
      double cv = 0.0, cs = 0.0;
      if (active) g.at(k, &cv, &cs);
      Pose<double> pz = warp_model_round<double>(cv, cs, active, g.dt, L, ratio, carry, lane);
      // lane l holds step base+l+1, i.e. offset base+l
      const int src = depth - base;
      const double px = __shfl_sync(FULL, pz.x, src & 31);
      const double py = __shfl_sync(FULL, pz.y, src & 31);
      if (mine && src >= 0 && src < 32) {
        x = nonfinite ? CUDART_NAN : px;
        y = nonfinite ? CUDART_NAN : py;
      }
    }
    if (mine) cover = -1;
    need = __ballot_sync(FULL, cover >= 0);
  }
  if (in_range) {
    out_x[f] = x; out_y[f] = y; out_th[f] = th; out_v[f] = v;
  }
}

// ---- a1 / a2: batched rollout, one warp per sequence, one step per lane ---------------------------
template <typename T>
__global__ void rollout_kernel(long long n_seq, int n_steps, const T* steer, const T* vel, T dt,
                               const T* state0, T L, T ratio, T max_steer, T max_accel, T* out,
                               int* fail) {
  const int lane = threadIdx.x & 31;
  const long long warps_per_grid = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_seq;
       q += warps_per_grid) {
    Pose<T> carry{state0[q * 4 + 0], state0[q * 4 + 1], state0[q * 4 + 2]};
    T v_prev_round = state0[q * 4 + 3];
    int fail_step = 0x7fffffff, fail_kind = 0;
    for (int base = 0; base < n_steps; base += 32) {
      const int k = base + lane;  // 0-based step
      const bool active = k < n_steps;
      T v = (T)0, s = (T)0;
      if (active) {
        v = vel[q * n_steps + k];
        s = steer[q * n_steps + k];
      }
      // the asserts of vmvo/bicycle_model.py:48-62, in the reference's order (steer first)
      T v_prev = __shfl_up_sync(FULL, v, 1);
      if (lane == 0) v_prev = v_prev_round;
      int kind = 0;
      if (active) {
        if (!(Num<T>::abs_(s) <= max_steer)) kind = VMVO_FAIL_STEER;
        else if (!(Num<T>::abs_(Num<T>::div(Num<T>::sub(v, v_prev), dt)) <= max_accel)) kind = VMVO_FAIL_ACCEL;
      }
      const unsigned bad = __ballot_sync(FULL, kind != 0);
      if (bad && fail_kind == 0) {
        const int l = __ffs(bad) - 1;
        fail_step = base + l;
        fail_kind = __shfl_sync(FULL, kind, l);
      }
      Pose<T> pz = warp_model_round<T>(v, s, active, dt, L, ratio, carry, lane);
      if (active) {
        T* o = out + (q * n_steps + k) * 3;
        o[0] = pz.x; o[1] = pz.y; o[2] = pz.th;
      }
      v_prev_round = __shfl_sync(FULL, v, 31);
    }
    if (lane == 0) {
      fail[q * 2 + 0] = fail_kind;
      fail[q * 2 + 1] = fail_kind ? fail_step : -1;
    }
  }
}

// ---- a10: cost closure of mpc_run for given steering sequences -------------------------------
__global__ void sequence_cost_kernel(long long n_seq, int n_steps, const double* steer,
                                     double velocity, double dt, const double* target_xy, double L,
                                     double ratio, double k_steer, double* cost) {
  const int lane = threadIdx.x & 31;
  const long long warps_per_grid = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_seq;
       q += warps_per_grid) {
    // start state (target[0], theta 0): vmvo/utils/mpc.py:83-85
    Pose<double> carry{target_xy[0], target_xy[1], 0.0};
    double J = 0.0;
    for (int base = 0; base < n_steps; base += 32) {
      const int k = base + lane;
      const bool active = k < n_steps;
      const double s = active ? steer[q * n_steps + k] : 0.0;
      Pose<double> pz = warp_model_round<double>(active ? velocity : 0.0, s, active, dt, L, ratio,
                                                 carry, lane);
      double term = 0.0;
      if (active) {
        // state after step k+1 against target k (mpc.py:70-78)
        const double ex = dsub(pz.x, target_xy[2 * k]), ey = dsub(pz.y, target_xy[2 * k + 1]);
        term = dadd(dadd(dmul(ex, ex), dmul(ey, ey)), dmul(k_steer, dmul(s, s)));
      }
      J = dadd(J, warp_sum(term));
    }
    if (lane == 0) cost[q] = J;
  }
}

// ---- a7 / a8 / a9 stand-alone ------------------------------------------------------------------
__global__ void extract_window_kernel(int n, const double* x, const double* y, const double* th,
                                      double* lx, double* ly, double* lth) {
  const double th0 = th[0], x0 = x[0], y0 = y[0];
  double sn, cs;
  sincos(th0, &sn, &cs);
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const double dx = dsub(x[m], x0), dy = dsub(y[m], y0);
    lx[m] = dadd(dmul(dx, cs), dmul(dy, sn));
    ly[m] = dadd(dmul(-dx, sn), dmul(dy, cs));
    lth[m] = dsub(th[m], th0);
  }
}

__global__ void time_extent_kernel(long long n, const double* time, double t0, double t1,
                                   long long* extent) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    extent[0] = lower_bound_f64(time, n, t0);
    extent[1] = upper_bound_f64(time, n, t1);
  }
}

__global__ void traverse_kernel(int n, const double* xy, double D, int* keep, int* count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int cnt = 0;
  if (n > 0) keep[cnt++] = 0;
  double dist = 0.0;
  for (int i = 1; i < n; ++i) {
    const double ddx = dsub(xy[2 * i], xy[2 * i - 2]), ddy = dsub(xy[2 * i + 1], xy[2 * i - 1]);
    const double seg = sqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy)));
    if (dadd(dist, seg) > D) {
      keep[cnt++] = i - 1;
      dist = seg;
    } else {
      dist = dadd(dist, seg);
    }
  }
  *count = cnt;
}

// ---- the tangent the search tabulates (exposed so that tests can check the compiled code) -----
__global__ void tan_steer_kernel(long long n, const float* x, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = tan_steer(x[i]);
}

// ---- issue-rate probes -------------------------------------------------------------------------
__global__ void peak_probe_kernel(int kind, int iters, float* sink) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (kind == 0) {  // MUFU: 8 independent sin/cos pairs per iteration = 16 MUFU ops
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.001f * (float)(tid & 1023) + 0.37f * q;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float s = __sinf(a[q]);
        float c = __cosf(a[q]);
        a[q] = s + c * 0.5f;   // 1 FFMA keeps the chain alive; MUFU is the bound
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += a[q];
  } else if (kind == 1) {  // FFMA: 16 independent chains
    float a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = 1.0f + 1e-3f * (float)((tid + q) & 255);
    const float m = 0.9999f, b = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = fmaf(a[q], m, b);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) acc += a[q];
  } else {  // DFMA: 8 independent chains
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 1.0 + 1e-3 * (double)((tid + q) & 255);
    const double m = 0.9999, b = 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = fma(a[q], m, b);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += (float)a[q];
  }
  sink[tid] = acc;
}

}  // namespace vmvo

using namespace vmvo;

// ---- context ---------------------------------------------------------------------------------
extern "C" int vmvo_abi_version(void) { return VMVO_ABI_VERSION; }

extern "C" int vmvo_ctx_create(int device, vmvo_ctx** out) {
  if (!out) return VMVO_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return VMVO_ERR_CUDA;
  if (device < 0 || device >= count) return VMVO_ERR_BAD_ARG;
  vmvo_ctx* ctx = new vmvo_ctx();
  ctx->device = device;
  ctx->launches = 0;
  ctx->err[0] = 0;
  ctx->d_counters = nullptr;
  ctx->slot_mutex = new std::mutex();
  ctx->h_stage = nullptr;
  ctx->h_stage_bytes = 0;
  ctx->host_stream = nullptr;
  ctx->tune = vmvo_tuning{-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  for (int q = 0; q < kLaunchSlots; ++q) ctx->slots[q] = vmvo_launch_slot{nullptr, nullptr, 0, nullptr, false, false};
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  bool ok = guard.err == cudaSuccess && cudaGetDeviceProperties(&prop, device) == cudaSuccess &&
            cudaMalloc(&ctx->d_counters, kLaunchSlots * 4 * sizeof(unsigned long long)) == cudaSuccess;
  for (int q = 0; ok && q < kLaunchSlots; ++q) {
    ctx->slots[q].d_counters = ctx->d_counters + 4 * q;
    ok = cudaEventCreateWithFlags(&ctx->slots[q].done, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    vmvo_ctx_destroy(ctx);
    return VMVO_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    vmvo_ctx_destroy(ctx);
    return VMVO_ERR_UNSUPPORTED;  // built for sm_100a only
  }
  *out = ctx;
  return VMVO_OK;
}

extern "C" int vmvo_ctx_destroy(vmvo_ctx* ctx) {
  if (!ctx) return VMVO_OK;
  {
    DeviceGuard guard(ctx->device);
    cudaFree(ctx->d_counters);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
    for (int q = 0; q < kLaunchSlots; ++q) {
      if (ctx->slots[q].d_defer) cudaFree(ctx->slots[q].d_defer);
      if (ctx->slots[q].done) cudaEventDestroy(ctx->slots[q].done);
    }
  }
  delete ctx->slot_mutex;
  delete ctx;
  return VMVO_OK;
}

extern "C" int vmvo_debug_set_tuning(vmvo_ctx* ctx, const char* key, int32_t value) {
  if (!ctx || !key) return VMVO_ERR_BAD_ARG;
  int* slot = !strcmp(key, "team_warps") ? &ctx->tune.team_warps
            : !strcmp(key, "fast_scan") ? &ctx->tune.fast_scan
            : !strcmp(key, "cand_cap") ? &ctx->tune.cand_cap
            : !strcmp(key, "defer_min") ? &ctx->tune.defer_min
            : !strcmp(key, "max_ctas_per_sm") ? &ctx->tune.max_ctas_per_sm
            : !strcmp(key, "defer_warps") ? &ctx->tune.defer_warps
            : !strcmp(key, "cta_teams") ? &ctx->tune.cta_teams
            : !strcmp(key, "pdl") ? &ctx->tune.pdl
            : !strcmp(key, "prep") ? &ctx->tune.prep
            : !strcmp(key, "prune") ? &ctx->tune.prune
            : !strcmp(key, "prune_every") ? &ctx->tune.prune_every
            : !strcmp(key, "lean") ? &ctx->tune.lean : nullptr;
  if (!slot) return fail(ctx, VMVO_ERR_BAD_ARG, "unknown tuning key '%s'", key);
  *slot = value < 0 ? -1 : value;
  return VMVO_OK;
}

// ---- peer buffers and arrival words (the exchange fused into the search / write-back) ----
extern "C" int vmvo_peer_buffer_create(vmvo_ctx* ctx, int64_t bytes, void** d_ptr, uint8_t* h_handle) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (bytes < 1 || !d_ptr || !h_handle) return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  VMVO_ON_DEVICE(ctx);
  void* p = nullptr;
  VMVO_CUDA(ctx, cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(ctx, VMVO_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(h_handle, &h, sizeof(h));
  *d_ptr = p;
  return VMVO_OK;
}

extern "C" int vmvo_peer_buffer_destroy(vmvo_ctx* ctx, void* d_ptr) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (!d_ptr) return VMVO_OK;
  VMVO_ON_DEVICE(ctx);
  VMVO_CUDA(ctx, cudaFree(d_ptr));
  return VMVO_OK;
}

extern "C" int vmvo_peer_buffer_open(vmvo_ctx* ctx, const uint8_t* h_handle, void** d_ptr) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (!h_handle || !d_ptr) return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  VMVO_ON_DEVICE(ctx);
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle, sizeof(h));
  VMVO_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VMVO_OK;
}

extern "C" int vmvo_peer_buffer_close(vmvo_ctx* ctx, void* d_ptr) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (!d_ptr) return VMVO_OK;
  VMVO_ON_DEVICE(ctx);
  VMVO_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
  return VMVO_OK;
}

extern "C" const char* vmvo_last_error(const vmvo_ctx* ctx) { return ctx ? ctx->err : "ctx is NULL"; }

extern "C" int64_t vmvo_launch_count(const vmvo_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" void vmvo_search_cfg_default(vmvo_search_cfg* c) {
  if (!c) return;
  c->grid_v = 32;
  c->grid_s = 32;
  c->window_mode = VMVO_WINDOW_FRAMES;
  c->window_frames = 30;
  c->horizon_frames = 60;
  c->target_mode = VMVO_TARGET_TIME;
  c->target_offset = 1;
  c->seed_mode = VMVO_SEED_DATA;
  c->primary = VMVO_PRIMARY_VO;
  c->max_window_poses = 128;
  c->horizon_time = 3.0;                      // optimize_trajectory_v2.py:35
  c->w_vo = 1.0; c->w_gps = 0.0; c->w_imu = 0.0;
  c->k_steer = 0.0;                           // mpc.py:31
  c->wheel_base = 2.83972;                    // constants.py:3-7
  c->steering_ratio = 13.27;
  c->max_steer = 460.0;
  c->max_accel = 10.0;
  c->max_steer_rate = 100.0;
}

extern "C" int64_t vmvo_window_count(const vmvo_search_cfg* cfg, int64_t n_frames) {
  if (!cfg) return 0;
  const int64_t h = cfg->window_mode == VMVO_WINDOW_FRAMES ? cfg->window_frames : cfg->horizon_frames;
  const int64_t n = n_frames - 2 * h;
  return n > 0 ? n : 0;
}

static inline unsigned grid_for(long long n, int threads, int cap) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

extern "C" int vmvo_plan_windows(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                                 const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                                 int64_t n_windows, const double* d_time, int64_t* d_win_start,
                                 int32_t* d_win_len, int32_t* d_win_drive, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  int rc = validate_cfg(ctx, cfg);
  if (rc) return rc;
  if (n_drives < 1 || n_windows < 0) return fail(ctx, VMVO_ERR_BAD_ARG, "n_drives < 1 or n_windows < 0");
  if (n_windows == 0) return VMVO_OK;
  if (!d_drive_offsets || !d_window_offsets || !d_win_start || !d_win_len || !d_win_drive)
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  if (cfg->window_mode == VMVO_WINDOW_TIME && !d_time)
    return fail(ctx, VMVO_ERR_BAD_ARG, "time mode needs d_time");
  VMVO_ON_DEVICE(ctx);
  plan_windows_kernel<<<grid_for(n_windows, 256, ctx->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(
      cfg->window_mode, cfg->window_frames, cfg->horizon_time, n_drives,
      (const long long*)d_drive_offsets, (const long long*)d_window_offsets, n_windows, d_time,
      (long long*)d_win_start, d_win_len, d_win_drive);
  return check_launch(ctx, "plan_windows_kernel");
}

static int exchange_dev(vmvo_ctx* ctx, const vmvo_exchange* ex, ExchangeDev* out) {
  memset(out, 0, sizeof(*out));
  if (!ex || ex->n_peers <= 0) return VMVO_OK;
  if (ex->world < 2 || ex->rank < 0 || ex->rank >= ex->world || ex->n_peers != ex->world - 1 ||
      ex->n_peers > VMVO_MAX_MIRRORS || !ex->local_flags || !ex->epoch)
    return fail(ctx, VMVO_ERR_BAD_ARG, "exchange: world %d rank %d n_peers %d, or NULL flags / epoch",
                ex->world, ex->rank, ex->n_peers);
  out->world = ex->world;
  out->rank = ex->rank;
  out->n_peers = ex->n_peers;
  for (int q = 0; q < ex->n_peers; ++q) {
    if (!ex->peer_flags[q]) return fail(ctx, VMVO_ERR_BAD_ARG, "exchange: peer flag %d is NULL", q);
    out->peer_flags[q] = ex->peer_flags[q];
  }
  out->local_flags = ex->local_flags;
  out->epoch = ex->epoch;
  return VMVO_OK;
}

extern "C" int vmvo_exchange_publish(vmvo_ctx* ctx, const vmvo_exchange* ex, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  ExchangeDev d;
  int rc = exchange_dev(ctx, ex, &d);
  if (rc || d.n_peers == 0) return rc;
  VMVO_ON_DEVICE(ctx);
  exchange_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d);
  return check_launch(ctx, "exchange_publish_kernel");
}

extern "C" int vmvo_exchange_wait(vmvo_ctx* ctx, const vmvo_exchange* ex, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  ExchangeDev d;
  int rc = exchange_dev(ctx, ex, &d);
  if (rc || d.n_peers == 0) return rc;
  VMVO_ON_DEVICE(ctx);
  exchange_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d);
  return check_launch(ctx, "exchange_wait_kernel");
}

template <typename Pose4>
static int write_back_impl(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                           int64_t total_frames, int64_t frame_lo, int64_t frame_hi,
                           const int64_t* d_drive_offsets,
                           const int64_t* d_window_offsets, const double* d_dt_per_drive,
                           const void* d_vo, const void* d_gps, const vmvo_window_result* d_results,
                           double* d_out_x, double* d_out_y, double* d_out_theta, double* d_out_vel,
                           const vmvo_exchange* ex, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  int rc = validate_cfg(ctx, cfg);
  if (rc) return rc;
  ExchangeDev exd;
  rc = exchange_dev(ctx, ex, &exd);
  if (rc) return rc;
  if (n_drives < 1) return fail(ctx, VMVO_ERR_BAD_ARG, "n_drives < 1");
  if (frame_lo < 0 || frame_hi > total_frames || frame_lo > frame_hi)
    return fail(ctx, VMVO_ERR_BAD_ARG, "frame range [%lld, %lld) outside [0, %lld)", (long long)frame_lo,
                (long long)frame_hi, (long long)total_frames);
  if (!d_drive_offsets || !d_window_offsets || !d_dt_per_drive || !d_vo || !d_results || !d_out_x ||
      !d_out_y || !d_out_theta || !d_out_vel)
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  if (((uintptr_t)d_vo | (uintptr_t)d_gps) & 15)
    return fail(ctx, VMVO_ERR_BAD_ARG, "pose streams must be 16-byte aligned");
  VMVO_ON_DEVICE(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = frame_hi - frame_lo;
  if (total <= 0 && exd.n_peers == 0) return VMVO_OK;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks < 1) blocks = 1;          // (an empty share still publishes and waits)
  write_back_kernel<Pose4><<<(unsigned)blocks, threads, 0, st>>>(
      cfg->grid_v, cfg->grid_s, cfg->wheel_base, cfg->steering_ratio, cfg->max_steer, cfg->max_accel,
      cfg->max_steer_rate, cfg->max_window_poses, n_drives, (const long long*)d_drive_offsets,
      (const long long*)d_window_offsets, d_dt_per_drive, (const Pose4*)d_vo, (const Pose4*)d_gps,
      d_results, frame_lo, frame_hi, d_out_x, d_out_y, d_out_theta, d_out_vel, exd);
  return check_launch(ctx, "write_back_kernel");
}

extern "C" int vmvo_write_back_f32(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                                   int64_t total_frames, const int64_t* d_drive_offsets,
                                   const int64_t* d_window_offsets, const double* d_dt_per_drive,
                                   const float* d_vo, const float* d_gps,
                                   const vmvo_window_result* d_results, double* d_out_x,
                                   double* d_out_y, double* d_out_theta, double* d_out_vel,
                                   void* stream) {
  return write_back_impl<float4>(ctx, cfg, n_drives, total_frames, 0, total_frames, d_drive_offsets,
                                 d_window_offsets, d_dt_per_drive, d_vo, d_gps, d_results, d_out_x,
                                 d_out_y, d_out_theta, d_out_vel, nullptr, stream);
}

extern "C" int vmvo_write_back_f64(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                                   int64_t total_frames, const int64_t* d_drive_offsets,
                                   const int64_t* d_window_offsets, const double* d_dt_per_drive,
                                   const double* d_vo, const double* d_gps,
                                   const vmvo_window_result* d_results, double* d_out_x,
                                   double* d_out_y, double* d_out_theta, double* d_out_vel,
                                   void* stream) {
  return write_back_impl<double4>(ctx, cfg, n_drives, total_frames, 0, total_frames, d_drive_offsets,
                                  d_window_offsets, d_dt_per_drive, d_vo, d_gps, d_results, d_out_x,
                                  d_out_y, d_out_theta, d_out_vel, nullptr, stream);
}

extern "C" int vmvo_write_back_range(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                                     int64_t total_frames, int64_t frame_lo, int64_t frame_hi,
                                     const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                                     const double* d_dt_per_drive, const void* d_vo, const void* d_gps,
                                     int32_t stream_f64, const vmvo_window_result* d_results,
                                     double* d_out_x, double* d_out_y, double* d_out_theta,
                                     double* d_out_vel, const vmvo_exchange* ex, void* stream) {
  return stream_f64
             ? write_back_impl<double4>(ctx, cfg, n_drives, total_frames, frame_lo, frame_hi,
                                        d_drive_offsets, d_window_offsets, d_dt_per_drive, d_vo, d_gps,
                                        d_results, d_out_x, d_out_y, d_out_theta, d_out_vel, ex, stream)
             : write_back_impl<float4>(ctx, cfg, n_drives, total_frames, frame_lo, frame_hi,
                                       d_drive_offsets, d_window_offsets, d_dt_per_drive, d_vo, d_gps,
                                       d_results, d_out_x, d_out_y, d_out_theta, d_out_vel, ex, stream);
}

template <typename T>
static int rollout_impl(vmvo_ctx* ctx, int64_t n_seq, int32_t n_steps, const T* d_steer,
                        const T* d_vel, T dt, const T* d_state0, T max_steer, T max_accel, T* d_out,
                        int32_t* d_fail, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_seq < 0 || n_steps < 0) return fail(ctx, VMVO_ERR_BAD_ARG, "negative size");
  if (n_seq == 0) return VMVO_OK;
  if (!d_state0 || !d_fail || (n_steps > 0 && (!d_steer || !d_vel || !d_out)))
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  VMVO_ON_DEVICE(ctx);
  const int threads = 128;
  const unsigned blocks = grid_for(n_seq * 32, threads, ctx->sm_count * 16);
  // the model uses the module constants, not the constructor arguments (quirk D2,
  // vmvo/bicycle_model.py:66-68)
  rollout_kernel<T><<<blocks, threads, 0, (cudaStream_t)stream>>>(
      n_seq, n_steps, d_steer, d_vel, dt, d_state0, (T)2.83972, (T)13.27, max_steer, max_accel,
      d_out, d_fail);
  return check_launch(ctx, "rollout_kernel");
}

extern "C" int vmvo_rollout_f64(vmvo_ctx* ctx, int64_t n_seq, int32_t n_steps, const double* d_steer,
                                const double* d_vel, double dt, const double* d_state0,
                                double max_steer, double max_accel, double* d_out, int32_t* d_fail,
                                void* stream) {
  return rollout_impl<double>(ctx, n_seq, n_steps, d_steer, d_vel, dt, d_state0, max_steer,
                              max_accel, d_out, d_fail, stream);
}

extern "C" int vmvo_rollout_f32(vmvo_ctx* ctx, int64_t n_seq, int32_t n_steps, const float* d_steer,
                                const float* d_vel, float dt, const float* d_state0, float max_steer,
                                float max_accel, float* d_out, int32_t* d_fail, void* stream) {
  return rollout_impl<float>(ctx, n_seq, n_steps, d_steer, d_vel, dt, d_state0, max_steer, max_accel,
                             d_out, d_fail, stream);
}

// ---- a1 / a2 for a scalar caller: host buffers in, host buffers out, synchronous ---------------------
// BicycleModel.run / run_sequence of the reference are called one step or one short sequence at a
// time from Python (vmvo/bicycle_model.py:40-92).  Through tensors that is an allocation, three
// copies and a launch per call; here the controls are written into a pinned buffer the GPU reads in
// place (mapped host memory), the same rollout_kernel runs on a stream of the ctx's own, and the poses
// come back through the same buffer: one launch and one stream synchronisation per call.
extern "C" int vmvo_rollout_host_f64(vmvo_ctx* ctx, int32_t n_steps, const double* h_steer,
                                     const double* h_vel, double dt, const double* h_state0,
                                     double max_steer, double max_accel, double* h_out,
                                     int32_t* h_fail) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_steps < 0 || !h_state0 || !h_fail || (n_steps > 0 && (!h_steer || !h_vel || !h_out)))
    return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  if (n_steps == 0) {
    h_fail[0] = VMVO_FAIL_NONE;
    h_fail[1] = -1;
    return VMVO_OK;
  }
  VMVO_ON_DEVICE(ctx);
  std::lock_guard<std::mutex> lock(*ctx->slot_mutex);
  const size_t n = (size_t)n_steps;
  const size_t need = (n * 5 + 4) * sizeof(double) + 16;
  if (ctx->h_stage_bytes < need) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    const size_t bytes = need < 65536 ? 65536 : need * 2;
    VMVO_CUDA(ctx, cudaHostAlloc(&ctx->h_stage, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    ctx->h_stage_bytes = bytes;
  }
  if (!ctx->host_stream) VMVO_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking));
  double* hs = reinterpret_cast<double*>(ctx->h_stage);      // [steer n][vel n][state0 4][out 3n][fail 2 x int]
  memcpy(hs, h_steer, n * sizeof(double));
  memcpy(hs + n, h_vel, n * sizeof(double));
  memcpy(hs + 2 * n, h_state0, 4 * sizeof(double));
  void* dv = nullptr;
  VMVO_CUDA(ctx, cudaHostGetDevicePointer(&dv, ctx->h_stage, 0));
  double* ds = reinterpret_cast<double*>(dv);
  int* d_fail = reinterpret_cast<int*>(ds + 5 * n + 4);
  rollout_kernel<double><<<1, 32, 0, ctx->host_stream>>>(1, n_steps, ds, ds + n, dt, ds + 2 * n, 2.83972, 13.27,
                                                         max_steer, max_accel, ds + 2 * n + 4, d_fail);
  int rc = check_launch(ctx, "rollout_kernel");
  if (rc) return rc;
  VMVO_CUDA(ctx, cudaStreamSynchronize(ctx->host_stream));
  memcpy(h_out, hs + 2 * n + 4, 3 * n * sizeof(double));
  memcpy(h_fail, reinterpret_cast<int*>(hs + 5 * n + 4), 2 * sizeof(int));
  return VMVO_OK;
}

extern "C" int vmvo_sequence_cost_f64(vmvo_ctx* ctx, int64_t n_seq, int32_t n_steps,
                                      const double* d_steer, double velocity, double dt,
                                      const double* d_target_xy, double k_steer, double* d_cost,
                                      void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_seq < 0 || n_steps < 0) return fail(ctx, VMVO_ERR_BAD_ARG, "negative size");
  if (n_seq == 0) return VMVO_OK;
  if (!d_target_xy || !d_cost || (n_steps > 0 && !d_steer))
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  VMVO_ON_DEVICE(ctx);
  const int threads = 128;
  sequence_cost_kernel<<<grid_for(n_seq * 32, threads, ctx->sm_count * 16), threads, 0,
                         (cudaStream_t)stream>>>(n_seq, n_steps, d_steer, velocity, dt, d_target_xy,
                                                 2.83972, 13.27, k_steer, d_cost);
  return check_launch(ctx, "sequence_cost_kernel");
}

extern "C" int vmvo_extract_window_f64(vmvo_ctx* ctx, int32_t n, const double* d_x, const double* d_y,
                                       const double* d_theta, double* d_lx, double* d_ly,
                                       double* d_lth, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n < 1) return fail(ctx, VMVO_ERR_BAD_ARG, "n < 1");
  if (!d_x || !d_y || !d_theta || !d_lx || !d_ly || !d_lth) return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  VMVO_ON_DEVICE(ctx);
  extract_window_kernel<<<grid_for(n, 256, ctx->sm_count * 4), 256, 0, (cudaStream_t)stream>>>(
      n, d_x, d_y, d_theta, d_lx, d_ly, d_lth);
  return check_launch(ctx, "extract_window_kernel");
}

extern "C" int vmvo_time_extent_f64(vmvo_ctx* ctx, int64_t n, const double* d_time, double t0,
                                    double t1, int64_t* d_extent, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n < 0 || !d_extent || (n > 0 && !d_time)) return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  VMVO_ON_DEVICE(ctx);
  time_extent_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(n, d_time, t0, t1, (long long*)d_extent);
  return check_launch(ctx, "time_extent_kernel");
}

extern "C" int vmvo_traverse_f64(vmvo_ctx* ctx, int32_t n, const double* d_xy, double D,
                                 int32_t* d_keep, int32_t* d_count, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n < 0 || !d_count || (n > 0 && (!d_xy || !d_keep))) return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  VMVO_ON_DEVICE(ctx);
  traverse_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(n, d_xy, D, d_keep, d_count);
  return check_launch(ctx, "traverse_kernel");
}

extern "C" int vmvo_tan_steer_f32(vmvo_ctx* ctx, int64_t n, const float* d_delta, float* d_out,
                                  void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n < 0 || (n > 0 && (!d_delta || !d_out))) return fail(ctx, VMVO_ERR_BAD_ARG, "bad argument");
  if (n == 0) return VMVO_OK;
  VMVO_ON_DEVICE(ctx);
  tan_steer_kernel<<<grid_for(n, 256, ctx->sm_count * 16), 256, 0, (cudaStream_t)stream>>>(n, d_delta, d_out);
  return check_launch(ctx, "tan_steer_kernel");
}

extern "C" int vmvo_peak_probe(vmvo_ctx* ctx, int32_t kind, int32_t blocks, int32_t threads,
                               int32_t iters, float* d_sink, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (kind < 0 || kind > 2 || blocks < 1 || threads < 32 || threads > 1024 || iters < 1 || !d_sink)
    return fail(ctx, VMVO_ERR_BAD_ARG, "bad probe argument");
  VMVO_ON_DEVICE(ctx);
  peak_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(kind, iters, d_sink);
  return check_launch(ctx, "peak_probe_kernel");
}
