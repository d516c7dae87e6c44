// Host-side internals of libvmvo_b200.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/vmvo_b200.h"

// Per-launch scratch of a search: the work-queue head and the deferred-window slot count (16 bytes,
// cleared by one memset) and the buffer the deferred windows are parked in.  A slot belongs to ONE
// launch at a time: it is handed out again only after the event recorded behind that launch has
// completed, so searches of one ctx on different streams never share scratch.  A launch recorded
// into a CUDA graph keeps its slot for good (pinned): the graph may be replayed at any time.
struct vmvo_launch_slot {
  unsigned long long* d_counters;   // [4]: queue head, deferred-window count, finished windows, -
  unsigned char* d_defer;
  size_t defer_bytes;
  cudaEvent_t done;
  bool used, pinned;
};
constexpr int kLaunchSlots = 256;

// test / tuning overrides (vmvo_debug_set_tuning); -1 = the library's own choice
struct vmvo_tuning {
  int team_warps, fast_scan, cand_cap, defer_min, max_ctas_per_sm, defer_warps, cta_teams, pdl, prep;
  int prune, prune_every, lean;
};

struct vmvo_ctx {
  int device;
  int sm_count;
  unsigned long long* d_counters;   // kLaunchSlots x 4
  vmvo_launch_slot slots[kLaunchSlots];
  std::mutex* slot_mutex;
  long long launches;
  vmvo_tuning tune;
  // vmvo_rollout_host_f64: a pinned, device-mapped staging buffer and a stream of the ctx's own
  void* h_stage;
  size_t h_stage_bytes;
  cudaStream_t host_stream;
  char err[512];
};

namespace vmvo {

inline int fail(vmvo_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

inline int check_launch(vmvo_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(ctx, VMVO_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  ctx->launches++;
  return VMVO_OK;
}

#define VMVO_CUDA(ctx, call)                                                        \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return vmvo::fail(ctx, VMVO_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// Makes ctx->device current for the duration of an entry point and restores the caller's device on
// the way out (the library never leaves the calling thread's current device changed).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    else if (err == cudaSuccess) prev = -1;      // already current: nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define VMVO_ON_DEVICE(ctx)                                                                   \
  vmvo::DeviceGuard dev_guard__((ctx)->device);                                               \
  if (dev_guard__.err != cudaSuccess)                                                         \
    return vmvo::fail(ctx, VMVO_ERR_CUDA, "cudaSetDevice(%d): %s", (ctx)->device,            \
                      cudaGetErrorString(dev_guard__.err))

inline int validate_cfg(vmvo_ctx* ctx, const vmvo_search_cfg* c) {
  if (!c) return fail(ctx, VMVO_ERR_BAD_ARG, "cfg is NULL");
  if (c->grid_v < 1 || c->grid_s < 1 || (int64_t)c->grid_v * c->grid_s > (1 << 30))
    return fail(ctx, VMVO_ERR_BAD_ARG, "grid %d x %d out of range", c->grid_v, c->grid_s);
  if (c->window_mode != VMVO_WINDOW_FRAMES && c->window_mode != VMVO_WINDOW_TIME)
    return fail(ctx, VMVO_ERR_BAD_ARG, "window_mode %d", c->window_mode);
  if (c->window_mode == VMVO_WINDOW_FRAMES && c->window_frames < 1)
    return fail(ctx, VMVO_ERR_BAD_ARG, "window_frames %d", c->window_frames);
  if (c->window_mode == VMVO_WINDOW_TIME && (c->horizon_frames < 1 || !(c->horizon_time > 0)))
    return fail(ctx, VMVO_ERR_BAD_ARG, "horizon_frames %d / horizon_time %g", c->horizon_frames,
                c->horizon_time);
  if (c->target_mode != VMVO_TARGET_TIME && c->target_mode != VMVO_TARGET_TRAVERSE)
    return fail(ctx, VMVO_ERR_BAD_ARG, "target_mode %d", c->target_mode);
  if (c->target_offset != 0 && c->target_offset != 1)
    return fail(ctx, VMVO_ERR_BAD_ARG, "target_offset %d", c->target_offset);
  if (c->seed_mode < VMVO_SEED_DATA || c->seed_mode > VMVO_SEED_CHAINED)
    return fail(ctx, VMVO_ERR_BAD_ARG, "seed_mode %d", c->seed_mode);
  if (c->primary != VMVO_PRIMARY_VO && c->primary != VMVO_PRIMARY_GPS)
    return fail(ctx, VMVO_ERR_BAD_ARG, "primary %d", c->primary);
  if (c->max_window_poses < 2 || c->max_window_poses > 256)
    return fail(ctx, VMVO_ERR_BAD_ARG, "max_window_poses %d not in [2, 256]", c->max_window_poses);
  if (c->window_mode == VMVO_WINDOW_FRAMES && c->window_frames + 1 > c->max_window_poses)
    return fail(ctx, VMVO_ERR_BAD_ARG, "window_frames + 1 = %d exceeds max_window_poses %d",
                c->window_frames + 1, c->max_window_poses);
  if (!(c->w_vo >= 0) || !(c->w_gps >= 0) || !(c->w_imu >= 0) || !(c->k_steer >= 0))
    return fail(ctx, VMVO_ERR_BAD_ARG, "weights must be finite and >= 0");
  if (c->w_vo == 0 && c->w_gps == 0 && c->w_imu == 0 && c->k_steer == 0)
    return fail(ctx, VMVO_ERR_BAD_ARG, "all cost weights are zero");
  if (!(c->wheel_base > 0) || !(c->steering_ratio > 0) || !(c->max_steer > 0) ||
      !(c->max_accel >= 0) || !(c->max_steer_rate >= 0))
    return fail(ctx, VMVO_ERR_BAD_ARG, "vehicle constants out of range");
  if (c->max_steer * 3.14159265358979323846 / 180.0 / c->steering_ratio >= 1.5)
    return fail(ctx, VMVO_ERR_BAD_ARG, "max road-wheel angle must stay below 1.5 rad");
  return VMVO_OK;
}

}  // namespace vmvo
