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
//
// This file: the host side (launch parameters, per-launch scratch, the C ABI) and the MODE 0 kernels.
#include "vmvo_search_kernels.cuh"

namespace vmvo {

int launch_search_mode0(vmvo_ctx* ctx, const SearchParams& p, cudaStream_t st, bool dual, bool imu, bool f64) {
  return launch_search_mode<0>(ctx, p, st, dual, imu, f64);
}

// ---- per-launch scratch (vmvo_launch_slot, vmvo_internal.h) -----------------------------------------
// A slot that no launch in flight and no captured graph owns, with a deferred-window buffer of at
// least `need` bytes when one can be had: a free slot that already has one, else a free slot whose
// buffer is (re)allocated; without memory the search runs without deferral, which is still correct.
// While the caller's stream is being captured the event queries and the allocation run with this
// thread's capture mode relaxed (they touch nothing the capture records).
struct RelaxedCapture {
  bool on;
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  explicit RelaxedCapture(bool capturing) : on(capturing) {
    if (on) cudaThreadExchangeStreamCaptureMode(&mode);
  }
  ~RelaxedCapture() {
    if (on) cudaThreadExchangeStreamCaptureMode(&mode);
  }
};

static vmvo_launch_slot* acquire_launch_slot(vmvo_ctx* ctx, size_t need, bool capturing) {
  std::lock_guard<std::mutex> lock(*ctx->slot_mutex);
  RelaxedCapture relaxed(capturing);
  vmvo_launch_slot* fit = nullptr;      // free, buffer large enough (the smallest such)
  vmvo_launch_slot* bare = nullptr;     // free, no buffer
  vmvo_launch_slot* small = nullptr;    // free, buffer too small (the largest such)
  for (int q = 0; q < kLaunchSlots; ++q) {
    vmvo_launch_slot* sl = &ctx->slots[q];
    if (sl->pinned) continue;
    if (sl->used) {
      if (cudaEventQuery(sl->done) != cudaSuccess) {
        cudaGetLastError();             // (cudaErrorNotReady is not sticky, but clear it anyway)
        continue;
      }
      sl->used = false;
    }
    if (sl->defer_bytes >= need && sl->d_defer) {
      if (!fit || sl->defer_bytes < fit->defer_bytes) fit = sl;
    } else if (!sl->d_defer) {
      if (!bare) bare = sl;
    } else if (!small || sl->defer_bytes > small->defer_bytes) {
      small = sl;
    }
  }
  if (need == 0) return bare ? bare : (small ? small : fit);
  if (fit) return fit;
  vmvo_launch_slot* sl = bare ? bare : small;
  if (!sl) return nullptr;
  if (sl->d_defer) {                    // free slot: nothing in flight reads its buffer
    cudaFree(sl->d_defer);
    sl->d_defer = nullptr;
    sl->defer_bytes = 0;
  }
  void* buf = nullptr;
  if (cudaMalloc(&buf, need) == cudaSuccess) {
    sl->d_defer = (unsigned char*)buf;
    sl->defer_bytes = need;
  } else {
    cudaGetLastError();                 // no room: search without deferral
  }
  return sl;
}

static void release_launch_slot(vmvo_ctx* ctx, vmvo_launch_slot* sl, cudaStream_t st, bool capturing) {
  std::lock_guard<std::mutex> lock(*ctx->slot_mutex);
  if (capturing) {
    sl->pinned = true;                  // the graph owns it from now on
  } else {
    sl->used = cudaEventRecord(sl->done, st) == cudaSuccess;
    if (!sl->used) {                    // cannot track it: wait it out rather than share it
      cudaGetLastError();
      cudaStreamSynchronize(st);
    }
  }
}

static int validate_exchange(vmvo_ctx* ctx, const vmvo_exchange* ex) {
  if (!ex) return VMVO_OK;
  if (ex->world < 1 || ex->rank < 0 || ex->rank >= ex->world)
    return fail(ctx, VMVO_ERR_BAD_ARG, "exchange: rank %d of world %d", ex->rank, ex->world);
  if (ex->block < 0 || (ex->block > 0 && log2_exact(ex->block) < 0))
    return fail(ctx, VMVO_ERR_BAD_ARG, "exchange: block %d is not a power of two", ex->block);
  if (ex->n_peers < 0 || ex->n_peers > VMVO_MAX_MIRRORS || (ex->n_peers > 0 && ex->n_peers != ex->world - 1))
    return fail(ctx, VMVO_ERR_BAD_ARG, "exchange: n_peers %d (world - 1 = %d, at most %d)", ex->n_peers,
                ex->world - 1, VMVO_MAX_MIRRORS);
  for (int q = 0; q < ex->n_peers; ++q)
    if (!ex->peer_records[q] || ((uintptr_t)ex->peer_records[q] & 15) || !ex->peer_flags[q])
      return fail(ctx, VMVO_ERR_BAD_ARG, "exchange: peer %d has a NULL or misaligned pointer", q);
  if (ex->n_peers > 0 && (!ex->local_flags || !ex->epoch))
    return fail(ctx, VMVO_ERR_BAD_ARG, "exchange: local_flags / epoch is NULL");
  return VMVO_OK;
}

static int grid_search_impl(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                            const int64_t* d_win_start, const int32_t* d_win_len,
                            const int32_t* d_win_drive, const double* d_dt_per_drive,
                            const void* d_vo, const void* d_gps, const void* d_imu, bool f64,
                            const double* d_seeds, vmvo_window_result* d_results,
                            double* d_out_poses, double* d_out_steer, double* d_out_vel,
                            int32_t out_stride, float* d_dbg_cost, float* d_dbg_err,
                            int64_t n_runs, const int64_t* d_run_offsets, const vmvo_exchange* ex,
                            int32_t n_drives, const int64_t* d_drive_offsets,
                            const int64_t* d_window_offsets, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  int rc = validate_cfg(ctx, cfg);
  if (rc) return rc;
  if (n_windows < 0) return fail(ctx, VMVO_ERR_BAD_ARG, "n_windows < 0");
  if (n_windows == 0) return VMVO_OK;
  rc = validate_exchange(ctx, ex);
  if (rc) return rc;
  if (!d_dt_per_drive || !d_results) return fail(ctx, VMVO_ERR_BAD_ARG, "NULL dt / result pointer");
  const bool planned = d_win_start || d_win_len || d_win_drive;
  if (planned && (!d_win_start || !d_win_len || !d_win_drive))
    return fail(ctx, VMVO_ERR_BAD_ARG, "the window plan is three arrays: pass all or none");
  if (!planned) {
    if (cfg->window_mode != VMVO_WINDOW_FRAMES)
      return fail(ctx, VMVO_ERR_UNSUPPORTED, "window extents are computed in the kernel in frames mode only: "
                  "time mode needs the plan of vmvo_plan_windows");
    if (n_drives < 1 || !d_drive_offsets || !d_window_offsets)
      return fail(ctx, VMVO_ERR_BAD_ARG, "without a plan the search needs the drive and window offsets");
  }
  if (cfg->seed_mode == VMVO_SEED_CHAINED && !d_run_offsets)
    return fail(ctx, VMVO_ERR_UNSUPPORTED,
                "seed_mode chained walks the windows of a drive in order: call "
                "vmvo_grid_search_chained with the per-drive window ranges");
  if (d_run_offsets && (cfg->seed_mode != VMVO_SEED_CHAINED || n_runs < 1))
    return fail(ctx, VMVO_ERR_BAD_ARG, "run offsets are for seed_mode chained with n_runs >= 1");
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
  if ((uintptr_t)d_imu & (f64 ? 7 : 3)) return fail(ctx, VMVO_ERR_BAD_ARG, "imu stream misaligned");
  if ((d_out_poses || d_out_steer || d_out_vel) && out_stride < 1)
    return fail(ctx, VMVO_ERR_BAD_ARG, "out_stride < 1");
  if ((d_dbg_cost == nullptr) != (d_dbg_err == nullptr))
    return fail(ctx, VMVO_ERR_BAD_ARG, "debug outputs come in pairs");

  constexpr int kC = 8;
  const int cta_warps = 8;
  SearchParams p;
  p.gv = cfg->grid_v;
  p.gs = cfg->grid_s;
  p.n_ic = (p.gv + kC - 1) / kC;
  p.n_items = p.n_ic * p.gs;
  // team size: about two passes of 32 items per warp (measured best on 32x32: two warps) -- but at most
  // four warps where the tables of two such teams still leave room for two CTAs per SM: with pruning a
  // pass ends with its slowest warp (the one that holds the best steering rates), and the warps that
  // stopped early wait for it at the barrier; four-warp teams measured 5 % faster on 128x128 (the
  // loop below grows the team again when the tables do not fit: 256x256 keeps eight warps)
  int tw = 1;
  while (tw < cta_warps / 2 && p.n_items > tw * 32 * 2) tw *= 2;
  // window preparation as a pass of its own (vmvo_window_prep_kernel): for the small teams of small
  // grids, where the serial float64 work of phases A1-A3 is a tenth of a window's time
  const bool want_prep = !d_run_offsets && !d_dbg_cost && ctx->tune.prep != 0;
  const int prep_rec_bytes = PrepLayout(cfg->max_window_poses, use_vo && use_gps, use_imu).total;
  // a small grid's whole VD table (every acceleration) fits beside the rest: one fill and one
  // barrier per window instead of one per pass
  const bool vd_whole = (size_t)p.n_ic * kC * cfg->max_window_poses * 4 <= 8192;
  // ... unless the per-team tables would not leave room for two CTAs per SM
  for (;;) {
    const int th = tw * 32;
    int ch = chunks_per_pass(th, p.gs);
    if (ch > p.n_ic || vd_whole) ch = p.n_ic;
    const SmemLayout probe(cfg->max_window_poses, p.gs, ch * kC, tw, (int)load_vo + (int)load_gps,
                           use_vo && use_gps, use_imu, cfg->target_mode == VMVO_TARGET_TRAVERSE,
                           f64 ? 32 : 16, (want_prep && tw <= 2) ? prep_rec_bytes : 0);
    if (tw == cta_warps || probe.total * (cta_warps / tw) <= 110 * 1024) break;
    tw *= 2;
  }
  if (const int v = ctx->tune.team_warps; (v == 1 || v == 2 || v == 4 || v == 8) && v <= cta_warps) tw = v;
  p.team_warps = tw;
  p.cand_cap = kCandPerWarp * tw;
  // the packed / rotation scan: 10 instead of 16 MUFU per 8 hypothesis-steps, at the price of a
  // ~2.5x wider trig term in the band.  Measured: +5 % on 256x256, +2 % on 32x32 (re-scores per
  // window 2.5 -> 2.8); on smaller grids the per-item preamble outweighs the loop.
  p.allow_fast = p.n_items >= 128;
  if (ctx->tune.fast_scan >= 0) p.allow_fast = ctx->tune.fast_scan != 0;                 // test hook
  // pruning votes every four steps (the debug export wants every cost in full)
  p.prune_every = d_dbg_cost ? 0 : 4;
  if (ctx->tune.prune == 0) p.prune_every = 0;                                            // test hook
  if (ctx->tune.prune_every >= 1 && p.prune_every) p.prune_every = ctx->tune.prune_every; // test hook
  if (ctx->tune.cand_cap >= 1 && ctx->tune.cand_cap < p.cand_cap) p.cand_cap = ctx->tune.cand_cap;   // test hook
  const int threads = tw * 32;
  // accelerations one pass can touch: items [pass*T, pass*T + T) span at most this many chunks
  int chunks = chunks_per_pass(threads, p.gs);
  if (chunks > p.n_ic) chunks = p.n_ic;
  if (vd_whole) chunks = p.n_ic;
  p.vd_cols = chunks * kC;
  p.n_pass = (p.n_items + threads - 1) / threads;
  p.tl_slices = tl_kpar(threads, p.gs);
  p.vd_kpar = threads >= p.vd_cols ? threads / p.vd_cols : 1;
  p.gs_sh = log2_exact(p.gs);
  p.vd_sh = log2_exact(p.vd_cols);
  p.t_sh = log2_exact(threads);
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
  p.kappa = 2 * p.delta_max / sin(2 * p.delta_max);
  p.acc_step = p.gv > 1 ? cfg->max_accel / (double)(p.gv - 1) : 0.0;
  p.rate_step = p.gs > 1 ? cfg->max_steer_rate / (double)(p.gs - 1) : 0.0;
  p.inv_L = (float)(1.0 / cfg->wheel_base);
  if (!(p.kappa > 0) || p.kappa > 64) p.kappa = 64;
  p.win_start = (const long long*)d_win_start;
  p.win_len = d_win_len;
  p.win_drive = d_win_drive;
  p.drive_off = (const long long*)d_drive_offsets;
  p.win_off = (const long long*)d_window_offsets;
  p.n_drives = n_drives;
  p.window_frames = cfg->window_frames;
  p.dt_drive = d_dt_per_drive;
  p.vo = d_vo;
  p.gps = d_gps;
  p.imu = d_imu;
  p.seeds = d_seeds;
  p.results = d_results;
  p.out_poses = d_out_poses;
  p.out_steer = d_out_steer;
  p.out_vel = d_out_vel;
  p.out_stride = out_stride;
  p.n_windows = n_windows;
  p.dbg_cost = d_dbg_cost;
  p.dbg_err = d_dbg_err;
  p.run_offsets = (const long long*)d_run_offsets;
  p.n_runs = d_run_offsets ? n_runs : 0;
  // the window deal and the record exchange of this call (vmvo_exchange)
  const bool dealt = ex && ex->world > 1;
  p.sh_world = dealt ? ex->world : 1;
  p.sh_rank = dealt ? ex->rank : 0;
  p.sh_block_sh = dealt && ex->block > 0 ? log2_exact(ex->block) : -1;
  p.n_mirrors = ex ? ex->n_peers : 0;
  for (int q = 0; q < VMVO_MAX_MIRRORS; ++q)
    p.mirrors[q] = q < p.n_mirrors ? (vmvo_window_result*)ex->peer_records[q] : nullptr;
  p.epoch = ex ? ex->epoch : nullptr;
  if (p.run_offsets) {          // whole runs are dealt: run r to rank r % world
    p.n_local = p.n_runs > p.sh_rank ? (p.n_runs - p.sh_rank + p.sh_world - 1) / p.sh_world : 0;
  } else if (dealt && p.sh_block_sh >= 0) {
    const long long nb = (n_windows + ex->block - 1) / ex->block;      // blocks of the global list
    const long long mine = nb > p.sh_rank ? (nb - p.sh_rank + p.sh_world - 1) / p.sh_world : 0;
    p.n_local = mine * ex->block;
  } else {
    p.n_local = n_windows;
  }

  cudaStream_t st = (cudaStream_t)stream;
  VMVO_ON_DEVICE(ctx);
  cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &capture);
  const bool capturing = capture != cudaStreamCaptureStatusNone;

  // Scratch budgets.  Small launches stay within 64 MiB of slots and 256 MiB of preparation records; a
  // large batch (BASELINE configs[3]: 3.0e6 windows) may take up to 1 GiB / 8 GiB where a quarter of
  // the device's free memory covers it -- measured on the 512-drive batch: 112.0 -> 91.7 ms with the
  // preparation records, 88.8 ms with the slots as well (one run in five or six reads 103-117 ms with
  // either: something about where the 6 GB land).  (cudaMemGetInfo only for such launches.)
  size_t mem_quarter = 0;
  auto free_quarter = [&]() {
    if (!mem_quarter) {
      RelaxedCapture relaxed(capturing);
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        free_b = 0;
      }
      mem_quarter = free_b / 4 + 1;
    }
    return mem_quarter;
  };
  // deferred windows: a slot per window up to the budget.  Not with chained seeds (the next
  // window needs this one's optimum at once), rollout outputs or the debug export.
  p.defer_buf = nullptr;
  p.defer_count = nullptr;
  p.defer_slots = p.defer_slot_bytes = 0;
  // Pays where a window's scan is short and its team small (a 32x32 window takes a two-warp team
  // ~30 us, a list of 32 near-ties another ~35 us); an eight-warp team on a dense grid re-scores
  // faster than the dump and the CTA of the second kernel would (measured: -5 %).
  p.defer_min = tw <= 2 ? 8 : 0;
  if (ctx->tune.defer_min >= 0) p.defer_min = ctx->tune.defer_min;   // test hook; 0 = never
  size_t slot_bytes = 0, need = 0;
  long long slots = 0;
  if (p.defer_min > 0 && !d_run_offsets && !d_out_poses && !d_out_steer && !d_out_vel && !d_dbg_cost) {
    const int n_arr = 2 + ((use_vo && use_gps) ? 2 : 0) + (use_imu ? 1 : 0);
    slot_bytes = ((size_t)kDeferHdrBytes + (size_t)n_arr * p.maxp * 8 + (size_t)p.cand_cap * 8 + 127) & ~(size_t)127;
    slots = p.run_offsets ? 0 : (p.n_local < n_windows ? p.n_local : n_windows);
    long long budget = (64LL << 20) / (long long)slot_bytes;
    if (slots > budget) {
      size_t big = (size_t)slots * slot_bytes;
      if (big > ((size_t)1 << 30)) big = (size_t)1 << 30;
      if (big <= free_quarter()) budget = (long long)(big / slot_bytes);
    }
    if (slots > budget) slots = budget;
    need = (size_t)slots * slot_bytes;
  }
  // the scratch buffer of a launch: [window preparation records][counters, 128 bytes][ready words, one
  // per slot][slots] -- counters and ready words side by side, so that ONE memset node clears both
  const size_t ready_bytes = (((size_t)slots * 4 + 127) & ~(size_t)127) + 128;
  if (need > 0) need += ready_bytes;
  size_t prep_total = 0;
  if (want_prep && tw <= 2) {
    prep_total = (size_t)p.n_local * (size_t)prep_rec_bytes;
    if (prep_total > (256ull << 20) && (prep_total > (8ull << 30) || prep_total + need > free_quarter()))
      prep_total = 0;                                      // (a batch this large keeps the fused phases)
  }
  const size_t need_all = need + prep_total;
  vmvo_launch_slot* ls = acquire_launch_slot(ctx, need_all, capturing);
  if (!ls)
    return fail(ctx, VMVO_ERR_UNSUPPORTED, "%d searches of this ctx are in flight or captured in graphs: "
                "no launch slot left", kLaunchSlots);
  // queue head, slot allocation count, finished teams
  unsigned long long* counters = ls->d_counters;
  bool counters_cleared = false;
  p.defer_ready = nullptr;
  p.n_todo = 0;
  p.prep = nullptr;
  p.prep_stride = prep_rec_bytes;
  if (need_all > 0 && ls->d_defer && ls->defer_bytes >= need_all) {
    unsigned char* base = ls->d_defer;
    if (prep_total > 0) {
      p.prep = base;
      base += prep_total;
    }
    if (need > 0 && slots > 0) {
      counters = reinterpret_cast<unsigned long long*>(base);
      p.defer_ready = reinterpret_cast<unsigned*>(base + 128);
      p.defer_buf = base + ready_bytes;
      p.defer_slots = (int)(slots > 0x7fffffff ? 0x7fffffff : slots);
      p.defer_slot_bytes = (int)slot_bytes;
      p.defer_count = reinterpret_cast<unsigned*>(counters + 1);
      VMVO_CUDA(ctx, cudaMemsetAsync(base, 0, 128 + (size_t)p.defer_slots * 4, st));
      counters_cleared = true;
    }
  }
  if (!counters_cleared) VMVO_CUDA(ctx, cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned long long), st));
  p.work_counter = counters;
  p.windows_done = counters + 2;

  const bool dual = use_vo && use_gps;
  // the kernel MODE (vmvo_search_kernels.cuh): lean when nothing outside the default path is asked for
  const bool lean = !d_run_offsets && !d_out_poses && !d_out_steer && !d_out_vel && !d_dbg_cost &&
                    cfg->k_steer == 0.0 && p.gs_sh >= 0 && p.vd_sh >= 0 && p.t_sh >= 0 && ctx->tune.lean != 0;
  rc = !lean ? launch_search_mode0(ctx, p, st, dual, use_imu, f64)
             : p.prep ? launch_search_mode2(ctx, p, st, dual, use_imu, f64)
                      : launch_search_mode1(ctx, p, st, dual, use_imu, f64);
  release_launch_slot(ctx, ls, st, capturing);
  return rc;
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
  return grid_search_impl(ctx, cfg, n_windows, d_win_start, d_win_len, d_win_drive, d_dt_per_drive,
                          d_vo, d_gps, d_imu, false, d_seeds, d_results, d_out_poses, d_out_steer,
                          d_out_vel, out_stride, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, stream);
}

extern "C" int vmvo_grid_search_f64(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                                    const int64_t* d_win_start, const int32_t* d_win_len,
                                    const int32_t* d_win_drive, const double* d_dt_per_drive,
                                    const double* d_vo, const double* d_gps, const double* d_imu,
                                    const double* d_seeds, vmvo_window_result* d_results,
                                    double* d_out_poses, double* d_out_steer, double* d_out_vel,
                                    int32_t out_stride, void* stream) {
  return grid_search_impl(ctx, cfg, n_windows, d_win_start, d_win_len, d_win_drive, d_dt_per_drive,
                          d_vo, d_gps, d_imu, true, d_seeds, d_results, d_out_poses, d_out_steer,
                          d_out_vel, out_stride, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, stream);
}

// stream_f64 != 0: the pose streams are double4 / double (as in vmvo_grid_search_f64)
extern "C" int vmvo_grid_search_chained(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                                        const int64_t* d_win_start, const int32_t* d_win_len,
                                        const int32_t* d_win_drive, const double* d_dt_per_drive,
                                        const void* d_vo, const void* d_gps, const void* d_imu,
                                        int32_t stream_f64, int64_t n_runs,
                                        const int64_t* d_run_offsets, vmvo_window_result* d_results,
                                        double* d_out_poses, double* d_out_steer, double* d_out_vel,
                                        int32_t out_stride, void* stream) {
  if (!d_run_offsets) return fail(ctx, VMVO_ERR_BAD_ARG, "d_run_offsets is NULL");
  return grid_search_impl(ctx, cfg, n_windows, d_win_start, d_win_len, d_win_drive, d_dt_per_drive,
                          d_vo, d_gps, d_imu, stream_f64 != 0, nullptr, d_results, d_out_poses,
                          d_out_steer, d_out_vel, out_stride, nullptr, nullptr, n_runs, d_run_offsets,
                          nullptr, 0, nullptr, nullptr, stream);
}

extern "C" int vmvo_grid_search_debug_f32(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                                          const int64_t* d_win_start, const int32_t* d_win_len,
                                          const int32_t* d_win_drive, const double* d_dt_per_drive,
                                          const float* d_vo, const float* d_gps, const float* d_imu,
                                          const double* d_seeds, vmvo_window_result* d_results,
                                          float* d_scan_cost, float* d_scan_err, void* stream) {
  if (!d_scan_cost || !d_scan_err) return fail(ctx, VMVO_ERR_BAD_ARG, "debug outputs are NULL");
  return grid_search_impl(ctx, cfg, n_windows, d_win_start, d_win_len, d_win_drive, d_dt_per_drive,
                          d_vo, d_gps, d_imu, false, d_seeds, d_results, nullptr, nullptr, nullptr, 0,
                          d_scan_cost, d_scan_err, 0, nullptr, nullptr, 0, nullptr, nullptr, stream);
}

extern "C" int vmvo_grid_search_sharded(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                                        const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                                        int64_t n_windows, const int64_t* d_win_start,
                                        const int32_t* d_win_len, const int32_t* d_win_drive,
                                        const double* d_dt_per_drive, const void* d_vo, const void* d_gps,
                                        const void* d_imu, int32_t stream_f64, const double* d_seeds,
                                        vmvo_window_result* d_results, const vmvo_exchange* ex,
                                        void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  const bool chained = cfg && cfg->seed_mode == VMVO_SEED_CHAINED;
  if (chained && (n_drives < 1 || !d_window_offsets))
    return fail(ctx, VMVO_ERR_BAD_ARG, "seed_mode chained walks the windows of each drive in order: "
                "pass n_drives and d_window_offsets");
  return grid_search_impl(ctx, cfg, n_windows, d_win_start, d_win_len, d_win_drive, d_dt_per_drive,
                          d_vo, d_gps, d_imu, stream_f64 != 0, d_seeds, d_results, nullptr, nullptr,
                          nullptr, 0, nullptr, nullptr, chained ? n_drives : 0,
                          chained ? d_window_offsets : nullptr, ex, n_drives, d_drive_offsets,
                          d_window_offsets, stream);
}
