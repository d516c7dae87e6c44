/*
 * vmvo_b200.h -- C ABI of the B200-native VMVO window search (libvmvo_b200.so).
 *
 * The reference (AdityaNG/VehicleModelVisualOdometry) is pure Python and has no FFI; the
 * boundary it offers is the set of Python signatures on the optimize_trajectory_v2 path.
 * Each entry point below names the reference interface it stands in for (file:line,
 * relative to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - plain C types; every `const T* d_xxx` / `T* d_xxx` is a DEVICE pointer owned by the
 *     caller (e.g. torch.Tensor.data_ptr()); the library allocates nothing persistent
 *     except the opaque vmvo_ctx (work counters, the scratch of deferred windows, error text);
 *   - calls are stream-ordered on the caller's `stream` (a cudaStream_t passed as void*)
 *     and asynchronous; one host thread per ctx at a time.  Every search launch owns its scratch
 *     (queue head, deferred-window slots) until it has completed, so searches of one ctx may
 *     overlap on different streams; the calling thread's current device is left as it was;
 *   - return value: vmvo_status (0 = ok); vmvo_last_error(ctx) gives the text.
 *   - pose streams are float4 (x [m], y [m], theta [rad], v [m/s]) per frame (double4 in the
 *     _f64 entry points), all drives concatenated; `d_drive_offsets[n_drives + 1]` delimits them; `d_time` is float64 [s].
 */
#ifndef VMVO_B200_H
#define VMVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMVO_ABI_VERSION 2

typedef struct vmvo_ctx vmvo_ctx;

typedef enum vmvo_status {
  VMVO_OK = 0,
  VMVO_ERR_BAD_ARG = 1,
  VMVO_ERR_CUDA = 2,
  VMVO_ERR_UNSUPPORTED = 3
} vmvo_status;

/* enumerations used inside vmvo_search_cfg */
enum { VMVO_WINDOW_FRAMES = 0, VMVO_WINDOW_TIME = 1 };
enum { VMVO_TARGET_TIME = 0, VMVO_TARGET_TRAVERSE = 1 };
enum { VMVO_SEED_DATA = 0, VMVO_SEED_GIVEN = 1, VMVO_SEED_CHAINED = 2 };
enum { VMVO_PRIMARY_VO = 0, VMVO_PRIMARY_GPS = 1 };

/* per-window status bits */
enum {
  VMVO_WIN_EMPTY = 1,     /* fewer than two targets, N == 0  (vmvo/utils/mpc.py:42-43)      */
  VMVO_WIN_NONFINITE = 2, /* NaN/Inf among the window's inputs: argmin degenerates to 0     */
  VMVO_WIN_TOO_LONG = 4,  /* more poses than cfg.max_window_poses: window not searched      */
  VMVO_WIN_NO_FRAMES = 8  /* no pose in the window's extent (set with EMPTY): the reference's
                             assert "No frames found", vmvo/schema.py:122                    */
};

/* per-file status bits of vmvo_csv_parse_f64 */
enum {
  VMVO_CSV_BAD_NUMBER = 1,      /* a wanted field is not a number (pandas: column dtype object)   */
  VMVO_CSV_TOO_MANY_FIELDS = 2, /* a row has more fields than the header (pandas: ParserError)    */
  VMVO_CSV_BAD_ROT = 4,         /* a rot field does not hold nine numbers                         */
  VMVO_CSV_UNSORTED = 8         /* the sorted_slot column decreases somewhere                     */
};

/* per-sequence failure kinds of vmvo_rollout_* (the two asserts of bicycle_model.py:48-62) */
enum { VMVO_FAIL_NONE = 0, VMVO_FAIL_STEER = 1, VMVO_FAIL_ACCEL = 2 };

/*
 * Search configuration: the derived spec of DESIGN.md section 2.
 * Constants default to vmvo/constants.py:3-7; horizon_time to optimize_trajectory_v2.py:35.
 */
typedef struct vmvo_search_cfg {
  int32_t grid_v;           /* G_v accelerations in [-max_accel, +max_accel]               */
  int32_t grid_s;           /* G_s steering rates in [-max_steer_rate, +max_steer_rate]    */
  int32_t window_mode;      /* VMVO_WINDOW_*                                               */
  int32_t window_frames;    /* W steps => W+1 poses per window (frames mode)               */
  int32_t horizon_frames;   /* int(horizon_time * FPS) (time mode): n_windows = n - 2*this */
  int32_t target_mode;      /* VMVO_TARGET_*                                               */
  int32_t target_offset;    /* 1: state k vs target k-1 (mpc.py:70-78); 0: state k vs k    */
  int32_t seed_mode;        /* VMVO_SEED_*                                                 */
  int32_t primary;          /* stream that defines the window frame and the seeds          */
  int32_t max_window_poses; /* capacity per window (<= 256)                                */
  double horizon_time;      /* seconds (time mode)                                         */
  double w_vo, w_gps, w_imu;/* cost weights; a zero weight drops the term                  */
  double k_steer;           /* K of mpc.py:31                                              */
  double wheel_base, steering_ratio, max_steer, max_accel, max_steer_rate;
} vmvo_search_cfg;

/* One record per window; also the unit of the multi-GPU gather (64 bytes). */
typedef struct vmvo_window_result {
  int32_t best_idx;      /* flat index i*G_s + j of the argmin hypothesis, -1 if none      */
  int32_t n_steps;       /* N = number of targets - 1                                      */
  int32_t status;        /* VMVO_WIN_* bits                                                */
  int32_t n_rescored;    /* hypotheses re-scored in float64 (diagnostic; may vary run to run) */
  double best_cost;      /* float64 cost of the argmin hypothesis                          */
  double v_seed, s_seed; /* window seeds V_w [m/s], S_w [deg]                              */
  double x1, y1, theta1; /* pose after the first step of the best rollout (local frame)    */
} vmvo_window_result;

/* ---- context ---------------------------------------------------------------------- */
int vmvo_abi_version(void);
int vmvo_ctx_create(int device, vmvo_ctx** out);
int vmvo_ctx_destroy(vmvo_ctx* ctx);
const char* vmvo_last_error(const vmvo_ctx* ctx);
/* fills *cfg with the defaults described above (32x32 grid, frames mode, W = 30) */
void vmvo_search_cfg_default(vmvo_search_cfg* cfg);
/* number of windows of a drive of n_frames frames: max(0, n - 2*horizon)
 * (vmvo/scripts/optimize_trajectory_v2.py:48) */
int64_t vmvo_window_count(const vmvo_search_cfg* cfg, int64_t n_frames);

/* ---- a8: window extents ------------------------------------------------------------
 * Replaces the per-window np.searchsorted pair of Trajectory.sub_trajectory_from_time
 * (vmvo/schema.py:117-127) for every window of every drive at once.
 * d_window_offsets[n_drives + 1] is the prefix sum of vmvo_window_count per drive.
 * Outputs (length d_window_offsets[n_drives]): absolute start frame, pose count, drive. */
int vmvo_plan_windows(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                      const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                      int64_t n_windows, const double* d_time, int64_t* d_win_start,
                      int32_t* d_win_len, int32_t* d_win_drive, void* stream);

/* ---- a7, a9, a10, a11: the fused window search --------------------------------------
 * Replaces mpc_run (vmvo/utils/mpc.py:14-122) and the window preparation around it
 * (schema.py:59-115 local frame, mpc.py:125-141 decimation, optimize_trajectory_v2.py:57-91
 * seed + re-rollout) for n_windows windows.  d_gps / d_imu / d_seeds may be NULL when the
 * cfg does not use them.  d_seeds is [n_windows][2] = (V_w, S_w) for VMVO_SEED_GIVEN.
 * Optional outputs (NULL to skip), row stride out_stride (>= max steps):
 *   d_out_poses [n_windows][out_stride][3]  rollout (x, y, theta) of the argmin hypothesis
 *   d_out_steer [n_windows][out_stride]     its steering sequence [deg] (mpc_run's return)
 *   d_out_vel   [n_windows][out_stride]     its velocity sequence [m/s]                  */
int vmvo_grid_search_f32(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                         const int64_t* d_win_start, const int32_t* d_win_len,
                         const int32_t* d_win_drive, const double* d_dt_per_drive,
                         const float* d_vo, const float* d_gps, const float* d_imu,
                         const double* d_seeds, vmvo_window_result* d_results,
                         double* d_out_poses, double* d_out_steer, double* d_out_vel,
                         int32_t out_stride, void* stream);

/* The same search over float64 pose streams: double4 (x, y, theta, v) per frame (32-byte
 * records, 16-byte aligned), d_imu float64.  The reference computes in float64 throughout
 * (vmvo/schema.py:21-28 holds List[float]); the Python facades use this entry point so that no
 * input is rounded on its way to the GPU.                                                   */
int vmvo_grid_search_f64(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                         const int64_t* d_win_start, const int32_t* d_win_len,
                         const int32_t* d_win_drive, const double* d_dt_per_drive,
                         const double* d_vo, const double* d_gps, const double* d_imu,
                         const double* d_seeds, vmvo_window_result* d_results,
                         double* d_out_poses, double* d_out_steer, double* d_out_vel,
                         int32_t out_stride, void* stream);

/* seed_mode = VMVO_SEED_CHAINED (optimize_trajectory_v2.py:46,72,146): the steering seed of a
 * window is the last steering angle of the previous window's optimum, 0 for the first window
 * of a drive.  Windows of a drive are therefore searched in order by one team; drives run in
 * parallel.  d_run_offsets[n_runs + 1] delimits, in window indices of this call, the runs
 * (normally one per drive: the d_window_offsets of vmvo_plan_windows).  stream_f64 = 0: the
 * streams are float4 / float as in vmvo_grid_search_f32; 1: double4 / double as in _f64.    */
int vmvo_grid_search_chained(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                             const int64_t* d_win_start, const int32_t* d_win_len,
                             const int32_t* d_win_drive, const double* d_dt_per_drive,
                             const void* d_vo, const void* d_gps, const void* d_imu,
                             int32_t stream_f64, int64_t n_runs, const int64_t* d_run_offsets,
                             vmvo_window_result* d_results, double* d_out_poses,
                             double* d_out_steer, double* d_out_vel, int32_t out_stride,
                             void* stream);

/* Test hook: the same search, additionally exporting the FP32 scan cost of EVERY hypothesis
 * and the width of its error band, [n_windows][grid_v * grid_s] each, so tests can check
 * |scan - float64| <= band against the oracle.  Not for production use.  */
int vmvo_grid_search_debug_f32(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int64_t n_windows,
                               const int64_t* d_win_start, const int32_t* d_win_len,
                               const int32_t* d_win_drive, const double* d_dt_per_drive,
                               const float* d_vo, const float* d_gps, const float* d_imu,
                               const double* d_seeds, vmvo_window_result* d_results,
                               float* d_scan_cost, float* d_scan_err, void* stream);

/* ---- sharding the window list over the GPUs of a box, and the one exchange ------
 * Windows are independent (optimize_trajectory_v2.py:48-146 is a loop over them), so R ranks share
 * one GLOBAL window list (the plan of all drives, resident on every rank) with no data-path
 * collective: the windows are dealt block-cyclically -- window w belongs to rank (w / block) %
 * world -- and every rank searches its own.  The only exchange is the 64-byte records, and it is
 * fused into the kernels that produce and consume them:
 *   - d_results is a gather buffer with one record per GLOBAL window on every rank; the epilogue
 *     of the search stores record w into its own buffer AND into peer_records[q][w] of every
 *     peer (buffers mapped through CUDA IPC, below): plain 64-byte stores over NVLink while the
 *     kernel is still searching;
 *   - arrival: every sharded search advances the step counter epoch[0] (device memory, so that a
 *     replayed CUDA graph counts on).  After its last record store a rank PUBLISHES: one
 *     st.release.sys of the step number into the word it owns in every peer's flag array
 *     (peer_flags[q]; stream order puts it behind every record store of the step).  A consumer
 *     WAITS: it spins (ld.acquire.sys) on local_flags[q], q != rank, until each holds the step
 *     number.  vmvo_exchange_publish / vmvo_exchange_wait are the stand-alone forms;
 *     vmvo_write_back_range fuses both into the kernel that consumes the records, so the step
 *     needs no collective, no extra launch and no host barrier.  A wait gives up after ~10 s
 *     (a dead peer must not hang the GPU) and sets epoch[1] != 0.
 * Two buffer sets (records + flags + epoch) used alternately make this race-free without any
 * other synchronisation: a rank starts step s+2 only after its step s+1 wait, i.e. after every
 * peer has finished reading buffer set s%2.
 * All pointers are device pointers; the struct itself is host memory, read at launch time.   */
#define VMVO_MAX_MIRRORS 16
typedef struct vmvo_exchange {
  int32_t world;         /* ranks sharing the window list (1: no sharding, no exchange)           */
  int32_t rank;
  int32_t block;         /* windows per block of the deal, a power of two; 0: the call's windows
                            are all this rank's (contiguous shards: pass offset pointers)        */
  int32_t n_peers;       /* valid entries below: world - 1, or 0 (no mirrors, no flags)           */
  void* peer_records[VMVO_MAX_MIRRORS];      /* peers' gather buffers, indexed like d_results     */
  uint32_t* peer_flags[VMVO_MAX_MIRRORS];    /* the word this rank owns in each peer's flags      */
  const uint32_t* local_flags;               /* this rank's flag array [world]                    */
  uint32_t* epoch;                           /* [0] step counter, [1] wait-timeout indicator      */
} vmvo_exchange;

/* The fused search over this rank's share of the n_windows GLOBAL windows of n_drives drives
 * (d_results, and the plan arrays when given, cover all of them).  d_drive_offsets / d_window_offsets
 * [n_drives + 1] are the frame and window prefix sums of vmvo_plan_windows.  The plan arrays may be
 * all NULL in VMVO_WINDOW_FRAMES mode: the kernel then derives each window's extent itself (window i
 * of a drive = poses i .. i + window_frames, optimize_trajectory_v2.py:48-56) and the step needs no
 * planning launch.  stream_f64 as in vmvo_grid_search_chained; d_seeds only for VMVO_SEED_GIVEN; with
 * VMVO_SEED_CHAINED whole drives are dealt (drive d to rank d % world) and walked in order.
 * ex == NULL or ex->world <= 1: every window, no exchange -- the single-GPU pipeline's entry point. */
int vmvo_grid_search_sharded(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                             const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                             int64_t n_windows, const int64_t* d_win_start, const int32_t* d_win_len,
                             const int32_t* d_win_drive, const double* d_dt_per_drive,
                             const void* d_vo, const void* d_gps, const void* d_imu,
                             int32_t stream_f64, const double* d_seeds,
                             vmvo_window_result* d_results, const vmvo_exchange* ex, void* stream);
/* Arrival word of the current step to every peer / wait for every peer's (see above).          */
int vmvo_exchange_publish(vmvo_ctx* ctx, const vmvo_exchange* ex, void* stream);
int vmvo_exchange_wait(vmvo_ctx* ctx, const vmvo_exchange* ex, void* stream);
/* A device buffer other processes of the box can map: cudaMalloc + cudaIpcGetMemHandle.
 * h_handle receives the 64-byte IPC handle (host memory) to send to the peers.               */
int vmvo_peer_buffer_create(vmvo_ctx* ctx, int64_t bytes, void** d_ptr, uint8_t* h_handle);
int vmvo_peer_buffer_destroy(vmvo_ctx* ctx, void* d_ptr);
/* Map / unmap a peer's buffer from its handle (cudaIpcOpenMemHandle with lazy peer access).  */
int vmvo_peer_buffer_open(vmvo_ctx* ctx, const uint8_t* h_handle, void** d_ptr);
int vmvo_peer_buffer_close(vmvo_ctx* ctx, void* d_ptr);

/* ---- a12: write-back and blends -----------------------------------------------------
 * Replaces optimize_trajectory_v2.py:32-33,122-137: output columns start as the VO stream,
 * window i overwrites x,y[i : i+N_i] with its local-frame rollout (later windows win),
 * theta[i] / velocity[i] become the VO/GPS blends when d_gps != NULL.
 * total_frames = d_drive_offsets[n_drives] (passed by value so the call stays
 * asynchronous).  Outputs are float64 [total_frames] each.                              */
int vmvo_write_back_f32(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                        int64_t total_frames, const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                        const double* d_dt_per_drive, const float* d_vo, const float* d_gps,
                        const vmvo_window_result* d_results, double* d_out_x, double* d_out_y,
                        double* d_out_theta, double* d_out_vel, void* stream);
/* float64 pose streams (double4 per frame), otherwise identical */
int vmvo_write_back_f64(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                        int64_t total_frames, const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                        const double* d_dt_per_drive, const double* d_vo, const double* d_gps,
                        const vmvo_window_result* d_results, double* d_out_x, double* d_out_y,
                        double* d_out_theta, double* d_out_vel, void* stream);

/* The same write-back for frames [frame_lo, frame_hi) only (outputs stay [total_frames]; other
 * frames are not touched), reading the records of ALL windows from a gather buffer: the consumer of
 * the sharded search.  stream_f64 selects float4 / double4 pose streams.  With ex != NULL and
 * ex->n_peers > 0 the kernel first publishes this rank's arrival word and then waits for every
 * peer's before it reads a record (vmvo_exchange above): the whole exchange rides on the two
 * kernels of the step.                                                                          */
int vmvo_write_back_range(vmvo_ctx* ctx, const vmvo_search_cfg* cfg, int32_t n_drives,
                          int64_t total_frames, int64_t frame_lo, int64_t frame_hi,
                          const int64_t* d_drive_offsets, const int64_t* d_window_offsets,
                          const double* d_dt_per_drive, const void* d_vo, const void* d_gps,
                          int32_t stream_f64, const vmvo_window_result* d_results, double* d_out_x,
                          double* d_out_y, double* d_out_theta, double* d_out_vel,
                          const vmvo_exchange* ex, void* stream);

/* ---- a1, a2: batched model rollout --------------------------------------------------
 * Replaces BicycleModel.run / run_sequence (vmvo/bicycle_model.py:40-92) for n_seq
 * independent sequences of n_steps steps.  d_steer / d_vel are [n_seq][n_steps]
 * (steering-wheel degrees, m/s), d_state0 [n_seq][4] = (x, y, theta, velocity).
 * d_out [n_seq][n_steps][3] = (x, y, theta) after each step.  d_fail [n_seq][2] =
 * (VMVO_FAIL_* kind, step index) of the first violated bound, kind 0 when feasible.   */
int vmvo_rollout_f64(vmvo_ctx* ctx, int64_t n_seq, int32_t n_steps, const double* d_steer,
                     const double* d_vel, double dt, const double* d_state0,
                     double max_steer, double max_accel, double* d_out, int32_t* d_fail,
                     void* stream);
int vmvo_rollout_f32(vmvo_ctx* ctx, int64_t n_seq, int32_t n_steps, const float* d_steer,
                     const float* d_vel, float dt, const float* d_state0, float max_steer,
                     float max_accel, float* d_out, int32_t* d_fail, void* stream);

/* The scalar caller's form of the same operator (BicycleModel.run / run_sequence are called one step
 * or one short sequence at a time, vmvo/bicycle_model.py:40-92): HOST pointers in and out, one
 * sequence, synchronous.  The controls go through a pinned buffer of the ctx that the GPU reads in
 * place; one launch and one stream synchronisation per call, no allocation, no torch.
 * h_state0 [4], h_out [n_steps][3], h_fail [2] as above.                                          */
int vmvo_rollout_host_f64(vmvo_ctx* ctx, int32_t n_steps, const double* h_steer, const double* h_vel,
                          double dt, const double* h_state0, double max_steer, double max_accel,
                          double* h_out, int32_t* h_fail);

/* ---- a10: cost of given control sequences -------------------------------------------
 * The closure `cost` inside mpc_run (vmvo/utils/mpc.py:56-85): n_seq steering sequences
 * [n_seq][n_steps] at constant speed `velocity`, against one target polyline
 * d_target_xy [n_steps + 1][2]; d_cost [n_seq].                                        */
int vmvo_sequence_cost_f64(vmvo_ctx* ctx, int64_t n_seq, int32_t n_steps, const double* d_steer,
                           double velocity, double dt, const double* d_target_xy, double k_steer,
                           double* d_cost, void* stream);

/* ---- a7, a8, a9 as stand-alone operators (used by the Trajectory facade) ------------ */
/* Trajectory.sub_trajectory (vmvo/schema.py:59-115): local frame of n poses.           */
int vmvo_extract_window_f64(vmvo_ctx* ctx, int32_t n, const double* d_x, const double* d_y,
                            const double* d_theta, double* d_lx, double* d_ly, double* d_lth,
                            void* stream);
/* np.searchsorted(time, t0, 'left'), np.searchsorted(time, t1, 'right')
 * (vmvo/schema.py:119-120); d_extent[2].                                               */
int vmvo_time_extent_f64(vmvo_ctx* ctx, int64_t n, const double* d_time, double t0, double t1,
                         int64_t* d_extent, void* stream);
/* traverse_trajectory (vmvo/utils/mpc.py:125-141): kept indices and their count.       */
int vmvo_traverse_f64(vmvo_ctx* ctx, int32_t n, const double* d_xy, double D, int32_t* d_keep,
                      int32_t* d_count, void* stream);

/* ---- next row: trajectory pre-processing, batched over drives ---------------
 * All arrays float64, drives concatenated, d_offsets[n_drives + 1] in frames.                */
/* smoothen_traj (vmvo/utils/trajectory.py:68-99): trailing moving average, per drive.       */
int vmvo_smooth_f64(vmvo_ctx* ctx, int32_t n_drives, int64_t total_frames, const int64_t* d_offsets,
                    const double* d_x, const double* d_y, int32_t window, double* d_out_x,
                    double* d_out_y, void* stream);
/* process_vo_trajectory (vmvo/utils/trajectory.py:13-65): d_rot [frames][9] row-major 3x3,
 * d_stamp_ms the Timestamp column.  Outputs [frames] each.  yaw_f32 != 0: the rotation entries
 * are float32 values (the cached trajectory, bdd_raw.py:163-164) and the yaw is atan2f in
 * float32, as np.arctan2 on float32 scalars is (trajectory.py:28); 0: float64 atan2.          */
int vmvo_vo_prepare_f64(vmvo_ctx* ctx, int32_t n_drives, int64_t total_frames,
                        const int64_t* d_offsets, const double* d_x, const double* d_y,
                        const double* d_rot, const double* d_stamp_ms, double scale, int32_t window,
                        int32_t yaw_f32,
                        double* d_out_x, double* d_out_y, double* d_out_theta, double* d_out_vel,
                        double* d_out_time, void* stream);
/* process_gps_trajectory (vmvo/utils/trajectory.py:177-335, with geodetic_to_euclidean
 * :120-174).  A drive of n fixes yields n + 1 points (the reference's leading duplicate of the
 * origin): outputs are [total_frames + n_drives], drive d starting at d_offsets[d] + d; theta has
 * n valid entries per drive (NaN in the last slot).  d_status[d] = 1 where the reference would
 * raise IndexError (log ending on a fresh fix).  d_scratch: vmvo_gps_prepare_scratch_bytes.   */
int64_t vmvo_gps_prepare_scratch_bytes(int64_t total_frames, int32_t n_drives);
int vmvo_gps_prepare_f64(vmvo_ctx* ctx, int32_t n_drives, int64_t total_frames,
                         const int64_t* d_offsets, const double* d_lat, const double* d_lon,
                         const double* d_speed, const double* d_stamp_ms, int32_t window,
                         void* d_scratch, double* d_out_x, double* d_out_y, double* d_out_theta,
                         double* d_out_vel, double* d_out_time, int32_t* d_status, void* stream);

/* ---- next row: the reference's on-disk formats -------------------------------
 * <id>.csv, the Android log read by pd.read_csv (vmvo/datasets/bdd/bdd_raw.py:53-55; columns
 * Timestamp [ms], Latitude, Longitude, heading, speed: vmvo/utils/trajectory.py:191-228), and
 * <id>_traj.csv, the cached VO trajectory (bdd_raw.py:150-168, 331-332: x, y, z, rot with
 * rot = str(3x3 ndarray), a quoted field spanning three lines).  The raw bytes of n_files files
 * sit concatenated in d_bytes (16-byte aligned; file f occupies [file_off[f], file_off[f] +
 * file_len[f]), every file_off a multiple of 16, the buffer readable up to the next multiple of
 * 16).  h_* are HOST copies of the same two arrays (grid sizes are derived from them).
 * Well-formed CSV is assumed: quotes delimit fields or are doubled inside quoted fields.
 *
 *   1. vmvo_csv_count_rows -> d_row_counts[n_files]: non-blank lines per file (header included;
 *      newlines inside quoted fields do not end a line).  d_scratch: vmvo_csv_scratch_bytes.
 *   2. the caller turns the counts into d_row_off[n_files + 1] (exclusive prefix sum) and calls
 *      vmvo_csv_index_rows (same d_scratch, untouched in between) -> d_row_starts[total_rows]:
 *      byte offset of every line.
 *   3. vmvo_csv_parse_f64: the first line of a file is its header; for every other line the
 *      field with index c of file f goes to output slot d_colmap[f * 64 + c]: -1 = ignored,
 *      0 .. n_slots-1 = number converted like pandas' default converter (precise_xstrtod:
 *      bit-identical to what pd.read_csv returns, including its last-digit behaviour on 16-17
 *      digit inputs), 1000 = the rot field (nine numbers, each the correctly rounded double
 *      rounded to float32 like np.array(tokens).astype(np.float32), bdd_raw.py:157-165).
 *      d_out [n_slots][n_data] and d_rot [n_data][9] (may be NULL), n_data = total_rows - n_files,
 *      data rows of all files back to back.  Empty / NA fields and short rows give NaN.
 *      d_n_fields[f] = fields in the header of file f.  sorted_slot >= 0: that slot is checked to
 *      be non-decreasing within each file (bdd_raw.py:55 sorts by Timestamp).
 *      d_status[n_files]: VMVO_CSV_* bits.                                                    */
int64_t vmvo_csv_scratch_bytes(int32_t n_files, const int64_t* h_file_len);
int vmvo_csv_count_rows(vmvo_ctx* ctx, const uint8_t* d_bytes, int32_t n_files,
                        const int64_t* h_file_off, const int64_t* h_file_len,
                        const int64_t* d_file_off, const int64_t* d_file_len, void* d_scratch,
                        int64_t* d_row_counts, void* stream);
int vmvo_csv_index_rows(vmvo_ctx* ctx, const uint8_t* d_bytes, int32_t n_files,
                        const int64_t* h_file_off, const int64_t* h_file_len,
                        const int64_t* d_file_off, const int64_t* d_file_len, void* d_scratch,
                        const int64_t* d_row_off, int64_t* d_row_starts, void* stream);
int vmvo_csv_parse_f64(vmvo_ctx* ctx, const uint8_t* d_bytes, int32_t n_files,
                       const int64_t* d_file_off, const int64_t* d_file_len,
                       const int64_t* d_row_off, const int64_t* d_row_starts, int64_t total_rows,
                       const int32_t* d_colmap, const int32_t* d_n_fields, int32_t n_slots,
                       int32_t sorted_slot, double* d_out, double* d_rot, int32_t* d_status,
                       void* stream);

/* tan of road-wheel angles [rad] exactly as the search tabulates it (tan delta of
 * vmvo/bicycle_model.py:66-68 in float32): a polynomial within 1 ulp for |delta| <= 0.62 rad
 * (steering is mechanically limited: MAX_STEER / ratio = 0.605 rad), tanf beyond.            */
int vmvo_tan_steer_f32(vmvo_ctx* ctx, int64_t n, const float* d_delta, float* d_out, void* stream);

/* ---- measurement helpers ------------------------------------------------------------ */
/* Issue-rate microbenchmarks used for the roofline denominators (bench.py): each thread
 * runs `iters` dependent-free MUFU (kind 0: sin+cos pairs), FFMA (kind 1) or DFMA (kind 2)
 * operations; d_sink receives one float per thread.                                     */
int vmvo_peak_probe(vmvo_ctx* ctx, int32_t kind, int32_t blocks, int32_t threads,
                    int32_t iters, float* d_sink, void* stream);
/* Test / tuning hook, NOT part of the product surface: overrides one of the launch heuristics of
 * this ctx ("team_warps", "fast_scan", "cand_cap", "defer_min", "max_ctas_per_sm", "defer_warps", "cta_teams", "pdl", "prep",
 * "prune" (0: every scan runs to its last step), "prune_every" (steps between two pruning votes), "lean" (0: the
 * generic kernel mode)); value < 0 restores the library's own choice.  None of them changes a result.  The
 * environment is never consulted.                                                               */
int vmvo_debug_set_tuning(vmvo_ctx* ctx, const char* key, int32_t value);
/* kernels launched by this ctx since creation (for bench.py's gpu_launches)             */
int64_t vmvo_launch_count(const vmvo_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VMVO_B200_H */
