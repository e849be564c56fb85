/* slslam_b200 — C ABI of the B200-native LBA / PO solver.
 *
 * This is the drop-in boundary: what the reference does inside `ceres::Solve(options, &problem, &summary)`
 * for an LBAProblem (reference src/slam.cpp:663, 944) or a POProblem (reference src/slam.cpp:1293) is one call
 * here.  Plain pointers and sizes only; no C++/torch types.  The reference-facing C++ classes
 * (include/lba_problem.h, include/po_problem.h, include/ceres/ceres.h) are thin wrappers over these calls;
 * see INTEGRATION.md.
 *
 * There is NO CPU fallback: every solve entry point returns SLSLAM_ERR_CUDA when no sm_100 device is usable.
 * Never throws.  On any error `params_inout` is left untouched.
 */
#ifndef SLSLAM_B200_H_
#define SLSLAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  SLSLAM_OK = 0,
  SLSLAM_ERR_INVALID = -1,      /* null pointer, negative size, index out of range */
  SLSLAM_ERR_UNSUPPORTED = -2,  /* problem exceeds a kernel limit (see slslam_lba_limits) */
  SLSLAM_ERR_CUDA = -3,         /* no device / CUDA runtime error (message via slslam_last_error) */
  SLSLAM_ERR_NUMERICAL = -4     /* non-finite input parameters */
};

/* Solver::Summary::termination_type as far as this path can produce it (Ceres 1.7.0 names). */
enum {
  SLSLAM_NO_CONVERGENCE = 0,      /* max_num_iterations reached */
  SLSLAM_GRADIENT_TOLERANCE = 1,
  SLSLAM_FUNCTION_TOLERANCE = 2,
  SLSLAM_PARAMETER_TOLERANCE = 3,
  SLSLAM_NUMERICAL_FAILURE = 4    /* max_num_consecutive_invalid_steps reached */
};

/* One LBA window: the arrays LBAProblem holds (reference src/lba_problem.h:188-196, filled at
 * src/slam.cpp:899-920) plus the constants that are literals in the reference.  Borrowed, never freed. */
typedef struct slslam_lba_desc {
  int32_t num_cameras;            /* lba_param_t::num_cameras       (lba_problem.h:123-130) */
  int32_t num_lines;              /* lba_param_t::num_lines */
  int32_t num_observations;       /* lba_param_t::num_observations */
  int32_t max_iterations;         /* lba_param_t::num_iterations -> Solver::Options::max_num_iterations (lba_problem.cpp:124) */
  const int32_t* camera_index;    /* [N]   LBAProblem::camera_index() */
  const int32_t* line_index;      /* [N]   LBAProblem::line_index() */
  const int32_t* fixed_index;     /* [2N]  [2i] camera constant, [2i+1] line constant; sticky per block (lba_problem.cpp:88-91) */
  const double* observations;     /* [8N]  normalised endpoints x0 y0 x1 y1 (cam A)  x2 y2 x3 y3 (cam B) */
  int32_t robust;                 /* FLAGS_robust (lba_problem.cpp:35): HuberLoss on the squared norm of the 4-vector */
  double huber_delta;             /* 1/406.05 at lba_problem.cpp:78-80; <= 0 selects that default */
  double baseline;                /* 0.12 literal at lba_problem.h:101; < 0 selects that default */
  /* Ceres 1.7.0 Solver::Options the reference leaves at their defaults; a value <= 0 selects the default
   * (1e-6, 1e-10, 1e-8, 1e4).  Pass a tiny positive number (e.g. 1e-300) to disable a tolerance. */
  double function_tolerance;
  double gradient_tolerance;
  double parameter_tolerance;
  double initial_trust_region_radius;
} slslam_lba_desc;

/* One pose graph: the arrays POProblem holds (reference src/po_problem.h:139-143, filled at src/slam.cpp:1261-1287). */
typedef struct slslam_po_desc {
  int32_t num_poses;              /* number of 6-vectors in the parameter array (kfs.size(), slam.cpp:1264) */
  int32_t num_edges;              /* POProblem::num_size() */
  int32_t max_iterations;         /* POProblem ctor arg n (=10 at slam.cpp:1283) */
  const int32_t* pose_index_1;    /* [E] */
  const int32_t* pose_index_2;    /* [E] */
  const double* constraints;      /* [6E] measured T_{2<-1} as (angle-axis, t) */
  double function_tolerance, gradient_tolerance, parameter_tolerance, initial_trust_region_radius;
} slslam_po_desc;

/* The fields of ceres::Solver::Summary the reference reads (src/slam.cpp:949-952) plus diagnostics. */
typedef struct slslam_summary {
  double initial_cost;            /* 1/2 sum rho, includes residual blocks whose parameter blocks are all constant */
  double final_cost;
  double fixed_cost;              /* the part contributed by all-constant residual blocks */
  double gradient_max_norm;       /* max |J^T r| at the last linearisation */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t termination_type;
  int32_t iterations;             /* linear solves performed = successful + unsuccessful (+1 if stopped by a tolerance) */
} slslam_summary;

/* Per-iteration record, SLSLAM_TRACE_WIDTH doubles:
 * cost, trial_cost, model_cost_change, radius, step_norm, accepted(1/0/-1 invalid), gradient_max_norm, relative_decrease */
#define SLSLAM_TRACE_WIDTH 8

/* Limits of the tiled solve kernel (the fast path; also what the device-resident, pipelined and slslam_lba_batch_* entry
 * points accept).  slslam_lba_solve / slslam_lba_solve_batch route windows beyond them -- the reference's
 * --ba_window_size 20 / 40 shapes: up to 2 W camera blocks, W free, lines seen by most cameras -- to a general kernel
 * (one CTA per window, no limit on camera blocks or observations per line, max_free_cameras_general free cameras). */
typedef struct slslam_lba_limits {
  int32_t max_cameras;            /* parameter blocks, free + constant */
  int32_t max_free_cameras;       /* reduced camera system is 6*max_free_cameras square */
  int32_t max_observations_per_line;
  int32_t max_cluster_size;
  int32_t max_free_cameras_general;   /* general kernel: reduced camera system up to 6*64 square */
} slslam_lba_limits;

int slslam_version(void);
const char* slslam_strerror(int code);
const char* slslam_last_error(void);          /* thread-local detail for the last SLSLAM_ERR_CUDA */
int slslam_device_count(void);                /* usable sm_100 devices; 0 means every solve call fails */
void slslam_lba_get_limits(slslam_lba_limits* out);
/* Diagnostic: measured fp64 FMA throughput (TFLOP/s, 2 flops per DFMA) of `device` (< 0: current) from a register-only
 * kernel of independent DFMA chains, best of a few repetitions (~10 ms).  Both solvers are fp64-issue / latency bound
 * (the whole LM loop runs out of shared memory), so this -- not HBM bandwidth -- is the ceiling their arithmetic sees. */
int slslam_measure_fp64_peak(int32_t device, double* tflops_out, double* sm_clock_mhz_out);

/* ---- what replaces ceres::Solve for one LBAProblem: H2D, device LM loop, D2H; parameters updated in place ---- */
int slslam_lba_solve(const slslam_lba_desc* desc, double* params_inout, slslam_summary* summary_out);

/* Independent windows in one launch (one thread-block cluster per window).  params_inout[i] has 6C_i+4L_i doubles. */
int slslam_lba_solve_batch(int32_t n, const slslam_lba_desc* descs, double* const* params_inout,
                           slslam_summary* summaries_out);

/* The same with DEVICE-resident inputs: every array the descs point to, and params_dev_inout[i], are device pointers on
 * the current device (typically slices of a buffer NCCL just received: SURVEY.md 8e, window w -> rank w mod G).  Nothing
 * is staged or copied: the plan kernel reads the arrays where they are, the parameters are updated IN PLACE, and the
 * summaries go to summaries_dev_out (device, may be NULL) and / or summaries_host_out (host, may be NULL).  With
 * summaries_host_out == NULL the call returns once the solve is enqueued on `cuda_stream`; otherwise it waits for it.
 * Index errors are found on the device (SLSLAM_ERR_INVALID); non-finite parameters are not screened; windows that need
 * the host planner (a camera observing one line twice) return SLSLAM_ERR_UNSUPPORTED; on a failure detected after the
 * launch the parameters may have been modified.  16-byte alignment of `observations` is required. */
int slslam_lba_solve_batch_device(int32_t n, const slslam_lba_desc* descs, double* const* params_dev_inout,
                                  slslam_summary* summaries_dev_out, slslam_summary* summaries_host_out, void* cuda_stream);

/* Host wall-clock split (ms) of this thread's last slslam_lba_solve / slslam_lba_solve_batch:
 * plan | staging + H2D enqueue | launch + device solve + D2H | copy-out | total, then device time by CUDA events:
 * H2D | kernel | D2H.  ms8 receives 8 doubles. */
void slslam_lba_last_timings(double* ms8);

/* ---- device-resident form: plan + upload once, solve any number of times from the uploaded initial guess ---- */
typedef struct slslam_lba_batch slslam_lba_batch;
/* cluster_size 0 = choose automatically; device < 0 = current device. */
int slslam_lba_batch_create(int32_t n, const slslam_lba_desc* descs, const double* const* params,
                            int32_t device, int32_t cluster_size, slslam_lba_batch** out);
/* Enqueue one solve of every window on `cuda_stream` (a cudaStream_t, may be NULL); asynchronous. */
int slslam_lba_batch_solve(slslam_lba_batch* b, void* cuda_stream);
/* Re-upload the initial parameters of every window (H2D only; same sizes as at creation). */
int slslam_lba_batch_upload_params(slslam_lba_batch* b, const double* const* params, void* cuda_stream);
/* Synchronise with `cuda_stream` and copy results back.  Any of the output pointers may be NULL.
 * trace_out[i] receives max_iterations_i * SLSLAM_TRACE_WIDTH doubles. */
int slslam_lba_batch_download(slslam_lba_batch* b, void* cuda_stream, double* const* params_out,
                              slslam_summary* summaries_out, double* const* trace_out);
int slslam_lba_batch_info(const slslam_lba_batch* b, int32_t* cluster_size, int32_t* threads_per_cta,
                          int32_t* smem_bytes_per_cta, int32_t* z_in_smem);
/* How many clusters of this batch's shape the device keeps resident at once (a batch larger than this runs in waves). */
int slslam_lba_batch_max_active_clusters(const slslam_lba_batch* b, int32_t* max_active);
/* Diagnostics: SM cycles CTA 0 of `window` spent per phase in the last solve
 * (init, linearise, pairs, fold, allreduce, gradient, reduced solve, trial, decide, total, then the reduced solve split:
 * prep, factor + panel, trailing update, back-substitution); n <= 14. */
int slslam_lba_batch_phase_cycles(slslam_lba_batch* b, void* cuda_stream, int32_t window, int64_t* cycles_out, int32_t n);
/* Diagnostics of the device planner: SM cycles from kernel start to the end of its phases (counts + constants, scan,
 * grouping by line, partition + tile packing, slot table, slot metadata + gather, pair lists, total) for `window`. */
int slslam_lba_batch_plan_cycles(const slslam_lba_batch* b, int32_t window, int32_t* cycles8);
/* Bytes one host-buffer solve of this batch moves: plan + parameters up, parameters + summaries down. */
int slslam_lba_batch_transfer_bytes(const slslam_lba_batch* b, int64_t* h2d_bytes, int64_t* d2h_bytes);
void slslam_lba_batch_destroy(slslam_lba_batch* b);
/* The launch shape batch creation chooses for `num_windows` windows of at most `max_observations` observations and
 * `max_lines` lines on a device that keeps `resident_ctas` CTAs of the solve kernel resident (148 on a B200) with
 * `smem_bytes_per_cta` of shared memory: CTAs per window and windows per launch (larger batches run in equally full
 * waves).  Pure host arithmetic, usable without a device. */
int slslam_lba_launch_shape(int32_t num_windows, int32_t max_observations, int32_t max_lines, int32_t resident_ctas,
                            int32_t smem_bytes_per_cta, int32_t* ctas_per_window, int32_t* windows_per_wave);
/* Host only (no device needed): which kernel slslam_lba_solve would hand this window to -- SLSLAM_ROUTE_TILED (the tiled
 * solve kernel: <= 32 camera blocks, <= 24 free, <= 32 observations per line), SLSLAM_ROUTE_MOTION_ONLY (one free camera,
 * every line constant: SLAM::motion_only_ba, reference src/slam.cpp:578-675) or SLSLAM_ROUTE_GENERAL (the general kernel:
 * the reference's --ba_window_size 20 / 40 shapes).  Returns the route (>= 0) or SLSLAM_ERR_INVALID / SLSLAM_ERR_UNSUPPORTED
 * (more than max_free_cameras_general free cameras, or a camera observing one line twice in a window of that size). */
#define SLSLAM_ROUTE_TILED 0
#define SLSLAM_ROUTE_MOTION_ONLY 1
#define SLSLAM_ROUTE_GENERAL 2
int slslam_lba_route(const slslam_lba_desc* desc);

/* Test hook: builds the plan of every window twice -- on the device (the product path) and with the host planner --
 * for the same group size and compares every array bit for bit.  0 = identical, 1 = the device planner deferred to the
 * host planner for this batch, 2 = mismatch (detail[0] window, detail[1] field, detail[2] index), < 0 = error. */
int slslam_lba_plan_check(int32_t n, const slslam_lba_desc* descs, const double* const* params, int32_t cluster_size,
                          int32_t* detail);

/* ---- pipelined host-buffer form (throughput over independent batches, BASELINE.json configs[3]) ----
 * submit() validates the arguments and hands the batch to one of `depth` slots, each with its own device pool, pinned
 * staging and CUDA stream; the batch is planned, staged, copied H2D, solved and read back D2H on that slot.  wait()
 * blocks until that batch has finished and only then overwrites the `params_inout` / `summaries_out` given to
 * submit(): they, and the arrays the descs point to, must stay valid and unmodified until then (the desc structs
 * themselves are copied).  When every slot is in flight, submit() first waits for the oldest batch.
 *   flags 0: submit() plans, stages and enqueues on the calling thread and returns without waiting for the device:
 *            the host work and the H2D copy of batch k+1 overlap the kernel of batch k.
 *   SLSLAM_PIPELINE_ASYNC_HOST: every slot also has its own host thread that does the planning / staging / enqueue,
 *            so submit() returns at once, two batches are prepared side by side, and an error found while planning
 *            is reported by wait() (or by the submit() that reuses the slot).
 * One pipeline per caller thread (submit / wait are not re-entrant).  ticket < 0 in wait() drains every slot. */
typedef struct slslam_lba_pipeline slslam_lba_pipeline;
#define SLSLAM_PIPELINE_ASYNC_HOST 2
int slslam_lba_pipeline_create(int32_t device, int32_t depth, int32_t flags, slslam_lba_pipeline** out);
int slslam_lba_pipeline_submit(slslam_lba_pipeline* p, int32_t n, const slslam_lba_desc* descs, double* const* params_inout,
                               slslam_summary* summaries_out, int64_t* ticket_out);
int slslam_lba_pipeline_wait(slslam_lba_pipeline* p, int64_t ticket);
void slslam_lba_pipeline_destroy(slslam_lba_pipeline* p);

/* ---- device-resident map: window assembly and write-back around the LBA solve ----
 * What SLAM::bundle_adjustment does on the host before and after ceres::Solve (reference src/slam.cpp:799-920, 957-972),
 * with the map kept in device memory between keyframes: keyframe poses keyframe_t::T as (R row-major 9 | t 3), landmark
 * lines landmark_t::line as (closest point, direction) in the frame of landmark_t::init_kfid, and every keyframe's
 * observations (landmark id + 8 normalised stereo endpoint coordinates, the obs_vec entries).  Per keyframe the caller
 * uploads the new keyframe with its observations, the new landmarks, and -- after its own metric_embedding
 * (slam.cpp:1317-1366, graph code that stays on the host) -- the re-anchored poses of the <= 2W window keyframes.
 * slslam_map_bundle_adjust then selects the landmarks seen by >= 2 free keyframes (slam.cpp:838-845), builds the index /
 * flag / observation arrays and the parameters (gc_Rt_to_wt, gc_line_from_pose + gc_av_to_orth, src/gc.cpp:24-50, 79-81,
 * 361-417) on the device, solves the window where it was assembled, and writes poses (gc_wt_to_Rt) and lines
 * (gc_orth_to_av + gc_line_to_pose, src/gc.cpp:419-460, 63-77) back into the map.  Differences from the reference's
 * packing that no result depends on: observations are emitted chronologically instead of grouped by landmark (their
 * order inside a line is the same), constant keyframes are numbered by id and kept even when they observe no selected
 * landmark (an unobserved block is never touched). */
typedef struct slslam_map slslam_map;
typedef struct slslam_map_timings {
  double assemble_ms;             /* host wall clock: small upload + selection kernels + read-back of the two sizes */
  double solve_and_writeback_ms;  /* emit + parameters + plan + solve + write-back, until the stream is idle */
  double total_ms;
  int64_t h2d_bytes;              /* bytes uploaded by the call (window keyframe list and ranges) */
  int32_t candidates;             /* observations of the window keyframes that were looked at */
} slslam_map_timings;
int slslam_map_create(int32_t device, int32_t max_keyframes, int32_t max_landmarks, int32_t max_observations, slslam_map** out);
void slslam_map_destroy(slslam_map* m);
int slslam_map_add_keyframe(slslam_map* m, int32_t kf_id, const double* T12, int32_t n_obs, const int32_t* lm_ids, const double* obs8);
int slslam_map_add_landmarks(slslam_map* m, int32_t n, const int32_t* lm_ids, const int32_t* init_kf_ids, const double* line_av6);
int slslam_map_set_poses(slslam_map* m, int32_t n, const int32_t* kf_ids, const double* T12);
int slslam_map_get_poses(slslam_map* m, int32_t n, const int32_t* kf_ids, double* T12_out);
int slslam_map_get_landmarks(slslam_map* m, int32_t n, const int32_t* lm_ids, double* line_av6_out);
/* ba_order[i] = graph-distance rank of ba_kf_ids[i] (ba_kfs, slam.cpp:1376-1382): rank < window_size => free camera.
 * sizes3_out (optional) receives cameras, lines, observations of the assembled window. */
int slslam_map_bundle_adjust(slslam_map* m, int32_t n_ba, const int32_t* ba_kf_ids, const int32_t* ba_order, int32_t window_size,
                             int32_t max_iterations, int32_t robust, slslam_summary* summary_out, int32_t* sizes3_out);
void slslam_map_last_timings(const slslam_map* m, slslam_map_timings* out);
/* Diagnostics / parity: the window the last slslam_map_bundle_adjust assembled (reference array layout; parameters as
 * assembled, before the solve).  Any pointer may be NULL.  line_landmark [L], camera_keyframe [C]. */
int slslam_map_last_window(slslam_map* m, int32_t* camera_index, int32_t* line_index, int32_t* fixed_index, double* observations,
                           double* parameters, int32_t* line_landmark, int32_t* camera_keyframe);
/* Batch geometry conversions on the device: mode 0 gc_av_to_orth (av[6n] -> orth[4n]), 1 gc_orth_to_av, 2 rotation
 * matrix (row-major 9) -> angle-axis (RotationMatrixToAngleAxis behind gc_Rt_to_wt), 3 angle-axis -> rotation matrix. */
int slslam_geometry_convert(int32_t mode, int32_t n, const double* in, double* out);

/* ---- K1 alone: residuals (Huber-unscaled) and analytic Jacobians of every observation, for parity tests ----
 * residuals [4N]; jac_camera [24N] row-major 4x6; jac_line [16N] row-major 4x4; cost_out = 1/2 sum rho. */
int slslam_lba_evaluate(const slslam_lba_desc* desc, const double* params, double* residuals, double* jac_camera,
                        double* jac_line, double* cost_out);

/* ---- what replaces ceres::Solve for one POProblem ---- */
int slslam_po_solve(const slslam_po_desc* desc, double* poses_inout, slslam_summary* summary_out);
int slslam_po_solve_trace(const slslam_po_desc* desc, double* poses_inout, slslam_summary* summary_out,
                          double* trace_out);
/* Device time (CUDA events, ms) of the LM loop of this thread's last slslam_po_solve*, transfers excluded. */
float slslam_po_last_solve_ms(void);
/* How this thread's last slslam_po_solve* factored the normal equations.  The reference selects SPARSE_NORMAL_CHOLESKY
 * (src/po_problem.cpp:68); here the free poses are ordered by minimum degree and the 6x6 blocks of the factor (fill
 * included) are eliminated by one CTA in one launch per LM iteration.  A graph whose factor is more than a third full,
 * or with more than max_column_blocks_sparse blocks in one column, goes to the dense blocked factorisation instead. */
typedef struct slslam_po_stats {
  int32_t sparse;                 /* 2 = block-sparse, level order (warp per column); 1 = block-sparse, minimum-degree order
                                     (column at a time); 0 = dense path */
  int32_t free_poses;
  int64_t factor_blocks;          /* 6x6 blocks of L (dense path: Kf (Kf + 1) / 2) */
  int64_t block_updates;          /* 6x6x6 block products of one numeric factorisation (sparse path) */
  int32_t max_column_rows;
  int32_t iterations_enqueued;    /* LM iterations whose kernels were launched (<= max_iterations: early stop) */
  int64_t factor_cycles[4];       /* last factorisation (sparse paths), SM cycles.  sparse = 1: panel phase, update phase,
                                     back-substitution, total; sparse = 2: stages with a warp per column, stages with the CTA per
                                     column, back-substitution, total */
} slslam_po_stats;
void slslam_po_last_stats(slslam_po_stats* out);
typedef struct slslam_po_limits {
  int32_t max_column_blocks_sparse;   /* off-diagonal blocks in one column of L on the sparse path */
  int32_t max_free_poses_dense;       /* dense fallback: 6 * free poses <= 32 * 8 * 64 */
} slslam_po_limits;
void slslam_po_get_limits(slslam_po_limits* out);
/* Host only (no device needed): the symbolic plan po_solve would build for this graph, with its invariants verified --
 * every column's rows are eliminated later; the columns of a stage (level order) are pairwise non-adjacent and update
 * disjoint blocks and right-hand-side rows; every update destination exists in the structure.  Returns 0 and fills
 * `out` (order: 2 = level order, 1 = minimum-degree order, 0 = not sparse enough: dense path), SLSLAM_ERR_INVALID for a bad
 * graph, SLSLAM_ERR_NUMERICAL when an invariant is violated (a bug). */
typedef struct slslam_po_plan_info {
  int32_t order, free_poses, stages, widest_stage, max_column_rows, reserved;
  int64_t factor_blocks, block_updates;
} slslam_po_plan_info;
int slslam_po_plan_check(const slslam_po_desc* desc, int32_t force_columns, slslam_po_plan_info* out);
/* residuals [6E], jac_pose1 / jac_pose2 [36E] row-major 6x6 */
int slslam_po_evaluate(const slslam_po_desc* desc, const double* poses, double* residuals, double* jac_pose1,
                       double* jac_pose2, double* cost_out);

/* ---- RANSAC hypothesis scoring: the inner loops of SLAM::ransac_motion (reference src/slam.cpp:398-412) ----
 * Every motion hypothesis h (poses[12h..]: R row-major 9, t 3; previous keyframe -> current frame) against every common
 * line k (lines[6k..]: closest point, direction, in the previous keyframe's frame; obs[8k..]: normalised stereo endpoints
 * in the current frame) through SLAM::reprojection_error (src/slam.cpp:691-726, its float / double mix kept).
 * scores[h] = number of lines with error < thr, or -1 when the hypothesis is skipped (|t| > 1, :400-401);
 * inlier [n_hyp][n_lines] (1/0) and errors [n_hyp][n_lines] may be NULL.  thr is parameter.h:56 error_thr (5 / focal). */
int slslam_ransac_score(int32_t n_hyp, const double* poses, int32_t n_lines, const double* lines, const double* obs,
                        double baseline, double thr, int32_t* scores, uint8_t* inlier, float* errors);

#ifdef __cplusplus
}
#endif
#endif /* SLSLAM_B200_H_ */
