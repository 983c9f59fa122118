/* deeparc_ba.h — C ABI of the B200-native bundle-adjustment engine (libdeeparc_ba.so).
 *
 * This is the drop-in boundary for the hot path of pureexe/deeparc-sfm.  The
 * reference has no FFI of its own: the seam is the body of `solve()`
 * (reference src/sfm.cc:31-75), the residual re-evaluation inside
 * `DeepArcManager::filterPoint3d` (src/DeepArcManager.cc:332-352) and the
 * hemisphere block of `main` (src/sfm.cc:86-103), i.e. the slice of the Ceres API
 * the reference calls.  Each entry point below names the reference call it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every call returns DBA_OK (0)
 *     or a negative dba_status and never throws; dba_last_error() gives the text.
 *   - all arithmetic is IEEE fp64; indices are int32; observation counts are int64.
 *   - the caller owns every host buffer (copied during the call); the library owns
 *     device memory, streams and the NCCL communicator.
 *   - a handle is bound to one CUDA device and is not thread-safe (one host thread
 *     per handle).  Multi-GPU = one handle per rank (one process or thread per GPU),
 *     points sharded across ranks, see dba_config.
 *   - there is NO CPU fallback: without a CUDA device dba_create fails with
 *     DBA_ERR_NO_DEVICE.
 */
#ifndef DEEPARC_BA_H_
#define DEEPARC_BA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBA_ABI_VERSION 2

typedef enum dba_status {
  DBA_OK = 0,
  DBA_ERR_INVALID_ARGUMENT = -1, /* null pointer, bad size, index out of range        */
  DBA_ERR_NO_DEVICE = -2,        /* no CUDA device / driver (there is no CPU path)     */
  DBA_ERR_CUDA = -3,             /* CUDA runtime error, text in dba_last_error()        */
  DBA_ERR_UNSUPPORTED = -4,      /* problem shape outside what the engine implements   */
  DBA_ERR_NO_PROBLEM = -5,       /* dba_problem_set has not been called                */
  DBA_ERR_NCCL = -6,             /* NCCL missing or failed (world_size > 1 only)       */
  DBA_ERR_NUMERIC = -7           /* non-finite cost at the initial point               */
} dba_status;

typedef struct dba_handle dba_handle;

/* ------------------------------------------------------------------ configuration */
typedef struct dba_config {
  int32_t device;     /* CUDA device ordinal                                            */
  int32_t rank;       /* 0 .. world_size-1                                              */
  int32_t world_size; /* 1 = single GPU.  >1: every rank passes the FULL problem to
                         dba_problem_set and keeps only its own contiguous point range.
                         The ranks must share one node: the camera-space vector of every
                         PCG iteration is exchanged through CUDA-IPC peer windows over
                         NVLink inside the PCG launch (ncclAllReduce when peer access is
                         not available); the once-per-iteration camera accumulators and
                         scalars go through ncclAllReduce.  dba_problem_set, dba_solve,
                         dba_eval and dba_params_get are collective calls.               */
  const void* nccl_unique_id; /* 128 bytes from dba_nccl_unique_id() on rank 0,
                                 broadcast by the caller; NULL when world_size == 1     */
  int32_t verbose;
} dba_config;

/* ------------------------------------------------------------------------ problem
 * SoA image of the reference's AoS pointer graph.  One observation == one
 * `ParameterBlock` (src/ParameterBlock.hh); the parameter-block order the residual
 * indexes is the one `ParameterBlock::get()` returns (src/ParameterBlock.hh:68-94):
 *   [0] point  [1] principal point  [2] focal  [3] distortion  [4] rot  [5] trans
 *   ([6] ring rot  [7] ring trans when compose_extrinsic(), src/ParameterBlock.hh:24-28)
 * obs_pose_a is the extrinsic behind params[4],[5] (applied LAST), obs_pose_b the one
 * behind params[6],[7] (applied FIRST) or -1:
 *   p = R(rot_a) * (R(rot_b) * X + t_b) + t_a      (snavely_reprojection_error.hh:96-108)
 *   p = R(rot_a) * X + t_a                          (snavely_reprojection_error.hh:110-115)
 */
typedef struct dba_problem {
  int64_t n_obs;
  int32_t n_pts, n_ext, n_intr;
  const double* obs_xy;       /* [n_obs][2] observed pixel (Point2d)                    */
  const int32_t* obs_pt;      /* [n_obs] point index                                    */
  const int32_t* obs_pose_a;  /* [n_obs] extrinsic index of params[4],[5]               */
  const int32_t* obs_pose_b;  /* [n_obs] extrinsic index of params[6],[7], -1 = none;
                                 NULL = no observation composes two poses               */
  const int32_t* obs_intr;    /* [n_obs] intrinsic index                                */
  const double* pts;          /* [n_pts][3]  Point3d::position()                        */
  const double* ext_rot;      /* [n_ext][3]  Extrinsic::rotation(), angle-axis          */
  const double* ext_trans;    /* [n_ext][3]  Extrinsic::translation()                   */
  const double* intr_center;  /* [n_intr][2] Intrinsic::center()                        */
  const double* intr_focal;   /* [n_intr][2] Intrinsic::focal() (second unused if nf=1) */
  const double* intr_dist;    /* [n_intr][2] Intrinsic::distrotion()                    */
  const int32_t* intr_nf;     /* [n_intr] focal_size() in {1,2}                         */
  const int32_t* intr_nd;     /* [n_intr] distrotion_size() in {0,1,2}                  */
  const uint8_t* ext_const;   /* [n_ext] 1 = SetParameterBlockConstant (gauge rule,
                                 sfm.cc:50-53); NULL = none                             */
  int32_t freeze_camera;      /* sfm.cc:54-57: everything but the points is constant    */
  int32_t free_intrinsics;    /* 0 = as shipped, sfm.cc:60-62 (centre, focal, distortion
                                 constant).  1 = focal and distortion are optimised with
                                 the pose as one 9-dof camera block [w,t,f,k0,k1]; needs
                                 nf=1, nd=2, obs_intr == obs_pose_a, no composed poses   */
} dba_problem;

/* -------------------------------------------------------------------------- solve
 * Mirrors ceres::Solver::Options as the reference leaves them (sfm.cc:66-71 sets
 * linear_solver_type, minimizer_progress_to_stdout, max_num_iterations, num_threads,
 * max_solver_time_in_seconds; everything else is the Ceres default).               */
typedef enum dba_linear_solver {
  DBA_LS_AUTO = 0,  /* DBA_LS_DENSE when the reduced system has at most dense_max_size
                       unknowns (and fits the dense path), else DBA_LS_PCG              */
  DBA_LS_PCG = 1,   /* implicit Schur complement + block-Jacobi preconditioned CG       */
  DBA_LS_DENSE = 2  /* the reference's DENSE_SCHUR (sfm.cc:67, :95): explicit reduced camera
                       system assembled on the device + dense Cholesky, an exact step.
                       Limits: at most 128 camera blocks and 1008 reduced unknowns (the
                       reference's rigs have 19..50 blocks); beyond them dba_solve returns
                       DBA_ERR_UNSUPPORTED — it never substitutes the PCG silently.      */
} dba_linear_solver;

typedef enum dba_loss {
  DBA_LOSS_NONE = 0,   /* cost = 1/2 sum |r|^2 (as shipped)                                    */
  DBA_LOSS_CAUCHY = 1  /* cost = 1/2 sum rho(|r|^2); residuals and Jacobians of every observation
                          are rescaled by sqrt(rho') inside the Jacobian kernel (Ceres' corrector;
                          rho'' < 0 for Cauchy, so the rank-one term vanishes).  The implicit Schur
                          product then runs on the materialised planes.                          */
} dba_loss;

typedef struct dba_solve_options {
  int32_t max_num_iterations;             /* reference: 100 (sfm.cc:111,121)            */
  double max_solver_time_in_seconds;      /* reference: 3600                            */
  double initial_trust_region_radius;     /* 1e4                                        */
  double max_trust_region_radius;         /* 1e16                                       */
  double min_trust_region_radius;         /* 1e-32                                      */
  double min_relative_decrease;           /* 1e-3                                       */
  double min_lm_diagonal;                 /* 1e-6                                       */
  double max_lm_diagonal;                 /* 1e32                                       */
  double function_tolerance;              /* 1e-6                                       */
  double gradient_tolerance;              /* 1e-10                                      */
  double parameter_tolerance;             /* 1e-8                                       */
  int32_t jacobi_scaling;                 /* 1                                          */
  int32_t max_num_consecutive_invalid_steps; /* 5                                       */
  int32_t linear_solver;                  /* dba_linear_solver                          */
  int32_t pcg_max_iterations;             /* 500                                        */
  int32_t pcg_min_iterations;             /* 0                                          */
  double pcg_rel_tolerance;               /* stop when r'z <= tol^2 * r0'z0; 1e-12      */
  int32_t dense_max_size;                 /* DBA_LS_AUTO picks DBA_LS_DENSE up to this many
                                             reduced unknowns (default 768)             */
  int32_t progress_to_stdout;             /* reference: true (sfm.cc:68)                */
  int32_t loss_type;                      /* dba_loss.  The reference passes NULL (sfm.cc:48) and
                                             keeps `new ceres::CauchyLoss(0.5)` in a comment (:49) */
  double loss_scale;                      /* a of CauchyLoss(a): rho(s) = a^2 log(1 + s / a^2)  */
} dba_solve_options;

typedef enum dba_termination {
  DBA_CONVERGENCE = 0,
  DBA_NO_CONVERGENCE = 1,
  DBA_FAILURE = 2
} dba_termination;

typedef struct dba_iteration {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_successful;
  int32_t linear_solver_iterations;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  double model_cost_change;
  double iteration_time_in_seconds;
} dba_iteration;

typedef struct dba_summary {
  int32_t termination; /* dba_termination */
  int32_t num_iterations; /* entries written to `iterations` (iteration 0 included)     */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t linear_solver_used; /* dba_linear_solver actually run                         */
  int32_t reduced_system_size;
  double initial_cost;
  double final_cost;
  double total_time_in_seconds;      /* host wall clock around the LM loop              */
  double device_time_in_seconds;     /* CUDA events around the whole call               */
  double loop_device_time_in_seconds;/* CUDA events around iterations >= 1 only (the
                                        initial evaluation, iteration 0, excluded)      */
  int64_t kernel_launches;           /* kernels of this library launched by the call    */
  int64_t jacobian_evaluations;
  int64_t residual_evaluations;
  int64_t pcg_iterations_total;
  char message[192];
  dba_iteration* iterations;  /* caller-provided array, may be NULL                     */
  int32_t iterations_capacity;
  int32_t linear_solver_failures;  /* DBA_LS_DENSE: factorisations that met a non-positive
                                      pivot (the step is then invalid, as in Ceres)          */
  int32_t pcg_unconverged_solves;  /* DBA_LS_PCG with pcg_rel_tolerance > 0: solves that hit
                                      pcg_max_iterations before reaching the tolerance       */
  int32_t reserved_;
} dba_summary;

/* Per-kernel accounting (CUDA events on the launching stream), for bench.py. */
typedef struct dba_kernel_stat {
  char name[48];
  int64_t launches;
  double total_ms;
  double algorithmic_bytes; /* per launch, SURVEY.md §8(d) model; 0 if not modelled     */
} dba_kernel_stat;

/* ---------------------------------------------------------------------- functions */
int dba_abi_version(void);

/* Number of visible CUDA devices, or a negative dba_status. */
int dba_device_count(void);

/* Rank 0 fills 128 bytes that every rank then passes in dba_config.nccl_unique_id. */
int dba_nccl_unique_id(void* out128);

/* Host-only helper (no GPU needed): the point range each rank of a world_size-rank job owns,
 * contiguous in point index and balanced by observation count.  pt_begin has world_size + 1
 * entries (rank r owns points [pt_begin[r], pt_begin[r+1])), obs_count world_size entries.
 * dba_problem_set uses exactly this plan.                                                  */
int dba_shard_plan(const dba_problem* p, int32_t world_size, int32_t* pt_begin, int64_t* obs_count);

int dba_create(dba_handle** out, const dba_config* cfg);
void dba_destroy(dba_handle* h);
const char* dba_last_error(const dba_handle* h); /* h may be NULL: last create error     */

/* Replaces the Problem construction loop of solve() (sfm.cc:36-65): copies the SoA
 * arrays to the device, validates indices, sorts observations by point, builds the
 * shard of this rank.  May be called again with a new problem on the same handle.
 * Input that already is sorted by point and does not compose poses (obs_pose_b all -1) is
 * staged shard by shard and indexed on the device; anything else is sorted and indexed by
 * the host cores first.  Same structures, same results either way.
 * Limits (DBA_ERR_UNSUPPORTED beyond them; the reference has none): a point may carry at
 * most 1024 observations (512 when observations compose two poses and cameras are free) —
 * a point and its observations are processed by one thread block; fewer than 2^30
 * observations per handle; free_intrinsics as described at dba_problem.                 */
int dba_problem_set(dba_handle* h, const dba_problem* p);

/* The outer loop of main() (sfm.cc:118-127: solve, filterPoint3d, solve, ...) without a second
 * upload: removes the observations / points flagged by dba_filter (arrays in the shapes and the
 * caller's order of the last dba_problem_set / dba_problem_update; either may be NULL) and rebuilds
 * the device structures from the point-sorted image the engine kept.  Observations of removed points
 * go with them; surviving points / observations keep their relative order and are re-indexed
 * 0 .. n-1, exactly as a fresh dba_problem_set of the filtered scene would index them.  Parameters
 * continue from their current device values.  freeze_camera may change (sfm.cc:111 vs :121).
 * Single-GPU handles only.                                                                  */
int dba_problem_update(dba_handle* h, const uint8_t* obs_remove, const uint8_t* pt_remove, int32_t freeze_camera,
                       int64_t* n_obs_out, int32_t* n_pts_out);

/* Restores the parameters uploaded by the last dba_problem_set (device-side copy). */
int dba_params_reset(dba_handle* h);

/* Replaces one Evaluate() of the Ceres problem / the plain-double functor call of
 * filterPoint3d (DeepArcManager.cc:335-347).  Outputs are in the CALLER's observation
 * order; any pointer may be NULL.
 *   cost       1/2 sum r^2 (all ranks' observations when world_size > 1)
 *   residuals  [n_obs][2]
 *   jac_pt     [n_obs][2][3]  d r / d point
 *   jac_pose_a [n_obs][2][6]  d r / d (rot_a, trans_a)
 *   jac_pose_b [n_obs][2][6]  d r / d (rot_b, trans_b); zero rows where pose_b == -1
 *   jac_intr   [n_obs][2][3]  d r / d (f, k0, k1).  nf = 2: column 0 holds the diagonal of the
 *                             2x2 focal Jacobian, (d r0 / d fx, d r1 / d fy); the off-diagonal
 *                             entries d r0 / d fy and d r1 / d fx are identically zero
 * Jacobians are UNSCALED and ignore constancy masks (raw derivatives).  Single-GPU
 * handles only for the per-observation outputs.                                       */
int dba_eval(dba_handle* h, double* cost, double* residuals, double* jac_pt, double* jac_pose_a,
             double* jac_pose_b, double* jac_intr);

void dba_solve_options_default(dba_solve_options* o);

/* Replaces ceres::Solve(options, &problem, &summary) (sfm.cc:73): GPU-resident
 * Levenberg-Marquardt.  Parameters stay on the device; fetch with dba_params_get.    */
int dba_solve(dba_handle* h, const dba_solve_options* o, dba_summary* s);

/* Scatter-back source: Ceres writes results in place through the double* it was
 * given (sfm.cc:47-48); the C++ wrapper copies these arrays back into
 * Point3d::position(), Extrinsic::rotation()/translation(), Intrinsic::focal()/
 * distrotion().  Any pointer may be NULL.  Arrays have the dba_problem shapes.        */
int dba_params_get(dba_handle* h, double* pts, double* ext_rot, double* ext_trans,
                   double* intr_focal, double* intr_dist);

/* Replaces the hemisphere problem of main() (sfm.cc:86-103): residual_k =
 * |c - centre_k|^2 - rho (hemisphere_radius.hh:18-28; rho is the SQUARED radius),
 * 4 unknowns, same LM options.  centre_io / rho_io hold the initial values
 * (reference: 0,0,0 and 1; sfm.cc:87-88) and receive the result.                     */
int dba_fit_hemisphere(dba_handle* h, const double* centres /* [n][3] */, int32_t n,
                       double centre_io[3], double* rho_io, const dba_solve_options* o,
                       dba_summary* s);

/* Replaces the per-observation loop of filterPoint3d (DeepArcManager.cc:332-352):
 * mse[i] = (r0^2 + r1^2) / 2 at the current device parameters, caller's order.      */
int dba_filter_mse(dba_handle* h, double* mse /* [n_obs] */);

/* The decisions of filterPoint3d (DeepArcManager.cc:331-424) taken on the device at the
 * current parameters, so that only one byte per observation / point crosses the bus:
 *  (1) observations with mse = (r0^2 + r1^2) / 2 < error_boundary go (:347-350; "<" as the
 *      reference writes it),
 *  (2) points left without observations go (:368-378),
 *  (3) points with |x - centre|^2 > rho / 2 go together with all their observations
 *      (:380-408; rho is the SQUARED radius fitted by dba_fit_hemisphere); centre == NULL
 *      skips this rule.
 * obs_remove[n_obs] / pt_remove[n_pts]: 1 = removed, caller's order; the counts are
 * optional.  The pointer-graph surgery itself stays with the caller.                  */
int dba_filter(dba_handle* h, double error_boundary, const double* centre /* [3] or NULL */,
               double rho, uint8_t* obs_remove /* [n_obs] */, uint8_t* pt_remove /* [n_pts] */,
               int64_t* n_obs_removed, int32_t* n_pts_removed);

/* Kernel statistics since the last dba_kernel_stats_reset; returns the number of
 * entries written (<= capacity).  Timing is only collected when enabled.            */
int dba_kernel_stats_enable(dba_handle* h, int32_t enable);
int dba_kernel_stats_reset(dba_handle* h);
int dba_kernel_stats(dba_handle* h, dba_kernel_stat* out, int32_t capacity);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* DEEPARC_BA_H_ */
