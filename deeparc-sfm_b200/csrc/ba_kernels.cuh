// Launch interface of the bundle-adjustment kernels (sm_100a).  See DESIGN.md for the data
// layout and the roofline of each kernel.
//
// Observation-ordered data is POINT-SORTED; every per-observation quantity is a "plane" of
// double2 (= the two residual rows of one Jacobian column) with plane stride `ld`, so each
// thread issues 16-byte loads and a warp touches 512 contiguous bytes per plane:
//   plane 0            r            (r0, r1)
//   plane 1..3         Jp[:,k]      d r / d point_k         (Jacobi-scaled)
//   plane 4..4+CB-1    JA[:,k]      d r / d block A (pose a [+ f,k0,k1 when CB = 9])
//   plane 4+CB..+5     JB[:,k]      d r / d block B (pose b), TWO only
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_math.cuh"

// Checked build (`make CHECKED=1` -> lib/libdeeparc_ba_checked.so): device-side assertions on every
// data-dependent index of the tile / product / gather / dense kernels.  compute-sanitizer is closed on the
// GPU pool this was developed on (profiles/r02_sanitizer.md); these checks are the substitute, run over
// scripts/sanitize_probe.py.  A failed check prints the site and traps (the launch fails with an error).
#ifdef DBA_CHECKED
#include <cstdio>
#define DBA_CHECK(cond)                                                                        \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      printf("DBA_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, \
             static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));                     \
      __trap();                                                                                \
    }                                                                                          \
  } while (0)
#else
#define DBA_CHECK(cond) \
  do {                  \
  } while (0)
#endif

namespace dba {

constexpr int kTile = 256;       // default tile capacity (observations per tile == threads per tile CTA);
                                 // 512 / 1024 are selected when a point has a longer track
constexpr int kMaxTile = 1024;
// whole points per tile of capacity `tile` (bounds the staged per-point data of k_spmv_mf)
__host__ __device__ constexpr int max_tile_points(int tile) { return tile >= 1024 ? 256 : 128; }
constexpr int kPlaneR = 0;
constexpr int kPlaneJp = 1;
constexpr int kPlaneJA = 4;

// One record per point tile: everything a tile CTA needs to find its data (32 B, one load).
struct alignas(16) TileMeta {
  int obs0, n_obs;    // observations [obs0, obs0 + n_obs)
  int pt0, n_pts;     // whole points [pt0, pt0 + n_pts)
  int g0, n_parts;    // partials (tile, camera block) [g0, g0 + n_parts)
  int item0, n_items; // incidence items of the tile, grouped by partial
};

// Everything a kernel needs to find its data.  Plain pointers, passed by value.
struct DeviceProblem {
  int64_t n_obs;   // local (this rank) observations
  int64_t ld;      // plane stride (double2 elements)
  int n_pts;       // local points
  int n_ext, n_intr, n_tiles;
  int tile;        // tile capacity actually used: 256, 512 or 1024
  int cb;          // camera block size: 0 (points only), 6, 9
  int two;         // 1: observations may carry a second pose block
  int n_blocks;    // camera blocks (== n_ext)
  const double2* obs_xy;    // [n_obs]
  const int2* obs_ip;       // [n_obs] (intrinsic, local point)
  const int* tile_obs;      // [n_tiles + 1]
  const int* tile_pt;       // [n_tiles + 1]
  const int* pt_first;      // [n_pts + 1] first observation of each local point
  // camera-sorted incidence (entries = obs * 2 + slot), chunked for the gather kernel
  const int* cam_entries;   // [n_entries]
  const int4* cam_chunks;   // [n_chunks] (block, first entry, last entry, 0)
  const int* cam_chunk_first;  // [n_blocks + 1] first chunk of each camera block
  int n_chunks;
  // static tile-local camera incidence for the implicit Schur product: one "partial" per
  // (tile, camera block present in the tile)
  const TileMeta* tile_meta;        // [n_tiles]
  const int2* obs_ab;               // [n_obs] (block a, block b or -1) of each observation
  const unsigned short* obs_lp;     // [n_obs] point index local to the tile
  const unsigned short* part_first_rel;  // per tile t, n_parts + 1 item offsets relative to item0, at [g0 + t]
  const unsigned short* items;      // [n_entries] local observation | slot << 15, grouped by partial
  const int* cam_part_first;        // [n_blocks + 1] camera block -> its partials
  int n_partials;
  double2* J;               // planes
  // matrix-free implicit Schur product (k_spmv_mf): observations of a tile re-ordered by camera
  // block ("columns"), so a warp touches few distinct camera rows and the reduce-by-camera runs
  // over contiguous columns
  const void* mf_cols;              // [n_tiles][tile] (padded, block a = -1) per column: CB = 9: int2 (block a, lp << 16 | lo);
                                    //   else int4 (block a, block b or -1, lp << 16 | lo, intrinsic);
                                    //   lo = point-sorted position in the tile, lp = tile-local point
  const int* items_mf;              // two-pose problems: `items` with lo replaced by the column
  const int* part_first;            // `part_first_rel` widened to int (cp.async granularity), same indexing
  const int* part_dst;              // [n_partials] row of partial g in the camera-grouped partial buffer
  const int* part_blk;              // [n_partials] camera block of partial g (rows staged by k_jacobian_tile)
  const ushort2* obs_lc;            // [n_obs] tile-local camera block (= partial) of block a / block b (0xffff: none)
  int intr_is_pose;                 // every observation uses intrinsic == block a (per-camera intrinsics)
  // robust loss of the running solve (dba_solve_options.loss_type): 0 none, 1 Cauchy with b = a^2, c = 1 / b
  int loss_type;
  double loss_b, loss_c;
};

__host__ __device__ constexpr int mf_row_len(int cb) { return cb == 9 ? 24 : 20; }  // multiples of 4 doubles (256-bit loads)

struct ParamSet {
  double* pts;        // [n_pts_local][3]
  double* ext_rot;    // [n_ext][3]
  double* ext_trans;  // [n_ext][3]
  double* focal;      // [n_intr][2]
  double* dist;       // [n_intr][2]
  const double* center;  // [n_intr][2] (never optimised)
  const int* nf;      // [n_intr]
  const int* nd;      // [n_intr]
  PoseRow* pose_rows; // [n_ext]   derived
  IntrRow* intr_rows; // [n_intr]  derived
};

// per-point and per-camera work arrays
struct WorkArrays {
  double* sp;      // [n_pts][3] Jacobi scale of point columns
  double* sc;      // [n_blocks][cb] Jacobi scale of camera columns
  double* cinv;    // [n_pts][6]  (E^T E + D^2)^-1
  double* tp;      // [n_pts][4]  C^-1 g_p (padded rows: one 32-byte load)
  double* dp;      // [n_pts][3]  point step (scaled space)
  // camera accumulators, ONE contiguous buffer (single allreduce):
  //   B [n_blocks][cb][cb] | diagF [n_blocks][cb] | gc [n_blocks][cb] | rhs [n_blocks][cb]
  double* cam_acc;
  double* cam_chunk_acc;  // [n_chunks][cb (cb + 1) / 2 + 3 cb] per-chunk sums of k_camera_gather, combined in chunk order
  double* minv;    // [n_blocks][cb][cb] inverse of the block-Jacobi preconditioner
  double* dc2;     // [n_blocks][cb] D_c^2
  // PCG vectors [n_blocks * cb]
  double *x, *r, *z, *p, *q;
  double* scalars;   // [32] reduced scalars, copied to the host
  int* pcg_state;    // [4] iter, done, peer exchange timed out, -
  double* pcg_scal;  // [4] rz, rz0, p.q, beta
  double* q_split;     // [n_split][n_blocks * cb] slices of q when the per-camera sum is split
  double* partials_q;  // [n_partials][cb] tile-local partial products of the implicit Schur product
  // matrix-free product: one row per camera block, rebuilt after every Jacobian evaluation
  //   CB = 6: R[9] t[3] p~[6] 0 0      CB = 9: R[9] t[3] f k0 k1 p~[9]
  //   (the last mantissa bit of t_x is the branch selector: 0 in Ceres' small-angle branch of
  //   AngleAxisRotatePoint, else 1)
  // p~ = T p is the PCG direction in "geometric" coordinates, T = blockdiag(J_l(w), I, I) diag(sc * free)
  // (J_l = left Jacobian of SO(3): d(R X) = (J_l dw) x (R X)); rewritten every PCG iteration
  double* mf_rows;   // [n_blocks][mf_row_len(cb)]
  double* mf_T;      // [n_blocks][9 + cb]  J_l row-major, then sc * free
  double* vec_partials;    // per-CTA partials of the PCG vector kernels
  unsigned int* counters;  // [4] "last block" arrival counters
  // DBA_TAIL_TRACE=1: device-timestamp accumulators of the fused PCG launch (NULL otherwise), ns:
  //   [0] launches  [1] product (kernel start -> CTA 0 done)  [2] grid sync 1 (incl. waiting for the slowest CTA)
  //   [3] phase 1 (per-camera sums + push)  [4] exchange wait + slot sum  [5] grid sync 2  [6] phase 2
  //   [7] grid sync 3  [8] phase 3  [9] spread of the CTAs' product-done times  [10] min  [11] max (scratch)
  unsigned long long* trace;
};

// Peer windows of the fused PCG tail (multi-GPU): every rank owns one window in its HBM,
//   ll [2 buffers][world slots][slot_len] 16-byte records   slot s = camera-space vector pushed by rank s
// mapped into the other ranks through CUDA IPC, so a rank PUSHES its partial q = S p straight into
// the HBM of every peer over NVLink.  A record is {low half, seq, high half, seq} of one double: each
// 8-byte half validates itself (NCCL's LL protocol), so the receiver simply polls the record it
// needs — no fence, no flag round trip, no grid-wide barrier around the exchange — and adds the
// slots in rank order (bit-identical on every rank, no NCCL call on the PCG path).
constexpr int kMaxPeers = 8;
struct PeerWin {
  int world = 1, rank = 0;
  unsigned long long seq = 0;   // number of this exchange (1, 2, ...), identical on every rank
  long long slot_len = 0;       // records per slot
  long long timeout_ns = 0;     // give up waiting for a peer after this long (error flag, no hang)
  uint4* ll[kMaxPeers] = {};
};

// Explicit reduced system + Cholesky (ba_dense.cu): DENSE_SCHUR for small camera counts.
constexpr int kDnPtsCap = 64;      // points per batch of k_schur_dense
constexpr int kDnEntCap = 384;     // (point, camera block) entries per batch
constexpr int kDnMaxBlocks = 128;  // camera blocks (the per-point block lookup is a 16-bit table in shared memory)
constexpr int kDnMaxSize = 1008;   // reduced unknowns (one CTA of 1024 threads owns the rows of the factorisation)
struct DenseWork {
  const int* batch_pt = nullptr;  // [n_batches + 1] first local point of each batch (whole points; sum of
                                  // min(2 k_i, n_blocks) over a batch <= kDnEntCap, points <= kDnPtsCap)
  int n_batches = 0;
  int n_pairs = 0;                // n_blocks (n_blocks + 1) / 2 camera-block pairs A <= B
  double* Z = nullptr;            // [n_batches][kDnEntCap][3 cb] Z entries of each batch, compacted (k_dense_z -> k_dense_pairs)
  unsigned int* Zent = nullptr;   // [n_batches][kDnEntCap] entry -> (point in batch << 16 | block)
  int* Zcount = nullptr;          // [n_batches] entries of each batch
  double* S_part = nullptr;       // [slices][n_pairs][cb * cb]
  double* S = nullptr;            // [n][n] reduced matrix without D_c^2, then its Cholesky factor (lower)
  int* fail_flag = nullptr;       // set to 1 when a pivot is not positive
  // two-pose problems: observations that compose two pose blocks, sorted by the block pair
  const int* pair_entries = nullptr;      // [n] observation * 2 + (1: block b is the lower-numbered block)
  const int4* pair_chunks = nullptr;      // [n_pair_chunks] (pair, first entry, last entry, 0), <= 1024 entries each
  const int* pair_chunk_first = nullptr;  // [n_pairs + 1] pair -> its chunks (NULL: no composed observations)
  double* pair_chunk_acc = nullptr;       // [n_pair_chunks][36]
  int n_pair_chunks = 0;
};
int dense_slices(const DenseWork& Q);
// S (both triangles, without D_c^2) from the Jacobian planes, W.cinv and the F^T F blocks that
// k_camera_gather (mode 2) left in W.cam_acc (added when add_diag != 0); returns 0 on success
int launch_schur_dense(const DeviceProblem& D, const WorkArrays& W, const DenseWork& Q, int add_diag, cudaStream_t st);
// W.x = (S + D_c^2)^-1 rhs  (rhs = the fourth region of W.cam_acc); NaN and *fail_flag = 1 when not positive definite
int launch_dense_cholesky(const DeviceProblem& D, const WorkArrays& W, const DenseWork& Q, cudaStream_t st);

// ---- launches (all asynchronous on `st`) ---------------------------------------------
void launch_pose_rows(const ParamSet& P, const uint8_t* ext_const, int freeze_all, int n_ext, int n_intr,
                      cudaStream_t st);
// residuals + Jacobians at P.  cb_store in {0,6,9}; store_two: also block B planes.
// unit_scale: ignore sp/sc and constancy masks (raw derivatives for dba_eval / column norms).
// partial_cost: per-CTA sums of r^2, cost_grid(D) entries.
void launch_jacobian(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, int cb_store, int store_two,
                     int unit_scale, double* partial_cost, cudaStream_t st);
// residual-only cost at P: per-CTA partial sums of r^2; optional per-observation mse.
void launch_cost(const DeviceProblem& D, const ParamSet& P, double* partial_cost, double* mse_out, cudaStream_t st);
// filterPoint3d decisions from the per-observation mse (sorted order): removal flags per observation / point
void launch_filter_flags(const DeviceProblem& D, const ParamSet& P, const double* mse, double boundary, const double* centre,
                         double rho, uint8_t* obs_remove, uint8_t* pt_remove, cudaStream_t st);
int cost_grid(const DeviceProblem& D);
int jacobian_partials(const DeviceProblem& D);  // entries launch_jacobian writes into partial_cost
int tile_grid(const DeviceProblem& D);
// per point: H = E^T E, g = E^T r.  mode 0: Jacobi scales sp.  mode 1: C = H + D^2, C^-1,
// t = C^-1 g, partials[3*tile + {0,1,2}] = {sum g^2, max |g|, #non-SPD blocks}.
void launch_point_prepare(const DeviceProblem& D, const WorkArrays& W, double radius, double min_diag,
                          double max_diag, int mode, double* partials, cudaStream_t st);
// camera-sorted gather into W.cam_acc (zeroed here).  mode 0: diag F^T F only; mode 1: everything;
// mode 2: as mode 1 with B = F^T F (no elimination term): the diagonal blocks of the dense reduced system.
void launch_camera_gather(const DeviceProblem& D, const WorkArrays& W, int mode, cudaStream_t st, const ParamSet* recompute = nullptr);
void launch_camera_scales(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st);
// D_c^2, block-Jacobi inverse; partials[3*cta + {0,1,2}] as above for the camera side
int camera_finalize_grid(const DeviceProblem& D);
// with_minv = 0 (dense reduced system): D_c^2 and the gradient norms only
void launch_camera_finalize(const DeviceProblem& D, const WorkArrays& W, double radius, double min_diag,
                            double max_diag, double* partials, int with_minv, cudaStream_t st);
void launch_pcg_init(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st);
// implicit Schur complement product: one pass over point tiles -> W.partials_q, then the
// per-camera fixed-order sum -> W.q
void launch_spmv_tile(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st);
void launch_partials_to_q(const DeviceProblem& D, const WorkArrays& W, int fuse_dot, int n_split, int mf, cudaStream_t st);
// matrix-free variant: camera rows (static part) after a Jacobian evaluation; p (+)= ..., p~ = T p; the product
void launch_mf_rows(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, cudaStream_t st);
void launch_mf_direction(const DeviceProblem& D, const WorkArrays& W, int init, cudaStream_t st);
// tail.fuse: cooperative launch whose epilogue is the rest of the PCG iteration (pcg_tail: per-camera
// partial sum, peer exchange, vector updates).  Returns 0 on success, -1 when the launch failed.
struct MfTail {
  int fuse = 0;  // 0: product only; n >= 1: run the tail with n slices per camera block
  int min_iter = 0;
  double tol2 = 0.0;
  PeerWin pw{};
};
int launch_spmv_mf(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, const MfTail& tail, cudaStream_t st);
// PCG vector phases: q += D_c^2 p and p.q; the x/r/z update with r.z; the new direction
void launch_fold_q(const DeviceProblem& D, const WorkArrays& W, int n_split, cudaStream_t st);
void launch_pcg_dot(const DeviceProblem& D, const WorkArrays& W, int n_split, cudaStream_t st);
void launch_pcg_step(const DeviceProblem& D, const WorkArrays& W, double tol2, int min_iter, cudaStream_t st);
void launch_pcg_direction(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st);
// single GPU: partial sum + all three vector phases in one cooperative launch; returns 0 on success
// pw.world > 1: the partial q of every rank is exchanged through the peer windows inside the launch
// n_split > 1: few camera blocks with long partial lists — n_split warps per block (W.q_split holds the slices)
int launch_pcg_fused(const DeviceProblem& D, const WorkArrays& W, int mf, double tol2, int min_iter, int n_split,
                     const PeerWin& pw, cudaStream_t st);
// dp = -t - C^-1 E^T F x ; partial_model[tile] = sum (J d).(r + J d / 2)
void launch_back_substitute(const DeviceProblem& D, const WorkArrays& W, double* partial_model, cudaStream_t st,
                            const ParamSet* recompute = nullptr);
// candidate = current + scale * step; partials[2*cta + {0,1}] = {sum step^2, sum x^2}
int update_points_grid(const DeviceProblem& D);
int update_cameras_grid(const DeviceProblem& D);
void launch_update_points(const DeviceProblem& D, const ParamSet& cur, const ParamSet& cand, const WorkArrays& W,
                          double* partials, cudaStream_t st);
void launch_update_cameras(const DeviceProblem& D, const ParamSet& cur, const ParamSet& cand, const WorkArrays& W,
                           double* partials, cudaStream_t st);
// deterministic single-CTA reductions of strided partials
void launch_reduce_sum(const double* partials, int n, int stride, int offset, double* out, cudaStream_t st);
void launch_reduce_max(const double* partials, int n, int stride, int offset, double* out, cudaStream_t st);
// up to 8 of them in one launch (one CTA per job)
struct ReduceJob {
  const double* in;
  double* out;
  int n, stride, offset, is_max;
};
struct ReduceJobs {
  ReduceJob job[8];
  int count = 0;
  void sum(const double* in, int n, int stride, int offset, double* out) { job[count++] = ReduceJob{in, out, n, stride, offset, 0}; }
  void max(const double* in, int n, int stride, int offset, double* out) { job[count++] = ReduceJob{in, out, n, stride, offset, 1}; }
};
void launch_reduce_multi(const ReduceJobs& jobs, cudaStream_t st);

}  // namespace dba
