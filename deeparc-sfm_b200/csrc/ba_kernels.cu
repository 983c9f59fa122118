// Hand-written sm_100a kernels of the bundle-adjustment hot path.  fp64 throughout; the path
// is streaming + segmented reductions (no dense contraction, so no tensor cores).
// Kernel <-> reference map (DESIGN.md has the byte models and the measured bounds):
//   k_pose_rows        per-extrinsic trig hoisting for rotatePoint (snavely_reprojection_error.hh:80-91)
//   k_jacobian_tile / k_jacobian (K1)  residual + analytic Jacobian of operator() (:93-118), replaces
//                      Ceres autodiff; per point tile with the tile's camera rows staged in shared memory
//   k_cost       (K2)  residual-only evaluation (rejected-step streaks; filterPoint3d, DeepArcManager.cc:335-347)
//   k_filter_flags     the three removal rules of filterPoint3d decided per point (DeepArcManager.cc:347-408)
//   k_point_prepare, k_camera_gather + k_camera_combine, k_camera_finalize (K3)  Schur elimination
//                      front half: C^-1, t per point; block-Jacobi blocks, gradient, rhs per camera (no atomics)
//   k_spmv_mf    (K5)  matrix-free implicit Schur complement product, persistent CTAs, cp.async tile
//                      staging; launched cooperatively with pcg_tail (per-camera sum of the partials,
//                      cross-rank exchange through NVLink peer windows, PCG vector updates) as its epilogue
//   k_spmv_tile + k_partials_to_q  the same product from the materialised planes (TMA-staged tiles)
//   k_pcg_fused        pcg_tail as its own cooperative launch
//   k_pcg_init / k_pcg_dot / k_pcg_step / k_pcg_direction / k_mf_direction / k_fold_q  plain-launch PCG
//                      vector work (DBA_PCG_FUSED=0, few-camera multi-GPU without peer windows)
//   k_back_substitute (K7), k_update_points / k_update_cameras, k_reduce_multi  point back-substitution,
//                      x + delta, norms, fixed-order reductions
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include <cooperative_groups.h>

#include "ba_kernels.cuh"

namespace dba {

namespace {

// butterfly: every lane ends with the same total (fixed order)
__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum over the CTA; result valid in thread 0.  `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum(v);
  return v;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_max(v);
  return v;
}
// Sum over the CTA, result broadcast to every thread.
__device__ __forceinline__ double block_sum_all(double v, double* red) {
  v = block_sum(v, red);
  __syncthreads();
  if (threadIdx.x == 0) red[0] = v;
  __syncthreads();
  v = red[0];
  __syncthreads();
  return v;
}

// "last block" pattern: deterministic grid-wide scalars without a cooperative launch
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  return is_last;
}
__device__ __forceinline__ double sum_partials(const volatile double* part, int n, double* red) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  return block_sum_all(acc, red);
}

__device__ __forceinline__ double dot2(const double2 a, const double2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// camera row -> registers with 256-bit loads (sm_100 LDG.256: one L1 tag lookup per 32 B of a row
// instead of one per 16 B; rows are 32-byte aligned and a multiple of 32 bytes long)
template <int N>
__device__ __forceinline__ void load_row(const double* __restrict__ row, double (&r)[N]) {
  static_assert(N % 4 == 0, "row length must be a multiple of 4 doubles");
#pragma unroll
  for (int i = 0; i < N / 4; ++i)
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r[4 * i]), "=d"(r[4 * i + 1]), "=d"(r[4 * i + 2]), "=d"(r[4 * i + 3])
                 : "l"(row + 4 * i));
}


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// cp.async (LDGSTS) helpers: per-thread 4/8/16-byte asynchronous copies global -> shared
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------- pose rows
__global__ void k_pose_rows(ParamSet P, const uint8_t* __restrict__ ext_const, int freeze_all, int n_ext, int n_intr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_ext) {
    const double free_flag = (freeze_all || (ext_const && ext_const[i])) ? 0.0 : 1.0;
    make_pose_row(P.ext_rot + 3 * i, P.ext_trans + 3 * i, free_flag, P.pose_rows + i);
  }
  if (i < n_intr) {
    IntrRow r;
    const int nf = P.nf[i], nd = P.nd[i];
    r.fx = P.focal[2 * i];
    r.fy = (nf == 2) ? P.focal[2 * i + 1] : P.focal[2 * i];
    r.cx = P.center[2 * i];
    r.cy = P.center[2 * i + 1];
    r.k0 = nd >= 1 ? P.dist[2 * i] : 0.0;
    r.k1 = nd >= 2 ? P.dist[2 * i + 1] : 0.0;
    r.nf = static_cast<double>(nf);
    r.nd = static_cast<double>(nd);
    P.intr_rows[i] = r;
  }
}

// ------------------------------------------------------------------ forward transform
struct Forward {
  double mid[3];  // point after the (optional) first pose
  double cam[3];  // camera-frame point
};
__device__ __forceinline__ void forward(const ParamSet& P, const int2 ab, const double X[3], Forward& f) {
  if (ab.y >= 0) {
    transform(P.pose_rows[ab.y], X, f.mid);
  } else {
    f.mid[0] = X[0];
    f.mid[1] = X[1];
    f.mid[2] = X[2];
  }
  transform(P.pose_rows[ab.x], f.mid, f.cam);
}

// ------------------------------------------------------------------------ K1 jacobian
// Residual + analytic Jacobian of ONE observation written to its planes (Jacobi scales and constancy
// masks applied).  The camera-side inputs are references so that the caller decides where they live
// (global tables, or the tile's rows staged in shared memory).
template <int CB, bool TWO>
__device__ __forceinline__ double jacobian_obs(const DeviceProblem& D, int64_t o, const PoseRow& A, const PoseRow* B,
                                               const IntrRow& I, const double* scA, const double* scB,
                                               const double* X, const double* sp, double2 xy) {
  ObsJacobian j;
  observation_jacobian(A, B, I, X, xy.x, xy.y, CB > 0, j);
  double2* J = D.J + o;
  const int64_t ld = D.ld;
  // robust loss (Ceres corrector.cc; Cauchy has rho'' < 0, so residuals and Jacobian are simply
  // scaled by sqrt(rho')): cost term rho(s), planes of the corrected r~, J~
  double cost_term = j.r0 * j.r0 + j.r1 * j.r1;
  double lw = 1.0;
  if (D.loss_type == 1) {
    const double sum = 1.0 + cost_term * D.loss_c;
    lw = sqrt(fmax(DBL_MIN, 1.0 / sum));
    cost_term = D.loss_b * log(sum);
    j.r0 *= lw;
    j.r1 *= lw;
  }
  J[kPlaneR * ld] = make_double2(j.r0, j.r1);
  {
    const double s0 = (sp ? sp[0] : 1.0) * lw, s1 = (sp ? sp[1] : 1.0) * lw, s2 = (sp ? sp[2] : 1.0) * lw;
    J[(kPlaneJp + 0) * ld] = make_double2(j.Jp[0][0] * s0, j.Jp[1][0] * s0);
    J[(kPlaneJp + 1) * ld] = make_double2(j.Jp[0][1] * s1, j.Jp[1][1] * s1);
    J[(kPlaneJp + 2) * ld] = make_double2(j.Jp[0][2] * s2, j.Jp[1][2] * s2);
  }
  if (CB >= 6) {
    double s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = lw;
    if (scA) {
#pragma unroll
      for (int k = 0; k < CB; ++k) s[k] = scA[k] * lw;
#pragma unroll
      for (int k = 0; k < 6; ++k) s[k] *= A.free_;  // constant pose: its six columns vanish
    }
    double2 FA[CB > 0 ? CB : 1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      FA[k] = make_double2(j.JwA[0][k] * s[k], j.JwA[1][k] * s[k]);
      FA[3 + k] = make_double2(j.JtA[0][k] * s[3 + k], j.JtA[1][k] * s[3 + k]);
    }
    if (CB == 9) {
      FA[6] = make_double2(j.df[0] * s[6], j.df[1] * s[6]);
      FA[7] = make_double2(j.dk0[0] * s[7], j.dk0[1] * s[7]);
      FA[8] = make_double2(j.dk1[0] * s[8], j.dk1[1] * s[8]);
    }
#pragma unroll
    for (int k = 0; k < CB; ++k) J[(kPlaneJA + k) * ld] = FA[k];
    if (TWO) {
      const int pb = kPlaneJA + CB;
      if (B) {
        double sb[6] = {lw, lw, lw, lw, lw, lw};
        if (scB) {
#pragma unroll
          for (int k = 0; k < 6; ++k) sb[k] = scB[k] * B->free_ * lw;
        }
        double2 FB[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          FB[k] = make_double2(j.JwB[0][k] * sb[k], j.JwB[1][k] * sb[k]);
          FB[3 + k] = make_double2(j.JtB[0][k] * sb[3 + k], j.JtB[1][k] * sb[3 + k]);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) J[(pb + k) * ld] = FB[k];
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) J[(pb + k) * ld] = make_double2(0.0, 0.0);
      }
    }
  }
  return cost_term;
}

// One thread per observation (point-sorted).  Reads 16 B (xy) + 8 B (indices) + L1/L2-resident
// tables; writes (1 + 3 + CB [+6]) double2 planes.
template <int CB, bool TWO, int MINB>
__global__ void __launch_bounds__(256, MINB) k_jacobian(DeviceProblem D, ParamSet P, WorkArrays W, int unit_scale,
                                                         double* __restrict__ partial_cost) {
  __shared__ double red[32];
  const int64_t o = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  double c = 0.0;
  if (o < D.n_obs) {
    const double2 xy = D.obs_xy[o];
    const int2 idx = D.obs_ip[o];  // (intrinsic, local point)
    const int2 ab = D.obs_ab[o];   // (block a, block b or -1)
    const double* Xp = P.pts + 3 * static_cast<int64_t>(idx.y);
    const double X[3] = {Xp[0], Xp[1], Xp[2]};
    const PoseRow* B = ab.y >= 0 ? P.pose_rows + ab.y : nullptr;  // (TWO only says whether block-B planes are stored)
    c = jacobian_obs<CB, TWO>(D, o, P.pose_rows[ab.x], B, P.intr_rows[idx.x],
                              (unit_scale || CB == 0) ? nullptr : W.sc + static_cast<int64_t>(ab.x) * CB,
                              (unit_scale || CB == 0 || !B) ? nullptr : W.sc + static_cast<int64_t>(ab.y) * CB, X,
                              unit_scale ? nullptr : W.sp + 3 * static_cast<int64_t>(idx.y), xy);
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0 && partial_cost) partial_cost[blockIdx.x] = c;
}

// The same per point tile, with the camera-side inputs of the tile's camera blocks staged in shared
// memory first.  In point order a warp's 32 observations belong to 32 different cameras: gathered
// from the global tables that is 32 L1 tag lookups per load instruction, and the flat kernel is
// bound by exactly that (ncu: L1 data pipe 67 %, DRAM 33 %).  A tile touches few distinct camera
// blocks (its partial list), so their rows (pose row 160 B, Jacobi scales, and the intrinsic row
// when intrinsics are per camera) are copied once per tile and read from shared memory by local index.
constexpr int kJacRowCap = 96;                  // staged camera blocks per tile; the rest falls back to the tables
constexpr int kJacRowLen = 20 + 8 + 9 + 2;      // PoseRow | IntrRow | scales | pad: 39 doubles (odd: spreads the banks)
constexpr int kJacThreads = 256;                // a tile of up to 1024 observations is walked in batches of 256
template <int CB, bool TWO, int MINB>
__global__ void __launch_bounds__(kJacThreads, MINB) k_jacobian_tile(DeviceProblem D, ParamSet P, WorkArrays W, int unit_scale,
                                                                      int intr_is_pose, double* __restrict__ partial_cost) {
  constexpr int T = kJacThreads;
  __shared__ double red[32];
  __shared__ double rows[kJacRowCap * kJacRowLen];
  const int t = blockIdx.x, tid = threadIdx.x;
  const TileMeta tm = D.tile_meta[t];
  const int n_stage = min(tm.n_parts, kJacRowCap);
  constexpr int kCopy = 20 + 8 + (CB > 0 ? CB : 1);
  for (int i = tid; i < n_stage * kCopy; i += T) {
    const int r = i / kCopy, f = i - r * kCopy;
    const int blk = D.part_blk[tm.g0 + r];
    double v;
    if (f < 20)
      v = reinterpret_cast<const double*>(P.pose_rows + blk)[f];
    else if (f < 28)
      v = intr_is_pose ? reinterpret_cast<const double*>(P.intr_rows + blk)[f - 20] : 0.0;
    else
      v = (CB > 0 && !unit_scale) ? W.sc[static_cast<int64_t>(blk) * CB + (f - 28)] : 1.0;
    rows[r * kJacRowLen + f] = v;
  }
  __syncthreads();
  double c = 0.0;
  for (int i = tid; i < tm.n_obs; i += T) {
    const int64_t o = tm.obs0 + i;
    const double2 xy = D.obs_xy[o];
    const int2 idx = D.obs_ip[o];
    const int2 ab = D.obs_ab[o];
    const ushort2 lc = D.obs_lc[o];
    DBA_CHECK(idx.y >= 0 && idx.y < D.n_pts && idx.x >= 0 && idx.x < D.n_intr);
    DBA_CHECK(ab.x >= 0 && ab.x < D.n_ext && ab.y >= -1 && ab.y < D.n_ext);
    DBA_CHECK(CB == 0 || (lc.x < tm.n_parts && (ab.y < 0 || lc.y < tm.n_parts)));
    const double* Xp = P.pts + 3 * static_cast<int64_t>(idx.y);
    const double X[3] = {Xp[0], Xp[1], Xp[2]};
    const double* sp = unit_scale ? nullptr : W.sp + 3 * static_cast<int64_t>(idx.y);
    const bool has_b = ab.y >= 0;  // (TWO only says whether block-B planes are stored)
    const bool no_scale = unit_scale || CB == 0;
    if (lc.x < kJacRowCap && (!has_b || lc.y < kJacRowCap)) {
      // everything camera-side out of shared memory
      const double* rowA = rows + lc.x * kJacRowLen;
      const double* rowB = rows + (has_b ? lc.y : 0) * kJacRowLen;
      const IntrRow& I = intr_is_pose ? *reinterpret_cast<const IntrRow*>(rowA + 20) : P.intr_rows[idx.x];
      c += jacobian_obs<CB, TWO>(D, o, *reinterpret_cast<const PoseRow*>(rowA), has_b ? reinterpret_cast<const PoseRow*>(rowB) : nullptr,
                                I, no_scale ? nullptr : rowA + 28, (no_scale || !has_b) ? nullptr : rowB + 28, X, sp, xy);
    } else {
      // a tile with more camera blocks than the staging area holds: the global tables
      c += jacobian_obs<CB, TWO>(D, o, P.pose_rows[ab.x], has_b ? P.pose_rows + ab.y : nullptr, P.intr_rows[idx.x],
                                no_scale ? nullptr : W.sc + static_cast<int64_t>(ab.x) * CB,
                                (no_scale || !has_b) ? nullptr : W.sc + static_cast<int64_t>(ab.y) * CB, X, sp, xy);
    }
  }
  c = block_sum(c, red);
  if (tid == 0 && partial_cost) partial_cost[t] = c;
}

// ---------------------------------------------------------------------------- K2 cost
__global__ void __launch_bounds__(256) k_cost(DeviceProblem D, ParamSet P, double* __restrict__ partial_cost,
                                               double* __restrict__ mse_out) {
  __shared__ double red[32];
  const int64_t o = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  double c = 0.0;
  if (o < D.n_obs) {
    const double2 xy = D.obs_xy[o];
    const int2 idx = D.obs_ip[o];
    const double* Xp = P.pts + 3 * static_cast<int64_t>(idx.y);
    const double X[3] = {Xp[0], Xp[1], Xp[2]};
    Forward f;
    forward(P, D.obs_ab[o], X, f);
    Projection pr;
    project<false>(P.intr_rows[idx.x], f.cam, xy.x, xy.y, pr);
    c = pr.r0 * pr.r0 + pr.r1 * pr.r1;
    if (mse_out) mse_out[o] = c / 2.0;
    if (D.loss_type == 1) c = D.loss_b * log(1.0 + c * D.loss_c);
  }
  c = block_sum(c, red);
  if (threadIdx.x == 0 && partial_cost) partial_cost[blockIdx.x] = c;
}

// --------------------------------------------------------------- K3 point-side prepare
// One CTA per tile of whole points (<= 256 observations).  Per point: H = E^T E (6 unique),
// g = E^T r; mode 0: Jacobi scale sp = 1/(1+sqrt(diag H)); mode 1: C = H + D^2, C^-1, t = C^-1 g
// and the point part of the gradient norms.  Reads Jp + r planes only (64 B / observation).
template <int T, int MB>
__global__ void __launch_bounds__(T, MB) k_point_prepare(DeviceProblem D, WorkArrays W, double radius,
                                                          double min_diag, double max_diag, int mode,
                                                          double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_pp[];
  double(*v)[T] = reinterpret_cast<double(*)[T]>(smem_pp);  // [9][T]
  __shared__ double red[32];
  const int t = blockIdx.x;
  const int obs0 = D.tile_obs[t], obs1 = D.tile_obs[t + 1];
  const int pt0 = D.tile_pt[t], pt1 = D.tile_pt[t + 1];
  const int tid = threadIdx.x;
  const int o = obs0 + tid;
  if (o < obs1) {
    const double2* J = D.J + o;
    const int64_t ld = D.ld;
    const double2 r = J[kPlaneR * ld];
    const double2 e0 = J[(kPlaneJp + 0) * ld], e1 = J[(kPlaneJp + 1) * ld], e2 = J[(kPlaneJp + 2) * ld];
    v[0][tid] = dot2(e0, e0);
    v[1][tid] = dot2(e0, e1);
    v[2][tid] = dot2(e0, e2);
    v[3][tid] = dot2(e1, e1);
    v[4][tid] = dot2(e1, e2);
    v[5][tid] = dot2(e2, e2);
    v[6][tid] = dot2(e0, r);
    v[7][tid] = dot2(e1, r);
    v[8][tid] = dot2(e2, r);
  }
  __syncthreads();
  double gsq = 0.0, gmax = 0.0, bad = 0.0;
  const int pt = pt0 + tid;
  if (pt < pt1) {
    const int a = D.pt_first[pt] - obs0, b = D.pt_first[pt + 1] - obs0;
    DBA_CHECK(a >= 0 && a <= b && b <= obs1 - obs0 && b <= T);
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = a; i < b; ++i) {
#pragma unroll
      for (int k = 0; k < 9; ++k) s[k] += v[k][i];
    }
    const int64_t p3 = 3 * static_cast<int64_t>(pt);
    if (mode == 0) {
      W.sp[p3 + 0] = 1.0 / (1.0 + sqrt(s[0]));
      W.sp[p3 + 1] = 1.0 / (1.0 + sqrt(s[3]));
      W.sp[p3 + 2] = 1.0 / (1.0 + sqrt(s[5]));
    } else {
      double C[6] = {s[0], s[1], s[2], s[3], s[4], s[5]};
      C[0] += fmin(fmax(s[0], min_diag), max_diag) / radius;
      C[3] += fmin(fmax(s[3], min_diag), max_diag) / radius;
      C[5] += fmin(fmax(s[5], min_diag), max_diag) / radius;
      double inv[6];
      if (!inv_sym3(C, inv)) {
        bad = 1.0;
        inv[0] = inv[1] = inv[2] = inv[3] = inv[4] = inv[5] = 0.0;
      }
      double* ci = W.cinv + 6 * static_cast<int64_t>(pt);
#pragma unroll
      for (int k = 0; k < 6; ++k) ci[k] = inv[k];
      double* tpo = W.tp + 4 * static_cast<int64_t>(pt);  // rows of 4 doubles: one 32-byte load per gather
      tpo[0] = inv[0] * s[6] + inv[1] * s[7] + inv[2] * s[8];
      tpo[1] = inv[1] * s[6] + inv[3] * s[7] + inv[4] * s[8];
      tpo[2] = inv[2] * s[6] + inv[4] * s[7] + inv[5] * s[8];
      tpo[3] = 0.0;
      // unscaled gradient = scaled gradient / scale
      const double g0 = s[6] / W.sp[p3 + 0], g1 = s[7] / W.sp[p3 + 1], g2 = s[8] / W.sp[p3 + 2];
      gsq = g0 * g0 + g1 * g1 + g2 * g2;
      gmax = fmax(fabs(g0), fmax(fabs(g1), fabs(g2)));
    }
  }
  if (mode == 1) {
    gsq = block_sum(gsq, red);
    gmax = block_max(gmax, red);
    bad = block_sum(bad, red);
    if (tid == 0) {
      partials[3 * t + 0] = gsq;
      partials[3 * t + 1] = gmax;
      partials[3 * t + 2] = bad;
    }
  }
}

// ------------------------------------------------------------- K3 camera-side gather
// Camera-sorted incidence list, chunked; one CTA per chunk.  Gathers the Jacobian planes
// of the chunk's observations (16 B per plane per observation), accumulates in registers
//   mode 0: diag(F^T F)                      -> Jacobi scales
//   mode 1: B = F^T (I - E C^-1 E^T) F (upper), diag(F^T F), g = F^T r, rhs = -F^T (r - E t)
// then reduces across the CTA into one row of sums per chunk (k_camera_combine adds the rows of a
// camera block in chunk order: no atomics).
// accumulation of one observation into the chunk sums (shared by the plane-reading and the recomputing kernel)
template <int CB, int MODE>
__device__ __forceinline__ void gather_accumulate(const double2 (&F)[CB], const double2 r, const double2 e0, const double2 e1,
                                                  const double2 e2, const double (&c)[6], const double (&t)[3],
                                                  double (&acc)[MODE == 0 ? CB : CB * (CB + 1) / 2 + 3 * CB]) {
  constexpr int NU = CB * (CB + 1) / 2;
  if (MODE == 0) {
#pragma unroll
    for (int k = 0; k < CB; ++k) acc[k] += dot2(F[k], F[k]);
    return;
  }
  const double c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4], c5 = c[5];
  // M = E C^-1 (2x3), P = I - M E^T (2x2 symmetric)
  double p00 = 1.0, p01 = 0.0, p11 = 1.0;  // mode 2: plain F^T F
  if (MODE == 1) {
    const double m00 = e0.x * c0 + e1.x * c1 + e2.x * c2, m01 = e0.x * c1 + e1.x * c3 + e2.x * c4,
                 m02 = e0.x * c2 + e1.x * c4 + e2.x * c5;
    const double m10 = e0.y * c0 + e1.y * c1 + e2.y * c2, m11 = e0.y * c1 + e1.y * c3 + e2.y * c4,
                 m12 = e0.y * c2 + e1.y * c4 + e2.y * c5;
    p00 = 1.0 - (m00 * e0.x + m01 * e1.x + m02 * e2.x);
    p01 = -(m00 * e0.y + m01 * e1.y + m02 * e2.y);
    p11 = 1.0 - (m10 * e0.y + m11 * e1.y + m12 * e2.y);
  }
  // rr = r - E t
  const double t0 = t[0], t1 = t[1], t2 = t[2];
  const double rr0 = r.x - (e0.x * t0 + e1.x * t1 + e2.x * t2);
  const double rr1 = r.y - (e0.y * t0 + e1.y * t1 + e2.y * t2);
  int u = 0;
#pragma unroll
  for (int i = 0; i < CB; ++i) {
    // (P F)_i
    const double pf0 = MODE == 1 ? p00 * F[i].x + p01 * F[i].y : F[i].x;
    const double pf1 = MODE == 1 ? p01 * F[i].x + p11 * F[i].y : F[i].y;
#pragma unroll
    for (int j = i; j < CB; ++j) {
      acc[u] += pf0 * F[j].x + pf1 * F[j].y;
      ++u;
    }
    acc[NU + i] += dot2(F[i], F[i]);
    acc[NU + CB + i] += dot2(F[i], r);
    acc[NU + 2 * CB + i] -= F[i].x * rr0 + F[i].y * rr1;
  }
}

// CTA-wide sum of the per-thread accumulators into the chunk's row (fixed order: warp tree, then warps 0..3)
template <int NACC>
__device__ __forceinline__ void gather_store(double (&acc)[NACC], double (*red)[NACC], double* __restrict__ out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    const double s = warp_sum(acc[k]);
    if (lane == 0) red[wid][k] = s;
  }
  __syncthreads();
  // one row of sums per chunk; k_camera_combine adds the rows of a camera block in chunk order
  // (no atomics: the accumulators, hence the whole solve, are bit-reproducible)
  for (int k = threadIdx.x; k < NACC; k += blockDim.x) out[k] = red[0][k] + red[1][k] + red[2][k] + red[3][k];
}

template <int CB, int MODE>
__global__ void __launch_bounds__(128) k_camera_gather(DeviceProblem D, WorkArrays W) {
  constexpr int NU = CB * (CB + 1) / 2;
  constexpr int NACC = MODE == 0 ? CB : NU + 3 * CB;
  __shared__ double red[4][NACC];
  const int4 ch = D.cam_chunks[blockIdx.x];  // (block, first entry, last entry, -)
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  const int64_t ld = D.ld;
  for (int e = ch.y + threadIdx.x; e < ch.z; e += blockDim.x) {
    const int ent = D.cam_entries[e];
    const int o = ent >> 1;
    const int slot = ent & 1;
    DBA_CHECK(o >= 0 && o < D.n_obs && (slot == 0 || D.two));
    DBA_CHECK((slot ? D.obs_ab[o].y : D.obs_ab[o].x) == ch.x);
    const double2* J = D.J + o;
    const int base = kPlaneJA + (slot ? D.cb : 0);
    double2 F[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) F[k] = (slot && k >= 6) ? make_double2(0.0, 0.0) : J[(base + k) * ld];
    double2 r = make_double2(0.0, 0.0), e0 = r, e1 = r, e2 = r;
    double c[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, t[3] = {0.0, 0.0, 0.0};
    if (MODE != 0) {
      r = J[kPlaneR * ld];
      e0 = J[(kPlaneJp + 0) * ld];
      e1 = J[(kPlaneJp + 1) * ld];
      e2 = J[(kPlaneJp + 2) * ld];
      const int pt = D.obs_ip[o].y;
      // the kernel is bound by L1 tag lookups (one per gathered sector): C^-1 in three 16-byte
      // loads, t in one 32-byte load
      const double2* ci2 = reinterpret_cast<const double2*>(W.cinv + 6 * static_cast<int64_t>(pt));
      const double2 ca = ci2[0], cbb = ci2[1], cc = ci2[2];
      c[0] = ca.x; c[1] = ca.y; c[2] = cbb.x; c[3] = cbb.y; c[4] = cc.x; c[5] = cc.y;
      double tp[4];
      load_row<4>(W.tp + 4 * static_cast<int64_t>(pt), tp);
      t[0] = tp[0]; t[1] = tp[1]; t[2] = tp[2];
    }
    gather_accumulate<CB, MODE>(F, r, e0, e1, e2, c, t, acc);
  }
  gather_store<NACC>(acc, red, W.cam_chunk_acc + static_cast<int64_t>(blockIdx.x) * NACC);
}

// The same sums with the Jacobian of every observation RECOMPUTED from the camera row (shared by the whole
// chunk: one CTA = one chunk of ONE camera block), the observation's pixel and its point, instead of being
// gathered from the point-sorted planes: 13 half-used sectors per observation (the planes in camera order)
// become xy + indices + X, sp, C^-1, t.  Single-pose problems; same expressions as jacobian_obs, so the
// values are those of the planes.
template <int CB, int MODE>
__global__ void __launch_bounds__(128) k_camera_gather_mf(DeviceProblem D, ParamSet P, WorkArrays W) {
  constexpr int NU = CB * (CB + 1) / 2;
  constexpr int NACC = MODE == 0 ? CB : NU + 3 * CB;
  constexpr int T = 128;
  constexpr int NPT = MODE == 0 ? 3 : 16;  // staged doubles per observation: X | sp | C^-1 | t (4)
  __shared__ double red[4][NACC];
  __shared__ PoseRow sA;
  __shared__ IntrRow sI;
  __shared__ double sSc[CB];
  __shared__ __align__(16) double stage[2][NPT][T];
  const int tid = threadIdx.x;
  const int4 ch = D.cam_chunks[blockIdx.x];  // (block, first entry, last entry, -)
  if (tid < 20) reinterpret_cast<double*>(&sA)[tid] = reinterpret_cast<const double*>(P.pose_rows + ch.x)[tid];
  if (tid >= 32 && tid < 40 && D.intr_is_pose)
    reinterpret_cast<double*>(&sI)[tid - 32] = reinterpret_cast<const double*>(P.intr_rows + ch.x)[tid - 32];
  if (tid >= 64 && tid < 64 + CB) sSc[tid - 64] = MODE == 0 ? 1.0 : W.sc[static_cast<int64_t>(ch.x) * CB + (tid - 64)];
  // Three dependent levels of gathered loads per observation (incidence entry -> pixel + indices -> point
  // data) at 8 resident warps per SM: software pipeline.  Entries are read three iterations ahead, pixel
  // and indices two ahead, and the point data lands one iteration ahead in this thread's own shared-memory
  // slot through cp.async (no registers held, no barrier: a thread reads only what it copied itself).
  auto load_entry = [&](int e) { return e < ch.z ? (D.cam_entries[e] >> 1) : -1; };
  struct Obs {
    double2 xy;
    int2 idx;
  };
  auto load_obs = [&](int o) {
    Obs b;
    b.xy = make_double2(0.0, 0.0);
    b.idx = make_int2(0, -1);
    if (o >= 0) {
      DBA_CHECK(o < D.n_obs && !D.two && D.obs_ab[o].x == ch.x);
      b.xy = D.obs_xy[o];
      b.idx = D.obs_ip[o];
    }
    return b;
  };
  auto issue_point = [&](int buf, int pt) {
    if (pt >= 0) {
      const double* Xp = P.pts + 3 * static_cast<int64_t>(pt);
#pragma unroll
      for (int k = 0; k < 3; ++k) cp_async8(&stage[buf][k][tid], Xp + k);
      if (MODE != 0) {
        const double* sp = W.sp + 3 * static_cast<int64_t>(pt);
#pragma unroll
        for (int k = 0; k < 3; ++k) cp_async8(&stage[buf][3 + k][tid], sp + k);
        const double* ci = W.cinv + 6 * static_cast<int64_t>(pt);
#pragma unroll
        for (int k = 0; k < 6; ++k) cp_async8(&stage[buf][6 + k][tid], ci + k);
        const double* tp = W.tp + 4 * static_cast<int64_t>(pt);
#pragma unroll
        for (int k = 0; k < 3; ++k) cp_async8(&stage[buf][12 + k][tid], tp + k);
      }
    }
    cp_async_commit();
  };
  int e = ch.y + tid;
  int o1 = load_entry(e + T), o2 = load_entry(e + 2 * T);
  Obs b0 = load_obs(load_entry(e));
  Obs b1 = load_obs(o1);
  issue_point(0, b0.idx.y);
  __syncthreads();  // camera row in shared memory
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  for (int it = 0; e < ch.z; ++it, e += T) {
    const int buf = it & 1;
    const int o3 = load_entry(e + 3 * T);
    const Obs b2 = load_obs(o2);
    issue_point(buf ^ 1, b1.idx.y);
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this iteration's point data landed
    const double X[3] = {stage[buf][0][tid], stage[buf][1][tid], stage[buf][2][tid]};
    ObsJacobian j;
    observation_jacobian(sA, nullptr, D.intr_is_pose ? sI : P.intr_rows[b0.idx.x], X, b0.xy.x, b0.xy.y, true, j);
    double lw = 1.0;
    if (D.loss_type == 1) {
      lw = sqrt(fmax(DBL_MIN, 1.0 / (1.0 + (j.r0 * j.r0 + j.r1 * j.r1) * D.loss_c)));
      j.r0 *= lw;
      j.r1 *= lw;
    }
    const double2 r = make_double2(j.r0, j.r1);
    double2 e0 = r, e1 = r, e2 = r;
    double c[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, t[3] = {0.0, 0.0, 0.0};
    double s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = lw;
    if (MODE != 0) {
      const double s0 = stage[buf][3][tid] * lw, s1 = stage[buf][4][tid] * lw, s2 = stage[buf][5][tid] * lw;
      e0 = make_double2(j.Jp[0][0] * s0, j.Jp[1][0] * s0);
      e1 = make_double2(j.Jp[0][1] * s1, j.Jp[1][1] * s1);
      e2 = make_double2(j.Jp[0][2] * s2, j.Jp[1][2] * s2);
#pragma unroll
      for (int k = 0; k < 6; ++k) c[k] = stage[buf][6 + k][tid];
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = stage[buf][12 + k][tid];
#pragma unroll
      for (int k = 0; k < CB; ++k) s[k] = sSc[k] * lw;
#pragma unroll
      for (int k = 0; k < 6; ++k) s[k] *= sA.free_;  // constant pose: its six columns vanish
    }
    double2 F[CB];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      F[k] = make_double2(j.JwA[0][k] * s[k], j.JwA[1][k] * s[k]);
      F[3 + k] = make_double2(j.JtA[0][k] * s[3 + k], j.JtA[1][k] * s[3 + k]);
    }
    if (CB == 9) {
      F[6] = make_double2(j.df[0] * s[6], j.df[1] * s[6]);
      F[7] = make_double2(j.dk0[0] * s[7], j.dk0[1] * s[7]);
      F[8] = make_double2(j.dk1[0] * s[8], j.dk1[1] * s[8]);
    }
    gather_accumulate<CB, MODE>(F, r, e0, e1, e2, c, t, acc);
    b0 = b1;
    b1 = b2;
    o2 = o3;
  }
  cp_async_wait_all();
  gather_store<NACC>(acc, red, W.cam_chunk_acc + static_cast<int64_t>(blockIdx.x) * NACC);
}

// Fixed-order sum of the chunk rows of every camera block into the accumulator layout
//   B [n_blocks][cb][cb] | diagF [n_blocks][cb] | gc [n_blocks][cb] | rhs [n_blocks][cb]
// one thread per (camera block, accumulator); blocks without observations get zeros.
template <int CB, int MODE>
__global__ void __launch_bounds__(128) k_camera_combine(DeviceProblem D, WorkArrays W) {
  constexpr int NU = CB * (CB + 1) / 2;
  constexpr int NACC = MODE == 0 ? CB : NU + 3 * CB;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = D.n_blocks;
  if (idx >= nb * NACC) return;
  const int blk = idx / NACC, k = idx - blk * NACC;
  double s = 0.0;
  for (int c = D.cam_chunk_first[blk]; c < D.cam_chunk_first[blk + 1]; ++c) s += W.cam_chunk_acc[static_cast<int64_t>(c) * NACC + k];
  if (MODE == 0) {
    W.cam_acc[static_cast<int64_t>(nb) * CB * CB + static_cast<int64_t>(blk) * CB + k] = s;
  } else if (k < NU) {
    // unpack upper-triangular index k -> (i, j)
    int i = 0, rem = k;
    while (rem >= CB - i) {
      rem -= CB - i;
      ++i;
    }
    const int j = i + rem;
    double* Bm = W.cam_acc + static_cast<int64_t>(blk) * CB * CB;
    Bm[i * CB + j] = s;
    if (i != j) Bm[j * CB + i] = s;
  } else {
    const int which = (k - NU) / CB, i = (k - NU) % CB;
    W.cam_acc[static_cast<int64_t>(nb) * CB * CB + (static_cast<int64_t>(which) * nb + blk) * CB + i] = s;
  }
}

// Jacobi scales of the camera columns from diag(F^T F) (iteration 0 only).
__global__ void k_camera_scales(DeviceProblem D, WorkArrays W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = D.n_blocks * D.cb;
  if (i < n) {
    const double d = W.cam_acc[static_cast<int64_t>(D.n_blocks) * D.cb * D.cb + i];
    W.sc[i] = 1.0 / (1.0 + sqrt(d));
  }
}

// CB threads per camera block (thread c owns column c of the inverse): D_c^2 = clamp(diag F^T F)/radius,
// M = B + D_c^2, Cholesky of M (every thread of the block's group, redundantly: 165 flops), M^-1 = L^-T L^-1
// column c, camera part of the gradient norms (thread 0 of the group).  The kernel is one serial chain per
// thread and replicated on every rank: a column per thread makes the chain 4x shorter than a camera per thread.
constexpr int kFinThreads = 64;
template <int CB, bool MINV>
__global__ void __launch_bounds__(kFinThreads) k_camera_finalize(DeviceProblem D, WorkArrays W, double radius, double min_diag,
                                                                  double max_diag, double* __restrict__ partials) {
  constexpr int kCams = kFinThreads / CB;  // camera blocks per CTA
  __shared__ double red[32];
  const int grp = threadIdx.x / CB, c = threadIdx.x - grp * CB;
  const int b = blockIdx.x * kCams + grp;
  const int nb = D.n_blocks;
  double gsq = 0.0, gmax = 0.0, bad = 0.0;
  if (grp < kCams && b < nb) {
    const double* Bm = W.cam_acc + static_cast<int64_t>(b) * CB * CB;
    const double* diagF = W.cam_acc + static_cast<int64_t>(nb) * CB * CB + static_cast<int64_t>(b) * CB;
    const double* gc = diagF + static_cast<int64_t>(nb) * CB;
    double L[CB][CB];
#pragma unroll
    for (int i = 0; i < CB; ++i) {
      const double d2 = fmin(fmax(diagF[i], min_diag), max_diag) / radius;
      if (c == 0) {
        W.dc2[static_cast<int64_t>(b) * CB + i] = d2;
        const double g = gc[i] / W.sc[static_cast<int64_t>(b) * CB + i];
        gsq += g * g;
        gmax = fmax(gmax, fabs(g));
      }
#pragma unroll
      for (int j = 0; j < CB; ++j) L[i][j] = Bm[i * CB + j] + (i == j ? d2 : 0.0);
    }
    if (MINV) {
      // in-place lower Cholesky
      bool ok = true;
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) {
          ok = false;
          d = 1.0;
        }
        d = sqrt(d);
        L[j][j] = d;
        const double id = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < CB; ++i) {
          double s = L[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
          L[i][j] = s * id;
        }
      }
      if (!ok && c == 0) bad = 1.0;
      // column c of M^-1 = L^-T L^-1
      double y[CB];
#pragma unroll
      for (int i = 0; i < CB; ++i) {
        double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
        y[i] = s / L[i][i];
      }
#pragma unroll
      for (int i = CB - 1; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < CB; ++k) s -= L[k][i] * y[k];
        y[i] = s / L[i][i];
      }
      double* Mi = W.minv + static_cast<int64_t>(b) * CB * CB;
#pragma unroll
      for (int i = 0; i < CB; ++i) Mi[i * CB + c] = ok ? y[i] : (i == c ? 1.0 : 0.0);
    }
  }
  gsq = block_sum(gsq, red);
  gmax = block_max(gmax, red);
  bad = block_sum(bad, red);
  if (threadIdx.x == 0) {
    partials[3 * blockIdx.x + 0] = gsq;
    partials[3 * blockIdx.x + 1] = gmax;
    partials[3 * blockIdx.x + 2] = bad;
  }
}

// ----------------------------------------------------------------- K5 implicit Schur
// q = (F^T F - F^T E C^-1 E^T F) p  (+ D_c^2 p, added by the PCG vector kernel), atomic-free:
//  k_spmv_tile, ONE pass over the point-sorted Jacobian planes (point tiles of <= 256 obs):
//     u_o = F_o p[blocks(o)]                 (gather of p through L1/L2)
//     y_i = C_i^-1 sum_{o in i} E_o^T u_o     (segmented reduction inside the tile, smem)
//     w_o = u_o - E_o y_i ;  c_o = F_o^T w_o  (per observation, registers -> smem)
//     tile-local reduce-by-camera of c_o through a STATIC per-tile incidence list (built once
//     on upload): one partial vector per (tile, camera present in the tile)
//  k_partials_to_q: per camera block, fixed-order sum of its partial vectors.
// HBM traffic: the Jacobian planes once (16*(3+CB[+6]) + 8 B/obs) plus 2 * 8*CB B per partial;
// with d observations per (tile, camera) pair that is 16*CB/d B/obs extra (d ~ 4.6 on the
// banded bal5m problem, d = 1 in the worst case, where it equals a second pass over Jc).
// History: fp64 RED to L2 serialises at ~150 cycles per cache line on B200 (3.65 ms per
// product); a camera-sorted second pass over a copy of Jc took 0.33 ms.
// --- TMA (cp.async.bulk) + mbarrier helpers: 1-D bulk copies global -> shared, sm_90+ PTX
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Shared-memory plan of k_spmv_tile (dynamic): NP = 3 + CB (+6) planes of (kTile + 1) double2
// (the +1 keeps the camera reduce, which reads 9 consecutive planes at one column, off a single
// bank), then the static tile incidence.  The three Jp planes are recycled as (v, y) pairs and
// the Jc planes as the per-observation contributions once their Jacobian column is consumed.
template <int CB, bool TWO, int T>
struct SpmvSmem {
  static constexpr int NP = 3 + CB + (TWO ? 6 : 0);
  static constexpr int kStride = T + 1;
  static constexpr int kMaxItems = TWO ? 2 * T : T;
  static constexpr size_t kPlaneBytes = static_cast<size_t>(NP) * kStride * sizeof(double2);
  static constexpr size_t kBytes = kPlaneBytes + sizeof(unsigned short) * (2 * kMaxItems + 2) + 16;
};

template <int CB, bool TWO, int T>
__global__ void __launch_bounds__(T, 1024 / T) k_spmv_tile(DeviceProblem D, WorkArrays W) {
  if (W.pcg_state[1]) return;
  constexpr int kTile = T;  // tile capacity == threads per CTA
  using L = SpmvSmem<CB, TWO, T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* sJ = reinterpret_cast<double2*>(smem_raw);  // [NP][kStride]; plane 0..2 = Jp, 3.. = Jc
  unsigned short* s_items = reinterpret_cast<unsigned short*>(smem_raw + L::kPlaneBytes);
  unsigned short* s_first = s_items + L::kMaxItems;
  __shared__ __align__(8) uint64_t bar;
  const int t = blockIdx.x;
  const int tid = threadIdx.x;
  const TileMeta tm = D.tile_meta[t];  // one 32-byte record: no dependent index chain
  const int obs0 = tm.obs0, n_tile = tm.n_obs;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_proxy_async();
    // one elected thread: arm the barrier with the byte count, then one bulk copy per plane
    mbar_expect_tx(&bar, static_cast<uint32_t>(L::NP * n_tile * sizeof(double2)));
    const double2* src = D.J + static_cast<int64_t>(kPlaneJp) * D.ld + obs0;
#pragma unroll
    for (int pl = 0; pl < L::NP; ++pl)
      tma_load_1d(sJ + pl * L::kStride, src + static_cast<int64_t>(pl) * D.ld, static_cast<uint32_t>(n_tile * sizeof(double2)), &bar);
  }
  const int o = obs0 + tid;
  const bool active = tid < n_tile;
  // everything else the CTA needs is requested now, while the bulk copies are in flight:
  // static tile incidence, block indices + p, and the per-point rows used after the first barrier
  for (int i = tid; i < tm.n_items; i += kTile) s_items[i] = D.items[tm.item0 + i];
  for (int i = tid; i <= tm.n_parts; i += kTile) s_first[i] = D.part_first_rel[tm.g0 + t + i];
  int seg_a = 0, seg_b = 0;
  double ci0 = 0.0, ci1 = 0.0, ci2 = 0.0, ci3 = 0.0, ci4 = 0.0, ci5 = 0.0;
  const bool is_pt = tid < tm.n_pts;
  if (is_pt) {
    const int pt = tm.pt0 + tid;
    seg_a = D.pt_first[pt] - obs0;
    seg_b = D.pt_first[pt + 1] - obs0;
    const double2* ci = reinterpret_cast<const double2*>(W.cinv + 6 * static_cast<int64_t>(pt));
    const double2 c01 = ci[0], c23 = ci[1], c45 = ci[2];
    ci0 = c01.x; ci1 = c01.y; ci2 = c23.x; ci3 = c23.y; ci4 = c45.x; ci5 = c45.y;
  }
  double pa[CB];
  double pb[TWO ? 6 : 1];
  int lp = 0;
  bool has_b = false;
  if (active) {
    const int2 ab = D.obs_ab[o];
    lp = D.obs_lp[o];
    const double* pap = W.p + static_cast<int64_t>(ab.x) * CB;
#pragma unroll
    for (int k = 0; k < CB; ++k) pa[k] = pap[k];
    if (TWO) {
      has_b = ab.y >= 0;
      const double* pbp = W.p + static_cast<int64_t>(has_b ? ab.y : 0) * CB;
#pragma unroll
      for (int k = 0; k < 6; ++k) pb[k] = has_b ? pbp[k] : 0.0;
    }
  }
  __syncthreads();  // barrier init visible to all waiters; tile incidence staged
  while (!mbar_try_wait(&bar, 0)) {
  }
  double2 e0 = make_double2(0.0, 0.0), e1 = e0, e2 = e0;
  double u0 = 0.0, u1 = 0.0;
  if (active) {
    e0 = sJ[0 * L::kStride + tid];
    e1 = sJ[1 * L::kStride + tid];
    e2 = sJ[2 * L::kStride + tid];
#pragma unroll
    for (int k = 0; k < CB; ++k) {
      const double2 F = sJ[(3 + k) * L::kStride + tid];
      u0 += F.x * pa[k];
      u1 += F.y * pa[k];
    }
    if (TWO) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double2 F = sJ[(3 + CB + k) * L::kStride + tid];
        u0 += F.x * pb[k];
        u1 += F.y * pb[k];
      }
    }
    // v = E^T u into the .x halves of the (consumed) Jp slots of this thread
    sJ[0 * L::kStride + tid].x = e0.x * u0 + e0.y * u1;
    sJ[1 * L::kStride + tid].x = e1.x * u0 + e1.y * u1;
    sJ[2 * L::kStride + tid].x = e2.x * u0 + e2.y * u1;
  }
  __syncthreads();
  if (is_pt) {
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
    for (int i = seg_a; i < seg_b; ++i) {
      z0 += sJ[0 * L::kStride + i].x;
      z1 += sJ[1 * L::kStride + i].x;
      z2 += sJ[2 * L::kStride + i].x;
    }
    // y of local point `tid` into the .y halves (nobody reads .y before the next barrier)
    sJ[0 * L::kStride + tid].y = ci0 * z0 + ci1 * z1 + ci2 * z2;
    sJ[1 * L::kStride + tid].y = ci1 * z0 + ci3 * z1 + ci4 * z2;
    sJ[2 * L::kStride + tid].y = ci2 * z0 + ci4 * z1 + ci5 * z2;
  }
  __syncthreads();
  if (active) {
    const double y0 = sJ[0 * L::kStride + lp].y, y1 = sJ[1 * L::kStride + lp].y, y2 = sJ[2 * L::kStride + lp].y;
    const double w0 = u0 - (e0.x * y0 + e1.x * y1 + e2.x * y2);
    const double w1 = u1 - (e0.y * y0 + e1.y * y1 + e2.y * y2);
    // contribution c_k = F_k . w overwrites the .x half of the Jacobian slot it came from
#pragma unroll
    for (int k = 0; k < L::NP - 3; ++k) {
      double2* slot = sJ + (3 + k) * L::kStride + tid;
      const double2 F = *slot;
      slot->x = F.x * w0 + F.y * w1;
    }
  }
  __syncthreads();
  // tile-local reduce-by-camera: partial g = g0 + local camera; one work item = (camera, 3 columns),
  // so the incidence list of a camera is walked CB/3 times instead of CB times
  constexpr int KG = CB / 3;
  const int n_work = tm.n_parts * KG;
  for (int wk = tid; wk < n_work; wk += kTile) {
    const int lc = wk / KG, k0 = (wk - lc * KG) * 3;
    const int i0 = s_first[lc], i1 = s_first[lc + 1];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int i = i0; i < i1; ++i) {
      const unsigned int it = s_items[i];
      const int lo = it & 0x7fffu;
      DBA_CHECK(lo < n_tile && i < tm.n_items);
      const int row = TWO ? 3 + (it >> 15) * CB + k0 : 3 + k0;  // slot-B items: planes 3+CB.. (CB == 6 there)
      a0 += sJ[row * L::kStride + lo].x;
      a1 += sJ[(row + 1) * L::kStride + lo].x;
      a2 += sJ[(row + 2) * L::kStride + lo].x;
    }
    DBA_CHECK(D.part_dst[tm.g0 + lc] >= 0 && D.part_dst[tm.g0 + lc] < D.n_partials);
    double* out = W.partials_q + static_cast<int64_t>(D.part_dst[tm.g0 + lc]) * CB + k0;
    out[0] = a0;
    out[1] = a1;
    out[2] = a2;
  }
}

// q_j = sum of the partial vectors of camera block j, in the fixed order of the static list
// (the tile kernels write partial g to row part_dst[g], so a camera's partials are contiguous).
// MF: the partials are in geometric coordinates; q = T^T (sum) = sc * free * (J_l^T . , . , .).
// fuse_dot (single GPU): also q += D_c^2 p and the p.q partial of this block; the last CTA
// publishes p.q, so k_pcg_dot is not launched.
// n_split > 1 (few camera blocks, long partial lists): blockIdx.y owns one slice of the list and
// writes W.q_split[slice]; k_pcg_dot adds the slices in order.
template <int CB, bool MF>
__global__ void __launch_bounds__(128) k_partials_to_q(DeviceProblem D, WorkArrays W, int fuse_dot, int n_split) {
  if (W.pcg_state[1]) return;
  __shared__ double red[4][CB];
  __shared__ double qg[CB];
  __shared__ double red2[32];
  const int blk = blockIdx.x;
  int i0 = D.cam_part_first[blk], i1 = D.cam_part_first[blk + 1];
  if (n_split > 1) {
    const int len = i1 - i0, per = (len + n_split - 1) / n_split;
    i0 = min(i0 + static_cast<int>(blockIdx.y) * per, i1);
    i1 = min(i0 + per, i1);
  }
  double acc[CB];
#pragma unroll
  for (int k = 0; k < CB; ++k) acc[k] = 0.0;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const double* pv = W.partials_q + static_cast<int64_t>(i) * CB;
#pragma unroll
    for (int k = 0; k < CB; ++k) acc[k] += pv[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < CB; ++k) {
    const double s = warp_sum(acc[k]);
    if (lane == 0) red[wid][k] = s;
  }
  __syncthreads();
  if (MF) {
    if (threadIdx.x < CB) qg[threadIdx.x] = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    __syncthreads();
  }
  double pq = 0.0;
  if (threadIdx.x < CB) {
    const int64_t i = static_cast<int64_t>(blk) * CB + threadIdx.x;
    double q;
    if (MF) {
      const double* Tm = W.mf_T + static_cast<int64_t>(blk) * (9 + CB);
      const int k = threadIdx.x;
      q = (k < 3) ? Tm[k] * qg[0] + Tm[3 + k] * qg[1] + Tm[6 + k] * qg[2] : qg[k];  // (J_l^T q)_k
      q *= Tm[9 + k];
    } else {
      q = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    }
    if (fuse_dot) {
      const double p = W.p[i];
      q += W.dc2[i] * p;
      pq = p * q;
    }
    if (n_split > 1)
      W.q_split[static_cast<int64_t>(blockIdx.y) * D.n_blocks * CB + i] = q;
    else
      W.q[i] = q;
  }
  if (!fuse_dot) return;
  if (threadIdx.x < 32) pq = warp_sum(pq);
  if (threadIdx.x == 0) W.vec_partials[blockIdx.x] = pq;
  if (last_block_done(W.counters + 1)) {
    const double tot = sum_partials(W.vec_partials, gridDim.x, red2);
    if (threadIdx.x == 0) {
      W.pcg_scal[2] = tot;
      W.counters[1] = 0;
    }
  }
}

// ------------------------------------------------ K5' matrix-free implicit Schur product
// Same operator as k_spmv_tile, but the Jacobian is RECOMPUTED per observation from the camera
// row and the point instead of being read (~130 fp64 operations against 200 B):
// the kernel reads 8-16 B of indices per observation, X / C^-1 per point and L1-resident camera
// rows, so it leaves the HBM roofline of the materialised product behind.
// Geometric coordinates: with q = R X (rotated point), G = d r / d p_cam (2x3),
//   F p = G (a x q + p~_t) + (d r/d f, k0, k1) p~_fk,   a = p~_w = J_l(w) (sc * free * p)_w
//   F^T w = T^T [q x g, g, (d r/d f,k0,k1)^T w],        g = G^T w
// (d(R(w) X) = (J_l dw) x (R X) is exact; in Ceres' small-angle branch, 0 < |w|^2 <= eps, it differs
// from -[X]x dw by |w| <= 1.5e-8 relative — inside the product only, never in residuals or gradient.)
// Threads of a tile map to observations in camera-block order (static `mf_cols`), so the
// reduce-by-camera runs over contiguous columns and a warp's row loads hit few distinct rows.
// branch selector of a camera row: the last mantissa bit of its t_x (see k_mf_rows)
__device__ __forceinline__ double mf_row_sel(double tx) { return static_cast<double>(__double2loint(tx) & 1); }

// G = d r / d p_cam for projectPoint (snavely_reprojection_error.hh:38-78); also u, v, r^2, d
__device__ __forceinline__ void project_G(double fx, double fy, double k0, double k1, const double c[3], double G[2][3],
                                          double& uu, double& vv, double& rr, double& d) {
  // reciprocal by MUFU seed + two Newton steps (~1 ulp; the product does not need IEEE division)
  double iz;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(iz) : "d"(c[2]));
  iz = iz * (2.0 - c[2] * iz);
  iz = iz * (2.0 - c[2] * iz);
  uu = c[0] * iz;
  vv = c[1] * iz;
  rr = uu * uu + vv * vv;
  const double kk = k0 + k1 * rr;
  d = 1.0 + rr * kk;
  const double kap = kk + k1 * rr;
  const double uv2 = 2.0 * uu * vv * kap;
  const double r0u = fx * (d + 2.0 * uu * uu * kap), r0v = fx * uv2;
  const double r1u = fy * uv2, r1v = fy * (d + 2.0 * vv * vv * kap);
  G[0][0] = r0u * iz;
  G[0][1] = r0v * iz;
  G[0][2] = -(r0u * uu + r0v * vv) * iz;
  G[1][0] = r1u * iz;
  G[1][1] = r1v * iz;
  G[1][2] = -(r1u * uu + r1v * vv) * iz;
}


// Shared-memory plan of k_spmv_mf.  Two stage buffers (tile k computes while tile k+1 lands):
//   cols [T] column records | ptd [12][PS] X, sp, C^-1 of the tile's points (<= kMaxTilePoints by construction)
//   | seg [PM+1] first observation of each point | first [MP+1] partial -> first column/item
//   | dst [MP] partial -> row of the partial buffer | items [2T] (two-pose only)
// and, single: sY [3][PS] per-point y', sV [3][S] v = E^T u in point-sorted order, sC [NC][S]
// contributions in column order.
template <int CB, bool TWO, int T>
struct MfSmem {
  static constexpr int NC = CB + (TWO ? 6 : 0);
  static constexpr int S = T + 1;
  static constexpr int PM = max_tile_points(T);  // points per tile (upload guarantees it)
  static constexpr int PS = PM + 1;
  static constexpr int MP = TWO ? 2 * T : T;  // partials per tile
  static constexpr int kColBytes = CB == 9 ? 8 : 16;
  static constexpr size_t oCols = 0;
  static constexpr size_t oPtd = oCols + static_cast<size_t>(T) * kColBytes;
  static constexpr size_t oSeg = oPtd + 12 * PS * sizeof(double);
  static constexpr size_t oFirst = oSeg + ((PM + 1 + 3) / 4) * 4 * sizeof(int);
  static constexpr size_t oDst = oFirst + ((MP + 1 + 3) / 4) * 4 * sizeof(int);
  static constexpr size_t oItems = oDst + static_cast<size_t>(MP) * sizeof(int);
  static constexpr size_t kStage = ((oItems + (TWO ? 2 * T * sizeof(int) : 0)) + 15) / 16 * 16;
  static constexpr size_t oY = 2 * kStage;                       // [3][PS] y' = sp * C^-1 (sp * sum v) per point
  // sV (phases 1-2) and sC (phases 3-4) have disjoint lifetimes separated by barriers: same memory
  // (12 KB per CTA back to the L1 cache that serves the camera-row loads)
  static constexpr size_t oC = (oY + 3 * PS * sizeof(double) + 15) / 16 * 16;
  static constexpr size_t oV = oC;
  static constexpr size_t kBytes = oC + static_cast<size_t>(NC) * S * sizeof(double);
};

template <int CB, bool MF>
__device__ __forceinline__ void pcg_tail(const DeviceProblem& D, const WorkArrays& W, double tol2, int min_iter, int n_split,
                                         const PeerWin& pw, double* red, int* s_flag);

// Measured in round 2 and not kept (device timestamps, DBA_TAIL_TRACE=1, bal5m, 1 GPU: product 157 us, wait
// for the slowest CTA 14, per-camera sums 17, dot-product phases + two grid barriers 9, launch + drain 9):
//  * all K iterations in one cooperative launch (grid barrier instead of launch + drain): the extra loop level
//    costs 416 bytes of spills inside the tile loop at the 64-register cap, product 157 -> 250 us;
//  * contiguous tile range per CTA with one shared-memory accumulator row per (range, camera): 30x fewer
//    partial rows and per-camera sums 17 -> 9 us, but the product loses 9 us (18 KB less L1 per CTA, wider
//    arrival spread): net zero;
//  * four warps per camera in the per-camera sums: same 17 us (latency of the dependent loads, not list length).
//  * tiles handed out through an atomic counter instead of the static round robin (first tiles static, the draw
//    issued before phase 4, published through shared memory after it): the wait for the slowest CTA drops
//    18 -> 4 us (spread of the CTAs' finish times 26 -> 6 us), but a drawn index is not warp-uniform, so the tile
//    records leave the uniform registers: 28 -> 56 bytes of spills (68 / 100 with the records in a shared-memory
//    ring), product 157 -> 168..176 us: net +15 us per launch.
//  * two observations per thread (T / 2 threads per tile at 128 registers, a pair of one camera shares its row
//    fill and is added in registers before the reduce-by-camera): shared-memory wavefronts 26.3 M -> 20.8 M,
//    instructions -10 %, barrier stalls 8.5 -> 3.8 per issue — and 16 instead of 32 warps per SM, 128 B of spills:
//    issue utilisation 31.6 -> 27 %, product 157 -> 168 us.
// Persistent CTAs (grid = resident CTAs), each walking tiles blockIdx.x, += gridDim.x.  Everything a
// tile needs from HBM is fetched one tile ahead with cp.async into the other stage buffer, so the
// only exposed latency per tile is the L1/L2-resident camera-row load.
template <int CB, bool TWO, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_spmv_mf(DeviceProblem D, WorkArrays W, const double* __restrict__ pts,
                                                      const IntrRow* __restrict__ intr_rows, int fuse_tail, double tol2,
                                                      int min_iter, PeerWin pw) {
  if (W.pcg_state[1]) return;
  const long long t_start = W.trace ? global_ns() : 0;
  using L = MfSmem<CB, TWO, T>;
  constexpr int S = L::S, PS = L::PS, ROW = mf_row_len(CB);
  extern __shared__ __align__(16) unsigned char smem_mf[];
  double* sY = reinterpret_cast<double*>(smem_mf + L::oY);
  double* sV = reinterpret_cast<double*>(smem_mf + L::oV);
  double* sC = reinterpret_cast<double*>(smem_mf + L::oC);
  const int tid = threadIdx.x;

  // stage `buf` <- tile t (record m): every thread issues its share of the copies
  auto issue = [&](int buf, int t, const TileMeta& m) {
    unsigned char* st = smem_mf + buf * L::kStage;
    if (CB == 9)
      cp_async8(st + L::oCols + tid * 8, reinterpret_cast<const int2*>(D.mf_cols) + static_cast<int64_t>(t) * T + tid);
    else
      cp_async16(st + L::oCols + tid * 16, reinterpret_cast<const int4*>(D.mf_cols) + static_cast<int64_t>(t) * T + tid);
    double* ptd = reinterpret_cast<double*>(st + L::oPtd);
    int* seg = reinterpret_cast<int*>(st + L::oSeg);
    if (tid < m.n_pts) {
      const int64_t pt = m.pt0 + tid;
      const double* Xp = pts + 3 * pt;
      const double* sp = W.sp + 3 * pt;
      const double* ci = W.cinv + 6 * pt;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        cp_async8(ptd + k * PS + tid, Xp + k);
        cp_async8(ptd + (3 + k) * PS + tid, sp + k);
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) cp_async8(ptd + (6 + k) * PS + tid, ci + k);
    }
    if (tid <= m.n_pts) cp_async4(seg + tid, D.pt_first + m.pt0 + tid);
    int* first = reinterpret_cast<int*>(st + L::oFirst);
    int* dst = reinterpret_cast<int*>(st + L::oDst);
    for (int i = tid; i <= m.n_parts; i += T) cp_async4(first + i, D.part_first + m.g0 + t + i);
    for (int i = tid; i < m.n_parts; i += T) cp_async4(dst + i, D.part_dst + m.g0 + i);
    if (TWO) {
      int* items = reinterpret_cast<int*>(st + L::oItems);
      for (int i = tid; i < m.n_items; i += T) cp_async4(items + i, D.items_mf + m.item0 + i);
    }
  };

  int t = blockIdx.x;
  TileMeta tm{};
  if (t < D.n_tiles) {
    tm = D.tile_meta[t];
    issue(0, t, tm);
  }
  cp_async_commit();
  int t_next = t + gridDim.x;
  TileMeta tm_next = tm;
  if (t_next < D.n_tiles) tm_next = D.tile_meta[t_next];

  for (int k = 0; t < D.n_tiles; ++k) {
    const int buf = k & 1;
    cp_async_wait_all();
    __syncthreads();  // [A] tile k landed for every thread; everyone is done with tile k-1
    const bool more = t_next < D.n_tiles;
    if (more) issue(buf ^ 1, t_next, tm_next);
    cp_async_commit();
    const TileMeta tm_after = (t_next + static_cast<int>(gridDim.x) < D.n_tiles) ? D.tile_meta[t_next + gridDim.x] : tm_next;

    unsigned char* st = smem_mf + buf * L::kStage;
    const double* ptd = reinterpret_cast<const double*>(st + L::oPtd);
    const int* seg = reinterpret_cast<const int*>(st + L::oSeg);
    const int* s_first = reinterpret_cast<const int*>(st + L::oFirst);
    const int* s_dst = reinterpret_cast<const int*>(st + L::oDst);
    const int* s_items = reinterpret_cast<const int*>(st + L::oItems);
    int blk_a, blk_b = -1, intr = 0;
    unsigned int lplo;
    if (CB == 9) {
      const int2 c = reinterpret_cast<const int2*>(st + L::oCols)[tid];
      blk_a = c.x;
      lplo = static_cast<unsigned int>(c.y);
    } else {
      const int4 c = reinterpret_cast<const int4*>(st + L::oCols)[tid];
      blk_a = c.x;
      blk_b = c.y;
      lplo = static_cast<unsigned int>(c.z);
      intr = c.w;
    }
    const bool active = blk_a >= 0;
    const bool has_b = TWO && blk_b >= 0;
    const int lp = lplo >> 16, lo = lplo & 0xffffu;
    DBA_CHECK(!active || (blk_a < D.n_blocks && blk_b < D.n_blocks && lp < tm.n_pts && lo < tm.n_obs && tid < T));
    DBA_CHECK(tm.n_pts <= L::PM && tm.n_parts <= L::MP && tm.n_obs <= T);
    // ---- phase 1: geometry of this observation, u = F p, v = E^T u
    // (xa, xb: the vector the rotation derivative crosses with: R X, or X itself in Ceres' small-angle branch)
    double G[2][3], E[2][3], GA[2][3];
    double xa[3] = {0.0, 0.0, 0.0}, xb[3] = {0.0, 0.0, 0.0};
    double u0 = 0.0, u1 = 0.0, uu = 0.0, vv = 0.0, rr = 0.0, dd = 0.0, ff = 0.0;
    if (active) {
      double ra[ROW];
      double rb[TWO ? ROW : 2];
      double fx = 0.0, fy = 0.0, k0 = 0.0, k1 = 0.0;
      load_row<ROW>(W.mf_rows + static_cast<int64_t>(blk_a) * ROW, ra);
      if constexpr (TWO) {
        if (has_b) load_row<ROW>(W.mf_rows + static_cast<int64_t>(blk_b) * ROW, rb);
      }
      if (CB != 9) {
        const double2* ir = reinterpret_cast<const double2*>(intr_rows + intr);
        const double2 f2 = __ldg(ir), k2 = __ldg(ir + 2);
        fx = f2.x; fy = f2.y; k0 = k2.x; k1 = k2.y;
      }
      const double X[3] = {ptd[0 * PS + lp], ptd[1 * PS + lp], ptd[2 * PS + lp]};
      double mid[3] = {X[0], X[1], X[2]};
      double dmid[3] = {0.0, 0.0, 0.0};
      if constexpr (TWO) {
        if (has_b) {
          double qb[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) qb[k] = rb[3 * k] * X[0] + rb[3 * k + 1] * X[1] + rb[3 * k + 2] * X[2];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            mid[k] = qb[k] + rb[9 + k];
            xb[k] = X[k] + mf_row_sel(rb[9]) * (qb[k] - X[k]);
          }
          dmid[0] = rb[13] * xb[2] - rb[14] * xb[1] + rb[15];
          dmid[1] = rb[14] * xb[0] - rb[12] * xb[2] + rb[16];
          dmid[2] = rb[12] * xb[1] - rb[13] * xb[0] + rb[17];
        }
      }
      double qa[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) qa[k] = ra[3 * k] * mid[0] + ra[3 * k + 1] * mid[1] + ra[3 * k + 2] * mid[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) xa[k] = mid[k] + mf_row_sel(ra[9]) * (qa[k] - mid[k]);
      const double cam[3] = {qa[0] + ra[9], qa[1] + ra[10], qa[2] + ra[11]};
      if (CB == 9) {
        fx = fy = ra[12];
        k0 = ra[13];
        k1 = ra[14];
      }
      ff = fx;
      project_G(fx, fy, k0, k1, cam, G, uu, vv, rr, dd);
      constexpr int PO = CB == 9 ? 15 : 12;  // p~ offset in the row
      double dc[3];
      dc[0] = ra[PO + 1] * xa[2] - ra[PO + 2] * xa[1] + ra[PO + 3];
      dc[1] = ra[PO + 2] * xa[0] - ra[PO + 0] * xa[2] + ra[PO + 4];
      dc[2] = ra[PO + 0] * xa[1] - ra[PO + 1] * xa[0] + ra[PO + 5];
      if (TWO) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dc[k] += ra[3 * k] * dmid[0] + ra[3 * k + 1] * dmid[1] + ra[3 * k + 2] * dmid[2];
      }
      u0 = G[0][0] * dc[0] + G[0][1] * dc[1] + G[0][2] * dc[2];
      u1 = G[1][0] * dc[0] + G[1][1] * dc[1] + G[1][2] * dc[2];
      if (CB == 9) {
        const double sI = dd * ra[21] + ff * rr * (ra[22] + rr * ra[23]);
        u0 += uu * sI;
        u1 += vv * sI;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) GA[i][j] = G[i][0] * ra[j] + G[i][1] * ra[3 + j] + G[i][2] * ra[6 + j];
      if (TWO && has_b) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) E[i][j] = GA[i][0] * rb[j] + GA[i][1] * rb[3 + j] + GA[i][2] * rb[6 + j];
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) E[i][j] = GA[i][j];
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) sV[j * S + lo] = E[0][j] * u0 + E[1][j] * u1;
    }
    __syncthreads();  // [B]
    // ---- phase 2: one thread per point, y' = sp * C^-1 (sp * sum_o v_o) (E is kept unscaled per
    // observation; the Jacobi scale of the point columns is applied here, once per point)
    if (tid < tm.n_pts) {
      const int sa = seg[tid] - tm.obs0, sb = seg[tid + 1] - tm.obs0;
      DBA_CHECK(sa >= 0 && sa <= sb && sb <= tm.n_obs);
      double z0 = 0.0, z1 = 0.0, z2 = 0.0;
      for (int i = sa; i < sb; ++i) {
        z0 += sV[0 * S + i];
        z1 += sV[1 * S + i];
        z2 += sV[2 * S + i];
      }
      const double p0 = ptd[3 * PS + tid], p1 = ptd[4 * PS + tid], p2 = ptd[5 * PS + tid];
      z0 *= p0;
      z1 *= p1;
      z2 *= p2;
      const double c0 = ptd[6 * PS + tid], c1 = ptd[7 * PS + tid], c2 = ptd[8 * PS + tid], c3 = ptd[9 * PS + tid],
                   c4 = ptd[10 * PS + tid], c5 = ptd[11 * PS + tid];
      sY[0 * PS + tid] = p0 * (c0 * z0 + c1 * z1 + c2 * z2);
      sY[1 * PS + tid] = p1 * (c1 * z0 + c3 * z1 + c4 * z2);
      sY[2 * PS + tid] = p2 * (c2 * z0 + c4 * z1 + c5 * z2);
    }
    __syncthreads();  // [B']
    // ---- phase 3: w = u - E y', contributions F^T w in geometric coordinates
    if (active) {
      const double y0 = sY[0 * PS + lp], y1 = sY[1 * PS + lp], y2 = sY[2 * PS + lp];
      const double w0 = u0 - (E[0][0] * y0 + E[0][1] * y1 + E[0][2] * y2);
      const double w1 = u1 - (E[1][0] * y0 + E[1][1] * y1 + E[1][2] * y2);
      const double g0 = G[0][0] * w0 + G[1][0] * w1, g1 = G[0][1] * w0 + G[1][1] * w1, g2 = G[0][2] * w0 + G[1][2] * w1;
      sC[0 * S + tid] = xa[1] * g2 - xa[2] * g1;
      sC[1 * S + tid] = xa[2] * g0 - xa[0] * g2;
      sC[2 * S + tid] = xa[0] * g1 - xa[1] * g0;
      sC[3 * S + tid] = g0;
      sC[4 * S + tid] = g1;
      sC[5 * S + tid] = g2;
      if (CB == 9) {
        const double sw = uu * w0 + vv * w1;
        const double ck0 = ff * rr * sw;
        sC[6 * S + tid] = dd * sw;
        sC[7 * S + tid] = ck0;
        sC[8 * S + tid] = rr * ck0;
      }
      if (TWO) {
        double h0 = 0.0, h1 = 0.0, h2 = 0.0;
        if (has_b) {
          h0 = GA[0][0] * w0 + GA[1][0] * w1;
          h1 = GA[0][1] * w0 + GA[1][1] * w1;
          h2 = GA[0][2] * w0 + GA[1][2] * w1;
        }
        sC[(CB + 0) * S + tid] = xb[1] * h2 - xb[2] * h1;
        sC[(CB + 1) * S + tid] = xb[2] * h0 - xb[0] * h2;
        sC[(CB + 2) * S + tid] = xb[0] * h1 - xb[1] * h0;
        sC[(CB + 3) * S + tid] = h0;
        sC[(CB + 4) * S + tid] = h1;
        sC[(CB + 5) * S + tid] = h2;
      }
    }
    __syncthreads();  // [C]
    // ---- phase 4: tile-local reduce-by-camera over contiguous columns; one work item = (partial, row):
    // consecutive lanes read consecutive rows of one column run and write one 8*CB-byte partial row
    const int n_work = tm.n_parts * CB;
    for (int wk = tid; wk < n_work; wk += T) {
      const int lc = wk / CB, k = wk - lc * CB;
      const int i0 = s_first[lc], i1 = s_first[lc + 1];
      DBA_CHECK(i0 >= 0 && i0 <= i1 && i1 <= (TWO ? tm.n_items : tm.n_obs) && s_dst[lc] >= 0 && s_dst[lc] < D.n_partials);
      double a0 = 0.0;
      if (TWO) {
        for (int i = i0; i < i1; ++i) {
          const unsigned int it = static_cast<unsigned int>(s_items[i]);
          a0 += sC[((it >> 15) * CB + k) * S + (it & 0x7fffu)];
        }
      } else {
        const double* src = sC + k * S;
        for (int i = i0; i < i1; ++i) a0 += src[i];
      }
      W.partials_q[static_cast<int64_t>(s_dst[lc]) * CB + k] = a0;
    }
    if (!more) break;
    t = t_next;
    t_next += gridDim.x;
    tm = tm_next;
    tm_next = tm_after;
  }
  // ---- epilogue (cooperative launch only): the rest of the PCG iteration in the same kernel —
  // per-camera sum of the partials just written, the cross-rank exchange, x / r / z / p updates
  if (fuse_tail) {
    long long t_done = 0;
    if (W.trace && tid == 0) {
      t_done = global_ns();
      atomicMin(W.trace + 10, static_cast<unsigned long long>(t_done));
      atomicMax(W.trace + 11, static_cast<unsigned long long>(t_done));
    }
    cooperative_groups::this_grid().sync();  // every partial of every tile is in memory
    if (W.trace && tid == 0 && blockIdx.x == 0) {
      const long long t = global_ns();
      W.trace[0] += 1;
      W.trace[1] += static_cast<unsigned long long>(t_done - t_start);
      W.trace[2] += static_cast<unsigned long long>(t - t_done);
      W.trace[9] += W.trace[11] - W.trace[10];
      W.trace[10] = ~0ull;
      W.trace[11] = 0ull;
    }
    pcg_tail<CB, true>(D, W, tol2, min_iter, fuse_tail, pw, reinterpret_cast<double*>(smem_mf), reinterpret_cast<int*>(smem_mf + 512));
  }
}

// Static part of the camera rows and the transform T, one thread per camera block; runs after
// every Jacobian evaluation (parameters and Jacobi scales are fixed until the next one).
template <int CB>
__global__ void __launch_bounds__(128) k_mf_rows(DeviceProblem D, ParamSet P, WorkArrays W) {
  constexpr int ROW = mf_row_len(CB);
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= D.n_blocks) return;
  const PoseRow pr = P.pose_rows[b];
  double* row = W.mf_rows + static_cast<int64_t>(b) * ROW;
#pragma unroll
  for (int k = 0; k < 9; ++k) row[k] = pr.R[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) row[9 + k] = pr.t[k];
  if (CB == 9) {
    const IntrRow ir = P.intr_rows[b];
    row[12] = ir.fx;
    row[13] = ir.k0;
    row[14] = ir.k1;
  }
  // J_l = I + b [w]x + c [w]x^2, c = (theta - sin)/theta^3 (series near 0); identity in the small-angle branch
  const double wx = pr.w[0], wy = pr.w[1], wz = pr.w[2];
  const double th2 = wx * wx + wy * wy + wz * wz;
  double bc = pr.b, cc;
  // branch selector (0: Ceres' small-angle branch, derivative -[X]x) in the last mantissa bit of t_x:
  // a 1-ulp change of the translation seen by the product only, and one 32-byte load less per row
  {
    long long bits = __double_as_longlong(pr.t[0]);
    bits = (th2 > DBL_EPSILON) ? (bits | 1LL) : (bits & ~1LL);
    row[9] = __longlong_as_double(bits);
  }
#pragma unroll
  for (int k = 9 + 3 + (CB == 9 ? 3 : 0) + CB; k < ROW; ++k) row[k] = 0.0;
  if (!(th2 > DBL_EPSILON)) {
    bc = 0.0;
    cc = 0.0;
  } else if (th2 < 1e-2) {
    cc = 1.0 / 6.0 + th2 * (-1.0 / 120.0 + th2 * (1.0 / 5040.0 + th2 * (-1.0 / 362880.0 + th2 * (1.0 / 39916800.0))));
  } else {
    cc = (1.0 - pr.a) / th2;
  }
  double* Tm = W.mf_T + static_cast<int64_t>(b) * (9 + CB);
  Tm[0] = 1.0 + cc * (wx * wx - th2);
  Tm[1] = -bc * wz + cc * wx * wy;
  Tm[2] = bc * wy + cc * wx * wz;
  Tm[3] = bc * wz + cc * wx * wy;
  Tm[4] = 1.0 + cc * (wy * wy - th2);
  Tm[5] = -bc * wx + cc * wy * wz;
  Tm[6] = -bc * wy + cc * wx * wz;
  Tm[7] = bc * wx + cc * wy * wz;
  Tm[8] = 1.0 + cc * (wz * wz - th2);
#pragma unroll
  for (int k = 0; k < CB; ++k) Tm[9 + k] = W.sc[static_cast<int64_t>(b) * CB + k] * (k < 6 ? pr.free_ : 1.0);
}

// PCG phase 3 of the matrix-free path, one thread per camera block: p = z + beta p (init: p as
// left by k_pcg_init), then p~ = T p into the camera row.
template <int CB>
__global__ void __launch_bounds__(128) k_mf_direction(DeviceProblem D, WorkArrays W, int init) {
  if (W.pcg_state[1]) return;
  constexpr int ROW = mf_row_len(CB), PO = CB == 9 ? 15 : 12;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= D.n_blocks) return;
  const double beta = init ? 0.0 : W.pcg_scal[3];
  const double* Tm = W.mf_T + static_cast<int64_t>(b) * (9 + CB);
  double ps[CB];
#pragma unroll
  for (int k = 0; k < CB; ++k) {
    const int64_t i = static_cast<int64_t>(b) * CB + k;
    double p = W.p[i];
    if (!init) {
      p = W.z[i] + beta * p;
      W.p[i] = p;
    }
    ps[k] = p * Tm[9 + k];
  }
  double* row = W.mf_rows + static_cast<int64_t>(b) * ROW + PO;
  row[0] = Tm[0] * ps[0] + Tm[1] * ps[1] + Tm[2] * ps[2];
  row[1] = Tm[3] * ps[0] + Tm[4] * ps[1] + Tm[5] * ps[2];
  row[2] = Tm[6] * ps[0] + Tm[7] * ps[1] + Tm[8] * ps[2];
#pragma unroll
  for (int k = 3; k < CB; ++k) row[k] = ps[k];
}

// x~ = T x into the p~ slots of the camera rows (after the PCG; the next PCG's init refills them with p~)
template <int CB>
__global__ void __launch_bounds__(128) k_mf_step(DeviceProblem D, WorkArrays W) {
  constexpr int ROW = mf_row_len(CB), PO = CB == 9 ? 15 : 12;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= D.n_blocks) return;
  const double* Tm = W.mf_T + static_cast<int64_t>(b) * (9 + CB);
  double xs[CB];
#pragma unroll
  for (int k = 0; k < CB; ++k) xs[k] = W.x[static_cast<int64_t>(b) * CB + k] * Tm[9 + k];
  double* row = W.mf_rows + static_cast<int64_t>(b) * ROW + PO;
  row[0] = Tm[0] * xs[0] + Tm[1] * xs[1] + Tm[2] * xs[2];
  row[1] = Tm[3] * xs[0] + Tm[4] * xs[1] + Tm[5] * xs[2];
  row[2] = Tm[6] * xs[0] + Tm[7] * xs[1] + Tm[8] * xs[2];
#pragma unroll
  for (int k = 3; k < CB; ++k) row[k] = xs[k];
}

// K7 without the camera planes (single-pose, matrix-free solves): u = F x of every observation from the
// camera row (with x~ = T x in its p~ slots, k_mf_step) and the point, as phase 1 of k_spmv_mf; the point
// side (E, r) still comes from the planes.  Lets the Jacobian kernel skip the camera planes altogether.
template <int CB, int T, int MB>
__global__ void __launch_bounds__(T, MB) k_back_substitute_mf(DeviceProblem D, WorkArrays W, const double* __restrict__ pts,
                                                               const IntrRow* __restrict__ intr_rows,
                                                               double* __restrict__ partial_model) {
  constexpr int ROW = mf_row_len(CB);
  extern __shared__ __align__(16) unsigned char smem_bs[];
  double(*v)[T] = reinterpret_cast<double(*)[T]>(smem_bs);  // [3][T]
  double(*y)[T] = v + 3;                                     // [3][T]
  double(*sPt)[T] = v + 6;                                   // [9][T] C^-1 (6), t (3) of this thread's point
  int(*sSeg)[T] = reinterpret_cast<int(*)[T]>(v + 15);       // [2][T] first observation of the point, of the next
  __shared__ double red[32];
  const int t = blockIdx.x;
  const int obs0 = D.tile_obs[t];
  [[maybe_unused]] const int obs1 = D.tile_obs[t + 1];
  const int pt0 = D.tile_pt[t], pt1 = D.tile_pt[t + 1];
  const int tid = threadIdx.x;
  // the point-side inputs of phase 2 do not depend on anything computed here: in flight (cp.async into this
  // thread's own slots) while the observation side runs, instead of a second chain of dependent loads
  const int pt = pt0 + tid;
  if (pt < pt1) {
    cp_async4(&sSeg[0][tid], D.pt_first + pt);
    cp_async4(&sSeg[1][tid], D.pt_first + pt + 1);
    const double* ci = W.cinv + 6 * static_cast<int64_t>(pt);
    const double* tp = W.tp + 4 * static_cast<int64_t>(pt);
#pragma unroll
    for (int k = 0; k < 6; ++k) cp_async8(&sPt[k][tid], ci + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) cp_async8(&sPt[6 + k][tid], tp + k);
  }
  cp_async_commit();
  // threads map to the tile's observations in camera-block order (the column records of k_spmv_mf): a
  // warp's row loads hit few distinct rows
  int blk, intr = 0;
  unsigned int lplo;
  if (CB == 9) {
    const int2 c = reinterpret_cast<const int2*>(D.mf_cols)[static_cast<int64_t>(t) * T + tid];
    blk = c.x;
    lplo = static_cast<unsigned int>(c.y);
  } else {
    const int4 c = reinterpret_cast<const int4*>(D.mf_cols)[static_cast<int64_t>(t) * T + tid];
    blk = c.x;
    lplo = static_cast<unsigned int>(c.z);
    intr = c.w;
  }
  const bool active = blk >= 0;
  const int lp = lplo >> 16, lo = lplo & 0xffffu;
  double2 e0, e1, e2, r;
  double u0 = 0.0, u1 = 0.0;
  if (active) {
    DBA_CHECK(lp < pt1 - pt0 && lo < obs1 - obs0 && blk < D.n_blocks);
    const double2* J = D.J + obs0 + lo;
    const int64_t ld = D.ld;
    r = J[kPlaneR * ld];
    e0 = J[(kPlaneJp + 0) * ld];
    e1 = J[(kPlaneJp + 1) * ld];
    e2 = J[(kPlaneJp + 2) * ld];
    double ra[ROW];
    load_row<ROW>(W.mf_rows + static_cast<int64_t>(blk) * ROW, ra);
    double fx, fy, k0, k1;
    if (CB == 9) {
      fx = fy = ra[12];
      k0 = ra[13];
      k1 = ra[14];
    } else {
      const double2* ir = reinterpret_cast<const double2*>(intr_rows + intr);
      const double2 f2 = __ldg(ir), k2 = __ldg(ir + 2);
      fx = f2.x; fy = f2.y; k0 = k2.x; k1 = k2.y;
    }
    const double* Xp = pts + 3 * static_cast<int64_t>(pt0 + lp);
    const double X[3] = {Xp[0], Xp[1], Xp[2]};
    double qa[3], xa[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) qa[k] = ra[3 * k] * X[0] + ra[3 * k + 1] * X[1] + ra[3 * k + 2] * X[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) xa[k] = X[k] + mf_row_sel(ra[9]) * (qa[k] - X[k]);
    const double cam[3] = {qa[0] + ra[9], qa[1] + ra[10], qa[2] + ra[11]};
    double G[2][3], uu, vv, rr, dd;
    project_G(fx, fy, k0, k1, cam, G, uu, vv, rr, dd);
    constexpr int PO = CB == 9 ? 15 : 12;  // x~ offset in the row
    double dc[3];
    dc[0] = ra[PO + 1] * xa[2] - ra[PO + 2] * xa[1] + ra[PO + 3];
    dc[1] = ra[PO + 2] * xa[0] - ra[PO + 0] * xa[2] + ra[PO + 4];
    dc[2] = ra[PO + 0] * xa[1] - ra[PO + 1] * xa[0] + ra[PO + 5];
    u0 = G[0][0] * dc[0] + G[0][1] * dc[1] + G[0][2] * dc[2];
    u1 = G[1][0] * dc[0] + G[1][1] * dc[1] + G[1][2] * dc[2];
    if (CB == 9) {
      const double sI = dd * ra[21] + fx * rr * (ra[22] + rr * ra[23]);
      u0 += uu * sI;
      u1 += vv * sI;
    }
    v[0][lo] = e0.x * u0 + e0.y * u1;
    v[1][lo] = e1.x * u0 + e1.y * u1;
    v[2][lo] = e2.x * u0 + e2.y * u1;
  }
  cp_async_wait_all();
  __syncthreads();
  if (pt < pt1) {
    const int a = sSeg[0][tid] - obs0, b = sSeg[1][tid] - obs0;
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
    for (int i = a; i < b; ++i) {
      z0 += v[0][i];
      z1 += v[1][i];
      z2 += v[2][i];
    }
    const double d0 = -sPt[6][tid] - (sPt[0][tid] * z0 + sPt[1][tid] * z1 + sPt[2][tid] * z2);
    const double d1 = -sPt[7][tid] - (sPt[1][tid] * z0 + sPt[3][tid] * z1 + sPt[4][tid] * z2);
    const double d2 = -sPt[8][tid] - (sPt[2][tid] * z0 + sPt[4][tid] * z1 + sPt[5][tid] * z2);
    double* dp = W.dp + 3 * static_cast<int64_t>(pt);
    dp[0] = d0;
    dp[1] = d1;
    dp[2] = d2;
    y[0][tid] = d0;
    y[1][tid] = d1;
    y[2][tid] = d2;
  }
  __syncthreads();
  double m = 0.0;
  if (active) {
    const double y0 = y[0][lp], y1 = y[1][lp], y2 = y[2][lp];
    const double jd0 = u0 + e0.x * y0 + e1.x * y1 + e2.x * y2;
    const double jd1 = u1 + e0.y * y0 + e1.y * y1 + e2.y * y2;
    m = jd0 * (r.x + 0.5 * jd0) + jd1 * (r.y + 0.5 * jd1);
  }
  m = block_sum(m, red);
  if (tid == 0) partial_model[t] = m;
}

// --------------------------------------------------------------------------- K6 PCG
// Block-Jacobi PCG vector work, multi-CTA (one thread per unknown), deterministic: every
// global scalar is a fixed-order sum of per-CTA partials done by the last CTA to finish
// ("last block" pattern with a device counter), so no cooperative launch is needed.
// x = 0, r = rhs, z = M^-1 r, p = z, rz = rz0 = r.z
__global__ void __launch_bounds__(256) k_pcg_init(DeviceProblem D, WorkArrays W) {
  __shared__ double red[32];
  const int n = D.n_blocks * D.cb, cb = D.cb;
  const double* rhs = W.cam_acc + static_cast<int64_t>(D.n_blocks) * cb * cb + 2 * static_cast<int64_t>(n);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n) {
    const int b = i / cb, row = i - b * cb;
    const double* Mi = W.minv + (static_cast<int64_t>(b) * cb + row) * cb;
    double z = 0.0;
    for (int k = 0; k < cb; ++k) z += Mi[k] * rhs[b * cb + k];
    W.x[i] = 0.0;
    W.r[i] = rhs[i];
    W.z[i] = z;
    W.p[i] = z;
    acc = rhs[i] * z;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) W.vec_partials[blockIdx.x] = acc;
  if (last_block_done(W.counters + 0)) {
    const double rz = sum_partials(W.vec_partials, gridDim.x, red);
    if (threadIdx.x == 0) {
      W.pcg_scal[0] = rz;
      W.pcg_scal[1] = rz;
      W.pcg_state[0] = 0;
      W.pcg_state[1] = (rz > 0.0) ? 0 : 1;  // zero right-hand side: nothing to do
      W.counters[0] = 0;
    }
  }
}

// phase 1: q (= sum of the slices when n_split > 1) += D_c^2 p, partial p.q; the last CTA publishes p.q
__global__ void __launch_bounds__(256) k_pcg_dot(DeviceProblem D, WorkArrays W, int n_split) {
  if (W.pcg_state[1]) return;
  __shared__ double red[32];
  const int n = D.n_blocks * D.cb;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n) {
    const double p = W.p[i];
    double q;
    if (n_split > 1) {
      q = 0.0;
      for (int sl = 0; sl < n_split; ++sl) q += W.q_split[static_cast<int64_t>(sl) * n + i];
    } else {
      q = W.q[i];
    }
    q += W.dc2[i] * p;
    W.q[i] = q;
    acc = p * q;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) W.vec_partials[blockIdx.x] = acc;
  if (last_block_done(W.counters + 1)) {
    const double pq = sum_partials(W.vec_partials, gridDim.x, red);
    if (threadIdx.x == 0) {
      W.pcg_scal[2] = pq;
      W.counters[1] = 0;
    }
  }
}

// phase 2: alpha = rz / p.q; x += alpha p; r -= alpha q; z = M^-1 r; partial r.z; the last CTA
// publishes rz_new, beta and the convergence flag.  One thread per unknown; a CTA owns whole
// camera blocks and stages their new residual in shared memory for the block mat-vec.
__global__ void __launch_bounds__(256) k_pcg_step(DeviceProblem D, WorkArrays W, double tol2, int min_iter) {
  if (W.pcg_state[1]) return;
  __shared__ double red[32];
  __shared__ double rn[256];
  const int cb = D.cb;
  const int rows_per_cta = (256 / cb) * cb;
  const int n = D.n_blocks * cb;
  const int i = blockIdx.x * rows_per_cta + threadIdx.x;
  const bool active = threadIdx.x < rows_per_cta && i < n;
  const double pq = W.pcg_scal[2], rz = W.pcg_scal[0];
  const bool breakdown = !(pq > 0.0) || !isfinite(pq);
  double acc = 0.0;
  if (!breakdown) {
    const double alpha = rz / pq;
    if (active) {
      W.x[i] += alpha * W.p[i];
      const double r = W.r[i] - alpha * W.q[i];
      W.r[i] = r;
      rn[threadIdx.x] = r;
    }
    __syncthreads();
    if (active) {
      const int lb = threadIdx.x / cb, row = threadIdx.x - lb * cb;
      const double* Mi = W.minv + static_cast<int64_t>(i) * cb;  // row i of the block-diagonal inverse
      double z = 0.0;
      for (int k = 0; k < cb; ++k) z += Mi[k] * rn[lb * cb + k];
      W.z[i] = z;
      acc = rn[threadIdx.x] * z;
      (void)row;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) W.vec_partials[blockIdx.x] = acc;
  if (last_block_done(W.counters + 2)) {
    const double rz_new = sum_partials(W.vec_partials, gridDim.x, red);
    if (threadIdx.x == 0) {
      W.counters[2] = 0;
      if (breakdown) {
        W.pcg_state[1] = 1;  // keep the current x
        W.pcg_scal[3] = 0.0;
      } else {
        const int it = W.pcg_state[0] + 1;
        W.pcg_state[0] = it;
        W.pcg_scal[3] = rz_new / rz;  // beta
        W.pcg_scal[0] = rz_new;
        if ((it >= min_iter && rz_new <= tol2 * W.pcg_scal[1]) || !(rz_new > 0.0)) W.pcg_state[1] = 1;
      }
    }
  }
}

// phase 3: p = z + beta p
__global__ void __launch_bounds__(256) k_pcg_direction(DeviceProblem D, WorkArrays W) {
  if (W.pcg_state[1]) return;
  const int n = D.n_blocks * D.cb;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) W.p[i] = W.z[i] + W.pcg_scal[3] * W.p[i];
}

// 16-byte records of the LL exchange: one vector store / one volatile vector load
__device__ __forceinline__ void ll_store(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// PCG iteration tail in ONE cooperative launch (grid = co-resident CTAs, grid.sync between the
// phases): per camera block q = T^T sum(partials) + D_c^2 p and p.q | alpha, x, r, z = M^-1 r and
// r.z | beta, p = z + beta p, p~ = T p.  Replaces k_partials_to_q + k_pcg_step + k_mf_direction
// (three launch latencies per PCG iteration).  All sums are fixed-order.
// Multi-GPU (pw.world > 1): the allreduce of q is part of the same launch — each rank stores its
// share of q into slot `rank` of every rank's peer window (NVLink P2P stores), raises one flag per
// peer, waits for its own flags and adds the slots in rank order; x, r, z, p then evolve
// bit-identically on every rank, so the dot products need no further communication.
// The tail of one PCG iteration, executed by EVERY thread of a cooperative grid (any grid / block
// shape; `red` = 32 doubles of shared memory).  One WARP per camera block, no block-level barrier
// inside the phases:
//   1. q = T^T sum(partials) [+ exchange with the peers] + D_c^2 p, p.q        | grid.sync
//   2. alpha = rz / p.q; x += alpha p; r -= alpha q; z = M^-1 r; r.z            | grid.sync
//   3. beta = r.z / rz; p = z + beta p; p~ = T p into the camera rows (MF)
// Every sum has a fixed order (lane-strided rows, butterfly, per-CTA partials in CTA order).
template <int CB, bool MF>
__device__ __forceinline__ void pcg_tail(const DeviceProblem& D, const WorkArrays& W, double tol2, int min_iter, int n_split,
                                         const PeerWin& pw, double* red, int* s_flag) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  constexpr unsigned kFull = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wpb = blockDim.x >> 5;
  const int gwarp = blockIdx.x * wpb + (tid >> 5), n_warps = gridDim.x * wpb;
  const int nb = D.n_blocks;
  const bool exchange = pw.world > 1;
  const long long buf_off = static_cast<long long>(pw.seq & 1ull) * pw.world * pw.slot_len;
  double acc_dot = 0.0;
  const bool tracer = W.trace && blockIdx.x == 0 && tid == 0;
  long long t_prev = tracer ? global_ns() : 0;
  auto stamp = [&](int slot) {
    if (tracer) {
      const long long t = global_ns();
      W.trace[slot] += static_cast<unsigned long long>(t - t_prev);
      t_prev = t;
    }
  };
  // ---- phase 1
  // streams rows [r0, r1) of the camera-grouped partial buffer: one contiguous run of (r1 - r0) * CB
  // doubles, fully coalesced 8-byte loads, kM loads (= one chunk of 32 * kM elements, a whole number
  // of rows) in flight per lane; element e of the run belongs to column e % CB, and within a chunk
  // lane l / load m always sees column (32 m + l) % CB.  Returns column `lane` in lane < CB.
  auto run_sum = [&](int r0, int r1) -> double {
    constexpr int kG = (CB % 2 == 0) ? 2 : 1;  // gcd(32, CB) for CB in {6, 9}
    constexpr int kM = CB / kG;
    const double* run = W.partials_q + static_cast<int64_t>(r0) * CB;
    const int n_el = (r1 - r0) * CB;
    double part[kM];
#pragma unroll
    for (int m = 0; m < kM; ++m) part[m] = 0.0;
    for (int e = lane; e < n_el; e += 32 * kM) {
#pragma unroll
      for (int m = 0; m < kM; ++m) {
        const int ee = e + 32 * m;
        if (ee < n_el) part[m] += __ldcg(run + ee);
      }
    }
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < CB; ++k) {
      double v = 0.0;
#pragma unroll
      for (int m = 0; m < kM; ++m) v = ((32 * m + lane) % CB == k) ? part[m] : v;
      v = warp_sum_all(v);
      if (lane == k) mine = v;
    }
    return mine;
  };
  if (n_split > 1) {
    // few camera blocks, long partial lists: n_split warps per block sum one slice each, then the
    // slices are added in slice order
    const int n_items = nb * n_split;
    for (int w = gwarp; w < n_items; w += n_warps) {
      const int blk = w / n_split, sl = w - blk * n_split;
      const int i0 = D.cam_part_first[blk], len = D.cam_part_first[blk + 1] - i0;
      const int r0 = i0 + static_cast<int>(static_cast<int64_t>(len) * sl / n_split);
      const int r1 = i0 + static_cast<int>(static_cast<int64_t>(len) * (sl + 1) / n_split);
      const double v = run_sum(r0, r1);
      if (lane < CB) W.q_split[static_cast<int64_t>(sl) * nb * CB + static_cast<int64_t>(blk) * CB + lane] = v;
    }
    grid.sync();
  }
  for (int blk = gwarp; blk < nb; blk += n_warps) {
    double q = 0.0;
    if (n_split > 1) {
      if (lane < CB)
        for (int sl = 0; sl < n_split; ++sl)
          q += __ldcg(W.q_split + static_cast<int64_t>(sl) * nb * CB + static_cast<int64_t>(blk) * CB + lane);
    } else {
      q = run_sum(D.cam_part_first[blk], D.cam_part_first[blk + 1]);
    }
    const double a0 = __shfl_sync(kFull, q, 0), a1 = __shfl_sync(kFull, q, 1), a2 = __shfl_sync(kFull, q, 2);
    if (lane < CB) {
      const int64_t i = static_cast<int64_t>(blk) * CB + lane;
      if (MF) {
        const double* Tm = W.mf_T + static_cast<int64_t>(blk) * (9 + CB);
        if (lane < 3) q = Tm[lane] * a0 + Tm[3 + lane] * a1 + Tm[6 + lane] * a2;
        q *= Tm[9 + lane];
      }
      if (exchange) {
        // push this rank's share of q into slot `rank` of every window (own window included) as a
        // self-validating record {low half, seq, high half, seq}: the flag travels with the data, so the
        // exchange needs no fence, no separate flag and no grid-wide barrier (NCCL's LL protocol)
        const long long off = buf_off + static_cast<long long>(pw.rank) * pw.slot_len + i;
        const unsigned int sq = static_cast<unsigned int>(pw.seq);
        const uint4 rec = make_uint4(static_cast<unsigned int>(__double2loint(q)), sq, static_cast<unsigned int>(__double2hiint(q)), sq);
#pragma unroll 1
        for (int r = 0; r < pw.world; ++r) ll_store(pw.ll[r] + off, rec);
      } else {
        const double p = W.p[i];
        q += W.dc2[i] * p;
        W.q[i] = q;
        acc_dot += p * q;
      }
    }
  }
  stamp(3);
  if (exchange) {
    // q = sum of the slots in rank order (the same order on every rank) + D_c^2 p; a record is consumed
    // as soon as both halves carry this exchange's sequence number.  A peer that never delivers raises
    // the error flag after the timeout; the launch still runs to its end (every grid.sync is attended).
    const uint4* mine = pw.ll[pw.rank] + buf_off;
    const unsigned int sq = static_cast<unsigned int>(pw.seq);
    for (int blk = gwarp; blk < nb; blk += n_warps) {
      if (lane < CB) {
        const int64_t i = static_cast<int64_t>(blk) * CB + lane;
        double q = 0.0;
#pragma unroll 1
        for (int r = 0; r < pw.world; ++r) {
          const uint4* src = mine + static_cast<long long>(r) * pw.slot_len + i;
          uint4 rec = ll_load(src);
          if (rec.y != sq || rec.w != sq) {
            const long long t0 = global_ns();
            do {
              rec = ll_load(src);
              if (*reinterpret_cast<volatile int*>(W.pcg_state + 2)) break;  // somebody already gave up: do not wait again
              if (global_ns() - t0 > pw.timeout_ns) {
                W.pcg_state[2] = 1;
                W.pcg_state[1] = 1;
                break;
              }
            } while (rec.y != sq || rec.w != sq);
          }
          q += __hiloint2double(static_cast<int>(rec.z), static_cast<int>(rec.x));
        }
        const double p = W.p[i];
        q += W.dc2[i] * p;
        W.q[i] = q;
        acc_dot += p * q;
      }
    }
  }
  acc_dot = block_sum(acc_dot, red);
  if (tid == 0) W.vec_partials[blockIdx.x] = acc_dot;
  stamp(4);
  grid.sync();
  stamp(5);
  // ---- phase 2
  const double pq = sum_partials(W.vec_partials, gridDim.x, red);
  const double rz = W.pcg_scal[0];
  const bool breakdown = !(pq > 0.0) || !isfinite(pq);
  acc_dot = 0.0;
  if (!breakdown) {
    const double alpha = rz / pq;
    for (int blk = gwarp; blk < nb; blk += n_warps) {
      const int64_t i = static_cast<int64_t>(blk) * CB + lane;
      double r = 0.0;
      if (lane < CB) {
        W.x[i] += alpha * W.p[i];
        r = W.r[i] - alpha * W.q[i];
        W.r[i] = r;
      }
      double z = 0.0;
      const double* Mi = W.minv + (lane < CB ? i : static_cast<int64_t>(blk) * CB) * CB;
#pragma unroll
      for (int k = 0; k < CB; ++k) z += Mi[k] * __shfl_sync(kFull, r, k);
      if (lane < CB) {
        W.z[i] = z;
        acc_dot += r * z;
      }
    }
  }
  acc_dot = block_sum(acc_dot, red);
  if (tid == 0) W.vec_partials[gridDim.x + blockIdx.x] = acc_dot;
  stamp(6);
  grid.sync();
  stamp(7);
  if (breakdown) {
    if (blockIdx.x == 0 && tid == 0) {
      W.pcg_state[1] = 1;  // keep the current x
      W.pcg_scal[3] = 0.0;
    }
    return;
  }
  // ---- phase 3
  const double rz_new = sum_partials(W.vec_partials + gridDim.x, gridDim.x, red);
  const double beta = rz_new / rz;
  for (int blk = gwarp; blk < nb; blk += n_warps) {
    const int64_t i = static_cast<int64_t>(blk) * CB + lane;
    double ps = 0.0;
    if (lane < CB) {
      const double p = W.z[i] + beta * W.p[i];
      W.p[i] = p;
      if (MF) ps = p * W.mf_T[static_cast<int64_t>(blk) * (9 + CB) + 9 + lane];
    }
    if (MF) {
      constexpr int ROW = mf_row_len(CB), PO = CB == 9 ? 15 : 12;
      const double p0 = __shfl_sync(kFull, ps, 0), p1 = __shfl_sync(kFull, ps, 1), p2 = __shfl_sync(kFull, ps, 2);
      if (lane < CB) {
        const double* Tm = W.mf_T + static_cast<int64_t>(blk) * (9 + CB);
        double v = ps;
        if (lane < 3) v = Tm[3 * lane] * p0 + Tm[3 * lane + 1] * p1 + Tm[3 * lane + 2] * p2;
        W.mf_rows[static_cast<int64_t>(blk) * ROW + PO + lane] = v;
      }
    }
  }
  stamp(8);
  if (blockIdx.x == 0 && tid == 0) {
    const int it = W.pcg_state[0] + 1;
    W.pcg_state[0] = it;
    W.pcg_scal[3] = beta;
    W.pcg_scal[0] = rz_new;
    if ((it >= min_iter && rz_new <= tol2 * W.pcg_scal[1]) || !(rz_new > 0.0)) W.pcg_state[1] = 1;
  }
}

// Stand-alone launch of the tail (after k_spmv_tile, or after a k_spmv_mf that could not be
// launched cooperatively); k_spmv_mf normally runs it as its own epilogue.
constexpr int kFusedThreads = 256;
template <int CB, bool MF>
__global__ void __launch_bounds__(kFusedThreads) k_pcg_fused(DeviceProblem D, WorkArrays W, double tol2, int min_iter, int n_split,
                                                             PeerWin pw) {
  if (W.pcg_state[1]) return;
  __shared__ double red[32];
  __shared__ int s_flag;
  pcg_tail<CB, MF>(D, W, tol2, min_iter, n_split, pw, red, &s_flag);
}

// ------------------------------------------------------------- K7 back-substitution
// dp_i = -t_i - C_i^-1 sum_o E_o^T F_o x ;  partial of sum (J d).(r + J d / 2)
template <int CB, bool TWO, int T, int MB>
__global__ void __launch_bounds__(T, MB) k_back_substitute(DeviceProblem D, WorkArrays W,
                                                            double* __restrict__ partial_model) {
  extern __shared__ __align__(16) unsigned char smem_bs[];
  double(*v)[T] = reinterpret_cast<double(*)[T]>(smem_bs);  // [3][T]
  double(*y)[T] = v + 3;                                     // [3][T]
  __shared__ double red[32];
  const int t = blockIdx.x;
  const int obs0 = D.tile_obs[t], obs1 = D.tile_obs[t + 1];
  const int pt0 = D.tile_pt[t], pt1 = D.tile_pt[t + 1];
  const int tid = threadIdx.x;
  const int o = obs0 + tid;
  const bool active = o < obs1;
  double2 e0, e1, e2, r;
  double u0 = 0.0, u1 = 0.0;
  int lp = 0;
  if (active) {
    lp = D.obs_ip[o].y - pt0;
    DBA_CHECK(lp >= 0 && lp < pt1 - pt0 && lp < T);
    const double2* J = D.J + o;
    const int64_t ld = D.ld;
    r = J[kPlaneR * ld];
    e0 = J[(kPlaneJp + 0) * ld];
    e1 = J[(kPlaneJp + 1) * ld];
    e2 = J[(kPlaneJp + 2) * ld];
    if (CB > 0) {
      const int2 ab = D.obs_ab[o];
      struct { int pose_a, pose_b; } vw = {ab.x, ab.y};
      const double* xa = W.x + static_cast<int64_t>(vw.pose_a) * CB;
#pragma unroll
      for (int k = 0; k < CB; ++k) {
        const double2 F = J[(kPlaneJA + k) * ld];
        u0 += F.x * xa[k];
        u1 += F.y * xa[k];
      }
      if (TWO) {
        if (vw.pose_b >= 0) {
          const double* xb = W.x + static_cast<int64_t>(vw.pose_b) * CB;
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const double2 F = J[(kPlaneJA + CB + k) * ld];
            u0 += F.x * xb[k];
            u1 += F.y * xb[k];
          }
        }
      }
    }
    v[0][tid] = e0.x * u0 + e0.y * u1;
    v[1][tid] = e1.x * u0 + e1.y * u1;
    v[2][tid] = e2.x * u0 + e2.y * u1;
  }
  __syncthreads();
  const int pt = pt0 + tid;
  if (pt < pt1) {
    const int a = D.pt_first[pt] - obs0, b = D.pt_first[pt + 1] - obs0;
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
    for (int i = a; i < b; ++i) {
      z0 += v[0][i];
      z1 += v[1][i];
      z2 += v[2][i];
    }
    const double* ci = W.cinv + 6 * static_cast<int64_t>(pt);
    const double* tp = W.tp + 4 * static_cast<int64_t>(pt);
    const double d0 = -tp[0] - (ci[0] * z0 + ci[1] * z1 + ci[2] * z2);
    const double d1 = -tp[1] - (ci[1] * z0 + ci[3] * z1 + ci[4] * z2);
    const double d2 = -tp[2] - (ci[2] * z0 + ci[4] * z1 + ci[5] * z2);
    double* dp = W.dp + 3 * static_cast<int64_t>(pt);
    dp[0] = d0;
    dp[1] = d1;
    dp[2] = d2;
    y[0][tid] = d0;
    y[1][tid] = d1;
    y[2][tid] = d2;
  }
  __syncthreads();
  double m = 0.0;
  if (active) {
    const double y0 = y[0][lp], y1 = y[1][lp], y2 = y[2][lp];
    const double jd0 = u0 + e0.x * y0 + e1.x * y1 + e2.x * y2;
    const double jd1 = u1 + e0.y * y0 + e1.y * y1 + e2.y * y2;
    m = jd0 * (r.x + 0.5 * jd0) + jd1 * (r.y + 0.5 * jd1);
  }
  m = block_sum(m, red);
  if (tid == 0) partial_model[t] = m;
}

// ---------------------------------------------------------------------- param update
// candidate = current + scale * step for points (thread per scalar); partial sums of
// step^2 and x^2 (free parameters only, as Ceres' reduced program).
__global__ void __launch_bounds__(256) k_update_points(int64_t n3, const double* __restrict__ cur,
                                                        double* __restrict__ cand, const double* __restrict__ sp,
                                                        const double* __restrict__ dp, double* __restrict__ partials) {
  __shared__ double red[32];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  double ssq = 0.0, xsq = 0.0;
  if (i < n3) {
    const double x = cur[i];
    const double d = sp[i] * dp[i];
    cand[i] = x + d;
    ssq = d * d;
    xsq = x * x;
  }
  ssq = block_sum(ssq, red);
  xsq = block_sum(xsq, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = ssq;
    partials[2 * blockIdx.x + 1] = xsq;
  }
}

__global__ void __launch_bounds__(256) k_update_cameras(DeviceProblem D, ParamSet cur, ParamSet cand, WorkArrays W,
                                                         double* __restrict__ partials) {
  __shared__ double red[32];
  const int b = blockIdx.x * 256 + threadIdx.x;
  const int cb = D.cb;
  double ssq = 0.0, xsq = 0.0;
  if (b < D.n_ext) {
    const double free_ = cur.pose_rows[b].free_;
    for (int k = 0; k < 3; ++k) {
      double d = 0.0, dt = 0.0;
      if (cb >= 6) {
        d = W.sc[b * cb + k] * W.x[b * cb + k] * free_;
        dt = W.sc[b * cb + 3 + k] * W.x[b * cb + 3 + k] * free_;
      }
      const double w = cur.ext_rot[3 * b + k], t = cur.ext_trans[3 * b + k];
      cand.ext_rot[3 * b + k] = w + d;
      cand.ext_trans[3 * b + k] = t + dt;
      ssq += d * d + dt * dt;
      xsq += free_ * (w * w + t * t);
    }
  }
  if (b < D.n_intr) {
    double f0 = cur.focal[2 * b], f1 = cur.focal[2 * b + 1], k0 = cur.dist[2 * b], k1 = cur.dist[2 * b + 1];
    if (cb == 9) {
      const double df = W.sc[b * cb + 6] * W.x[b * cb + 6];
      const double dk0 = W.sc[b * cb + 7] * W.x[b * cb + 7];
      const double dk1 = W.sc[b * cb + 8] * W.x[b * cb + 8];
      ssq += df * df + dk0 * dk0 + dk1 * dk1;
      xsq += f0 * f0 + k0 * k0 + k1 * k1;
      f0 += df;
      k0 += dk0;
      k1 += dk1;
    }
    cand.focal[2 * b] = f0;
    cand.focal[2 * b + 1] = f1;
    cand.dist[2 * b] = k0;
    cand.dist[2 * b + 1] = k1;
  }
  ssq = block_sum(ssq, red);
  xsq = block_sum(xsq, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = ssq;
    partials[2 * blockIdx.x + 1] = xsq;
  }
}

// ------------------------------------------------- filterPoint3d decisions (DeepArcManager.cc:331-424)
// One thread per point over its (point-sorted) observations: an observation goes when its mse is
// below the boundary, a point goes when nothing is left of it or when it lies outside the
// hemisphere, and then takes its remaining observations along.  NaN compares false, as on the host.
__global__ void __launch_bounds__(256) k_filter_flags(DeviceProblem D, const double* __restrict__ pts,
                                                       const double* __restrict__ mse, double boundary, int use_sphere,
                                                       double cx, double cy, double cz, double half_rho,
                                                       uint8_t* __restrict__ obs_remove, uint8_t* __restrict__ pt_remove) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= D.n_pts) return;
  const int a = D.pt_first[pt], b = D.pt_first[pt + 1];
  DBA_CHECK(a >= 0 && a <= b && b <= D.n_obs);
  int kept = 0;
  for (int o = a; o < b; ++o) kept += (mse[o] < boundary) ? 0 : 1;
  bool gone = kept == 0;
  if (!gone && use_sphere) {
    const double dx = pts[3 * static_cast<int64_t>(pt)] - cx, dy = pts[3 * static_cast<int64_t>(pt) + 1] - cy,
                 dz = pts[3 * static_cast<int64_t>(pt) + 2] - cz;
    // summed in the reference's order: ((dx^2 + dy^2) + dz^2)
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    gone = d2 > half_rho;
  }
  pt_remove[pt] = gone ? 1 : 0;
  for (int o = a; o < b; ++o) obs_remove[o] = (gone || mse[o] < boundary) ? 1 : 0;
}

// ------------------------------------------------------------------------ reductions
// Deterministic single-CTA reductions of per-CTA partials.
__global__ void __launch_bounds__(1024) k_reduce_sum(const double* __restrict__ in, int n, int stride, int offset,
                                                      double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += in[static_cast<int64_t>(i) * stride + offset];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *out = acc;
}
__global__ void __launch_bounds__(1024) k_reduce_max(const double* __restrict__ in, int n, int stride, int offset,
                                                      double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fmax(acc, in[static_cast<int64_t>(i) * stride + offset]);
  acc = block_max(acc, red);
  if (threadIdx.x == 0) *out = acc;
}

// several of those in one launch: CTA j runs job j (same summation order as the single kernels)
__global__ void __launch_bounds__(1024) k_reduce_multi(ReduceJobs J) {
  __shared__ double red[32];
  const ReduceJob job = J.job[blockIdx.x];
  double acc = 0.0;
  if (job.is_max) {
    for (int i = threadIdx.x; i < job.n; i += blockDim.x) acc = fmax(acc, job.in[static_cast<int64_t>(i) * job.stride + job.offset]);
    acc = block_max(acc, red);
  } else {
    for (int i = threadIdx.x; i < job.n; i += blockDim.x) acc += job.in[static_cast<int64_t>(i) * job.stride + job.offset];
    acc = block_sum(acc, red);
  }
  if (threadIdx.x == 0) *job.out = acc;
}

}  // namespace

// ================================================================== launch wrappers
// true the first time it is called on the current device (function attributes such as the dynamic
// shared-memory limit are per device; a process may hold handles on several devices)
static bool first_use_on_device(unsigned long long* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

int tile_grid(const DeviceProblem& D) { return D.n_tiles; }
int cost_grid(const DeviceProblem& D) { return static_cast<int>((D.n_obs + 255) / 256); }

void launch_pose_rows(const ParamSet& P, const uint8_t* ext_const, int freeze_all, int n_ext, int n_intr,
                      cudaStream_t st) {
  const int n = n_ext > n_intr ? n_ext : n_intr;
  if (n <= 0) return;
  k_pose_rows<<<(n + 127) / 128, 128, 0, st>>>(P, ext_const, freeze_all, n_ext, n_intr);
}

// resident 512-thread CTAs per SM for the per-tile kernels of the Schur front/back half (DBA_TILE_MINB = 2 | 3)
static int tile_minb() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("DBA_TILE_MINB");
    v = e ? std::atoi(e) : 3;
    if (v != 2) v = 3;
  }
  return v;
}
// resident 256-thread CTAs per SM asked of the register allocator for the Jacobian kernels (2 = 128
// registers, 3 = 80, 4 = 64 with more spills).  Measured on bal5m: tiled 403 / 352 / 382 us, flat
// 424 / 400 / 385 us (DBA_JAC_MINB; DBA_JAC=flat selects the flat kernel)
static int jac_minb() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("DBA_JAC_MINB");
    v = e ? std::atoi(e) : 3;
    if (v < 2 || v > 4) v = 3;
  }
  return v;
}
// Tiled (camera rows of the tile staged in shared memory) or flat (one thread per observation, gathers from the
// global tables) Jacobian kernel.  The staging pays when a warp's 32 observations would hit 32 different rows of a
// table that does not fit L1; with a few dozen pose blocks (the reference's rigs: 19) the tables ARE L1-resident
// and the flat kernel wins (arc1m: 90 vs 115 us).  DBA_JAC=flat|tiled forces either.
static bool jac_tiled(const DeviceProblem& D) {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("DBA_JAC");
    v = (e && std::strcmp(e, "flat") == 0) ? 0 : ((e && std::strcmp(e, "tiled") == 0) ? 1 : 2);
  }
  if (!D.obs_lc) return false;
  return v == 2 ? D.n_ext > 64 : v != 0;
}
int jacobian_partials(const DeviceProblem& D) { return jac_tiled(D) ? D.n_tiles : cost_grid(D); }

// resident 256-thread CTAs per SM of the per-tile kernels on 256-observation tiles (the reference's rigs):
// DBA_TILE256_MINB = 1 (whatever the register allocator takes: 2-3 CTAs) | 4 (64 registers) | 6 (40).
// Measured on arc1m: k_point_prepare 69 / 43 / 37 us, k_back_substitute 62 / 61 / 67 us -> defaults 6 and 4.
static int tile256_minb(int dflt) {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("DBA_TILE256_MINB");
    v = e ? std::atoi(e) : 0;
    if (v != 1 && v != 4 && v != 6) v = 0;
  }
  return v ? v : dflt;
}
// the variant without camera planes (points-only solves, plane-less matrix-free solves) is light enough for more
// resident CTAs, and the kernel is bound by waves x depth of its load chain (DBA_JAC0_MINB = 3..8; measured on
// bal5m: 3: 206 us, 4: 168 us)
static int jac0_minb() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("DBA_JAC0_MINB");
    v = e ? std::atoi(e) : 4;
    if (v != 3 && v != 4 && v != 5 && v != 6 && v != 8) v = 4;
  }
  return v;
}
template <int CB, bool TWO>
static void launch_jacobian_t(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, int unit_scale, double* partial_cost,
                              cudaStream_t st) {
  if constexpr (CB == 0) {
    if (jac_tiled(D)) {
      switch (jac0_minb()) {
        case 3: k_jacobian_tile<0, false, 3><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost); return;
        case 5: k_jacobian_tile<0, false, 5><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost); return;
        case 6: k_jacobian_tile<0, false, 6><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost); return;
        case 8: k_jacobian_tile<0, false, 8><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost); return;
        default: k_jacobian_tile<0, false, 4><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost); return;
      }
    }
  }
  if (jac_tiled(D)) {
    const int mb = jac_minb();
    if (mb == 2) k_jacobian_tile<CB, TWO, 2><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost);
    else if (mb == 3) k_jacobian_tile<CB, TWO, 3><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost);
    else k_jacobian_tile<CB, TWO, 4><<<D.n_tiles, kJacThreads, 0, st>>>(D, P, W, unit_scale, D.intr_is_pose, partial_cost);
    return;
  }
  const int grid = cost_grid(D);
  const int mb = jac_minb();
  if (mb == 2) k_jacobian<CB, TWO, 2><<<grid, 256, 0, st>>>(D, P, W, unit_scale, partial_cost);
  else if (mb == 3) k_jacobian<CB, TWO, 3><<<grid, 256, 0, st>>>(D, P, W, unit_scale, partial_cost);
  else k_jacobian<CB, TWO, 4><<<grid, 256, 0, st>>>(D, P, W, unit_scale, partial_cost);
}
void launch_jacobian(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, int cb_store, int store_two,
                     int unit_scale, double* partial_cost, cudaStream_t st) {
  if (D.n_obs == 0) return;
  if (cb_store == 0)
    launch_jacobian_t<0, false>(D, P, W, unit_scale, partial_cost, st);
  else if (cb_store == 6 && !store_two)
    launch_jacobian_t<6, false>(D, P, W, unit_scale, partial_cost, st);
  else if (cb_store == 6 && store_two)
    launch_jacobian_t<6, true>(D, P, W, unit_scale, partial_cost, st);
  else if (cb_store == 9 && !store_two)
    launch_jacobian_t<9, false>(D, P, W, unit_scale, partial_cost, st);
  else
    launch_jacobian_t<9, true>(D, P, W, unit_scale, partial_cost, st);
}

void launch_filter_flags(const DeviceProblem& D, const ParamSet& P, const double* mse, double boundary, const double* centre,
                         double rho, uint8_t* obs_remove, uint8_t* pt_remove, cudaStream_t st) {
  if (D.n_pts == 0) return;
  k_filter_flags<<<(D.n_pts + 255) / 256, 256, 0, st>>>(D, P.pts, mse, boundary, centre ? 1 : 0, centre ? centre[0] : 0.0,
                                                         centre ? centre[1] : 0.0, centre ? centre[2] : 0.0, rho / 2, obs_remove,
                                                         pt_remove);
}

void launch_cost(const DeviceProblem& D, const ParamSet& P, double* partial_cost, double* mse_out, cudaStream_t st) {
  if (D.n_obs == 0) return;
  k_cost<<<cost_grid(D), 256, 0, st>>>(D, P, partial_cost, mse_out);
}

void launch_point_prepare(const DeviceProblem& D, const WorkArrays& W, double radius, double min_diag,
                          double max_diag, int mode, double* partials, cudaStream_t st) {
  if (D.n_tiles == 0) return;
  auto go = [&](auto tag, auto mb) {
    constexpr int T = decltype(tag)::value;
    constexpr int MB = decltype(mb)::value;
    constexpr size_t smem = 9 * T * sizeof(double);
    static unsigned long long configured = 0;  // per device: function attributes are
    if (first_use_on_device(&configured)) {
      cudaFuncSetAttribute(k_point_prepare<T, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    }
    k_point_prepare<T, MB><<<D.n_tiles, T, smem, st>>>(D, W, radius, min_diag, max_diag, mode, partials);
  };
  using std::integral_constant;
  if (D.tile == 256) {
    const int mb = tile256_minb(6);
    if (mb == 6) go(integral_constant<int, 256>(), integral_constant<int, 6>());
    else if (mb == 4) go(integral_constant<int, 256>(), integral_constant<int, 4>());
    else go(integral_constant<int, 256>(), integral_constant<int, 1>());
  }
  else if (D.tile == 512) {
    static int prep4 = -1;  // four resident 512-thread CTAs (32 registers, 120 B of spills): 131 vs 139 us on bal5m (DBA_PREP_MINB=3 restores three)
    if (prep4 < 0) {
      const char* e = std::getenv("DBA_PREP_MINB");
      prep4 = (e && std::atoi(e) != 4) ? 0 : 1;
    }
    if (prep4) go(integral_constant<int, 512>(), integral_constant<int, 4>());
    else if (tile_minb() == 3) go(integral_constant<int, 512>(), integral_constant<int, 3>());
    else go(integral_constant<int, 512>(), integral_constant<int, 2>());
  } else go(integral_constant<int, 1024>(), integral_constant<int, 1>());
}

static size_t cam_acc_doubles(const DeviceProblem& D) {
  return static_cast<size_t>(D.n_blocks) * D.cb * D.cb + 3 * static_cast<size_t>(D.n_blocks) * D.cb;
}

template <int CB, int MODE>
static void launch_camera_gather_t(const DeviceProblem& D, const WorkArrays& W, const ParamSet* P, cudaStream_t st) {
  constexpr int NACC = MODE == 0 ? CB : CB * (CB + 1) / 2 + 3 * CB;
  if (D.n_chunks > 0) {
    if (P) k_camera_gather_mf<CB, MODE><<<D.n_chunks, 128, 0, st>>>(D, *P, W);
    else k_camera_gather<CB, MODE><<<D.n_chunks, 128, 0, st>>>(D, W);
  }
  const int n = D.n_blocks * NACC;
  k_camera_combine<CB, MODE><<<(n + 127) / 128, 128, 0, st>>>(D, W);
}
// recompute != nullptr: the Jacobian is recomputed from these parameters instead of read from the planes
// (single-pose problems only)
void launch_camera_gather(const DeviceProblem& D, const WorkArrays& W, int mode, cudaStream_t st, const ParamSet* recompute) {
  if (D.cb == 0 || D.n_blocks == 0) return;
  const ParamSet* P = (recompute && !D.two) ? recompute : nullptr;
  if (mode == 0) cudaMemsetAsync(W.cam_acc, 0, cam_acc_doubles(D) * sizeof(double), st);  // mode 0 fills diagF only
  if (D.cb == 6) {
    if (mode == 0) launch_camera_gather_t<6, 0>(D, W, P, st);
    else if (mode == 1) launch_camera_gather_t<6, 1>(D, W, P, st);
    else launch_camera_gather_t<6, 2>(D, W, P, st);
  } else {
    if (mode == 0) launch_camera_gather_t<9, 0>(D, W, P, st);
    else if (mode == 1) launch_camera_gather_t<9, 1>(D, W, P, st);
    else launch_camera_gather_t<9, 2>(D, W, P, st);
  }
}

void launch_camera_scales(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st) {
  const int n = D.n_blocks * D.cb;
  if (n == 0) return;
  k_camera_scales<<<(n + 255) / 256, 256, 0, st>>>(D, W);
}

int camera_finalize_grid(const DeviceProblem& D) {
  const int per_cta = kFinThreads / (D.cb > 0 ? D.cb : 6);
  return (D.n_blocks + per_cta - 1) / per_cta;
}

void launch_camera_finalize(const DeviceProblem& D, const WorkArrays& W, double radius, double min_diag,
                            double max_diag, double* partials, int with_minv, cudaStream_t st) {
  if (D.cb == 0 || D.n_blocks == 0) return;
  const int grid = camera_finalize_grid(D);
  if (D.cb == 6) {
    if (with_minv) k_camera_finalize<6, true><<<grid, kFinThreads, 0, st>>>(D, W, radius, min_diag, max_diag, partials);
    else k_camera_finalize<6, false><<<grid, kFinThreads, 0, st>>>(D, W, radius, min_diag, max_diag, partials);
  } else {
    if (with_minv) k_camera_finalize<9, true><<<grid, kFinThreads, 0, st>>>(D, W, radius, min_diag, max_diag, partials);
    else k_camera_finalize<9, false><<<grid, kFinThreads, 0, st>>>(D, W, radius, min_diag, max_diag, partials);
  }
}

void launch_pcg_init(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st) {
  const int n = D.n_blocks * D.cb;
  k_pcg_init<<<(n + 255) / 256, 256, 0, st>>>(D, W);
}

template <int CB, bool TWO, int T>
static void launch_spmv_tile_tt(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st) {
  static unsigned long long configured = 0;  // per device: function attributes are
  constexpr size_t smem = SpmvSmem<CB, TWO, T>::kBytes;
  static_assert(smem <= 227 * 1024, "tile does not fit in shared memory");
  if (first_use_on_device(&configured)) {
    cudaFuncSetAttribute(k_spmv_tile<CB, TWO, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  k_spmv_tile<CB, TWO, T><<<D.n_tiles, T, smem, st>>>(D, W);
}
template <int CB, bool TWO>
static void launch_spmv_tile_t(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st) {
  if (D.tile == 256) launch_spmv_tile_tt<CB, TWO, 256>(D, W, st);
  else if (D.tile == 512) launch_spmv_tile_tt<CB, TWO, 512>(D, W, st);
  else if constexpr (!TWO) launch_spmv_tile_tt<CB, TWO, 1024>(D, W, st);  // two-pose tiles of 1024 exceed 227 KB
}

void launch_spmv_tile(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st) {
  if (D.n_tiles == 0) return;
  if (D.cb == 6 && !D.two)
    launch_spmv_tile_t<6, false>(D, W, st);
  else if (D.cb == 6)
    launch_spmv_tile_t<6, true>(D, W, st);
  else
    launch_spmv_tile_t<9, false>(D, W, st);
}

void launch_partials_to_q(const DeviceProblem& D, const WorkArrays& W, int fuse_dot, int n_split, int mf, cudaStream_t st) {
  if (D.n_blocks == 0) return;
  const dim3 grid(D.n_blocks, n_split > 1 ? n_split : 1);
  if (n_split > 1) fuse_dot = 0;
  if (D.cb == 6) {
    if (mf) k_partials_to_q<6, true><<<grid, 128, 0, st>>>(D, W, fuse_dot, n_split);
    else k_partials_to_q<6, false><<<grid, 128, 0, st>>>(D, W, fuse_dot, n_split);
  } else {
    if (mf) k_partials_to_q<9, true><<<grid, 128, 0, st>>>(D, W, fuse_dot, n_split);
    else k_partials_to_q<9, false><<<grid, 128, 0, st>>>(D, W, fuse_dot, n_split);
  }
}

template <int CB, bool MF>
static int launch_pcg_fused_t(const DeviceProblem& D, const WorkArrays& W, double tol2, int min_iter, int n_split,
                              const PeerWin& pw, cudaStream_t st) {
  static int resident = 0;
  if (resident == 0) {
    int dev = 0, n_sm = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_fused<CB, MF>, kFusedThreads, 0);
    resident = std::max(1, n_sm * per_sm);
  }
  // one warp per (camera block, slice)
  const int grid = std::max(1, std::min((D.n_blocks * std::max(n_split, 1) + kFusedThreads / 32 - 1) / (kFusedThreads / 32), resident));
  DeviceProblem d = D;
  WorkArrays w = W;
  PeerWin win = pw;
  void* args[] = {&d, &w, &tol2, &min_iter, &n_split, &win};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_pcg_fused<CB, MF>), dim3(grid), dim3(kFusedThreads), args, 0, st) == cudaSuccess ? 0 : -1;
}
int launch_pcg_fused(const DeviceProblem& D, const WorkArrays& W, int mf, double tol2, int min_iter, int n_split,
                     const PeerWin& pw, cudaStream_t st) {
  if (D.n_blocks == 0) return 0;
  if (D.cb == 6)
    return mf ? launch_pcg_fused_t<6, true>(D, W, tol2, min_iter, n_split, pw, st)
              : launch_pcg_fused_t<6, false>(D, W, tol2, min_iter, n_split, pw, st);
  return mf ? launch_pcg_fused_t<9, true>(D, W, tol2, min_iter, n_split, pw, st)
            : launch_pcg_fused_t<9, false>(D, W, tol2, min_iter, n_split, pw, st);
}

void launch_mf_rows(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, cudaStream_t st) {
  if (D.cb == 0 || D.n_blocks == 0) return;
  const int grid = (D.n_blocks + 127) / 128;
  if (D.cb == 6) k_mf_rows<6><<<grid, 128, 0, st>>>(D, P, W);
  else k_mf_rows<9><<<grid, 128, 0, st>>>(D, P, W);
}

void launch_mf_direction(const DeviceProblem& D, const WorkArrays& W, int init, cudaStream_t st) {
  if (D.cb == 0 || D.n_blocks == 0) return;
  const int grid = (D.n_blocks + 127) / 128;
  if (D.cb == 6) k_mf_direction<6><<<grid, 128, 0, st>>>(D, W, init);
  else k_mf_direction<9><<<grid, 128, 0, st>>>(D, W, init);
}

template <int CB, bool TWO, int T, int MINB>
static int launch_spmv_mf_tt(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, const MfTail& tail, cudaStream_t st) {
  static unsigned long long configured = 0;  // per device: function attributes are
  constexpr size_t smem = MfSmem<CB, TWO, T>::kBytes;
  static_assert(smem <= 227 * 1024, "tile does not fit in shared memory");
  if (first_use_on_device(&configured)) {
    cudaFuncSetAttribute(k_spmv_mf<CB, TWO, T, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  int resident = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_spmv_mf<CB, TWO, T, MINB>, T, smem);
  const int grid = std::min(D.n_tiles, n_sm * std::max(resident, 1));
  if (tail.fuse) {
    // the grid is exactly the co-resident CTAs, so it qualifies for a cooperative launch (grid.sync)
    DeviceProblem d = D;
    WorkArrays w = W;
    const double* pts = P.pts;
    const IntrRow* ir = P.intr_rows;
    int fuse = std::max(tail.fuse, 1), min_iter = tail.min_iter;  // the flag doubles as the slice count of the tail
    double tol2 = tail.tol2;
    PeerWin pw = tail.pw;
    void* args[] = {&d, &w, &pts, &ir, &fuse, &tol2, &min_iter, &pw};
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_spmv_mf<CB, TWO, T, MINB>), dim3(grid), dim3(T), args, smem, st) == cudaSuccess ? 0 : -1;
  }
  k_spmv_mf<CB, TWO, T, MINB><<<grid, T, smem, st>>>(D, W, P.pts, P.intr_rows, 0, 0.0, 0, PeerWin{});
  return 0;
}
// resident CTAs per SM requested from the register allocator for 256-thread tiles (tuning knob DBA_MF_MINB)
static int mf_minb() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("DBA_MF_MINB");
    v = e ? std::atoi(e) : 3;
    if (v < 2 || v > 4) v = 3;
  }
  return v;
}
template <int CB, bool TWO>
static int launch_spmv_mf_t(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, const MfTail& tail, cudaStream_t st) {
  if (D.tile == 256) {
    const int mb = mf_minb();
    if (mb == 2) return launch_spmv_mf_tt<CB, TWO, 256, 2>(D, P, W, tail, st);
    if (mb == 3) return launch_spmv_mf_tt<CB, TWO, 256, 3>(D, P, W, tail, st);
    return launch_spmv_mf_tt<CB, TWO, 256, 4>(D, P, W, tail, st);
  } else if (D.tile == 512) {
    return launch_spmv_mf_tt<CB, TWO, 512, TWO ? 1 : 2>(D, P, W, tail, st);
  } else if constexpr (!TWO) {
    return launch_spmv_mf_tt<CB, TWO, 1024, 1>(D, P, W, tail, st);
  }
  return -1;
}

int launch_spmv_mf(const DeviceProblem& D, const ParamSet& P, const WorkArrays& W, const MfTail& tail, cudaStream_t st) {
  if (D.n_tiles == 0) return tail.fuse ? -1 : 0;
  if (D.cb == 6 && !D.two) return launch_spmv_mf_t<6, false>(D, P, W, tail, st);
  if (D.cb == 6) return launch_spmv_mf_t<6, true>(D, P, W, tail, st);
  return launch_spmv_mf_t<9, false>(D, P, W, tail, st);
}

// q = sum of the slices (multi-GPU with n_split > 1: input of the allreduce)
__global__ void __launch_bounds__(256) k_fold_q(DeviceProblem D, WorkArrays W, int n_split) {
  if (W.pcg_state[1]) return;
  const int n = D.n_blocks * D.cb;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double q = 0.0;
    for (int sl = 0; sl < n_split; ++sl) q += W.q_split[static_cast<int64_t>(sl) * n + i];
    W.q[i] = q;
  }
}
void launch_fold_q(const DeviceProblem& D, const WorkArrays& W, int n_split, cudaStream_t st) {
  const int n = D.n_blocks * D.cb;
  k_fold_q<<<(n + 255) / 256, 256, 0, st>>>(D, W, n_split);
}

void launch_pcg_dot(const DeviceProblem& D, const WorkArrays& W, int n_split, cudaStream_t st) {
  const int n = D.n_blocks * D.cb;
  k_pcg_dot<<<(n + 255) / 256, 256, 0, st>>>(D, W, n_split);
}
void launch_pcg_step(const DeviceProblem& D, const WorkArrays& W, double tol2, int min_iter, cudaStream_t st) {
  const int blocks_per_cta = 256 / D.cb;
  k_pcg_step<<<(D.n_blocks + blocks_per_cta - 1) / blocks_per_cta, 256, 0, st>>>(D, W, tol2, min_iter);
}
void launch_pcg_direction(const DeviceProblem& D, const WorkArrays& W, cudaStream_t st) {
  const int n = D.n_blocks * D.cb;
  k_pcg_direction<<<(n + 255) / 256, 256, 0, st>>>(D, W);
}

template <int CB, bool TWO>
static void launch_back_substitute_t(const DeviceProblem& D, const WorkArrays& W, double* partial_model, cudaStream_t st) {
  auto go = [&](auto tag, auto mb) {
    constexpr int T = decltype(tag)::value;
    constexpr int MB = decltype(mb)::value;
    constexpr size_t smem = 6 * T * sizeof(double);
    static unsigned long long configured = 0;  // per device: function attributes are
    if (first_use_on_device(&configured)) {
      cudaFuncSetAttribute(k_back_substitute<CB, TWO, T, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    }
    k_back_substitute<CB, TWO, T, MB><<<D.n_tiles, T, smem, st>>>(D, W, partial_model);
  };
  using std::integral_constant;
  if (D.tile == 256) {
    const int mb = tile256_minb(4);
    if (mb == 6) go(integral_constant<int, 256>(), integral_constant<int, 6>());
    else if (mb == 4) go(integral_constant<int, 256>(), integral_constant<int, 4>());
    else go(integral_constant<int, 256>(), integral_constant<int, 1>());
  }
  else if (D.tile == 512) {
    if (tile_minb() == 3) go(integral_constant<int, 512>(), integral_constant<int, 3>());
    else go(integral_constant<int, 512>(), integral_constant<int, 2>());
  } else go(integral_constant<int, 1024>(), integral_constant<int, 1>());
}

template <int CB>
static void launch_back_substitute_mf_t(const DeviceProblem& D, const WorkArrays& W, const ParamSet& P, double* partial_model,
                                        cudaStream_t st) {
  k_mf_step<CB><<<(D.n_blocks + 127) / 128, 128, 0, st>>>(D, W);
  auto go = [&](auto tag, auto mb) {
    constexpr int T = decltype(tag)::value;
    constexpr int MB = decltype(mb)::value;
    constexpr size_t smem = 16 * T * sizeof(double);
    static unsigned long long configured = 0;  // per device: function attributes are
    if (first_use_on_device(&configured)) {
      cudaFuncSetAttribute(k_back_substitute_mf<CB, T, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    }
    k_back_substitute_mf<CB, T, MB><<<D.n_tiles, T, smem, st>>>(D, W, P.pts, P.intr_rows, partial_model);
  };
  using std::integral_constant;
  if (D.tile == 256) go(integral_constant<int, 256>(), integral_constant<int, 2>());
  else if (D.tile == 512) go(integral_constant<int, 512>(), integral_constant<int, 2>());
  else go(integral_constant<int, 1024>(), integral_constant<int, 1>());
}

// recompute != nullptr (single-pose, matrix-free solves): u = F x from the camera rows instead of the camera planes
void launch_back_substitute(const DeviceProblem& D, const WorkArrays& W, double* partial_model, cudaStream_t st,
                            const ParamSet* recompute) {
  if (D.n_tiles == 0) return;
  if (recompute && D.cb > 0 && !D.two) {
    if (D.cb == 6) launch_back_substitute_mf_t<6>(D, W, *recompute, partial_model, st);
    else launch_back_substitute_mf_t<9>(D, W, *recompute, partial_model, st);
    return;
  }
  if (D.cb == 0)
    launch_back_substitute_t<0, false>(D, W, partial_model, st);
  else if (D.cb == 6 && !D.two)
    launch_back_substitute_t<6, false>(D, W, partial_model, st);
  else if (D.cb == 6)
    launch_back_substitute_t<6, true>(D, W, partial_model, st);
  else
    launch_back_substitute_t<9, false>(D, W, partial_model, st);
}

int update_points_grid(const DeviceProblem& D) { return static_cast<int>((3 * static_cast<int64_t>(D.n_pts) + 255) / 256); }
int update_cameras_grid(const DeviceProblem& D) {
  const int n = D.n_ext > D.n_intr ? D.n_ext : D.n_intr;
  return (n + 255) / 256;
}

void launch_update_points(const DeviceProblem& D, const ParamSet& cur, const ParamSet& cand, const WorkArrays& W,
                          double* partials, cudaStream_t st) {
  const int64_t n3 = 3 * static_cast<int64_t>(D.n_pts);
  if (n3 == 0) return;
  k_update_points<<<update_points_grid(D), 256, 0, st>>>(n3, cur.pts, cand.pts, W.sp, W.dp, partials);
}

void launch_update_cameras(const DeviceProblem& D, const ParamSet& cur, const ParamSet& cand, const WorkArrays& W,
                           double* partials, cudaStream_t st) {
  const int g = update_cameras_grid(D);
  if (g == 0) return;
  k_update_cameras<<<g, 256, 0, st>>>(D, cur, cand, W, partials);
}

void launch_reduce_sum(const double* partials, int n, int stride, int offset, double* out, cudaStream_t st) {
  k_reduce_sum<<<1, 1024, 0, st>>>(partials, n, stride, offset, out);
}
void launch_reduce_max(const double* partials, int n, int stride, int offset, double* out, cudaStream_t st) {
  k_reduce_max<<<1, 1024, 0, st>>>(partials, n, stride, offset, out);
}
void launch_reduce_multi(const ReduceJobs& jobs, cudaStream_t st) {
  if (jobs.count > 0) k_reduce_multi<<<jobs.count, 1024, 0, st>>>(jobs);
}

}  // namespace dba
