// Explicit reduced camera system + on-device Cholesky: the reference's DENSE_SCHUR linear solver
// (reference src/sfm.cc:67, :95; upstream Ceres schur_eliminator_impl.h + DenseSchurComplementSolver,
// restated on the CPU in oracle/mini_ceres.cc::EliminateAndSolve) for problems whose reduced system is
// small — the reference's own rigs have 19..50 pose blocks, i.e. 114..300 reduced unknowns.
//
//   S = F^T F + D_c^2 - sum_points Y_i C_i^-1 Y_i^T,   Y_i[A] = sum_{o in i, o uses A} F_{o,A}^T E_o
//
// k_dense_z      (K4a)  one CTA per batch of points: the compact list of camera blocks each point touches
//                       and Z_i[A] = Y_i[A] L_i with C_i^-1 = L_i L_i^T, written to a global image
// k_dense_pairs  (K4b)  one THREAD per camera-block pair (A <= B) keeps the CB x CB block S_AB in
//                       registers and walks a fixed slice of the batches: the elimination term of a
//                       pair is the 3-deep product Z_i[A] Z_i[B]^T.  No atomics: every pair is owned by
//                       one thread, slices are summed in slice order.
// k_pair_gather         F_A^T F_B of composed (arc, ring) observations, camera-pair-sorted chunks
//                       (the diagonal blocks F_A^T F_A come from k_camera_gather, mode 2)
// k_dense_combine       fixed-order sum of all of it into the dense n x n matrix (both triangles)
// k_dense_cholesky (K9) one CTA: blocked left-looking Cholesky of S + D_c^2 with the right-hand side
//                       carried as an extra row (so L^-1 rhs falls out of the factorisation), then the
//                       blocked back substitution; writes the camera step W.x.  A non-positive pivot
//                       makes the step NaN, which the LM driver treats as a linear-solver failure
//                       exactly like the CPU oracle does.
#include <cstdio>

#include "ba_kernels.cuh"

namespace dba {

namespace {

constexpr int kDnThreads = 256;  // pair threads per CTA == pairs per pair group
constexpr int kDnIncObs = 512;   // observations of a batch that the incidence path of k_dense_z stages (else: one thread per entry)

// packed upper-triangular pair index -> (A, B), A <= B, row-major over A
__device__ __forceinline__ void pair_decode(int q, int nb, int& A, int& B) {
  int a = 0, rem = q;
  while (rem >= nb - a) {
    rem -= nb - a;
    ++a;
  }
  A = a;
  B = a + rem;
}

// Shared-memory plan of k_dense_z: per-point segments, entry capacities, the compact entry list and the
// per-point block lookup of one batch.
struct DzSmem {
  static constexpr size_t oL = 0;                                               // [points][6] L with C^-1 = L L^T
  static constexpr size_t oBase = oL + sizeof(double) * kDnPtsCap * 6;          // capacity-based entry base of each point
  static constexpr size_t oSeg = oBase + sizeof(int) * (kDnPtsCap + 4);         // first observation of each point
  static constexpr size_t oCnt = oSeg + sizeof(int) * (kDnPtsCap + 4);          // distinct blocks of each point, then their exclusive scan
  static constexpr size_t oEnt = oCnt + sizeof(int) * (kDnPtsCap + 4);          // entry -> (point << 16 | block), 0xffffffff = unused
  static constexpr size_t oMul = oEnt + sizeof(unsigned int) * kDnEntCap;       // incidences seen so far of each entry
  static constexpr size_t oInc = oMul + sizeof(int) * kDnEntCap;                // (observation, slot) -> entry | rank << 16
  static constexpr size_t oAB = oInc + sizeof(unsigned int) * 2 * kDnIncObs;    // (block a, block b) of the batch's observations
  static constexpr size_t oLook = oAB + sizeof(int2) * kDnIncObs;               // [point][block] -> entry index + 1 within the point
  __host__ __device__ static size_t look_bytes(int nb) { return (sizeof(unsigned short) * kDnPtsCap * ((nb + 1) & ~1) + 15) / 16 * 16; }
  // + the Z tile [kDnEntCap][3 CB + 1] of the incidence path behind the lookup
  static size_t bytes(int nb, int cb) { return oLook + look_bytes(nb) + sizeof(double) * kDnEntCap * (3 * cb + 1); }
};
template <int CB>
__host__ __device__ constexpr int dn_zs() { return 3 * CB; }  // doubles per entry of the global Z image

// K4a: one CTA per batch of points.  Compact list of the camera blocks each point touches, Y = sum F^T E over the
// point's observations that use the block, Z = Y L with C^-1 = L L^T — written to the global Z image of the batch,
// entries compacted (Q.Z, Q.Zent, Q.Zcount).  One thread per INCIDENCE (observation, pose slot): consecutive lanes
// read consecutive observations of the planes, form (F^T E) L in registers and add it to the entry's row of a
// shared-memory Z tile in rounds — round r takes the r-th incidence of every entry, so no two threads meet in a
// row and each row is summed in observation order (bit-reproducible).  (One thread per entry scanning its point's
// observations, as the fused kernel of the first version did, spends its time on serial scattered plane loads:
// 393 us on arc1m; it remains the path for batches of more than kDnIncObs observations.)
template <int CB>
__global__ void __launch_bounds__(kDnThreads) k_dense_z(DeviceProblem D, WorkArrays W, DenseWork Q) {
  using L = DzSmem;
  constexpr int ZS = dn_zs<CB>();
  extern __shared__ __align__(16) unsigned char smem_dz[];
  double* sL = reinterpret_cast<double*>(smem_dz + L::oL);
  int* sBase = reinterpret_cast<int*>(smem_dz + L::oBase);
  int* sSeg = reinterpret_cast<int*>(smem_dz + L::oSeg);
  int* sCnt = reinterpret_cast<int*>(smem_dz + L::oCnt);
  unsigned int* sEnt = reinterpret_cast<unsigned int*>(smem_dz + L::oEnt);
  int* sMul = reinterpret_cast<int*>(smem_dz + L::oMul);
  unsigned int* sInc = reinterpret_cast<unsigned int*>(smem_dz + L::oInc);
  int2* sAB = reinterpret_cast<int2*>(smem_dz + L::oAB);
  unsigned short* sLook = reinterpret_cast<unsigned short*>(smem_dz + L::oLook);
  __shared__ int s_rounds;
  const int tid = threadIdx.x;
  const int nb = D.n_blocks, nbs = (nb + 1) & ~1;
  constexpr int ZT = 3 * CB + 1;  // odd row stride of the Z tile
  double* sZt = reinterpret_cast<double*>(smem_dz + L::oLook + L::look_bytes(nb));
  const int64_t ld = D.ld;
  const int b = blockIdx.x;
  const int p0 = Q.batch_pt[b], np = Q.batch_pt[b + 1] - p0;
  DBA_CHECK(np > 0 && np <= kDnPtsCap && p0 >= 0 && p0 + np <= D.n_pts);
  // ---- the batch's points: segments, entry capacities, cleared lookup
  for (int i = tid; i < (np * nbs) / 2; i += kDnThreads) reinterpret_cast<unsigned int*>(sLook)[i] = 0u;
  for (int i = tid; i < kDnEntCap; i += kDnThreads) sMul[i] = 0;
  for (int i = tid; i < kDnEntCap * ZT; i += kDnThreads) sZt[i] = 0.0;
  if (tid == 0) s_rounds = 0;
  if (tid <= np) sSeg[tid] = D.pt_first[p0 + tid];
  __syncthreads();
  {
    // the batch's (block a, block b) pairs in one coalesced pass: the per-point scan below then runs out of
    // shared memory instead of a chain of dependent global loads per thread
    const int o_first = sSeg[0], nob = sSeg[np] - o_first;
    if (nob <= kDnIncObs)
      for (int i = tid; i < nob; i += kDnThreads) sAB[i] = D.obs_ab[o_first + i];
  }
  if (tid < 32) {
    // exclusive scan of the capacities min(2 k_i, nb) (np <= kDnPtsCap = 2 per lane)
    int c[kDnPtsCap / 32], tot = 0;
#pragma unroll
    for (int j = 0; j < kDnPtsCap / 32; ++j) {
      const int p = tid * (kDnPtsCap / 32) + j;
      c[j] = p < np ? min(2 * (sSeg[p + 1] - sSeg[p]), nb) : 0;
      tot += c[j];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += v;
    }
    int run = incl - tot;
#pragma unroll
    for (int j = 0; j < kDnPtsCap / 32; ++j) {
      const int p = tid * (kDnPtsCap / 32) + j;
      if (p <= np) sBase[p] = run;
      run += c[j];
    }
    if (tid == 31 && np == kDnPtsCap) sBase[np] = run;
  }
  __syncthreads();
  // ---- one thread per point: compact list of the camera blocks it touches; L with C^-1 = L L^T
  if (tid < np) {
    const int base = sBase[tid], cap = sBase[tid + 1] - base;
    DBA_CHECK(base >= 0 && base + cap <= kDnEntCap);
    int cnt = 0, max_rank = 0;
    unsigned short* look = sLook + tid * nbs;
    const int o_first = sSeg[0];
    const bool stage_inc = sSeg[np] - o_first <= kDnIncObs;
    for (int o = sSeg[tid]; o < sSeg[tid + 1]; ++o) {
      const int2 ab = stage_inc ? sAB[o - o_first] : D.obs_ab[o];
#pragma unroll
      for (int slot = 0; slot < 2; ++slot) {
        const int blk = slot ? ab.y : ab.x;
        unsigned int inc = 0xffffffffu;
        if (blk >= 0) {
          if (!look[blk]) {
            DBA_CHECK(blk < nb && cnt < cap);
            look[blk] = static_cast<unsigned short>(++cnt);
            sEnt[base + cnt - 1] = (static_cast<unsigned int>(tid) << 16) | static_cast<unsigned int>(blk);
          }
          const int e = base + look[blk] - 1;
          const int rank = sMul[e]++;  // entries of a point are touched by its own thread only
          max_rank = max(max_rank, rank);
          inc = static_cast<unsigned int>(e) | (static_cast<unsigned int>(rank) << 16);
        }
        if (stage_inc) sInc[2 * (o - o_first) + slot] = inc;
      }
    }
    for (int e = cnt; e < cap; ++e) sEnt[base + e] = 0xffffffffu;
    sCnt[tid] = cnt;
    if (stage_inc) atomicMax(&s_rounds, max_rank + 1);
    const double* ci = W.cinv + 6 * static_cast<int64_t>(p0 + tid);
    double l00 = 0.0, l10 = 0.0, l20 = 0.0, l11 = 0.0, l21 = 0.0, l22 = 0.0;
    if (ci[0] > 0.0) {  // a point whose C was not positive definite carries C^-1 = 0 (flagged by k_point_prepare)
      l00 = sqrt(ci[0]);
      l10 = ci[1] / l00;
      l20 = ci[2] / l00;
      l11 = sqrt(fmax(ci[3] - l10 * l10, 0.0));
      l21 = l11 > 0.0 ? (ci[4] - l20 * l10) / l11 : 0.0;
      l22 = sqrt(fmax(ci[5] - l20 * l20 - l21 * l21, 0.0));
    }
    double* Lp = sL + tid * 6;
    Lp[0] = l00; Lp[1] = l10; Lp[2] = l20; Lp[3] = l11; Lp[4] = l21; Lp[5] = l22;
  }
  __syncthreads();
  // ---- compact entry index of each point: exclusive scan of the distinct-block counts
  if (tid < 32) {
    int c[kDnPtsCap / 32], tot = 0;
#pragma unroll
    for (int j = 0; j < kDnPtsCap / 32; ++j) {
      const int p = tid * (kDnPtsCap / 32) + j;
      c[j] = p < np ? sCnt[p] : 0;
      tot += c[j];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += v;
    }
    int run = incl - tot;
#pragma unroll
    for (int j = 0; j < kDnPtsCap / 32; ++j) {
      const int p = tid * (kDnPtsCap / 32) + j;
      if (p < np) sCnt[p] = run;
      run += c[j];
    }
    if (tid == 31) Q.Zcount[b] = run;  // lane 31 ends with the total
  }
  __syncthreads();
  const int n_ent = sBase[np];
  double* Zb = Q.Z + static_cast<int64_t>(b) * kDnEntCap * ZS;
  unsigned int* Eb = Q.Zent + static_cast<int64_t>(b) * kDnEntCap;
  const int o_first = sSeg[0], n_inc = 2 * (sSeg[np] - o_first);
  if (n_inc <= 2 * kDnIncObs) {
    // ---- one thread per incidence, chunks of kDnThreads in observation order, rounds by rank within the entry
    const int rounds = s_rounds;
    for (int i0 = 0; i0 < n_inc; i0 += kDnThreads) {
      const int i = i0 + tid;
      const unsigned int inc = i < n_inc ? sInc[i] : 0xffffffffu;
      const int e = inc & 0xffffu, rank = inc >> 16;
      double zc[3 * CB];
      if (inc != 0xffffffffu) {
        const int o = o_first + (i >> 1), slot = i & 1;
        const double2* J = D.J + o;
        const double2 e0 = J[(kPlaneJp + 0) * ld], e1 = J[(kPlaneJp + 1) * ld], e2 = J[(kPlaneJp + 2) * ld];
        const int base_pl = slot ? kPlaneJA + CB : kPlaneJA;
        const double* Lp = sL + (sEnt[e] >> 16) * 6;
        const double l00 = Lp[0], l10 = Lp[1], l20 = Lp[2], l11 = Lp[3], l21 = Lp[4], l22 = Lp[5];
#pragma unroll
        for (int k = 0; k < CB; ++k) {
          double y0 = 0.0, y1 = 0.0, y2 = 0.0;
          if (k < 6 || !slot) {
            const double2 F = J[(base_pl + k) * ld];
            y0 = F.x * e0.x + F.y * e0.y;
            y1 = F.x * e1.x + F.y * e1.y;
            y2 = F.x * e2.x + F.y * e2.y;
          }
          zc[3 * k + 0] = y0 * l00 + y1 * l10 + y2 * l20;
          zc[3 * k + 1] = y1 * l11 + y2 * l21;
          zc[3 * k + 2] = y2 * l22;
        }
      }
      for (int r = 0; r < rounds; ++r) {
        if (inc != 0xffffffffu && rank == r) {
          double* z = sZt + e * ZT;
#pragma unroll
          for (int k = 0; k < 3 * CB; ++k) z[k] += zc[k];
        }
        __syncthreads();
      }
    }
    // compacted write-out (the point's entries are the first of its capacity)
    for (int idx = tid; idx < n_ent * ZS; idx += kDnThreads) {
      const int e = idx / ZS, k = idx - e * ZS;
      const unsigned int ent = sEnt[e];
      if (ent == 0xffffffffu) continue;
      const int pt = ent >> 16;
      const int ce = sCnt[pt] + (e - sBase[pt]);
      DBA_CHECK(ce >= 0 && ce < kDnEntCap);
      Zb[static_cast<int64_t>(ce) * ZS + k] = sZt[e * ZT + k];
      if (k == 0) Eb[ce] = ent;
    }
    return;
  }
  // ---- (large batches) one thread per (point, block) entry: Y = sum F^T E over the point's observations that use
  // the block, Z = Y L
  for (int e = tid; e < n_ent; e += kDnThreads) {
    const unsigned int ent = sEnt[e];
    if (ent == 0xffffffffu) continue;
    const int pt = ent >> 16, blk = ent & 0xffffu;
    double y[CB][3];
#pragma unroll
    for (int i = 0; i < CB; ++i) y[i][0] = y[i][1] = y[i][2] = 0.0;
    for (int o = sSeg[pt]; o < sSeg[pt + 1]; ++o) {
      const int2 ab = D.obs_ab[o];
      const bool is_a = ab.x == blk;
      if (!is_a && ab.y != blk) continue;
      const double2* J = D.J + o;
      const double2 e0 = J[(kPlaneJp + 0) * ld], e1 = J[(kPlaneJp + 1) * ld], e2 = J[(kPlaneJp + 2) * ld];
      const int base_pl = is_a ? kPlaneJA : kPlaneJA + CB;
#pragma unroll
      for (int i = 0; i < CB; ++i) {
        if (i < 6 || is_a) {
          const double2 F = J[(base_pl + i) * ld];
          y[i][0] += F.x * e0.x + F.y * e0.y;
          y[i][1] += F.x * e1.x + F.y * e1.y;
          y[i][2] += F.x * e2.x + F.y * e2.y;
        }
      }
    }
    const double* Lp = sL + pt * 6;
    const double l00 = Lp[0], l10 = Lp[1], l20 = Lp[2], l11 = Lp[3], l21 = Lp[4], l22 = Lp[5];
    const int ce = sCnt[pt] + (e - sBase[pt]);  // compact index: the point's entries are the first of its capacity
    DBA_CHECK(ce >= 0 && ce < kDnEntCap);
    double* z = Zb + static_cast<int64_t>(ce) * ZS;
#pragma unroll
    for (int i = 0; i < CB; ++i) {
      z[3 * i + 0] = y[i][0] * l00 + y[i][1] * l10 + y[i][2] * l20;
      z[3 * i + 1] = y[i][1] * l11 + y[i][2] * l21;
      z[3 * i + 2] = y[i][2] * l22;
    }
    Eb[ce] = ent;
  }
}

// Shared-memory plan of k_dense_pairs: the batch's Z entries (odd stride: entries of different blocks on
// different banks), the entry list and the per-point block lookup.
template <int CB>
struct DnSmem {
  static constexpr int ZS = (3 * CB) | 1;
  static constexpr size_t oZ = 0;
  static constexpr size_t oEnt = oZ + sizeof(double) * kDnEntCap * ZS;
  static constexpr size_t oLook = oEnt + sizeof(unsigned int) * kDnEntCap;      // [point][block] -> entry index + 1 within the batch
  static size_t bytes(int nb) { return oLook + sizeof(unsigned short) * kDnPtsCap * ((nb + 1) & ~1); }
};

// K4b: one THREAD per camera-block pair (A <= B) keeps S_AB in registers and walks a fixed slice of the batches:
// per batch the CTA copies the Z entries (one coalesced stream) and rebuilds the block lookup, then
// S_AB -= Z_A Z_B^T over the batch's points.  (F^T F is added by k_dense_combine from the camera gather and the
// camera-pair gather: adding it here would run one lane per matching observation.)
template <int CB, int MINB>
__global__ void __launch_bounds__(kDnThreads, MINB) k_dense_pairs(DeviceProblem D, DenseWork Q) {
  using L = DnSmem<CB>;
  constexpr int ZS = L::ZS, ZG = dn_zs<CB>();
  extern __shared__ __align__(16) unsigned char smem_dn[];
  double* sZ = reinterpret_cast<double*>(smem_dn + L::oZ);
  unsigned int* sEnt = reinterpret_cast<unsigned int*>(smem_dn + L::oEnt);
  unsigned short* sLook = reinterpret_cast<unsigned short*>(smem_dn + L::oLook);
  const int tid = threadIdx.x;
  const int nb = D.n_blocks, nbs = (nb + 1) & ~1;
  const int q = blockIdx.y * kDnThreads + tid;
  const bool has_pair = q < Q.n_pairs;
  int A = 0, B = 0;
  if (has_pair) pair_decode(q, nb, A, B);
  double acc[CB * CB];
#pragma unroll
  for (int k = 0; k < CB * CB; ++k) acc[k] = 0.0;
  const int b0 = static_cast<int>(static_cast<int64_t>(Q.n_batches) * blockIdx.x / gridDim.x);
  const int b1 = static_cast<int>(static_cast<int64_t>(Q.n_batches) * (blockIdx.x + 1) / gridDim.x);
  for (int b = b0; b < b1; ++b) {
    const int np = Q.batch_pt[b + 1] - Q.batch_pt[b];
    const int n_ent = Q.Zcount[b];
    DBA_CHECK(np > 0 && np <= kDnPtsCap && n_ent >= 0 && n_ent <= kDnEntCap);
    const double* Zb = Q.Z + static_cast<int64_t>(b) * kDnEntCap * ZG;
    const unsigned int* Eb = Q.Zent + static_cast<int64_t>(b) * kDnEntCap;
    for (int i = tid; i < (np * nbs) / 2; i += kDnThreads) reinterpret_cast<unsigned int*>(sLook)[i] = 0u;
    for (int i = tid; i < n_ent * ZG; i += kDnThreads) {
      const int e = i / ZG, k = i - e * ZG;
      sZ[e * ZS + k] = Zb[i];
    }
    for (int e = tid; e < n_ent; e += kDnThreads) sEnt[e] = Eb[e];
    __syncthreads();
    for (int e = tid; e < n_ent; e += kDnThreads) {
      const unsigned int ent = sEnt[e];
      DBA_CHECK((ent >> 16) < static_cast<unsigned int>(np) && (ent & 0xffffu) < static_cast<unsigned int>(nb));
      sLook[(ent >> 16) * nbs + (ent & 0xffffu)] = static_cast<unsigned short>(e + 1);
    }
    __syncthreads();
    if (has_pair) {
      for (int p = 0; p < np; ++p) {
        const int ia = sLook[p * nbs + A];
        if (!ia) continue;
        const int ib = (A == B) ? ia : sLook[p * nbs + B];
        if (!ib) continue;
        const double* za = sZ + (ia - 1) * ZS;
        const double* zb = sZ + (ib - 1) * ZS;
        double a[3 * CB];
#pragma unroll
        for (int k = 0; k < 3 * CB; ++k) a[k] = za[k];
#pragma unroll
        for (int j = 0; j < CB; ++j) {
          const double z0 = zb[3 * j], z1 = zb[3 * j + 1], z2 = zb[3 * j + 2];
#pragma unroll
          for (int i = 0; i < CB; ++i)
            acc[i * CB + j] = fma(-a[3 * i + 2], z2, fma(-a[3 * i + 1], z1, fma(-a[3 * i], z0, acc[i * CB + j])));
        }
      }
    }
    __syncthreads();  // the next batch overwrites the staging
  }
  if (has_pair) {
    double* out = Q.S_part + (static_cast<int64_t>(blockIdx.x) * Q.n_pairs + q) * (CB * CB);
#pragma unroll
    for (int k = 0; k < CB * CB; ++k) out[k] = acc[k];
  }
}

// F_lo^T F_hi of the observations that compose the pose blocks (lo, hi), lo < hi: camera-pair-sorted
// incidence in chunks (one CTA per chunk, like k_camera_gather); one row of 36 sums per chunk.
__global__ void __launch_bounds__(128) k_pair_gather(DeviceProblem D, DenseWork Q) {
  __shared__ double red[4][36];
  const int4 ch = Q.pair_chunks[blockIdx.x];  // (pair, first entry, last entry, -)
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
  const int64_t ld = D.ld;
  for (int e = ch.y + threadIdx.x; e < ch.z; e += blockDim.x) {
    const int ent = Q.pair_entries[e];
    const int o = ent >> 1, swap = ent & 1;
    DBA_CHECK(o >= 0 && o < D.n_obs && D.obs_ab[o].y >= 0);  // swap: block b of the observation is the lower-numbered block
    const double2* J = D.J + o;
    const int plo = kPlaneJA + (swap ? 6 : 0), phi = kPlaneJA + (swap ? 0 : 6);
    double2 FH[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) FH[k] = J[(phi + k) * ld];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double2 f = J[(plo + i) * ld];
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[i * 6 + j] += f.x * FH[j].x + f.y * FH[j].y;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 36; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 36)
    Q.pair_chunk_acc[static_cast<int64_t>(blockIdx.x) * 36 + threadIdx.x] =
        red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

// S[A-block rows, B-block cols] = sum over slices (slice order) of the pair's elimination blocks
//   + F_A^T F_A from the camera gather (diagonal pairs; add_diag = 0 on ranks > 0, whose accumulators
//     already hold the sum over all ranks) + the chunk rows of the camera-pair gather; mirrored.
template <int CB>
__global__ void __launch_bounds__(256) k_dense_combine(DenseWork Q, const double* __restrict__ cam_B, int add_diag, int nb,
                                                        int n_slices) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(Q.n_pairs) * (CB * CB)) return;
  const int q = static_cast<int>(idx / (CB * CB)), e = static_cast<int>(idx - static_cast<int64_t>(q) * (CB * CB));
  double s = 0.0;
  for (int sl = 0; sl < n_slices; ++sl) s += Q.S_part[(static_cast<int64_t>(sl) * Q.n_pairs + q) * (CB * CB) + e];
  int A, B;
  pair_decode(q, nb, A, B);
  const int i = e / CB, j = e - i * CB;
  if (A == B) {
    if (add_diag) s += cam_B[static_cast<int64_t>(A) * CB * CB + e];
  } else if (Q.pair_chunk_first && i < 6 && j < 6) {
    for (int c = Q.pair_chunk_first[q]; c < Q.pair_chunk_first[q + 1]; ++c) s += Q.pair_chunk_acc[static_cast<int64_t>(c) * 36 + i * 6 + j];
  }
  const int64_t n = static_cast<int64_t>(nb) * CB;
  const int64_t r = static_cast<int64_t>(A) * CB + i, c = static_cast<int64_t>(B) * CB + j;
  Q.S[r * n + c] = s;
  if (A != B) Q.S[c * n + r] = s;
}

// ----------------------------------------------------------------- K9 dense Cholesky + solve
// One CTA of 1024 threads, thread t <-> row j0 + t of the current panel (panel width 16).  Row n is
// the right-hand side, so after the last panel it holds y = L^-1 rhs.  The factor overwrites the
// lower triangle of S.
// (Measured and not kept: the right-looking panel scheme of k_dense_ldlt_small — diagonal block factored per thread
// in registers, rank-8 trailing update in 4 x 4 tiles — on the L2-resident matrix: every panel then pays two round
// trips to L2 per tile, n = 306: 0.60 ms against 0.48 ms for this left-looking version.  The matrix has to sit in
// shared memory for that scheme to pay: beyond n = 152 that takes a thread-block cluster.)
constexpr int kChNB = 16;
constexpr int kChLS = 18;  // row stride of the staged panel rows: 16-byte aligned rows, spread over the banks
constexpr int kChThreads = 1024;

__global__ void __launch_bounds__(kChThreads, 1) k_dense_cholesky(double* __restrict__ S, int n, const double* __restrict__ dc2,
                                                                   const double* __restrict__ rhs, double* __restrict__ x,
                                                                   int* __restrict__ fail_flag) {
  extern __shared__ __align__(16) unsigned char smem_ch[];
  double* sLjT = reinterpret_cast<double*>(smem_ch);          // [j0][kChLS]: L[j0 + c][k] at [k][c]
  double* sy = sLjT + static_cast<size_t>(n) * kChLS;         // [n + 16] y, then x
  double* sInv = sy + n + kChNB;                              // [n] 1 / L_kk (no division on the serial chains)
  __shared__ double sD[kChNB][kChNB + 1];
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  if (tid == 0) s_ok = 1;
  for (int j0 = 0; j0 < n; j0 += kChNB) {
    const int w = min(kChNB, n - j0);
    const int row = j0 + tid;
    const bool active = row <= n;
    __syncthreads();  // previous panel written back (global + sy)
    // stage L[j0 .. j0+w) x [0, j0) transposed
    for (int i = tid; i < w * j0; i += kChThreads) {
      const int c = i / j0, k = i - c * j0;
      sLjT[k * kChLS + c] = S[static_cast<int64_t>(j0 + c) * n + k];
    }
    double p[kChNB];
    if (active) {
#pragma unroll
      for (int c = 0; c < kChNB; ++c) {
        double v = 0.0;
        if (c < w) {
          if (row < n) {
            if (j0 + c <= row) v = S[static_cast<int64_t>(row) * n + j0 + c] + (j0 + c == row ? dc2[row] : 0.0);
          } else {
            v = rhs[j0 + c];
          }
        }
        p[c] = v;
      }
    }
    __syncthreads();
    if (active && j0 > 0) {
      const double* lrow = row < n ? S + static_cast<int64_t>(row) * n : sy;
#pragma unroll 4
      for (int k = 0; k < j0; ++k) {
        const double l = lrow[k];
        const double2* lj = reinterpret_cast<const double2*>(sLjT + k * kChLS);
#pragma unroll
        for (int c = 0; c < kChNB / 2; ++c) {
          const double2 v = lj[c];
          p[2 * c] -= l * v.x;
          p[2 * c + 1] -= l * v.y;
        }
      }
    }
    // diagonal block: lane r of warp 0 keeps row r in registers (p[]), columns travel by shuffle —
    // no shared-memory round trip on the serial chain
    if (tid < 32) {
      bool good_all = true;
#pragma unroll
      for (int k = 0; k < kChNB; ++k) {
        if (k < w) {
          const double d = __shfl_sync(0xffffffffu, p[k], k);
          const bool good = d > 0.0;
          good_all = good_all && good;
          const double rs = rsqrt(good ? d : 1.0);
          if (tid == k) sInv[j0 + k] = rs;
          p[k] = (tid == k) ? d * rs : p[k] * rs;  // L[r][k] for r > k (rows above the diagonal carry zeros)
#pragma unroll
          for (int c = k + 1; c < kChNB; ++c) {
            const double lck = __shfl_sync(0xffffffffu, p[k], c);
            if (tid >= c) p[c] -= p[k] * lck;
          }
        }
      }
      if (tid == 0 && !good_all) s_ok = 0;
      if (tid < w) {
#pragma unroll
        for (int c = 0; c < kChNB; ++c) sD[tid][c] = p[c];
      }
    }
    __syncthreads();
    if (!s_ok) break;
    if (active) {
      if (tid < w) {
        // a row of the diagonal block: its factor comes from shared memory
#pragma unroll
        for (int c = 0; c < kChNB; ++c)
          if (c <= tid) S[static_cast<int64_t>(row) * n + j0 + c] = sD[tid][c];
      } else {
        // triangular solve of the row against the diagonal block (the rows that sit in warp 0 were
        // carried along by the factorisation above)
        if (tid >= 32) {
#pragma unroll
          for (int c = 0; c < kChNB; ++c) {
            if (c < w) {
              double v = p[c];
#pragma unroll
              for (int k = 0; k < c; ++k) v -= p[k] * sD[c][k];
              p[c] = v * sInv[j0 + c];
            }
          }
        }
        if (row < n) {
#pragma unroll
          for (int c = 0; c < kChNB; ++c)
            if (c < w) S[static_cast<int64_t>(row) * n + j0 + c] = p[c];
        } else {
#pragma unroll
          for (int c = 0; c < kChNB; ++c)
            if (c < w) sy[j0 + c] = p[c];
        }
      }
    }
  }
  __syncthreads();
  if (!s_ok) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = tid; i < n; i += kChThreads) x[i] = nan;
    if (tid == 0 && fail_flag) *fail_flag = 1;
    return;
  }
  // ---- back substitution L^T x = y, panels from the bottom; thread i owns unknown i
  double v = 0.0;
  if (tid < n) v = sy[tid];
  const int last = ((n - 1) / kChNB) * kChNB;
  for (int j0 = last; j0 >= 0; j0 -= kChNB) {
    const int w = min(kChNB, n - j0);
    __syncthreads();
    if (tid >= j0 && tid < j0 + w) sy[tid] = v;  // current right-hand side of the panel's unknowns
    if (tid < kChNB * kChNB) {                   // the panel's diagonal block, so that the serial part stays in shared memory
      const int r = tid / kChNB, c = tid - r * kChNB;
      if (r < w && c <= r) sD[r][c] = S[static_cast<int64_t>(j0 + r) * n + j0 + c];
    }
    __syncthreads();
    if (tid < 32) {
      // lane i: unknown j0 + i; column i of the block (L[k][i], k > i) in registers
      double col[kChNB];
#pragma unroll
      for (int k = 0; k < kChNB; ++k) col[k] = (tid < w && k < w && k > tid) ? sD[k][tid] : 0.0;
      double yi = tid < w ? sy[j0 + tid] : 0.0;
      const double inv = tid < w ? sInv[j0 + tid] : 0.0;
#pragma unroll
      for (int k = kChNB - 1; k >= 0; --k) {
        if (k < w) {
          const double xk = __shfl_sync(0xffffffffu, yi * inv, k);
          if (tid == k) yi = xk;
          else if (tid < k) yi -= col[k] * xk;
        }
      }
      if (tid < w) sy[j0 + tid] = yi;
    }
    __syncthreads();
    if (tid < j0) {
      for (int k = 0; k < w; ++k) v -= S[static_cast<int64_t>(j0 + k) * n + tid] * sy[j0 + k];
    }
  }
  __syncthreads();
  if (tid < n) x[tid] = sy[tid];
}

// Small systems (the whole augmented matrix fits in shared memory: n <= kChSmallMax): blocked right-looking
// L D L^T without square roots, panels of 8 columns.  Row n is the right-hand side, so its multipliers are
// w = D^-1 L^-1 rhs; then the column-oriented back substitution L^T x = w.  Same step as the Cholesky path up
// to rounding.  The kernel is ONE SM running a chain of dependent steps, so what counts is the length of the
// chain, not the flops: per panel every thread factors the 8 x 8 diagonal block by itself in registers (85
// multiply-adds, no communication) and solves its own row against it, then one rank-8 update of the trailing
// block in 4 x 4 register tiles — two block barriers per panel, 2 n / 8 in all, and the matrix passes through
// shared memory once per panel.  (The unblocked version — a barrier and a read-modify-write of the whole
// trailing block per column — took 156 us for n = 114.)
constexpr int kChSmallMax = 152;
constexpr int kChSmallThreads = 256;
constexpr int kLdNB = 8;
constexpr int kLdPS = kLdNB + 1;  // odd row stride of the panel buffers
__global__ void __launch_bounds__(kChSmallThreads, 1) k_dense_ldlt_small(const double* __restrict__ S, int n, const double* __restrict__ dc2,
                                                                     const double* __restrict__ rhs, double* __restrict__ x,
                                                                     int* __restrict__ fail_flag) {
  extern __shared__ __align__(16) unsigned char smem_ch[];
  const int ldm = n | 1;  // odd row stride: a column walk touches every bank
  double* A = reinterpret_cast<double*>(smem_ch);          // [n + 1][ldm] lower triangle + rhs row; column k ends as l_rk d_k
  double* sw = A + static_cast<size_t>(n + 1) * ldm;       // [n] w, then x
  double* sP = sw + n;                                     // [n + 1][kLdPS] multipliers l_rc of the current panel
  double* sQ = sP + static_cast<size_t>(n + 1) * kLdPS;    // [n + 1][kLdPS] l_rc d_c of the current panel
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  if (tid == 0) s_ok = 1;
  for (int i = tid; i < n * n; i += kChSmallThreads) {
    const int r = i / n, c = i - r * n;
    if (c <= r) A[r * ldm + c] = S[i] + (r == c ? dc2[r] : 0.0);
  }
  if (tid < n) A[n * ldm + tid] = rhs[tid];
  __syncthreads();
  bool ok = true;
  for (int k0 = 0; k0 < n; k0 += kLdNB) {
    const int w = min(kLdNB, n - k0);
    // ---- the diagonal block, factored by every thread in registers (padded with the identity beyond w)
    double B[kLdNB][kLdNB], inv[kLdNB];
#pragma unroll
    for (int i = 0; i < kLdNB; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) B[i][j] = (i < w) ? A[(k0 + i) * ldm + k0 + j] : (i == j ? 1.0 : 0.0);
#pragma unroll
    for (int c = 0; c < kLdNB; ++c) {
      const double d = B[c][c];
      if (!(d > 0.0)) ok = false;
      double iv;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(iv) : "d"(d));
      iv = iv * (2.0 - d * iv);
      iv = iv * (2.0 - d * iv);
      inv[c] = iv;
#pragma unroll
      for (int i = c + 1; i < kLdNB; ++i) {
        const double l = B[i][c] * iv;
#pragma unroll
        for (int j = c + 1; j <= i; ++j) B[i][j] = fma(-l, B[j][c], B[i][j]);
      }
    }
    if (!ok) break;  // uniform: every thread factors the same block
    // ---- this thread's row against the block: a'_rc = a_rc - sum_{m<c} l_rm a'_{cm},  l_rc = a'_rc / d_c
    const int r = k0 + w + tid;  // rows below the block (the block's own rows are B itself)
    if (r <= n) {
      double a[kLdNB], l[kLdNB];
#pragma unroll
      for (int c = 0; c < kLdNB; ++c) a[c] = c < w ? A[r * ldm + k0 + c] : 0.0;
#pragma unroll
      for (int c = 0; c < kLdNB; ++c) {
        double v = a[c];
#pragma unroll
        for (int m = 0; m < c; ++m) v = fma(-l[m], B[c][m], v);
        a[c] = v;
        l[c] = v * inv[c];
      }
#pragma unroll
      for (int c = 0; c < kLdNB; ++c) {
        if (c < w) A[r * ldm + k0 + c] = a[c];
        sP[r * kLdPS + c] = l[c];
        sQ[r * kLdPS + c] = a[c];
      }
    }
    if (tid < w) {  // the block's own rows: updated entries back into A (row k0 + tid, columns k0 .. k0 + tid)
#pragma unroll
      for (int i = 0; i < kLdNB; ++i)
        if (i == tid) {
#pragma unroll
          for (int j = 0; j <= i; ++j) A[(k0 + i) * ldm + k0 + j] = B[i][j];
        }
    }
    __syncthreads();
    // ---- rank-w update of the trailing block (rows k0 + w .. n, columns k0 + w .. min(row, n - 1)), 4 x 4 tiles
    const int base = k0 + w, m = n + 1 - base;  // trailing rows (the right-hand side row included)
    const int mt = (m + 3) >> 2, n_tiles = mt * (mt + 1) / 2;
    for (int t = tid; t < n_tiles; t += kChSmallThreads) {
      int tr = static_cast<int>((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
      while (tr * (tr + 1) / 2 > t) --tr;
      const int tj = t - tr * (tr + 1) / 2;
      const int r0 = base + 4 * tr, j0 = base + 4 * tj;
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll
      for (int c = 0; c < kLdNB; ++c) {
        double lr[4], qj[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) lr[i] = (r0 + i <= n) ? sP[(r0 + i) * kLdPS + c] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) qj[j] = (j0 + j < n) ? sQ[(j0 + j) * kLdPS + c] : 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(lr[i], qj[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = r0 + i, jj = j0 + j;
          if (rr <= n && jj < n && jj <= rr) A[rr * ldm + jj] -= acc[i][j];
        }
    }
    __syncthreads();
  }
  if (!ok && tid == 0) s_ok = 0;
  __syncthreads();
  if (!s_ok) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = tid; i < n; i += kChSmallThreads) x[i] = nan;
    if (tid == 0 && fail_flag) *fail_flag = 1;
    return;
  }
  // w_k = A[n][k] / d_k; back substitution with unit-lower L: l_ki = A[k][i] / d_i  (one barrier per unknown;
  // eight unknowns per barrier with the 8 x 8 triangle solved by every thread measured slower: 68 vs 62 us)
  double v = 0.0, inv_i = 0.0;
  if (tid < n) {
    inv_i = 1.0 / A[tid * ldm + tid];
    v = A[n * ldm + tid] * inv_i;
  }
  for (int k = n - 1; k >= 0; --k) {
    if (tid == k) sw[k] = v;
    __syncthreads();
    if (tid < k) v -= A[k * ldm + tid] * inv_i * sw[k];
  }
  if (tid < n) x[tid] = v;
}

}  // namespace

int dense_slices(const DenseWork& Q) {
  // a fixed function of the problem (not of the device), so that runs are reproducible: enough
  // CTAs to fill 148 SMs twice over, at least four batches per slice
  const int groups = (Q.n_pairs + kDnThreads - 1) / kDnThreads;
  int s = std::max(1, 296 / std::max(groups, 1));
  s = std::min(s, std::max(1, Q.n_batches / 2));
  return std::max(s, 1);
}

template <int CB, int MINB>
static int launch_schur_dense_t(const DeviceProblem& D, const WorkArrays& W, const DenseWork& Q, int add_diag, cudaStream_t st) {
  const size_t smem = DnSmem<CB>::bytes(D.n_blocks);
  static unsigned long long configured = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(configured & (1ull << (dev & 63)))) {
    configured |= 1ull << (dev & 63);
    cudaFuncSetAttribute(k_dense_pairs<CB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_dense_z<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  }
  const int groups = (Q.n_pairs + kDnThreads - 1) / kDnThreads;
  const int slices = dense_slices(Q);
  if (Q.n_batches > 0) k_dense_z<CB><<<Q.n_batches, kDnThreads, DzSmem::bytes(D.n_blocks, CB), st>>>(D, W, Q);
  k_dense_pairs<CB, MINB><<<dim3(slices, groups), kDnThreads, smem, st>>>(D, Q);
  if (Q.n_pair_chunks > 0) k_pair_gather<<<Q.n_pair_chunks, 128, 0, st>>>(D, Q);
  const int64_t total = static_cast<int64_t>(Q.n_pairs) * CB * CB;
  k_dense_combine<CB><<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(Q, W.cam_acc, add_diag, D.n_blocks, slices);
  return 0;
}

int launch_schur_dense(const DeviceProblem& D, const WorkArrays& W, const DenseWork& Q, int add_diag, cudaStream_t st) {
  if (D.n_blocks == 0 || Q.n_pairs == 0) return 0;
  if (D.cb == 6) return launch_schur_dense_t<6, 2>(D, W, Q, add_diag, st);
  if (D.cb == 9) return launch_schur_dense_t<9, 1>(D, W, Q, add_diag, st);
  return -1;
}

int launch_dense_cholesky(const DeviceProblem& D, const WorkArrays& W, const DenseWork& Q, cudaStream_t st) {
  const int n = D.n_blocks * D.cb;
  if (n == 0) return 0;
  const size_t smem = sizeof(double) * (static_cast<size_t>(n) * kChLS + 2 * static_cast<size_t>(n) + kChNB);
  static unsigned long long configured = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(configured & (1ull << (dev & 63)))) {
    configured |= 1ull << (dev & 63);
    cudaFuncSetAttribute(k_dense_cholesky, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_dense_ldlt_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);  // (+ 1 KB static <= 227 KB)
  }
  const double* rhs = W.cam_acc + static_cast<int64_t>(D.n_blocks) * D.cb * D.cb + 2 * static_cast<int64_t>(n);
  if (n <= kChSmallMax) {
    const size_t small = sizeof(double) * (static_cast<size_t>(n + 1) * (n | 1) + n + 2 * static_cast<size_t>(n + 1) * 9);
    k_dense_ldlt_small<<<1, kChSmallThreads, small, st>>>(Q.S, n, W.dc2, rhs, W.x, Q.fail_flag);
    return 0;
  }
  k_dense_cholesky<<<1, kChThreads, smem, st>>>(Q.S, n, W.dc2, rhs, W.x, Q.fail_flag);
  return 0;
}

}  // namespace dba
