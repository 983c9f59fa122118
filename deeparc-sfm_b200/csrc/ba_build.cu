// Device-side problem construction: what the reference does per solve() call in
// `sfm.cc:36-65` (one residual block per observation, parameter blocks by pointer) and what
// dba_problem_set otherwise builds on the host cores — the per-observation records, the CSR offsets of
// the points, the tile-local camera incidence (partials, items, matrix-free columns), the partials
// grouped by camera and the camera-sorted incidence in chunks — built on the GPU from the caller's raw
// arrays.  Handles the case that matters for throughput: point-sorted observations, one pose per
// observation; anything else reports "not handled" and the host build runs.
//
// Sorts and scans of index arrays use CUB (toolkit header library): this is plumbing executed once per
// upload, not the solve path.  Every step is deterministic (stable sorts, fixed-order scans), and the
// structures are the ones the host build produces up to the order of a tile's camera blocks (by block id
// here, by first appearance there), which no sum depends on: solves are bit-identical (GPU test).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>

#include "ba_build.cuh"

namespace dba {

namespace {

// ---- per-observation records + validation flags
//   flags: 1 index out of range, 2 not sorted by point, 4 some observation composes two poses, 8 intrinsic != pose a
__global__ void __launch_bounds__(256) k_bld_obs(int64_t n, const double* __restrict__ xy, const int* __restrict__ pt,
                                                  const int* __restrict__ pa, const int* __restrict__ pb, const int* __restrict__ in,
                                                  int pt_lo, int n_pts_local, int n_ext, int n_intr, int prev_pt, double2* __restrict__ o_xy,
                                                  int2* __restrict__ o_ip, int2* __restrict__ o_ab, int* __restrict__ o_a, int* __restrict__ flags) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (k >= n) return;
  const int p = pt[k], a = pa[k], i = in[k], b = pb ? pb[k] : -1;
  int f = 0;
  if (p < pt_lo || p >= pt_lo + n_pts_local || a < 0 || a >= n_ext || i < 0 || i >= n_intr || b < -1 || b >= n_ext) f |= 1;
  if ((k > 0 ? pt[k - 1] : prev_pt) > p) f |= 2;
  if (b >= 0) f |= 4;
  if (i != a) f |= 8;
  if (f) atomicOr(flags, f);
  o_xy[k] = make_double2(xy[2 * k], xy[2 * k + 1]);
  o_ip[k] = make_int2(i, p - pt_lo);
  o_ab[k] = make_int2(a, b);
  o_a[k] = a;
}

// first[i] = first position whose key >= base + i, i = 0 .. m (keys sorted ascending)
__global__ void __launch_bounds__(256) k_bld_lower_bound(const int* __restrict__ keys, int64_t n, int base, int m, int* __restrict__ first) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i > m) return;
  const int target = base + i;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  first[i] = static_cast<int>(lo);
}

// ---- tile-local camera incidence.  One CTA per tile: the tile's observations sorted by (block, position)
// with a bitonic network in shared memory; runs of one block = the tile's partials.
constexpr int kBldThreads = 256;
template <int CAP>  // tile capacity: 256, 512 or 1024 (power of two)
__device__ __forceinline__ int tile_sort(const DeviceBuild& B, const TileMeta& tm, unsigned long long* keys) {
  const int tid = threadIdx.x;
  for (int i = tid; i < CAP; i += kBldThreads)
    keys[i] = i < tm.n_obs ? (static_cast<unsigned long long>(static_cast<unsigned int>(B.obs_ab[tm.obs0 + i].x)) << 32) | static_cast<unsigned int>(i)
                           : ~0ull;
  __syncthreads();
  for (int k = 2; k <= CAP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < CAP; i += kBldThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  return tm.n_obs;
}

template <int CAP>
__global__ void __launch_bounds__(kBldThreads) k_bld_tile_count(DeviceBuild B) {
  __shared__ unsigned long long keys[CAP];
  __shared__ int s_cnt;
  const TileMeta tm = B.tile_meta[blockIdx.x];
  if (threadIdx.x == 0) s_cnt = 0;
  tile_sort<CAP>(B, tm, keys);
  int mine = 0;
  for (int i = threadIdx.x; i < tm.n_obs; i += kBldThreads)
    mine += (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1 : 0;
  if (mine) atomicAdd(&s_cnt, mine);  // integer count: order does not matter
  __syncthreads();
  if (threadIdx.x == 0) B.tile_np[blockIdx.x] = s_cnt;
}

template <int CAP, int CB>
__global__ void __launch_bounds__(kBldThreads) k_bld_tile_fill(DeviceBuild B) {
  __shared__ unsigned long long keys[CAP];
  __shared__ int s_run[CAP];  // run index (local block) of every sorted position, via a scan of the run heads
  const int t = blockIdx.x, tid = threadIdx.x;
  TileMeta tm = B.tile_meta[t];
  tile_sort<CAP>(B, tm, keys);
  const int n = tm.n_obs;
  // inclusive scan of the run heads (CAP <= 1024: one pass per thread strip + a serial fix-up by thread 0)
  for (int i = tid; i < n; i += kBldThreads) s_run[i] = (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1 : 0;
  __syncthreads();
  for (int off = 1; off < n; off <<= 1) {  // Hillis-Steele, CAP / 256 elements per thread
    int v[CAP / kBldThreads];
#pragma unroll
    for (int q = 0; q < CAP / kBldThreads; ++q) {
      const int i = tid + q * kBldThreads;
      v[q] = (i < n && i >= off) ? s_run[i - off] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CAP / kBldThreads; ++q) {
      const int i = tid + q * kBldThreads;
      if (i < n) s_run[i] += v[q];
    }
    __syncthreads();
  }
  const int g0 = B.tile_g0[t];
  const int n_parts = n > 0 ? s_run[n - 1] : 0;
  if (tid == 0) {
    tm.g0 = g0;
    tm.n_parts = n_parts;
    tm.item0 = tm.obs0;  // one slot-0 item per observation
    tm.n_items = n;
    B.tile_meta[t] = tm;
    B.part_first[g0 + t + n_parts] = n;
    B.part_first_rel[g0 + t + n_parts] = static_cast<unsigned short>(n);
  }
  for (int i = tid; i < n; i += kBldThreads) {
    const int lo = static_cast<int>(keys[i] & 0xffffffffu), blk = static_cast<int>(keys[i] >> 32), lc = s_run[i] - 1;
    const int o = tm.obs0 + lo;
    const int lp = B.obs_ip[o].y - tm.pt0;
    B.items[tm.obs0 + i] = static_cast<unsigned short>(lo);
    B.obs_lc[o] = make_ushort2(static_cast<unsigned short>(lc), 0xffff);
    B.obs_lp[o] = static_cast<unsigned short>(lp);
    if (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) {
      B.part_first[g0 + t + lc] = i;
      B.part_first_rel[g0 + t + lc] = static_cast<unsigned short>(i);
      B.part_blk[g0 + lc] = blk;
      B.part_key[g0 + lc] = blk;
      B.part_id[g0 + lc] = g0 + lc;
    }
    const int lplo = static_cast<int>((static_cast<unsigned int>(lp) << 16) | static_cast<unsigned int>(lo));
    if (CB == 9) {
      reinterpret_cast<int2*>(B.mf_cols)[static_cast<int64_t>(t) * CAP + i] = make_int2(blk, lplo);
    } else if (CB == 6) {
      reinterpret_cast<int4*>(B.mf_cols)[static_cast<int64_t>(t) * CAP + i] = make_int4(blk, -1, lplo, B.obs_ip[o].x);
    }
  }
  for (int i = n + tid; i < CAP; i += kBldThreads) {  // padding columns
    if (CB == 9) reinterpret_cast<int2*>(B.mf_cols)[static_cast<int64_t>(t) * CAP + i] = make_int2(-1, 0);
    else if (CB == 6) reinterpret_cast<int4*>(B.mf_cols)[static_cast<int64_t>(t) * CAP + i] = make_int4(-1, 0, 0, 0);
  }
}

// points-only problems (freeze_camera): no camera incidence; the tile kernels still want the tile-local point of
// every observation, and "no staged camera row" in obs_lc
__global__ void __launch_bounds__(kBldThreads) k_bld_tile_plain(DeviceBuild B) {
  TileMeta tm = B.tile_meta[blockIdx.x];
  for (int i = threadIdx.x; i < tm.n_obs; i += kBldThreads) {
    const int o = tm.obs0 + i;
    B.obs_lp[o] = static_cast<unsigned short>(B.obs_ip[o].y - tm.pt0);
    B.obs_lc[o] = make_ushort2(0xffff, 0xffff);
  }
  if (threadIdx.x == 0) {
    tm.g0 = 0;
    tm.n_parts = 0;
    tm.item0 = 0;
    tm.n_items = 0;
    B.tile_meta[blockIdx.x] = tm;
  }
}

__global__ void __launch_bounds__(256) k_bld_iota2(int64_t n, int* __restrict__ out) {  // out[k] = 2 k (slot-0 entries)
  const int64_t k = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (k < n) out[k] = static_cast<int>(2 * k);
}
__global__ void __launch_bounds__(256) k_bld_scatter_pos(int n, const int* __restrict__ order, int* __restrict__ dst) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[order[i]] = i;  // partial order[i] sits at row i of the camera-grouped buffer
}
__global__ void __launch_bounds__(256) k_bld_chunk_counts(int n_ext, const int* __restrict__ first, int* __restrict__ cnt) {
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b < n_ext) cnt[b] = (first[b + 1] - first[b] + 1023) / 1024;
}
__global__ void __launch_bounds__(256) k_bld_chunks(int n_ext, const int* __restrict__ first, const int* __restrict__ chunk_first,
                                                     int4* __restrict__ chunks) {
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b >= n_ext) return;
  int c = chunk_first[b];
  for (int e = first[b]; e < first[b + 1]; e += 1024) chunks[c++] = make_int4(b, e, min(e + 1024, first[b + 1]), 0);
}

template <int CAP>
void launch_tiles(const DeviceBuild& B, int cb, int n_tiles, bool fill, cudaStream_t st) {
  if (n_tiles == 0) return;
  if (!fill) {
    k_bld_tile_count<CAP><<<n_tiles, kBldThreads, 0, st>>>(B);
  } else if (cb == 9) {
    k_bld_tile_fill<CAP, 9><<<n_tiles, kBldThreads, 0, st>>>(B);
  } else if (cb == 6) {
    k_bld_tile_fill<CAP, 6><<<n_tiles, kBldThreads, 0, st>>>(B);
  } else {
    k_bld_tile_fill<CAP, 0><<<n_tiles, kBldThreads, 0, st>>>(B);
  }
}

}  // namespace

void bld_observations(int64_t n, const double* xy, const int* pt, const int* pa, const int* pb, const int* in, int pt_lo,
                      int n_pts_local, int n_ext, int n_intr, int prev_pt, double2* o_xy, int2* o_ip, int2* o_ab, int* o_a, int* flags,
                      cudaStream_t st) {
  if (n > 0)
    k_bld_obs<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(n, xy, pt, pa, pb, in, pt_lo, n_pts_local, n_ext, n_intr, prev_pt,
                                                                       o_xy, o_ip, o_ab, o_a, flags);
}

void bld_lower_bound(const int* keys, int64_t n, int base, int m, int* first, cudaStream_t st) {
  k_bld_lower_bound<<<(m + 1 + 255) / 256, 256, 0, st>>>(keys, n, base, m, first);
}

void bld_tiles_plain(const DeviceBuild& B, int n_tiles, cudaStream_t st) {
  if (n_tiles > 0) k_bld_tile_plain<<<n_tiles, kBldThreads, 0, st>>>(B);
}

void bld_tiles(const DeviceBuild& B, int cb, int tile_cap, int n_tiles, bool fill, cudaStream_t st) {
  if (tile_cap == 256) launch_tiles<256>(B, cb, n_tiles, fill, st);
  else if (tile_cap == 512) launch_tiles<512>(B, cb, n_tiles, fill, st);
  else launch_tiles<1024>(B, cb, n_tiles, fill, st);
}

size_t bld_temp_bytes(int64_t n_obs, int n_partials_max, int n_tiles, int n_ext) {
  size_t a = 0, b = 0, c = 0, d = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, static_cast<const int*>(nullptr), static_cast<int*>(nullptr), static_cast<const int*>(nullptr),
                                  static_cast<int*>(nullptr), static_cast<int>(n_obs));
  cub::DeviceRadixSort::SortPairs(nullptr, b, static_cast<const int*>(nullptr), static_cast<int*>(nullptr), static_cast<const int*>(nullptr),
                                  static_cast<int*>(nullptr), n_partials_max);
  cub::DeviceScan::ExclusiveSum(nullptr, c, static_cast<const int*>(nullptr), static_cast<int*>(nullptr), n_tiles + 1);
  cub::DeviceScan::ExclusiveSum(nullptr, d, static_cast<const int*>(nullptr), static_cast<int*>(nullptr), n_ext + 1);
  return std::max(std::max(a, b), std::max(c, d)) + 256;
}

int bld_exclusive_sum(void* temp, size_t temp_bytes, const int* in, int* out, int n, cudaStream_t st) {
  return cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, n, st) == cudaSuccess ? 0 : -1;
}

int bld_sort_pairs(void* temp, size_t temp_bytes, const int* keys_in, int* keys_out, const int* vals_in, int* vals_out, int n, int key_bits,
                   cudaStream_t st) {
  return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, key_bits, st) == cudaSuccess ? 0 : -1;
}

void bld_iota2(int64_t n, int* out, cudaStream_t st) {
  if (n > 0) k_bld_iota2<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(n, out);
}
void bld_scatter_pos(int n, const int* order, int* dst, cudaStream_t st) {
  if (n > 0) k_bld_scatter_pos<<<(n + 255) / 256, 256, 0, st>>>(n, order, dst);
}
void bld_chunk_counts(int n_ext, const int* first, int* cnt, cudaStream_t st) {
  if (n_ext > 0) k_bld_chunk_counts<<<(n_ext + 255) / 256, 256, 0, st>>>(n_ext, first, cnt);
}
void bld_chunks(int n_ext, const int* first, const int* chunk_first, int4* chunks, cudaStream_t st) {
  if (n_ext > 0) k_bld_chunks<<<(n_ext + 255) / 256, 256, 0, st>>>(n_ext, first, chunk_first, chunks);
}

}  // namespace dba
