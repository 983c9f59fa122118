// Device-side problem construction (ba_build.cu): launch interface used by dba_problem_set.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_kernels.cuh"

namespace dba {

// Outputs of the tile-incidence kernels (device pointers; same arrays the host build uploads)
struct DeviceBuild {
  const int2* obs_ab;               // in: (block a, -1) per observation
  const int2* obs_ip;               // in: (intrinsic, local point)
  TileMeta* tile_meta;              // in: obs0, n_obs, pt0, n_pts; out: g0, n_parts, item0, n_items
  int* tile_np;                     // [n_tiles + 1] partials per tile (count pass)
  const int* tile_g0;               // [n_tiles + 1] exclusive scan of tile_np
  unsigned short* items;            // [n_obs]
  unsigned short* part_first_rel;   // [n_partials + n_tiles + 1]
  int* part_first;                  // [n_partials + n_tiles + 1]
  int* part_blk;                    // [n_partials]
  int* part_key;                    // [n_partials] sort key (block) ...
  int* part_id;                     // [n_partials] ... and value (partial index)
  ushort2* obs_lc;                  // [n_obs]
  unsigned short* obs_lp;           // [n_obs]
  void* mf_cols;                    // [n_tiles][tile] int2 (CB = 9) / int4 (CB = 6)
};

void bld_observations(int64_t n, const double* xy, const int* pt, const int* pa, const int* pb, const int* in, int pt_lo,
                      int n_pts_local, int n_ext, int n_intr, int prev_pt, double2* o_xy, int2* o_ip, int2* o_ab, int* o_a, int* flags,
                      cudaStream_t st);
void bld_lower_bound(const int* keys, int64_t n, int base, int m, int* first, cudaStream_t st);
void bld_tiles(const DeviceBuild& B, int cb, int tile_cap, int n_tiles, bool fill, cudaStream_t st);
void bld_tiles_plain(const DeviceBuild& B, int n_tiles, cudaStream_t st);  // points-only problems
size_t bld_temp_bytes(int64_t n_obs, int n_partials_max, int n_tiles, int n_ext);
int bld_exclusive_sum(void* temp, size_t temp_bytes, const int* in, int* out, int n, cudaStream_t st);
int bld_sort_pairs(void* temp, size_t temp_bytes, const int* keys_in, int* keys_out, const int* vals_in, int* vals_out, int n, int key_bits,
                   cudaStream_t st);
void bld_iota2(int64_t n, int* out, cudaStream_t st);
void bld_scatter_pos(int n, const int* order, int* dst, cudaStream_t st);
void bld_chunk_counts(int n_ext, const int* first, int* cnt, cudaStream_t st);
void bld_chunks(int n_ext, const int* first, const int* chunk_first, int4* chunks, cudaStream_t st);

}  // namespace dba
