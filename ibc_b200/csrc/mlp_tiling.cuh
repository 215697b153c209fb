// Work distribution of the persistent width-128 kernels (mlp_umma_tmem.cu, mlp_fwd_halves.cu):
// two 128-row tiles in flight per CTA.
#pragma once
#include <stdint.h>

namespace ibc {

// Work items of one CTA: `rounds` full rounds in which every CTA takes two
// consecutive 128-row tiles, then one tail round.  When no more than gridDim.x
// tiles are left the tail hands them out one per CTA (a lone tile runs without
// MMA/epilogue overlap, but every SM gets one, instead of half the SMs getting
// two: 131 584 rows = 1028 tiles over 148 SMs are 7 tiles per SM rather than 8).
struct Tiling {
  int64_t tiles;      // ceil(R / 128)
  int rounds;         // full two-tile rounds
  int tail_single;    // tail round: one tile per CTA
};

inline Tiling make_tiling(int64_t R, int grid) {
  Tiling w;
  w.tiles = (R + 127) / 128;
  w.rounds = (int)(w.tiles / (2 * (int64_t)grid));
  const int64_t rem = w.tiles - (int64_t)w.rounds * 2 * grid;
  w.tail_single = rem <= grid ? 1 : 0;
  return w;
}

// item k of this CTA -> first tile and tile count (0 = no more work)
__device__ __forceinline__ int work_item(const Tiling& w, int k, int64_t& tile0) {
  if (k < w.rounds) {
    tile0 = ((int64_t)k * gridDim.x + blockIdx.x) * 2;
    return 2;
  }
  if (k > w.rounds) return 0;
  const int64_t base = (int64_t)w.rounds * 2 * gridDim.x;
  tile0 = base + (w.tail_single ? (int64_t)blockIdx.x : 2 * (int64_t)blockIdx.x);
  const int64_t left = w.tiles - tile0;
  return left <= 0 ? 0 : (w.tail_single ? 1 : (left < 2 ? 1 : 2));
}

}  // namespace ibc
