// Packed tcgen05 operand buffer of one MLPEBM (built by pack_kernel in
// mlp_umma.cu) and small helpers shared by the fused kernels.
#pragma once

#include "common.cuh"

namespace ibc {

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// tcgen05 kind::tf32 reads the top 19 bits of each fp32 operand (truncation);
// rounding to nearest (ties away from zero, like cvt.rna.tf32.f32) on the way
// in halves the error and removes the bias.  Done with two integer ALU ops:
// the F2F conversion instruction runs at a quarter of the ALU rate and was the
// epilogue's bottleneck (3750 -> cycles per layer, scripts/timeline.py).
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// relu + round-to-nearest-tf32 of an accumulator in ONE integer instruction
// (VIADDMNMX): max(bits + 0x1000, 0).  The low 13 bits are left for the tensor
// core to truncate, which completes the rounding; negative inputs (sign bit =
// negative integer) clamp to +0.  The result is only meant to be an MMA operand.
__device__ __forceinline__ uint32_t relu_tf32_bits(uint32_t x) {
  return (uint32_t)__viaddmax_s32((int)x, 0x1000, 0);
}
// round-to-nearest-tf32 of a signed value, low bits left for the MMA to truncate
__device__ __forceinline__ uint32_t tf32_rn_bits(uint32_t x) { return x + 0x1000u; }
__device__ __forceinline__ float4 tf32_rn4(float4 v) {
  return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
}

// Offsets (in floats) inside the packed operand buffer.
struct PackedLayout {
  int W, nb, KP, AP;
  int64_t fwd_p;     // [KP/4][W][4]           act rows of Wp (B operand of step 0)
  int64_t fwd_l;     // 2*nb matrices [W/4][W][4]: W1_0, W2_0, W1_1, ...
  int64_t rev_l;     // 2*nb matrices: W2_{nb-1}^T, W1_{nb-1}^T, ..., W2_0^T, W1_0^T
  int64_t rev_p;     // [W/4][AP][4]           Wp act rows, B operand of dE/dact
  int64_t vecs;      // [(2nb+1)][W] epilogue bias vectors, then we[W], be, pad
  int64_t raw;       // [(2nb+1)][W] per-step raw biases: zeros, b1_0, b2_0, b1_1, ...
  int64_t fwd_bias;  // [(2nb+1)][2][W][4] bias as one extra K=8 step: (hi, lo, 0, 0 | 0...)
  int64_t fwd_lb;    // 2*nb blocks [W/4][W][4] + [2][W][4]: layer i's operand followed by its
                     // bias slice, so that one bulk copy brings a whole layer
  int64_t total;
};

inline PackedLayout make_packed_layout(const Layout& L) {
  PackedLayout p;
  p.W = L.W; p.nb = L.nb;
  p.KP = round_up(L.A, 8);
  p.AP = round_up(L.A, 16);
  const int64_t WW = (int64_t)L.W * L.W;
  p.fwd_p = 0;
  p.fwd_l = p.fwd_p + (int64_t)L.W * p.KP;
  p.rev_l = p.fwd_l + 2 * L.nb * WW;
  p.rev_p = p.rev_l + 2 * L.nb * WW;
  p.vecs = p.rev_p + (int64_t)L.W * p.AP;
  p.raw = p.vecs + (int64_t)(2 * L.nb + 2) * L.W + 4;
  p.fwd_bias = p.raw + (int64_t)(2 * L.nb + 1) * L.W;
  p.fwd_lb = p.fwd_bias + (int64_t)(2 * L.nb + 1) * 8 * L.W;
  p.total = p.fwd_lb + (int64_t)(2 * L.nb) * (WW + 8 * L.W);
  return p;
}


// Offsets (in halves) inside the fp16 operand buffer of a width-128 stack (langevin_f16.cu)
struct Packed16Layout {
  int KP, AP;
  int64_t p0;        // [KP/8][W][8]   act rows of Wp (B operand of step 0)
  int64_t fwd;       // 2*nb blocks [W/8][W][8] + [2][W][8]: layer i's operand + its bias slice
  int64_t rev;       // 2*nb matrices [W/8][W][8]: W2_{nb-1}^T, W1_{nb-1}^T, ..., W1_0^T
  int64_t revp;      // [W/8][AP][8]   Wp act rows, B operand of dE/dact
  int64_t total;
};
Packed16Layout make_packed16_layout(const Layout& L);     // langevin_f16.cu

// MN-major tf32 operand layout (UMMA SWIZZLE_128B_BASE32B, layout type 1; the
// only shared-memory layout tcgen05 accepts for MN-major 32-bit operands).
// Atoms of 32 MN-elements x 4 K-rows = 512 B: K-row kr is a 128-byte line made
// of four 32-byte units (8 consecutive MN-elements each); the unit that holds
// MN-octet o of line kr is (o ^ kr) (Swizzle<2,5,2> on byte addresses).
// Decoded on hardware with one-hot probes (scripts/diag_mn.py).
// Byte offset of element (mn, k) for atom strides lbo (MN) and sbo (K):
__host__ __device__ inline uint32_t mn32_offset(int mn, int k, uint32_t lbo, uint32_t sbo) {
  const int kr = k & 3, oct = (mn >> 3) & 3;
  return (uint32_t)(mn >> 5) * lbo + (uint32_t)(k >> 2) * sbo + (uint32_t)kr * 128u +
         (uint32_t)((oct ^ kr) << 5) + (uint32_t)(mn & 7) * 4u;
}

// Saved operand tiles of the training path: [tile][slot][sub-tile (4)][chunk
// (W/4)][32 rows][4 floats].  A warp of the epilogue (32 consecutive rows = one
// sub-tile) writing one 16-byte chunk per thread covers 512 contiguous bytes, so
// saves and reloads are fully coalesced (the earlier MN-major-in-global layout
// cost one 128-byte line per THREAD per store and was L2-request bound); one
// sub-tile (32 x W floats) is a single contiguous bulk copy.  The weight-gradient
// kernel re-arranges a sub-tile into the MN-major UMMA atoms in shared memory.
template <int W>
struct TrainTile {
  static constexpr int kSubRows = 32;
  static constexpr int kSubTiles = 128 / kSubRows;
  static constexpr int kSubFloats = kSubRows * W;
  static constexpr int kTileFloats = 128 * W;
  // MN-major operand image of one sub-tile: [W/32 MN atoms][8 K atoms][512 B]
  static constexpr uint32_t kSbo = 512;
  static constexpr uint32_t kLbo = (kSubRows / 4) * 512;
  __host__ __device__ static int64_t slot_base(int64_t tile, int slots, int slot) {
    return (tile * slots + slot) * (int64_t)kTileFloats;
  }
  // float4 index inside a slot of chunk c (features 4c .. 4c+3) of row `row`
  __host__ __device__ static int chunk_f4(int row, int c) {
    return (row >> 5) * (kSubFloats / 4) + c * kSubRows + (row & 31);
  }
  // ReLU masks of the 2nb W x W layers: [tile][layer slot][128 rows][W/32 words],
  // bit b of word w = (operand column 32w + b > 0).  The reverse chain only needs
  // these 16 bytes per row and layer.
  __host__ __device__ static int64_t mask_index(int64_t tile, int nb, int slot, int row) {
    return ((tile * (2 * nb) + slot) * 128 + row) * (int64_t)(W / 32);
  }
  // float4 index of the same piece inside the MN-major image of its sub-tile
  __host__ __device__ static int mn_f4(int rr, int c) {
    const int ka = rr >> 2, kr = rr & 3, ma = c >> 3, oct = (c >> 1) & 3, half = c & 1;
    return (ma * (kSubRows / 4) + ka) * 32 + kr * 8 + ((oct ^ kr) << 1) + half;
  }
};

}  // namespace ibc
