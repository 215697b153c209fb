// kind::f16 tcgen05 helpers for the weight-gradient GEMM (fp16 operands, fp32
// accumulate) and the MN-major SWIZZLE_128B image of a 16-bit operand.
//
// Canonical MN-major 128-byte-swizzle layout of a 16-bit operand (CUTLASS
// cute/atom/mma_traits_sm100.hpp "LayoutType::B128": Swizzle<3,4,3> o
// ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)), T = 8 halfs per 16 bytes): an atom is 64
// MN-elements x 8 K-rows = 8 lines of 128 bytes; inside line r the 16-byte piece
// that holds MN-elements 8c .. 8c+7 sits at piece (c ^ (r & 7)).  Atoms along MN are
// LBO bytes apart, atoms along K (groups of 8 K-rows) SBO bytes apart.
#pragma once

#include <cuda_fp16.h>

#include "umma.cuh"

namespace ibc {
namespace umma {

// kind::f16, fp16 operands, fp32 accumulate, M x N
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] . B[smem]^T with 16-bit A packed two K-elements per 32-bit column
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Image of a [K rows][MN] 16-bit operand block, K a multiple of 8, MN a multiple of 64:
// [MN / 64 atoms][K / 8 atoms][8 lines][128 B].
struct ImgF16 {
  static constexpr uint32_t kSbo = 1024;                        // K-atom stride
  __host__ __device__ static constexpr uint32_t lbo(int K) { return (uint32_t)(K / 8) * 1024u; }
  // byte offset of the 16-byte piece holding MN-elements 8c .. 8c+7 of K-row r
  __host__ __device__ static uint32_t piece(int r, int c, int K) {
    return (uint32_t)(c >> 3) * lbo(K) + (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 128u +
           (uint32_t)(((c & 7) ^ (r & 7)) << 4);
  }
};

// two floats -> packed fp16 pair (round to nearest, saturating to +-65504 instead of inf)
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// max of two packed fp16 pairs (relu on a packed pair: against 0; NaN -> the other operand)
__device__ __forceinline__ uint32_t hmax2_bits(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// the same toward zero: for an operand that already carries the tf32 round-to-nearest offset
// (bits + 0x1000, umma_pack.cuh) truncating to fp16's 10-bit mantissa completes the rounding
__device__ __forceinline__ uint32_t pack_half2_rz(float lo, float hi) {
  uint32_t r;
  asm("cvt.rz.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace umma
}  // namespace ibc
