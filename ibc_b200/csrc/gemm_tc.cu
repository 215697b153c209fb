// General tf32 tensor-core GEMM (tcgen05, accumulator in tensor memory) over plain
// row-major fp32 operands with the fused prologue / epilogue of GemmP (gemm_simt.cuh).
// It is the IBC_PRECISION_TF32 body of every dense contraction that is not one of the
// layer-fused MLPEBM chains: the ConvMaxpool encoder's convolutions (im2col rows) and
// the DenseResnetValue head of PixelEBM (networks/layers/conv_maxpool.py:20-48,
// networks/layers/dense_resnet_value.py:31-69) with their data / weight gradients, and
// the second-order passes of the gradient penalty (ibc/losses/gradient_loss.py:23-72).
//
// One CTA = one 128 x BN block of C over one K split.  The operands have no packed copy,
// so K-slices of 32 are brought in by 8 loader warps with asynchronous 16-byte copies
// (cp.async: S - 1 stages in flight without holding registers) straight into the layout
// the MMA wants -- chosen per operand so that neither side transposes -- and each thread
// then applies the prologue (ReLU / gate, round-to-nearest tf32) in place to the pieces
// it copied:
//   * operand contiguous along K (A[m, k], B^T[n, k]):  K-major SWIZZLE_128B, a row of 32
//     K-elements is one 128-byte line; eight lanes take a row (128 contiguous bytes of
//     global memory -> one conflict-free line; the unswizzled chunk-major layout made
//     the asynchronous copies 8-way bank conflicted, ncu r1);
//   * operand contiguous along M / N (A^T[k, m], B[k, n]):  MN-major
//     SWIZZLE_128B_BASE32B atoms (umma_pack.cuh::mn32_offset); a warp takes one K-row
//     (512 contiguous bytes of global memory, conflict-free 128-byte lines).
// The A operand can also be the im2col matrix of an NHWC tensor that is never written out
// (implicit 3x3 convolution, GemmP::conv_*): the copy addresses are the gather.
// A 3-4 stage ring (full / empty mbarriers) feeds one issuer warp (4 x 128xBNx8 MMAs per
// stage); two CTAs per SM overlap one block's epilogue with the other's main loop.  The
// epilogue re-reads the accumulator by 32-column chunks, transposes each chunk through
// shared memory (the ring is free by then) so that a warp's loads of gate / residual and
// its stores (or atomics, split-K) cover 128 contiguous bytes of a C row.
#include "gemm_simt.cuh"
#include "umma.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kBM = 128, kBK = 32;
constexpr int kLoaderWarps = 8;
constexpr int kLoaderThreads = kLoaderWarps * 32;
constexpr int kThreads = kLoaderThreads + 32;
constexpr uint32_t kMnLbo = (kBK / 4) * 512u;   // between 32-element MN atoms of a stage
constexpr uint32_t kMnSbo = 512u;               // between 4-row K atoms

template <int BN>
struct TcCfg {
  static constexpr int kStages = BN <= 64 ? 4 : 3;
  static constexpr int kABytes = kBM * kBK * 4;
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRing = kStages * kStageBytes;
  static constexpr int kStageFloats = kLoaderWarps * 32 * 33;   // epilogue transposing stage
  static_assert(kStageFloats * 4 <= kRing, "epilogue stage must fit in the ring");
  static constexpr size_t kSmem = (size_t)kRing + 1024 + 128;
};

// ---- Ampere-style asynchronous 16-byte copies (LDGSTS): global -> shared without
// registers, any 16-byte-aligned destination, src_bytes = 0 writes zeros ---------------
template <bool L1>
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  if (L1)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
                 : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st_shared16(uint32_t dst, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_shared16(uint32_t src) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(src)
               : "memory");
  return v;
}

// Bounded mbarrier wait without a suspend-time hint (the hardware's default time limit):
// the waits of this kernel sit on the critical path of every K-slice, and the 20 us hint
// of umma::mbar_wait (chosen for the layer-fused kernels' long idle waits) delays wake-ups.
__device__ __forceinline__ bool tc_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag,
                                        int* status, int code) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((it & 63) == 63 && *abort_flag) return false;
  }
  *abort_flag = 1;
  if (status) atomicCAS(status, 0, code);
  return false;
}

// Four consecutive elements along the contiguous dimension of a row-major operand.
//   KC  (contiguous along K):   element j = src[r * ld + k + j]   (valid: r < rlim, k + j < klim)
//   !KC (contiguous along M/N): element j = src[k * ld + r + j]   (valid: k < klim, r + j < rlim)
template <bool KC>
__device__ __forceinline__ float4 fetch4(const float* __restrict__ src, int64_t ld, int r, int k,
                                         int rlim, int klim, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= rlim || k >= klim) return v;
  const float* q = KC ? src + (int64_t)r * ld + k : src + (int64_t)k * ld + r;
  const int left = KC ? klim - k : rlim - r;
  if (vec && left >= 4) {
    v = __ldg(reinterpret_cast<const float4*>(q));
  } else {
    v.x = __ldg(q);
    if (left > 1) v.y = __ldg(q + 1);
    if (left > 2) v.z = __ldg(q + 2);
    if (left > 3) v.w = __ldg(q + 3);
  }
  return v;
}

// The same four elements, on their way to shared memory: an asynchronous copy when the
// piece is whole and aligned, zeros when it lies outside the matrix, else scalar loads.
template <bool KC>
__device__ __forceinline__ void issue4(uint32_t dst, const float* __restrict__ src, int64_t ld, int r,
                                       int k, int rlim, int klim, bool vec) {
  if (r >= rlim || k >= klim) {
    if (vec) cp_async16<false>(dst, src, 0u);
    else st_shared16(dst, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const float* q = KC ? src + (int64_t)r * ld + k : src + (int64_t)k * ld + r;
  const int left = KC ? klim - k : rlim - r;
  if (vec && left >= 4) cp_async16<false>(dst, q, 16u);
  else st_shared16(dst, fetch4<KC>(src, ld, r, k, rlim, klim, false));
}

// Piece `idx` (0 .. ROWS*8-1) of a ROWS x 32 stage tile: its (row, k) origin and the byte
// offset of its 16-byte slot in the operand image.
template <bool KC, int ROWS>
__device__ __forceinline__ uint32_t piece(int idx, int& r, int& k) {
  if (KC) {
    // K-major SWIZZLE_128B: row r is one 128-byte line (32 K-elements), its 16-byte chunk c
    // sits at position c ^ (r & 7).  Eight lanes take one row: 128 contiguous bytes of
    // global memory land in one conflict-free shared-memory line.
    r = idx >> 3;
    const int chunk = idx & 7;
    k = chunk * 4;
    return (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
  } else {
    const int mn4 = idx % (ROWS / 4);
    k = idx / (ROWS / 4);
    r = mn4 * 4;
    return mn32_offset(r, k, kMnLbo, kMnSbo);
  }
}

// ---- implicit 3x3 'same' convolution: the A operand gathered from an NHWC tensor --------
// One element of the logical im2col matrix (pixel row pr, column kc = (ky, kx, c)); the
// slow path of channel counts that are not multiples of 4 (the first convolution: C*T = 6).
__device__ __forceinline__ float conv_elem(const GemmP& p, int pr, int kc) {
  const int C = p.conv_C, W = p.conv_W, H = p.conv_H;
  const int tap = kc / C, c = kc - tap * C;
  const int ky = (tap * 11) >> 5, kx = tap - ky * 3;
  const int x = pr % W, y = (pr / W) % H;
  const int yy = y + ky - 1, xx = x + kx - 1;
  if ((unsigned)yy >= (unsigned)H || (unsigned)xx >= (unsigned)W) return 0.f;
  return __ldg(p.A + ((int64_t)pr + (ky - 1) * W + (kx - 1)) * C + c);
}
__device__ __forceinline__ float4 conv_scalar4(const GemmP& p, int pr, int kc, int prlim, int kclim) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pr >= prlim) return v;
  if (kc < kclim) v.x = conv_elem(p, pr, kc);
  if (kc + 1 < kclim) v.y = conv_elem(p, pr, kc + 1);
  if (kc + 2 < kclim) v.z = conv_elem(p, pr, kc + 2);
  if (kc + 3 < kclim) v.w = conv_elem(p, pr, kc + 3);
  return v;
}
// Walking state of one 16-byte piece (4 consecutive channels of one tap at one pixel):
// the loader advances either the column (forward / data-gradient form: 32 columns per
// stage) or the pixel row (weight-gradient form: 32 rows per stage) without divisions.
struct ConvPiece {
  int x, y;        // pixel
  int tap, c;      // column = tap * C + c
  int pr;          // pixel row index
  bool ok;         // the fixed coordinate is inside the matrix
};
__device__ __forceinline__ ConvPiece conv_piece(const GemmP& p, int pr, int kc) {
  ConvPiece s;
  s.pr = pr;
  s.x = pr % p.conv_W; s.y = (pr / p.conv_W) % p.conv_H;
  s.tap = kc / p.conv_C; s.c = kc - s.tap * p.conv_C;
  s.ok = true;
  return s;
}
__device__ __forceinline__ void conv_issue4(uint32_t dst, const GemmP& p, const ConvPiece& s,
                                            bool inside) {
  const int ky = (s.tap * 11) >> 5, kx = s.tap - ky * 3;
  const int yy = s.y + ky - 1, xx = s.x + kx - 1;
  if (!inside || !s.ok || (unsigned)yy >= (unsigned)p.conv_H || (unsigned)xx >= (unsigned)p.conv_W)
    cp_async16<true>(dst, p.A, 0u);
  else   // neighbouring taps re-read the same pixels: keep them in L1
    cp_async16<true>(dst, p.A + ((int64_t)s.pr + (ky - 1) * p.conv_W + (kx - 1)) * p.conv_C + s.c, 16u);
}

template <bool TA, bool TB, int BN, bool CONV>
__global__ void __launch_bounds__(kThreads, 2) gemm_tc_kernel(GemmP p, int* status) {
  using C = TcCfg<BN>;
  constexpr int S = C::kStages;
  constexpr bool AKC = !TA, BKC = TB;
  constexpr int NA = kBM / 32, NB = BN / 32;     // 16-byte pieces per loader thread and stage
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>(
      ((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + C::kRing);
  uint64_t* empty = full + S;
  uint64_t* d_ready = empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 1);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  const int kc = p.k_chunk > 0 ? p.k_chunk : p.K;
  const int kbeg = blockIdx.z * kc;
  const int kend = min(p.K, kbeg + kc);
  const int nks = (kend - kbeg + kBK - 1) / kBK;

  if (tid == 0) {
    s_abort = 0;
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], kLoaderWarps); mbar_init(&empty[s], 1); }
    mbar_init(d_ready, 1);
    fence_mbar_init();
  }
  if (warp == kLoaderWarps) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;

  if (warp < kLoaderWarps) {
    // ===================== loaders ===================================================
    const bool avec = ((p.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) &&
                      (p.aop != AOP_GATE || (reinterpret_cast<uintptr_t>(p.gateA) & 15) == 0);
    const bool bvec = ((p.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);
    // implicit convolution: per-piece walking state (columns advance in the forward /
    // data-gradient form, pixel rows in the weight-gradient form)
    const bool cvec = CONV && ((p.conv_C & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    ConvPiece cs[CONV ? NA : 1];
    if (CONV && cvec) {
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        int r, k;
        piece<AKC, kBM>(j * kLoaderThreads + tid, r, k);
        if (AKC) { cs[j] = conv_piece(p, m0 + r, kbeg + k); cs[j].ok = (m0 + r) < p.M; }
        else { cs[j] = conv_piece(p, kbeg + k, m0 + r); cs[j].ok = (m0 + r) < p.M; }
      }
    }
    const uint32_t ring_addr = smem_u32(ring);
    // Fast path of whole K-slices (every slice but a ragged last one): per piece, the
    // source of slice 0, the shared-memory offset and whether the piece is inside the
    // matrix are fixed for the thread, so a slice costs one address add and one copy
    // per piece (the generic path re-derives rows, limits and alignment: ~30
    // instructions per piece, which made the loaders issue-bound; ncu r1).
    const float* pa_src[NA]; uint32_t pa_dst[NA]; uint32_t pa_ok = 0, pa_fast = 1;
    const float* pb_src[NB]; uint32_t pb_dst[NB]; uint32_t pb_ok = 0, pb_fast = 1;
    const int64_t a_step = AKC ? (int64_t)kBK : (int64_t)kBK * p.lda;
    const int64_t b_step = BKC ? (int64_t)kBK : (int64_t)kBK * p.ldb;
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      int r, k;
      pa_dst[j] = piece<AKC, kBM>(j * kLoaderThreads + tid, r, k);
      const bool in = (m0 + r) < p.M;
      if (!AKC && in && m0 + r + 4 > p.M) pa_fast = 0;       // ragged 16-byte piece
      if (in) pa_ok |= 1u << j;
      pa_src[j] = !in ? p.A
                      : (AKC ? p.A + (int64_t)(m0 + r) * p.lda + kbeg + k
                             : p.A + (int64_t)(kbeg + k) * p.lda + m0 + r);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      int r, k;
      pb_dst[j] = C::kABytes + piece<BKC, BN>(j * kLoaderThreads + tid, r, k);
      const bool in = (n0 + r) < p.N;
      if (!BKC && in && n0 + r + 4 > p.N) pb_fast = 0;
      if (in) pb_ok |= 1u << j;
      pb_src[j] = !in ? p.B
                      : (BKC ? p.B + (int64_t)(n0 + r) * p.ldb + kbeg + k
                             : p.B + (int64_t)(kbeg + k) * p.ldb + n0 + r);
    }
    if (!avec || CONV) pa_fast = 0;
    if (!bvec) pb_fast = 0;
    // Stage ks on its way: asynchronous copies straight into the operand images.
    auto issue = [&](int ks) {
      const int k0 = kbeg + ks * kBK;
      const uint32_t sa = ring_addr + (uint32_t)((ks % S) * C::kStageBytes);
      const uint32_t sb = sa + C::kABytes;
      const bool whole = k0 + kBK <= kend;
      if (whole && pa_fast) {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          const bool in = (pa_ok >> j) & 1u;
          cp_async16<false>(sa + pa_dst[j], in ? pa_src[j] + (int64_t)ks * a_step : p.A,
                            in ? 16u : 0u);
        }
      }
      if (whole && pb_fast) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const bool in = (pb_ok >> j) & 1u;
          cp_async16<false>(sa + pb_dst[j], in ? pb_src[j] + (int64_t)ks * b_step : p.B,
                            in ? 16u : 0u);
        }
      }
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        if (whole && pa_fast) break;
        int r, k;
        const uint32_t dst = sa + piece<AKC, kBM>(j * kLoaderThreads + tid, r, k);
        if (CONV) {
          if (cvec) {
            conv_issue4(dst, p, cs[j], k0 + k < kend);
            if (AKC) {          // next stage: 32 columns further
              cs[j].c += kBK;
              while (cs[j].c >= p.conv_C) { cs[j].c -= p.conv_C; ++cs[j].tap; }
            } else {            // next stage: 32 pixel rows further
              cs[j].pr += kBK; cs[j].x += kBK;
              while (cs[j].x >= p.conv_W) {
                cs[j].x -= p.conv_W;
                if (++cs[j].y == p.conv_H) cs[j].y = 0;
              }
            }
          } else {
            st_shared16(dst, AKC ? conv_scalar4(p, m0 + r, k0 + k, p.M, kend)
                                 : conv_scalar4(p, k0 + k, m0 + r, kend, p.M));
          }
        } else {
          issue4<AKC>(dst, p.A, p.lda, m0 + r, k0 + k, p.M, kend, avec);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (whole && pb_fast) break;
        int r, k;
        const uint32_t dst = sb + piece<BKC, BN>(j * kLoaderThreads + tid, r, k);
        issue4<BKC>(dst, p.B, p.ldb, n0 + r, k0 + k, p.N, kend, bvec);
      }
    };
    // Stage ks has landed: each thread applies the prologue to the pieces it copied
    // itself (ReLU / gate on A, round-to-nearest tf32 on both), in place.
    const bool fin_a = !(p.a_tf32 && p.aop == AOP_ID), fin_b = !p.b_tf32;
    // Round-to-nearest tf32 is one integer add per element (the MMA truncates the low 13
    // bits), ReLU + rounding one VIADDMNMX (umma_pack.cuh).
    auto finish = [&](int ks) {
      const uint32_t sa = ring_addr + (uint32_t)((ks % S) * C::kStageBytes);
      if (fin_a) {
        if (p.aop == AOP_GATE) {
          const int k0 = kbeg + ks * kBK;
#pragma unroll
          for (int j = 0; j < NA; ++j) {
            int r, k;
            piece<AKC, kBM>(j * kLoaderThreads + tid, r, k);
            float4 v = ld_shared16(sa + pa_dst[j]);
            const float4 g = fetch4<AKC>(p.gateA, p.lda, m0 + r, k0 + k, p.M, kend, avec);
            if (!(g.x > 0.f)) v.x = 0.f;
            if (!(g.y > 0.f)) v.y = 0.f;
            if (!(g.z > 0.f)) v.z = 0.f;
            if (!(g.w > 0.f)) v.w = 0.f;
            st_shared16(sa + pa_dst[j], tf32_rn4(v));
          }
        } else if (p.aop == AOP_RELU) {
#pragma unroll
          for (int j = 0; j < NA; ++j) {
            const float4 v = ld_shared16(sa + pa_dst[j]);
            st_shared16(sa + pa_dst[j],
                        make_float4(__uint_as_float(relu_tf32_bits(__float_as_uint(v.x))),
                                    __uint_as_float(relu_tf32_bits(__float_as_uint(v.y))),
                                    __uint_as_float(relu_tf32_bits(__float_as_uint(v.z))),
                                    __uint_as_float(relu_tf32_bits(__float_as_uint(v.w)))));
          }
        } else {
#pragma unroll
          for (int j = 0; j < NA; ++j) {
            const float4 v = ld_shared16(sa + pa_dst[j]);
            st_shared16(sa + pa_dst[j],
                        make_float4(__uint_as_float(tf32_rn_bits(__float_as_uint(v.x))),
                                    __uint_as_float(tf32_rn_bits(__float_as_uint(v.y))),
                                    __uint_as_float(tf32_rn_bits(__float_as_uint(v.z))),
                                    __uint_as_float(tf32_rn_bits(__float_as_uint(v.w)))));
          }
        }
      }
      if (fin_b) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const float4 v = ld_shared16(sa + pb_dst[j]);
          st_shared16(sa + pb_dst[j],
                      make_float4(__uint_as_float(tf32_rn_bits(__float_as_uint(v.x))),
                                  __uint_as_float(tf32_rn_bits(__float_as_uint(v.y))),
                                  __uint_as_float(tf32_rn_bits(__float_as_uint(v.z))),
                                  __uint_as_float(tf32_rn_bits(__float_as_uint(v.w)))));
        }
      }
    };
    for (int ks = 0; ks < S - 1; ++ks) {
      if (ks < nks) issue(ks);
      cp_async_commit();
    }
    for (int ks = 0; ks < nks; ++ks) {
      cp_async_wait<S - 2>();            // this thread's copies of stage ks are complete
      finish(ks);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[ks % S]);
      // refill the stage MMA ks-1 read (it ran while this stage was being finished)
      const int nxt = ks + S - 1;
      if (nxt < nks) {
        tc_wait(&empty[nxt % S], ((nxt / S) & 1) ^ 1, abort_flag, status, 930);
        issue(nxt);
      }
      cp_async_commit();
    }

    // ===================== epilogue ==================================================
    if (nks > 0) tc_wait(d_ready, 0, abort_flag, status, 931);
    tc_fence_after();
    const int q = warp & 3, hh = warp >> 2;
    float* stg = reinterpret_cast<float*>(ring) + warp * (32 * 33);
    const bool first = (blockIdx.z == 0);
    const int rpg = (int)p.rows_per_group;
    for (int c32 = hh; c32 < NB; c32 += 2) {
      uint32_t v[32];
      if (nks > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c32 * 32), v);
        tmem_ld_wait(v);
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = 0u;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) stg[lane * 33 + c] = __uint_as_float(v[c]);
      __syncwarp();
      const int n = n0 + c32 * 32 + lane;
      const bool nok = n < p.N;
      const float bias_n = (first && p.bias && nok) ? p.bias[n] : 0.f;
      const int mrows = min(32, p.M - (m0 + q * 32));
      for (int rr = 0; rr < mrows; ++rr) {
        const int m = m0 + q * 32 + rr;
        float x = stg[rr * 33 + lane] * p.alpha;
        if (nok) {
          if (first) {
            x += bias_n;
            if (p.rowbias) x += p.rowbias[(int64_t)(m / rpg) * p.ldrb + n];
          }
          if (p.gateC) x = (p.gateC[(int64_t)m * p.ldgc + n] > 0.f) ? x : 0.f;
          if (first && p.res) x += p.res[(int64_t)m * p.ldres + n];
          if (p.relu_out) x = fmaxf(x, 0.f);
          float* c = p.C + (int64_t)m * p.ldc + n;
          if (p.atomic) atomicAdd(c, x); else *c = x;
        }
      }
      __syncwarp();
    }
  } else {
    // ===================== MMA issuer ================================================
    constexpr uint32_t idesc = idesc_tf32(kBM, BN) | (TA ? kIdescAMajorMN : 0u) |
                               (TB ? 0u : kIdescBMajorMN);
    // one K = 8 step: 32 bytes along the swizzled line (K-major), two 512-byte K atoms (MN-major)
    constexpr uint32_t a_step16 = AKC ? 32u >> 4 : 1024u >> 4;
    constexpr uint32_t b_step16 = BKC ? 32u >> 4 : 1024u >> 4;
    for (int ks = 0; ks < nks; ++ks) {
      const int st = ks % S;
      if (elect_one()) {
        tc_wait(&full[st], (ks / S) & 1, abort_flag, status, 932);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(ring + st * C::kStageBytes);
        const uint32_t b_addr = a_addr + C::kABytes;
        const uint64_t ad = AKC ? smem_desc_lt(a_addr, 16, 1024, 2)
                                : smem_desc_lt(a_addr, kMnLbo, kMnSbo, 1);
        const uint64_t bd = BKC ? smem_desc_lt(b_addr, 16, 1024, 2)
                                : smem_desc_lt(b_addr, kMnLbo, kMnSbo, 1);
        mma_tf32_ss_steps<kBK / 8>(tmem_base, ad, bd, a_step16, b_step16, idesc,
                                   ks == 0 ? 0u : 1u);
        mma_commit(&empty[st]);
        if (ks == nks - 1) mma_commit(d_ready);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kLoaderWarps) tmem_dealloc(tmem_base, BN);
}

template <bool TA, bool TB, int BN, bool CONV>
int launch_tc(ibc_handle* h, const GemmP& p, dim3 grid, cudaStream_t s) {
  using C = TcCfg<BN>;
  static bool attr = false;
  if (!attr) {
    IBC_CUDA(h, cudaFuncSetAttribute(gemm_tc_kernel<TA, TB, BN, CONV>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    attr = true;
  }
  gemm_tc_kernel<TA, TB, BN, CONV><<<grid, kThreads, C::kSmem, s>>>(p, h ? h->dev_status : nullptr);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

template <bool TA, bool TB, bool CONV>
int launch_tc_bn(ibc_handle* h, const GemmP& p, int bn, dim3 grid, cudaStream_t s) {
  if (bn == 32) return launch_tc<TA, TB, 32, CONV>(h, p, grid, s);
  if (bn == 64) return launch_tc<TA, TB, 64, CONV>(h, p, grid, s);
  return launch_tc<TA, TB, 128, CONV>(h, p, grid, s);
}

}  // namespace

bool gemm_tc_supported(const GemmP& p) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return false;
  if (p.k_chunk > 0 && p.k_chunk < p.K && !p.atomic) return false;
  if (p.conv_C > 0 && (p.conv_H < 1 || p.conv_W < 1 || p.aop == AOP_GATE)) return false;
  return true;
}

int gemm_tc(ibc_handle* h, const GemmP& p_in, bool ta, bool tb, cudaStream_t s) {
  if (!gemm_tc_supported(p_in)) IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "gemm_tc: unsupported shape");
  GemmP p = p_in;
  const int bn = p.N <= 32 ? 32 : (p.N <= 64 ? 64 : 128);
  const int tm = (p.M + kBM - 1) / kBM, tn = (p.N + bn - 1) / bn;
  int splits = 1;
  if (p.atomic) {
    // split K so that the (usually few) output blocks of a weight gradient fill the GPU
    const int sms = h ? h->sm_count : 148;
    const int64_t tiles = (int64_t)tm * tn;
    int64_t want = (4 * (int64_t)sms + tiles - 1) / tiles;
    const int64_t max_splits = (p.K + 255) / 256;          // at least 8 stages per split
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    int kc = (int)((p.K + want - 1) / want);
    kc = (kc + kBK - 1) / kBK * kBK;
    p.k_chunk = kc;
    splits = (p.K + kc - 1) / kc;
  } else {
    p.k_chunk = 0;
  }
  if (tn > 65535 || splits > 65535) IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "gemm_tc: grid too large");
  dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)splits);
  if (p.conv_C > 0) {       // implicit convolution: the weights are always [K, N] row-major
    if (tb) IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "gemm_tc: implicit convolution takes B as [K, N]");
    const int kdim = ta ? p.M : p.K;
    if (kdim != 9 * p.conv_C) IBC_FAIL(h, IBC_ERR_INVALID, "gemm_tc: convolution needs 9 * C columns");
    if (ta) return launch_tc_bn<true, false, true>(h, p, bn, grid, s);
    return launch_tc_bn<false, false, true>(h, p, bn, grid, s);
  }
  if (ta && tb) return launch_tc_bn<true, true, false>(h, p, bn, grid, s);
  if (ta) return launch_tc_bn<true, false, false>(h, p, bn, grid, s);
  if (tb) return launch_tc_bn<false, true, false>(h, p, bn, grid, s);
  return launch_tc_bn<false, false, false>(h, p, bn, grid, s);
}

}  // namespace ibc
