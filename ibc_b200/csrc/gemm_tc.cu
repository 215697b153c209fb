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
#include <stdlib.h>

#include "gemm_simt.cuh"
#include "umma.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kBM = 128, kBK = 32;
constexpr int kLoaderWarps = 16;
constexpr int kLoaderThreads = kLoaderWarps * 32;
constexpr int kIssuerWarp = kLoaderWarps;           // one MMA-issuer warp
constexpr int kFirstEpiWarp = kLoaderWarps + 1;     // eight epilogue warps (two per TMEM lane quarter)
constexpr int kEpiWarps = 8;
constexpr int kEpiLd = 36;                          // floats per row of an epilogue staging tile
constexpr int kThreads = (kLoaderWarps + 1 + kEpiWarps) * 32;
constexpr uint32_t kMnLbo = (kBK / 4) * 512u;   // between 32-element MN atoms of a stage
constexpr uint32_t kMnSbo = 512u;               // between 4-row K atoms

template <int BN>
struct TcCfg {
  static constexpr int kStages = BN > 64 ? 5 : 6;   // 32 KB stages at BN = 128: 160 KB + 36 KB epilogue staging
  static constexpr int kABytes = kBM * kBK * 4;
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRing = kStages * kStageBytes;
  // accumulators in tensor memory: the issuer runs that many output blocks ahead of the epilogue
  static constexpr int kBufs = BN > 64 ? 2 : 4;
  static constexpr int kEpiStage = kEpiWarps * 32 * kEpiLd * 4;   // epilogue transposing stage
  static constexpr size_t kSmem = (size_t)kRing + kEpiStage + 1024 + 256;
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
// Vector gather (channel counts that are multiples of 4): everything that does not change
// from slice to slice is worked out once per output block, so a piece costs a mask test, an
// add and the copy (the first version re-derived taps, bounds and a 64-bit product per piece:
// the large convolutions were bound by instruction issue, profiles/r1_ncu_full_pixel_step_gemms.csv).
//   forward / data-gradient form (rows = pixels, K walks (tap, channel)): per piece the
//     pixel's base pointer and a 9-bit mask of the taps that fall inside the frame; per
//     thread the (tap, channel) of the slice -- all pieces of a thread share the K chunk;
//   weight-gradient form (rows = (tap, channel), K walks pixels): per piece the tap's
//     (dy, dx) and element offset, and a walking (x, y, row pointer).
struct ConvRow {            // forward form, per piece
  const float* pix;         // &X[pixel, 0]
  uint32_t taps;            // bit t: tap t of this pixel is inside the frame (0: row outside M)
};
struct ConvCol {            // weight-gradient form, per piece
  const float* ptr;         // &X[pixel row of the current slice, 0] + tap / channel offset
  int x, y, dy, dx;
  bool ok;                  // the (tap, channel) column is inside the matrix
};
__device__ __forceinline__ uint32_t conv_tap_mask(int x, int y, int W, int H) {
  const uint32_t colm = (x > 0 ? 1u : 0u) | 2u | (x < W - 1 ? 4u : 0u);
  return (y > 0 ? colm : 0u) | (colm << 3) | (y < H - 1 ? colm << 6 : 0u);
}

// One output block of the persistent schedule.
struct TileCoord {
  int m0, n0, kbeg, kend, nks;
  bool first;        // first K split: bias / row bias / residual are added here
};
template <int BN>
__device__ __forceinline__ TileCoord tile_coord(const GemmP& p, int t, int tm, int tn) {
  TileCoord c;
  const int per = tm * tn;
  const int split = t / per, rem = t - split * per;
  const int nt = rem / tm, mt = rem - nt * tm;      // consecutive blocks share the B tile
  const int kc = p.k_chunk > 0 ? p.k_chunk : p.K;
  c.m0 = mt * kBM; c.n0 = nt * BN;
  c.kbeg = split * kc; c.kend = min(p.K, c.kbeg + kc);
  c.nks = (c.kend - c.kbeg + kBK - 1) / kBK;
  c.first = split == 0;
  return c;
}

// Persistent: CTA b walks blocks b, b + grid, ... with three sets of warps running ahead
// of one another -- loaders (K-slices of the next blocks are already in flight while a
// block finishes), the MMA issuer (2-4 accumulators in tensor memory) and the epilogue.
template <bool TA, bool TB, int BN, bool CONV>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(GemmP p, int* status, int tm, int tn,
                                                              int total, int dbg) {
  using C = TcCfg<BN>;
  constexpr int S = C::kStages;
  constexpr bool AKC = !TA, BKC = TB;
  // 16-byte pieces per loader thread and stage (B: the last threads idle when BN = 32)
  constexpr int NA = kBM * 8 / kLoaderThreads, NB = (BN * 8 + kLoaderThreads - 1) / kLoaderThreads;
  constexpr bool kBAll = (BN * 8) % kLoaderThreads == 0;   // every thread has B pieces
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>(
      ((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* epi_stage = reinterpret_cast<float*>(ring + C::kRing);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + C::kRing + C::kEpiStage);
  uint64_t* empty = full + S;
  constexpr int NBUF = C::kBufs;
  uint64_t* acc_full = empty + S;            // [NBUF] issuer -> epilogue
  uint64_t* acc_empty = acc_full + NBUF;     // [NBUF] epilogue -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NBUF);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    s_abort = 0;
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], kLoaderWarps); mbar_init(&empty[s], 1); }
    for (int s = 0; s < NBUF; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kIssuerWarp) tmem_alloc(tmem_slot, NBUF * BN);
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
    const bool cvec = CONV && ((p.conv_C & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const uint32_t ring_addr = smem_u32(ring);
    const int64_t a_step = AKC ? (int64_t)kBK : (int64_t)kBK * p.lda;
    const int64_t b_step = BKC ? (int64_t)kBK : (int64_t)kBK * p.ldb;
    // fixed per thread: where its pieces sit in a stage
    uint32_t pa_dst[NA], pb_dst[NB];
#pragma unroll
    for (int j = 0; j < NA; ++j) { int r, k; pa_dst[j] = piece<AKC, kBM>(j * kLoaderThreads + tid, r, k); }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      int r, k;
      pb_dst[j] = C::kABytes + piece<BKC, BN>(j * kLoaderThreads + tid, r, k);
    }
    const bool has_b = kBAll || tid < BN * 8;      // (NB == 1 whenever some threads have none)
    // Per block (issue side).  Fast path of whole K-slices: per piece, the source of slice
    // 0 and whether the piece is inside the matrix, so a slice costs one address add and
    // one copy per piece (the generic path re-derives rows, limits and alignment: ~30
    // instructions per piece, which made the loaders issue-bound; ncu r1).  Implicit
    // convolution: per-piece walking state (columns advance in the forward / data-gradient
    // form, pixel rows in the weight-gradient form).
    const float* pa_src[NA]; uint32_t pa_ok = 0, pa_fast = 0;
    const float* pb_src[NB]; uint32_t pb_ok = 0, pb_fast = 0;
    ConvRow crow[CONV && AKC ? NA : 1];
    ConvCol ccol[CONV && !AKC ? NA : 1];
    int ctap = 0, cch = 0;      // forward form: (tap, channel) of this thread's chunk in the next slice
    auto setup = [&](const TileCoord& c) {
      pa_ok = 0; pb_ok = 0; pa_fast = (avec && !CONV) ? 1u : 0u; pb_fast = bvec ? 1u : 0u;
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        int r, k;
        piece<AKC, kBM>(j * kLoaderThreads + tid, r, k);
        const bool in = (c.m0 + r) < p.M;
        if (!AKC && in && c.m0 + r + 4 > p.M) pa_fast = 0;       // ragged 16-byte piece
        if (in) pa_ok |= 1u << j;
        if (CONV) {
          if (cvec && AKC) {
            const int pr = c.m0 + r;
            crow[j].pix = p.A + (int64_t)pr * p.conv_C;
            crow[j].taps = in ? conv_tap_mask(pr % p.conv_W, (pr / p.conv_W) % p.conv_H,
                                              p.conv_W, p.conv_H) : 0u;
            const int kc0 = c.kbeg + k;                  // the same chunk for all of a thread's pieces
            ctap = kc0 / p.conv_C; cch = kc0 - ctap * p.conv_C;
          } else if (cvec) {
            const int col = c.m0 + r, pr = c.kbeg + k;
            const int tap = col / p.conv_C, ch = col - tap * p.conv_C;
            const int ky = (tap * 11) >> 5, kx = tap - ky * 3;
            ccol[j].dy = ky - 1; ccol[j].dx = kx - 1;
            ccol[j].x = pr % p.conv_W; ccol[j].y = (pr / p.conv_W) % p.conv_H;
            ccol[j].ptr = p.A + ((int64_t)pr + (ky - 1) * p.conv_W + (kx - 1)) * p.conv_C + ch;
            ccol[j].ok = in;
          }
        } else {
          pa_src[j] = !in ? p.A
                          : (AKC ? p.A + (int64_t)(c.m0 + r) * p.lda + c.kbeg + k
                                 : p.A + (int64_t)(c.kbeg + k) * p.lda + c.m0 + r);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        int r, k;
        piece<BKC, BN>(j * kLoaderThreads + tid, r, k);
        const bool in = has_b && (c.n0 + r) < p.N;
        if (!BKC && in && c.n0 + r + 4 > p.N) pb_fast = 0;
        if (in) pb_ok |= 1u << j;
        pb_src[j] = !in ? p.B
                        : (BKC ? p.B + (int64_t)(c.n0 + r) * p.ldb + c.kbeg + k
                               : p.B + (int64_t)(c.kbeg + k) * p.ldb + c.n0 + r);
      }
    };
    // K-slice ks of block c on its way into stage st: asynchronous copies straight into
    // the operand images.
    auto issue = [&](const TileCoord& c, int ks, int st) {
      const int k0 = c.kbeg + ks * kBK;
      const uint32_t sa = ring_addr + (uint32_t)(st * C::kStageBytes);
      const bool whole = k0 + kBK <= c.kend;
      if (whole && pa_fast) {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          const bool in = (pa_ok >> j) & 1u;
          cp_async16<false>(sa + pa_dst[j], in ? pa_src[j] + (int64_t)ks * a_step : p.A,
                            in ? 16u : 0u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          int r, k;
          piece<AKC, kBM>(j * kLoaderThreads + tid, r, k);
          const uint32_t dst = sa + pa_dst[j];
          if (CONV) {
            if (cvec && AKC) {
              const bool ok = (k0 + k < c.kend) && ((crow[j].taps >> ctap) & 1u);
              const int ky = (ctap * 11) >> 5, kx = ctap - ky * 3;
              const int off = ((ky - 1) * p.conv_W + (kx - 1)) * p.conv_C + cch;
              cp_async16<true>(dst, ok ? crow[j].pix + off : p.A, ok ? 16u : 0u);
            } else if (cvec) {
              ConvCol& q = ccol[j];
              const bool ok = q.ok && (k0 + k < c.kend) &&
                              (unsigned)(q.y + q.dy) < (unsigned)p.conv_H &&
                              (unsigned)(q.x + q.dx) < (unsigned)p.conv_W;
              cp_async16<true>(dst, ok ? q.ptr : p.A, ok ? 16u : 0u);
              q.ptr += (int64_t)kBK * p.conv_C;       // next slice: 32 pixel rows further
              q.x += kBK;
              while (q.x >= p.conv_W) {
                q.x -= p.conv_W;
                if (++q.y == p.conv_H) q.y = 0;
              }
            } else {
              st_shared16(dst, AKC ? conv_scalar4(p, c.m0 + r, k0 + k, p.M, c.kend)
                                   : conv_scalar4(p, k0 + k, c.m0 + r, c.kend, p.M));
            }
          } else {
            issue4<AKC>(dst, p.A, p.lda, c.m0 + r, k0 + k, p.M, c.kend, avec);
          }
        }
        if (CONV && AKC && cvec) {       // next slice: 32 columns further
          cch += kBK;
          while (cch >= p.conv_C) { cch -= p.conv_C; ++ctap; }
        }
      }
      if (!has_b) {
      } else if (whole && pb_fast) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const bool in = (pb_ok >> j) & 1u;
          cp_async16<false>(sa + pb_dst[j], in ? pb_src[j] + (int64_t)ks * b_step : p.B,
                            in ? 16u : 0u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          int r, k;
          piece<BKC, BN>(j * kLoaderThreads + tid, r, k);
          issue4<BKC>(sa + pb_dst[j], p.B, p.ldb, c.n0 + r, k0 + k, p.N, c.kend, bvec);
        }
      }
    };
    // A slice has landed: each thread applies the prologue to the pieces it copied itself
    // (ReLU / gate on A, round-to-nearest tf32 on both), in place.  Rounding is one
    // integer add per element (the MMA truncates the low 13 bits), ReLU + rounding one
    // VIADDMNMX (umma_pack.cuh).
    const bool fin_a = !(p.a_tf32 && p.aop == AOP_ID), fin_b = !p.b_tf32;
    auto round4 = [](float4 v) {
      return make_float4(__uint_as_float(tf32_rn_bits(__float_as_uint(v.x))),
                         __uint_as_float(tf32_rn_bits(__float_as_uint(v.y))),
                         __uint_as_float(tf32_rn_bits(__float_as_uint(v.z))),
                         __uint_as_float(tf32_rn_bits(__float_as_uint(v.w))));
    };
    auto finish = [&](const TileCoord& c, int ks, int st) {
      const uint32_t sa = ring_addr + (uint32_t)(st * C::kStageBytes);
      if (fin_a) {
        if (p.aop == AOP_GATE) {
          const int k0 = c.kbeg + ks * kBK;
#pragma unroll
          for (int j = 0; j < NA; ++j) {
            int r, k;
            piece<AKC, kBM>(j * kLoaderThreads + tid, r, k);
            float4 v = ld_shared16(sa + pa_dst[j]);
            const float4 g = fetch4<AKC>(p.gateA, p.lda, c.m0 + r, k0 + k, p.M, c.kend, avec);
            if (!(g.x > 0.f)) v.x = 0.f;
            if (!(g.y > 0.f)) v.y = 0.f;
            if (!(g.z > 0.f)) v.z = 0.f;
            if (!(g.w > 0.f)) v.w = 0.f;
            st_shared16(sa + pa_dst[j], round4(v));
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
          for (int j = 0; j < NA; ++j) st_shared16(sa + pa_dst[j], round4(ld_shared16(sa + pa_dst[j])));
        }
      }
      if (fin_b && has_b) {
#pragma unroll
        for (int j = 0; j < NB; ++j) st_shared16(sa + pb_dst[j], round4(ld_shared16(sa + pb_dst[j])));
      }
    };

    // issue cursor (runs S - 1 slices ahead, across block boundaries) and finish cursor
    int ti = blockIdx.x, ks_i = 0;
    bool have_i = ti < total;
    TileCoord ci = tile_coord<BN>(p, have_i ? ti : 0, tm, tn);
    if (have_i) setup(ci);
    int tf = ti, ks_f = 0;
    bool have_f = have_i;
    TileCoord cf = ci;
    uint32_t it_i = 0, it_f = 0;
    auto issue_next = [&]() {
      if (!(dbg & 1)) issue(ci, ks_i, (int)(it_i % S));
      ++it_i;
      if (++ks_i == ci.nks) {
        ti += gridDim.x; ks_i = 0;
        have_i = ti < total;
        if (have_i) { ci = tile_coord<BN>(p, ti, tm, tn); setup(ci); }
      }
    };
    for (int s = 0; s < S - 1; ++s) {
      if (have_i) issue_next();
      cp_async_commit();
    }
    while (have_f) {
      cp_async_wait<S - 2>();            // this thread's copies of the oldest slice are complete
      const int st = (int)(it_f % S);
      if (!(dbg & 1)) finish(cf, ks_f, st);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
      ++it_f;
      if (++ks_f == cf.nks) {
        tf += gridDim.x; ks_f = 0;
        have_f = tf < total;
        if (have_f) cf = tile_coord<BN>(p, tf, tm, tn);
      }
      // refill the stage the previous slice's MMAs read (they ran during this finish)
      if (have_i) {
        tc_wait(&empty[it_i % S], ((it_i / S) & 1) ^ 1, abort_flag, status, 930);
        issue_next();
      }
      cp_async_commit();
    }
  } else if (warp == kIssuerWarp) {
    // ===================== MMA issuer ================================================
    constexpr uint32_t idesc = idesc_tf32(kBM, BN) | (TA ? kIdescAMajorMN : 0u) |
                               (TB ? 0u : kIdescBMajorMN);
    // one K = 8 step: 32 bytes along the swizzled line (K-major), two 512-byte K atoms (MN-major)
    constexpr uint32_t a_step16 = AKC ? 32u >> 4 : 1024u >> 4;
    constexpr uint32_t b_step16 = BKC ? 32u >> 4 : 1024u >> 4;
    uint32_t it = 0;
    int li = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++li) {
      const TileCoord c = tile_coord<BN>(p, t, tm, tn);
      const int buf = li % NBUF;
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
      for (int ks = 0; ks < c.nks; ++ks, ++it) {
        const int st = (int)(it % S);
        if (elect_one()) {
          if (ks == 0)      // the epilogue has drained this accumulator (NBUF blocks ago)
            tc_wait(&acc_empty[buf], ((li / NBUF) & 1) ^ 1, abort_flag, status, 933);
          tc_wait(&full[st], (it / S) & 1, abort_flag, status, 932);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + st * C::kStageBytes);
          const uint32_t b_addr = a_addr + C::kABytes;
          const uint64_t ad = AKC ? smem_desc_lt(a_addr, 16, 1024, 2)
                                  : smem_desc_lt(a_addr, kMnLbo, kMnSbo, 1);
          const uint64_t bd = BKC ? smem_desc_lt(b_addr, 16, 1024, 2)
                                  : smem_desc_lt(b_addr, kMnLbo, kMnSbo, 1);
          if (!(dbg & 2))
            mma_tf32_ss_steps<kBK / 8>(d_tmem, ad, bd, a_step16, b_step16, idesc, ks == 0 ? 0u : 1u);
          mma_commit(&empty[st]);
          if (ks == c.nks - 1) mma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue ==================================================
    // Two warps per accumulator lane quarter (32 rows) share its 32-column chunks.  A chunk
    // is re-read from tensor memory (thread = row), transposed through shared memory, and
    // leaves as 16-byte pieces: a warp instruction covers 4 rows x 128 contiguous bytes of
    // C (and of the gate / residual operands).
    const int ew = warp - kFirstEpiWarp;
    const int q = warp & 3, hh = ew >> 2;
    float* stg = epi_stage + ew * (32 * kEpiLd);
    const int rpg = (int)p.rows_per_group;
    const bool vec_c = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                       (!p.gateC || (((p.ldgc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.gateC) & 15) == 0))) &&
                       (!p.res || (((p.ldres & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0))) &&
                       (!p.rowbias || (((p.ldrb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.rowbias) & 15) == 0))) &&
                       (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
    int li = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++li) {
      const TileCoord c = tile_coord<BN>(p, t, tm, tn);
      const int buf = li % NBUF;
      // a whole main loop away: wait with the suspend-time hint (umma::mbar_wait) so that the
      // 8 idle warps do not take issue slots from the loaders by polling
      mbar_wait(&acc_full[buf], (li / NBUF) & 1, abort_flag, status, 931);
      tc_fence_after();
      constexpr int kChunks = BN / 32;
      // The two warps of a lane quarter take alternate (block, chunk) units: with one chunk
      // per block (BN = 32) they alternate blocks, so two epilogues are in flight per quarter.
      const int c_first = (hh + li * kChunks) & 1;     // this warp's first chunk of the block
      const int my_last = (kChunks - 1 - c_first >= 0) ? c_first + ((kChunks - 1 - c_first) / 2) * 2 : -1;
      if (my_last < 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      for (int c32 = c_first; c32 < kChunks; c32 += 2) {
        {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c32 * 32), v);
          tmem_ld_wait(v);
#pragma unroll
          for (int cc = 0; cc < 8; ++cc)
            *reinterpret_cast<float4*>(stg + lane * kEpiLd + cc * 4) =
                make_float4(__uint_as_float(v[cc * 4]), __uint_as_float(v[cc * 4 + 1]),
                            __uint_as_float(v[cc * 4 + 2]), __uint_as_float(v[cc * 4 + 3]));
        }
        if (c32 == my_last) {            // accumulator drained by this warp: the issuer may reuse it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        __syncwarp();
        const int nbase = c.n0 + c32 * 32;
        const int mbase = c.m0 + q * 32;
        if (vec_c && nbase + 32 <= p.N) {
          const int l8 = lane & 7, r4 = lane >> 3;
          const int n = nbase + l8 * 4;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c.first && p.bias) b4 = *reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll 2
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + r4;
            const int m = mbase + rr;
            if (m >= p.M) continue;
            float4 x = *reinterpret_cast<const float4*>(stg + rr * kEpiLd + l8 * 4);
            x.x *= p.alpha; x.y *= p.alpha; x.z *= p.alpha; x.w *= p.alpha;
            if (c.first) {
              x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
              if (p.rowbias) {
                const float4 rb = *reinterpret_cast<const float4*>(p.rowbias + (int64_t)(m / rpg) * p.ldrb + n);
                x.x += rb.x; x.y += rb.y; x.z += rb.z; x.w += rb.w;
              }
            }
            if (p.gateC) {
              const float4 g = *reinterpret_cast<const float4*>(p.gateC + (int64_t)m * p.ldgc + n);
              if (!(g.x > 0.f)) x.x = 0.f;
              if (!(g.y > 0.f)) x.y = 0.f;
              if (!(g.z > 0.f)) x.z = 0.f;
              if (!(g.w > 0.f)) x.w = 0.f;
            }
            if (c.first && p.res) {
              const float4 rs = *reinterpret_cast<const float4*>(p.res + (int64_t)m * p.ldres + n);
              x.x += rs.x; x.y += rs.y; x.z += rs.z; x.w += rs.w;
            }
            if (p.relu_out) {
              x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
            }
            float* cp = p.C + (int64_t)m * p.ldc + n;
            if (p.atomic) {
              atomicAdd(cp, x.x); atomicAdd(cp + 1, x.y); atomicAdd(cp + 2, x.z); atomicAdd(cp + 3, x.w);
            } else {
              *reinterpret_cast<float4*>(cp) = x;
            }
          }
        } else {
          const int n = nbase + lane;
          const bool nok = n < p.N;
          const float bias_n = (c.first && p.bias && nok) ? p.bias[n] : 0.f;
          const int mrows = min(32, p.M - mbase);
          for (int rr = 0; rr < mrows; ++rr) {
            const int m = mbase + rr;
            float x = stg[rr * kEpiLd + lane] * p.alpha;
            if (nok) {
              if (c.first) {
                x += bias_n;
                if (p.rowbias) x += p.rowbias[(int64_t)(m / rpg) * p.ldrb + n];
              }
              if (p.gateC) x = (p.gateC[(int64_t)m * p.ldgc + n] > 0.f) ? x : 0.f;
              if (c.first && p.res) x += p.res[(int64_t)m * p.ldres + n];
              if (p.relu_out) x = fmaxf(x, 0.f);
              float* cp = p.C + (int64_t)m * p.ldc + n;
              if (p.atomic) atomicAdd(cp, x); else *cp = x;
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) tmem_dealloc(tmem_base, NBUF * BN);
}

// IBC_TC_DBG (diagnostics only, results are garbage): bit 0 = loaders skip the copies and
// the prologue (barrier traffic only), bit 1 = the issuer skips the MMAs.  Separates the
// loaders' rate from the issuer's (scripts/dbg_gemm_tc.py): with both off a K-slice still
// costs ~570 cycles of mbarrier hand-overs between the 16 loader warps and the issuer,
// which is what bounds blocks with few K-slices today.
int debug_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("IBC_TC_DBG");
    mode = e ? atoi(e) : 0;
  }
  return mode;
}

template <bool TA, bool TB, int BN, bool CONV>
int launch_tc(ibc_handle* h, const GemmP& p, int tm, int tn, int splits, cudaStream_t s) {
  using C = TcCfg<BN>;
  IBC_RAISE_SMEM(h, (gemm_tc_kernel<TA, TB, BN, CONV>), C::kSmem);
  const int64_t total = (int64_t)tm * tn * splits;
  if (total > 0x7fffffff) IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "gemm_tc: too many output blocks");
  const int sms = h ? h->sm_count : 148;
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  gemm_tc_kernel<TA, TB, BN, CONV><<<grid, kThreads, C::kSmem, s>>>(
      p, h ? h->dev_status : nullptr, tm, tn, (int)total, debug_mode());
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

template <bool TA, bool TB, bool CONV>
int launch_tc_bn(ibc_handle* h, const GemmP& p, int bn, int tm, int tn, int splits, cudaStream_t s) {
  if (bn == 32) return launch_tc<TA, TB, 32, CONV>(h, p, tm, tn, splits, s);
  if (bn == 64) return launch_tc<TA, TB, 64, CONV>(h, p, tm, tn, splits, s);
  return launch_tc<TA, TB, 128, CONV>(h, p, tm, tn, splits, s);
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
  int bn = p.N <= 32 ? 32 : (p.N <= 64 ? 64 : 128);
  if (!p.atomic && p.conv_C == 0) {
    // Few-row products (sampler chains: 1024-2048 rows) leave most SMs idle with 128-column
    // blocks and a block's K loop is latency-bound, so narrower blocks on more SMs finish
    // sooner: halve the block width while fewer than half of the SMs would have a block.
    static const bool narrow = getenv("IBC_TC_WIDE_BLOCKS") == nullptr;
    const int sms = h ? h->sm_count : 148;
    const int64_t rows_m = (p.M + kBM - 1) / kBM;
    while (narrow && bn > 32 && rows_m * ((p.N + bn - 1) / bn) * 2 < sms) bn >>= 1;
  }
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
  if (p.conv_C > 0) {       // implicit convolution: the weights are always [K, N] row-major
    if (tb) IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "gemm_tc: implicit convolution takes B as [K, N]");
    const int kdim = ta ? p.M : p.K;
    if (kdim != 9 * p.conv_C) IBC_FAIL(h, IBC_ERR_INVALID, "gemm_tc: convolution needs 9 * C columns");
    if (ta) return launch_tc_bn<true, false, true>(h, p, bn, tm, tn, splits, s);
    return launch_tc_bn<false, false, true>(h, p, bn, tm, tn, splits, s);
  }
  if (ta && tb) return launch_tc_bn<true, true, false>(h, p, bn, tm, tn, splits, s);
  if (ta) return launch_tc_bn<true, false, false>(h, p, bn, tm, tn, splits, s);
  if (tb) return launch_tc_bn<false, true, false>(h, p, bn, tm, tn, splits, s);
  return launch_tc_bn<false, false, false>(h, p, bn, tm, tn, splits, s);
}

}  // namespace ibc
