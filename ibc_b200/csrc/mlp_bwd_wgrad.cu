// Width-128 MLPEBM training backward as ONE kernel: reverse chain + weight
// gradients, block-major (tape.gradient of ibc/agents/base_agent.py:43-48 through
// networks/layers/resnet.py:135-153).
//
// Why.  Round 1 ran the reverse chain tile-major (mlp_bwd_tmem_kernel) and wrote
// every layer's output gradient dY to HBM (1.1 GB per C2 step) for a separate
// weight-gradient kernel that read it back together with the saved operands
// (2.2 GB): both kernels sat on the HBM roofline with the tensor pipe 17-26 %
// active.  dW_i = A_i^T dY_i is a sum over ALL rows, its accumulator is 128
// tensor-memory columns, and a tile-major chain visits all 2nb layers of a tile
// before the next tile -- 2nb accumulators do not fit 512 columns, so dY had to
// leave the SM.
//
// Here the loop nest is turned around: a persistent CTA owns a fixed set of
// 128-row tiles (tile = blockIdx.x + k * gridDim.x) and walks the ResNet blocks
// top-down; inside block j it visits all of its tiles.  Only TWO accumulators
// (dW2_j, dW1_j) are live, they stay in tensor memory across all the CTA's tiles
// and are flushed once per block, both of the block's transposed weights stay
// resident in shared memory (loaded once per block instead of once per tile), and
// dY never leaves the SM: the epilogue thread that produces a row of dY writes it
// into the MN-major operand image the weight-gradient MMA reads.  What is parked
// between blocks is one fp32 tile per (tile, block): the gradient of the residual
// stream, 64 KB, written and re-read by the same thread with L2 evict-last hints
// (67 MB for the C2 step: L2-resident); the saved forward operands stream through
// with evict-first.
//
// The weight-gradient GEMM runs in kind::f16 (fp16 operands, fp32 accumulate).  Shared
// memory forces it -- 128 KB of resident weights leave 96 KB, and in tf32 one tile's dY
// image (64 KB) plus landing zones for the saved operands do not fit: with half a
// tile of dY the fill -> MMA -> refill hops serialised four times per (tile, block), 17 k
// cycles against a 4 k tensor floor (timeline of the tf32 version) -- and
// the arithmetic allows it: fp16 carries tf32's 10-bit mantissa.  The saved operands are
// relu(.) rounded to tf32, i.e. exact fp16 values (the forward writes them as fp16);
// dY is scaled by a power of two so that max |dE| * max |we| lands at 2^3, which leaves 2^12
// of head room for growth down the chain and 2^-17 of the largest element before fp16
// runs out of normal numbers; conversions saturate.  Bias gradients are fp32 column sums
// of the unrounded dY (warp transposing reduction), as before.
//
// Per (tile, block j), g = dE/dh_{j+1} (fp32, registers of the row's threads):
//   G   X <- tf32(g); image(g)                       -> chain MMA  Y = X . W2_j^T
//                                                       wgrad MMA  dW2_j += A_{2j+1}^T . g
//   U   u = Y (.) m_{2j+1}; Y <- tf32(u); image(u)   -> chain MMA  X = Y . W1_j^T
//                                                       wgrad MMA  dW1_j += A_{2j}^T . u
//   G'  g <- g + X (.) m_{2j}, parked (after block 0 this is g0)
// Warp roles (24 warps): 0-15 epilogue (warp e: rows of sub-tile e & 3, columns
// 32 (e >> 2) ..: four warps per scheduler -- a lone warp per scheduler ran these phases
// at 3-5 cycles per instruction), 16-19 operand warps (saved A sub-tiles: 16-byte
// cp.async straight into the MN-major image, no registers: register prefetch shared a
// scoreboard with the loop's other variable-latency instructions and never overlapped),
// 20 weight producer, 21 chain MMA issuer, 22 weight-gradient MMA issuer (an mbarrier
// wait issued behind a thread's own tcgen05.mma does not return until they drain, so the
// two MMA streams have their own threads and the weight-gradient MMAs go eight per wait),
// 23 idle.
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "mlp.cuh"
#include "umma.cuh"
#include "umma_f16.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kW = 128;
constexpr int kLayerWBytes = kW * kW * 2;          // one transposed layer (fp16, K-major chunks): 32 KB
constexpr int kImgBytes = 32 * kW * 2;             // fp16 image of a 32-row sub-tile: 8 KB
constexpr int kKSteps = kW / 16;                   // kind::f16: K = 16 per MMA
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 8) * 32;
// 768 threads start with 80 registers; the CTA's pool is redistributed by setmaxnreg
constexpr int kEpiRegs = 96, kOpRegs = 32, kServiceRegs = 64;
constexpr int kASlots = 1;                         // saved-tile landing zones (32 KB each): see below
constexpr int kAStages = 4 * kASlots;              // in 8 KB sub-tile images
constexpr int kDbgWords = 2048;

constexpr size_t kOffW = 0;                                        // [2][32 KB]  W2_j^T, W1_j^T (fp16)
constexpr size_t kOffDy = kOffW + 2 * (size_t)kLayerWBytes;        // [2][4][8 KB] dY images (g | u), one per sub-tile
constexpr size_t kOffA = kOffDy + 8 * (size_t)kImgBytes;           // saved-tile landing zones
constexpr size_t kOffWe = kOffA + kAStages * (size_t)kImgBytes;    // we[128], scale, 1 / scale
constexpr size_t kOffBars = kOffWe + (kW + 4) * 4;
// a_ready, d_ready, w_full[2], w_empty[2], dy_full, dy_empty, a_full[2], a_empty[2],
// acc_done, acc_free.  (A passed mbarrier wait still costs 150-300 cycles; the weight-gradient
// issuer waited on 8-9 per batch of eight MMAs before the four sub-tiles of a slot shared one.)
constexpr int kNumBars = 2 + 4 + 4 + 4 + 2;
constexpr size_t kSmemBytes = kOffBars + kNumBars * 8 + 16;
static_assert(kSmemBytes <= 227 * 1024 - 1024, "shared memory budget");

struct BwdWgradParams {
  const float* dE;            // [R]
  const uint32_t* dE_absmax;  // bits of max |dE| (written by the loss kernel)
  const uint4* xa16;          // forward operand tiles, fp16: [tile][2nb][16 chunks of 8][128 rows][16 B]
  const uint32_t* mask_save;  // ReLU masks of the 2nb layers
  float* gpark;               // per tile 128 x 128 floats: the residual gradient between blocks (scaled fp16,
                              // [16 chunks of 8][128 rows][16 B] in the first half of the region);
                              // after the kernel: the gradient of the first Dense output (g0)
  const float* hn;            // forward's final residual stream, [tile][32 chunks][128 rows][4] (for dwe)
  float* acc_w;               // [2nb][32 chunks][128 in][4]: weight gradients, accumulated with coalesced
                              // vector reductions (permuted into the flat gradient afterwards)
  float* grads;               // flat gradient: b1 / b2 / we entries accumulated here directly
  Layout L;
  const float* packed;
  const __half* packed16;     // fp16 operand images (Packed16Layout): the transposed layers
  int64_t off_rev16, off_we;
  int64_t R;
  int nb, tiles;
  int* status;
  long long* dbg;
  int dbg_mode;   // timing experiments (wrong results): 1 no park loads, 2 no park stores, 4 no column
                  // sums, 8 no dY images, 256 no copies of the saved operands
};

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u32_hint(const uint32_t* ptr, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void red_add_v4(float* ptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void stg_hint(float4* ptr, const float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_u4_hint(const uint4* ptr, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_u4_hint(uint4* ptr, const uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ float2 half2_bits_to_float2(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
// 16-byte asynchronous copy global -> shared (LDGSTS): no destination registers, so the
// copies of several sub-tiles stay in flight without a scoreboard holding the warp
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst),
               "l"(gsrc), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One row's 32 consecutive columns as 16 packed fp16 pairs (round to nearest, saturating).
__device__ __forceinline__ void pack32_scaled(const uint32_t (&v)[32], float scale, uint32_t (&w)[16]) {
#pragma unroll
  for (int q = 0; q < 16; ++q)
    w[q] = pack_half2(__uint_as_float(v[2 * q]) * scale, __uint_as_float(v[2 * q + 1]) * scale);
}

// ... into the image of the row's sub-tile: four 16-byte pieces of the row's line
// (umma_f16.cuh).  The 128-byte swizzle spreads the eight rows of a quarter-warp over the eight
// 16-byte bank groups: conflict-free.
__device__ __forceinline__ void sts_image_f16(unsigned char* img, int lane, int cc, const uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(img + ImgF16::piece(lane, cc * 4 + i, 32)) =
        make_uint4(w[4 * i + 0], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// Transposing warp reduction: v[i] of every lane summed over the 32 lanes, the sum of
// v[l] ends in lane l (31 shuffles instead of 32 x 5).  Destroys v.
__device__ __forceinline__ float warp_colsum32(uint32_t (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = __uint_as_float(up ? v[i + s] : v[i]);
      const float send = __uint_as_float(up ? v[i] : v[i + s]);
      v[i] = __float_as_uint(keep + __shfl_xor_sync(0xffffffffu, send, s));
    }
  }
  return __uint_as_float(v[0]);
}

__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_wgrad_tmem_kernel(BwdWgradParams p) {
  using TT = TrainTile<kW>;
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* wbuf = smem + kOffW;
  unsigned char* dyb = smem + kOffDy;
  unsigned char* amn = smem + kOffA;
  float* we = reinterpret_cast<float*>(smem + kOffWe);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* a_ready = bars + 0;       // epilogue threads: chain A operand written
  uint64_t* d_ready = bars + 1;       // commit: chain accumulator complete
  uint64_t* w_full = bars + 2;        // [2] weights of this block landed
  uint64_t* w_empty = bars + 4;       // [2] last chain MMA of the block has read them
  uint64_t* dy_full = bars + 6;       // [2] the tile's dY image (g | u) written (16 epilogue warps)
  uint64_t* dy_empty = bars + 8;      // [2] its MMAs done
  uint64_t* a_full = bars + 10;       // [2] a slot's four A images landed (4 x 128 cp.async arrivals)
  uint64_t* a_empty = bars + 12;      // [2] its MMAs done
  uint64_t* acc_done = bars + 14;     // commit: the block's last weight-gradient MMA
  uint64_t* acc_free = bars + 15;     // epilogue threads: accumulators flushed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsl = 2 * p.nb;
  const int ntile = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // Boustrophedon: odd blocks visit the CTA's tiles in reverse, so the last tile of a block is
  // the first of the next and its residual gradient stays in registers (no park round trip).
  auto tile_of = [&](int blk, int i) -> int64_t {
    const int k = (blk & 1) ? ntile - 1 - i : i;
    return (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
  };

  if (tid == 0) {
    s_abort = 0;
    mbar_init(a_ready, kEpiWarps);        // one arrival per epilogue warp
    mbar_init(d_ready, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&dy_full[i], kEpiWarps); mbar_init(&dy_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(acc_done, 1);
    mbar_init(acc_free, kEpiWarps * 32);
    fence_mbar_init();
  }
  for (int i = tid; i < W; i += kThreads) we[i] = p.packed[p.off_we + i];
  if (warp == 1) {
    // dY scale: max |dE| * max |we| -> [8, 16)
    float m = 0.f;
    for (int i = lane; i < W; i += 32) m = fmaxf(m, fabsf(p.packed[p.off_we + i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
      m *= __uint_as_float(*p.dE_absmax);
      int e = (int)((__float_as_uint(m) >> 23) & 255u) - 127;       // floor(log2 m)
      e = max(-60, min(60, e));
      we[W] = __uint_as_float((uint32_t)(127 + 3 - e) << 23);
      we[W + 1] = __uint_as_float((uint32_t)(127 - 3 + e) << 23);
    }
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  // tensor-memory columns: X [0,128) | Y [128,256) | dW2 [256,384) | dW1 [384,512)

  if (warp < kEpiWarps) {
    // ============ epilogue: thread = (row, 32-column chunk) of the tile ============
    setmaxnreg_inc<kEpiRegs>();
    const int wq = warp & 3, cc = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t col_x = tmem_base + lane_off + (uint32_t)(cc * 32), col_y = col_x + W;
    unsigned char* dy_img = dyb + (size_t)wq * kImgBytes;       // g image; u image 4 sub-tiles further
    uint32_t item = 0;                                          // items done: the images' barrier phase
    const float scale = we[W], inv_scale = we[W + 1];
    const uint64_t pol_keep = l2_policy_evict_last();
    const uint64_t pol_once = l2_policy_evict_first();
    uint32_t ph_d = 0;
    float gr[32];                      // gradient of the residual stream: this row, this chunk
    float bsum_g = 0.f, bsum_u = 0.f;  // lane l: column 32 cc + l of db2_j / db1_j over this CTA's rows
    float bsum_e = 0.f;                // the same for dwe (top block only)
    for (int blk = 0; blk < p.nb; ++blk) {
      const int jb = p.nb - 1 - blk;
      for (int k = 0; k < ntile; ++k) {
        const int64_t tile = tile_of(blk, k);
        const bool carry_out = blk < p.nb - 1 && k == ntile - 1;     // ... and stays there
        const int64_t r = tile * 128 + row;
        const bool stamp = p.dbg && blockIdx.x == 0 && tid == 0 && (blk * ntile + k) < 24;
        long long* dbg = stamp ? p.dbg + (blk * ntile + k) * 8 : nullptr;
        if (stamp) dbg[0] = clock64();
        // read once: evict-first, so that they do not displace the parked tiles in L2
        const uint32_t m1 = ldg_u32_hint(p.mask_save + TT::mask_index(tile, p.nb, 2 * jb + 1, row) + cc, pol_once);
        const uint32_t m0 = ldg_u32_hint(p.mask_save + TT::mask_index(tile, p.nb, 2 * jb, row) + cc, pol_once);
        float4* park = reinterpret_cast<float4*>(p.gpark + tile * (int64_t)TT::kTileFloats) +
                       cc * 8 * 128 + row;
        uint32_t v[32];
        // ---- G: g into X (tf32), then its image and column sums -----------------
        float de = 0.f;
        if (blk == 0) {
          de = (r < p.R) ? p.dE[r] : 0.f;                  // g_top = dE (x) we
#pragma unroll
          for (int i = 0; i < 32; ++i) gr[i] = de * we[cc * 32 + i];
        }
        // blk > 0: g is already in registers -- carried over the block boundary (k == 0) or
        // loaded from the park at the end of the previous item (below)
        if (stamp) dbg[1] = clock64() + (long long)(gr[0] == 12345.f);   // after the loads landed
        // the chain's operand and the weight-gradient image are the same scaled fp16 values
        uint32_t w16[16];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(gr[q]);
        pack32_scaled(v, scale, w16);
        tmem_st16(col_x, w16);               // packed pairs in the first 16 columns of this chunk
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
        if (stamp) dbg[2] = clock64();
        mbar_wait(&dy_empty[0], (item & 1u) ^ 1u, abort_flag, p.status, 1000);   // previous g image consumed
        if (!(p.dbg_mode & 8)) sts_image_f16(dy_img, lane, cc, w16);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dy_full[0]);
        if (!(p.dbg_mode & 4)) bsum_g += warp_colsum32(v, lane);
        if (blk == 0) {
          // Dense(1) weight gradient: dwe += dE (x) h_nb, this row's chunk of the forward's last slot
          const float4* hrow = reinterpret_cast<const float4*>(p.hn) + tile * (int64_t)(32 * 128) +
                               cc * 8 * 128 + row;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 x = ldg_hint(hrow + q * 128, pol_once);
            v[q * 4 + 0] = __float_as_uint(de * x.x); v[q * 4 + 1] = __float_as_uint(de * x.y);
            v[q * 4 + 2] = __float_as_uint(de * x.z); v[q * 4 + 3] = __float_as_uint(de * x.w);
          }
          bsum_e += warp_colsum32(v, lane);
        }

        // ---- U: u = Y (.) m1 in place; then its image and column sums -----------
        mbar_wait(d_ready, ph_d, abort_flag, p.status, 1001);
        ph_d ^= 1;
        tc_fence_after();
        if (stamp) dbg[3] = clock64();
        tmem_ld32(col_y, v);
        tmem_ld_wait(v);
        // Y = (scale g) . W2^T: u stays scaled -- operand, image and (until the flush) bias sums
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = ((m1 >> q) & 1u) ? v[q] : 0u;
        pack32_scaled(v, 1.0f, w16);
        tmem_st16(col_y, w16);               // in place: the first 16 columns of this thread's own chunk
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
        if (stamp) dbg[4] = clock64();
        mbar_wait(&dy_empty[1], (item & 1u) ^ 1u, abort_flag, p.status, 1002);   // previous u image consumed
        if (!(p.dbg_mode & 8)) sts_image_f16(dy_img + 4 * kImgBytes, lane, cc, w16);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dy_full[1]);
        ++item;
        if (stamp) dbg[5] = clock64();
        if (!(p.dbg_mode & 4)) bsum_u += warp_colsum32(v, lane);

        // The next item's g (another tile of this CTA, parked by these same threads one block
        // ago) is requested here, while the second chain MMA runs: v is free until the end of
        // the item (G' reads its accumulator through the 16 registers of w16), and the loads
        // are ahead of this item's 64 KB of park stores, which drain at the SM's ~32 B/clk.
        const bool load_next = blk > 0 && k + 1 < ntile && !(p.dbg_mode & 1);
        if (load_next) {
          // (parked between blocks as scaled fp16: [16 chunks of 8 columns][128 rows][16 B] in
          // the first half of the tile's region; v holds the packed pairs until G' unpacks them)
          const uint4* pn = reinterpret_cast<const uint4*>(p.gpark + tile_of(blk, k + 1) * (int64_t)TT::kTileFloats) +
                            cc * 4 * 128 + row;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 x = ldg_u4_hint(pn + q * 128, pol_keep);
            v[q * 4 + 0] = x.x; v[q * 4 + 1] = x.y; v[q * 4 + 2] = x.z; v[q * 4 + 3] = x.w;
          }
        }

        // ---- G': g += X (.) m0, parked (after the last block this is g0) -----------
        mbar_wait(d_ready, ph_d, abort_flag, p.status, 1003);
        ph_d ^= 1;
        tc_fence_after();
        if (stamp) dbg[6] = clock64();
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          tmem_ld16(col_x + (uint32_t)(hq * 16), w16);
          tmem_ld_wait16(w16);
#pragma unroll
          for (int q = 0; q < 16; ++q)
            gr[hq * 16 + q] += ((m0 >> (hq * 16 + q)) & 1u) ? __uint_as_float(w16[q]) * inv_scale : 0.f;
        }
        if (!carry_out && !(p.dbg_mode & 2)) {
          if (blk == p.nb - 1) {      // g0, the kernel's output: fp32
#pragma unroll
            for (int q = 0; q < 8; ++q)
              stg_hint(park + q * 128, make_float4(gr[q * 4 + 0], gr[q * 4 + 1], gr[q * 4 + 2], gr[q * 4 + 3]),
                       pol_keep);
          } else {
            // Between blocks the residual gradient is parked as fp16 of scale * g -- the 11-bit
            // significand the next block's chain operand and image round it to anyway; what the
            // parking adds is one such rounding per block on the running sum (per-tensor gradient
            // errors of this stack against fp32 are 1e-2 .. 8e-2, from mask flips).  Half the
            // bytes through the SM's store port, and the 1 028 parked tiles are 34 MB instead of
            // 67: they stay in L2.
            uint4* pk = reinterpret_cast<uint4*>(p.gpark + tile * (int64_t)TT::kTileFloats) + cc * 4 * 128 + row;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack_half2(gr[q * 8 + 0] * scale, gr[q * 8 + 1] * scale);
              w.y = pack_half2(gr[q * 8 + 2] * scale, gr[q * 8 + 3] * scale);
              w.z = pack_half2(gr[q * 8 + 4] * scale, gr[q * 8 + 5] * scale);
              w.w = pack_half2(gr[q * 8 + 6] * scale, gr[q * 8 + 7] * scale);
              stg_u4_hint(pk + q * 128, w, pol_keep);
            }
          }
        }
        if (load_next) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float2 f = half2_bits_to_float2(v[q]);
            gr[2 * q] = f.x * inv_scale;
            gr[2 * q + 1] = f.y * inv_scale;
          }
        }
        tc_fence_before();
        if (stamp) dbg[7] = clock64();
      }
      // ---- flush dW2_j, dW1_j (unscaled) and the bias sums of this CTA's rows -------
      mbar_wait(acc_done, (uint32_t)(blk & 1), abort_flag, p.status, 1005);
      tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
        const int layer = 2 * jb + 1 - a;        // a = 0: dW2_j (slot 2j+1), a = 1: dW1_j (slot 2j)
        float* pw = p.acc_w + (int64_t)layer * (W * W) + ((cc * 8 * 128) + row) * 4;
        uint32_t v[32];
        tmem_ld32(col_x + (uint32_t)((2 + a) * W), v);
        tmem_ld_wait(v);
        // 512 contiguous bytes per warp instruction; 148 CTAs reduce into the same 64 KB
#pragma unroll
        for (int q = 0; q < 8; ++q)
          red_add_v4(pw + q * 128 * 4, __uint_as_float(v[q * 4 + 0]) * inv_scale, __uint_as_float(v[q * 4 + 1]) * inv_scale,
                     __uint_as_float(v[q * 4 + 2]) * inv_scale, __uint_as_float(v[q * 4 + 3]) * inv_scale);
        atomicAdd(p.grads + (a ? p.L.b1(jb) : p.L.b2(jb)) + cc * 32 + lane, a ? bsum_u * inv_scale : bsum_g);
      }
      if (blk == 0) atomicAdd(p.grads + p.L.we + cc * 32 + lane, bsum_e);
      bsum_g = 0.f; bsum_u = 0.f;
      tc_fence_before();
      mbar_arrive(acc_free);
    }
  } else if (warp < kEpiWarps + 4) {
    // ===================== operand warps ===========================================
    // The forward's saved fp16 tile of a (tile, layer) -- [chunk of 8 features][128 rows][16 B],
    // 32 KB contiguous -- IS a valid MN-major operand image without swizzle: a core matrix is 8
    // rows x 16 B, row groups 128 B apart (LBO), feature chunks 2 KB apart (SBO); validated on
    // hardware (ibc_umma_f16_selftest mode 200).  So one bulk copy per (tile, layer) brings the A
    // operand of its eight weight-gradient MMAs, on the TMA engine: the 16-byte cp.async copies
    // that used to re-arrange it into a swizzled image went through the LSU, 128 warp
    // instructions per item, and cost the kernel 0.066 ms (IBC_DBG_MODE=256).
    setmaxnreg_dec<kOpRegs>();
    const int ow = warp - kEpiWarps;
    const uint64_t pol_stream = l2_policy_evict_first();     // read once: do not displace the parked tiles
    const int nitems = p.nb * ntile;
    // L2 prefetch of a whole (tile, block)'s saved operands (two adjacent slots x 32 KB) by
    // the TMA engine, two (tile, block)s ahead: the copies below then hit L2 (~2 us from
    // HBM under load against the few hundred cycles the landing zones can cover)
    auto prefetch_item = [&](int T) {
      if (T >= nitems) return;
      const int blk = T / ntile, k = T - blk * ntile, jb = p.nb - 1 - blk;
      const int64_t tile = tile_of(blk, k);
      const uint4* src = p.xa16 + (tile * nsl + 2 * jb) * (int64_t)(16 * 128);
#pragma unroll 1
      for (int i = 0; i < 4; ++i)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (int64_t)i * 1024),
                     "r"(16384u)
                     : "memory");
      if (blk > 0 && k > 0) {     // the parked residual gradient of the same tile (L2-resident unless evicted)
        const float* pk = p.gpark + tile * (int64_t)TT::kTileFloats;
#pragma unroll 1
        for (int i = 0; i < 2; ++i)      // fp16: the first 32 KB of the tile's region
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pk + (int64_t)i * 4096),
                       "r"(16384u)
                       : "memory");
      }
    };
    if (ow == 0) {
      if (lane == 0) { prefetch_item(0); prefetch_item(1); }
      for (int u = 0; u < 2 * nitems; ++u) {       // unit = (item, half): slot 2jb+1 (dW2), then 2jb (dW1)
        const int T = u >> 1, half = u & 1;
        const int blk = T / ntile, k = T - blk * ntile, jb = p.nb - 1 - blk;
        const int64_t tile = tile_of(blk, k);
        const uint4* src = p.xa16 + (tile * nsl + (half ? 2 * jb : 2 * jb + 1)) * (int64_t)(16 * 128);
        const int st = u % kASlots;
        if (elect_one()) {
          if (half == 0) prefetch_item(T + 2);
          mbar_wait(&a_empty[st], ((uint32_t)(u / kASlots) & 1u) ^ 1u, abort_flag, p.status, 1100);
          mbar_arrive_expect_tx(&a_full[st], 4u * kImgBytes);
          if (p.dbg_mode & 256) {
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&a_full[st])), "r"(4u * kImgBytes) : "memory");
          } else {
            bulk_g2s_hint(amn + (size_t)st * 4 * kImgBytes, src, 4u * kImgBytes, &a_full[st], pol_stream);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == kEpiWarps + 4) {
    // ===================== producer: the block's two transposed weights =============
    setmaxnreg_dec<kServiceRegs>();
    for (int blk = 0; blk < p.nb; ++blk) {
      for (int half = 0; half < 2; ++half) {
        if (elect_one()) {
          mbar_wait(&w_empty[half], ((uint32_t)blk & 1u) ^ 1u, abort_flag, p.status, 1200 + half);
          mbar_arrive_expect_tx(&w_full[half], (uint32_t)kLayerWBytes);
          bulk_g2s(wbuf + (size_t)half * kLayerWBytes,
                   p.packed16 + p.off_rev16 + (int64_t)(2 * blk + half) * W * W,
                   (uint32_t)kLayerWBytes, &w_full[half]);
        }
        __syncwarp();
      }
    }
  } else if (warp == kEpiWarps + 5) {
    // ===================== chain MMA issuer (kind::f16) =================================
    // The chain's operands are the scaled fp16 values of the weight-gradient images (tf32's
    // significand; the range is the images'), packed two K-elements per tensor-memory column in
    // the first 16 columns of each 32-column chunk: K-step ks reads columns 32 (ks / 2) + 8 (ks % 2).
    setmaxnreg_dec<kServiceRegs>();
    const uint32_t idesc = idesc_f16(128, W);
    const uint32_t b_lbo = W * 16, sbo = 128;
    constexpr uint32_t b_step16 = (2 * W * 16) >> 4;
    uint32_t ph_a = 0;
    for (int blk = 0; blk < p.nb; ++blk) {
      for (int k = 0; k < ntile; ++k) {
        for (int half = 0; half < 2; ++half) {
          if (elect_one()) {
            if (k == 0) mbar_wait(&w_full[half], (uint32_t)blk & 1u, abort_flag, p.status, 1310 + half);
            mbar_wait(a_ready, ph_a, abort_flag, p.status, 1300 + half);
            tc_fence_after();
            // half 0: Y = X . W2^T ; half 1: X = Y . W1^T
            const uint32_t a_tmem = tmem_base + (uint32_t)(half ? W : 0);
            const uint32_t d_tmem = tmem_base + (uint32_t)(half ? 0 : W);
            const uint64_t bd = smem_desc(smem_u32(wbuf + (size_t)half * kLayerWBytes), b_lbo, sbo);
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks)
              mma_f16_ts(d_tmem, a_tmem + (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8),
                         bd + (uint64_t)(ks * b_step16), idesc, (uint32_t)(ks > 0));
            if (k == ntile - 1) mma_commit(&w_empty[half]);
            mma_commit(d_ready);
          }
          __syncwarp();
          ph_a ^= 1;
        }
      }
    }
  } else if (warp == kEpiWarps + 6) {
    // ===================== weight-gradient MMA issuer (kind::f16) ======================
    setmaxnreg_dec<kServiceRegs>();
    // both operands MN-major, 128-byte swizzle (umma_f16.cuh); K = 16 rows per MMA = two
    // 8-row atoms; a batch = the four sub-tiles of one slot = 8 MMAs
    const uint32_t idesc = idesc_f16(128, W) | kIdescAMajorMN | kIdescBMajorMN;
    const uint32_t lbo = ImgF16::lbo(32);
    uint32_t u = 0;
    for (int blk = 0; blk < p.nb; ++blk) {
      for (int k = 0; k < ntile; ++k) {
        for (int half = 0; half < 2; ++half, u += 4) {
          if (elect_one()) {
            const bool wstamp = p.dbg && blockIdx.x == 0 && u < 384;
            long long* wd = wstamp ? p.dbg + 1024 + (u >> 2) * 4 : nullptr;
            if (wstamp) wd[0] = clock64();
            if (blk > 0 && k == 0 && half == 0)
              mbar_wait(acc_free, (uint32_t)(blk - 1) & 1u, abort_flag, p.status, 1400);
            mbar_wait(&dy_full[half], (u >> 3) & 1u, abort_flag, p.status, 1401);
            if (wstamp) wd[1] = clock64();
            mbar_wait(&a_full[(u >> 2) % kASlots], ((u >> 2) / kASlots) & 1u, abort_flag, p.status, 1402);
            if (wstamp) wd[2] = clock64();
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)((2 + half) * W);
            // A: the saved tile as it is (MN-major, no swizzle: LBO 128 B between 8-row groups,
            // SBO 2 KB between feature chunks), 16 rows = 256 B per MMA; B: the four swizzled
            // 32-row dY images
            const uint64_t ad0 = smem_desc_lt(smem_u32(amn + (size_t)((u >> 2) % kASlots) * 4 * kImgBytes), 128u, 2048u, 0);
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
              const uint64_t bd0 = smem_desc_lt(smem_u32(dyb + (size_t)(half * 4 + sub) * kImgBytes), lbo, ImgF16::kSbo, 2);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                mma_f16_ss(acc, ad0 + (uint64_t)((sub * 2 + ks) * (256 >> 4)),
                           bd0 + (uint64_t)(ks * (2 * ImgF16::kSbo >> 4)), idesc,
                           (uint32_t)(k > 0 || sub > 0 || ks > 0));
            }
            mma_commit(&a_empty[(u >> 2) % kASlots]);
            mma_commit(&dy_empty[half]);
            if (k == ntile - 1 && half == 1) mma_commit(acc_done);
            if (wstamp) wd[3] = clock64();
          }
          __syncwarp();
        }
      }
    }
  } else {
    setmaxnreg_dec<kServiceRegs>();     // idle member of the service warpgroup
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// The accumulated weight gradients from the reduction layout [layer][chunk c][input row m][4]
// into the flat gradient's Keras layout [in][out].
__global__ void __launch_bounds__(256) wgrad_permute_kernel(const float* __restrict__ acc_w, Layout L,
                                                            float* __restrict__ grads) {
  constexpr int W = kW;
  const int layer = blockIdx.x >> 4;
  const int f = ((blockIdx.x & 15) << 8) + threadIdx.x;          // float4 index: chunk c, input row m
  const int c = f >> 7, m = f & 127;
  const int j = layer >> 1;
  const float4 v = reinterpret_cast<const float4*>(acc_w)[(int64_t)layer * (W * W / 4) + f];
  float* dw = grads + ((layer & 1) ? L.w2(j) : L.w1(j)) + (int64_t)m * W + c * 4;
  *reinterpret_cast<float4*>(dw) = v;
}

}  // namespace

size_t bwd_wgrad_scratch_floats(const ibc_handle* h, int64_t R, size_t* park, size_t* accw) {
  const Layout& L = h->L;
  const int64_t tiles = (R + 127) / 128;
  *park = (size_t)tiles * 128 * kW;
  *accw = (size_t)(2 * L.nb) * kW * kW;
  return *park + *accw;
}

// Reverse chain + weight / bias gradients of the 2nb W x W layers and of Dense(1)'s kernel.
// grads must be zeroed (or hold what is to be accumulated onto).  On return gpark holds g0, the
// gradient of the first Dense output, in the tile layout [tile][chunk (32)][128 rows][4].
int tmem_reverse_wgrad(ibc_handle* h, const float* dE, const uint32_t* dE_absmax, int64_t R,
                       const void* xa16, const float* hn, const uint32_t* mask_save, float* gpark,
                       float* acc_w, float* grads, cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  BwdWgradParams p;
  p.dE = dE; p.dE_absmax = dE_absmax; p.xa16 = reinterpret_cast<const uint4*>(xa16);
  p.mask_save = mask_save; p.gpark = gpark; p.hn = hn;
  p.acc_w = acc_w; p.grads = grads; p.L = L; p.packed = h->packed;
  p.packed16 = reinterpret_cast<const __half*>(h->packed16);
  p.off_rev16 = make_packed16_layout(L).rev; p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.R = R; p.nb = L.nb;
  const int64_t tiles = (R + 127) / 128;
  p.tiles = (int)tiles;
  const int grid = tiles < h->sm_count ? (int)tiles : h->sm_count;
  p.status = h->dev_status;
  p.dbg = nullptr;
  p.dbg_mode = getenv("IBC_DBG_MODE") ? atoi(getenv("IBC_DBG_MODE")) : 0;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&p.dbg, kDbgWords * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(p.dbg, 0, kDbgWords * sizeof(long long), s));
  }
  IBC_RAISE_SMEM(h, mlp_bwd_wgrad_tmem_kernel, kSmemBytes);
  IBC_CUDA(h, cudaMemsetAsync(acc_w, 0, (size_t)(2 * L.nb) * kW * kW * 4, s));
  mlp_bwd_wgrad_tmem_kernel<<<grid, kThreads, kSmemBytes, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  wgrad_permute_kernel<<<2 * L.nb * 16, 256, 0, s>>>(acc_w, L, grads);
  IBC_LAUNCH_CHECK(h);
  if (p.dbg) {
    cudaStreamSynchronize(s);
    std::vector<long long> hb(kDbgWords);
    cudaMemcpy(hb.data(), p.dbg, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const long long t0 = hb[0];
    fprintf(stderr, "[ibc timeline bwd_wgrad] item | G_start loaded G_done | U_start U_p1 U_img | G'_start G'_done\n");
    for (int i = 0; i < 24; ++i) {
      const long long* e = &hb[i * 8];
      if (!e[0]) break;
      fprintf(stderr, "[ibc timeline bwd_wgrad] %3d | %8lld %8lld %8lld | %8lld %8lld %8lld | %8lld %8lld\n", i,
              e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0, e[6] - t0, e[7] - t0);
    }
    fprintf(stderr, "[ibc timeline bwd_wgrad wgrad issuer] batch | start dy_full a_full issued\n");
    for (int u = 0; u < 48; ++u) {      // batches of four sub-tiles (one slot)
      const long long* e = &hb[1024 + u * 4];
      fprintf(stderr, "[ibc timeline bwd_wgrad wg] %3d | %8lld %8lld %8lld %8lld\n", u, e[0] - t0, e[1] - t0,
              e[2] - t0, e[3] - t0);
    }
  }
  return IBC_OK;
}

}  // namespace ibc

namespace ibc {
using namespace umma;

// ---------------------------------------------------------------------------
// Probe of the kind::f16 MN-major descriptor: D[128, N] = A[K, 128]^T . B[K, N] with
// both operands converted to fp16 and laid out by ImgF16; descriptor fields are
// arguments so that one run can try several.
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128, 1)
umma_f16_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                      uint32_t lbo_a, uint32_t lbo_b, uint32_t sbo, uint32_t layout_type,
                      uint32_t kstep_bytes, float* __restrict__ D, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_img = smem;
  unsigned char* b_img = smem + (size_t)K * 128 * 2;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { s_abort = 0; mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  // thread = K-row r (r < K): all its MN pieces
  for (int r = tid; r < K; r += 128) {
    for (int c = 0; c < 128 / 8; ++c) {
      uint4 v;
      const float* s = A + (int64_t)r * 128 + c * 8;
      v.x = pack_half2(s[0], s[1]); v.y = pack_half2(s[2], s[3]);
      v.z = pack_half2(s[4], s[5]); v.w = pack_half2(s[6], s[7]);
      if (layout_type >= 200) *reinterpret_cast<uint4*>(a_img + (size_t)c * K * 16 + (size_t)r * 16) = v;
      else *reinterpret_cast<uint4*>(a_img + ImgF16::piece(r, c, K)) = v;
    }
    for (int c = 0; c < N / 8; ++c) {
      uint4 v;
      const float* s = B + (int64_t)r * N + c * 8;
      v.x = pack_half2(s[0], s[1]); v.y = pack_half2(s[2], s[3]);
      v.z = pack_half2(s[4], s[5]); v.w = pack_half2(s[6], s[7]);
      *reinterpret_cast<uint4*>(b_img + ImgF16::piece(r, c, K)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_f16(128, N) | kIdescAMajorMN | kIdescBMajorMN;
    // layout_type 200 / 201: A as the plain saved-tile layout [chunk of 8 MN][K rows][16 B] =
    // MN-major without swizzle (core matrix = 8 K-rows x 16 B, contiguous); 200: LBO = 128
    // (K groups), SBO = 16 K (MN chunks); 201: the two swapped.  B stays the swizzled image.
    const bool plain_a = layout_type >= 200;
    const uint32_t a_lbo = plain_a ? (layout_type == 200 ? 128u : (uint32_t)K * 16u) : lbo_a;
    const uint32_t a_sbo = plain_a ? (layout_type == 200 ? (uint32_t)K * 16u : 128u) : sbo;
    const uint32_t b_lt = plain_a ? 2u : layout_type;
    const uint32_t b_lbo = plain_a ? ImgF16::lbo(K) : lbo_b, b_sbo = plain_a ? ImgF16::kSbo : sbo;
    const uint32_t b_step = plain_a ? 2 * ImgF16::kSbo : kstep_bytes;
    const uint32_t a_step = plain_a ? 256u : kstep_bytes;
    const uint64_t ad0 = smem_desc_lt(smem_u32(a_img), a_lbo, a_sbo, plain_a ? 0u : layout_type);
    const uint64_t bd0 = smem_desc_lt(smem_u32(b_img), b_lbo, b_sbo, b_lt);
    for (int ks = 0; ks < K / 16; ++ks)
      mma_f16_ss(tmem_base, ad0 + (uint64_t)(ks * (a_step >> 4)), bd0 + (uint64_t)(ks * (b_step >> 4)),
                 idesc, (uint32_t)(ks > 0));
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0, &s_abort, status, 77);
  tc_fence_after();
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  for (int c32 = 0; c32 < N / 32; ++c32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + lane_off + c32 * 32, v);
    tmem_ld_wait(v);
    for (int q = 0; q < 32; ++q) D[(int64_t)tid * N + c32 * 32 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// K-major variants of the probe (layout_type >= 100): D[m][n] = sum_k A[m][k] B[n][k] with
// A [128][K] and B [N][K] row-major fp32 inputs rounded to fp16.  B is the no-swizzle K-major
// image [K/8][N][8 halves] (the tf32 chain's byte geometry: 16-byte chunks, LBO = N * 16,
// SBO = 128, K = 16 per MMA = two chunks).  mode 100: A the same image in shared memory (SS);
// 101: A in tensor memory, two K-elements per 32-bit column, even k in the low half (TS);
// 102: the same with odd k in the low half.
__global__ void __launch_bounds__(128, 1)
umma_f16_kmajor_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                             int mode, float* __restrict__ D, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_img = smem;                                  // [K/8][128][16 B]
  unsigned char* b_img = smem + (size_t)K * 128 * 2;            // [K/8][N][16 B]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { s_abort = 0; mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  for (int c = 0; c < K / 8; ++c) {
    const float* sa = A + (int64_t)tid * K + c * 8;
    uint4 v;
    v.x = pack_half2(sa[0], sa[1]); v.y = pack_half2(sa[2], sa[3]);
    v.z = pack_half2(sa[4], sa[5]); v.w = pack_half2(sa[6], sa[7]);
    *reinterpret_cast<uint4*>(a_img + ((size_t)c * 128 + tid) * 16) = v;
    for (int n = tid; n < N; n += 128) {
      const float* sb = B + (int64_t)n * K + c * 8;
      uint4 w;
      w.x = pack_half2(sb[0], sb[1]); w.y = pack_half2(sb[2], sb[3]);
      w.z = pack_half2(sb[4], sb[5]); w.w = pack_half2(sb[6], sb[7]);
      *reinterpret_cast<uint4*>(b_img + ((size_t)c * N + n) * 16) = w;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t col_a = tmem_base + 128;                       // A operand behind the accumulator
  if (mode != 100) {
    for (int c = 0; c < K / 2; ++c) {
      const float lo = A[(int64_t)tid * K + 2 * c + (mode == 102 ? 1 : 0)];
      const float hi = A[(int64_t)tid * K + 2 * c + (mode == 102 ? 0 : 1)];
      uint32_t v[1] = {pack_half2(lo, hi)};
      tmem_st_n<1>(col_a + lane_off + c, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = idesc_f16(128, N);
    const uint64_t ad0 = smem_desc(smem_u32(a_img), 128 * 16, 128);
    const uint64_t bd0 = smem_desc(smem_u32(b_img), (uint32_t)N * 16, 128);
    for (int ks = 0; ks < K / 16; ++ks) {
      if (mode == 100)
        mma_f16_ss(tmem_base, ad0 + (uint64_t)(ks * ((2 * 128 * 16) >> 4)),
                   bd0 + (uint64_t)(ks * ((2 * N * 16) >> 4)), idesc, (uint32_t)(ks > 0));
      else
        mma_f16_ts(tmem_base, col_a + (uint32_t)(ks * 8), bd0 + (uint64_t)(ks * ((2 * N * 16) >> 4)),
                   idesc, (uint32_t)(ks > 0));
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0, &s_abort, status, 78);
  tc_fence_after();
  for (int c32 = 0; c32 < N / 32; ++c32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + lane_off + c32 * 32, v);
    tmem_ld_wait(v);
    for (int q = 0; q < 32; ++q) D[(int64_t)tid * N + c32 * 32 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}
}  // namespace

int umma_f16_selftest(const float* A, const float* B, int N, int K, uint32_t lbo_a, uint32_t lbo_b,
                      uint32_t sbo, uint32_t layout_type, uint32_t kstep_bytes, float* D,
                      int* status_host, cudaStream_t s) {
  ibc_handle* h = nullptr;
  if (N < 64 || N > 256 || (N % 64) || K < 16 || K > 128 || (K % 16))
    IBC_FAIL(h, IBC_ERR_INVALID, "umma_f16_selftest: need N in {64,128,192,256}, K a multiple of 16 <= 128");
  if (layout_type >= 100 && layout_type < 200) {      // K-major operands, A from shared or tensor memory (see the kernel)
    if (N > 128) IBC_FAIL(h, IBC_ERR_INVALID, "umma_f16_selftest: K-major modes need N <= 128");
    int* st = nullptr;
    IBC_CUDA(h, cudaMalloc((void**)&st, sizeof(int)));
    IBC_CUDA(h, cudaMemsetAsync(st, 0, sizeof(int), s));
    IBC_CUDA(h, cudaFuncSetAttribute(umma_f16_kmajor_probe_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    umma_f16_kmajor_probe_kernel<<<1, 128, (size_t)K * 128 * 2 + (size_t)K * N * 2, s>>>(
        A, B, N, K, (int)layout_type, D, st);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    int st_host = -1;
    if (e == cudaSuccess) e = cudaMemcpy(&st_host, st, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(st);
    if (status_host) *status_host = st_host;
    if (e != cudaSuccess) IBC_FAIL(h, IBC_ERR_CUDA, "umma_f16_selftest: %s", cudaGetErrorString(e));
    return IBC_OK;
  }
  if (!lbo_a) lbo_a = ImgF16::lbo(K);
  if (!lbo_b) lbo_b = ImgF16::lbo(K);
  if (!sbo) sbo = ImgF16::kSbo;
  if (!layout_type) layout_type = 2;          // SWIZZLE_128B (200 / 201: see the kernel)
  if (!kstep_bytes) kstep_bytes = 2 * ImgF16::kSbo;   // K = 16 per MMA = two 8-row atoms
  int* st = nullptr;
  IBC_CUDA(h, cudaMalloc((void**)&st, sizeof(int)));
  IBC_CUDA(h, cudaMemsetAsync(st, 0, sizeof(int), s));
  const size_t smem = (size_t)K * 128 * 2 + (size_t)K * N * 2;
  IBC_CUDA(h, cudaFuncSetAttribute(umma_f16_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   128 * 1024));
  umma_f16_probe_kernel<<<1, 128, smem, s>>>(A, B, N, K, lbo_a, lbo_b, sbo, layout_type, kstep_bytes, D, st);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  int st_host = -1;
  if (e == cudaSuccess) e = cudaMemcpy(&st_host, st, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(st);
  if (status_host) *status_host = st_host;
  if (e != cudaSuccess) IBC_FAIL(h, IBC_ERR_CUDA, "umma_f16_selftest: %s", cudaGetErrorString(e));
  return IBC_OK;
}

}  // namespace ibc
