// Width-128 MLPEBM forward, activations in tensor memory, TWO tiles per CTA with EIGHT epilogue
// warps per tile (networks/layers/resnet.py:135-153, networks/mlp_ebm.py:72-96).
//
// Same scheme as mlp_fwd_tmem_kernel (mlp_umma_tmem.cu: X / Y regions alternate as A operand and
// accumulator, rewritten in place; whole-layer weight ring; bias by one extra K = 8 MMA; one MMA
// issuer per tile and a tensor-pipe token), different epilogue split.  There a thread owns a whole
// row (128 columns) and a tile's four epilogue warps each sit alone on their scheduler; measured
// (scripts/timeline.py) the fp16-saving epilogue of one layer then takes ~4 500 cycles against
// ~1 500 for the layer's MMAs, so the tensor pipe idles 63 % of the time.  A one-tile CTA with
// sixteen warps (thread = row x 32 columns) does the same epilogue in ~1 800 cycles -- the integer
// pipe issues every other cycle per scheduler, ~100 of its operations per thread and layer -- but
// with nothing to overlap it with was slower still (4 170 cycles per tile and layer).  Here both:
// two tiles in flight, each with eight warps (thread = row x 64 columns, 64 registers of residual
// stream), four warps per scheduler, one tile's epilogue under the other's MMAs.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "mlp.cuh"
#include "mlp_tiling.cuh"
#include "umma.cuh"
#include "umma_f16.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kW = 128;
constexpr int kNT = 2;
constexpr int kLayerWBytes = kW * kW * 4;
constexpr int kBiasBytes = 8 * kW * 4;
constexpr int kLayerBytes = kLayerWBytes + kBiasBytes;   // 68 KB
constexpr int kKSteps = kW / 8;
constexpr int kRing = 2;     // two layers of prefetch keep the pipe fed; a third only takes L1 away from the spills
constexpr int kLayerHBytes = kW * kW * 2 + 16 * kW * 2;  // kF16: fp16 layer + K = 16 bias slice: 36 KB
constexpr int kEpiWarps = 8 * kNT;
constexpr int kThreads = (kEpiWarps + 4) * 32;           // + producer, two issuers, one idle warp
// 640 threads launch with 96 registers (61 440 in the CTA's pool); setmaxnreg then gives the
// epilogue warpgroups 112 and leaves the service warpgroup 32
constexpr int kEpiRegs = 112, kServiceRegs = 32;

struct FwdHalvesParams {
  const float* act;      // [R, A]
  const float* hp;       // [B, W]  obs . Wp[:O] + bp
  const float* packed;
  const __half* packed16;  // kF16: fp16 operand images (Packed16Layout: p0, forward layer blocks)
  int64_t off16_p0, off16_fwd;
  float* E;              // [R]
  uint4* xa16;           // kSave == 2: fp16 operands [tile][2nb slots][16 chunks of 8][128 rows][16 B]
  float* hn_save;        // kSave == 2: final residual stream [tile][32 chunks][128 rows][4] (fp32)
  uint32_t* mask_save;   // kSave == 2: ReLU masks of the 2nb W x W layers
  int64_t R, n_per_obs;
  int nb, A, KP;
  Tiling work;
  int64_t off_fwd_p, off_fwd_lb, off_we;
  int* status;
  long long* dbg;
};

// ring (2 x 68 KB) | step-0 operand (KP x W) | ones tile (4 KB) | we, be | partial energies | barriers
inline size_t halves_smem_bytes(int KP) {
  return (size_t)kRing * kLayerBytes + (size_t)KP * kW * 4 + 4096 + (kW + 4) * 4 + kNT * 128 * 4 +
         (size_t)(2 * kNT + 2 * kRing + 1) * 8 + 64;
}

__device__ __forceinline__ void pipe_acquire(int* lock, volatile int* abort_flag) {
  while (atomicCAS(lock, 0, 1) != 0) {
    if (*abort_flag) return;
    __nanosleep(32);
  }
}
__device__ __forceinline__ void pipe_release(int* lock) { atomicExch(lock, 0); }

// kSave: 0 = energies only; 2 = fp16 operand tiles + h_nb + masks (the block-major backward's
// inputs, mlp_bwd_wgrad.cu)
// kF16: the dependent GEMMs in kind::f16 (the training forward: its operands ARE the fp16 tiles
// it saves; same 11-bit significand as tf32, see langevin_f16.cu): half the MMA time per layer
// and fewer integer-pipe operations per element (pack once, relu on packed pairs).  The A operand
// is packed two K-elements per column in the first 16 columns of each 32-column chunk.
template <int kSave, bool kF16>
__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_halves_kernel(FwdHalvesParams p) {
  using TT = TrainTile<kW>;
  constexpr int W = kW;
  extern __shared__ __align__(128) unsigned char smem[];
  const int nsteps = 2 * p.nb + 1;
  constexpr int kSlot = kF16 ? kLayerHBytes : kLayerBytes;
  unsigned char* ring = smem;                                         // [kRing][68 KB] (fp16: 36 KB)
  unsigned char* p0_buf = ring + (size_t)kRing * kLayerBytes;         // step-0 operand [KP/4][W][4] (fp16: [2][W][8])
  float4* ones_tile = reinterpret_cast<float4*>(p0_buf + (size_t)p.KP * W * 4);   // [2][128]
  float* we = reinterpret_cast<float*>(ones_tile + 256);              // we[W], be
  float* e_part = we + W + 4;                                         // [kNT][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(e_part + kNT * 128);
  uint64_t* a_ready = bars;                    // [kNT]
  uint64_t* d_ready = bars + kNT;              // [kNT]
  uint64_t* w_full = bars + 2 * kNT;           // [kRing]
  uint64_t* w_empty = w_full + kRing;          // [kRing]
  uint64_t* p0_full = w_empty + kRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p0_full + 1);
  __shared__ int s_abort;
  __shared__ int s_pipe;    // tensor-pipe token: one tile's layer is issued at a time

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    s_abort = 0;
    s_pipe = 0;
    for (int t = 0; t < kNT; ++t) { mbar_init(&a_ready[t], 8); mbar_init(&d_ready[t], 1); }
    for (int s = 0; s < kRing; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], kNT); }
    mbar_init(p0_full, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 256; i += kThreads) {
    if (kF16)      // (1, 1, 0 x 6) in fp16 per row of chunk 0
      reinterpret_cast<uint4*>(ones_tile)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    else
      ones_tile[i] = (i < 128) ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = tid; i < W + 4; i += kThreads) we[i] = p.packed[p.off_we + i];
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;

  if (warp < kEpiWarps) {
    // ============ epilogue of tile t: thread = (row, 64-column half) ==================
    setmaxnreg_inc<kEpiRegs>();
    const int t = warp >> 3;
    const int wq = warp & 3, half = (warp >> 2) & 1;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t col_x = tmem_base + lane_off + (uint32_t)(t * 2 * W + half * 64);
    const uint32_t col_y = col_x + W;
    const float* we_h = we + half * 64;
    uint32_t ph_d = 0;
    uint32_t peak2 = 0;     // kF16: running max of the packed operands (range check)
    for (int k = 0;; ++k) {
      int64_t tile0;
      const int n = work_item(p.work, k, tile0);
      if (n == 0) break;
      if (t >= n) continue;
      const int64_t tile_id = tile0 + t;
      const int64_t r = tile_id * 128 + row;
      const bool valid = r < p.R;
      if (half == 0) {   // step-0 operand: this row's action (zero padded) into X[:, 0:32)
        if (kF16) {      // 16 action dimensions as 8 packed columns
          uint32_t v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            v[q] = pack_half2((valid && 2 * q < p.A) ? p.act[r * p.A + 2 * q] : 0.f,
                              (valid && 2 * q + 1 < p.A) ? p.act[r * p.A + 2 * q + 1] : 0.f);
          tmem_st_n<8>(col_x, v);
        } else {
          uint32_t v[16];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              v[q] = (valid && c * 16 + q < p.A) ? tf32_rn_bits(__float_as_uint(p.act[r * p.A + c * 16 + q])) : 0u;
            tmem_st16(col_x + c * 16, v);
          }
        }
        tmem_st_wait();
      }
      // one arrival per warp: the warp's tensor-memory stores are complete and fenced
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[t]);
      // residual stream of this row's half, seeded with the hoisted observation projection
      float h[64];
      {
        const float4* hp4 = reinterpret_cast<const float4*>(
            p.hp + (valid ? (r / p.n_per_obs) : 0) * W + half * 64);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 x = __ldg(hp4 + c);
          h[c * 4 + 0] = x.x; h[c * 4 + 1] = x.y; h[c * 4 + 2] = x.z; h[c * 4 + 3] = x.w;
        }
      }
      // One epilogue step (see mlp_fwd_tmem_kernel): even = the accumulator is this block's
      // increment of the residual stream, h += D, next operand relu(h) -> X; odd = relu, round,
      // in place in Y; last = Dense(1) on the final residual stream.
      auto step = [&](auto even_tag, auto last_tag, int i) {
        constexpr bool kEven = decltype(even_tag)::value;
        constexpr bool kLast = decltype(last_tag)::value;
        mbar_wait(&d_ready[t], ph_d, abort_flag, p.status, 4000 + i);
        ph_d ^= 1;
        tc_fence_after();
        const uint32_t col_d = kEven ? ((i == 0) ? col_y : col_x) : col_y;
        const uint32_t col_a = kEven ? col_x : col_y;
        uint32_t mb[2];
        uint32_t nbits = 0;     // bit (31 - q) = operand q > 0
        float e = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          tmem_ld16(col_d + c * 16, v);
          tmem_ld_wait16(v);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            uint32_t xb = v[q];
            if (kEven) {
              const float x = h[c * 16 + q] + __uint_as_float(v[q]);
              h[c * 16 + q] = x;
              xb = __float_as_uint(x);
              if (kLast) e = fmaf(x, we_h[c * 16 + q], e);
            }
            if (!kLast) {
              // mask bit from the FMA pipe: the sign of 0 - x is set exactly when x > 0
              if (kSave) nbits = __funnelshift_l(__float_as_uint(__fsub_rn(0.f, __uint_as_float(xb))), nbits, 1);
              if (!kF16) xb = relu_tf32_bits(xb);     // relu + round in one VIADDMNMX
            }
            v[q] = xb;
          }
          if (kSave && !kLast && (c & 1)) { mb[c >> 1] = __brev(nbits); nbits = 0; }
          if (kF16 && !kLast) {
            // pack once: the packed pairs are the next operand (relu on the pairs) AND the saved
            // tile; 16 columns = 8 words = two 16-byte pieces of the tile
            uint32_t w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              w[q] = hmax2_bits(pack_half2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), 0u);
              peak2 = hmax2_bits(peak2, w[q]);
            }
            tmem_st_n<8>(col_a + (uint32_t)((c >> 1) * 32 + (c & 1) * 8), w);
            if (kSave == 2) {
              uint4* h_slot = p.xa16 + (tile_id * (nsteps - 1) + i) * (int64_t)(16 * 128) + row;
              h_slot[(half * 8 + c * 2 + 0) * 128] = make_uint4(w[0], w[1], w[2], w[3]);
              h_slot[(half * 8 + c * 2 + 1) * 128] = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
          if (!kF16 && !kLast) tmem_st16(col_a + c * 16, v);
          if (!kF16 && kSave == 2 && !kLast) {
            // what the tensor core reads of the operand (low 13 bits truncated), as fp16: the
            // operand carries the round-to-nearest offset, so convert toward zero
            uint4* h_slot = p.xa16 + (tile_id * (nsteps - 1) + i) * (int64_t)(16 * 128) + row;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint4 w;
              w.x = pack_half2_rz(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
              w.y = pack_half2_rz(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
              w.z = pack_half2_rz(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
              w.w = pack_half2_rz(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
              h_slot[(half * 8 + c * 2 + q) * 128] = w;
            }
          }
          if (kSave == 2 && kLast) {
            float4* hn_slot = reinterpret_cast<float4*>(p.hn_save) + tile_id * (int64_t)(32 * 128) + row;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              hn_slot[(half * 16 + c * 4 + q) * 128] =
                  make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                              __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
          }
        }
        if (kLast) {
          // Dense(1), mlp_ebm.py:91-94: the two halves of a row meet in shared memory
          if (half == 1) e_part[t * 128 + row] = e;
          tc_fence_before();
          asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");     // this tile's eight warps
          if (half == 0 && valid) p.E[r] = (e + e_part[t * 128 + row]) + we[W];
          // e_part[t] is next written 2nb + 1 barrier-ordered steps later
        } else {
          if (kSave)
            *reinterpret_cast<uint2*>(p.mask_save + TT::mask_index(tile_id, p.nb, i, row) + half * 2) =
                make_uint2(mb[0], mb[1]);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[t]);
        }
      };
      using T_ = std::true_type;
      using F_ = std::false_type;
#pragma unroll 1
      for (int j = 0; j < p.nb; ++j) {
        step(T_{}, F_{}, 2 * j);        // h_j (j = 0: action projection + hp)
        step(F_{}, F_{}, 2 * j + 1);    // first Dense of block j
      }
      step(T_{}, T_{}, 2 * p.nb);       // last block's second Dense + Dense(1)
    }
    // cvt.satfinite clamps at 65 504 = 0x7BFF: an operand that reached it left fp16's range
    if (kF16 && ((peak2 & 0xFFFFu) >= 0x7BFFu || (peak2 >> 16) >= 0x7BFFu) && p.status)
      atomicCAS(p.status, 0, 5002);
  } else if (warp == kEpiWarps) {
    // ===================== producer: one bulk copy per layer ==================
    setmaxnreg_dec<kServiceRegs>();
    if (elect_one()) {    // the action rows of Wp stay resident for the whole kernel
      if (kF16) {
        mbar_arrive_expect_tx(p0_full, (uint32_t)(W * 16 * 2));
        bulk_g2s(p0_buf, p.packed16 + p.off16_p0, (uint32_t)(W * 16 * 2), p0_full);
      } else {
        mbar_arrive_expect_tx(p0_full, (uint32_t)(W * p.KP * 4));
        bulk_g2s(p0_buf, p.packed + p.off_fwd_p, (uint32_t)(W * p.KP * 4), p0_full);
      }
    }
    __syncwarp();
    uint32_t q = 0;
    for (int k = 0;; ++k) {
      int64_t tile0;
      if (work_item(p.work, k, tile0) == 0) break;
      for (int i = 1; i < nsteps; ++i, ++q) {
        const int slot = q % kRing;
        const uint32_t ph = (q / kRing) & 1;
        if (elect_one()) {
          mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 4100 + i);
          mbar_arrive_expect_tx(&w_full[slot], (uint32_t)kSlot);
          if (kF16)
            bulk_g2s(ring + (size_t)slot * kSlot, p.packed16 + p.off16_fwd + (int64_t)(i - 1) * (kSlot / 2),
                     (uint32_t)kSlot, &w_full[slot]);
          else
            bulk_g2s(ring + (size_t)slot * kSlot,
                     p.packed + p.off_fwd_lb + (int64_t)(i - 1) * (kSlot / 4),
                     (uint32_t)kSlot, &w_full[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kEpiWarps + 1 + kNT) {
    setmaxnreg_dec<kServiceRegs>();     // idle member of the service warpgroup
  } else {
    // ===================== MMA issuer of tile t ================================
    setmaxnreg_dec<kServiceRegs>();
    const int t = warp - (kEpiWarps + 1);
    const uint32_t idesc = idesc_tf32(128, W);
    const uint32_t a_lbo = 128 * 16, b_lbo = W * 16, sbo = 128;
    constexpr uint32_t b_step16 = (2 * W * 16) >> 4;
    const uint64_t ones_desc = smem_desc(smem_u32(ones_tile), a_lbo, sbo);
    const uint32_t col_x = tmem_base + (uint32_t)(t * 2 * W);
    const uint32_t col_y = col_x + W;
    uint32_t q = 0;
    uint32_t ph_a = 0;
    for (int k = 0;; ++k) {
      int64_t tile0;
      const int n = work_item(p.work, k, tile0);
      if (n == 0) break;
      if (t >= n) {
        // no tile for this issuer in the item: still hand every layer's ring slot back
        for (int i = 1; i < nsteps; ++i, ++q) {
          if (elect_one()) {
            mbar_wait(&w_full[q % kRing], (q / kRing) & 1, abort_flag, p.status, 4450 + i);
            mbar_arrive(&w_empty[q % kRing]);
          }
          __syncwarp();
        }
        continue;
      }
      for (int i = 0; i < nsteps; ++i) {
        // step 0 and odd steps read X and write Y; even steps >= 2 read Y, write X
        const bool x_to_y = (i == 0) || (i & 1);
        const uint32_t a_tmem = x_to_y ? col_x : col_y;
        const uint32_t d_tmem = x_to_y ? col_y : col_x;
        const uint32_t slot = q % kRing;
        if (elect_one()) {
          // the weights first: they landed long ago, and a passed mbarrier wait still costs
          // 150-300 cycles -- spent while the epilogue is busy instead of behind its arrival
          if (i == 0) mbar_wait(p0_full, 0, abort_flag, p.status, 4400);
          else mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 4400 + i);
          mbar_wait(&a_ready[t], ph_a, abort_flag, p.status, 4300 + i);
          tc_fence_after();
          pipe_acquire(&s_pipe, abort_flag);
          if (p.dbg && blockIdx.x == 0 && k == 1 && i < 40) p.dbg[(i * kNT + t) * 2] = clock64();
          if (kF16) {
            const uint32_t idesc16 = idesc_f16(128, W);
            if (i == 0) {        // K = 16 action dimensions: one MMA
              mma_f16_ts(d_tmem, a_tmem, smem_desc(smem_u32(p0_buf), b_lbo, sbo), idesc16, 0u);
            } else {
              const uint32_t layer = smem_u32(ring + (size_t)slot * kSlot);
              const uint64_t bd = smem_desc(layer, b_lbo, sbo);
#pragma unroll
              for (int ks = 0; ks < W / 16; ++ks)      // packed A: 8 columns per K-step, two per chunk
                mma_f16_ts(d_tmem, a_tmem + (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8),
                           bd + (uint64_t)(ks * b_step16), idesc16, (uint32_t)(ks > 0));
              mma_f16_ss(d_tmem, ones_desc, smem_desc(layer + W * W * 2, b_lbo, sbo), idesc16, 1u);   // + bias
              mma_commit(&w_empty[slot]);
            }
          } else if (i == 0) {
            const uint64_t b_desc0 = smem_desc(smem_u32(p0_buf), b_lbo, sbo);
            for (int ks = 0; ks < p.KP / 8; ++ks)
              mma_tf32_ts(d_tmem, a_tmem + (uint32_t)(ks * 8), b_desc0 + (uint64_t)(ks * b_step16),
                          idesc, (uint32_t)(ks > 0));
          } else {
            const uint32_t layer = smem_u32(ring + (size_t)slot * kSlot);
            mma_tf32_ts_steps<kKSteps>(d_tmem, a_tmem, smem_desc(layer, b_lbo, sbo), b_step16,
                                       idesc, 0u);
            // + bias: (1, 1, 0, ...) . (hi, lo, 0, ...)^T, the slice behind the weights
            mma_tf32_ss(d_tmem, ones_desc, smem_desc(layer + kLayerWBytes, b_lbo, sbo), idesc, 1u);
            mma_commit(&w_empty[slot]);     // free once every tile's issuer has arrived
          }
          mma_commit(&d_ready[t]);
          pipe_release(&s_pipe);
          if (p.dbg && blockIdx.x == 0 && k == 1 && i < 40) p.dbg[(i * kNT + t) * 2 + 1] = clock64();
        }
        __syncwarp();
        ph_a ^= 1;
        if (i > 0) ++q;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int kSave, bool kF16>
int launch_halves(ibc_handle* h, FwdHalvesParams& p, int grid, cudaStream_t s) {
  IBC_RAISE_SMEM(h, (mlp_fwd_halves_kernel<kSave, kF16>), 227 * 1024 - 512);
  mlp_fwd_halves_kernel<kSave, kF16><<<grid, kThreads, halves_smem_bytes(p.KP), s>>>(p);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

}  // namespace

// Energies only (xa16 == null) or the training forward of the block-major backward (fp16 operand
// tiles, fp32 h_nb, masks).
bool train_forward_is_f16(const Layout& L) {
  static const bool off = getenv("IBC_FWD_TF32") != nullptr || getenv("IBC_FWD_ROWS") != nullptr;
  return !off && L.W == 128 && L.A <= 16;
}

int tmem_forward_halves(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp,
                        float* E, void* xa16, float* hn_save, uint32_t* mask_save, cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  if (xa16 && (!hn_save || !mask_save))
    IBC_FAIL(h, IBC_ERR_INVALID, "tmem_forward_halves: fp16 tiles come with hn_save and mask_save");
  FwdHalvesParams p;
  p.act = act; p.hp = hp; p.packed = h->packed; p.E = E;
  p.xa16 = reinterpret_cast<uint4*>(xa16); p.hn_save = hn_save; p.mask_save = mask_save;
  p.R = R; p.n_per_obs = n; p.nb = L.nb; p.A = L.A; p.KP = pl.KP;
  const int64_t tiles = (R + 127) / 128;
  const int grid = tiles < h->sm_count ? (int)tiles : h->sm_count;
  p.work = make_tiling(R, grid);
  p.off_fwd_p = pl.fwd_p; p.off_fwd_lb = pl.fwd_lb;
  p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.status = h->dev_status;
  p.packed16 = nullptr; p.off16_p0 = 0; p.off16_fwd = 0;
  p.dbg = nullptr;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&p.dbg, 1024 * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(p.dbg, 0, 1024 * sizeof(long long), s));
  }
  // the training forward runs its GEMMs in kind::f16 (act_dim <= 16; IBC_FWD_TF32=1 keeps tf32)
  const bool f16 = xa16 && h->packed16 && train_forward_is_f16(L);
  if (f16) {
    const Packed16Layout p16 = make_packed16_layout(L);
    p.packed16 = reinterpret_cast<const __half*>(h->packed16);
    p.off16_p0 = p16.p0; p.off16_fwd = p16.fwd;
    IBC_TRY((launch_halves<2, true>(h, p, grid, s)));
  } else if (xa16) {
    IBC_TRY((launch_halves<2, false>(h, p, grid, s)));
  } else {
    IBC_TRY((launch_halves<0, false>(h, p, grid, s)));
  }
  if (p.dbg) {
    cudaStreamSynchronize(s);
    std::vector<long long> hb(1024);
    cudaMemcpy(hb.data(), p.dbg, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const long long t0 = hb[0];
    fprintf(stderr, "[ibc timeline halves] step | tile 0: mma_start mma_issued | tile 1: mma_start mma_issued\n");
    for (int i = 0; i < 2 * L.nb + 1 && i < 40; ++i)
      fprintf(stderr, "[ibc timeline halves] %3d | %7lld %7lld | %7lld %7lld\n", i, hb[(i * 2) * 2] - t0,
              hb[(i * 2) * 2 + 1] - t0, hb[(i * 2 + 1) * 2] - t0, hb[(i * 2 + 1) * 2 + 1] - t0);
  }
  return IBC_OK;
}

}  // namespace ibc
