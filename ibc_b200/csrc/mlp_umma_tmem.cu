// Width-128 MLPEBM forward and reverse chain with the activations held in
// TENSOR MEMORY end to end (tcgen05 "TS" form: A operand from TMEM, B = weights
// from shared memory) -- networks/layers/resnet.py:135-153 and its reverse,
// ibc/agents/mcmc.py:161-191.
//
// Why (measured on B200 with scripts/mma_rate.py and scripts/timeline.py): a
// 128x128x8 kind::tf32 MMA runs at its 64-cycle floor from either operand
// source and under any concurrent shared-memory / tensor-memory / bulk-copy
// traffic, so the tensor pipe is never the problem; what limited the
// shared-memory-operand kernel (mlp_umma.cu, still used for width 256) was that
// two 64 KB activation tiles left room for only one layer of weights in the
// ring, so every layer waited ~800 cycles for its first slice (a 16 KB bulk
// copy takes ~780 cycles round trip from L2).  Here the activations never touch
// shared memory:
//   * two TMEM regions X, Y (128 columns each) per 128-row tile alternate as
//     "A operand of this layer" / "accumulator of this layer";
//   * the epilogue rewrites an accumulator IN PLACE into the next A operand
//     (tcgen05.ld -> ReLU/round -> tcgen05.st, lane-local);
//   * the residual stream h (forward) / its gradient g (reverse) lives in the
//     epilogue thread's registers (128 fp32 per row);
//   * shared memory holds the weight ring (up to 13 x 16 KB K-slices = three
//     layers of prefetch), the per-layer bias operands (2 KB each; the bias is
//     added by one extra K = 8 MMA against a constant (1, 1, 0, ...) tile) and
//     nothing else.
// Two tiles per CTA (8 epilogue warps) overlap one tile's MMAs with the other's
// epilogue; with few rows (policy inference) a CTA takes a single tile so that
// more SMs are busy.  The MMA and producer warps stay converged and issue under
// elect.sync (umma::elect_one).
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
constexpr int kLayerWBytes = kW * kW * 4;               // one layer's operand: 64 KB
constexpr int kBiasBytes = 8 * kW * 4;                   // its K = 8 bias slice: 4 KB
constexpr int kKSteps = kW / 8;                          // 16 MMAs (K = 8) per layer
// 8 epilogue warps + one full service warpgroup (producer, one MMA issuer per
// tile, one idle warp) so that setmaxnreg can move registers from the service
// warpgroup to the epilogue warpgroups: each epilogue thread holds a 128-float
// residual row.
constexpr int kThreads = (4 * kNT + 4) * 32;
// register budgets after setmaxnreg (per instance: chosen so that ptxas does not spill)
constexpr int kFwdRegs = 232, kFwdSaveRegs = 224, kBwdRegs = 232, kServiceRegs = 40;
// The weight ring holds whole layers: one bulk copy and one full/empty barrier
// pair per layer.  (With 16 KB slices the producer, which does not get past an
// mbarrier wait while its previous bulk copy is in flight, could not stay ahead
// of the tensor pipe.)
constexpr int kRing = 3;
// reverse kernel: per-warp staging to write g0 row-major with coalesced stores
constexpr int kStageBytes = 32 * 16 * 4;
constexpr int kDbgWords = 1024;

// forward: ring (3 x 68 KB) | step-0 operand (KP x W) | ones tile (4 KB) | we, be | barriers
// reverse: ring (3 x 64 KB) | we | dwe accumulator | barriers
inline size_t smem_bytes(bool forward, int KP) {
  const size_t ring = (size_t)kRing * (kLayerWBytes + (forward ? kBiasBytes : 0));
  const size_t tail = forward ? (size_t)KP * kW * 4 + 4096 + (kW + 4) * 4
                              : 2 * (size_t)kW * 4 + (size_t)4 * kNT * kStageBytes;
  return ring + tail + (size_t)(2 * kNT + 2 * kRing + 1) * 8 + 64;
}

// Tensor-pipe token.  Two issuer threads sharing the pipe 1:1 finish their layers
// together, which puts both tiles' epilogues (and then both MMA phases) in lockstep
// and loses the MMA/epilogue overlap; handing the pipe over a layer at a time keeps
// the tiles half a period apart.
__device__ __forceinline__ void pipe_acquire(int* lock, volatile int* abort_flag) {
  while (atomicCAS(lock, 0, 1) != 0) {
    if (*abort_flag) return;
    __nanosleep(32);
  }
}
__device__ __forceinline__ void pipe_release(int* lock) { atomicExch(lock, 0); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
struct FwdTmemParams {
  const float* act;      // [R, A]
  const float* hp;       // [B, W]  obs . Wp[:O] + bp
  const float* packed;
  float* E;              // [R]
  float* xa_save;        // nullable, 2nb+1 slots per tile (training, fp32 tiles: kSave == 1)
  uint4* xa16;           // kSave == 2: fp16 operands [tile][2nb slots][16 chunks of 8][128 rows][16 B]
  float* hn_save;        // kSave == 2: final residual stream [tile][32 chunks][128 rows][4] (fp32)
  uint32_t* mask_save;   // nullable, ReLU masks of the 2nb W x W layers
  int64_t R, n_per_obs;
  int nb, A, KP;
  Tiling work;
  int64_t off_fwd_p, off_fwd_lb, off_we;
  int* status;
  long long* dbg;
  int dbg_mode;   // timeline experiments: 1 = epilogue does no work (MMA chain alone)
};

// kSave: 0 = energies only; 1 = fp32 operand tiles + masks (round 1's split backward, the
// dE/dact paths save masks only); 2 = fp16 operand tiles + masks for the block-major
// backward (mlp_bwd_wgrad.cu): a saved operand is relu(.) rounded to tf32 -- non-negative
// with a 10-bit mantissa, i.e. exactly an fp16 value inside fp16's exponent range -- so the
// weight-gradient GEMM runs in kind::f16 on half the bytes.
template <int kSave>
__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_tmem_kernel(FwdTmemParams p) {
  using TT = TrainTile<kW>;
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nsteps = 2 * p.nb + 1;
  constexpr int kLayerBytes = kLayerWBytes + kBiasBytes;
  unsigned char* ring = smem;                                         // [kRing][68 KB]
  unsigned char* p0_buf = ring + (size_t)kRing * kLayerBytes;         // step-0 operand [KP/4][W][4]
  float4* ones_tile = reinterpret_cast<float4*>(p0_buf + (size_t)p.KP * W * 4);   // [2][128]
  float* we = reinterpret_cast<float*>(ones_tile + 256);              // we[W], be
  uint64_t* bars = reinterpret_cast<uint64_t*>(we + W + 4);
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
    for (int t = 0; t < kNT; ++t) { mbar_init(&a_ready[t], 128); mbar_init(&d_ready[t], 1); }
    for (int s = 0; s < kRing; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], kNT); }
    mbar_init(p0_full, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 256; i += kThreads)
    ones_tile[i] = (i < 128) ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < W + 4; i += kThreads) we[i] = p.packed[p.off_we + i];
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  const float be = we[W];

  if (warp < 4 * kNT) {
    // ===================== epilogue warpgroup t ==============================
    setmaxnreg_inc<(kSave ? kFwdSaveRegs : kFwdRegs)>();
    const int t = warp >> 2;
    {
      const int row = tid & 127;
      const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t col_x = tmem_base + lane_off + (uint32_t)(t * 2 * W);
      const uint32_t col_y = col_x + W;
      uint32_t ph_d = 0;
      for (int k = 0;; ++k) {
        int64_t tile0;
        const int n = work_item(p.work, k, tile0);
        if (n == 0) break;
        if (t >= n) continue;
        const int64_t tile_id = tile0 + t;
        const int64_t r = tile_id * 128 + row;
        const bool valid = r < p.R;
        {   // step-0 operand: this row's action (zero padded) into X[:, 0:32)
          uint32_t v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k)
            v[k] = (valid && k < p.A) ? tf32_rn_bits(__float_as_uint(p.act[r * p.A + k])) : 0u;
          tmem_st32(col_x, v);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&a_ready[t]);
        }
        // residual stream of this row, seeded with the hoisted observation projection
        float h[W];
        {
          const float4* hp4 = reinterpret_cast<const float4*>(
              p.hp + (valid ? (r / p.n_per_obs) : 0) * W);
#pragma unroll
          for (int c = 0; c < W / 4; ++c) {
            const float4 v = hp4[c];
            h[c * 4 + 0] = v.x; h[c * 4 + 1] = v.y; h[c * 4 + 2] = v.z; h[c * 4 + 3] = v.w;
          }
        }
        // One epilogue step.  kEven: the accumulator (Y after the action projection,
        // X after a second Dense) is this block's increment of the residual stream:
        // h += D, next operand relu(h) -> X.  Odd: Y holds relu(h) . W1 + b1 -> ReLU,
        // round, in place (the A operand of the second Dense).  kLast: Dense(1) on
        // the final residual stream instead of a next operand.  TMEM loads run one
        // 32-column chunk ahead of the arithmetic (two register buffers): under MMA
        // load a tcgen05.ld round trip is ~300 cycles.
        auto step = [&](auto even_tag, auto last_tag, int i) {
          constexpr bool kEven = decltype(even_tag)::value;
          constexpr bool kLast = decltype(last_tag)::value;
          float4* g_slot = (kSave == 1 && p.xa_save) ? reinterpret_cast<float4*>(
              p.xa_save + TT::slot_base(tile_id, nsteps, i)) : nullptr;
          uint4* h_slot = (kSave == 2 && !kLast)
                              ? p.xa16 + (tile_id * (nsteps - 1) + i) * (int64_t)(16 * 128) + row : nullptr;
          float4* hn_slot = (kSave == 2 && kLast)
                                ? reinterpret_cast<float4*>(p.hn_save) + tile_id * (int64_t)(32 * 128) + row
                                : nullptr;
          mbar_wait(&d_ready[t], ph_d, abort_flag, p.status, 100 + i);
          ph_d ^= 1;
          tc_fence_after();
          const bool stamp = p.dbg && blockIdx.x == 0 && k == 0 && row == 0 && i < 64;
          if (stamp) p.dbg[(i * kNT + t) * 5 + 0] = clock64();
          if (p.dbg_mode & 1) {
            if (!kLast) { tc_fence_before(); mbar_arrive(&a_ready[t]); }
            return;
          }
          const uint32_t col_d = kEven ? ((i == 0) ? col_y : col_x) : col_y;
          const uint32_t col_a = kEven ? col_x : col_y;
          uint32_t mb[W / 32];
          float e = 0.f;
          auto chunk = [&](int c32, uint32_t (&v)[32]) {
            uint32_t nbits = 0;     // bit (31 - q) = operand q > 0
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              uint32_t xb = v[q];
              if (kEven) {
                const float x = h[c32 * 32 + q] + __uint_as_float(v[q]);
                h[c32 * 32 + q] = x;
                xb = __float_as_uint(x);
                if (kLast) e = fmaf(x, we[c32 * 32 + q], e);
              }
              if (!kLast) {
                // relu + round in one VIADDMNMX.  The epilogue is bound by the ALU pipe (integer
                // add / min-max / shift issue every other cycle per scheduler), so the mask bit
                // comes from the FMA pipe: the sign of 0 - x is set exactly when x > 0 (+0 - (+-0)
                // = +0), shifted in with one funnel shift.
                if (kSave) nbits = __funnelshift_l(__float_as_uint(__fsub_rn(0.f, __uint_as_float(xb))), nbits, 1);
                xb = relu_tf32_bits(xb);
              }
              v[q] = xb;
            }
            if (kSave && !kLast) mb[c32] = __brev(nbits);
            if (!kLast) tmem_st32(col_a + c32 * 32, v);
            if (g_slot) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                g_slot[TT::chunk_f4(row, c32 * 8 + q)] =
                    make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
            }
            if (kSave == 2 && !kLast) {
              // what the tensor core reads of the operand (low 13 bits truncated), as fp16:
              // the operand carries the round-to-nearest offset, so convert toward zero
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 w;
                w.x = pack_half2_rz(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
                w.y = pack_half2_rz(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
                w.z = pack_half2_rz(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
                w.w = pack_half2_rz(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
                h_slot[(c32 * 4 + q) * 128] = w;
              }
            }
            if (kSave == 2 && kLast) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                hn_slot[(c32 * 8 + q) * 128] =
                    make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
            }
          };
          uint32_t va[32], vb[32];
          tmem_ld32(col_d, va);
          tmem_ld_wait(va); tmem_ld32(col_d + 32, vb); chunk(0, va);
          tmem_ld_wait(vb); tmem_ld32(col_d + 64, va); chunk(1, vb);
          tmem_ld_wait(va); tmem_ld32(col_d + 96, vb); chunk(2, va);
          tmem_ld_wait(vb); chunk(3, vb);
          if (kLast) {
            if (valid) p.E[r] = e + be;          // Dense(1), mlp_ebm.py:91-94
            tc_fence_before();
          } else {
            if (kSave && p.mask_save)
              *reinterpret_cast<uint4*>(p.mask_save + TT::mask_index(tile_id, p.nb, i, row)) =
                  make_uint4(mb[0], mb[1], mb[2], mb[3]);
            if (stamp) p.dbg[(i * kNT + t) * 5 + 1] = clock64();
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&a_ready[t]);
          }
          if (stamp) p.dbg[(i * kNT + t) * 5 + 2] = clock64();
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
    }
  } else if (warp == 4 * kNT) {
    // ===================== producer: one bulk copy per layer ==================
    setmaxnreg_dec<kServiceRegs>();
    if (elect_one()) {    // the action rows of Wp stay resident for the whole kernel
      mbar_arrive_expect_tx(p0_full, (uint32_t)(W * p.KP * 4));
      bulk_g2s(p0_buf, p.packed + p.off_fwd_p, (uint32_t)(W * p.KP * 4), p0_full);
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
          mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 200 + i);
          mbar_arrive_expect_tx(&w_full[slot], (uint32_t)kLayerBytes);
          bulk_g2s(ring + (size_t)slot * kLayerBytes,
                   p.packed + p.off_fwd_lb + (int64_t)(i - 1) * (kLayerBytes / 4),
                   (uint32_t)kLayerBytes, &w_full[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 * kNT + 1 + kNT) {
    setmaxnreg_dec<kServiceRegs>();     // idle members of the service warpgroup
  } else {
    // ===================== MMA issuer of tile t ================================
    // One issuer warp per tile: the thread that issued a tile's MMAs does not get
    // past its next mbarrier wait until they have drained (~400 cycles, measured),
    // so a single issuer serialises the two tiles and idles the tensor pipe between
    // them; with two issuers the other tile's MMAs are already queued.
    setmaxnreg_dec<kServiceRegs>();
    const int t = warp - (4 * kNT + 1);
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
            mbar_wait(&w_full[q % kRing], (q / kRing) & 1, abort_flag, p.status, 450 + i);
            mbar_arrive(&w_empty[q % kRing]);
          }
          __syncwarp();
        }
        continue;
      }
      for (int i = 0; i < nsteps; ++i) {
        const bool mstamp = p.dbg && blockIdx.x == 0 && k == 0 && i < 64;
        // step 0 and odd steps read X and write Y; even steps >= 2 read Y, write X
        const bool x_to_y = (i == 0) || (i & 1);
        const uint32_t a_tmem = x_to_y ? col_x : col_y;
        const uint32_t d_tmem = x_to_y ? col_y : col_x;
        const uint32_t slot = q % kRing;
        if (elect_one()) {
          // The layer's weights before the first MMA: an mbarrier wait issued behind
          // tcgen05.mma instructions does not return until they have (nearly) drained
          // (~100 cycles each, measured) -- and before the operand wait: a passed wait still
          // costs 150-300 cycles, better spent while the epilogue is busy.
          if (i == 0) mbar_wait(p0_full, 0, abort_flag, p.status, 400);
          else mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 400 + i);
          mbar_wait(&a_ready[t], ph_a, abort_flag, p.status, 300 + i);
          tc_fence_after();
          pipe_acquire(&s_pipe, abort_flag);
          if (mstamp) p.dbg[(i * kNT + t) * 5 + 3] = clock64();
          if (i == 0) {
            const uint64_t b_desc0 = smem_desc(smem_u32(p0_buf), b_lbo, sbo);
            for (int ks = 0; ks < p.KP / 8; ++ks)
              mma_tf32_ts(d_tmem, a_tmem + (uint32_t)(ks * 8), b_desc0 + (uint64_t)(ks * b_step16),
                          idesc, (uint32_t)(ks > 0));
          } else {
            const uint32_t layer = smem_u32(ring + (size_t)slot * kLayerBytes);
            mma_tf32_ts_steps<kKSteps>(d_tmem, a_tmem, smem_desc(layer, b_lbo, sbo), b_step16,
                                       idesc, 0u);
            // + bias: (1, 1, 0, ...) . (hi, lo, 0, ...)^T, the slice behind the weights
            mma_tf32_ss(d_tmem, ones_desc, smem_desc(layer + kLayerWBytes, b_lbo, sbo), idesc, 1u);
            mma_commit(&w_empty[slot]);     // free once every tile's issuer has arrived
          }
          mma_commit(&d_ready[t]);
          pipe_release(&s_pipe);
          if (mstamp) p.dbg[(i * kNT + t) * 5 + 4] = clock64();
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

// ---------------------------------------------------------------------------
// reverse chain
// ---------------------------------------------------------------------------
struct BwdTmemParams {
  const float* dE;            // [R]; null = ones (dE/dact of the plain energy)
  const float* xa_save;       // forward tiles; only slot 2nb (h_nb) is read, for dwe
  const uint32_t* mask_save;  // ReLU masks of the 2nb layers
  float* gy_save;             // nullable: output-gradient tiles, 2nb slots (training)
  float* g0;                  // [R, W] row-major gradient of the first Dense output
  const float* packed;
  int64_t off_rev_l, off_we;
  float* dwe;                 // nullable [W], atomics
  int64_t R;
  int nb;
  Tiling work;
  int* status;
  long long* dbg;   // optional timeline, same layout as the forward's
  int dbg_item;     // which work item of CTA 0 is stamped
};

__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_tmem_kernel(BwdTmemParams p) {
  using TT = TrainTile<kW>;
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;                                         // [kRing][64 KB]
  float* we = reinterpret_cast<float*>(ring + (size_t)kRing * kLayerWBytes);
  float* dwe_acc = we + W;
  float4* stage_all = reinterpret_cast<float4*>(dwe_acc + W);          // [8 warps][32 rows][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + 4 * kNT * (kStageBytes / 16));
  uint64_t* a_ready = bars;
  uint64_t* d_ready = bars + kNT;
  uint64_t* w_full = bars + 2 * kNT;
  uint64_t* w_empty = w_full + kRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kRing + 1);
  __shared__ int s_abort;
  __shared__ int s_pipe;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsteps = 2 * p.nb;
  const int xa_slots = 2 * p.nb + 1, gy_slots = 2 * p.nb;

  if (tid == 0) {
    s_abort = 0;
    s_pipe = 0;
    for (int t = 0; t < kNT; ++t) { mbar_init(&a_ready[t], 128); mbar_init(&d_ready[t], 1); }
    for (int s = 0; s < kRing; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], kNT); }
    fence_mbar_init();
  }
  for (int i = tid; i < W; i += kThreads) { we[i] = p.packed[p.off_we + i]; dwe_acc[i] = 0.f; }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;

  if (warp < 4 * kNT) {
    // ===================== epilogue warpgroup t ==============================
    setmaxnreg_inc<kBwdRegs>();
    const int t = warp >> 2;
    {
      const int row = tid & 127;
      const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t col_x = tmem_base + lane_off + (uint32_t)(t * 2 * W);
      const uint32_t col_y = col_x + W;
      uint32_t ph_d = 0;
      if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[999] = clock64();
      for (int k = 0;; ++k) {
        int64_t tile0;
        const int n = work_item(p.work, k, tile0);
        if (n == 0) break;
        if (t >= n) continue;
        const int64_t tile_id = tile0 + t;
        const int64_t r = tile_id * 128 + row;
        const bool valid = r < p.R;
        if (p.dbg && blockIdx.x == 0 && row == 0 && t == 0 && k < 8) p.dbg[960 + k * 4 + 0] = clock64();
        const float de = valid ? (p.dE ? p.dE[r] : 1.0f) : 0.f;
        // masks of the topmost layer now, every other layer one step ahead
        uint4 mnext = *reinterpret_cast<const uint4*>(
            p.mask_save + TT::mask_index(tile_id, p.nb, nsteps - 1, row));
        float gr[W];                        // gradient of the residual stream of this row
        {
          // g_top = dE (x) we ; dwe += h_nb^T dE (h_nb = the forward's last slot)
          const float4* hn = p.dwe ? reinterpret_cast<const float4*>(
              p.xa_save + TT::slot_base(tile_id, xa_slots, 2 * p.nb)) : nullptr;
          float4* gy_top = p.gy_save ? reinterpret_cast<float4*>(
              p.gy_save + TT::slot_base(tile_id, gy_slots, 2 * p.nb - 1)) : nullptr;
#pragma unroll
          for (int c32 = 0; c32 < W / 32; ++c32) {
            uint32_t v[32];
            // the chunk's eight h_nb loads together: one exposed latency per chunk, not
            // per float4 (the serial version cost ~60k cycles per work item, measured)
            float4 hbuf[8];
            if (hn) {
#pragma unroll
              for (int q = 0; q < 8; ++q) hbuf[q] = hn[TT::chunk_f4(row, c32 * 8 + q)];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = c32 * 8 + q;
              const float4 w4 = *reinterpret_cast<const float4*>(we + c * 4);
              const float4 g4 = make_float4(de * w4.x, de * w4.y, de * w4.z, de * w4.w);
              gr[c * 4 + 0] = g4.x; gr[c * 4 + 1] = g4.y; gr[c * 4 + 2] = g4.z; gr[c * 4 + 3] = g4.w;
              v[q * 4 + 0] = tf32_rn_bits(__float_as_uint(g4.x));
              v[q * 4 + 1] = tf32_rn_bits(__float_as_uint(g4.y));
              v[q * 4 + 2] = tf32_rn_bits(__float_as_uint(g4.z));
              v[q * 4 + 3] = tf32_rn_bits(__float_as_uint(g4.w));
              if (gy_top) gy_top[TT::chunk_f4(row, c)] = g4;   // fp32: exact bias sums
              if (hn) {
                const float4 h4 = hbuf[q];
                const float s0 = warp_sum(de * h4.x), s1 = warp_sum(de * h4.y);
                const float s2 = warp_sum(de * h4.z), s3 = warp_sum(de * h4.w);
                if (lane == 0) {
                  atomicAdd(&dwe_acc[c * 4 + 0], s0); atomicAdd(&dwe_acc[c * 4 + 1], s1);
                  atomicAdd(&dwe_acc[c * 4 + 2], s2); atomicAdd(&dwe_acc[c * 4 + 3], s3);
                }
              }
            }
            tmem_st32(col_x + c32 * 32, v);
          }
          if (p.dbg && blockIdx.x == 0 && row == 0 && t == 0 && k < 8) p.dbg[960 + k * 4 + 1] = clock64();
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&a_ready[t]);
          if (p.dbg && blockIdx.x == 0 && row == 0 && t == 0 && k < 8) p.dbg[960 + k * 4 + 2] = clock64();
        }
        // One reverse step for forward layer i (2nb .. 1).
        //   even layer (W2_j): D = Y, masked in place -> A operand of the W1_j step
        //   odd  layer (W1_j): D = X, g += masked D  -> A operand of the next W2 step
        //   last (i = 1): g is the gradient of the first Dense output -> g0
        auto step = [&](auto odd_tag, auto last_tag, int i) {
          constexpr bool kOdd = decltype(odd_tag)::value;
          constexpr bool kLast = decltype(last_tag)::value;
          float4* gy_out = (kLast || !p.gy_save) ? nullptr : reinterpret_cast<float4*>(
              p.gy_save + TT::slot_base(tile_id, gy_slots, i - 2));
          const uint32_t mbits[4] = {mnext.x, mnext.y, mnext.z, mnext.w};
          if (!kLast)
            mnext = *reinterpret_cast<const uint4*>(
                p.mask_save + TT::mask_index(tile_id, p.nb, i - 2, row));
          mbar_wait(&d_ready[t], ph_d, abort_flag, p.status, 500 + i);
          ph_d ^= 1;
          tc_fence_after();
          const bool stamp = p.dbg && blockIdx.x == 0 && k == p.dbg_item && row == 0;
          const int si = 2 * p.nb - i;          // step number 0 .. 2nb-1
          if (stamp) p.dbg[(si * kNT + t) * 5 + 0] = clock64();
          const uint32_t col_d = kOdd ? col_x : col_y;
          auto chunk = [&](int c32, uint32_t (&v)[32]) {
            const uint32_t bits = mbits[c32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = c32 * 8 + q;
              float4 o;
              o.x = (bits >> (q * 4 + 0)) & 1u ? __uint_as_float(v[q * 4 + 0]) : 0.f;
              o.y = (bits >> (q * 4 + 1)) & 1u ? __uint_as_float(v[q * 4 + 1]) : 0.f;
              o.z = (bits >> (q * 4 + 2)) & 1u ? __uint_as_float(v[q * 4 + 2]) : 0.f;
              o.w = (bits >> (q * 4 + 3)) & 1u ? __uint_as_float(v[q * 4 + 3]) : 0.f;
              if (kOdd) {
                o.x += gr[c * 4 + 0]; o.y += gr[c * 4 + 1]; o.z += gr[c * 4 + 2]; o.w += gr[c * 4 + 3];
                gr[c * 4 + 0] = o.x; gr[c * 4 + 1] = o.y; gr[c * 4 + 2] = o.z; gr[c * 4 + 3] = o.w;
              }
              if (!kLast) {
                if (gy_out) gy_out[TT::chunk_f4(row, c)] = o;   // fp32: exact bias sums
                v[q * 4 + 0] = tf32_rn_bits(__float_as_uint(o.x));
                v[q * 4 + 1] = tf32_rn_bits(__float_as_uint(o.y));
                v[q * 4 + 2] = tf32_rn_bits(__float_as_uint(o.z));
                v[q * 4 + 3] = tf32_rn_bits(__float_as_uint(o.w));
              } else {
                v[q * 4 + 0] = __float_as_uint(o.x); v[q * 4 + 1] = __float_as_uint(o.y);
                v[q * 4 + 2] = __float_as_uint(o.z); v[q * 4 + 3] = __float_as_uint(o.w);
              }
            }
            if (!kLast) {
              tmem_st32(col_d + c32 * 32, v);
            } else {
              // g0 is row-major [R, W]: a thread owns a row, so its own stores would hit
              // one 128-byte line per lane.  Transpose 32 rows x 16 columns at a time
              // through a swizzled per-warp staging buffer: 8 rows x 64 B per store.
              float4* stage = stage_all + warp * (kStageBytes / 16);
              const int64_t row0 = tile_id * 128 + (warp & 3) * 32;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  stage[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_float4(
                      __uint_as_float(v[half * 16 + j * 4 + 0]), __uint_as_float(v[half * 16 + j * 4 + 1]),
                      __uint_as_float(v[half * 16 + j * 4 + 2]), __uint_as_float(v[half * 16 + j * 4 + 3]));
                __syncwarp();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const int rr = kk * 8 + (lane >> 2), jj = lane & 3;
                  const float4 x = stage[rr * 4 + (jj ^ ((rr >> 1) & 3))];
                  if (row0 + rr < p.R)
                    *reinterpret_cast<float4*>(p.g0 + (row0 + rr) * W + c32 * 32 + half * 16 + jj * 4) = x;
                }
                __syncwarp();
              }
            }
          };
          // TMEM loads one chunk ahead of the arithmetic (see the forward kernel)
          uint32_t va[32], vb[32];
          tmem_ld32(col_d, va);
          tmem_ld_wait(va); tmem_ld32(col_d + 32, vb); chunk(0, va);
          tmem_ld_wait(vb); tmem_ld32(col_d + 64, va); chunk(1, vb);
          tmem_ld_wait(va); tmem_ld32(col_d + 96, vb); chunk(2, va);
          tmem_ld_wait(vb); chunk(3, vb);
          if (stamp) p.dbg[(si * kNT + t) * 5 + 1] = clock64();
          if (!kLast) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&a_ready[t]);
          } else {
            tc_fence_before();
          }
          if (stamp) p.dbg[(si * kNT + t) * 5 + 2] = clock64();
        };
        using T_ = std::true_type;
        using F_ = std::false_type;
#pragma unroll 1
        for (int j = p.nb - 1; j >= 1; --j) {
          step(F_{}, F_{}, 2 * j + 2);
          step(T_{}, F_{}, 2 * j + 1);
        }
        step(F_{}, F_{}, 2);
        step(T_{}, T_{}, 1);
        if (p.dbg && blockIdx.x == 0 && row == 0 && k < 8) p.dbg[1000 + k * 2 + t] = clock64();
      }
    }
  } else if (warp == 4 * kNT) {
    // ===================== producer: one bulk copy per layer ==================
    setmaxnreg_dec<kServiceRegs>();
    uint32_t q = 0;
    for (int k = 0;; ++k) {
      int64_t tile0;
      if (work_item(p.work, k, tile0) == 0) break;
      for (int s = 0; s < nsteps; ++s, ++q) {
        const int slot = q % kRing;
        const uint32_t ph = (q / kRing) & 1;
        if (elect_one()) {
          mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 600 + s);
          mbar_arrive_expect_tx(&w_full[slot], (uint32_t)kLayerWBytes);
          bulk_g2s(ring + (size_t)slot * kLayerWBytes, p.packed + p.off_rev_l + (int64_t)s * W * W,
                   (uint32_t)kLayerWBytes, &w_full[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 * kNT + 1 + kNT) {
    setmaxnreg_dec<kServiceRegs>();     // idle members of the service warpgroup
  } else {
    // ===================== MMA issuer of tile t (see the forward kernel) =======
    setmaxnreg_dec<kServiceRegs>();
    const int t = warp - (4 * kNT + 1);
    const uint32_t idesc = idesc_tf32(128, W);
    const uint32_t b_lbo = W * 16, sbo = 128;
    constexpr uint32_t b_step16 = (2 * W * 16) >> 4;
    const uint32_t col_x = tmem_base + (uint32_t)(t * 2 * W);
    const uint32_t col_y = col_x + W;
    uint32_t q = 0;
    uint32_t ph_a = 0;
    for (int k = 0;; ++k) {
      int64_t tile0;
      const int n = work_item(p.work, k, tile0);
      if (n == 0) break;
      if (t >= n) {
        for (int s = 0; s < nsteps; ++s, ++q) {
          if (elect_one()) {
            mbar_wait(&w_full[q % kRing], (q / kRing) & 1, abort_flag, p.status, 850 + s);
            mbar_arrive(&w_empty[q % kRing]);
          }
          __syncwarp();
        }
        continue;
      }
      for (int s = 0; s < nsteps; ++s, ++q) {
        const int i = nsteps - s;
        // even layer: A = X (g), D = Y ; odd layer: A = Y (u), D = X
        const uint32_t a_tmem = (i & 1) ? col_y : col_x;
        const uint32_t d_tmem = (i & 1) ? col_x : col_y;
        const uint32_t slot = q % kRing;
        if (elect_one()) {
          mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 800 + s);
          mbar_wait(&a_ready[t], ph_a, abort_flag, p.status, 700 + s);
          tc_fence_after();
          pipe_acquire(&s_pipe, abort_flag);
          const bool mstamp = p.dbg && blockIdx.x == 0 && k == p.dbg_item;
          if (mstamp) p.dbg[(s * kNT + t) * 5 + 3] = clock64();
          mma_tf32_ts_steps<kKSteps>(
              d_tmem, a_tmem, smem_desc(smem_u32(ring + (size_t)slot * kLayerWBytes), b_lbo, sbo),
              b_step16, idesc, 0u);
          mma_commit(&w_empty[slot]);
          mma_commit(&d_ready[t]);
          pipe_release(&s_pipe);
          if (mstamp) p.dbg[(s * kNT + t) * 5 + 4] = clock64();
        }
        __syncwarp();
        ph_a ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  if (p.dwe)
    for (int i = tid; i < W; i += kThreads) atomicAdd(p.dwe + i, dwe_acc[i]);
}

template <typename K>
int raise_smem(ibc_handle* h, K kernel) {
  IBC_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024 - 1024));
  return IBC_OK;
}

}  // namespace

bool tmem_path_supported(const Layout& L) {
  return L.W == kW && L.A <= 32 && smem_bytes(true, round_up(L.A, 8)) <= 227 * 1024 - 1024;
}

int tmem_forward(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp, float* E,
                 float* xa_save, uint32_t* mask_save, cudaStream_t s) {
  return tmem_forward_f16(h, act, R, n, hp, E, xa_save, nullptr, nullptr, mask_save, s);
}

// xa16 / hn_save non-null: fp16 operand tiles for the block-major backward (and no fp32 tiles)
int tmem_forward_f16(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp, float* E,
                     float* xa_save, void* xa16, float* hn_save, uint32_t* mask_save,
                     cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  const bool saving = xa_save || mask_save;
  if (xa16 && (!hn_save || !mask_save || xa_save))
    IBC_FAIL(h, IBC_ERR_INVALID, "tmem_forward_f16: fp16 tiles come with hn_save and mask_save only");
  // energies only and the fp16-saving forward: eight epilogue warps per tile (mlp_fwd_halves.cu);
  // IBC_FWD_ROWS=1 keeps this file's one-thread-per-row kernel for comparison
  static const bool rows = getenv("IBC_FWD_ROWS") != nullptr;
  if (!rows && !xa_save && (xa16 || !mask_save))
    return tmem_forward_halves(h, act, R, n, hp, E, xa16, hn_save, mask_save, s);
  FwdTmemParams p;
  p.act = act; p.hp = hp; p.packed = h->packed; p.E = E; p.xa_save = xa_save;
  p.xa16 = reinterpret_cast<uint4*>(xa16); p.hn_save = hn_save;
  p.mask_save = mask_save;
  p.R = R; p.n_per_obs = n; p.nb = L.nb; p.A = L.A; p.KP = pl.KP;
  const int64_t tiles = (R + 127) / 128;
  const int grid = tiles < h->sm_count ? (int)tiles : h->sm_count;
  p.work = make_tiling(R, grid);
  p.off_fwd_p = pl.fwd_p; p.off_fwd_lb = pl.fwd_lb;
  p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.status = h->dev_status;
  p.dbg = nullptr;
  p.dbg_mode = getenv("IBC_DBG_MODE") ? atoi(getenv("IBC_DBG_MODE")) : 0;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&p.dbg, kDbgWords * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(p.dbg, 0, kDbgWords * sizeof(long long), s));
  }
  constexpr int kMaxDev = 64;
  static bool attr[kMaxDev] = {false};          // the attribute is per device
  if (h->device >= 0 && h->device < kMaxDev && !attr[h->device]) {
    IBC_TRY(raise_smem(h, mlp_fwd_tmem_kernel<0>));
    IBC_TRY(raise_smem(h, mlp_fwd_tmem_kernel<1>));
    IBC_TRY(raise_smem(h, mlp_fwd_tmem_kernel<2>));
    attr[h->device] = true;
  }
  const size_t smem = smem_bytes(true, p.KP);
  if (xa16) mlp_fwd_tmem_kernel<2><<<grid, kThreads, smem, s>>>(p);
  else if (saving) mlp_fwd_tmem_kernel<1><<<grid, kThreads, smem, s>>>(p);
  else mlp_fwd_tmem_kernel<0><<<grid, kThreads, smem, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  if (p.dbg) {
    cudaStreamSynchronize(s);
    std::vector<long long> hb(kDbgWords);
    cudaMemcpy(hb.data(), p.dbg, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const long long t0 = hb[3];
    fprintf(stderr, "[ibc timeline tmem rounds=%d tail_single=%d] step tile | mma_start mma_issued | epi_start epi_done arrived\n",
            p.work.rounds, p.work.tail_single);
    for (int i = 0; i < 2 * L.nb + 1 && i < 64; ++i)
      for (int t = 0; t < kNT; ++t) {
        const long long* e = &hb[(i * kNT + t) * 5];
        fprintf(stderr, "[ibc timeline tmem] %3d %d | %8lld %8lld | %8lld %8lld %8lld\n", i, t,
                e[3] - t0, e[4] - t0, e[0] - t0, e[1] - t0, e[2] - t0);
      }
  }
  return IBC_OK;
}

int tmem_reverse(ibc_handle* h, const float* dE, int64_t R, const float* xa_save,
                 const uint32_t* mask_save, float* gy_save, float* g0, float* dwe,
                 cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  BwdTmemParams p;
  p.dE = dE; p.xa_save = xa_save; p.mask_save = mask_save; p.gy_save = gy_save; p.g0 = g0;
  p.packed = h->packed;
  p.off_rev_l = pl.rev_l; p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.dwe = dwe; p.R = R; p.nb = L.nb;
  const int64_t tiles = (R + 127) / 128;
  const int grid = tiles < h->sm_count ? (int)tiles : h->sm_count;
  p.work = make_tiling(R, grid);
  p.status = h->dev_status;
  constexpr int kMaxDev = 64;
  static bool attr[kMaxDev] = {false};          // the attribute is per device
  if (h->device >= 0 && h->device < kMaxDev && !attr[h->device]) {
    IBC_TRY(raise_smem(h, mlp_bwd_tmem_kernel));
    attr[h->device] = true;
  }
  p.dbg = nullptr;
  p.dbg_item = getenv("IBC_DBG_ITEM") ? atoi(getenv("IBC_DBG_ITEM")) : 0;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&p.dbg, kDbgWords * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(p.dbg, 0, kDbgWords * sizeof(long long), s));
  }
  mlp_bwd_tmem_kernel<<<grid, kThreads, smem_bytes(false, 0), s>>>(p);
  IBC_LAUNCH_CHECK(h);
  if (p.dbg) {
    cudaStreamSynchronize(s);
    std::vector<long long> hb(kDbgWords);
    cudaMemcpy(hb.data(), p.dbg, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const long long t0 = hb[3];
    for (int k = 0; k < 8; ++k)
      if (hb[960 + k * 4])
        fprintf(stderr, "[ibc timeline rev] CTA 0 tile 0 item %d: start %lld, top-of-chain done %lld, arrived %lld\n", k,
                hb[960 + k * 4] - hb[999], hb[960 + k * 4 + 1] - hb[999], hb[960 + k * 4 + 2] - hb[999]);
    for (int k = 0; k < 8; ++k)
      if (hb[1000 + k * 2] || hb[1000 + k * 2 + 1])
        fprintf(stderr, "[ibc timeline rev] CTA 0 item %d done at %lld / %lld cycles after kernel start\n", k,
                hb[1000 + k * 2] ? hb[1000 + k * 2] - hb[999] : 0,
                hb[1000 + k * 2 + 1] ? hb[1000 + k * 2 + 1] - hb[999] : 0);
    fprintf(stderr, "[ibc timeline rev] step tile | mma_start mma_issued | epi_start epi_done arrived\n");
    for (int i = 0; i < 2 * L.nb && i < 64; ++i)
      for (int t = 0; t < kNT; ++t) {
        const long long* e = &hb[(i * kNT + t) * 5];
        fprintf(stderr, "[ibc timeline rev] %3d %d | %8lld %8lld | %8lld %8lld %8lld\n", i, t,
                e[3] - t0, e[4] - t0, e[0] - t0, e[1] - t0, e[2] - t0);
      }
  }
  return IBC_OK;
}

}  // namespace ibc
