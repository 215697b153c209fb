// Width-128 MLPEBM forward and reverse chain with the activations held in
// TENSOR MEMORY end to end (tcgen05 "TS" form: A operand from TMEM, B = weights
// from shared memory) -- networks/layers/resnet.py:135-153 and its reverse.
//
// Why: measured on B200 (scripts/mma_rate.py, scripts/timeline.py) a 128x128x8
// kind::tf32 MMA costs 139 cycles with both operands in un-swizzled shared
// memory and 217 when the other tile's epilogue is writing its next operand
// into shared memory at the same time, but 86 cycles with A in tensor memory.
// So the activations never touch shared memory here:
//   * two TMEM regions P, Q (128 columns each) per 128-row tile ping-pong as
//     "A operand of this layer" / "accumulator of this layer";
//   * the epilogue rewrites an accumulator IN PLACE into the next A operand
//     (tcgen05.ld -> bias/ReLU/round -> tcgen05.st, lane-local);
//   * the residual stream h (forward) / its gradient g (reverse) lives in the
//     epilogue thread's registers (128 fp32 per row), so no third region and
//     no bias folding are needed;
//   * shared memory only holds the weight ring (11 x 16 KB K-slices, TMA
//     bulk copies) and the tiny step-0 action operand.
// Two tiles per CTA (8 epilogue warps) overlap one tile's MMA with the other's
// epilogue; with few rows (policy inference) a CTA takes a single tile so that
// more SMs are busy.
#include <stdlib.h>

#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "ops.cuh"
#include "umma.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kW = 128;
constexpr int kNT = 2;
constexpr int kSliceBytes = 16384;
constexpr int kSliceK = kSliceBytes / (kW * 4);        // 32 K elements per slice
constexpr int kSlicesPerLayer = kW / kSliceK;          // 4
constexpr int kStepsPerSlice = kSliceK / 8;            // 4 MMAs (K = 8) per slice
constexpr int kThreads = (4 * kNT + 2) * 32;
constexpr int kActBytes = 128 * 32 * 4;                // step-0 operand tile (KP <= 32)
constexpr int kMaxSlots = 12;

struct TsSmemPlan {
  int nvec_floats;   // per-step bias vectors (+ we, be)
  int slots;
  size_t bytes;
};

inline TsSmemPlan ts_plan(int nb, bool forward) {
  TsSmemPlan p;
  p.nvec_floats = forward ? (2 * nb + 1) * kW + kW + 4 : kW + kW;   // reverse: we + dwe_acc
  const size_t fixed = (size_t)kNT * kActBytes + (size_t)p.nvec_floats * 4 + 512;
  const size_t budget = 227 * 1024 - 1024 /* static */;
  int slots = (int)((budget - fixed) / kSliceBytes);
  if (slots > kMaxSlots) slots = kMaxSlots;
  p.slots = slots;
  p.bytes = fixed + (size_t)slots * kSliceBytes;
  return p;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
struct FwdTsParams {
  const float* act;      // [R, A]
  const float* hp;       // [B, W]  obs . Wp[:O] + bp
  const float* packed;
  float* E;              // [R]
  float* xa_save;        // nullable, 2nb+1 slots per tile (training)
  int64_t R, n_per_obs;
  int nb, A, KP, nt, slots, num_groups;
  int64_t off_fwd_p, off_fwd_l, off_raw, off_we;
  int* status;
  long long* dbg;
};

__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_ts_kernel(FwdTsParams p) {
  using TT = TrainTile<kW>;
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* act_base = smem;                                   // [kNT][kActBytes]
  unsigned char* ring = smem + kNT * kActBytes;
  float* vecs = reinterpret_cast<float*>(ring + (size_t)p.slots * kSliceBytes);
  const int nsteps = 2 * p.nb + 1;
  const float* we = vecs + (size_t)nsteps * W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(vecs + (size_t)nsteps * W + W + 4);
  uint64_t* a_ready = bars;                    // [kNT]
  uint64_t* d_ready = bars + kNT;              // [kNT]
  uint64_t* w_full = bars + 2 * kNT;           // [kMaxSlots]
  uint64_t* w_empty = w_full + kMaxSlots;      // [kMaxSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kMaxSlots);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = p.nt;

  if (tid == 0) {
    s_abort = 0;
    for (int t = 0; t < kNT; ++t) { mbar_init(&a_ready[t], 128); mbar_init(&d_ready[t], 1); }
    for (int s = 0; s < kMaxSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < nsteps * W; i += kThreads) vecs[i] = p.packed[p.off_raw + i];
  for (int i = tid; i < W + 4; i += kThreads) vecs[nsteps * W + i] = p.packed[p.off_we + i];
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  const float be = vecs[nsteps * W + W];

  if (warp < 4 * kNT) {
    // ===================== epilogue warpgroup t ==============================
    const int t = warp >> 2;
    if (t < nt) {
      const int row = tid & 127;
      const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t col_p = (uint32_t)(t * 2 * W), col_q = (uint32_t)(t * 2 * W + W);
      float4* act_tile = reinterpret_cast<float4*>(act_base + t * kActBytes);
      uint32_t ph_d = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        const int64_t tile_id = (int64_t)g * nt + t;
        const int64_t r = tile_id * 128 + row;
        const bool valid = r < p.R;
        for (int c = 0; c < p.KP / 4; ++c) {          // step-0 operand: the action, zero padded
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) {
            const float* a = p.act + r * p.A;
            const int k = c * 4;
            if (k + 0 < p.A) o.x = a[k + 0];
            if (k + 1 < p.A) o.y = a[k + 1];
            if (k + 2 < p.A) o.z = a[k + 2];
            if (k + 3 < p.A) o.w = a[k + 3];
          }
          act_tile[c * 128 + row] = tf32_rn4(o);
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_ready[t]);
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
#pragma unroll 1
        for (int i = 0; i < nsteps; ++i) {
          const float* bias = vecs + (size_t)i * W;
          float4* g_slot = p.xa_save ? reinterpret_cast<float4*>(
              p.xa_save + TT::slot_base(tile_id, nsteps, i)) : nullptr;
          mbar_wait(&d_ready[t], ph_d, abort_flag, p.status, 100 + i);
          ph_d ^= 1;
          tc_fence_after();
          const bool stamp = p.dbg && blockIdx.x == 0 && g == (int)blockIdx.x && row == 0 && i < 64;
          if (stamp) p.dbg[(i * kNT + t) * 5 + 0] = clock64();
          const bool last = (i == nsteps - 1);
          if ((i & 1) == 0) {
            // P holds this block's increment of the residual stream
            float e = 0.f;
#pragma unroll
            for (int c32 = 0; c32 < W / 32; ++c32) {
              uint32_t v[32];
              tmem_ld32(tmem_base + lane_off + col_p + c32 * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                const float x = h[c32 * 32 + q] + (__uint_as_float(v[q]) + bias[c32 * 32 + q]);
                h[c32 * 32 + q] = x;
                if (last) e = fmaf(x, we[c32 * 32 + q], e);
                v[q] = __float_as_uint(last ? x : tf32_rn(fmaxf(x, 0.f)));
              }
              if (!last) tmem_st32(tmem_base + lane_off + col_p + c32 * 32, v);
              if (g_slot) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  g_slot[TT::chunk_f4(row, c32 * 8 + q)] =
                      make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                  __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
              }
            }
            if (last && valid) p.E[r] = e + be;          // Dense(1), mlp_ebm.py:91-94
          } else {
            // Q holds relu(h) . W1: bias, ReLU, round -> A operand of the second Dense
#pragma unroll
            for (int c32 = 0; c32 < W / 32; ++c32) {
              uint32_t v[32];
              tmem_ld32(tmem_base + lane_off + col_q + c32 * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 32; ++q)
                v[q] = __float_as_uint(
                    tf32_rn(fmaxf(__uint_as_float(v[q]) + bias[c32 * 32 + q], 0.f)));
              tmem_st32(tmem_base + lane_off + col_q + c32 * 32, v);
              if (g_slot) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  g_slot[TT::chunk_f4(row, c32 * 8 + q)] =
                      make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                                  __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
              }
            }
          }
          if (stamp) p.dbg[(i * kNT + t) * 5 + 1] = clock64();
          if (!last) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&a_ready[t]);
          } else {
            tc_fence_before();
          }
          if (stamp) p.dbg[(i * kNT + t) * 5 + 2] = clock64();
        }
      }
    }
  } else if (warp == 4 * kNT) {
    // ===================== producer: weight slices ============================
    if (lane == 0) {
      uint32_t q = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int i = 0; i < nsteps; ++i) {
          const int nsl = (i == 0) ? 1 : kSlicesPerLayer;
          const uint32_t bytes = (i == 0) ? (uint32_t)(W * p.KP * 4) : (uint32_t)kSliceBytes;
          const float* src = (i == 0) ? p.packed + p.off_fwd_p
                                      : p.packed + p.off_fwd_l + (int64_t)(i - 1) * W * W;
          for (int s = 0; s < nsl; ++s, ++q) {
            const int slot = q % p.slots;
            const uint32_t ph = (q / p.slots) & 1;
            mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 200 + i);
            mbar_arrive_expect_tx(&w_full[slot], bytes);
            bulk_g2s(ring + (size_t)slot * kSliceBytes, src + (int64_t)s * (kSliceBytes / 4),
                     bytes, &w_full[slot]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== MMA issuer ==========================================
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, W);
      const uint32_t a_lbo = 128 * 16, b_lbo = W * 16, sbo = 128;
      uint32_t q = 0;
      uint32_t ph_a[kNT] = {0, 0};
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int i = 0; i < nsteps; ++i) {
          const int nsl = (i == 0) ? 1 : kSlicesPerLayer;
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&a_ready[t], ph_a[t], abort_flag, p.status, 300 + i);
            ph_a[t] ^= 1;
            tc_fence_after();
            const bool mstamp = p.dbg && blockIdx.x == 0 && g == (int)blockIdx.x && i < 64;
            if (mstamp) p.dbg[(i * kNT + t) * 5 + 3] = clock64();
            const uint32_t col_p = tmem_base + (uint32_t)(t * 2 * W);
            const uint32_t col_q = col_p + W;
            // even steps write P (reading Q), odd steps write Q (reading P)
            const uint32_t d_tmem = (i & 1) ? col_q : col_p;
            const uint32_t a_tmem = (i & 1) ? col_p : col_q;
            const uint32_t act_addr = smem_u32(act_base + t * kActBytes);
            constexpr uint32_t b_step16 = (2 * W * 16) >> 4;
            if (i == 0) {
              const uint32_t slot = q % p.slots;
              if (t == 0) mbar_wait(&w_full[slot], (q / p.slots) & 1, abort_flag, p.status, 400);
              const uint64_t b_desc0 = smem_desc(smem_u32(ring + (size_t)slot * kSliceBytes), b_lbo, sbo);
              const uint64_t a_desc0 = smem_desc(act_addr, a_lbo, sbo);
              for (int ks = 0; ks < p.KP / 8; ++ks)
                mma_tf32_ss(d_tmem, a_desc0 + (uint64_t)(ks * ((2 * 128 * 16) >> 4)),
                            b_desc0 + (uint64_t)(ks * b_step16), idesc, (uint32_t)(ks > 0));
              if (t == nt - 1) mma_commit(&w_empty[slot]);
            } else {
#pragma unroll
              for (int s = 0; s < kSlicesPerLayer; ++s) {
                const uint32_t qq = q + s;
                const uint32_t slot = qq % p.slots;
                if (t == 0) mbar_wait(&w_full[slot], (qq / p.slots) & 1, abort_flag, p.status, 400 + i);
                const uint64_t b_desc0 = smem_desc(smem_u32(ring + (size_t)slot * kSliceBytes), b_lbo, sbo);
                mma_tf32_ts_steps<kStepsPerSlice>(d_tmem, a_tmem + (uint32_t)(s * kStepsPerSlice * 8),
                                                  b_desc0, b_step16, idesc, s == 0 ? 0u : 1u);
                if (t == nt - 1) mma_commit(&w_empty[slot]);
              }
            }
            mma_commit(&d_ready[t]);
            if (mstamp) { p.dbg[(i * kNT + t) * 5 + 4] = clock64(); p.dbg[640 + i * kNT + t] = 0; }
          }
          q += nsl;
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// reverse chain (training)
// ---------------------------------------------------------------------------
struct BwdTsParams {
  const float* dE;        // [R]
  const float* xa_save;   // forward operand tiles, 2nb+1 slots
  float* gy_save;         // output-gradient tiles, 2nb slots
  float* g0;              // [R, W] row-major gradient of the first Dense output
  const float* packed;
  int64_t off_rev_l, off_we;
  float* dwe;             // [W], atomics
  int64_t R;
  int nb, nt, slots, num_groups;
  int* status;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_ts_kernel(BwdTsParams p) {
  using TT = TrainTile<kW>;
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem + kNT * kActBytes;       // same carve-up as the forward
  float* we = reinterpret_cast<float*>(ring + (size_t)p.slots * kSliceBytes);
  float* dwe_acc = we + W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dwe_acc + W);
  uint64_t* a_ready = bars;
  uint64_t* d_ready = bars + kNT;
  uint64_t* w_full = bars + 2 * kNT;
  uint64_t* w_empty = w_full + kMaxSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kMaxSlots);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = p.nt;
  const int nsteps = 2 * p.nb;
  const int xa_slots = 2 * p.nb + 1, gy_slots = 2 * p.nb;

  if (tid == 0) {
    s_abort = 0;
    for (int t = 0; t < kNT; ++t) { mbar_init(&a_ready[t], 128); mbar_init(&d_ready[t], 1); }
    for (int s = 0; s < kMaxSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
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
    const int t = warp >> 2;
    if (t < nt) {
      const int row = tid & 127;
      const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t col_p = (uint32_t)(t * 2 * W), col_q = (uint32_t)(t * 2 * W + W);
      uint32_t ph_d = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        const int64_t tile_id = (int64_t)g * nt + t;
        const int64_t r = tile_id * 128 + row;
        const bool valid = r < p.R;
        const float de = valid ? p.dE[r] : 0.f;
        float gr[W];                        // gradient of the residual stream of this row
        {
          // g_top = dE (x) we ; dwe += h_nb^T dE (h_nb = the forward's last slot)
          const float4* hn = reinterpret_cast<const float4*>(
              p.xa_save + TT::slot_base(tile_id, xa_slots, 2 * p.nb));
          float4* gy_top = reinterpret_cast<float4*>(
              p.gy_save + TT::slot_base(tile_id, gy_slots, 2 * p.nb - 1));
#pragma unroll
          for (int c32 = 0; c32 < W / 32; ++c32) {
            uint32_t v[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = c32 * 8 + q;
              const float4 w4 = *reinterpret_cast<const float4*>(we + c * 4);
              const float4 g4 = make_float4(de * w4.x, de * w4.y, de * w4.z, de * w4.w);
              gr[c * 4 + 0] = g4.x; gr[c * 4 + 1] = g4.y; gr[c * 4 + 2] = g4.z; gr[c * 4 + 3] = g4.w;
              v[q * 4 + 0] = __float_as_uint(tf32_rn(g4.x));
              v[q * 4 + 1] = __float_as_uint(tf32_rn(g4.y));
              v[q * 4 + 2] = __float_as_uint(tf32_rn(g4.z));
              v[q * 4 + 3] = __float_as_uint(tf32_rn(g4.w));
              gy_top[TT::chunk_f4(row, c)] = g4;
              const float4 h4 = hn[TT::chunk_f4(row, c)];
              const float s0 = warp_sum(de * h4.x), s1 = warp_sum(de * h4.y);
              const float s2 = warp_sum(de * h4.z), s3 = warp_sum(de * h4.w);
              if (lane == 0) {
                atomicAdd(&dwe_acc[c * 4 + 0], s0); atomicAdd(&dwe_acc[c * 4 + 1], s1);
                atomicAdd(&dwe_acc[c * 4 + 2], s2); atomicAdd(&dwe_acc[c * 4 + 3], s3);
              }
            }
            tmem_st32(tmem_base + lane_off + col_p + c32 * 32, v);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&a_ready[t]);
        }
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
          const int i = nsteps - s;                   // forward layer index 2nb .. 1
          const bool odd = (i & 1) != 0;              // W1 layer: result adds onto g
          const bool last = (i == 1);
          const float4* xm = reinterpret_cast<const float4*>(
              p.xa_save + TT::slot_base(tile_id, xa_slots, i - 1));
          float4* gy_out = last ? nullptr : reinterpret_cast<float4*>(
              p.gy_save + TT::slot_base(tile_id, gy_slots, i - 2));
          mbar_wait(&d_ready[t], ph_d, abort_flag, p.status, 500 + s);
          ph_d ^= 1;
          tc_fence_after();
          // even layer (W2_j): D = Q, masked in place -> A operand of the W1_j step
          // odd  layer (W1_j): D = P, g += masked D -> A operand of the next W2 step
          const uint32_t col_d = odd ? col_p : col_q;
#pragma unroll
          for (int c32 = 0; c32 < W / 32; ++c32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_off + col_d + c32 * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = c32 * 8 + q;
              const float4 m = xm[TT::chunk_f4(row, c)];
              float4 o;
              o.x = m.x > 0.f ? __uint_as_float(v[q * 4 + 0]) : 0.f;
              o.y = m.y > 0.f ? __uint_as_float(v[q * 4 + 1]) : 0.f;
              o.z = m.z > 0.f ? __uint_as_float(v[q * 4 + 2]) : 0.f;
              o.w = m.w > 0.f ? __uint_as_float(v[q * 4 + 3]) : 0.f;
              if (odd) {
                o.x += gr[c * 4 + 0]; o.y += gr[c * 4 + 1]; o.z += gr[c * 4 + 2]; o.w += gr[c * 4 + 3];
                gr[c * 4 + 0] = o.x; gr[c * 4 + 1] = o.y; gr[c * 4 + 2] = o.z; gr[c * 4 + 3] = o.w;
              }
              if (!last) {
                gy_out[TT::chunk_f4(row, c)] = o;      // fp32: exact bias sums, MMA truncates
                v[q * 4 + 0] = __float_as_uint(tf32_rn(o.x));
                v[q * 4 + 1] = __float_as_uint(tf32_rn(o.y));
                v[q * 4 + 2] = __float_as_uint(tf32_rn(o.z));
                v[q * 4 + 3] = __float_as_uint(tf32_rn(o.w));
              } else if (valid) {
                *reinterpret_cast<float4*>(p.g0 + r * W + c * 4) = o;
              }
            }
            if (!last) tmem_st32(tmem_base + lane_off + col_d + c32 * 32, v);
          }
          if (!last) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&a_ready[t]);
          } else {
            tc_fence_before();
          }
        }
      }
    }
  } else if (warp == 4 * kNT) {
    if (lane == 0) {
      uint32_t q = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int s = 0; s < nsteps; ++s) {
          const float* src = p.packed + p.off_rev_l + (int64_t)s * W * W;
          for (int sl = 0; sl < kSlicesPerLayer; ++sl, ++q) {
            const int slot = q % p.slots;
            const uint32_t ph = (q / p.slots) & 1;
            mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 600 + s);
            mbar_arrive_expect_tx(&w_full[slot], (uint32_t)kSliceBytes);
            bulk_g2s(ring + (size_t)slot * kSliceBytes, src + (int64_t)sl * (kSliceBytes / 4),
                     (uint32_t)kSliceBytes, &w_full[slot]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, W);
      const uint32_t b_lbo = W * 16, sbo = 128;
      uint32_t q = 0;
      uint32_t ph_a[kNT] = {0, 0};
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int s = 0; s < nsteps; ++s) {
          const int i = nsteps - s;
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&a_ready[t], ph_a[t], abort_flag, p.status, 700 + s);
            ph_a[t] ^= 1;
            tc_fence_after();
            const uint32_t col_p = tmem_base + (uint32_t)(t * 2 * W);
            const uint32_t col_q = col_p + W;
            // even layer: A = P (g), D = Q ; odd layer: A = Q (u), D = P
            const uint32_t a_tmem = (i & 1) ? col_q : col_p;
            const uint32_t d_tmem = (i & 1) ? col_p : col_q;
            constexpr uint32_t b_step16 = (2 * W * 16) >> 4;
#pragma unroll
            for (int sl = 0; sl < kSlicesPerLayer; ++sl) {
              const uint32_t qq = q + sl;
              const uint32_t slot = qq % p.slots;
              if (t == 0) mbar_wait(&w_full[slot], (qq / p.slots) & 1, abort_flag, p.status, 800 + s);
              const uint64_t b_desc0 = smem_desc(smem_u32(ring + (size_t)slot * kSliceBytes), b_lbo, sbo);
              mma_tf32_ts_steps<kStepsPerSlice>(d_tmem, a_tmem + (uint32_t)(sl * kStepsPerSlice * 8),
                                                b_desc0, b_step16, idesc, sl == 0 ? 0u : 1u);
              if (t == nt - 1) mma_commit(&w_empty[slot]);
            }
            mma_commit(&d_ready[t]);
          }
          q += kSlicesPerLayer;
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  for (int i = tid; i < W; i += kThreads) atomicAdd(p.dwe + i, dwe_acc[i]);
}

// With few rows one tile per CTA keeps more SMs busy than two.
inline int choose_nt(int64_t R, int sm_count) {
  const int64_t tiles = (R + 127) / 128;
  return tiles <= sm_count ? 1 : 2;
}

}  // namespace

int ts_tiles_per_cta(ibc_handle* h, int64_t R) { return choose_nt(R, h->sm_count); }

int umma_forward_ts(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp,
                    float* E, float* xa_save, int nt, cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  const TsSmemPlan plan = ts_plan(L.nb, true);
  if (plan.slots < kSlicesPerLayer + 1)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "tcgen05 forward: depth %d leaves no room for the weight ring",
             L.D);
  FwdTsParams p;
  p.act = act; p.hp = hp; p.packed = h->packed; p.E = E; p.xa_save = xa_save;
  p.R = R; p.n_per_obs = n; p.nb = L.nb; p.A = L.A; p.KP = pl.KP; p.nt = nt;
  p.slots = plan.slots;
  p.num_groups = (int)((R + 128 * nt - 1) / (128 * nt));
  p.off_fwd_p = pl.fwd_p; p.off_fwd_l = pl.fwd_l; p.off_raw = pl.raw;
  p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.status = h->dev_status;
  p.dbg = nullptr;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&p.dbg, (640 + 128) * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(p.dbg, 0, (640 + 128) * sizeof(long long), s));
  }
  static bool attr = false;
  if (!attr) {
    IBC_CUDA(h, cudaFuncSetAttribute(mlp_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - 1024));
    IBC_CUDA(h, cudaFuncSetAttribute(mlp_bwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - 1024));
    attr = true;
  }
  const int grid = p.num_groups < h->sm_count ? p.num_groups : h->sm_count;
  mlp_fwd_ts_kernel<<<grid, kThreads, plan.bytes, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  if (p.dbg) {
    cudaStreamSynchronize(s);
    std::vector<long long> hb(640 + 128);
    cudaMemcpy(hb.data(), p.dbg, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const long long t0 = hb[3];
    fprintf(stderr, "[ibc timeline ts nt=%d] step tile | mma_start mma_issued | epi_start epi_done arrived\n", nt);
    for (int i = 0; i < 2 * L.nb + 1 && i < 64; ++i)
      for (int t = 0; t < nt; ++t) {
        const long long* e = &hb[(i * kNT + t) * 5];
        fprintf(stderr, "[ibc timeline ts] %3d %d | %8lld %8lld (w-wait %6lld) | %8lld %8lld %8lld\n", i, t,
                e[3] - t0, e[4] - t0, hb[640 + i * kNT + t], e[0] - t0, e[1] - t0, e[2] - t0);
      }
  }
  return IBC_OK;
}

int umma_backward_ts(ibc_handle* h, const float* dE, int64_t R, const float* xa_save,
                     float* gy_save, float* g0, float* dwe, int nt, cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  const TsSmemPlan plan = ts_plan(L.nb, false);
  BwdTsParams p;
  p.dE = dE; p.xa_save = xa_save; p.gy_save = gy_save; p.g0 = g0; p.packed = h->packed;
  p.off_rev_l = pl.rev_l; p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.dwe = dwe; p.R = R; p.nb = L.nb; p.nt = nt; p.slots = plan.slots;
  p.num_groups = (int)((R + 128 * nt - 1) / (128 * nt));
  p.status = h->dev_status;
  static bool attr = false;
  if (!attr) {
    IBC_CUDA(h, cudaFuncSetAttribute(mlp_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - 1024));
    IBC_CUDA(h, cudaFuncSetAttribute(mlp_bwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - 1024));
    attr = true;
  }
  const int grid = p.num_groups < h->sm_count ? p.num_groups : h->sm_count;
  mlp_bwd_ts_kernel<<<grid, kThreads, plan.bytes, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

}  // namespace ibc
