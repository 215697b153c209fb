// mcmc.langevin_actions_given_obs (ibc/agents/mcmc.py:330-425) for width-128 stacks with up to 8
// action dimensions (pushing_states/mlp_ebm_langevin.gin: W = 128, D = 16, A = 2): the persistent
// chain of langevin_tmem.cu with its dependent GEMMs in tcgen05 kind::f16 and a sixteen-warp
// epilogue.
//
// Why.  A CTA owns a 128-row tile and runs K iterations x (2nb + 1 forward + 2nb reverse) layers
// one after the other: the chain is a latency chain, 3 300 layer steps for K = 100, nb = 8.
// langevin_tmem.cu spends ~2 740 cycles per step: 1 088 for the sixteen K = 8 kind::tf32 MMAs and
// the bias MMA, ~460 until their commit is seen, ~1 000 in the four epilogue warps (a lone warp per
// scheduler, 128 columns per thread), ~200 to hand the operand back.  4.76 ms per chain against a
// floor of 2.7 ms for MMA + commit alone, whatever the epilogue does.
//   * The operands of those GEMMs are activations rounded to tf32 (an 11-bit significand) and
//     weights rounded the same way.  fp16 carries the same significand; what it lacks is exponent
//     range, and this path does not need it: relu outputs and residual gradients of these nets are
//     O(1e-4 .. 1e2) against fp16's 6e-5 (normal) .. 65 504 (values below 6e-5 keep an absolute
//     error <= 3e-8, values above 60 000 raise kernel status 5001 instead of saturating silently).
//     kind::f16 takes K = 16 per instruction at the same 128 x 128 rate: 8 MMAs, ~290 cycles.
//     Accumulation stays fp32 in tensor memory, the residual stream and its gradient stay fp32 in
//     registers, biases are added by an MMA against a (hi, lo) fp16 split.
//     The A operand lives in tensor memory packed two K-elements per 32-bit column (even k in the
//     low half; validated on hardware, scripts/probe_f16.py), the weights in shared memory as
//     K-major 16-byte chunks [K/8][N][8 halves] (ring of four 36 KB layers).
//   * Sixteen epilogue warps, thread = (row, 32 columns): four warps per scheduler.
//   * dE/dact is one more MMA (N = 16) against the action rows of Wp, as in langevin256.cu; each
//     thread then updates two action dimensions of its row.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include <algorithm>
#include "mlp.cuh"
#include "philox.cuh"
#include "umma.cuh"
#include "umma_f16.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kW = 128;
constexpr int kAMax = 8;                                   // action dimensions (two per thread)
constexpr int kKP = 16;                                    // K of the action projection
constexpr int kAP = 16;                                    // N of the dE/dact product
constexpr int kLayerHBytes = kW * kW * 2;                  // 32 KB
constexpr int kBiasHBytes = 16 * kW * 2;                   // K = 16 bias slice: 4 KB
constexpr int kSlotBytes = kLayerHBytes + kBiasHBytes;     // 36 KB (reverse layers use 32)
constexpr int kRing = 4;
constexpr int kKSteps = kW / 16;                           // 8 MMAs per layer
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 2) * 32;             // + producer, MMA issuer
constexpr int kColD = 0, kColA = 128;                      // tensor memory: accumulator | packed operand

struct LangevinF16Params {
  const float* hp;          // [B, W]  obs . Wp[:O] + bp
  float* actions;           // [R, A] in / out
  const __half* packed16;   // fp16 operand images (Packed16Layout)
  const float* we;          // we[W], be (fp32, from the tf32 pack's vector block)
  const float* mn;          // [A]
  const float* mx;
  const float* normals;     // [K, R, A] or null (Philox)
  const float* steps;       // [K] device copy of the step-size schedule
  const uint64_t* seed_ptr; // optional device seed
  uint64_t seed;
  float* chain_actions;     // [K, R, A] or null
  float* chain_energies;    // [K, R] or null
  float* chain_grad_norms;  // [K, R] or null
  int64_t R, n_per_obs;
  int nb, A, K;
  int64_t off_p0, off_fwd, off_rev, off_revp;       // in halves
  float noise_scale, grad_clip, dclip_scale;
  int norm_type, apply_exp;
  int* status;
  long long* dbg;   // optional timeline (IBC_DBG_TIMELINE)
};

// ring | p0 (4 KB) | Wp act rows for dE/dact (4 KB) | ones tile (4 KB) | we, be | masks | e_part |
// n_part | acts | barriers
inline size_t f16_smem_bytes(int nb) {
  return (size_t)kRing * kSlotBytes + (size_t)kKP * kW * 2 + (size_t)kW * kAP * 2 + 4096 +
         (kW + 4) * 4 + (size_t)2 * nb * 128 * 16 + 4 * 128 * 4 + 4 * 128 * 4 + kAMax * 128 * 4 +
         (2 + 2 * kRing + 1) * 8 + 64;
}

// The warp's tensor-memory stores are complete and ordered: one arrival per warp (512 arrivals
// on one barrier cost ~200 cycles more per layer step than 16).
__device__ __forceinline__ void operand_ready(uint64_t* bar, int lane) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// 32 consecutive fp32 values of a row -> 16 packed fp16 pairs (round to nearest)
__device__ __forceinline__ void pack32(const float (&x)[32], uint32_t (&w)[16]) {
#pragma unroll
  for (int q = 0; q < 16; ++q) w[q] = pack_half2(x[2 * q], x[2 * q + 1]);
}

__global__ void __launch_bounds__(kThreads, 1) langevin_f16_kernel(LangevinF16Params p) {
  constexpr int W = kW;
  constexpr int DPT = kAMax / 4;                    // action dimensions per thread
  extern __shared__ __align__(128) unsigned char smem[];
  const int nsteps = 2 * p.nb + 1;                  // forward steps; the reverse chain has 2nb + 1 MMAs
  unsigned char* ring = smem;                                              // [kRing][36 KB]
  unsigned char* p0_buf = ring + (size_t)kRing * kSlotBytes;               // [KP/8][W][8] halves
  unsigned char* pr_buf = p0_buf + (size_t)kKP * W * 2;                    // [W/8][AP][8] halves
  uint4* ones_tile = reinterpret_cast<uint4*>(pr_buf + (size_t)W * kAP * 2);   // [2][128] x 8 halves
  float* we = reinterpret_cast<float*>(ones_tile + 256);                   // we[W], be
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(we + W + 4);              // [2nb][4][128]
  float* e_part = reinterpret_cast<float*>(mask_s + (size_t)2 * p.nb * 128 * 4);   // [4][128]
  float* n_part = e_part + 4 * 128;                                        // [4][128]
  float* acts = n_part + 4 * 128;                                          // [kAMax][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(acts + kAMax * 128);
  uint64_t* a_ready = bars;
  uint64_t* d_ready = bars + 1;
  uint64_t* w_full = bars + 2;             // [kRing]
  uint64_t* w_empty = w_full + kRing;      // [kRing]
  uint64_t* p_full = w_empty + kRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_full + 1);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (p.R + 127) / 128;

  if (tid == 0) {
    s_abort = 0;
    mbar_init(a_ready, kEpiWarps);
    mbar_init(d_ready, 1);
    for (int s = 0; s < kRing; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(p_full, 1);
    fence_mbar_init();
  }
  // A operand of the bias MMA: row m = (1, 1, 0, ...) over K = 16 -> chunk 0 holds (1, 1, 0 x 6)
  for (int i = tid; i < 256; i += kThreads)
    ones_tile[i] = (i < 128) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < W + 1; i += kThreads) we[i] = p.we[i];       // we[W] = be
  if (warp == 1) {        // max |we|: the scale of the reverse chain's top gradient
    float m = 0.f;
    for (int i = lane; i < W; i += 32) m = fmaxf(m, fabsf(p.we[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncwarp();
    if (lane == 0) we[W + 1] = m;
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  const float be = we[W];
  const uint64_t seed = p.seed_ptr ? *p.seed_ptr : p.seed;

  if (warp < kEpiWarps) {
    // ============ epilogue: thread = (row, 32-column chunk) ===========================
    const int wq = warp & 3, cc = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t col_d = tmem_base + lane_off + kColD + (uint32_t)(cc * 32);   // this chunk of D
    const uint32_t col_a = tmem_base + lane_off + kColA + (uint32_t)(cc * 16);   // ... of the operand
    const uint32_t col_act = tmem_base + lane_off + kColA;      // action operand: columns 0 .. 7
    const uint32_t col_g = tmem_base + lane_off + kColD;        // dE/dact: columns 0 .. A-1 of D
    const int a0 = cc * DPT;                                    // first action dimension of this thread
    uint32_t ph_d = 0;
    uint32_t peak2 = 0;     // running max of the packed operands' magnitudes (range check)
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int64_t r = tile * 128 + row;
      const bool valid = r < p.R;
      const float4* hp4 = reinterpret_cast<const float4*>(
          p.hp + (valid ? (r / p.n_per_obs) : 0) * W + cc * 32);
      {
        // this row's actions into shared memory and, packed, into the operand columns 0 .. 7
        // (group cc writes column cc = dimensions 2cc, 2cc + 1, and zeroes column 4 + cc)
        float x[DPT];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
          const int a = a0 + j;
          x[j] = (valid && a < p.A) ? p.actions[r * p.A + a] : 0.f;
          acts[a * 128 + row] = x[j];
        }
        uint32_t v[1] = {pack_half2(x[0], x[1])}, z[1] = {0u};
        tmem_st_n<1>(col_act + (uint32_t)cc, v);
        tmem_st_n<1>(col_act + 4 + (uint32_t)cc, z);      // dimensions 8 .. 15 of the K = 16 operand
        operand_ready(a_ready, lane);
      }
#pragma unroll 1
      for (int k = 0; k < p.K; ++k) {
        // the noise of this thread's action dimensions, while the first MMAs run
        float nz[DPT];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
          const int a = a0 + j;
          nz[j] = 0.f;
          if (valid && a < p.A) {
            const int64_t i = r * p.A + a;
            nz[j] = p.normals ? p.normals[(int64_t)k * p.R * p.A + i]
                              : philox_normal(seed, (uint64_t)k + 1, (uint64_t)i);
          }
        }
        float h[32];          // residual row chunk (forward), then its gradient (reverse)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = __ldg(hp4 + c);
          h[c * 4 + 0] = x.x; h[c * 4 + 1] = x.y; h[c * 4 + 2] = x.z; h[c * 4 + 3] = x.w;
        }
        float e = 0.f;
        // ---- forward steps (resnet.py:135-153): even = h += D, operand relu(h); odd = relu(D);
        // last = Dense(1) on the final residual stream instead of a next operand ----
        auto fstep = [&](auto even_tag, auto last_tag, int i) {
          constexpr bool kEven = decltype(even_tag)::value;
          constexpr bool kLast = decltype(last_tag)::value;
          mbar_wait(d_ready, ph_d, abort_flag, p.status, 5100 + i);
          ph_d ^= 1;
          tc_fence_after();
          const bool stamp = p.dbg && blockIdx.x == 0 && k == 1 && lane == 0 && (warp == 0 || warp == 15);
          long long* dbg = stamp ? p.dbg + (i * 2 + (warp == 15)) * 4 : nullptr;
          if (stamp) dbg[0] = clock64();
          uint32_t v[32];
          tmem_ld32(col_d, v);
          tmem_ld_wait(v);
          if (stamp) dbg[1] = clock64();
          float x[32];
          uint32_t nbits = 0;     // bit (31 - q) = operand q > 0
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            float t = __uint_as_float(v[q]);
            if (kEven) { t += h[q]; h[q] = t; }
            if (kLast) {
              e = fmaf(t, we[cc * 32 + q], e);
            } else {
              // mask bit from the FMA pipe: the sign of 0 - t is set exactly when t > 0
              nbits = __funnelshift_l(__float_as_uint(__fsub_rn(0.f, t)), nbits, 1);
              x[q] = t;
            }
          }
          if (!kLast) {
            // the epilogue is bound by instruction issue: relu and the range check run on the
            // packed pairs (one instruction per two elements each)
            uint32_t w[16];
            pack32(x, w);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              w[q] = hmax2_bits(w[q], 0u);
              peak2 = hmax2_bits(peak2, w[q]);
            }
            tmem_st16(col_a, w);
            mask_s[((size_t)i * 4 + cc) * 128 + row] = __brev(nbits);
            if (stamp) dbg[2] = clock64();
            operand_ready(a_ready, lane);
            if (stamp) dbg[3] = clock64();
          }
        };
        using T_ = std::true_type;
        using F_ = std::false_type;
#pragma unroll 1
        for (int j = 0; j < p.nb; ++j) {
          fstep(T_{}, F_{}, 2 * j);
          fstep(F_{}, F_{}, 2 * j + 1);
        }
        fstep(T_{}, T_{}, 2 * p.nb);
        // ---- Dense(1) (mlp_ebm.py:91-94): the four chunks of a row meet in shared memory ----
        e_part[cc * 128 + row] = e;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");      // epilogue warps only
        const float energy = ((e_part[row] + e_part[128 + row]) + (e_part[256 + row] + e_part[384 + row])) + be;
        const float de = valid ? (p.apply_exp ? expf(energy) : 1.0f) : 0.f;
        if (cc == 0 && p.chain_energies && valid)
          p.chain_energies[(int64_t)k * p.R + r] = p.apply_exp ? de : energy;

        // ---- reverse chain (mcmc.py:161-191): g_top = dE (x) we ----
        // The gradient of a row is carried times a power of two that puts its largest top
        // element in [256, 512): residual gradients shrink through the layers (1e-5 and below
        // on the test nets), under fp16's normal range 6e-5 where the significand thins out --
        // scaled, there are 2^23 below the top before that happens and 2^7 of headroom above
        // (checked: status 5001).  Rows are independent in the GEMMs, the scaling is exact, and
        // dE/dact is scaled back at the end.
        float sc = 1.f, isc = 1.f;
        {
          const float m = fabsf(de) * we[W + 1];
          if (m > 0.f && m < 3.0e38f) {
            int ex = (int)((__float_as_uint(m) >> 23) & 255u) - 127;       // floor(log2 m)
            ex = max(-100, min(100, ex));
            sc = __uint_as_float((uint32_t)(127 + 8 - ex) << 23);
            isc = __uint_as_float((uint32_t)(127 - 8 + ex) << 23);
          }
        }
        {
          float x[32];
          const float des = de * sc;
#pragma unroll
          for (int q = 0; q < 32; ++q) { h[q] = des * we[cc * 32 + q]; x[q] = h[q]; }
          uint32_t w[16];
          pack32(x, w);
          tmem_st16(col_a, w);
          operand_ready(a_ready, lane);
        }
        auto rstep = [&](auto odd_tag, int i) {
          constexpr bool kOdd = decltype(odd_tag)::value;
          const uint32_t bits = mask_s[((size_t)(i - 1) * 4 + cc) * 128 + row];
          mbar_wait(d_ready, ph_d, abort_flag, p.status, 5200 + i);
          ph_d ^= 1;
          tc_fence_after();
          const bool stamp = p.dbg && blockIdx.x == 0 && k == 1 && lane == 0 && (warp == 0 || warp == 15);
          long long* dbg = stamp ? p.dbg + ((nsteps + 2 * p.nb - i) * 2 + (warp == 15)) * 4 : nullptr;
          if (stamp) dbg[0] = clock64();
          uint32_t v[32];
          tmem_ld32(col_d, v);
          tmem_ld_wait(v);
          if (stamp) dbg[1] = clock64();
          float x[32];
          if (kOdd) {        // g += D (.) m: a predicated add per element
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              if ((bits >> q) & 1u) h[q] += __uint_as_float(v[q]);
              x[q] = h[q];
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) x[q] = ((bits >> q) & 1u) ? __uint_as_float(v[q]) : 0.f;
          }
          uint32_t w[16];
          pack32(x, w);
          if (kOdd) {
#pragma unroll
            for (int q = 0; q < 16; ++q) peak2 = hmax2_bits(peak2, w[q] & 0x7FFF7FFFu);
          }
          tmem_st16(col_a, w);              // after i == 1 this is g0, the operand of dE/dact
          if (stamp) dbg[2] = clock64();
          operand_ready(a_ready, lane);
          if (stamp) dbg[3] = clock64();
        };
#pragma unroll 1
        for (int j = p.nb; j >= 1; --j) {
          rstep(F_{}, 2 * j);
          rstep(T_{}, 2 * j - 1);
        }

        // ---- dE/dact = D[:, 0:A), langevin_step on this thread's action dimensions ----
        mbar_wait(d_ready, ph_d, abort_flag, p.status, 5300);
        ph_d ^= 1;
        tc_fence_after();
        float nrm = 0.f;
        {
          uint32_t v[DPT];
          tmem_ld_n<DPT>(col_g + (uint32_t)a0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float stepsize = p.steps[k];
          float nv[DPT];
#pragma unroll
          for (int j = 0; j < DPT; ++j) {
            const int a = a0 + j;
            nv[j] = 0.f;
            if (a < p.A) {
              // langevin_step (mcmc.py:238-262), as langevin_dact_update_kernel
              const int64_t i = r * p.A + a;
              float g = -(__uint_as_float(v[j]) * isc);        // de_dact = -dE/dact (mcmc.py:190)
              const float t = fabsf(g);
              if (p.norm_type == IBC_NORM_INF) nrm = fmaxf(nrm, t);
              else if (p.norm_type == IBC_NORM_L2) nrm += t * t;
              else nrm += t;
              if (p.grad_clip > 0.f) g = fminf(fmaxf(g, -p.grad_clip), p.grad_clip);
              g = __fadd_rn(__fmul_rn(0.5f, g), __fmul_rn(nz[j], p.noise_scale));
              float dl = __fmul_rn(stepsize, g);
              const float dc = __fmul_rn(p.dclip_scale, __fadd_rn(p.mx[a], -p.mn[a]));
              dl = fminf(fmaxf(dl, -dc), dc);
              float u = __fadd_rn(acts[a * 128 + row], -dl);
              u = fminf(fmaxf(u, p.mn[a]), p.mx[a]);
              nv[j] = u;
              acts[a * 128 + row] = u;
              if (valid) {
                if (p.chain_actions) p.chain_actions[(int64_t)k * p.R * p.A + i] = u;
                if (k + 1 == p.K) p.actions[i] = u;
              }
            }
          }
          if (k + 1 < p.K) {
            // the operand region held g0 until now: rewrite all eight columns of the action operand
            uint32_t w[1] = {pack_half2(nv[0], nv[1])}, z[1] = {0u};
            tmem_st_n<1>(col_act + (uint32_t)cc, w);
            tmem_st_n<1>(col_act + 4 + (uint32_t)cc, z);
          }
        }
        if (p.chain_grad_norms) {
          // the row's norm from the four groups' partial norms (the next write of n_part is
          // ordered behind this read by a whole iteration of barriers)
          n_part[cc * 128 + row] = nrm;
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (cc == 0 && valid) {
            float t;
            if (p.norm_type == IBC_NORM_INF)
              t = fmaxf(fmaxf(n_part[row], n_part[128 + row]), fmaxf(n_part[256 + row], n_part[384 + row]));
            else
              t = n_part[row] + n_part[128 + row] + n_part[256 + row] + n_part[384 + row];
            if (p.norm_type == IBC_NORM_L2) t = sqrtf(t);
            if (p.norm_type == IBC_NORM_NONE) t = 0.f;
            p.chain_grad_norms[(int64_t)k * p.R + r] = t;
          }
        }
        if (k + 1 < p.K) operand_ready(a_ready, lane);
      }
    }
    // cvt.satfinite clamps at 65 504 = 0x7BFF: an operand that reached it left fp16's range
    if (((peak2 & 0xFFFFu) >= 0x7BFFu || (peak2 >> 16) >= 0x7BFFu) && p.status) atomicCAS(p.status, 0, 5001);
  } else if (warp == kEpiWarps) {
    // ===================== producer: one bulk copy per layer ==================
    if (elect_one()) {
      const uint32_t b0 = (uint32_t)(kKP * W * 2), b1 = (uint32_t)(W * kAP * 2);
      mbar_arrive_expect_tx(p_full, b0 + b1);
      bulk_g2s(p0_buf, p.packed16 + p.off_p0, b0, p_full);
      bulk_g2s(pr_buf, p.packed16 + p.off_revp, b1, p_full);
    }
    __syncwarp();
    uint32_t q = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int k = 0; k < p.K; ++k) {
        for (int l = 0; l < 4 * p.nb; ++l, ++q) {       // 2nb forward layers, 2nb reverse layers
          const int slot = q % kRing;
          const uint32_t ph = (q / kRing) & 1;
          const bool fwd = l < 2 * p.nb;
          const __half* src = fwd ? p.packed16 + p.off_fwd + (int64_t)l * (kSlotBytes / 2)
                                  : p.packed16 + p.off_rev + (int64_t)(l - 2 * p.nb) * W * W;
          const uint32_t bytes = fwd ? (uint32_t)kSlotBytes : (uint32_t)kLayerHBytes;
          if (elect_one()) {
            mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 5400);
            mbar_arrive_expect_tx(&w_full[slot], bytes);
            bulk_g2s(ring + (size_t)slot * kSlotBytes, src, bytes, &w_full[slot]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== MMA issuer ==========================================
    const uint32_t idesc = idesc_f16(128, W), idesc_g = idesc_f16(128, kAP);
    const uint32_t sbo = 128;
    constexpr uint32_t b_step16 = (2 * W * 16) >> 4;            // K = 16 = two 16-byte chunks of [N] rows
    const uint64_t ones_desc = smem_desc(smem_u32(ones_tile), 128 * 16, sbo);
    const uint32_t col_d = tmem_base + kColD, col_a = tmem_base + kColA;
    uint32_t q = 0, ph_a = 0;
    auto layer = [&](uint32_t base, bool bias) {
#pragma unroll
      for (int ks = 0; ks < kKSteps; ++ks)
        mma_f16_ts(col_d, col_a + (uint32_t)(ks * 8), smem_desc(base, W * 16, sbo) + (uint64_t)(ks * b_step16),
                   idesc, (uint32_t)(ks > 0));
      // + bias: (1, 1, 0, ...) . (hi, lo, 0, ...)^T, the slice behind the weights
      if (bias) mma_f16_ss(col_d, ones_desc, smem_desc(base + kLayerHBytes, W * 16, sbo), idesc, 1u);
    };
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int k = 0; k < p.K; ++k) {
        for (int i = 0; i < nsteps; ++i) {
          const uint32_t slot = q % kRing;
          if (elect_one()) {
            // the weights first: they landed long ago, and a passed mbarrier wait still costs
            // 150-300 cycles -- spent here while the epilogue is busy, not behind its arrival
            if (i == 0) mbar_wait(p_full, 0, abort_flag, p.status, 5550);
            else mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 5560);
            mbar_wait(a_ready, ph_a, abort_flag, p.status, 5500 + i);
            tc_fence_after();
            if (p.dbg && blockIdx.x == 0 && k == 1) p.dbg[600 + i * 2] = clock64();
            if (i == 0) {
              mma_f16_ts(col_d, col_a, smem_desc(smem_u32(p0_buf), W * 16, sbo), idesc, 0u);   // K = 16
            } else {
              layer(smem_u32(ring + (size_t)slot * kSlotBytes), true);
              mma_commit(&w_empty[slot]);
            }
            mma_commit(d_ready);
            if (p.dbg && blockIdx.x == 0 && k == 1) p.dbg[600 + i * 2 + 1] = clock64();
          }
          __syncwarp();
          ph_a ^= 1;
          if (i > 0) ++q;
        }
        for (int s = 0; s < 2 * p.nb; ++s, ++q) {
          const uint32_t slot = q % kRing;
          if (elect_one()) {
            mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 5660);
            mbar_wait(a_ready, ph_a, abort_flag, p.status, 5600 + s);
            tc_fence_after();
            if (p.dbg && blockIdx.x == 0 && k == 1) p.dbg[600 + (nsteps + s) * 2] = clock64();
            layer(smem_u32(ring + (size_t)slot * kSlotBytes), false);
            mma_commit(&w_empty[slot]);
            mma_commit(d_ready);
            if (p.dbg && blockIdx.x == 0 && k == 1) p.dbg[600 + (nsteps + s) * 2 + 1] = clock64();
          }
          __syncwarp();
          ph_a ^= 1;
        }
        // dE/dact = g0 . Wp[O:]^T  (N = 16)
        if (elect_one()) {
          mbar_wait(a_ready, ph_a, abort_flag, p.status, 5700);
          tc_fence_after();
          const uint64_t bd = smem_desc(smem_u32(pr_buf), kAP * 16, sbo);
          constexpr uint32_t g_step16 = (2 * kAP * 16) >> 4;
#pragma unroll
          for (int ks = 0; ks < kKSteps; ++ks)
            mma_f16_ts(col_d, col_a + (uint32_t)(ks * 8), bd + (uint64_t)(ks * g_step16), idesc_g,
                       (uint32_t)(ks > 0));
          mma_commit(d_ready);
        }
        __syncwarp();
        ph_a ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// fp16 operand images of a width-128 stack: [KP/8][W][8] action rows of Wp | 2nb x ([W/8][W][8]
// layer + [2][W][8] bias slice (hi, lo, 0 ...)) | 2nb x [W/8][W][8] transposed layers, top first |
// [W/8][AP][8] action rows of Wp for dE/dact.  Element (n, k) of a K-major image sits at
// ((k / 8) * N + n) * 8 + k % 8.
// Elements [lo, hi) of the buffer, without [hole_lo, hole_lo + hole_len).
// Eight consecutive halves per thread -- every image and every range boundary is a multiple of
// eight, and a group is one (n, k / 8) cell of a K-major image: the source of its eight values is
// one base address and one stride (W for the forward images, 1 for the transposed ones), worked
// out once; eight independent loads, one 16-byte store.
__global__ void pack16_kernel(const float* __restrict__ P, Layout L, Packed16Layout pl,
                              __half* __restrict__ out, int64_t lo, int64_t hi, int64_t hole_lo,
                              int64_t hole_len) {
  const int W = L.W;
  const int64_t WW = (int64_t)W * W;
  const int64_t groups = (hi - lo - hole_len) >> 3;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j0 = lo + (g << 3);
    const int64_t i = j0 < hole_lo ? j0 : j0 + hole_len;
    float v[8];
    const float* src = nullptr;
    int64_t stride = 0;
    int valid = 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
    if (i < pl.fwd) {                       // step-0 operand: B[n][k] = Wp[O + k][n]
      const int64_t c = (i - pl.p0) >> 3;
      const int n = (int)(c % W), k0 = (int)(c / W) * 8;
      src = P + L.wp + (int64_t)(L.O + k0) * W + n;
      stride = W;
      valid = min(max(L.A - k0, 0), 8);
    } else if (i < pl.rev) {                // forward layer m + bias slice of step m + 1
      const int64_t e = i - pl.fwd;
      const int64_t blk = WW + 16 * W;
      const int m = (int)(e / blk);
      const int64_t r = e - (int64_t)m * blk;
      const int j = m >> 1;
      if (r < WW) {
        const int64_t c = r >> 3;
        const int n = (int)(c % W), k0 = (int)(c / W) * 8;
        src = P + ((m & 1) ? L.w2(j) : L.w1(j)) + (int64_t)k0 * W + n;
        stride = W;
      } else {                              // (hi, lo, 0 ...) of the bias, first K chunk only
        const int64_t c = (r - WW) >> 3;
        const int n = (int)(c % W), kc = (int)(c / W);
        valid = 0;
        if (kc == 0) {
          const float bv = (m & 1) ? P[L.b2(j) + n] : P[L.b1(j) + n];
          const float hi16 = __half2float(__float2half_rn(bv));
          v[0] = hi16;
          v[1] = bv - hi16;
        }
      }
    } else if (i < pl.revp) {               // reverse layers: B[n = in][k = out] = W[in][out]
      const int64_t e = i - pl.rev;
      const int m = (int)(e / WW);          // 0: W2_{nb-1}, 1: W1_{nb-1}, 2: W2_{nb-2}, ...
      const int64_t c = (e - (int64_t)m * WW) >> 3;
      const int irow = (int)(c % W), o0 = (int)(c / W) * 8;
      const int j = L.nb - 1 - (m >> 1);
      src = P + ((m & 1) ? L.w1(j) : L.w2(j)) + (int64_t)irow * W + o0;
      stride = 1;
    } else {                                // dE/dact operand: B[n = a][k = w] = Wp[O + a][w]
      const int64_t c = (i - pl.revp) >> 3;
      const int a = (int)(c % pl.AP), w0 = (int)(c / pl.AP) * 8;
      src = P + L.wp + (int64_t)(L.O + a) * W + w0;
      stride = 1;
      valid = a < L.A ? 8 : 0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < valid) v[q] = src[q * stride];
    __align__(16) __half2 h2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) h2[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(h2);
  }
}

}  // namespace

Packed16Layout make_packed16_layout(const Layout& L) {
  Packed16Layout p;
  const int64_t WW = (int64_t)L.W * L.W;
  p.KP = kKP; p.AP = kAP;
  p.p0 = 0;
  p.fwd = p.p0 + (int64_t)L.W * kKP;
  p.rev = p.fwd + (int64_t)(2 * L.nb) * (WW + 16 * L.W);
  p.revp = p.rev + (int64_t)(2 * L.nb) * WW;
  p.total = p.revp + (int64_t)L.W * kAP;
  return p;
}

bool langevin_f16_supported(const Layout& L) {
  const bool off = getenv("IBC_LANGEVIN_TF32") != nullptr;    // keep the kind::tf32 kernel
  return !off && L.W == kW && L.A <= kAMax && L.nb >= 1 && f16_smem_bytes(L.nb) <= 227 * 1024 - 512;
}

int pack16_params(ibc_handle* h, cudaStream_t s, bool core) {
  const Packed16Layout pl = make_packed16_layout(h->L);
  if (!h->packed16) IBC_CUDA(h, cudaMalloc((void**)&h->packed16, (size_t)pl.total * 2));
  __half* out = reinterpret_cast<__half*>(h->packed16);
  // core: action rows + forward layers (the training forward) + transposed layers (the backward);
  // the rest: the dE/dact operand of the Langevin chain
  // latency-bound gathers out of L2 (the parameters Adam just wrote): one group of eight per
  // thread in one wave instead of fifteen dependent rounds of single halves on 148 CTAs
  const int64_t cnt = (core ? pl.revp : pl.total - pl.revp) / 8;       // groups of eight halves
  if ((pl.fwd | pl.rev | pl.revp | pl.total) & 7)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "pack16: image sizes must be multiples of eight halves");
  const int grid = (int)std::min<int64_t>(std::max<int64_t>((cnt + 255) / 256, 1), 148 * 8);
  if (core) pack16_kernel<<<grid, 256, 0, s>>>(h->params, h->L, pl, out, 0, pl.revp, pl.revp, 0);
  else pack16_kernel<<<grid, 256, 0, s>>>(h->params, h->L, pl, out, pl.revp, pl.total, pl.total, 0);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// hp: [B, W] observation projection; steps_dev: [K] step sizes on the device.
int langevin_f16(ibc_handle* h, const float* hp, int64_t B, float* actions, int64_t n,
                 const float* mn, const float* mx, const ibc_langevin_cfg& c, const float* normals,
                 uint64_t seed, const uint64_t* seed_ptr, const float* steps_dev,
                 float* chain_actions, float* chain_energies, float* chain_grad_norms,
                 cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  const Packed16Layout p16 = make_packed16_layout(L);
  LangevinF16Params p;
  p.hp = hp; p.actions = actions; p.packed16 = reinterpret_cast<const __half*>(h->packed16);
  p.we = h->packed + pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.mn = mn; p.mx = mx; p.normals = normals; p.steps = steps_dev; p.seed_ptr = seed_ptr;
  p.seed = seed; p.chain_actions = chain_actions; p.chain_energies = chain_energies;
  p.chain_grad_norms = chain_grad_norms;
  p.R = B * n; p.n_per_obs = n; p.nb = L.nb; p.A = L.A; p.K = c.num_iterations;
  p.off_p0 = p16.p0; p.off_fwd = p16.fwd; p.off_rev = p16.rev; p.off_revp = p16.revp;
  p.noise_scale = c.noise_scale; p.grad_clip = c.grad_clip;
  p.dclip_scale = (float)((double)c.delta_action_clip * 0.5);
  p.norm_type = c.grad_norm_type; p.apply_exp = c.apply_exp;
  p.status = h->dev_status;
  IBC_RAISE_SMEM(h, langevin_f16_kernel, 227 * 1024 - 512);
  const int64_t tiles = (p.R + 127) / 128;
  const int grid = tiles < h->sm_count ? (int)tiles : h->sm_count;
  p.dbg = nullptr;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&p.dbg, 1024 * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(p.dbg, 0, 1024 * sizeof(long long), s));
  }
  langevin_f16_kernel<<<grid, kThreads, f16_smem_bytes(p.nb), s>>>(p);
  IBC_LAUNCH_CHECK(h);
  if (p.dbg) {
    cudaStreamSynchronize(s);
    long long hb[1024];
    cudaMemcpy(hb, p.dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const long long t0 = hb[600];
    fprintf(stderr, "[ibc timeline langevin_f16] step | mma_start mma_issued | warp 0: d_seen ld_done packed arrived | warp 15: ...\n");
    for (int i = 0; i < 4 * L.nb + 1 && i < 70; ++i)
      fprintf(stderr, "[ibc timeline langevin_f16] %3d | %7lld %7lld | %7lld %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", i,
              hb[600 + i * 2] - t0, hb[600 + i * 2 + 1] - t0, hb[i * 8] - t0, hb[i * 8 + 1] - t0, hb[i * 8 + 2] - t0,
              hb[i * 8 + 3] - t0, hb[i * 8 + 4] - t0, hb[i * 8 + 5] - t0, hb[i * 8 + 6] - t0, hb[i * 8 + 7] - t0);
  }
  return IBC_OK;
}

}  // namespace ibc
