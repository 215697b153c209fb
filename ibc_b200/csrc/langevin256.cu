// mcmc.langevin_actions_given_obs (ibc/agents/mcmc.py:330-425) for the width-256, depth-2
// stacks (particle/mlp_ebm_langevin.gin:42-44: C1 / C3) as ONE persistent kernel: a CTA owns
// a 128-row tile and runs all K iterations of forward -> dE/dact -> langevin_step on it.
//
// Round 1 replayed 3 launches per iteration from a CUDA graph (9.3 ms per 100-iteration
// chain on 4096 rows: each launch pays prologue, cold ring and drain for ~10 us of work).
// Two collapses make one iteration two 256 x 256 products plus the two narrow action
// projections (h0 = hp + act . Wp[O:], v = relu(h0) . W1 + b1, h1 = h0 + relu(v) . W2 + b2,
// E = h1 . we + be, networks/layers/resnet.py:135-153 with one block):
//   * E = h0 . we + relu(v) . c + (b2 . we + be) with c = W2 . we: the forward never forms
//     relu(v) . W2;
//   * dE/dh1 = we for every row, so the top reverse product (we . W2^T) is the same constant
//     row c: u = c (.) [v > 0], g0 = we + (u . W1^T) (.) [h0 > 0], dE/dact = g0 . Wp[O:]^T.
// (Both are exact in fp32 where the per-layer path rounds W2 / we to tf32.)
//
// Layout: tensor memory X [0, 256) = A operand, Y [256, 512) = accumulator (N = 256 per
// MMA); 16 epilogue warps (warp e: rows of sub-tile e & 3, columns 64 (e >> 2) ..: four
// warps per scheduler, 64 values per thread per phase); W1 and W1^T streamed every iteration
// from L2 in 64 KB K-slices (K = 64) through a two-slot ring by the TMA engine; the two
// action-projection operands, we, b1, c and the tile's actions stay in shared memory.
#include <stdlib.h>

#include "mlp.cuh"
#include "philox.cuh"
#include "umma.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kW = 256;
constexpr int kSliceK = 64;
constexpr int kSliceBytes = kSliceK * kW * 4;          // 64 KB
constexpr int kSlices = kW / kSliceK;                  // 4 per layer
constexpr int kRing = 2;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 2) * 32;         // + producer, MMA issuer (18 warps: 5 on one scheduler, so <= 96 registers)

struct L256Params {
  const float* hp;          // [B, W]  obs . Wp[:O] + bp
  float* actions;           // [R, A] in / out
  const float* packed;
  const float* cvec;        // c[W] = W2 . we, then b2 . we + be
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
  int A, KP, AP, K;
  int64_t off_fwd_p, off_w1, off_w1t, off_rev_p, off_we, off_b1;
  float noise_scale, grad_clip, dclip_scale;
  int norm_type;
  int* status;
  long long* dbg;   // optional timeline of CTA 0 (IBC_DBG_TIMELINE)
};

inline size_t l256_smem_bytes(int KP, int AP, int AMAX) {
  return (size_t)kRing * kSliceBytes + (size_t)KP * kW * 4 + (size_t)AP * kW * 4 +
         4 * (size_t)kW * 4 + 16 + 2 * 4 * 128 * 4 + (size_t)AMAX * 128 * 4 + (4 + 2 * kRing + 2) * 8 + 64;
}

// c = W2 . we and b2 . we + be (one block of 256 threads; once per chain)
__global__ void __launch_bounds__(kW) l256_prep_kernel(const float* __restrict__ params, Layout L,
                                                       float* __restrict__ out) {
  __shared__ float s_we[kW];
  __shared__ float s_red[kW];
  const int k = threadIdx.x;
  s_we[k] = params[L.we + k];
  __syncthreads();
  const float* w2 = params + L.w2(0) + (int64_t)k * kW;
  float acc = 0.f;
  for (int n = 0; n < kW; ++n) acc = fmaf(w2[n], s_we[n], acc);
  out[k] = acc;
  s_red[k] = params[L.b2(0) + k] * s_we[k];
  __syncthreads();
  if (k == 0) {
    float s = params[L.be];
    for (int n = 0; n < kW; ++n) s += s_red[n];
    out[kW] = s;
  }
}

template <int AMAX>
__global__ void __launch_bounds__(kThreads, 1) langevin256_kernel(L256Params p) {
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;                                              // [kRing][64 KB]
  unsigned char* p0_buf = ring + (size_t)kRing * kSliceBytes;              // [KP/4][W][4]
  unsigned char* pr_buf = p0_buf + (size_t)p.KP * W * 4;                   // [W/4][AP][4]
  float* we = reinterpret_cast<float*>(pr_buf + (size_t)p.AP * W * 4);     // we[W]
  float* b1 = we + W;
  float* cv = b1 + W;                                                      // c
  float* crn = cv + W;                                                     // c rounded to tf32 (operand)
  float* consts = crn + W;                                                 // [0] = b2 . we + be
  float* e_part = consts + 4;                                              // [4][128]
  float* n_part = e_part + 4 * 128;                                        // [4][128] partial gradient norms
  float* acts = n_part + 4 * 128;                                          // [AMAX][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(acts + AMAX * 128);
  uint64_t* a_all = bars;                  // 512 epilogue threads: the A operand is written
  uint64_t* a_act = bars + 1;              // 512 epilogue threads: the next actions are written
  uint64_t* d_ready = bars + 2;
  uint64_t* p_full = bars + 3;             // the two resident projection operands landed
  uint64_t* w_full = bars + 4;             // [kRing]
  uint64_t* w_empty = w_full + kRing;      // [kRing]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kRing);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (p.R + 127) / 128;

  if (tid == 0) {
    s_abort = 0;
    mbar_init(a_all, kEpiWarps * 32);
    mbar_init(a_act, kEpiWarps * 32);
    mbar_init(d_ready, 1);
    mbar_init(p_full, 1);
    for (int s = 0; s < kRing; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < W; i += kThreads) {
    we[i] = p.packed[p.off_we + i];
    b1[i] = p.packed[p.off_b1 + i];
    const float c = p.cvec[i];
    cv[i] = c;
    crn[i] = tf32_rn(c);
  }
  if (tid == 0) consts[0] = p.cvec[W];
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  const uint64_t seed = p.seed_ptr ? *p.seed_ptr : p.seed;

  if (warp < kEpiWarps) {
    // ============ epilogue: thread = (row, 64-column group) ===================
    // The action update is spread over the four column groups too: group cc owns the action
    // dimensions [cc * DPT, (cc + 1) * DPT) of its row -- their normals are drawn into registers
    // while the W1 product runs (one thread per row drawing 32 Philox / Box-Muller normals took
    // 55 k cycles per iteration against 23 k for everything else: first version, timeline).
    constexpr int DPT = AMAX / 4;                          // action dimensions per thread (AMAX = 8 or 32,
                                                           // so that the groups rewrite all KP <= AMAX columns)
    const int wq = warp & 3, cc = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t col_x = tmem_base + lane_off, col_y = col_x + W;
    const int c0 = cc * 64;
    const int a0 = cc * DPT;                                // first action dimension of this thread
    constexpr bool has_dims = true;
    uint32_t ph_d = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int64_t r = tile * 128 + row;
      const bool valid = r < p.R;
      const float* hprow = p.hp + (valid ? (r / p.n_per_obs) : 0) * W + c0;
      {
        // this row's actions into shared memory and, tf32, into X[:, 0:32) (group 0 zero-fills
        // the padding columns once: the update below only rewrites columns < AMAX)
        if (cc == 0) {
          uint32_t z[32];
#pragma unroll
          for (int a = 0; a < 32; ++a) z[a] = 0u;
          tmem_st32(col_x, z);
          tmem_st_wait();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");      // epilogue warps only
        if (has_dims) {
          uint32_t v[DPT];
#pragma unroll
          for (int j = 0; j < DPT; ++j) {
            const int a = a0 + j;
            const float x = (valid && a < p.A) ? p.actions[r * p.A + a] : 0.f;
            acts[a * 128 + row] = x;
            v[j] = tf32_rn_bits(__float_as_uint(x));
          }
          tmem_st_n<DPT>(col_x + (uint32_t)a0, v);
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(a_act);
      }
#pragma unroll 1
      for (int k = 0; k < p.K; ++k) {
        float e = 0.f;
        uint32_t m0[2];
        float nz[DPT];
        const bool stamp = p.dbg && blockIdx.x == 0 && tid == 0 && k < 4 && tile == blockIdx.x;
        long long* dbg = stamp ? p.dbg + k * 8 : nullptr;
        if (stamp) dbg[0] = clock64();
        // ---- E0: h0 = Y + hp; e += h0 . we; X <- relu(h0) ------------------------
        float4 hv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) hv[q] = __ldg(reinterpret_cast<const float4*>(hprow) + q);
        mbar_wait(d_ready, ph_d, abort_flag, p.status, 2000);
        ph_d ^= 1;
        tc_fence_after();
        if (stamp) dbg[1] = clock64();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = c0 + half * 32;
          uint32_t v[32];
          tmem_ld32(col_y + col, v);
          tmem_ld_wait(v);
          uint32_t nbits = 0;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 w4 = *reinterpret_cast<const float4*>(we + col + q4 * 4);
            const float hq[4] = {hv[q4].x, hv[q4].y, hv[q4].z, hv[q4].w};
            const float wq4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int q = q4 * 4 + i;
              const float x = __uint_as_float(v[q]) + hq[i];
              e = fmaf(x, wq4[i], e);
              const uint32_t o = relu_tf32_bits(__float_as_uint(x));
              nbits = __funnelshift_l(o - 0x1001u, nbits, 1);
              v[q] = o;
            }
          }
          m0[half] = ~__brev(nbits);
          tmem_st32(col_x + col, v);
          if (half == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) hv[q] = __ldg(reinterpret_cast<const float4*>(hprow + 32) + q);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(a_all);
        if (stamp) dbg[2] = clock64();
        // this iteration's normals of this thread's action dimensions, while the W1 product runs
        if (has_dims) {
#pragma unroll
          for (int j = 0; j < DPT; ++j) {
            const int a = a0 + j;
            const int64_t i = r * p.A + a;
            nz[j] = (!valid || a >= p.A) ? 0.f
                    : (p.normals ? p.normals[(int64_t)k * p.R * p.A + i]
                                 : philox_normal(seed, (uint64_t)k + 1, (uint64_t)i));
          }
        }

        // ---- E1: v = Y + b1; e += relu(v) . c; X <- c (.) [v > 0] -------------------
        mbar_wait(d_ready, ph_d, abort_flag, p.status, 2001);
        ph_d ^= 1;
        tc_fence_after();
        if (stamp) dbg[3] = clock64();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = c0 + half * 32;
          uint32_t v[32];
          tmem_ld32(col_y + col, v);
          tmem_ld_wait(v);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b4 = *reinterpret_cast<const float4*>(b1 + col + q4 * 4);
            const float4 c4 = *reinterpret_cast<const float4*>(cv + col + q4 * 4);
            const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
            const float cq[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int q = q4 * 4 + i;
              const float x = __uint_as_float(v[q]) + bq[i];
              const bool on = x > 0.f;
              e = fmaf(on ? x : 0.f, cq[i], e);
              v[q] = on ? tf32_rn_bits(__float_as_uint(cq[i])) : 0u;
            }
          }
          tmem_st32(col_x + col, v);
        }
        e_part[cc * 128 + row] = e;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(a_all);
        if (stamp) dbg[4] = clock64();

        // ---- E2: g0 = we + Y (.) [h0 > 0]; X <- tf32(g0) ------------------------------
        mbar_wait(d_ready, ph_d, abort_flag, p.status, 2002);
        ph_d ^= 1;
        tc_fence_after();
        if (stamp) dbg[5] = clock64();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = c0 + half * 32;
          uint32_t v[32];
          tmem_ld32(col_y + col, v);
          tmem_ld_wait(v);
          const uint32_t bits = m0[half];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 w4 = *reinterpret_cast<const float4*>(we + col + q4 * 4);
            const float wq4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int q = q4 * 4 + i;
              const float g = wq4[i] + (((bits >> q) & 1u) ? __uint_as_float(v[q]) : 0.f);
              v[q] = tf32_rn_bits(__float_as_uint(g));
            }
          }
          tmem_st32(col_x + col, v);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(a_all);
        if (stamp) dbg[6] = clock64();

        // ---- E3: dE/dact = Y[:, 0:A), langevin_step on this thread's action dimensions ------
        mbar_wait(d_ready, ph_d, abort_flag, p.status, 2003);
        ph_d ^= 1;
        tc_fence_after();
        if (stamp) dbg[7] = clock64();
        if (cc == 0) {
          const float energy = e_part[row] + e_part[128 + row] + e_part[256 + row] + e_part[384 + row] + consts[0];
          if (p.chain_energies && valid) p.chain_energies[(int64_t)k * p.R + r] = energy;
        }
        float nrm = 0.f;
        if (has_dims) {
          uint32_t v[DPT];
          tmem_ld_n<DPT>(col_y + (uint32_t)a0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float stepsize = p.steps[k];
#pragma unroll
          for (int j = 0; j < DPT; ++j) {
            const int a = a0 + j;
            float nv = 0.f;
            if (a < p.A) {
              // langevin_step (mcmc.py:238-262), as langevin_dact_update_kernel
              const int64_t i = r * p.A + a;
              float g = -__uint_as_float(v[j]);                // de_dact = -dE/dact (mcmc.py:190)
              const float t = fabsf(g);
              if (p.norm_type == IBC_NORM_INF) nrm = fmaxf(nrm, t);
              else if (p.norm_type == IBC_NORM_L2) nrm += t * t;
              else nrm += t;
              if (p.grad_clip > 0.f) g = fminf(fmaxf(g, -p.grad_clip), p.grad_clip);
              g = __fadd_rn(__fmul_rn(0.5f, g), __fmul_rn(nz[j], p.noise_scale));
              float dl = __fmul_rn(stepsize, g);
              const float dc = __fmul_rn(p.dclip_scale, __fadd_rn(p.mx[a], -p.mn[a]));
              dl = fminf(fmaxf(dl, -dc), dc);
              nv = __fadd_rn(acts[a * 128 + row], -dl);
              nv = fminf(fmaxf(nv, p.mn[a]), p.mx[a]);
              acts[a * 128 + row] = nv;
              if (valid) {
                if (p.chain_actions) p.chain_actions[(int64_t)k * p.R * p.A + i] = nv;
                if (k + 1 == p.K) p.actions[i] = nv;
              }
            }
            v[j] = tf32_rn_bits(__float_as_uint(nv));
          }
          if (k + 1 < p.K) {
            tmem_st_n<DPT>(col_x + (uint32_t)a0, v);
            tmem_st_wait();
          }
        }
        if (p.chain_grad_norms) {
          // the row's norm from the four groups' partial norms (the next E0 is ordered behind
          // this through the a_act / d_ready barriers, so n_part is free again by then)
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
        tc_fence_before();
        if (k + 1 < p.K) mbar_arrive(a_act);
      }
    }
  } else if (warp == kEpiWarps) {
    // ===================== producer ==============================================
    if (elect_one()) {
      const uint32_t b0 = (uint32_t)(p.KP * W * 4), b1b = (uint32_t)(p.AP * W * 4);
      mbar_arrive_expect_tx(p_full, b0 + b1b);
      bulk_g2s(p0_buf, p.packed + p.off_fwd_p, b0, p_full);
      bulk_g2s(pr_buf, p.packed + p.off_rev_p, b1b, p_full);
    }
    __syncwarp();
    uint32_t q = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int k = 0; k < p.K; ++k) {
        for (int l = 0; l < 2 * kSlices; ++l, ++q) {       // W1 slices, then W1^T slices
          const int slot = q % kRing;
          const uint32_t ph = (q / kRing) & 1;
          const float* src = p.packed + (l < kSlices ? p.off_w1 : p.off_w1t) +
                             (int64_t)(l % kSlices) * (kSliceBytes / 4);
          if (elect_one()) {
            mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 2100);
            mbar_arrive_expect_tx(&w_full[slot], (uint32_t)kSliceBytes);
            bulk_g2s(ring + (size_t)slot * kSliceBytes, src, (uint32_t)kSliceBytes, &w_full[slot]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer ==============================================
    const uint32_t idesc = idesc_tf32(128, W);
    const uint32_t idesc_p = idesc_tf32(128, p.AP);
    const uint32_t b_lbo = W * 16, sbo = 128;
    constexpr uint32_t b_step16 = (2 * W * 16) >> 4;
    const uint32_t p_lbo = (uint32_t)p.AP * 16, p_step16 = (2u * (uint32_t)p.AP * 16u) >> 4;
    const uint32_t col_x = tmem_base, col_y = tmem_base + W;
    uint32_t q = 0, ph_all = 0, ph_act = 0;
    bool first = true;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int k = 0; k < p.K; ++k) {
        // S0: Y = act . Wp[O:]
        if (elect_one()) {
          mbar_wait(a_act, ph_act, abort_flag, p.status, 2200);
          if (first) mbar_wait(p_full, 0, abort_flag, p.status, 2201);
          tc_fence_after();
          const uint64_t bd = smem_desc(smem_u32(p0_buf), b_lbo, sbo);
          for (int ks = 0; ks < p.KP / 8; ++ks)
            mma_tf32_ts(col_y, col_x + (uint32_t)(ks * 8), bd + (uint64_t)(ks * b_step16), idesc,
                        (uint32_t)(ks > 0));
          mma_commit(d_ready);
        }
        __syncwarp();
        first = false;
        ph_act ^= 1;
        // S1: Y = relu(h0) . W1 ; S2: Y = u . W1^T   (K = 256 in four streamed slices)
        for (int layer = 0; layer < 2; ++layer) {
          if (elect_one()) mbar_wait(a_all, ph_all, abort_flag, p.status, 2210 + layer);
          __syncwarp();
          ph_all ^= 1;
          for (int sl = 0; sl < kSlices; ++sl, ++q) {
            const uint32_t slot = q % kRing;
            if (elect_one()) {
              const bool ist = p.dbg && blockIdx.x == 0 && q < 32;
              if (ist) p.dbg[64 + q * 3] = clock64();
              mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 2220 + layer);
              if (ist) p.dbg[64 + q * 3 + 1] = clock64();
              tc_fence_after();
              const uint64_t bd = smem_desc(smem_u32(ring + (size_t)slot * kSliceBytes), b_lbo, sbo);
#pragma unroll
              for (int ks = 0; ks < kSliceK / 8; ++ks)
                mma_tf32_ts(col_y, col_x + (uint32_t)(sl * kSliceK + ks * 8),
                            bd + (uint64_t)(ks * b_step16), idesc, (uint32_t)(sl > 0 || ks > 0));
              mma_commit(&w_empty[slot]);
              if (sl == kSlices - 1) mma_commit(d_ready);
              if (ist) p.dbg[64 + q * 3 + 2] = clock64();
            }
            __syncwarp();
          }
        }
        // S3: Y[:, 0:AP) = g0 . Wp[O:]^T
        if (elect_one()) {
          mbar_wait(a_all, ph_all, abort_flag, p.status, 2230);
          tc_fence_after();
          const uint64_t bd = smem_desc(smem_u32(pr_buf), p_lbo, sbo);
          for (int ks = 0; ks < W / 8; ++ks)
            mma_tf32_ts(col_y, col_x + (uint32_t)(ks * 8), bd + (uint64_t)(ks * p_step16), idesc_p,
                        (uint32_t)(ks > 0));
          mma_commit(d_ready);
        }
        __syncwarp();
        ph_all ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int AMAX>
int launch_l256(ibc_handle* h, const L256Params& p, cudaStream_t s) {
  const size_t smem = l256_smem_bytes(p.KP, p.AP, AMAX);
  IBC_RAISE_SMEM(h, langevin256_kernel<AMAX>, 227 * 1024 - 1024);
  const int64_t tiles = (p.R + 127) / 128;
  const int grid = tiles < h->sm_count ? (int)tiles : h->sm_count;
  L256Params q = p;
  q.dbg = nullptr;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&q.dbg, 512 * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(q.dbg, 0, 512 * sizeof(long long), s));
  }
  langevin256_kernel<AMAX><<<grid, kThreads, smem, s>>>(q);
  IBC_LAUNCH_CHECK(h);
  if (q.dbg) {
    cudaStreamSynchronize(s);
    long long hb[512];
    cudaMemcpy(hb, q.dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    cudaFree(q.dbg);
    const long long t0 = hb[0];
    fprintf(stderr, "[ibc timeline l256] iter | start S0done E0done | S1done E1done | S2done E2done | S3done\n");
    for (int k = 0; k < 4; ++k) {
      const long long* e = &hb[k * 8];
      fprintf(stderr, "[ibc timeline l256] %d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld | %7lld\n", k, e[0] - t0,
              e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0, e[6] - t0, e[7] - t0);
    }
    for (int q2 = 0; q2 < 24; ++q2)
      fprintf(stderr, "[ibc timeline l256 slice] %2d | wait_start %7lld got %7lld issued %7lld\n", q2,
              hb[64 + q2 * 3] - t0, hb[64 + q2 * 3 + 1] - t0, hb[64 + q2 * 3 + 2] - t0);
  }
  return IBC_OK;
}

}  // namespace

bool langevin256_supported(const Layout& L) {
  if (L.W != kW || L.nb != 1 || L.A > 32) return false;
  const int amax = L.A <= 8 ? 8 : 32;
  return l256_smem_bytes(round_up(L.A, 8), round_up(L.A, 16), amax) <= 227 * 1024 - 1024;
}

// hp: [B, W] observation projection; steps_dev: [K] step sizes; cvec: W + 1 floats of scratch
// (filled here).  apply_exp is not handled (callers keep the three-kernel chain for it).
int langevin256_fused(ibc_handle* h, const float* hp, int64_t B, float* actions, int64_t n,
                      const float* mn, const float* mx, const ibc_langevin_cfg& c,
                      const float* normals, uint64_t seed, const uint64_t* seed_ptr,
                      const float* steps_dev, float* cvec, float* chain_actions,
                      float* chain_energies, float* chain_grad_norms, cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  l256_prep_kernel<<<1, kW, 0, s>>>(h->params, L, cvec);
  IBC_LAUNCH_CHECK(h);
  L256Params p;
  p.hp = hp; p.actions = actions; p.packed = h->packed; p.cvec = cvec;
  p.mn = mn; p.mx = mx; p.normals = normals; p.steps = steps_dev; p.seed_ptr = seed_ptr;
  p.seed = seed; p.chain_actions = chain_actions; p.chain_energies = chain_energies;
  p.chain_grad_norms = chain_grad_norms;
  p.R = B * n; p.n_per_obs = n; p.A = L.A; p.KP = pl.KP; p.AP = pl.AP; p.K = c.num_iterations;
  p.off_fwd_p = pl.fwd_p;
  p.off_w1 = pl.fwd_l;                                  // W1_0
  p.off_w1t = pl.rev_l + (int64_t)L.W * L.W;            // rev_l: W2_0^T, W1_0^T
  p.off_rev_p = pl.rev_p;
  p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.off_b1 = pl.raw + (int64_t)L.W;                     // raw: zeros, b1_0, b2_0
  p.noise_scale = c.noise_scale; p.grad_clip = c.grad_clip;
  p.dclip_scale = (float)((double)c.delta_action_clip * 0.5);
  p.norm_type = c.grad_norm_type;
  p.status = h->dev_status;
  if (L.A <= 8) return launch_l256<8>(h, p, s);
  return launch_l256<32>(h, p, s);
}

}  // namespace ibc
