// mcmc.langevin_actions_given_obs (ibc/agents/mcmc.py:330-425) for width-128 stacks as
// ONE persistent kernel: a CTA owns a 128-row tile and runs all K iterations of
//   forward (resnet.py:135-153) -> dE/dact (mcmc.py:161-191) -> langevin_step (mcmc.py:209-264)
// on it without leaving the SM.  Rows are independent, so there is no grid-wide step.
//
// Why: the chain is latency bound.  With one launch per forward / reverse / update (even
// replayed from a CUDA graph) an iteration costs ~80 us at 4096 rows, of which the
// dependent layers of a lone tile are ~40 us; the rest is 300 kernel prologues (TMEM
// allocation, barrier and table set-up, cold weight ring) and drains per chain.
//
// Same machinery as mlp_umma_tmem.cu: activations in tensor memory (two 128-column
// regions X / Y alternating as A operand and accumulator, rewritten in place by the
// epilogue), the residual row h (forward) or its gradient g (reverse) in 128 registers per
// thread, weights streamed a layer at a time (forward: operand + bias slice, reverse:
// transposed operand) through a two-slot shared-memory ring, bias by an extra K = 8 MMA.
// The ReLU masks of the 2nb layers stay in shared memory (16 B per row and layer), and
// dE/dact (g0 . Wp[O:]^T, act_dim <= 32 columns), the noise, both clips and the action
// update are done by the row's own epilogue thread.  Arithmetic and its order are those of
// the three-kernel path, so the result is bit-identical to it (tested).
#include <stdlib.h>

#include <type_traits>

#include "mlp.cuh"
#include "philox.cuh"
#include "umma.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kW = 128;
constexpr int kLayerWBytes = kW * kW * 4;
constexpr int kBiasBytes = 8 * kW * 4;
constexpr int kSlotBytes = kLayerWBytes + kBiasBytes;      // 68 KB (reverse layers use 64)
constexpr int kRing = 2;
constexpr int kKSteps = kW / 8;
constexpr int kThreads = 6 * 32;                           // 4 epilogue warps, producer, issuer

struct LangevinParams {
  const float* hp;          // [B, W]  obs . Wp[:O] + bp
  float* actions;           // [R, A] in / out
  const float* packed;
  const float* wpa;         // Wp[O:, :]  [A, W] fp32 (dE/dact)
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
  int nb, A, KP, K;
  int64_t off_fwd_p, off_fwd_lb, off_rev_l, off_we;
  float noise_scale, grad_clip, dclip_scale;
  int norm_type, apply_exp;
  int* status;
};

inline size_t langevin_smem_bytes(int nb, int KP, int AMAX) {
  return (size_t)kRing * kSlotBytes + (size_t)KP * kW * 4 + 4096 + (kW + 4) * 4 +
         (size_t)AMAX * kW * 4 + (size_t)2 * nb * 128 * 16 + (2 + 2 * kRing + 1) * 8 + 64;
}

template <int AMAX>
__global__ void __launch_bounds__(kThreads, 1) langevin_tmem_kernel(LangevinParams p) {
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nsteps = 2 * p.nb + 1;            // forward steps; the reverse chain has 2nb
  unsigned char* ring = smem;                                            // [kRing][68 KB]
  unsigned char* p0_buf = ring + (size_t)kRing * kSlotBytes;             // [KP/4][W][4]
  float4* ones_tile = reinterpret_cast<float4*>(p0_buf + (size_t)p.KP * W * 4);
  float* we = reinterpret_cast<float*>(ones_tile + 256);                 // we[W], be
  float* wpa = we + W + 4;                                               // [AMAX][W]
  uint4* mask_s = reinterpret_cast<uint4*>(wpa + AMAX * W);              // [2nb][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(mask_s + 2 * p.nb * 128);
  uint64_t* a_ready = bars;
  uint64_t* d_ready = bars + 1;
  uint64_t* w_full = bars + 2;             // [kRing]
  uint64_t* w_empty = w_full + kRing;      // [kRing]
  uint64_t* p0_full = w_empty + kRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p0_full + 1);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t tiles = (p.R + 127) / 128;

  if (tid == 0) {
    s_abort = 0;
    mbar_init(a_ready, 128);
    mbar_init(d_ready, 1);
    for (int s = 0; s < kRing; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(p0_full, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 256; i += kThreads)
    ones_tile[i] = (i < 128) ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < W + 4; i += kThreads) we[i] = p.packed[p.off_we + i];
  for (int i = tid; i < AMAX * W; i += kThreads) wpa[i] = (i / W < p.A) ? p.wpa[i] : 0.f;
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  const float be = we[W];
  const uint64_t seed = p.seed_ptr ? *p.seed_ptr : p.seed;

  if (warp < 4) {
    // ===================== epilogue: thread = row =============================
    const int row = tid;
    const uint32_t col_x = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t col_y = col_x + W;
    uint32_t ph_d = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int64_t r = tile * 128 + row;
      const bool valid = r < p.R;
      float act[AMAX];
#pragma unroll
      for (int a = 0; a < AMAX; ++a) act[a] = (valid && a < p.A) ? p.actions[r * p.A + a] : 0.f;
      const float4* hp4 = reinterpret_cast<const float4*>(
          p.hp + (valid ? (r / p.n_per_obs) : 0) * W);
#pragma unroll 1
      for (int k = 0; k < p.K; ++k) {
        // ---- step-0 operand: the action (zero padded) into X[:, 0:32) ----
        {
          uint32_t v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q)
            v[q] = (q < AMAX && q < p.A) ? tf32_rn_bits(__float_as_uint(act[q < AMAX ? q : 0])) : 0u;
          if (!valid) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0u;
          }
          tmem_st32(col_x, v);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(a_ready);
        }
        float h[W];          // residual row (forward), then its gradient (reverse)
#pragma unroll
        for (int c = 0; c < W / 4; ++c) {
          const float4 v = hp4[c];
          h[c * 4 + 0] = v.x; h[c * 4 + 1] = v.y; h[c * 4 + 2] = v.z; h[c * 4 + 3] = v.w;
        }
        float energy = 0.f;
        // ---- forward steps (see mlp_fwd_tmem_kernel) ----
        auto fstep = [&](auto even_tag, auto last_tag, int i) {
          constexpr bool kEven = decltype(even_tag)::value;
          constexpr bool kLast = decltype(last_tag)::value;
          mbar_wait(d_ready, ph_d, abort_flag, p.status, 1100 + i);
          ph_d ^= 1;
          tc_fence_after();
          const uint32_t col_d = kEven ? ((i == 0) ? col_y : col_x) : col_y;
          const uint32_t col_a = kEven ? col_x : col_y;
          uint32_t mb[4];
          float e = 0.f;
          auto chunk = [&](int c32, uint32_t (&v)[32]) {
            uint32_t nbits = 0;
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
                const uint32_t o = relu_tf32_bits(xb);
                nbits = __funnelshift_l(o - 0x1001u, nbits, 1);
                xb = o;
              }
              v[q] = xb;
            }
            if (!kLast) {
              mb[c32] = ~__brev(nbits);
              tmem_st32(col_a + c32 * 32, v);
            }
          };
          uint32_t va[32], vb[32];
          tmem_ld32(col_d, va);
          tmem_ld_wait(va); tmem_ld32(col_d + 32, vb); chunk(0, va);
          tmem_ld_wait(vb); tmem_ld32(col_d + 64, va); chunk(1, vb);
          tmem_ld_wait(va); tmem_ld32(col_d + 96, vb); chunk(2, va);
          tmem_ld_wait(vb); chunk(3, vb);
          if (kLast) {
            energy = e + be;                       // Dense(1), mlp_ebm.py:91-94
          } else {
            mask_s[i * 128 + row] = make_uint4(mb[0], mb[1], mb[2], mb[3]);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a_ready);
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
        if (p.chain_energies && valid) p.chain_energies[(int64_t)k * p.R + r] = energy;

        // ---- reverse chain (see mlp_bwd_tmem_kernel): g_top = dE (x) we ----
        const float de = valid ? (p.apply_exp ? expf(energy) : 1.0f) : 0.f;
        if (p.apply_exp && p.chain_energies && valid)
          p.chain_energies[(int64_t)k * p.R + r] = de;       // the three-kernel path stores exp(E)
#pragma unroll
        for (int c32 = 0; c32 < W / 32; ++c32) {
          uint32_t v[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const float g = de * we[c32 * 32 + q];
            h[c32 * 32 + q] = g;
            v[q] = tf32_rn_bits(__float_as_uint(g));
          }
          tmem_st32(col_x + c32 * 32, v);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(a_ready);
        auto rstep = [&](auto odd_tag, auto last_tag, int i) {
          constexpr bool kOdd = decltype(odd_tag)::value;
          constexpr bool kLast = decltype(last_tag)::value;
          const uint4 m4 = mask_s[(i - 1) * 128 + row];
          const uint32_t mbits[4] = {m4.x, m4.y, m4.z, m4.w};
          mbar_wait(d_ready, ph_d, abort_flag, p.status, 1200 + i);
          ph_d ^= 1;
          tc_fence_after();
          const uint32_t col_d = kOdd ? col_x : col_y;
          auto chunk = [&](int c32, uint32_t (&v)[32]) {
            const uint32_t bits = mbits[c32];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              float o = (bits >> q) & 1u ? __uint_as_float(v[q]) : 0.f;
              if (kOdd) {
                o += h[c32 * 32 + q];
                h[c32 * 32 + q] = o;
              }
              v[q] = tf32_rn_bits(__float_as_uint(o));
            }
            if (!kLast) tmem_st32(col_d + c32 * 32, v);
          };
          uint32_t va[32], vb[32];
          tmem_ld32(col_d, va);
          tmem_ld_wait(va); tmem_ld32(col_d + 32, vb); chunk(0, va);
          tmem_ld_wait(vb); tmem_ld32(col_d + 64, va); chunk(1, vb);
          tmem_ld_wait(va); tmem_ld32(col_d + 96, vb); chunk(2, va);
          tmem_ld_wait(vb); chunk(3, vb);
          if (!kLast) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a_ready);
          } else {
            tc_fence_before();
          }
        };
#pragma unroll 1
        for (int j = p.nb - 1; j >= 1; --j) {
          rstep(F_{}, F_{}, 2 * j + 2);
          rstep(T_{}, F_{}, 2 * j + 1);
        }
        rstep(F_{}, F_{}, 2);
        rstep(T_{}, T_{}, 1);

        // ---- dE/dact = g0 . Wp[O:]^T, then langevin_step (langevin_dact_update_kernel) ----
        float gact[AMAX];
        float nrm = 0.f;
#pragma unroll
        for (int a = 0; a < AMAX; ++a) {
          float d = 0.f;
          const float4* w4 = reinterpret_cast<const float4*>(wpa + a * W);
#pragma unroll
          for (int c = 0; c < W / 4; ++c) {
            const float4 w = w4[c];
            d = fmaf(h[c * 4 + 0], w.x, d); d = fmaf(h[c * 4 + 1], w.y, d);
            d = fmaf(h[c * 4 + 2], w.z, d); d = fmaf(h[c * 4 + 3], w.w, d);
          }
          const float g = -d;                          // de_dact = -dE/dact (mcmc.py:190)
          gact[a] = g;
          if (a < p.A) {
            const float t = fabsf(g);
            if (p.norm_type == IBC_NORM_INF) nrm = fmaxf(nrm, t);
            else if (p.norm_type == IBC_NORM_L2) nrm += t * t;
            else nrm += t;
          }
        }
        if (p.norm_type == IBC_NORM_L2) nrm = sqrtf(nrm);
        if (p.norm_type == IBC_NORM_NONE) nrm = 0.f;
        if (p.chain_grad_norms && valid) p.chain_grad_norms[(int64_t)k * p.R + r] = nrm;
        const float stepsize = p.steps[k];
#pragma unroll
        for (int a = 0; a < AMAX; ++a) {
          if (a < p.A && valid) {
            const int64_t i = r * p.A + a;
            float g = gact[a];
            if (p.grad_clip > 0.f) g = fminf(fmaxf(g, -p.grad_clip), p.grad_clip);
            const float z = p.normals ? p.normals[(int64_t)k * p.R * p.A + i]
                                      : philox_normal(seed, (uint64_t)k + 1, (uint64_t)i);
            g = __fadd_rn(__fmul_rn(0.5f, g), __fmul_rn(z, p.noise_scale));
            float dl = __fmul_rn(stepsize, g);
            const float dc = __fmul_rn(p.dclip_scale, __fadd_rn(p.mx[a], -p.mn[a]));
            dl = fminf(fmaxf(dl, -dc), dc);
            float v = __fadd_rn(act[a], -dl);
            v = fminf(fmaxf(v, p.mn[a]), p.mx[a]);
            act[a] = v;
            if (p.chain_actions) p.chain_actions[(int64_t)k * p.R * p.A + i] = v;
          }
        }
      }
#pragma unroll
      for (int a = 0; a < AMAX; ++a)
        if (a < p.A && valid) p.actions[r * p.A + a] = act[a];
    }
  } else if (warp == 4) {
    // ===================== producer: one bulk copy per layer ==================
    if (elect_one()) {
      mbar_arrive_expect_tx(p0_full, (uint32_t)(W * p.KP * 4));
      bulk_g2s(p0_buf, p.packed + p.off_fwd_p, (uint32_t)(W * p.KP * 4), p0_full);
    }
    __syncwarp();
    uint32_t q = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int k = 0; k < p.K; ++k) {
        for (int l = 0; l < 4 * p.nb; ++l, ++q) {       // 2nb forward layers, 2nb reverse layers
          const int slot = q % kRing;
          const uint32_t ph = (q / kRing) & 1;
          const bool fwd = l < 2 * p.nb;
          const float* src = fwd ? p.packed + p.off_fwd_lb + (int64_t)l * (kSlotBytes / 4)
                                 : p.packed + p.off_rev_l + (int64_t)(l - 2 * p.nb) * W * W;
          const uint32_t bytes = fwd ? (uint32_t)kSlotBytes : (uint32_t)kLayerWBytes;
          if (elect_one()) {
            mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 1300);
            mbar_arrive_expect_tx(&w_full[slot], bytes);
            bulk_g2s(ring + (size_t)slot * kSlotBytes, src, bytes, &w_full[slot]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== MMA issuer ==========================================
    const uint32_t idesc = idesc_tf32(128, W);
    const uint32_t a_lbo = 128 * 16, b_lbo = W * 16, sbo = 128;
    constexpr uint32_t b_step16 = (2 * W * 16) >> 4;
    const uint64_t ones_desc = smem_desc(smem_u32(ones_tile), a_lbo, sbo);
    const uint32_t col_x = tmem_base, col_y = tmem_base + W;
    uint32_t q = 0, ph_a = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int k = 0; k < p.K; ++k) {
        // forward: step 0 and odd steps read X, write Y; even steps >= 2 read Y, write X
        for (int i = 0; i < nsteps; ++i) {
          const bool x_to_y = (i == 0) || (i & 1);
          const uint32_t a_tmem = x_to_y ? col_x : col_y;
          const uint32_t d_tmem = x_to_y ? col_y : col_x;
          const uint32_t slot = q % kRing;
          if (elect_one()) {
            // weights first: a passed wait still costs 150-300 cycles (langevin_f16.cu)
            if (i == 0) mbar_wait(p0_full, 0, abort_flag, p.status, 1450);
            else mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 1460);
            mbar_wait(a_ready, ph_a, abort_flag, p.status, 1400 + i);
            tc_fence_after();
            if (i == 0) {
              const uint64_t b_desc0 = smem_desc(smem_u32(p0_buf), b_lbo, sbo);
              for (int ks = 0; ks < p.KP / 8; ++ks)
                mma_tf32_ts(d_tmem, a_tmem + (uint32_t)(ks * 8), b_desc0 + (uint64_t)(ks * b_step16),
                            idesc, (uint32_t)(ks > 0));
            } else {
              const uint32_t layer = smem_u32(ring + (size_t)slot * kSlotBytes);
              mma_tf32_ts_steps<kKSteps>(d_tmem, a_tmem, smem_desc(layer, b_lbo, sbo), b_step16,
                                         idesc, 0u);
              mma_tf32_ss(d_tmem, ones_desc, smem_desc(layer + kLayerWBytes, b_lbo, sbo), idesc, 1u);
              mma_commit(&w_empty[slot]);
            }
            mma_commit(d_ready);
          }
          __syncwarp();
          ph_a ^= 1;
          if (i > 0) ++q;
        }
        // reverse: even layers read X (g), write Y; odd layers read Y (u), write X
        for (int s = 0; s < 2 * p.nb; ++s, ++q) {
          const int i = 2 * p.nb - s;
          const uint32_t a_tmem = (i & 1) ? col_y : col_x;
          const uint32_t d_tmem = (i & 1) ? col_x : col_y;
          const uint32_t slot = q % kRing;
          if (elect_one()) {
            mbar_wait(&w_full[slot], (q / kRing) & 1, abort_flag, p.status, 1560);
            mbar_wait(a_ready, ph_a, abort_flag, p.status, 1500 + s);
            tc_fence_after();
            mma_tf32_ts_steps<kKSteps>(
                d_tmem, a_tmem, smem_desc(smem_u32(ring + (size_t)slot * kSlotBytes), b_lbo, sbo),
                b_step16, idesc, 0u);
            mma_commit(&w_empty[slot]);
            mma_commit(d_ready);
          }
          __syncwarp();
          ph_a ^= 1;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

template <int AMAX>
int launch_langevin(ibc_handle* h, const LangevinParams& p, cudaStream_t s) {
  const size_t smem = langevin_smem_bytes(p.nb, p.KP, AMAX);
  IBC_RAISE_SMEM(h, langevin_tmem_kernel<AMAX>, 227 * 1024 - 1024);
  const int64_t tiles = (p.R + 127) / 128;
  const int grid = tiles < h->sm_count ? (int)tiles : h->sm_count;
  langevin_tmem_kernel<AMAX><<<grid, kThreads, smem, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

}  // namespace

bool langevin_fused_supported(const Layout& L) {
  if (L.W != kW || L.A > 32 || L.nb < 1) return false;
  const int amax = L.A <= 2 ? 2 : (L.A <= 8 ? 8 : 32);
  return langevin_smem_bytes(L.nb, round_up(L.A, 8), amax) <= 227 * 1024 - 1024;
}

// hp: [B, W] observation projection; steps_dev: [K] step sizes on the device.
int langevin_fused(ibc_handle* h, const float* hp, int64_t B, float* actions, int64_t n,
                   const float* mn, const float* mx, const ibc_langevin_cfg& c,
                   const float* normals, uint64_t seed, const uint64_t* seed_ptr,
                   const float* steps_dev, float* chain_actions, float* chain_energies,
                   float* chain_grad_norms, cudaStream_t s) {
  const Layout& L = h->L;
  const PackedLayout pl = make_packed_layout(L);
  LangevinParams p;
  p.hp = hp; p.actions = actions; p.packed = h->packed;
  p.wpa = h->params + L.wp + (int64_t)L.O * L.W;
  p.mn = mn; p.mx = mx; p.normals = normals; p.steps = steps_dev; p.seed_ptr = seed_ptr;
  p.seed = seed; p.chain_actions = chain_actions; p.chain_energies = chain_energies;
  p.chain_grad_norms = chain_grad_norms;
  p.R = B * n; p.n_per_obs = n; p.nb = L.nb; p.A = L.A; p.KP = pl.KP; p.K = c.num_iterations;
  p.off_fwd_p = pl.fwd_p; p.off_fwd_lb = pl.fwd_lb; p.off_rev_l = pl.rev_l;
  p.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  p.noise_scale = c.noise_scale; p.grad_clip = c.grad_clip;
  p.dclip_scale = (float)((double)c.delta_action_clip * 0.5);
  p.norm_type = c.grad_norm_type; p.apply_exp = c.apply_exp;
  p.status = h->dev_status;
  if (L.A <= 2) return launch_langevin<2>(h, p, s);
  if (L.A <= 8) return launch_langevin<8>(h, p, s);
  return launch_langevin<32>(h, p, s);
}

}  // namespace ibc
