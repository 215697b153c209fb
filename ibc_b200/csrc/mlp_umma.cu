// Layer-fused MLPEBM energy kernels on tcgen05 (kind::tf32, fp32 accumulate in
// tensor memory) -- the IBC_PRECISION_TF32 path of
//   networks/mlp_ebm.py:72-96 + networks/layers/resnet.py:135-153 (forward)
//   ibc/agents/mcmc.py:161-191 (dE/dact, reverse chain)
//
// One persistent CTA per SM walks groups of NT row tiles (128 rows each).  The
// residual stream h of a tile never leaves tensor memory: the second Dense of
// every block accumulates straight onto it (D_h += relu(v) . W2), and the
// biases are applied when the epilogue reads it (cumulative-bias vectors
// precomputed at pack time).  Weights are pre-packed into the UMMA operand
// layout (umma.cuh) and streamed from L2 in 16 KB K-slices through a
// shared-memory ring by the TMA engine (cp.async.bulk + mbarrier tx-count).
//
// Warp roles: 4*NT epilogue warps (warpgroup t owns tile t: TMEM -> registers
// -> bias/ReLU -> next A operand in shared memory), 1 producer warp (bulk
// copies), 1 MMA warp (single elected thread issues tcgen05.mma / commit).
// With NT = 2 the MMA of one tile overlaps the epilogue of the other and each
// weight slice is read once per 256 rows.
#include <stdlib.h>

#include <algorithm>
#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "ops.cuh"
#include "umma.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kSlots = 5;
constexpr int kSliceBytes = 16384;

// Elements [lo, hi) of the packed buffer, without [hole_lo, hole_lo + hole_len).
__global__ void pack_kernel(const float* __restrict__ P, Layout L, PackedLayout pl,
                            float* __restrict__ out, int64_t lo, int64_t hi, int64_t hole_lo,
                            int64_t hole_len) {
  const int W = L.W;
  const int64_t WW = (int64_t)W * W;
  for (int64_t j = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < hi - hole_len;
       j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j < hole_lo ? j : j + hole_len;
    float v = 0.f;
    if (i < pl.fwd_l) {                    // step-0 operand: [KP/4][W][4]
      const int64_t e = i - pl.fwd_p;
      const int kq = (int)(e & 3), n = (int)((e >> 2) % W), kc = (int)((e >> 2) / W);
      const int k = kc * 4 + kq;
      v = (k < L.A) ? P[L.wp + (int64_t)(L.O + k) * W + n] : 0.f;
    } else if (i < pl.rev_l) {             // forward W x W operands
      const int64_t e = i - pl.fwd_l;
      const int m = (int)(e / WW);         // 0: W1_0, 1: W2_0, ...
      const int64_t r = e - (int64_t)m * WW;
      const int kq = (int)(r & 3), n = (int)((r >> 2) % W), kc = (int)((r >> 2) / W);
      const int k = kc * 4 + kq;
      const int j = m >> 1;
      const int64_t base = (m & 1) ? L.w2(j) : L.w1(j);
      v = P[base + (int64_t)k * W + n];
    } else if (i < pl.rev_p) {             // reverse operands (transposed use)
      const int64_t e = i - pl.rev_l;
      const int m = (int)(e / WW);         // 0: W2_{nb-1}, 1: W1_{nb-1}, 2: W2_{nb-2}, ...
      const int64_t r = e - (int64_t)m * WW;
      const int oq = (int)(r & 3), irow = (int)((r >> 2) % W), oc = (int)((r >> 2) / W);
      const int o = oc * 4 + oq;
      const int j = L.nb - 1 - (m >> 1);
      const int64_t base = (m & 1) ? L.w1(j) : L.w2(j);
      v = P[base + (int64_t)irow * W + o];
    } else if (i < pl.vecs) {              // dE/dact operand: [W/4][AP][4]
      const int64_t e = i - pl.rev_p;
      const int wq = (int)(e & 3), a = (int)((e >> 2) % pl.AP), wc = (int)((e >> 2) / pl.AP);
      const int w = wc * 4 + wq;
      v = (a < L.A) ? P[L.wp + (int64_t)(L.O + a) * W + w] : 0.f;
    } else if (i >= pl.fwd_lb) {
      // layer block m: the forward operand of layer m (as in fwd_l) + the bias slice of
      // step m + 1 (as in fwd_bias)
      const int64_t e = i - pl.fwd_lb;
      const int64_t blk = WW + 8 * W;
      const int m = (int)(e / blk);
      const int64_t r = e - (int64_t)m * blk;
      const int j = m >> 1;
      if (r < WW) {
        const int kq = (int)(r & 3), n = (int)((r >> 2) % W), kc = (int)((r >> 2) / W);
        const int k = kc * 4 + kq;
        const int64_t base = (m & 1) ? L.w2(j) : L.w1(j);
        v = tf32_rn(P[base + (int64_t)k * W + n]);
      } else {
        const int rb = (int)(r - WW);
        const int kc = rb / (4 * W), n = (rb / 4) % W, kq = rb & 3;
        const float b = (m & 1) ? P[L.b2(j) + n] : P[L.b1(j) + n];
        const float hi = tf32_rn(b);
        v = (kc == 0 && kq == 0) ? hi : ((kc == 0 && kq == 1) ? tf32_rn(b - hi) : 0.f);
      }
      out[i] = v;
      continue;
    } else if (i >= pl.fwd_bias) {
      // bias of step s as a K = 8 operand slice [2 chunks][W][4]: row n of chunk 0 holds
      // (hi, lo, 0, 0) with hi + lo = bias to ~2^-21; the matching A operand is a
      // constant (1, 1, 0, 0 | 0) tile, so the MMA adds the bias in fp32
      const int64_t e = i - pl.fwd_bias;
      const int s = (int)(e / (8 * W));
      const int r = (int)(e % (8 * W));
      const int kc = r / (4 * W), n = (r / 4) % W, kq = r & 3;
      float b = 0.f;
      if (s > 0) b = (s & 1) ? P[L.b1((s - 1) >> 1) + n] : P[L.b2((s >> 1) - 1) + n];
      const float hi = tf32_rn(b);
      v = (kc == 0 && kq == 0) ? hi : ((kc == 0 && kq == 1) ? tf32_rn(b - hi) : 0.f);
    } else if (i >= pl.raw) {              // raw per-step biases (residual kept in registers)
      const int64_t e = i - pl.raw;
      const int s = (int)(e / W), w = (int)(e % W);
      if (s == 0) v = 0.f;
      else if (s & 1) v = P[L.b1((s - 1) >> 1) + w];
      else v = P[L.b2((s >> 1) - 1) + w];
    } else {
      const int64_t e = i - pl.vecs;
      const int s = (int)(e / W), w = (int)(e % W);
      if (s == 0) {
        v = 0.f;                            // bp is folded into the hoisted obs projection
      } else if (s <= 2 * L.nb) {
        if (s & 1) {
          v = P[L.b1((s - 1) >> 1) + w];
        } else {                            // cumulative b2 of blocks 0 .. s/2-1
          float c = 0.f;
          for (int j = 0; j < (s >> 1); ++j) c += P[L.b2(j) + w];
          v = c;
        }
      } else if (s == 2 * L.nb + 1) {
        v = P[L.we + w];
      } else {
        v = (w == 0) ? P[L.be] : 0.f;
      }
    }
    out[i] = (i < pl.vecs) ? tf32_rn(v) : v;   // MMA operands rounded to nearest tf32
  }
}

struct FwdParams {
  const float* act;      // [R, A]
  const float* hp;       // [B, W]
  const float* packed;
  float* E;              // [R]
  int64_t R, n_per_obs;
  int nb, A, KP;
  int64_t off_fwd_p, off_fwd_l, off_vecs, off_fwd_bias;
  int* status;
  int num_groups;
  // Training: when non-null the A operand of every W x W layer (slots
  // 0..2nb-1, tf32-rounded) and the final residual stream h_nb (slot 2nb,
  // fp32) are also written to global memory in the tile layout of
  // train_tile_offset() for the backward kernels.
  float* xa_save;
  uint32_t* mask_save;   // ReLU masks of the 2nb W x W layers (TrainTile::mask_index)
  long long* dbg;   // optional timeline of CTA 0, group 0: [step][tile][5] clock64 stamps
};

template <int W, int NT>
struct FwdSmem {
  static constexpr int kABytes = 128 * W * 4;
  static constexpr int kSliceK = kSliceBytes / (W * 4);      // K elements per ring slot
  static constexpr int kSlicesPerLayer = W / kSliceK;
  static constexpr int kStepsPerSlice = kSliceK / 8;
  static constexpr int kThreads = (4 * NT + 2) * 32;
  static size_t bytes(int nb) {
    (void)nb;
    return (size_t)NT * kABytes + (size_t)kSlots * kSliceBytes + 4096 + (size_t)W * 4 + 16 + 256;
  }
};

// 32 accumulator columns of one row (bias already added by the MMA): ReLU, round
// to tf32, write the 8 operand chunks.
template <bool kRelu, bool kMask, int W>
__device__ __forceinline__ uint32_t emit_chunks(const uint32_t (&v)[32], float4* __restrict__ a_tile,
                                                int c32, int row, float4* __restrict__ g_slot) {
  uint32_t bits = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 o = make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                           __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
    if (kRelu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    const float4 r4 = tf32_rn4(o);
    a_tile[(c32 * 8 + q) * 128 + row] = r4;
    if (g_slot) g_slot[TrainTile<W>::chunk_f4(row, c32 * 8 + q)] = r4;
    if (kMask) {
      bits |= (r4.x > 0.f ? 1u : 0u) << (q * 4 + 0);
      bits |= (r4.y > 0.f ? 1u : 0u) << (q * 4 + 1);
      bits |= (r4.z > 0.f ? 1u : 0u) << (q * 4 + 2);
      bits |= (r4.w > 0.f ? 1u : 0u) << (q * 4 + 3);
    }
  }
  return bits;
}

// kSave: the instance that also writes operand tiles / ReLU masks for a reverse
// pass; the inference instance carries none of that work.
template <int W, int NT, bool kSave>
__global__ void __launch_bounds__((4 * NT + 2) * 32, 1) mlp_fwd_umma_kernel(FwdParams p) {
  using S = FwdSmem<W, NT>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_base = smem;
  unsigned char* ring = smem + NT * S::kABytes;
  float4* ones_tile = reinterpret_cast<float4*>(ring + kSlots * kSliceBytes);   // [2][128]
  float* vecs = reinterpret_cast<float*>(ones_tile + 256);                        // we[W], be
  uint64_t* bars = reinterpret_cast<uint64_t*>(vecs + W + 4);
  uint64_t* a_ready = bars;                 // [NT]
  uint64_t* d_ready = bars + NT;            // [NT]
  uint64_t* w_full = bars + 2 * NT;         // [kSlots]
  uint64_t* w_empty = bars + 2 * NT + kSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NT + 2 * kSlots);
  __shared__ int s_abort;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nsteps = 2 * p.nb + 1;

  if (tid == 0) {
    s_abort = 0;
    for (int t = 0; t < NT; ++t) { mbar_init(&a_ready[t], 128); mbar_init(&d_ready[t], 1); }
    for (int s = 0; s < kSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < W + 4; i += S::kThreads)
    vecs[i] = p.packed[p.off_vecs + (int64_t)(2 * p.nb + 1) * W + i];
  for (int i = tid; i < 256; i += S::kThreads)
    ones_tile[i] = (i < 128) ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  const float* we = vecs;
  const float be = vecs[W];

  if (warp < 4 * NT) {
    // ===================== epilogue warpgroup t ==============================
    const int t = warp >> 2;
    const int row = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col_h = (uint32_t)(t * 2 * W), col_v = (uint32_t)(t * 2 * W + W);
    float4* a_tile = reinterpret_cast<float4*>(a_base + t * S::kABytes);
    uint32_t ph_d = 0;
    for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
      const int64_t r = ((int64_t)g * NT + t) * 128 + row;
      const bool valid = r < p.R;
      // step-0 A operand: this row's action, zero padded to KP
      for (int c = 0; c < p.KP / 4; ++c) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          const float* a = p.act + r * p.A;
          const int k = c * 4;
          if (k + 0 < p.A) o.x = a[k + 0];
          if (k + 1 < p.A) o.y = a[k + 1];
          if (k + 2 < p.A) o.z = a[k + 2];
          if (k + 3 < p.A) o.w = a[k + 3];
        }
        a_tile[c * 128 + row] = tf32_rn4(o);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_ready[t]);
      const float* hp_row = p.hp + (valid ? (r / p.n_per_obs) : 0) * W;
      const int64_t tile_id = (int64_t)g * NT + t;
      for (int i = 0; i < nsteps; ++i) {
        float4* g_slot = (kSave && p.xa_save)
            ? reinterpret_cast<float4*>(p.xa_save + TrainTile<W>::slot_base(tile_id, nsteps, i))
            : nullptr;
        uint32_t* m_slot = (kSave && p.mask_save && i < nsteps - 1)
            ? p.mask_save + TrainTile<W>::mask_index(tile_id, p.nb, i, row)
            : nullptr;
        mbar_wait(&d_ready[t], ph_d, abort_flag, p.status, 100 + i);
        ph_d ^= 1;
        tc_fence_after();
        const bool stamp = p.dbg && blockIdx.x == 0 && g == (int)blockIdx.x && row == 0 && i < 64;
        if (stamp) p.dbg[(i * NT + t) * 5 + 0] = clock64();
        const bool is_h = (i & 1) == 0;
        const uint32_t taddr = tmem_base + lane_off + (is_h ? col_h : col_v);
        if (i == nsteps - 1) {
          // Dense(1): E = (h + cb) . we + be   (mlp_ebm.py:91-94)
          float e = 0.f;
#pragma unroll 1
          for (int cb = 0; cb < W / 128; ++cb) {
            // four 32-column loads in flight, one wait: the TMEM latency is paid once
            uint32_t v[4][32];
#pragma unroll
            for (int j = 0; j < 4; ++j) tmem_ld32(taddr + cb * 128 + j * 32, v[j]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c32 = cb * 4 + j;
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                float x = __uint_as_float(v[j][q]);
                if (i == 0) x += hp_row[c32 * 32 + q];
                e = fmaf(x, we[c32 * 32 + q], e);
                v[j][q] = __float_as_uint(x);
              }
              if (g_slot) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  g_slot[TrainTile<W>::chunk_f4(row, c32 * 8 + q)] = make_float4(
                      __uint_as_float(v[j][q * 4 + 0]), __uint_as_float(v[j][q * 4 + 1]),
                      __uint_as_float(v[j][q * 4 + 2]), __uint_as_float(v[j][q * 4 + 3]));
              }
            }
          }
          if (valid) p.E[r] = e + be;
          tc_fence_before();
        } else {
#pragma unroll 1
          for (int cb = 0; cb < W / 128; ++cb) {
            uint32_t v[4][32];
            uint32_t mb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tmem_ld32(taddr + cb * 128 + j * 32, v[j]);
            tmem_ld_wait();
            if (stamp && cb == 0) p.dbg[640 + i * NT + t] = clock64();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c32 = cb * 4 + j;
              if (i == 0) {
                // h0 = act . Wp[O:] + (obs . Wp[:O] + bp); keep it in tensor memory
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 hv = *reinterpret_cast<const float4*>(hp_row + c32 * 32 + q * 4);
                  v[j][q * 4 + 0] = __float_as_uint(__uint_as_float(v[j][q * 4 + 0]) + hv.x);
                  v[j][q * 4 + 1] = __float_as_uint(__uint_as_float(v[j][q * 4 + 1]) + hv.y);
                  v[j][q * 4 + 2] = __float_as_uint(__uint_as_float(v[j][q * 4 + 2]) + hv.z);
                  v[j][q * 4 + 3] = __float_as_uint(__uint_as_float(v[j][q * 4 + 3]) + hv.w);
                }
                tmem_st32(taddr + c32 * 32, v[j]);
              }
              mb[j] = emit_chunks<true, kSave, W>(v[j], a_tile, c32, row, g_slot);
            }
            if (m_slot)
              *reinterpret_cast<uint4*>(m_slot + cb * 4) = make_uint4(mb[0], mb[1], mb[2], mb[3]);
          }
          if (i == 0) tmem_st_wait();
          if (stamp) p.dbg[(i * NT + t) * 5 + 1] = clock64();
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(&a_ready[t]);
          if (stamp) p.dbg[(i * NT + t) * 5 + 2] = clock64();
        }
      }
    }
  } else if (warp == 4 * NT) {
    // ===================== producer: weight slices via bulk copy ==============
    // converged warp, elected lane (see umma::elect_one)
    {
      uint32_t q = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int i = 0; i < nsteps; ++i) {
          const int nsl = (i == 0) ? (p.KP + S::kSliceK - 1) / S::kSliceK : S::kSlicesPerLayer;
          const float* src =
              (i == 0) ? p.packed + p.off_fwd_p : p.packed + p.off_fwd_l + (int64_t)(i - 1) * W * W;
          for (int s = 0; s < nsl; ++s, ++q) {
            const int ke = (i == 0) ? min(S::kSliceK, p.KP - s * S::kSliceK) : S::kSliceK;
            const uint32_t bytes = (uint32_t)(W * ke * 4);
            const int slot = q % kSlots;
            const uint32_t ph = (q / kSlots) & 1;
            if (elect_one()) {
              mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 200 + i);
              if (p.dbg && blockIdx.x == 0 && g == (int)blockIdx.x && i < 32 && s < 8)
                p.dbg[1024 + i * 8 + s] = clock64();
              mbar_arrive_expect_tx(&w_full[slot], bytes);
              bulk_g2s(ring + slot * kSliceBytes, src + (int64_t)s * (kSliceBytes / 4), bytes,
                       &w_full[slot]);
            }
            __syncwarp();
          }
          if (i > 0) {   // this step's bias as one more K = 8 operand slice
            const int slot = q % kSlots;
            const uint32_t ph = (q / kSlots) & 1;
            if (elect_one()) {
              mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 250 + i);
              mbar_arrive_expect_tx(&w_full[slot], (uint32_t)(8 * W * 4));
              bulk_g2s(ring + slot * kSliceBytes, p.packed + p.off_fwd_bias + (int64_t)i * 8 * W,
                       (uint32_t)(8 * W * 4), &w_full[slot]);
            }
            __syncwarp();
            ++q;
          }
        }
      }
    }
  } else {
    // ===================== MMA issuer ==========================================
    // The warp stays converged; every tcgen05 instruction is issued by the lane
    // elect.sync picks (see umma::elect_one).
    {
      const uint32_t idesc = idesc_tf32(128, W);
      const uint32_t a_lbo = 128 * 16, b_lbo = W * 16, sbo = 128;
      uint32_t q = 0;
      uint32_t ph_a[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) ph_a[t] = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int i = 0; i < nsteps; ++i) {
          const int nsl = (i == 0) ? (p.KP + S::kSliceK - 1) / S::kSliceK : S::kSlicesPerLayer;
          const bool to_h = (i & 1) == 0;
          const uint32_t acc0 = (to_h && i > 0) ? 1u : 0u;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const bool mstamp = p.dbg && blockIdx.x == 0 && g == (int)blockIdx.x && i < 32;
            long long wf_wait = 0;
            const long long t_top = mstamp ? clock64() : 0;
            if (elect_one()) mbar_wait(&a_ready[t], ph_a[t], abort_flag, p.status, 300 + i);
            __syncwarp();
            ph_a[t] ^= 1;
            tc_fence_after();
            const long long t_start = mstamp ? clock64() : 0;
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * 2 * W + (to_h ? 0 : W));
            const uint64_t a_desc0 = smem_desc(smem_u32(a_base + t * S::kABytes), a_lbo, sbo);
            constexpr uint32_t a_step16 = (2 * 128 * 16) >> 4, b_step16 = (2 * W * 16) >> 4;
            if (i == 0) {
              for (int s = 0; s < nsl; ++s) {
                const uint32_t qq = q + s;
                const int slot = qq % kSlots;
                if (t == 0) mbar_wait(&w_full[slot], (qq / kSlots) & 1, abort_flag, p.status, 400);
                const uint64_t b_desc0 = smem_desc(smem_u32(ring + slot * kSliceBytes), b_lbo, sbo);
                const int kps = min(S::kSliceK, p.KP - s * S::kSliceK) / 8;
                if (elect_one()) {
                  for (int ks = 0; ks < kps; ++ks) {
                    const int kk = s * S::kStepsPerSlice + ks;
                    mma_tf32_ss(d_tmem, a_desc0 + (uint64_t)(kk * a_step16),
                                b_desc0 + (uint64_t)(ks * b_step16), idesc, (uint32_t)(kk > 0));
                  }
                  if (t == NT - 1) mma_commit(&w_empty[slot]);
                }
                __syncwarp();
              }
            } else {
              // one elected lane issues the whole layer for this tile back to back (the
              // tensor pipe's issue queue is shallow: instructions between two MMAs
              // show up as pipe bubbles); tile 0 polls each slice's barrier itself
              uint32_t b_addr[S::kSlicesPerLayer + 1];
              uint32_t b_par[S::kSlicesPerLayer + 1];
              uint64_t* b_full[S::kSlicesPerLayer + 1];
              uint64_t* b_empty[S::kSlicesPerLayer + 1];
#pragma unroll
              for (int s = 0; s <= S::kSlicesPerLayer; ++s) {
                const uint32_t qq = q + s;
                const uint32_t slot = qq % kSlots;
                b_addr[s] = smem_u32(ring + slot * kSliceBytes);
                b_par[s] = (qq / kSlots) & 1;
                b_full[s] = &w_full[slot];
                b_empty[s] = &w_empty[slot];
              }
              const uint64_t ones_desc = smem_desc(smem_u32(ones_tile), a_lbo, sbo);
              if (elect_one()) {
#pragma unroll
                for (int s = 0; s < S::kSlicesPerLayer; ++s) {
                  if (t == 0) {
                    mbar_wait(b_full[s], b_par[s], abort_flag, p.status, 400 + i);
                    if (mstamp) p.dbg[1280 + i * 8 + s] = clock64();
                  }
                  mma_tf32_ss_steps<S::kStepsPerSlice>(
                      d_tmem, a_desc0 + (uint64_t)(s * S::kStepsPerSlice * a_step16),
                      smem_desc(b_addr[s], b_lbo, sbo), a_step16, b_step16, idesc,
                      s == 0 ? acc0 : 1u);
                  if (t == NT - 1) mma_commit(b_empty[s]);
                }
                // + bias: (1, 1, 0, ...) . (hi, lo, 0, ...)^T
                if (t == 0)
                  mbar_wait(b_full[S::kSlicesPerLayer], b_par[S::kSlicesPerLayer], abort_flag,
                            p.status, 450 + i);
                mma_tf32_ss(d_tmem, ones_desc, smem_desc(b_addr[S::kSlicesPerLayer], b_lbo, sbo),
                            idesc, 1u);
                if (t == NT - 1) mma_commit(b_empty[S::kSlicesPerLayer]);
                mma_commit(&d_ready[t]);
              }
              __syncwarp();
            }
            if (i == 0) {
              if (elect_one()) mma_commit(&d_ready[t]);
              __syncwarp();
            }
            if (mstamp && lane == 0) {
              p.dbg[(i * NT + t) * 5 + 3] = t_start;
              p.dbg[(i * NT + t) * 5 + 4] = clock64();
              p.dbg[768 + (i * NT + t) * 2 + 0] = t_start - t_top;
              p.dbg[768 + (i * NT + t) * 2 + 1] = wf_wait;
              (void)wf_wait;
            }
          }
          q += nsl + (i > 0 ? 1 : 0);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// Self test: one 128 x N x K tile through the same operand layout.
// ---------------------------------------------------------------------------
__global__ void pack_rows_kernel(const float* __restrict__ X, int rows, int K, int mn_major,
                                 float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  if (!mn_major) {   // [K/4][rows][4]
    const int kq = i & 3, r = (i >> 2) % rows, kc = (i >> 2) / rows;
    out[i] = X[(int64_t)r * K + kc * 4 + kq];
  } else if (mn_major == 1) {   // [rows/4][K][4]
    const int rq = i & 3, k = (i >> 2) % K, rc = (i >> 2) / K;
    out[i] = X[(int64_t)(rc * 4 + rq) * K + k];
  } else {           // MN-major 128B/32B-base swizzle atoms
    const int r = i / K, k = i % K;
    out[mn32_offset(r, k, (uint32_t)(K / 4) * 512u, 512u) / 4] = X[i];
  }
}

__global__ void __launch_bounds__(160, 1)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bpacked, int N, int K,
                     int mode, float* __restrict__ D, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float4* a_tile = reinterpret_cast<float4*>(smem);                       // [K/4][128]
  unsigned char* b_tile = smem + 128 * K * 4;                             // [K/4][N][16B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_tile + (size_t)N * K * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  __shared__ int s_abort;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    s_abort = 0;
    mbar_init(&bars[0], 1);   // B landed
    mbar_init(&bars[1], 1);   // MMA done
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  // tensor-memory copy of A for the TS modes (23 / 24 probe other column offsets)
  const uint32_t a_col = (mode == 23) ? 128 : ((mode == 24) ? 384 : 256);

  if (warp < 4) {
    const int row = tid;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    if (mode == 0 || mode == 20 || mode == 22 || (mode >= 25 && mode <= 27) || mode == 30 || mode == 32 || mode == 34 || mode == 36 || mode == 38 || mode == 39 || mode == 40 || mode == 42 || mode == 43) {
      for (int c = 0; c < K / 4; ++c)
        a_tile[c * 128 + row] = *reinterpret_cast<const float4*>(A + (int64_t)row * K + c * 4);
      fence_proxy_async_smem();
    } else if (mode == 3) {   // MN-major, SWIZZLE_128B_BASE32B atoms
      float* af = reinterpret_cast<float*>(smem);
      for (int k = 0; k < K; ++k)
        af[mn32_offset(row, k, (uint32_t)(K / 4) * 512u, 512u) / 4] = A[(int64_t)row * K + k];
      fence_proxy_async_smem();
    } else if (mode == 2) {   // MN-major staging: [row/4][K][4]
      float* af = reinterpret_cast<float*>(smem);
      for (int k = 0; k < K; ++k) af[((row >> 2) * K + k) * 4 + (row & 3)] = A[(int64_t)row * K + k];
      fence_proxy_async_smem();
    } else {
      for (int c32 = 0; c32 < K / 32; ++c32) {
        uint32_t v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(A[(int64_t)row * K + c32 * 32 + q]);
        tmem_st32(tmem_base + lane_off + a_col + c32 * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
    }
  } else if (lane == 0) {
    mbar_arrive_expect_tx(&bars[0], (uint32_t)(N * K * 4));
    bulk_g2s(b_tile, Bpacked, (uint32_t)(N * K * 4), &bars[0]);
  }
  __syncthreads();
  tc_fence_after();
  __shared__ volatile int s_done;
  if (tid == 0) s_done = 0;
  __syncthreads();
  __shared__ __align__(8) uint64_t s_hbar;
  if (mode >= 40 && tid == 0) { mbar_init(&s_hbar, 1); fence_mbar_init(); }
  __syncthreads();
  if (warp == 0 && mode >= 40) {
    // bulk-copy flood (40: SS MMA, 41: TS MMA): 4 x 16 KB global -> shared copies
    // in flight at all times while the MMA loop runs
    unsigned char* scratch = b_tile + (size_t)N * K * 4 + 1024;
    uint32_t ph = 0;
    // 40/41: 4 x 16 KB per round; 42: 1 x 16 KB per round (latency); 43: 1 x 64 KB
    const int ncopy = (mode == 42 || mode == 43) ? 1 : 4;
    const uint32_t cbytes = (mode == 43) ? 65536u : 16384u;
    long long rounds = 0, issue_cycles = 0;
    const long long f0 = clock64();
    while (!s_done) {
      if (lane == 0) {
        const long long i0 = clock64();
        mbar_arrive_expect_tx(&s_hbar, ncopy * cbytes);
        for (int c = 0; c < ncopy; ++c)
          bulk_g2s(scratch + c * 16384, Bpacked + (c & 1) * 4096, cbytes, &s_hbar);
        issue_cycles += clock64() - i0;
      }
      mbar_wait(&s_hbar, ph, abort_flag, status, 3);
      ph ^= 1;
      ++rounds;
    }
    if (lane == 0)
      printf("[bulk probe] mode %d: %lld rounds of %d x %u B in %lld cycles = %.1f B/cycle, %.0f cycles/round, issue %.0f cycles/round\n",
             mode, rounds, ncopy, cbytes, clock64() - f0,
             (double)rounds * ncopy * cbytes / (double)(clock64() - f0),
             (double)(clock64() - f0) / (double)rounds, (double)issue_cycles / (double)rounds);
  }
  if (warp < 4 && mode >= 34 && mode <= 37) {
    // contention probes: while the MMA thread runs its clean loop, the four
    // epilogue warps hammer shared memory (34: SS MMA, 37: TS MMA) or tensor
    // memory (35: TS MMA, 36: SS MMA) the way an epilogue would
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float4* scratch = reinterpret_cast<float4*>(b_tile + (size_t)N * K * 4 + 1024);   // 64 KB
    uint32_t v[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = tid + q;
    while (!s_done) {
      if (mode == 34 || mode == 37) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          scratch[c * 128 + tid] = make_float4(__uint_as_float(v[c]), 0.f, 1.f, 2.f);
      } else {
#pragma unroll 1
        for (int c32 = 0; c32 < 4; ++c32) {
          tmem_ld32(tmem_base + lane_off + 128 + c32 * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(fmaxf(__uint_as_float(v[q]), 0.f));
          tmem_st32(tmem_base + lane_off + 128 + c32 * 32, v);
          tmem_st_wait();
        }
      }
    }
    tc_fence_before();
  }
  if (warp == 4 && lane == 0 && mode >= 20) {
    // MMA issue-rate microbenchmark: 256 passes over the K steps, results unused.
    // 20: A,B from shared memory (no swizzle); 21: A from tensor memory;
    // 22: A,B descriptors in the 128B-swizzle K-major form (timing only)
    mbar_wait(&bars[0], 0, abort_flag, status, 1);
    const uint32_t idesc = idesc_tf32(128, N);
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)N * 16, sbo = 128;
    bool return_early = false;
    const int nk = K / 8;
    if (mode >= 30) {
      // Clean pipe-rate probes (K = 128: 16 unrolled steps, descriptors hoisted).
      // 30: SS one accumulator; 31: TS one accumulator; 32: SS two alternating
      // accumulators; 33: TS two alternating accumulators (N <= 128)
      const uint64_t ad0 = smem_desc(smem_u32(a_tile), a_lbo, sbo);
      const uint64_t bd0 = smem_desc(smem_u32(b_tile), b_lbo, sbo);
      const uint32_t a_step16 = (2 * a_lbo) >> 4, b_step16 = (2 * b_lbo) >> 4;
      const uint32_t d1 = (N <= 128) ? 128u : 256u;
      const long long t0 = clock64();
      if (mode == 30 || mode == 34 || mode == 36 || mode == 40 || mode == 42 || mode == 43) {
        for (int rep = 0; rep < 64; ++rep)
          mma_tf32_ss_steps<16>(tmem_base, ad0, bd0, a_step16, b_step16, idesc, 1u);
      } else if (mode == 31 || mode == 35 || mode == 37 || mode == 41) {
        for (int rep = 0; rep < 64; ++rep)
          mma_tf32_ts_steps<16>(tmem_base, tmem_base + a_col, bd0, b_step16, idesc, 1u);
      } else if (mode == 38 || mode == 39) {
        // commit-cost probe: the same MMAs with a tcgen05.commit (to a barrier nobody
        // waits on) after every 4 (38) or 16 (39) of them
        uint64_t* dummy = &bars[0];   // B already landed; extra arrivals are harmless here
        for (int rep = 0; rep < 64; ++rep) {
          if (mode == 38) {
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              mma_tf32_ss_steps<4>(tmem_base, ad0 + (uint64_t)(s4 * 4 * a_step16),
                                   bd0 + (uint64_t)(s4 * 4 * b_step16), a_step16, b_step16, idesc, 1u);
              mma_commit(dummy);
            }
          } else {
            mma_tf32_ss_steps<16>(tmem_base, ad0, bd0, a_step16, b_step16, idesc, 1u);
            mma_commit(dummy);
          }
        }
      } else if (mode == 32) {
        for (int rep = 0; rep < 32; ++rep) {
          mma_tf32_ss_steps<16>(tmem_base, ad0, bd0, a_step16, b_step16, idesc, 1u);
          mma_tf32_ss_steps<16>(tmem_base + d1, ad0, bd0, a_step16, b_step16, idesc, 1u);
        }
      } else {
        for (int rep = 0; rep < 32; ++rep) {
          mma_tf32_ts_steps<16>(tmem_base, tmem_base + a_col, bd0, b_step16, idesc, 1u);
          mma_tf32_ts_steps<16>(tmem_base + d1, tmem_base + a_col, bd0, b_step16, idesc, 1u);
        }
      }
      mma_commit(&bars[1]);
      mbar_wait(&bars[1], 0, abort_flag, status, 2);
      const long long t1 = clock64();
      if (blockIdx.x == 0) atomicExch(status, (int)((t1 - t0) * 10 / (64 * 16)));
      s_done = 1;
      return_early = true;
    }
    const long long t0 = clock64();
    for (int rep = 0; rep < (return_early ? 0 : 256); ++rep)
      for (int ks = 0; ks < nk; ++ks) {
        if (mode >= 25 && mode <= 27) {
          // accumulator-dependency probes: the same work as mode 20 split over several
          // independent accumulators.  25: two N halves, 26: four N quarters,
          // 27: two full-N accumulators alternating (N <= 128 so both fit)
          const uint64_t ad = smem_desc(smem_u32(a_tile) + (uint32_t)ks * 2 * a_lbo, a_lbo, sbo);
          if (mode == 27) {
            const uint64_t bd = smem_desc(smem_u32(b_tile) + (uint32_t)ks * 2 * b_lbo, b_lbo, sbo);
            mma_tf32_ss(tmem_base + ((rep & 1) ? 256u : 0u), ad, bd, idesc, 1);
          } else {
            const int parts = (mode == 25) ? 2 : 4;
            const int np = N / parts;
            const uint32_t idp = idesc_tf32(128, np);
            for (int part = 0; part < parts; ++part) {
              const uint64_t bd = smem_desc(smem_u32(b_tile) + (uint32_t)ks * 2 * b_lbo +
                                                (uint32_t)(part * np) * 16, b_lbo, sbo);
              mma_tf32_ss(tmem_base + (uint32_t)(part * np), ad, bd, idp, 1);
            }
          }
        } else if (mode == 22) {
          const uint32_t ao = (uint32_t)(ks & 3) * 32 + (uint32_t)(ks >> 2) * (128 * 128);
          const uint32_t bo = (uint32_t)(ks & 3) * 32 + (uint32_t)(ks >> 2) * ((uint32_t)N * 128);
          mma_tf32_ss(tmem_base, smem_desc_lt(smem_u32(a_tile) + ao, 16, 1024, 2),
                      smem_desc_lt(smem_u32(b_tile) + bo, 16, 1024, 2), idesc, 1);
        } else if (mode == 21 || mode == 23 || mode == 24) {
          const uint64_t bd = smem_desc(smem_u32(b_tile) + (uint32_t)ks * 2 * b_lbo, b_lbo, sbo);
          mma_tf32_ts(tmem_base, tmem_base + a_col + ks * 8, bd, idesc, 1);
        } else {
          const uint64_t bd = smem_desc(smem_u32(b_tile) + (uint32_t)ks * 2 * b_lbo, b_lbo, sbo);
          const uint64_t ad = smem_desc(smem_u32(a_tile) + (uint32_t)ks * 2 * a_lbo, a_lbo, sbo);
          mma_tf32_ss(tmem_base, ad, bd, idesc, 1);
        }
      }
    if (!return_early) {
      mma_commit(&bars[1]);
      mbar_wait(&bars[1], 0, abort_flag, status, 2);
      const long long t1 = clock64();
      atomicExch(status, (int)((t1 - t0) * 10 / (256 * nk)));      // 0.1 cycles per MMA
    }
  } else if (warp == 4 && lane == 0) {
    mbar_wait(&bars[0], 0, abort_flag, status, 1);
    const uint32_t idesc = idesc_tf32(128, N);
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)N * 16, sbo = 128;
    for (int ks = 0; ks < K / 8; ++ks) {
      if (mode == 3) {
        // both operands MN-major in the 128B / 32-byte-base swizzled layout:
        // LBO = stride between 32-element MN atoms, SBO = between 4-row K atoms
        const uint32_t lbo = (uint32_t)(K / 4) * 512u;
        const uint64_t ad = smem_desc_lt(smem_u32(a_tile) + (uint32_t)ks * 1024, lbo, 512, 1);
        const uint64_t bd = smem_desc_lt(smem_u32(b_tile) + (uint32_t)ks * 1024, lbo, 512, 1);
        mma_tf32_ss(tmem_base, ad, bd, idesc | kIdescAMajorMN | kIdescBMajorMN, ks > 0);
        continue;
      }
      if (mode == 2) {
        // both operands MN-major: 8 K-rows per step are 128 contiguous bytes per
        // 4-element MN chunk; SBO = stride between MN chunks, LBO = between 8-row groups
        const uint32_t mn_sbo = (uint32_t)K * 16;
        const uint64_t ad = smem_desc(smem_u32(a_tile) + (uint32_t)ks * 128, 128, mn_sbo);
        const uint64_t bd = smem_desc(smem_u32(b_tile) + (uint32_t)ks * 128, 128, mn_sbo);
        mma_tf32_ss(tmem_base, ad, bd, idesc | kIdescAMajorMN | kIdescBMajorMN, ks > 0);
        continue;
      }
      const uint64_t bd = smem_desc(smem_u32(b_tile) + (uint32_t)ks * 2 * b_lbo, b_lbo, sbo);
      if (mode == 0) {
        const uint64_t ad = smem_desc(smem_u32(a_tile) + (uint32_t)ks * 2 * a_lbo, a_lbo, sbo);
        mma_tf32_ss(tmem_base, ad, bd, idesc, ks > 0);
      } else {
        mma_tf32_ts(tmem_base, tmem_base + a_col + ks * 8, bd, idesc, ks > 0);
      }
    }
    mma_commit(&bars[1]);
  }
  __syncwarp();
  if (warp < 4) {
    mbar_wait(&bars[1], 0, abort_flag, status, 2);
    tc_fence_after();
    const int row = tid;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int c16 = 0; c16 < N / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + lane_off + c16 * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) D[(int64_t)row * N + c16 * 16 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int W, int NT, bool kSave>
int launch_fwd_impl(ibc_handle* h, const FwdParams& p, cudaStream_t s) {
  using S = FwdSmem<W, NT>;
  const size_t smem = S::bytes(p.nb);
  if (smem > 227 * 1024)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "tcgen05 forward: depth %d needs %zu B of shared memory",
             2 * p.nb, smem);
  IBC_RAISE_SMEM(h, (mlp_fwd_umma_kernel<W, NT, kSave>), 227 * 1024 - 1024);    // per template instance
  const int grid = p.num_groups < h->sm_count ? p.num_groups : h->sm_count;
  mlp_fwd_umma_kernel<W, NT, kSave><<<grid, S::kThreads, smem, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

template <int W, int NT>
int launch_fwd(ibc_handle* h, const FwdParams& p, cudaStream_t s) {
  return (p.xa_save || p.mask_save) ? launch_fwd_impl<W, NT, true>(h, p, s)
                                    : launch_fwd_impl<W, NT, false>(h, p, s);
}

}  // namespace

bool umma_supported(const Layout& L) {
  return (L.W == 128 || L.W == 256) && L.nb >= 1 && L.A >= 1 && L.A <= 32 && L.O >= 1;
}

// The operand copies are refreshed in two parts, because a training step re-packs after every
// optimizer update (25 us of two launches for all of them, 5 % of the C2 step): the CORE -- what
// the layer-fused forward and the block-major backward read: action rows, forward layers with
// their bias slices, vectors, and (width 128) the fp16 transposed layers -- and the REST: the
// tf32 reverse layers and the shared-memory-operand kernels' copies (fwd_l, rev_l, rev_p:
// contiguous in the buffer), the fp16 forward layers.  umma_pack_core is for the two hot paths
// that need nothing else; every other caller uses umma_pack_params, which brings both up to date.
static int ensure_packed_buffer(ibc_handle* h, const PackedLayout& pl) {
  if (!h->packed) {
    IBC_CUDA(h, cudaMalloc((void**)&h->packed, (size_t)pl.total * 4));
    h->packed_bytes = (size_t)pl.total * 4;
  }
  return IBC_OK;
}

static int pack_range(ibc_handle* h, const PackedLayout& pl, int64_t lo, int64_t hi, int64_t hole_lo,
                      int64_t hole_len, cudaStream_t s) {
  // (grid: one element per thread up to a full resident wave -- a latency-bound gather out of L2)
  const int64_t cnt = hi - lo - hole_len;
  if (cnt <= 0) return IBC_OK;
  const int grid = (int)std::min<int64_t>((cnt + 255) / 256, 148 * 8);
  pack_kernel<<<grid, 256, 0, s>>>(h->params, h->L, pl, h->packed, lo, hi, hole_lo, hole_len);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// With the training forward in kind::f16 (train_forward_is_f16: width 128, act_dim <= 16) the
// layer-fused step reads its MMA operands from the fp16 images only; of this buffer it reads the
// epilogue vectors (we, be, biases: [vecs, fwd_lb)).  The tf32 forward layer blocks -- 1.1 MB,
// the bulk of the re-pack -- then wait for the first caller that needs them (umma_pack_params).
static bool core_is_slim(const ibc_handle* h) { return h->L.W == 128 && train_forward_is_f16(h->L); }

int umma_pack_core(ibc_handle* h, cudaStream_t s) {
  if (!h->packed_stale) return IBC_OK;
  const PackedLayout pl = make_packed_layout(h->L);
  IBC_TRY(ensure_packed_buffer(h, pl));
  if (core_is_slim(h)) IBC_TRY(pack_range(h, pl, pl.vecs, pl.fwd_lb, pl.fwd_lb, 0, s));
  else IBC_TRY(pack_range(h, pl, 0, pl.total, pl.fwd_l, pl.vecs - pl.fwd_l, s));
  if (h->L.W == 128) IBC_TRY(pack16_params(h, s, true));     // fp16 transposed layers (backward)
  h->packed_stale = false;
  return IBC_OK;
}

int umma_pack_params(ibc_handle* h, cudaStream_t s) {
  IBC_TRY(umma_pack_core(h, s));
  if (!h->packed_rest_stale) return IBC_OK;
  const PackedLayout pl = make_packed_layout(h->L);
  if (core_is_slim(h)) IBC_TRY(pack_range(h, pl, 0, pl.total, pl.vecs, pl.fwd_lb - pl.vecs, s));
  else IBC_TRY(pack_range(h, pl, pl.fwd_l, pl.vecs, pl.vecs, 0, s));
  if (h->L.W == 128) IBC_TRY(pack16_params(h, s, false));    // fp16 forward layers (Langevin chain)
  h->packed_rest_stale = false;
  return IBC_OK;
}

int mlp_energy_umma(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                    float* E, float* dEda, bool apply_exp, cudaStream_t s) {
  const Layout& L = h->L;
  if (wide_supported(L)) return mlp_energy_wide(h, obs, B, act, n, E, dEda, apply_exp, s);
  if (!umma_supported(L))
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED,
             "tcgen05 path supports width 128/256 (act_dim <= 32) and multiples of 128 above");
  if (dEda) return mlp_energy_dact_umma(h, obs, B, act, n, E, dEda, apply_exp, s);
  IBC_CUDA(h, h->ws.reserve(Workspace::rounded((size_t)B * L.W * 4)));
  h->ws.reset();
  float* hp = h->ws.take<float>((size_t)B * L.W);
  return umma_forward(h, obs, B, act, n, hp, E, nullptr, nullptr, s);
}

// Hoisted observation projection + fused forward.  hp [B, W] is caller scratch;
// xa_save (nullable) receives the saved operand tiles for the backward pass.
int umma_forward(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                 float* hp, float* E, float* xa_save, uint32_t* mask_save, cudaStream_t s) {
  const Layout& L = h->L;
  IBC_TRY(umma_pack_params(h, s));
  {  // hp = obs . Wp[:O] + bp   (fp32)
    GemmP g;
    g.A = obs; g.lda = L.O; g.B = h->params + L.wp; g.ldb = L.W; g.C = hp; g.ldc = L.W;
    g.bias = h->params + L.bp; g.M = (int)B; g.N = L.W; g.K = L.O;
    IBC_TRY(gemm(h, g, false, false, s));
  }
  return umma_forward_hp(h, act, B * n, n, hp, E, xa_save, mask_save, s);
}

// Fused forward given the hoisted observation projection hp [R / n, W].
int umma_forward_hp(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp,
                    float* E, float* xa_save, uint32_t* mask_save, cudaStream_t s) {
  const Layout& L = h->L;
  IBC_TRY(umma_pack_params(h, s));
  const PackedLayout pl = make_packed_layout(L);
  // width 128: activations in tensor memory; IBC_SMEM_OPERANDS=1 selects the
  // shared-memory-operand kernel below instead (kept for width 256 and for A/B runs)
  static const bool force_smem = getenv("IBC_SMEM_OPERANDS") != nullptr;
  if (!force_smem && tmem_path_supported(L))
    return tmem_forward(h, act, R, n, hp, E, xa_save, mask_save, s);
  FwdParams p;
  p.act = act; p.hp = hp; p.packed = h->packed; p.E = E; p.R = R; p.n_per_obs = n;
  p.nb = L.nb; p.A = L.A; p.KP = pl.KP;
  p.off_fwd_p = pl.fwd_p; p.off_fwd_l = pl.fwd_l; p.off_vecs = pl.vecs;
  p.off_fwd_bias = pl.fwd_bias;
  p.status = h->dev_status;
  p.xa_save = xa_save;
  p.mask_save = mask_save;
  p.dbg = nullptr;
  static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
  if (want_dbg) {
    IBC_CUDA(h, cudaMalloc((void**)&p.dbg, (640 + 128 + 256 + 512) * sizeof(long long)));
    IBC_CUDA(h, cudaMemsetAsync(p.dbg, 0, (640 + 128 + 256 + 512) * sizeof(long long), s));
  }
  struct DbgDump {
    long long* d; int nsteps, nt; cudaStream_t s;
    ~DbgDump() {
      if (!d) return;
      cudaStreamSynchronize(s);
      std::vector<long long> hbuf(640 + 128 + 256 + 512);
      cudaMemcpy(hbuf.data(), d, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      cudaFree(d);
      long long t0 = hbuf[3];
      fprintf(stderr, "[ibc timeline] step tile | mma_start mma_issued | epi_start epi_done arrived (cycles from first MMA)\n");
      for (int i = 0; i < nsteps && i < 64; ++i)
        for (int t = 0; t < nt; ++t) {
          const long long* e = &hbuf[(i * nt + t) * 5];
          fprintf(stderr, "[ibc timeline] %3d %d | %8lld %8lld | %8lld (tmem loaded +%5lld) %8lld %8lld | a_wait %5lld w_wait %5lld\n", i, t,
                  e[3] - t0, e[4] - t0, e[0] - t0, hbuf[640 + i * nt + t] ? hbuf[640 + i * nt + t] - e[0] : 0,
                  e[1] ? e[1] - t0 : 0, e[2] ? e[2] - t0 : 0,
                  i < 32 ? hbuf[768 + (i * nt + t) * 2] : 0, i < 32 ? hbuf[768 + (i * nt + t) * 2 + 1] : 0);
        }
      for (int i = 8; i < 11 && i < nsteps; ++i)
        for (int sl = 0; sl < 4; ++sl)
          fprintf(stderr, "[ibc ring] step %d slice %d: slot free seen %8lld, data seen by MMA %8lld (+%lld)\n",
                  i, sl, hbuf[1024 + i * 8 + sl] - t0, hbuf[1280 + i * 8 + sl] - t0,
                  hbuf[1280 + i * 8 + sl] - hbuf[1024 + i * 8 + sl]);
    }
  } dump{p.dbg, 2 * L.nb + 1, L.W == 128 ? 2 : 1, s};
  if (L.W == 128) {
    p.num_groups = (int)((R + 255) / 256);
    IBC_TRY((launch_fwd<128, 2>(h, p, s)));
  } else {
    p.num_groups = (int)((R + 127) / 128);
    IBC_TRY((launch_fwd<256, 1>(h, p, s)));
  }
  return IBC_OK;
}

int umma_selftest(const float* A, const float* Bt, int N, int K, int mode, float* D,
                  int* status_host, cudaStream_t s) {
  ibc_handle* h = nullptr;
  // modes 50+: the same probe as mode - 20 on every SM at once (chip-wide rate)
  int grid = 1;
  if (mode >= 50) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&grid, cudaDevAttrMultiProcessorCount, dev);
    mode -= 20;
  }
  if (N < 16 || N > 256 || (N % 16) || K < 8 || K > 128 || (K % 8) ||
      ((mode == 1 || mode == 21 || mode == 23 || mode == 24 || mode == 31 || mode == 33 || mode == 35 || mode == 37 || mode == 41) && (K % 32)))
    IBC_FAIL(h, IBC_ERR_INVALID, "umma_selftest: need 16<=N<=256 (x16), 8<=K<=128 (x8; x32 for mode 1)");
  float* bp = nullptr;
  int* st = nullptr;
  IBC_CUDA(h, cudaMalloc((void**)&bp, (size_t)N * K * 4));
  IBC_CUDA(h, cudaMalloc((void**)&st, sizeof(int)));
  IBC_CUDA(h, cudaMemsetAsync(st, 0, sizeof(int), s));
  pack_rows_kernel<<<(N * K + 255) / 256, 256, 0, s>>>(Bt, N, K, mode == 2 ? 1 : (mode == 3 ? 2 : 0), bp);
  count_launch();
  const size_t smem = (size_t)128 * K * 4 + (size_t)N * K * 4 + 64 + ((mode >= 34 && mode <= 37) || mode >= 40 ? 66 * 1024 : 0);
  IBC_CUDA(h, cudaFuncSetAttribute(umma_selftest_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  umma_selftest_kernel<<<grid, 160, smem, s>>>(A, bp, N, K, mode, D, st);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  int st_host = -1;
  if (e == cudaSuccess) e = cudaMemcpy(&st_host, st, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(bp);
  cudaFree(st);
  if (status_host) *status_host = st_host;
  if (e != cudaSuccess) IBC_FAIL(h, IBC_ERR_CUDA, "umma_selftest: %s", cudaGetErrorString(e));
  return IBC_OK;
}

}  // namespace ibc
