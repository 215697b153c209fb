// Training backward of the MLPEBM on tcgen05 (kind::tf32) for width 128:
//   tape.gradient(loss, variables)       ibc/agents/base_agent.py:43-48
//   of the InfoNCE loss                  ibc/losses/ebm_loss.py:21-48
//   through the ResNetPreActivation body networks/layers/resnet.py:135-153
//
// Three fused kernels after the forward (mlp_umma.cu, which saves the A
// operand of every W x W layer per 128-row tile):
//   mlp_bwd_umma_kernel   reverse chain per tile.  The gradient of the
//       residual stream g stays in tensor memory; every layer's output
//       gradient GY is written once (tf32, tile layout) for the weight
//       gradient, ReLU masks come from the saved operands (x > 0).
//   mlp_wgrad_umma_kernel dW_i = XA_i^T . GY_i as one tcgen05 GEMM per layer
//       with K = all rows: both saved tiles are consumed as MN-major operands
//       straight from their bulk-copied (TMA engine) shared-memory image,
//       accumulating in tensor memory across all tiles of the CTA; the bias
//       gradient (fp32 column sums of GY) is taken by the otherwise idle
//       epilogue warps while the tiles sit in shared memory.
//   first / last layer pieces (K = act_dim, obs_dim; Dense(1)) reuse the fp32
//       kernels: they are O(R * W) work.
#include <stdlib.h>

#include "chain_graph.cuh"
#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "ops.cuh"
#include "umma.cuh"
#include "philox.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

inline const float* act_const(const float* p) { return p; }

constexpr int kSlots = 5;
constexpr int kSliceBytes = 16384;
constexpr int kW = 128;          // the only width this path is instantiated for
constexpr int kNT = 2;

// ---------------------------------------------------------------------------
// reverse chain
// ---------------------------------------------------------------------------
struct BwdParams {
  const float* dE;        // [R]   d(sum loss / global_batch) / dE
  const float* xa_save;   // forward operand tiles, 2nb+1 slots (only slot 2nb is read, for dwe)
  const uint32_t* mask_save;  // ReLU mask bits of the 2nb layers
  float* gy_save;         // output-gradient tiles, 2nb slots (slot i-1 = GY of layer i)
  float* g0;              // [R, W] row-major gradient of the first Dense output
  const float* packed;
  int64_t off_rev_l, off_we;
  float* dwe;             // [W] accumulated with atomics
  int64_t R;
  int nb, num_groups;
  int* status;
};

template <int W_, int NT_>
struct BwdSmem {
  static constexpr int kABytes = 128 * W_ * 4;
  static constexpr int kSliceK = kSliceBytes / (W_ * 4);
  static constexpr int kSlicesPerLayer = W_ / kSliceK;
  static constexpr int kStepsPerSlice = kSliceK / 8;
  static constexpr int kThreads = (4 * NT_ + 2) * 32;
  static constexpr size_t kBytes = (size_t)NT_ * kABytes + (size_t)kSlots * kSliceBytes +
                                   2 * W_ * 4 + 256;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int W_, int NT_>
__global__ void __launch_bounds__((4 * NT_ + 2) * 32, 1) mlp_bwd_umma_kernel(BwdParams p) {
  using S = BwdSmem<W_, NT_>;
  using TT = TrainTile<W_>;
  constexpr int W = W_, NT = NT_;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_base = smem;
  unsigned char* ring = smem + NT * S::kABytes;
  float* we = reinterpret_cast<float*>(ring + kSlots * kSliceBytes);
  float* dwe_acc = we + W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dwe_acc + W);
  uint64_t* a_ready = bars;
  uint64_t* d_ready = bars + NT;
  uint64_t* w_full = bars + 2 * NT;
  uint64_t* w_empty = bars + 2 * NT + kSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NT + 2 * kSlots);
  __shared__ int s_abort;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nsteps = 2 * p.nb;
  const int xa_slots = 2 * p.nb + 1, gy_slots = 2 * p.nb;

  if (tid == 0) {
    s_abort = 0;
    for (int t = 0; t < NT; ++t) { mbar_init(&a_ready[t], 128); mbar_init(&d_ready[t], 1); }
    for (int s = 0; s < kSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < W; i += S::kThreads) { we[i] = p.packed[p.off_we + i]; dwe_acc[i] = 0.f; }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;

  if (warp < 4 * NT) {
    // ===================== epilogue warpgroup t ==============================
    const int t = warp >> 2;
    const int row = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col_g = (uint32_t)(t * 2 * W), col_d = (uint32_t)(t * 2 * W + W);
    float4* a_tile = reinterpret_cast<float4*>(a_base + t * S::kABytes);
    uint32_t ph_d = 0;
    for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
      const int64_t tile_id = (int64_t)g * NT + t;
      const int64_t r = tile_id * 128 + row;
      const bool valid = r < p.R;
      const float de = valid ? (p.dE ? p.dE[r] : 1.0f) : 0.f;
      const uint32_t* mask_base = p.mask_save;
      uint32_t mnext[W / 32];
      {
        const uint4* mp = reinterpret_cast<const uint4*>(
            mask_base + TT::mask_index(tile_id, p.nb, nsteps - 1, row));
#pragma unroll
        for (int c4 = 0; c4 < W / 128; ++c4) {
          const uint4 m = mp[c4];
          mnext[c4 * 4 + 0] = m.x; mnext[c4 * 4 + 1] = m.y;
          mnext[c4 * 4 + 2] = m.z; mnext[c4 * 4 + 3] = m.w;
        }
      }
      {
        // g_top = dE (x) we   (gradient of the residual stream entering Dense(1));
        // dwe += h_nb^T dE with h_nb from the forward's last slot
        const float4* hn = reinterpret_cast<const float4*>(
            p.xa_save + TT::slot_base(tile_id, xa_slots, 2 * p.nb));
        float4* gy_top = p.gy_save ? reinterpret_cast<float4*>(
            p.gy_save + TT::slot_base(tile_id, gy_slots, 2 * p.nb - 1)) : nullptr;
#pragma unroll 1
        for (int c32 = 0; c32 < W / 32; ++c32) {
          uint32_t v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c32 * 8 + q;
            const float4 w4 = *reinterpret_cast<const float4*>(we + c * 4);
            const float4 g4 = make_float4(de * w4.x, de * w4.y, de * w4.z, de * w4.w);
            v[q * 4 + 0] = __float_as_uint(g4.x); v[q * 4 + 1] = __float_as_uint(g4.y);
            v[q * 4 + 2] = __float_as_uint(g4.z); v[q * 4 + 3] = __float_as_uint(g4.w);
            a_tile[c * 128 + row] = tf32_rn4(g4);
            if (gy_top) gy_top[TT::chunk_f4(row, c)] = g4;   // fp32: exact bias sums, MMA truncates
            if (p.dwe) {
              const float4 h4 = hn[TT::chunk_f4(row, c)];
              const float s0 = warp_sum(de * h4.x), s1 = warp_sum(de * h4.y);
              const float s2 = warp_sum(de * h4.z), s3 = warp_sum(de * h4.w);
              if (lane == 0) {
                atomicAdd(&dwe_acc[c * 4 + 0], s0); atomicAdd(&dwe_acc[c * 4 + 1], s1);
                atomicAdd(&dwe_acc[c * 4 + 2], s2); atomicAdd(&dwe_acc[c * 4 + 3], s3);
              }
            }
          }
          tmem_st32(tmem_base + lane_off + col_g + c32 * 32, v);
        }
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&a_ready[t]);
      }
      for (int s = 0; s < nsteps; ++s) {
        const int i = nsteps - s;                     // forward layer index 2nb .. 1
        const bool odd = (i & 1) != 0;                // W1 layer: result adds onto g
        const bool last = (i == 1);
        float4* gy_out = (last || !p.gy_save) ? nullptr : reinterpret_cast<float4*>(
            p.gy_save + TT::slot_base(tile_id, gy_slots, i - 2));
        // ReLU masks of this layer (bits written by the forward); the next layer's
        // are requested now so that their latency hides behind this step
        uint32_t mbits[W / 32];
#pragma unroll
        for (int c32 = 0; c32 < W / 32; ++c32) mbits[c32] = mnext[c32];
        if (!last) {
          const uint4* mp = reinterpret_cast<const uint4*>(
              mask_base + TT::mask_index(tile_id, p.nb, i - 2, row));
#pragma unroll
          for (int c4 = 0; c4 < W / 128; ++c4) {
            const uint4 m = mp[c4];
            mnext[c4 * 4 + 0] = m.x; mnext[c4 * 4 + 1] = m.y;
            mnext[c4 * 4 + 2] = m.z; mnext[c4 * 4 + 3] = m.w;
          }
        }
        mbar_wait(&d_ready[t], ph_d, abort_flag, p.status, 500 + s);
        ph_d ^= 1;
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < W / 64; ++cb) {
          // two 32-column chunks of D (and of g on W1 layers) in flight per wait
          uint32_t v[2][32], gp[2][32];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            tmem_ld32(tmem_base + lane_off + col_d + (cb * 2 + j) * 32, v[j]);
            if (odd) tmem_ld32(tmem_base + lane_off + col_g + (cb * 2 + j) * 32, gp[j]);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c32 = cb * 2 + j;
            const uint32_t bits = mbits[c32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = c32 * 8 + q;
              float4 o;
              o.x = (bits >> (q * 4 + 0)) & 1u ? __uint_as_float(v[j][q * 4 + 0]) : 0.f;
              o.y = (bits >> (q * 4 + 1)) & 1u ? __uint_as_float(v[j][q * 4 + 1]) : 0.f;
              o.z = (bits >> (q * 4 + 2)) & 1u ? __uint_as_float(v[j][q * 4 + 2]) : 0.f;
              o.w = (bits >> (q * 4 + 3)) & 1u ? __uint_as_float(v[j][q * 4 + 3]) : 0.f;
              if (odd) {
                o.x += __uint_as_float(gp[j][q * 4 + 0]); o.y += __uint_as_float(gp[j][q * 4 + 1]);
                o.z += __uint_as_float(gp[j][q * 4 + 2]); o.w += __uint_as_float(gp[j][q * 4 + 3]);
                gp[j][q * 4 + 0] = __float_as_uint(o.x); gp[j][q * 4 + 1] = __float_as_uint(o.y);
                gp[j][q * 4 + 2] = __float_as_uint(o.z); gp[j][q * 4 + 3] = __float_as_uint(o.w);
              }
              if (!last) {
                a_tile[c * 128 + row] = tf32_rn4(o);
                if (gy_out) gy_out[TT::chunk_f4(row, c)] = o;
              } else if (valid) {
                *reinterpret_cast<float4*>(p.g0 + r * W + c * 4) = o;
              }
            }
            if (odd && !last) tmem_st32(tmem_base + lane_off + col_g + c32 * 32, gp[j]);
          }
        }
        if (odd && !last) tmem_st_wait();
        tc_fence_before();
        if (!last) {
          fence_proxy_async_smem();
          mbar_arrive(&a_ready[t]);
        }
      }
    }
  } else if (warp == 4 * NT) {
    // ===================== producer ===========================================
    if (lane == 0) {
      uint32_t q = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int s = 0; s < nsteps; ++s) {
          const float* src = p.packed + p.off_rev_l + (int64_t)s * W * W;
          for (int sl = 0; sl < S::kSlicesPerLayer; ++sl, ++q) {
            const int slot = q % kSlots;
            const uint32_t ph = (q / kSlots) & 1;
            mbar_wait(&w_empty[slot], ph ^ 1, abort_flag, p.status, 600 + s);
            mbar_arrive_expect_tx(&w_full[slot], (uint32_t)kSliceBytes);
            bulk_g2s(ring + slot * kSliceBytes, src + (int64_t)sl * (kSliceBytes / 4),
                     (uint32_t)kSliceBytes, &w_full[slot]);
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
      uint32_t ph_a[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) ph_a[t] = 0;
      for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        for (int s = 0; s < nsteps; ++s) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            mbar_wait(&a_ready[t], ph_a[t], abort_flag, p.status, 700 + s);
            ph_a[t] ^= 1;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * 2 * W + W);
            const uint64_t a_desc0 = smem_desc(smem_u32(a_base + t * S::kABytes), a_lbo, sbo);
            constexpr uint32_t a_step16 = (2 * 128 * 16) >> 4, b_step16 = (2 * W * 16) >> 4;
#pragma unroll
            for (int sl = 0; sl < S::kSlicesPerLayer; ++sl) {
              const uint32_t qq = q + sl;
              const int slot = qq % kSlots;
              if (t == 0) mbar_wait(&w_full[slot], (qq / kSlots) & 1, abort_flag, p.status, 800 + s);
              const uint64_t b_desc0 = smem_desc(smem_u32(ring + slot * kSliceBytes), b_lbo, sbo);
              mma_tf32_ss_steps<S::kStepsPerSlice>(
                  d_tmem, a_desc0 + (uint64_t)(sl * S::kStepsPerSlice * a_step16), b_desc0,
                  a_step16, b_step16, idesc, sl == 0 ? 0u : 1u);
              if (t == NT - 1) mma_commit(&w_empty[slot]);
            }
            mma_commit(&d_ready[t]);
          }
          q += S::kSlicesPerLayer;
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  if (p.dwe)
    for (int i = tid; i < W; i += S::kThreads) atomicAdd(p.dwe + i, dwe_acc[i]);
}

// dWp_act[a, w] += sum_r act[r, a] * g0[r, w]: one thread per output feature w
// streams 256 rows (coalesced g0 reads, broadcast action reads), A <= 32
// accumulators in registers, one atomic per (a, w) per block.
template <int AMAX>
__global__ void __launch_bounds__(kW) first_layer_grads_kernel(
    const float* __restrict__ act, const float* __restrict__ g0, const float* __restrict__ dE,
    int64_t R, int A, int64_t n1, float* __restrict__ dw, float* __restrict__ gsum,
    float* __restrict__ dbe) {
  // One pass over g0 [R, W]: dWp[O:] += act^T . g0 (per-thread accumulators, thread =
  // column), gsum[b] += sum of the rows of observation b (for dWp[:O] and dbp), and
  // dbe += sum(dE).  Rows of a 256-row block span at most a few observations.
  __shared__ float s_act[64 * AMAX];
  const int w = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * 256;
  const int64_t r1 = min(R, r0 + 256);
  float acc[AMAX];
#pragma unroll
  for (int a = 0; a < AMAX; ++a) acc[a] = 0.f;
  float gs = 0.f;
  int64_t cur_b = r0 / n1;
  int64_t next_edge = (cur_b + 1) * n1;
  for (int64_t rb = r0; rb < r1; rb += 64) {
    const int nr = (int)min((int64_t)64, r1 - rb);
    __syncthreads();
    for (int i = w; i < nr * AMAX; i += kW) {
      const int k = i / AMAX, a = i % AMAX;
      s_act[i] = (a < A) ? act[(rb + k) * A + a] : 0.f;
    }
    __syncthreads();
    const float* gp = g0 + rb * kW + w;
    // first row (relative to rb) of the next observation's group
    int edge = (int)min(next_edge - rb, (int64_t)1 << 20);
    const int n1i = (int)min(n1, (int64_t)1 << 20);
    for (int k0 = 0; k0 < nr; k0 += 8) {
      float g[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) g[u] = (k0 + u < nr) ? gp[(int64_t)(k0 + u) * kW] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u;
        if (k == edge) {
          atomicAdd(gsum + cur_b * kW + w, gs);
          gs = 0.f;
          ++cur_b;
          next_edge += n1;
          edge += n1i;
        }
        gs += g[u];
#pragma unroll
        for (int a = 0; a < AMAX; ++a) acc[a] = fmaf(s_act[(k < nr ? k : 0) * AMAX + a], g[u], acc[a]);
      }
    }
  }
  if (r1 > r0) atomicAdd(gsum + cur_b * kW + w, gs);
#pragma unroll
  for (int a = 0; a < AMAX; ++a)
    if (a < A) atomicAdd(dw + (int64_t)a * kW + w, acc[a]);
  if (w < 32) {            // dbe: the block's 256 dE values through one warp
    float sde = 0.f;
    for (int64_t r = r0 + w; r < r1; r += 32) sde += dE[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sde += __shfl_xor_sync(0xffffffffu, sde, o);
    if (w == 0) atomicAdd(dbe, sde);
  }
}

// The same pass over g0 in the TILE layout the block-major backward leaves it in
// ([tile][chunk (32)][128 rows][4], mlp_bwd_wgrad.cu): one block per 128-row tile, 32 rows
// at a time through shared memory (coalesced 512-byte reads; rows of 33 float4 make the
// per-column reads of thread = column conflict-free).
template <int AMAX>
__global__ void __launch_bounds__(kW) first_layer_grads_tiled_kernel(
    const float* __restrict__ act, const float* __restrict__ g0t, const float* __restrict__ dE,
    int64_t R, int A, int64_t n1, float* __restrict__ dw, float* __restrict__ gsum,
    float* __restrict__ dbe) {
  __shared__ float4 s_g[32 * 33];
  __shared__ float s_act[32 * AMAX];
  const int w = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * 128;
  const int64_t r1 = min(R, r0 + 128);
  const float4* tile = reinterpret_cast<const float4*>(g0t) + (int64_t)blockIdx.x * (128 * kW / 4);
  float acc[AMAX];
#pragma unroll
  for (int a = 0; a < AMAX; ++a) acc[a] = 0.f;
  float gs = 0.f;
  int64_t cur_b = r0 / n1;
  int64_t next_edge = (cur_b + 1) * n1;
  const float* sg = reinterpret_cast<const float*>(s_g) + (w >> 2) * (33 * 4) + (w & 3);
  // the next 32 rows travel (L2 / HBM -> registers) while the current 32 are being summed: a
  // block is a chain of four load -> barrier -> sum -> barrier rounds, and the whole grid is one
  // resident wave, so the chain's latency is the kernel's duration
  float4 pre[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = i * kW + w, c = idx >> 5, rr = idx & 31;
    pre[i] = tile[c * 128 + rr];
  }
  for (int64_t rb = r0; rb < r1; rb += 32) {
    const int nr = (int)min((int64_t)32, r1 - rb);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = i * kW + w, c = idx >> 5, rr = idx & 31;
      s_g[c * 33 + rr] = pre[i];
    }
    for (int i = w; i < nr * AMAX; i += kW) {
      const int k = i / AMAX, a = i % AMAX;
      s_act[i] = (a < A) ? act[(rb + k) * A + a] : 0.f;
    }
    __syncthreads();
    if (rb + 32 < r1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = i * kW + w, c = idx >> 5, rr = idx & 31;
        pre[i] = tile[c * 128 + (int)(rb + 32 - r0) + rr];
      }
    }
    for (int k = 0; k < nr; ++k) {
      if (rb + k == next_edge) {
        atomicAdd(gsum + cur_b * kW + w, gs);
        gs = 0.f;
        ++cur_b;
        next_edge += n1;
      }
      const float g = sg[k * 4];
      gs += g;
#pragma unroll
      for (int a = 0; a < AMAX; ++a) acc[a] = fmaf(s_act[k * AMAX + a], g, acc[a]);
    }
  }
  if (r1 > r0) atomicAdd(gsum + cur_b * kW + w, gs);
#pragma unroll
  for (int a = 0; a < AMAX; ++a)
    if (a < A) atomicAdd(dw + (int64_t)a * kW + w, acc[a]);
  if (w < 32) {            // dbe: the tile's dE values through one warp
    float sde = 0.f;
    for (int64_t r = r0 + w; r < r1; r += 32) sde += dE[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sde += __shfl_xor_sync(0xffffffffu, sde, o);
    if (w == 0) atomicAdd(dbe, sde);
  }
}

// dwe[w] += sum_r dE[r] * h_nb[r, w] over the forward's last saved slot (tile layout
// [sub-tile][chunk][32 rows][4]).  Lane = row of a sub-tile, warp = 4 of the 32
// chunks: loads are 512 contiguous bytes per warp, every lane keeps 16 partial sums
// across all the tiles its block visits, and the cross-row reduction happens once
// per block.  (Doing this inside the reverse kernel cost 128 warp reductions per
// tile in front of every chain: ~20% of that kernel.)
// hn_tiled: h_nb as the fp16-saving forward leaves it, [tile][chunk (32)][128 rows][4].
__global__ void __launch_bounds__(256) dwe_kernel(const float* __restrict__ xa_save,
                                                  const float* __restrict__ dE, int64_t R,
                                                  int num_tiles, int nb, float* __restrict__ dwe,
                                                  int hn_tiled) {
  using TT = TrainTile<kW>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const float4* hn = hn_tiled ? reinterpret_cast<const float4*>(xa_save) + (int64_t)tile * (32 * 128)
                                : reinterpret_cast<const float4*>(
                                      xa_save + TT::slot_base(tile, 2 * nb + 1, 2 * nb));
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int64_t r = (int64_t)tile * 128 + sub * 32 + lane;
      const float de = r < R ? dE[r] : 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float4 h4 = hn[hn_tiled ? (warp * 4 + cc) * 128 + sub * 32 + lane
                                      : TT::chunk_f4(sub * 32 + lane, warp * 4 + cc)];
        acc[cc * 4 + 0] = fmaf(de, h4.x, acc[cc * 4 + 0]);
        acc[cc * 4 + 1] = fmaf(de, h4.y, acc[cc * 4 + 1]);
        acc[cc * 4 + 2] = fmaf(de, h4.z, acc[cc * 4 + 2]);
        acc[cc * 4 + 3] = fmaf(de, h4.w, acc[cc * 4 + 3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(dwe + warp * 16 + i, v);
  }
}

// ---------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------
struct WgradParams {
  const float* xa_save;   // 2nb+1 slots
  const float* gy_save;   // 2nb slots
  float* grads;           // flat gradient vector (atomic accumulate)
  Layout L;
  int nb, num_tiles, splits;
  int* status;
};

struct WgradSmem {
  static constexpr int kRawStages = 3;                        // bulk-copied sub-tiles (chunk layout)
  static constexpr int kMmaStages = 2;                        // re-arranged MN-major operand images
  static constexpr int kSubBytes = TrainTile<kW>::kSubFloats * 4;   // one operand sub-tile (16 KB)
  static constexpr int kThreads = 6 * 32;
  static constexpr size_t kBytes =
      (size_t)(kRawStages + kMmaStages) * 2 * kSubBytes + 256;
};

// grid = (2nb layers, row splits).  Warp 4: bulk copies of the saved XA / GY
// sub-tiles; warps 0-3: re-arrange each sub-tile into the MN-major UMMA atoms
// (conflict-free 16-byte moves) and take the fp32 column sums of GY (bias
// gradient) on the way; warp 5: tcgen05.mma, accumulating dW in tensor memory
// across all rows of the CTA; warps 0-3 flush at the end.
__global__ void __launch_bounds__(WgradSmem::kThreads, 1) mlp_wgrad_umma_kernel(WgradParams p) {
  using S = WgradSmem;
  using TT = TrainTile<kW>;
  constexpr int W = kW;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw_base = smem;                                           // [raw stage][xa | gy]
  unsigned char* mma_base = smem + (size_t)S::kRawStages * 2 * S::kSubBytes; // [mma stage][xa | gy]
  uint64_t* bars = reinterpret_cast<uint64_t*>(mma_base + (size_t)S::kMmaStages * 2 * S::kSubBytes);
  uint64_t* raw_full = bars;                          // [kRawStages]
  uint64_t* raw_empty = raw_full + S::kRawStages;     // [kRawStages]
  uint64_t* mma_full = raw_empty + S::kRawStages;     // [kMmaStages]
  uint64_t* mma_empty = mma_full + S::kMmaStages;     // [kMmaStages]
  uint64_t* done = mma_empty + S::kMmaStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  __shared__ int s_abort;
  __shared__ float s_bias[kW];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int layer = blockIdx.x + 1;                            // forward layer index 1 .. 2nb
  const int per = (p.num_tiles + p.splits - 1) / p.splits;
  const int t0 = blockIdx.y * per;
  const int t1 = min(p.num_tiles, t0 + per);
  const int ntiles = max(0, t1 - t0);
  const int xa_slots = 2 * p.nb + 1, gy_slots = 2 * p.nb;

  if (tid == 0) {
    s_abort = 0;
    for (int s = 0; s < S::kRawStages; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 4); }
    for (int s = 0; s < S::kMmaStages; ++s) { mbar_init(&mma_full[s], 4); mbar_init(&mma_empty[s], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (tid < kW) s_bias[tid] = 0.f;
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;
  const int nsub = ntiles * TT::kSubTiles;

  if (warp == 4) {
    if (lane == 0) {
      for (int u = 0; u < nsub; ++u) {
        const int stage = u % S::kRawStages;
        const uint32_t ph = (u / S::kRawStages) & 1;
        const int64_t tile = t0 + u / TT::kSubTiles;
        const int sub = u % TT::kSubTiles;
        mbar_wait(&raw_empty[stage], ph ^ 1, abort_flag, p.status, 900);
        mbar_arrive_expect_tx(&raw_full[stage], 2u * S::kSubBytes);
        unsigned char* dst = raw_base + (size_t)stage * 2 * S::kSubBytes;
        bulk_g2s(dst, p.xa_save + TT::slot_base(tile, xa_slots, layer - 1) +
                          (int64_t)sub * TT::kSubFloats,
                 S::kSubBytes, &raw_full[stage]);
        bulk_g2s(dst + S::kSubBytes, p.gy_save + TT::slot_base(tile, gy_slots, layer - 1) +
                                         (int64_t)sub * TT::kSubFloats,
                 S::kSubBytes, &raw_full[stage]);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0 && nsub > 0) {
      // both operands MN-major (SWIZZLE_128B_BASE32B atoms, see umma_pack.cuh):
      // one K step = 8 rows = two 512-byte K atoms
      const uint32_t idesc = idesc_tf32(128, W) | kIdescAMajorMN | kIdescBMajorMN;
      for (int u = 0; u < nsub; ++u) {
        const int stage = u % S::kMmaStages;
        mbar_wait(&mma_full[stage], (u / S::kMmaStages) & 1, abort_flag, p.status, 901);
        const uint32_t xa_addr = smem_u32(mma_base + (size_t)stage * 2 * S::kSubBytes);
        const uint32_t gy_addr = xa_addr + S::kSubBytes;
        const uint64_t ad0 = smem_desc_lt(xa_addr, TT::kLbo, TT::kSbo, 1);
        const uint64_t bd0 = smem_desc_lt(gy_addr, TT::kLbo, TT::kSbo, 1);
#pragma unroll
        for (int ks = 0; ks < TT::kSubRows / 8; ++ks)
          mma_tf32_ss(tmem_base, ad0 + (uint64_t)(ks * (1024 >> 4)), bd0 + (uint64_t)(ks * (1024 >> 4)),
                      idesc, (uint32_t)(u > 0 || ks > 0));   // dW += XA^T GY
        mma_commit(&mma_empty[stage]);
      }
      mma_commit(done);
    }
    __syncwarp();
  } else if (nsub > 0) {
    // ---- warps 0-3: chunk layout -> MN-major atoms, bias sums, final flush ----
    // lane = p * 8 + r8: a quarter-warp reads 8 consecutive rows of one chunk (128
    // contiguous bytes) and writes two K atoms' lines (2-way conflicts at worst)
    const int pq = lane >> 3, r8 = lane & 7;
    const int rr = warp * 8 + r8;                              // row inside the sub-tile
    float4 bacc[W / 16];
#pragma unroll
    for (int i = 0; i < W / 16; ++i) bacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int u = 0; u < nsub; ++u) {
      const int rs = u % S::kRawStages, ms = u % S::kMmaStages;
      mbar_wait(&raw_full[rs], (u / S::kRawStages) & 1, abort_flag, p.status, 903);
      mbar_wait(&mma_empty[ms], ((u / S::kMmaStages) & 1) ^ 1, abort_flag, p.status, 904);
      const float4* raw_xa = reinterpret_cast<const float4*>(raw_base + (size_t)rs * 2 * S::kSubBytes);
      const float4* raw_gy = raw_xa + S::kSubBytes / 16;
      float4* mma_xa = reinterpret_cast<float4*>(mma_base + (size_t)ms * 2 * S::kSubBytes);
      float4* mma_gy = mma_xa + S::kSubBytes / 16;
#pragma unroll
      for (int cg = 0; cg < W / 16; ++cg) {
        const int c = cg * 4 + pq;
        const float4 x = raw_xa[c * TT::kSubRows + rr];
        const float4 g = raw_gy[c * TT::kSubRows + rr];
        mma_xa[TT::mn_f4(rr, c)] = x;
        mma_gy[TT::mn_f4(rr, c)] = g;
        bacc[cg].x += g.x; bacc[cg].y += g.y; bacc[cg].z += g.z; bacc[cg].w += g.w;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&mma_full[ms]); mbar_arrive(&raw_empty[rs]); }
    }
    // bias gradient: reduce over the 8 rows of the quarter-warp, then over warps
#pragma unroll
    for (int cg = 0; cg < W / 16; ++cg) {
      float4 v = bacc[cg];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
      }
      if (r8 == 0) {
        const int f = (cg * 4 + pq) * 4;
        atomicAdd(&s_bias[f + 0], v.x); atomicAdd(&s_bias[f + 1], v.y);
        atomicAdd(&s_bias[f + 2], v.z); atomicAdd(&s_bias[f + 3], v.w);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");            // warps 0-3 only
    mbar_wait(done, 0, abort_flag, p.status, 902);
    tc_fence_after();
    const int row = tid;                                         // input feature k
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int j = (layer - 1) >> 1;
    const bool is_w2 = (layer & 1) == 0;
    float* dw = p.grads + (is_w2 ? p.L.w2(j) : p.L.w1(j)) + (int64_t)row * W;
    float* db = p.grads + (is_w2 ? p.L.b2(j) : p.L.b1(j));
    atomicAdd(db + tid, s_bias[tid]);
#pragma unroll 1
    for (int c32 = 0; c32 < W / 32; ++c32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + c32 * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) atomicAdd(dw + c32 * 32 + q, __uint_as_float(v[q]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace

template <int W_, int NT_>
static int launch_bwd(ibc_handle* h, const BwdParams& bp, cudaStream_t s) {
  using S = BwdSmem<W_, NT_>;
  IBC_RAISE_SMEM(h, (mlp_bwd_umma_kernel<W_, NT_>), S::kBytes);    // per template instance
  const int grid = bp.num_groups < h->sm_count ? bp.num_groups : h->sm_count;
  mlp_bwd_umma_kernel<W_, NT_><<<grid, S::kThreads, S::kBytes, s>>>(bp);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// Fused reverse chain over the tiles saved by umma_forward(..., xa_save): g0 [R, W]
// = d(sum_r dE[r] * E[r]) / d(first Dense output).  gy_save / dwe are optional
// (training only).
int umma_reverse(ibc_handle* h, const float* dE, int64_t R, const float* xa_save,
                 const uint32_t* mask_save, float* gy_save, float* g0, float* dwe,
                 cudaStream_t s) {
  if (dwe && !xa_save) IBC_FAIL(h, IBC_ERR_INVALID, "umma_reverse: dwe needs the saved tiles");
  const Layout& L = h->L;
  static const bool force_smem = getenv("IBC_SMEM_OPERANDS") != nullptr;
  if (!force_smem && tmem_path_supported(L))
    return tmem_reverse(h, dE, R, xa_save, mask_save, gy_save, g0, dwe, s);
  const PackedLayout pl = make_packed_layout(L);
  BwdParams bp;
  bp.dE = dE; bp.xa_save = xa_save; bp.mask_save = mask_save; bp.gy_save = gy_save; bp.g0 = g0; bp.packed = h->packed;
  bp.off_rev_l = pl.rev_l; bp.off_we = pl.vecs + (int64_t)(2 * L.nb + 1) * L.W;
  bp.dwe = dwe; bp.R = R; bp.nb = L.nb; bp.status = h->dev_status;
  if (L.W == 128) {
    bp.num_groups = (int)((R + 255) / 256);
    return launch_bwd<128, 2>(h, bp, s);
  }
  bp.num_groups = (int)((R + 127) / 128);
  return launch_bwd<256, 1>(h, bp, s);
}

__global__ void fill_or_exp_kernel(float* __restrict__ e, float* __restrict__ de, int64_t R,
                                   int apply_exp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  if (apply_exp) { const float v = expf(e[i]); e[i] = v; de[i] = v; } else { de[i] = 1.0f; }
}

// mcmc.gradient_wrt_act (ibc/agents/mcmc.py:161-191) on tensor cores: fused
// forward (operands saved per tile), fused reverse chain, then
// dE/dact = g0 . Wp[O:]^T (K = W, N = act_dim: fp32 GEMM).
int mlp_energy_dact_umma(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                         float* E, float* dEda, bool apply_exp, cudaStream_t s) {
  const Layout& L = h->L;
  const int64_t R = B * n;
  const size_t mask_words = umma_mask_words(L, R);
  const size_t bytes = Workspace::rounded(mask_words * 4) + Workspace::rounded((size_t)R * L.W * 4) +
                       Workspace::rounded((size_t)B * L.W * 4) + Workspace::rounded((size_t)R * 4);
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  uint32_t* masks = h->ws.take<uint32_t>(mask_words);
  float* g0 = h->ws.take<float>((size_t)R * L.W);
  float* hp = h->ws.take<float>((size_t)B * L.W);
  float* dE = h->ws.take<float>(R);
  IBC_TRY(umma_forward(h, obs, B, act, n, hp, E, nullptr, masks, s));
  fill_or_exp_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(E, dE, R, apply_exp ? 1 : 0);
  IBC_LAUNCH_CHECK(h);
  IBC_TRY(umma_reverse(h, dE, R, nullptr, masks, nullptr, g0, nullptr, s));
  GemmP g;
  g.A = g0; g.lda = L.W; g.B = h->params + L.wp + (int64_t)L.O * L.W; g.ldb = L.W;
  g.C = dEda; g.ldc = L.A; g.M = (int)R; g.N = L.A; g.K = L.W;
  return gemm(h, g, false, true, s);
}

// One warp per row: dE/dact = g0[r, :] . Wp[O:]^T (lane a owns action dimension a),
// then the mcmc.langevin_step update (mcmc.py:238-262) -- replaces a GEMM launch
// and the update launch of every chain iteration.
__global__ void __launch_bounds__(256)
langevin_dact_update_kernel(const float* __restrict__ g0, const float* __restrict__ wpa, int W,
                            float* __restrict__ actions, int64_t R, int A,
                            const float* __restrict__ normals, uint64_t seed, uint64_t rng_stream,
                            float noise_scale, float grad_clip, float dclip_scale, float stepsize,
                            const float* __restrict__ mn, const float* __restrict__ mx,
                            int norm_type, float* __restrict__ grad_norms,
                            float* __restrict__ chain_actions,
                            const uint64_t* __restrict__ seed_ptr) {
  if (seed_ptr) seed = *seed_ptr;     // captured chains read the call's seed from memory
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int a = threadIdx.x & 31;
  if (r >= R) return;
  float d = 0.f;
  if (a < A) {
    const float4* g4 = reinterpret_cast<const float4*>(g0 + r * W);
    const float4* w4 = reinterpret_cast<const float4*>(wpa + (int64_t)a * W);
    for (int k = 0; k < W / 4; ++k) {
      const float4 g = g4[k], w = w4[k];
      d = fmaf(g.x, w.x, d); d = fmaf(g.y, w.y, d); d = fmaf(g.z, w.z, d); d = fmaf(g.w, w.w, d);
    }
  }
  float g = -d;                                   // de_dact = -dE/dact (mcmc.py:190)
  float nrm = (a < A) ? fabsf(g) : 0.f;
  if (norm_type == IBC_NORM_L2) nrm = nrm * nrm;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, nrm, o);
    nrm = (norm_type == IBC_NORM_INF) ? fmaxf(nrm, t) : nrm + t;
  }
  if (norm_type == IBC_NORM_L2) nrm = sqrtf(nrm);
  if (norm_type == IBC_NORM_NONE) nrm = 0.f;
  if (a == 0 && grad_norms) grad_norms[r] = nrm;
  if (a < A) {
    const int64_t i = r * A + a;
    if (grad_clip > 0.f) g = fminf(fmaxf(g, -grad_clip), grad_clip);
    const float z = normals ? normals[i] : philox_normal(seed, rng_stream, (uint64_t)i);
    g = __fadd_rn(__fmul_rn(0.5f, g), __fmul_rn(z, noise_scale));
    float dl = __fmul_rn(stepsize, g);
    const float dc = __fmul_rn(dclip_scale, __fadd_rn(mx[a], -mn[a]));
    dl = fminf(fmaxf(dl, -dc), dc);
    float v = __fadd_rn(actions[i], -dl);
    v = fminf(fmaxf(v, mn[a]), mx[a]);
    actions[i] = v;
    if (chain_actions) chain_actions[i] = v;
  }
}

// The chain itself: every launch goes to `s` (directly or under stream capture).
static int langevin_chain(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
                          const float* mn, const float* mx, const ibc_langevin_cfg& c,
                          const float* normals, uint64_t seed, const uint64_t* seed_ptr,
                          const std::vector<float>& steps, float* chain_actions,
                          float* chain_energies, float* chain_grad_norms, cudaStream_t s) {
  const Layout& L = h->L;
  const int64_t R = B * n;
  const int A = L.A;
  h->ws.reset();
  uint32_t* masks = h->ws.take<uint32_t>(umma_mask_words(L, R));
  float* g0 = h->ws.take<float>((size_t)R * L.W);
  float* hp = h->ws.take<float>((size_t)B * L.W);
  float* E = h->ws.take<float>(R);
  float* dE = h->ws.take<float>(R);
  float* nrm = h->ws.take<float>(R);
  {  // hp = obs . Wp[:O] + bp, once per chain
    GemmP g;
    g.A = obs; g.lda = L.O; g.B = h->params + L.wp; g.ldb = L.W; g.C = hp; g.ldc = L.W;
    g.bias = h->params + L.bp; g.M = (int)B; g.N = L.W; g.K = L.O;
    IBC_TRY(gemm(h, g, false, false, s));
  }
  const float dclip_scale = (float)((double)c.delta_action_clip * 0.5);
  const float* wpa = h->params + L.wp + (int64_t)L.O * L.W;
  for (int k = 0; k < c.num_iterations; ++k) {
    float* Ek = chain_energies ? chain_energies + (int64_t)k * R : E;
    float* Nk = chain_grad_norms ? chain_grad_norms + (int64_t)k * R : nrm;
    IBC_TRY(umma_forward_hp(h, act_const(actions), R, n, hp, Ek, nullptr, masks, s));
    const float* de = nullptr;
    if (c.apply_exp) {
      fill_or_exp_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(Ek, dE, R, 1);
      IBC_LAUNCH_CHECK(h);
      de = dE;
    }
    IBC_TRY(umma_reverse(h, de, R, nullptr, masks, nullptr, g0, nullptr, s));
    langevin_dact_update_kernel<<<(unsigned)((R + 7) / 8), 256, 0, s>>>(
        g0, wpa, L.W, actions, R, A, normals ? normals + (int64_t)k * R * A : nullptr, seed,
        (uint64_t)k + 1, c.noise_scale, c.grad_clip, dclip_scale, steps[k], mn, mx,
        c.grad_norm_type, Nk, chain_actions ? chain_actions + (int64_t)k * R * A : nullptr,
        seed_ptr);
    IBC_LAUNCH_CHECK(h);
  }
  return IBC_OK;
}

// mcmc.langevin_actions_given_obs on tensor cores: per iteration one fused forward
// (ReLU masks saved per tile), one fused reverse chain and one dE/dact + update kernel;
// the observation projection is hoisted out of the chain.  With Philox noise (no
// caller-supplied normals) the whole chain is replayed from a CUDA graph: the second
// call with the same shapes / config captures it, later calls stage their inputs into
// the graph's buffers and launch it (IBC_NO_GRAPHS=1 disables this).
// The observation half of the first Dense layer's gradient from the per-observation sums of g0:
// dWp[:O] += obs^T . gsum and dbp += column sums of gsum, in one small launch (thread = output
// column, a block = a slice of the observations; O <= 32 accumulators per thread).  Replaces a
// split-K SIMT GEMM launch and a column-sum launch of ~9 us each on a 512 x 128 matrix.
template <int OMAX>
__global__ void __launch_bounds__(kW) first_layer_obs_grads_kernel(
    const float* __restrict__ obs, const float* __restrict__ gsum, int64_t B, int O, int rows_per_block,
    float* __restrict__ dwp, float* __restrict__ dbp) {
  __shared__ float s_obs[OMAX];
  const int w = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * rows_per_block, b1 = min(B, b0 + rows_per_block);
  float acc[OMAX];
#pragma unroll
  for (int o = 0; o < OMAX; ++o) acc[o] = 0.f;
  float bs = 0.f;
  for (int64_t b = b0; b < b1; ++b) {
    __syncthreads();
    if (w < O) s_obs[w] = obs[b * O + w];
    __syncthreads();
    const float g = gsum[b * kW + w];
    bs += g;
#pragma unroll
    for (int o = 0; o < OMAX; ++o)
      if (o < O) acc[o] = fmaf(s_obs[o], g, acc[o]);
  }
#pragma unroll
  for (int o = 0; o < OMAX; ++o)
    if (o < O) atomicAdd(dwp + (int64_t)o * kW + w, acc[o]);
  atomicAdd(dbp + w, bs);
}

int langevin_umma(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
                  const float* mn, const float* mx, const ibc_langevin_cfg& c,
                  const float* normals, uint64_t seed, const std::vector<float>& steps,
                  float* chain_actions, float* chain_energies, float* chain_grad_norms,
                  cudaStream_t s) {
  const Layout& L = h->L;
  const int64_t R = B * n;
  const int K = c.num_iterations;
  const size_t bytes = Workspace::rounded(umma_mask_words(L, R) * 4) +
                       Workspace::rounded((size_t)R * L.W * 4) +
                       Workspace::rounded((size_t)B * L.W * 4) + 3 * Workspace::rounded((size_t)R * 4);
  IBC_CUDA(h, h->ws.reserve(bytes));
  IBC_TRY(umma_pack_params(h, s));
  static const bool no_fused = getenv("IBC_NO_FUSED_LANGEVIN") != nullptr;
  if (!no_fused && K > 0 && langevin_fused_supported(L)) {
    // width 128: the whole chain in one persistent kernel (langevin_tmem.cu)
    h->ws.reset();
    float* hp = h->ws.take<float>((size_t)B * L.W);
    float* steps_dev = h->ws.take<float>((size_t)K);
    GemmP g;   // hp = obs . Wp[:O] + bp, once per chain
    g.A = obs; g.lda = L.O; g.B = h->params + L.wp; g.ldb = L.W; g.C = hp; g.ldc = L.W;
    g.bias = h->params + L.bp; g.M = (int)B; g.N = L.W; g.K = L.O;
    IBC_TRY(gemm(h, g, false, false, s));
    IBC_CUDA(h, cudaMemcpyAsync(steps_dev, steps.data(), (size_t)K * 4, cudaMemcpyHostToDevice, s));
    if (langevin_f16_supported(L))     // act_dim <= 8: kind::f16 chain, sixteen epilogue warps
      return langevin_f16(h, hp, B, actions, n, mn, mx, c, normals, seed, nullptr, steps_dev,
                          chain_actions, chain_energies, chain_grad_norms, s);
    return langevin_fused(h, hp, B, actions, n, mn, mx, c, normals, seed, nullptr, steps_dev,
                          chain_actions, chain_energies, chain_grad_norms, s);
  }
  if (!no_fused && K > 0 && !c.apply_exp && langevin256_supported(L)) {
    // width 256, depth 2 (C1 / C3): the whole chain in one persistent kernel (langevin256.cu)
    h->ws.reset();
    float* hp = h->ws.take<float>((size_t)B * L.W);
    float* steps_dev = h->ws.take<float>((size_t)K);
    float* cvec = h->ws.take<float>((size_t)L.W + 4);
    GemmP g;   // hp = obs . Wp[:O] + bp, once per chain
    g.A = obs; g.lda = L.O; g.B = h->params + L.wp; g.ldb = L.W; g.C = hp; g.ldc = L.W;
    g.bias = h->params + L.bp; g.M = (int)B; g.N = L.W; g.K = L.O;
    IBC_TRY(gemm(h, g, false, false, s));
    IBC_CUDA(h, cudaMemcpyAsync(steps_dev, steps.data(), (size_t)K * 4, cudaMemcpyHostToDevice, s));
    return langevin256_fused(h, hp, B, actions, n, mn, mx, c, normals, seed, nullptr, steps_dev, cvec,
                             chain_actions, chain_energies, chain_grad_norms, s);
  }
  if (normals || !chain_graphs_enabled() || K < 4)
    return langevin_chain(h, obs, B, actions, n, mn, mx, c, normals, seed, nullptr, steps,
                          chain_actions, chain_energies, chain_grad_norms, s);

  // ---- graph path (chain_graph.cuh) --------------------------------------------
  return run_chain_graphed(
      h, obs, B, actions, n, mn, mx, c, seed, steps, chain_actions, chain_energies,
      chain_grad_norms, s,
      [&](const float* o, float* a, const uint64_t* seed_ptr, float* ca, float* ce, float* cn,
          cudaStream_t cs) {
        return langevin_chain(h, o, B, a, n, mn, mx, c, nullptr, 0, seed_ptr, steps, ca, ce, cn, cs);
      });
}

static int64_t umma_tiles(const Layout& L, int64_t R) {
  const int nt = (L.W == 128) ? 2 : 1;
  return ((R + 128 * nt - 1) / (128 * nt)) * nt;
}

size_t umma_save_floats(const Layout& L, int64_t R) {
  return (size_t)(umma_tiles(L, R) * (2 * L.nb + 1) * (int64_t)128 * L.W);
}

size_t umma_mask_words(const Layout& L, int64_t R) {
  return (size_t)(umma_tiles(L, R) * (2 * L.nb) * 128 * (int64_t)(L.W / 32));
}

bool umma_train_supported(const Layout& L) {
  return umma_supported(L) && L.W == kW;
}

// ImplicitBCAgent._loss + tape.gradient, tensor-core path (no gradient penalty).
int train_grads_umma(ibc_handle* h, const float* obs, int64_t B, const float* actions, int64_t N,
                     const ibc_loss_cfg& c, int64_t global_batch, float* per_example,
                     float* grad_loss_out, float* energies_out, float* grads, cudaStream_t s) {
  const Layout& L = h->L;
  using TT = TrainTile<kW>;
  const float* P = h->params;
  const int W = L.W, A = L.A, nb = L.nb;
  const int64_t n1 = N + 1, R = B * n1;
  const float gscale = 1.0f / (float)global_batch;
  const int num_groups = (int)((R + 255) / 256);
  const int64_t num_tiles = (int64_t)num_groups * kNT;
  size_t xa_floats = umma_save_floats(L, R);
  // IBC_TRAIN_SPLIT=1: round 1's three-kernel backward (tile-major reverse chain writing
  // every layer's output gradient + a separate weight-gradient kernel), kept for A/B runs
  const bool split_env = getenv("IBC_TRAIN_SPLIT") != nullptr;
  static const bool force_smem = getenv("IBC_SMEM_OPERANDS") != nullptr;
  const bool fused_bwd = !split_env && !force_smem && tmem_path_supported(L);
  size_t park_floats = 0, pw_floats = 0, pb_floats = 0;   // pw: the reduction-layout weight gradients
  // block-major backward: fp16 operand tiles (2nb slots x 32 KB per tile) + the fp32 h_nb tile
  const int64_t tiles128 = (R + 127) / 128;
  size_t xa16_bytes = 0, hn_floats = 0;
  if (fused_bwd) {
    bwd_wgrad_scratch_floats(h, R, &park_floats, &pw_floats);
    xa16_bytes = (size_t)tiles128 * (2 * nb) * 128 * kW * 2;
    hn_floats = (size_t)tiles128 * 128 * kW;
    xa_floats = 0;
  }
  const size_t gy_floats = fused_bwd ? 0 : (size_t)num_tiles * (2 * nb) * TT::kTileFloats;
  const size_t mask_words = umma_mask_words(L, R);
  const size_t bytes = Workspace::rounded(xa_floats * 4) + Workspace::rounded(gy_floats * 4) +
                       Workspace::rounded(park_floats * 4) + Workspace::rounded(pw_floats * 4) +
                       Workspace::rounded(pb_floats * 4) + Workspace::rounded(xa16_bytes) +
                       Workspace::rounded(hn_floats * 4) + 256 +
                       Workspace::rounded(mask_words * 4) + Workspace::rounded(fused_bwd ? 0 : (size_t)R * W * 4) +
                       2 * Workspace::rounded((size_t)B * W * 4) +
                       2 * Workspace::rounded((size_t)R * 4) + 2 * Workspace::rounded((size_t)B * 4);
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  float* xa = h->ws.take<float>(xa_floats);
  float* gy = h->ws.take<float>(gy_floats);
  float* gpark = h->ws.take<float>(park_floats);
  float* part_w = h->ws.take<float>(pw_floats);
  float* part_b = h->ws.take<float>(pb_floats);
  unsigned char* xa16 = h->ws.take<unsigned char>(xa16_bytes);
  float* hn = h->ws.take<float>(hn_floats);
  uint32_t* de_absmax = h->ws.take<uint32_t>(1);
  uint32_t* masks = h->ws.take<uint32_t>(mask_words);
  float* g0 = h->ws.take<float>(fused_bwd ? 0 : (size_t)R * W);
  float* hp = h->ws.take<float>((size_t)B * W);
  float* gsum = h->ws.take<float>((size_t)B * W);
  float* E = h->ws.take<float>(R);
  float* dE = h->ws.take<float>(R);
  float* gl = h->ws.take<float>(B);

  auto mark = [&](int i) { if (h->profiling) cudaEventRecord(h->ev[i], s); };
  IBC_CUDA(h, cudaMemsetAsync(grads, 0, (size_t)L.count * 4, s));
  mark(0);
  if (fused_bwd) {
    IBC_TRY(umma_pack_core(h, s));     // forward layers + fp16 transposed layers: all this path reads
    GemmP g;   // hp = obs . Wp[:O] + bp   (fp32)
    g.A = obs; g.lda = L.O; g.B = h->params + L.wp; g.ldb = L.W; g.C = hp; g.ldc = L.W;
    g.bias = h->params + L.bp; g.M = (int)B; g.N = L.W; g.K = L.O;
    IBC_TRY(gemm(h, g, false, false, s));
    IBC_TRY(tmem_forward_f16(h, actions, R, n1, hp, E, nullptr, xa16, hn, masks, s));
    IBC_CUDA(h, cudaMemsetAsync(de_absmax, 0, 4, s));
  } else {
    IBC_TRY(umma_forward(h, obs, B, actions, n1, hp, E, xa, masks, s));
  }
  mark(1);
  if (energies_out)
    IBC_CUDA(h, cudaMemcpyAsync(energies_out, E, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
  IBC_TRY(ebm_loss(h, c, E, actions, B, n1, L.A, gscale, per_example, dE, s,
                   fused_bwd ? de_absmax : nullptr));
  if (grad_loss_out) IBC_CUDA(h, cudaMemsetAsync(grad_loss_out, 0, (size_t)B * 4, s));
  (void)gl;
  // the per-example losses are final: a caller may read the loss while the backward runs
  if (!h->loss_event) IBC_CUDA(h, cudaEventCreateWithFlags(&h->loss_event, cudaEventDisableTiming));
  IBC_CUDA(h, cudaEventRecord(h->loss_event, s));
  h->loss_event_recorded = true;
  mark(2);

  IBC_RAISE_SMEM(h, mlp_wgrad_umma_kernel, WgradSmem::kBytes);
  if (fused_bwd)
    IBC_TRY(tmem_reverse_wgrad(h, dE, de_absmax, R, xa16, hn, masks, gpark, part_w, grads, s));
  else
    IBC_TRY(umma_reverse(h, dE, R, xa, masks, gy, g0, nullptr, s));
  if (!fused_bwd) {  // Dense(1) weight gradient from the forward's last slot (the block-major
                     // backward takes it in its top block)
    const int tiles = (int)((R + 127) / 128);
    const int blocks = tiles < 2 * h->sm_count ? tiles : 2 * h->sm_count;
    dwe_kernel<<<blocks, 256, 0, s>>>(fused_bwd ? hn : xa, dE, R, tiles, nb, grads + L.we, fused_bwd ? 1 : 0);
    IBC_LAUNCH_CHECK(h);
  }
  mark(3);
  if (!fused_bwd) {
    WgradParams wp;
    wp.xa_save = xa; wp.gy_save = gy; wp.grads = grads; wp.L = L; wp.nb = nb;
    // only tiles that hold rows: the forward / reverse kernels never touch a padding tile
    wp.num_tiles = (int)((R + 127) / 128);
    int splits = h->sm_count / (2 * nb);
    if (splits < 1) splits = 1;
    if (splits > wp.num_tiles) splits = wp.num_tiles;
    wp.splits = splits;
    wp.status = h->dev_status;
    dim3 grid(2 * nb, splits);
    mlp_wgrad_umma_kernel<<<grid, WgradSmem::kThreads, WgradSmem::kBytes, s>>>(wp);
    IBC_LAUNCH_CHECK(h);
  }
  mark(4);
  // Dense(1) bias and the first Dense layer (x = [obs tiled, act]) in fp32
  {  // one pass over g0: dWp[O:] = act^T . g0, per-observation sums, dbe
    IBC_CUDA(h, cudaMemsetAsync(gsum, 0, (size_t)B * W * 4, s));
    const int blocks = (int)((R + 255) / 256);
    float* dwa = grads + L.wp + (int64_t)L.O * W;
    float* dbe = grads + L.be;
    if (fused_bwd) {     // g0 in the tile layout, left in gpark by the block-major backward
      const int tblocks = (int)((R + 127) / 128);
      if (A <= 2)
        first_layer_grads_tiled_kernel<2><<<tblocks, kW, 0, s>>>(actions, gpark, dE, R, A, n1, dwa, gsum, dbe);
      else if (A <= 8)
        first_layer_grads_tiled_kernel<8><<<tblocks, kW, 0, s>>>(actions, gpark, dE, R, A, n1, dwa, gsum, dbe);
      else
        first_layer_grads_tiled_kernel<32><<<tblocks, kW, 0, s>>>(actions, gpark, dE, R, A, n1, dwa, gsum, dbe);
    } else if (A <= 2)
      first_layer_grads_kernel<2><<<blocks, kW, 0, s>>>(actions, g0, dE, R, A, n1, dwa, gsum, dbe);
    else if (A <= 8)
      first_layer_grads_kernel<8><<<blocks, kW, 0, s>>>(actions, g0, dE, R, A, n1, dwa, gsum, dbe);
    else
      first_layer_grads_kernel<32><<<blocks, kW, 0, s>>>(actions, g0, dE, R, A, n1, dwa, gsum, dbe);
    IBC_LAUNCH_CHECK(h);
  }
  if (L.O <= 32 && W == kW) {     // dWp[:O] = obs^T . sum_n g0 and dbp in one launch
    const int rpb = 4;      // 128 blocks for B = 512: the launch is latency, not work
    first_layer_obs_grads_kernel<32><<<(unsigned)((B + rpb - 1) / rpb), kW, 0, s>>>(
        obs, gsum, B, L.O, rpb, grads + L.wp, grads + L.bp);
    IBC_LAUNCH_CHECK(h);
  } else {
    GemmP g;   // dWp[:O] = obs^T . sum_n g0
    g.A = obs; g.lda = L.O; g.B = gsum; g.ldb = W; g.C = grads + L.wp; g.ldc = W;
    g.M = L.O; g.N = W; g.K = (int)B; g.k_chunk = 32; g.atomic = 1;   // many small splits: latency-bound
    IBC_TRY(gemm(h, g, true, false, s));
    IBC_TRY(colsum_add(h, gsum, W, B, W, grads + L.bp, s));
  }
  mark(5);
  if (h->profiling) h->ev_valid = true;
  return IBC_OK;
}

}  // namespace ibc
