// Wide MLPEBM stacks (width a multiple of 128 above 256: d4rl W = 512, particle "best"
// W = 2048, the PixelEBM value head's widths) on tcgen05: energy and dE/dact
// (networks/layers/resnet.py:135-153, ibc/agents/mcmc.py:161-191), one tensor-core GEMM
// kernel per Dense layer instead of the layer-fused chains of mlp_umma*.cu (a 512-wide
// tile's activations and a layer of 512 x 512 weights do not fit on one SM).
//
// Activations travel between layers in the MMA's own K-major operand layout ("T1"):
//   element (row r, col c) of [R, W]  ->  tile r/128, offset (c/4)*512 + (r%128)*4 + c%4
// so a 128-row x 32-column K-slice is 16 contiguous KB (one cp.async.bulk, like a weight
// slice) and an epilogue thread (one row) writes 16-byte chunks that are contiguous
// across the warp.  Per layer, CTA (tile, column block) accumulates a 128 x 128 output
// block in tensor memory over K-slices streamed through a 6-stage ring, then applies
// bias / ReLU mask / residual in the epilogue and writes the fp32 stream and/or the
// tf32-rounded operand of the next layer (+ ReLU mask bits for the reverse pass).
//
// Roofline: a 128 x 128 tf32 tile consumes 32 KB of operands per 256 MMA cycles (128
// B/cycle) while one SM's bulk copies from L2 deliver ~50 B/cycle (scripts/mma_rate.py),
// so this kernel is L2->shared bound at ~40 % of the tensor rate (measured: 0.36 ms for
// energy + dE/dact of the d4rl net on 4096 rows vs 1.84 ms on the fp32 FMA path).
// Grouping the issuer's waits did not help; 256-wide tiles / CTA pairs with multicast
// are the next step.
#include <stdlib.h>

#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "umma.cuh"
#include "umma_pack.cuh"

namespace ibc {

using namespace umma;

namespace {

constexpr int kStages = 6;
constexpr int kSliceBytes = 16384;         // 128 rows x 32 K-columns x 4 B
constexpr int kStageBytes = 2 * kSliceBytes;
constexpr int kThreads = 6 * 32;            // 4 epilogue warps, producer, MMA issuer

struct DenseParams {
  const float* a;          // T1 operand [tiles][K/4][128][4]
  const float* w;          // packed layer: [N/128][K/4][128][4]
  const float* bias;       // [N] or null
  const float* res;        // T1 fp32 [tiles][N/4][128][4] or null (may alias out_f)
  const uint32_t* mask_in; // [tiles][128][N/32] bits applied to the product, or null
  float* out_f;            // T1 fp32 result, or null
  float* out_op;           // T1 tf32-rounded (optionally ReLU'd) operand, or null
  uint32_t* mask_out;      // [tiles][128][N/32] bits of (result > 0), or null
  int K, N, relu;
  int* status;
};

__global__ void __launch_bounds__(kThreads, 1) dense_wide_kernel(DenseParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* ring = smem;
  float* s_bias = reinterpret_cast<float*>(ring + kStages * kStageBytes);   // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 128);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* d_ready = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 1);
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, bn = blockIdx.y;
  const int nks = p.K / 32;

  if (tid == 0) {
    s_abort = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(d_ready, 1);
    fence_mbar_init();
  }
  if (tid < 128) s_bias[tid] = p.bias ? p.bias[bn * 128 + tid] : 0.f;
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* abort_flag = &s_abort;

  if (warp < 4) {
    // ===================== epilogue: thread = output row =======================
    const int row = tid;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int64_t blk = ((int64_t)tile * (p.N / 128) + bn) * (128 * 128);   // floats
    const float4* res4 = p.res ? reinterpret_cast<const float4*>(p.res + blk) : nullptr;
    float4* of4 = p.out_f ? reinterpret_cast<float4*>(p.out_f + blk) : nullptr;
    float4* op4 = p.out_op ? reinterpret_cast<float4*>(p.out_op + blk) : nullptr;
    const int64_t mrow = ((int64_t)tile * 128 + row) * (p.N / 32) + bn * 4;
    uint4 min4 = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (p.mask_in) min4 = *reinterpret_cast<const uint4*>(p.mask_in + mrow);
    const uint32_t mbits[4] = {min4.x, min4.y, min4.z, min4.w};
    uint32_t mout[4];
    mbar_wait(d_ready, 0, abort_flag, p.status, 900);
    tc_fence_after();
    auto chunk = [&](int c32, uint32_t (&v)[32]) {
      const uint32_t bits = mbits[c32];
      uint32_t nb = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = c32 * 8 + q;                      // 16-byte chunk of the block
        float4 x = make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                               __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c * 4);
        x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
        if (!((bits >> (q * 4 + 0)) & 1u)) x.x = 0.f;
        if (!((bits >> (q * 4 + 1)) & 1u)) x.y = 0.f;
        if (!((bits >> (q * 4 + 2)) & 1u)) x.z = 0.f;
        if (!((bits >> (q * 4 + 3)) & 1u)) x.w = 0.f;
        if (res4) {
          const float4 r4 = res4[c * 128 + row];
          x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
        }
        if (of4) of4[c * 128 + row] = x;
        nb |= (x.x > 0.f ? 1u : 0u) << (q * 4 + 0);
        nb |= (x.y > 0.f ? 1u : 0u) << (q * 4 + 1);
        nb |= (x.z > 0.f ? 1u : 0u) << (q * 4 + 2);
        nb |= (x.w > 0.f ? 1u : 0u) << (q * 4 + 3);
        if (op4) {
          uint4 o;
          if (p.relu) {
            o = make_uint4(relu_tf32_bits(__float_as_uint(x.x)), relu_tf32_bits(__float_as_uint(x.y)),
                           relu_tf32_bits(__float_as_uint(x.z)), relu_tf32_bits(__float_as_uint(x.w)));
          } else {
            o = make_uint4(tf32_rn_bits(__float_as_uint(x.x)), tf32_rn_bits(__float_as_uint(x.y)),
                           tf32_rn_bits(__float_as_uint(x.z)), tf32_rn_bits(__float_as_uint(x.w)));
          }
          op4[c * 128 + row] = make_float4(__uint_as_float(o.x), __uint_as_float(o.y),
                                           __uint_as_float(o.z), __uint_as_float(o.w));
        }
      }
      mout[c32] = nb;
    };
    uint32_t va[32], vb[32];
    tmem_ld32(taddr, va);
    tmem_ld_wait(va); tmem_ld32(taddr + 32, vb); chunk(0, va);
    tmem_ld_wait(vb); tmem_ld32(taddr + 64, va); chunk(1, vb);
    tmem_ld_wait(va); tmem_ld32(taddr + 96, vb); chunk(2, va);
    tmem_ld_wait(vb); chunk(3, vb);
    if (p.mask_out)
      *reinterpret_cast<uint4*>(p.mask_out + mrow) = make_uint4(mout[0], mout[1], mout[2], mout[3]);
    tc_fence_before();
  } else if (warp == 4) {
    // ===================== producer: A and B K-slices ==========================
    const float* a_tile = p.a + (int64_t)tile * 128 * p.K;
    const float* w_blk = p.w + (int64_t)bn * 128 * p.K;
    for (int ks = 0; ks < nks; ++ks) {
      const int st = ks % kStages;
      const uint32_t ph = (ks / kStages) & 1;
      if (elect_one()) {
        mbar_wait(&empty[st], ph ^ 1, abort_flag, p.status, 910);
        mbar_arrive_expect_tx(&full[st], (uint32_t)kStageBytes);
        bulk_g2s(ring + st * kStageBytes, a_tile + (int64_t)ks * (kSliceBytes / 4),
                 (uint32_t)kSliceBytes, &full[st]);
        bulk_g2s(ring + st * kStageBytes + kSliceBytes, w_blk + (int64_t)ks * (kSliceBytes / 4),
                 (uint32_t)kSliceBytes, &full[st]);
      }
      __syncwarp();
    }
  } else {
    // ===================== MMA issuer ==========================================
    const uint32_t idesc = idesc_tf32(128, 128);
    const uint32_t lbo = 128 * 16, sbo = 128;
    constexpr uint32_t step16 = (2 * 128 * 16) >> 4;     // 8 K-elements = two 2 KB chunks
    for (int ks = 0; ks < nks; ++ks) {
      const int st = ks % kStages;
      if (elect_one()) {
        mbar_wait(&full[st], (ks / kStages) & 1, abort_flag, p.status, 920);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(ring + st * kStageBytes);
        mma_tf32_ss_steps<4>(tmem_base, smem_desc(a_addr, lbo, sbo),
                             smem_desc(a_addr + kSliceBytes, lbo, sbo), step16, step16, idesc,
                             ks == 0 ? 0u : 1u);
        mma_commit(&empty[st]);
        if (ks == nks - 1) mma_commit(d_ready);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

constexpr size_t kDenseSmem = (size_t)kStages * kStageBytes + 512 + (2 * kStages + 2) * 8 + 64;

int dense_wide(ibc_handle* h, const DenseParams& p, int64_t tiles, cudaStream_t s) {
  IBC_RAISE_SMEM(h, dense_wide_kernel, kDenseSmem);
  dim3 grid((unsigned)tiles, (unsigned)(p.N / 128));
  dense_wide_kernel<<<grid, kThreads, kDenseSmem, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// packed wide operands: 2*nb forward layers then 2*nb reverse layers, each
// [N/128 blocks][K/4][128][4] (B operand rows = output features of that product)
__global__ void pack_wide_kernel(const float* __restrict__ P, Layout L, float* __restrict__ out) {
  const int W = L.W;
  const int64_t WW = (int64_t)W * W;
  const int64_t total = 4 * (int64_t)L.nb * WW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool rev = i >= 2 * (int64_t)L.nb * WW;
    const int64_t e = rev ? i - 2 * (int64_t)L.nb * WW : i;
    const int m = (int)(e / WW);
    const int64_t r = e - (int64_t)m * WW;
    const int bn = (int)(r / ((int64_t)128 * W));
    const int64_t q = r - (int64_t)bn * 128 * W;
    const int kq = (int)(q & 3), nl = (int)((q >> 2) & 127), kc = (int)(q >> 9);
    const int k = kc * 4 + kq, n = bn * 128 + nl;
    float v;
    if (!rev) {            // forward layer m: W1_0, W2_0, W1_1, ...:  B[n][k] = Wm[k][n]
      const int j = m >> 1;
      const int64_t base = (m & 1) ? L.w2(j) : L.w1(j);
      v = P[base + (int64_t)k * W + n];
    } else {               // reverse step m: W2_{nb-1}^T, W1_{nb-1}^T, ...:  B[n][k] = Wm[n][k]
      const int j = L.nb - 1 - (m >> 1);
      const int64_t base = (m & 1) ? L.w1(j) : L.w2(j);
      v = P[base + (int64_t)n * W + k];
    }
    out[i] = tf32_rn(v);
  }
}

// row-major [R, W] -> T1 fp32 stream + ReLU'd rounded operand + mask bits (rows >= R: zeros)
__global__ void t1_from_rowmajor_kernel(const float* __restrict__ x, int64_t R, int W,
                                        float* __restrict__ hs, float* __restrict__ hop,
                                        uint32_t* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (tile, chunk, row)
  const int64_t per_tile = (int64_t)(W / 4) * 128;
  const int64_t tile = i / per_tile;
  const int64_t rem = i - tile * per_tile;
  const int c = (int)(rem >> 7), row = (int)(rem & 127);
  const int64_t r = tile * 128 + row;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < R) v = *reinterpret_cast<const float4*>(x + r * W + c * 4);
  reinterpret_cast<float4*>(hs)[i] = v;
  const uint4 o = make_uint4(relu_tf32_bits(__float_as_uint(v.x)), relu_tf32_bits(__float_as_uint(v.y)),
                             relu_tf32_bits(__float_as_uint(v.z)), relu_tf32_bits(__float_as_uint(v.w)));
  reinterpret_cast<uint4*>(hop)[i] = o;
  if (mask) {
    const uint32_t nib = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) |
                         (v.w > 0.f ? 8u : 0u);
    atomicOr(mask + r * (W / 32) + (c >> 3), nib << ((c & 7) * 4));
  }
}

// E[r] = h[r, :] . we + be over the T1 stream; optional g = dE (x) we for the reverse pass
__global__ void t1_rowdot_kernel(const float* __restrict__ hs, const float* __restrict__ we,
                                 const float* __restrict__ be, int64_t R, int W,
                                 float* __restrict__ E) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float4* t = reinterpret_cast<const float4*>(hs) + (r >> 7) * (int64_t)(W / 4) * 128 + (r & 127);
  float e = 0.f;
  for (int c = 0; c < W / 4; ++c) {
    const float4 v = t[(int64_t)c * 128];
    const float4 w = *reinterpret_cast<const float4*>(we + c * 4);
    e = fmaf(v.x, w.x, e); e = fmaf(v.y, w.y, e); e = fmaf(v.z, w.z, e); e = fmaf(v.w, w.w, e);
  }
  E[r] = e + be[0];
}

// G = dE (x) we (T1 fp32) and its rounded operand; dE null = ones
__global__ void t1_gtop_kernel(const float* __restrict__ dE, const float* __restrict__ we,
                               int64_t R, int W, float* __restrict__ g, float* __restrict__ gop) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_tile = (int64_t)(W / 4) * 128;
  const int64_t tile = i / per_tile;
  const int64_t rem = i - tile * per_tile;
  const int c = (int)(rem >> 7), row = (int)(rem & 127);
  const int64_t r = tile * 128 + row;
  const float de = r < R ? (dE ? dE[r] : 1.0f) : 0.f;
  const float4 w = *reinterpret_cast<const float4*>(we + c * 4);
  const float4 v = make_float4(de * w.x, de * w.y, de * w.z, de * w.w);
  reinterpret_cast<float4*>(g)[i] = v;
  reinterpret_cast<uint4*>(gop)[i] =
      make_uint4(tf32_rn_bits(__float_as_uint(v.x)), tf32_rn_bits(__float_as_uint(v.y)),
                 tf32_rn_bits(__float_as_uint(v.z)), tf32_rn_bits(__float_as_uint(v.w)));
}

// Energy without the last block's second product: h_L = h + a . W2 + b2 gives
// E = h . we + a . (W2 we) + (b2 . we + be), so with c2[0:W] = W2_last . we and
// c2[W] = b2_last . we + be (fp32, two row-dot launches)
//   E[r] = hs[r, :] . we + aop[r, :] . c2[0:W] + c2[W]
// over the T1 stream hs and the T1 operand aop = relu(h . W1 + b1) of the last block.
__global__ void t1_rowdot_top_kernel(const float* __restrict__ hs, const float* __restrict__ we,
                                     const float* __restrict__ aop, const float* __restrict__ c2,
                                     int64_t R, int W, float* __restrict__ E) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int64_t base = (r >> 7) * (int64_t)(W / 4) * 128 + (r & 127);
  const float4* t = reinterpret_cast<const float4*>(hs) + base;
  const float4* a = reinterpret_cast<const float4*>(aop) + base;
  float e = 0.f, f = 0.f;
  for (int c = 0; c < W / 4; ++c) {
    const float4 v = t[(int64_t)c * 128];
    const float4 w = *reinterpret_cast<const float4*>(we + c * 4);
    e = fmaf(v.x, w.x, e); e = fmaf(v.y, w.y, e); e = fmaf(v.z, w.z, e); e = fmaf(v.w, w.w, e);
    const float4 u = a[(int64_t)c * 128];
    const float4 q = *reinterpret_cast<const float4*>(c2 + c * 4);
    f = fmaf(u.x, q.x, f); f = fmaf(u.y, q.y, f); f = fmaf(u.z, q.z, f); f = fmaf(u.w, q.w, f);
  }
  E[r] = e + f + c2[W];
}

// The reverse chain's first product on the same vector: every row of G_top is dE[r] * we, so
// mask(v_last) (.) (G_top . W2_last^T) = mask (.) (dE[r] * c2): the T1 operand without a GEMM.
__global__ void t1_utop_kernel(const float* __restrict__ dE, const float* __restrict__ c2,
                               const uint32_t* __restrict__ mask, int64_t R, int W,
                               float* __restrict__ uop) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (tile, chunk, row)
  const int64_t per_tile = (int64_t)(W / 4) * 128;
  const int64_t tile = i / per_tile;
  const int64_t rem = i - tile * per_tile;
  const int c = (int)(rem >> 7), row = (int)(rem & 127);
  const int64_t r = tile * 128 + row;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (r < R) {
    const float de = dE ? dE[r] : 1.0f;
    const uint32_t nib = (mask[r * (W / 32) + (c >> 3)] >> ((c & 7) * 4)) & 15u;
    const float4 q = *reinterpret_cast<const float4*>(c2 + c * 4);
    if (nib & 1u) o.x = tf32_rn_bits(__float_as_uint(de * q.x));
    if (nib & 2u) o.y = tf32_rn_bits(__float_as_uint(de * q.y));
    if (nib & 4u) o.z = tf32_rn_bits(__float_as_uint(de * q.z));
    if (nib & 8u) o.w = tf32_rn_bits(__float_as_uint(de * q.w));
  }
  reinterpret_cast<uint4*>(uop)[i] = o;
}

// T1 -> row-major [R, W]
__global__ void t1_to_rowmajor_kernel(const float* __restrict__ t1, int64_t R, int W,
                                      float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row, chunk)
  const int64_t r = i / (W / 4);
  const int c = (int)(i - r * (W / 4));
  if (r >= R) return;
  const float4 v = reinterpret_cast<const float4*>(t1)[(r >> 7) * (int64_t)(W / 4) * 128 +
                                                         (int64_t)c * 128 + (r & 127)];
  *reinterpret_cast<float4*>(x + r * W + c * 4) = v;
}

__global__ void exp_fill_kernel(float* __restrict__ e, float* __restrict__ de, int64_t R) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) { const float v = expf(e[i]); e[i] = v; de[i] = v; }
}

}  // namespace

bool wide_supported(const Layout& L) {
  return L.W % 128 == 0 && L.W > 256 && L.nb >= 1 && L.A >= 1 && L.O >= 1;
}

static int wide_pack(ibc_handle* h, cudaStream_t s) {
  const Layout& L = h->L;
  const size_t floats = 4 * (size_t)L.nb * L.W * L.W;
  if (!h->packed) {
    IBC_CUDA(h, cudaMalloc((void**)&h->packed, floats * 4));
    h->packed_bytes = floats * 4;
  }
  pack_wide_kernel<<<592, 256, 0, s>>>(h->params, L, h->packed);
  IBC_LAUNCH_CHECK(h);
  h->packed_stale = false;
  return IBC_OK;
}

// Repack the tcgen05 operand copies if the parameters changed.  Callers that capture the wide
// path into a CUDA graph (samplers.cu::langevin) do this BEFORE the capture / replay, on their
// own stream: a pack recorded inside -- or left out of -- a captured chain would replay with
// whatever was packed at capture time.
int wide_pack_if_stale(ibc_handle* h, cudaStream_t s) {
  if (h->packed_stale) IBC_TRY(wide_pack(h, s));
  return IBC_OK;
}

// Energy (and optionally dE/dact) of the wide stack.  mcmc.gradient_wrt_act semantics
// as in mlp_energy_dact_fp32: dEda = d(sum E or sum exp E)/d act.
int mlp_energy_wide(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                    float* E, float* dEda, bool apply_exp, cudaStream_t s) {
  const Layout& L = h->L;
  const float* P = h->params;
  const int W = L.W, nb = L.nb;
  const int64_t R = B * n;
  const int64_t tiles = (R + 127) / 128;
  const size_t t1 = (size_t)tiles * 128 * W;                       // floats per T1 buffer
  const size_t mask_words = dEda ? (size_t)tiles * 128 * (W / 32) * (2 * nb) : 0;
  const size_t bytes = Workspace::rounded((size_t)B * W * 4) + Workspace::rounded((size_t)R * W * 4) +
                       3 * Workspace::rounded(t1 * 4) + Workspace::rounded(mask_words * 4) +
                       Workspace::rounded((size_t)R * 4) + Workspace::rounded((size_t)(W + 4) * 4);
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  float* c2 = h->ws.take<float>((size_t)W + 4);      // W2_last . we, then b2_last . we + be
  float* hp = h->ws.take<float>((size_t)B * W);
  float* rowm = h->ws.take<float>((size_t)R * W);     // first-layer output / g0, row-major
  float* hs = h->ws.take<float>(t1);                  // residual stream (forward), G (reverse)
  float* hop = h->ws.take<float>(t1);                 // operand of the first Dense of a block
  float* aop = h->ws.take<float>(t1);                 // operand of the second Dense of a block
  uint32_t* masks = mask_words ? h->ws.take<uint32_t>(mask_words) : nullptr;
  float* dE = h->ws.take<float>(R);
  if (h->packed_stale) IBC_TRY(wide_pack(h, s));
  // the last block's second product collapses onto the vector W2 we in both directions
  // Measured (scripts/time_langevin.py): 0.92 -> 0.85 ms for energy + dE/dact at W = 2048, D = 2,
  // nothing at W = 512, D = 8 (two of sixteen 17 us products against four small launches), so
  // it is on from W = 1024; IBC_WIDE_TOP=1 / IBC_WIDE_NO_TOP=1 force it on / off.
  static const int top_env = getenv("IBC_WIDE_TOP") ? 1 : (getenv("IBC_WIDE_NO_TOP") ? -1 : 0);
  const bool top = top_env > 0 || (top_env == 0 && W >= 1024);
  const size_t mslot = (size_t)tiles * 128 * (W / 32);
  const int64_t WW = (int64_t)W * W;

  {  // first Dense in fp32: obs half hoisted per observation (as in mlp_forward_fp32)
    GemmP g;
    g.A = obs; g.lda = L.O; g.B = P + L.wp; g.ldb = W; g.C = hp; g.ldc = W;
    g.bias = P + L.bp; g.M = (int)B; g.N = W; g.K = L.O;
    IBC_TRY(gemm(h, g, false, false, s));
    GemmP g2;
    g2.A = act; g2.lda = L.A; g2.B = P + L.wp + (int64_t)L.O * W; g2.ldb = W;
    g2.C = rowm; g2.ldc = W; g2.rowbias = hp; g2.ldrb = W; g2.rows_per_group = n;
    g2.M = (int)R; g2.N = W; g2.K = L.A;
    IBC_TRY(gemm(h, g2, false, false, s));
  }
  const int64_t t1_f4 = (int64_t)tiles * (W / 4) * 128;
  if (masks) IBC_CUDA(h, cudaMemsetAsync(masks, 0, mslot * 4, s));     // slot 0 is built with atomicOr
  t1_from_rowmajor_kernel<<<(unsigned)((t1_f4 + 255) / 256), 256, 0, s>>>(rowm, R, W, hs, hop, masks);
  IBC_LAUNCH_CHECK(h);
  for (int j = 0; j < nb; ++j) {
    DenseParams v;      // a = relu(h . W1 + b1)
    v.a = hop; v.w = h->packed + (int64_t)(2 * j) * WW; v.bias = P + L.b1(j); v.res = nullptr;
    v.mask_in = nullptr; v.out_f = nullptr; v.out_op = aop;
    v.mask_out = masks ? masks + (size_t)(2 * j + 1) * mslot : nullptr;
    v.K = W; v.N = W; v.relu = 1; v.status = h->dev_status;
    IBC_TRY(dense_wide(h, v, tiles, s));
    const bool last = (j == nb - 1);
    if (last && top) break;   // E below from hs, aop and c2: h_L is never formed
    DenseParams u;      // h += a . W2 + b2 ; next operand relu(h)
    u.a = aop; u.w = h->packed + (int64_t)(2 * j + 1) * WW; u.bias = P + L.b2(j); u.res = hs;
    u.mask_in = nullptr; u.out_f = hs; u.out_op = last ? nullptr : hop;
    u.mask_out = (masks && !last) ? masks + (size_t)(2 * j + 2) * mslot : nullptr;
    u.K = W; u.N = W; u.relu = 1; u.status = h->dev_status;
    IBC_TRY(dense_wide(h, u, tiles, s));
  }
  if (top) {
    IBC_TRY(rowdot(h, P + L.w2(nb - 1), W, P + L.we, nullptr, W, W, c2, s));
    IBC_TRY(rowdot(h, P + L.b2(nb - 1), W, P + L.we, P + L.be, 1, W, c2 + W, s));
    t1_rowdot_top_kernel<<<(unsigned)((R + 127) / 128), 128, 0, s>>>(hs, P + L.we, aop, c2, R, W, E);
  } else {
    t1_rowdot_kernel<<<(unsigned)((R + 127) / 128), 128, 0, s>>>(hs, P + L.we, P + L.be, R, W, E);
  }
  IBC_LAUNCH_CHECK(h);
  if (!dEda) return IBC_OK;

  // ---- reverse chain ----------------------------------------------------------------
  const float* de = nullptr;
  if (apply_exp) {
    exp_fill_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(E, dE, R);
    IBC_LAUNCH_CHECK(h);
    de = dE;
  }
  float* G = hs;          // the forward stream is dead: reuse its buffers
  float* gop = hop;
  float* uop = aop;
  t1_gtop_kernel<<<(unsigned)((t1_f4 + 255) / 256), 256, 0, s>>>(de, P + L.we, R, W, G, gop);
  IBC_LAUNCH_CHECK(h);
  const float* rev = h->packed + 2 * (int64_t)nb * WW;
  for (int j = nb - 1; j >= 0; --j) {
    const int m = 2 * (nb - 1 - j);
    if (top && j == nb - 1) {   // constant-row product (t1_utop_kernel)
      t1_utop_kernel<<<(unsigned)((t1_f4 + 255) / 256), 256, 0, s>>>(
          de, c2, masks + (size_t)(2 * j + 1) * mslot, R, W, uop);
      IBC_LAUNCH_CHECK(h);
    } else {
    DenseParams u;      // u = mask(v_j) (.) (g . W2_j^T)
    u.a = gop; u.w = rev + (int64_t)m * WW; u.bias = nullptr; u.res = nullptr;
    u.mask_in = masks + (size_t)(2 * j + 1) * mslot; u.out_f = nullptr; u.out_op = uop;
    u.mask_out = nullptr; u.K = W; u.N = W; u.relu = 0; u.status = h->dev_status;
    IBC_TRY(dense_wide(h, u, tiles, s));
    }
    DenseParams g;      // g += mask(h_j) (.) (u . W1_j^T)
    g.a = uop; g.w = rev + (int64_t)(m + 1) * WW; g.bias = nullptr; g.res = G;
    g.mask_in = masks + (size_t)(2 * j) * mslot; g.out_f = G; g.out_op = (j > 0) ? gop : nullptr;
    g.mask_out = nullptr; g.K = W; g.N = W; g.relu = 0; g.status = h->dev_status;
    IBC_TRY(dense_wide(h, g, tiles, s));
  }
  const int64_t rc = R * (W / 4);
  t1_to_rowmajor_kernel<<<(unsigned)((rc + 255) / 256), 256, 0, s>>>(G, R, W, rowm);
  IBC_LAUNCH_CHECK(h);
  GemmP ga;               // dE/dact = g0 . Wp[O:]^T  (N = act_dim: fp32)
  ga.A = rowm; ga.lda = W; ga.B = P + L.wp + (int64_t)L.O * W; ga.ldb = W;
  ga.C = dEda; ga.ldc = L.A; ga.M = (int)R; ga.N = L.A; ga.K = W;
  return gemm(h, ga, false, true, s);
}

}  // namespace ibc
