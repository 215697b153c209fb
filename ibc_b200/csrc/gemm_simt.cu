// fp32 FMA GEMM + small reductions (IBC_PRECISION_FP32 path of the dense
// layers: networks/layers/resnet.py:135-153 forward, its reverse chain and
// weight gradients).
#include "gemm_simt.cuh"

namespace ibc {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS = 68;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_kernel(GemmP p) {
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kc = p.k_chunk > 0 ? p.k_chunk : p.K;
  const int kbeg = blockIdx.z * kc;
  const int kend = min(p.K, kbeg + kc);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (TA) { m = e & 63; k = e >> 6; } else { k = e & 15; m = e >> 4; }
      float v = 0.f;
      if (m0 + m < p.M && k0 + k < kend) {
        const int64_t idx = TA ? (int64_t)(k0 + k) * p.lda + (m0 + m)
                               : (int64_t)(m0 + m) * p.lda + (k0 + k);
        v = p.A[idx];
        if (p.aop == AOP_RELU) v = fmaxf(v, 0.f);
        else if (p.aop == AOP_GATE) v = (p.gateA[idx] > 0.f) ? v : 0.f;
      }
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (TB) { k = e & 15; n = e >> 4; } else { n = e & 63; k = e >> 6; }
      float v = 0.f;
      if (n0 + n < p.N && k0 + k < kend) {
        const int64_t idx = TB ? (int64_t)(n0 + n) * p.ldb + (k0 + k)
                               : (int64_t)(k0 + k) * p.ldb + (n0 + n);
        v = p.B[idx];
      }
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool first = (blockIdx.z == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (first) {
        if (p.bias) v += p.bias[n];
        if (p.rowbias) v += p.rowbias[(int64_t)(m / p.rows_per_group) * p.ldrb + n];
      }
      if (p.gateC) v = (p.gateC[(int64_t)m * p.ldgc + n] > 0.f) ? v : 0.f;
      if (first && p.res) v += p.res[(int64_t)m * p.ldres + n];
      if (p.relu_out) v = fmaxf(v, 0.f);
      float* c = p.C + (int64_t)m * p.ldc + n;
      if (p.atomic) atomicAdd(c, v); else *c = v;
    }
  }
}

__global__ void rowdot_kernel(const float* __restrict__ X, int64_t ldx,
                              const float* __restrict__ v, const float* __restrict__ b,
                              int64_t R, int W, float* __restrict__ E) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float s = 0.f;
  for (int w = lane; w < W; w += 32) s = fmaf(X[r * ldx + w], v[w], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) E[r] = s + (b ? b[0] : 0.f);
}

// X[r, :] = encp[r / n, :] + sum_a act[r, a] * W[a, :]   (the action half of the value head's
// first Dense layer: K = act_dim is far too short for an MMA tile, the product is a
// bandwidth-bound rank-A update; one thread = 4 columns of one row, 16-byte stores)
__global__ void act_rows_kernel(const float* __restrict__ act, const float* __restrict__ W,
                                const float* __restrict__ encp, int64_t R, int64_t n, int A,
                                int Wv, float* __restrict__ X) {
  const int W4 = Wv / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * W4) return;
  const int c4 = (int)(i % W4);
  const int64_t r = i / W4;
  float4 v = *reinterpret_cast<const float4*>(encp + (r / n) * Wv + c4 * 4);
  for (int a = 0; a < A; ++a) {
    const float s = act[r * A + a];
    const float4 w = *reinterpret_cast<const float4*>(W + (int64_t)a * Wv + c4 * 4);
    v.x = fmaf(s, w.x, v.x); v.y = fmaf(s, w.y, v.y); v.z = fmaf(s, w.z, v.z); v.w = fmaf(s, w.w, v.w);
  }
  *reinterpret_cast<float4*>(X + r * Wv + c4 * 4) = v;
}

// Its reverse in one pass over G [B*n, Wv]: gsum[b, :] = sum of the group's rows (the
// gradient of the tiled encoding projection) and dW[a, :] += sum_r act[r, a] * G[r, :].
// grid (B, Wv / 128), block 256 = 32 column quads x 8 row lanes; A <= kMaxActRows.
__global__ void act_rows_bwd_kernel(const float* __restrict__ G, const float* __restrict__ act,
                                    int64_t n, int A, int Wv, float* __restrict__ gsum,
                                    float* __restrict__ dW) {
  __shared__ float4 part[8][32];
  const int b = blockIdx.x, c4 = blockIdx.y * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dw[kMaxActRows];
#pragma unroll
  for (int a = 0; a < kMaxActRows; ++a) dw[a] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = rl; i < n; i += 8) {
    const int64_t r = (int64_t)b * n + i;
    const float4 g = *reinterpret_cast<const float4*>(G + r * Wv + c4 * 4);
    gs.x += g.x; gs.y += g.y; gs.z += g.z; gs.w += g.w;
#pragma unroll
    for (int a = 0; a < kMaxActRows; ++a) {
      if (a < A) {
        const float s = act[r * A + a];
        dw[a].x = fmaf(s, g.x, dw[a].x); dw[a].y = fmaf(s, g.y, dw[a].y);
        dw[a].z = fmaf(s, g.z, dw[a].z); dw[a].w = fmaf(s, g.w, dw[a].w);
      }
    }
  }
  auto reduce = [&](float4 v) -> float4 {      // sum over the 8 row lanes, result in lane row 0
    part[rl][threadIdx.x & 31] = v;
    __syncthreads();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 q = part[k][threadIdx.x & 31];
        t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
      }
    }
    __syncthreads();
    return t;
  };
  const float4 gt = reduce(gs);
  if (rl == 0) *reinterpret_cast<float4*>(gsum + (int64_t)b * Wv + c4 * 4) = gt;
  for (int a = 0; a < A; ++a) {
    float4 pick = dw[0];
#pragma unroll
    for (int k = 1; k < kMaxActRows; ++k) if (k == a) pick = dw[k];
    const float4 t = reduce(pick);
    if (rl == 0) {
      float* o = dW + (int64_t)a * Wv + c4 * 4;
      atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
    }
  }
}


// float4 form of rowfill_kernel (W4 = W / 4 column quads per row)
__global__ void rowfill4_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                int64_t R, int W4, float* __restrict__ G) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * W4) return;
  const int64_t r = i / W4;
  const int w4 = (int)(i - r * W4);
  const float sc = scale ? scale[r] : 1.f;
  const float4 x = *reinterpret_cast<const float4*>(v + w4 * 4);
  *reinterpret_cast<float4*>(G + i * 4) = make_float4(sc * x.x, sc * x.y, sc * x.z, sc * x.w);
}

__global__ void rowfill_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                               int64_t R, int W, float* __restrict__ G) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * W) return;
  const int64_t r = i / W;
  const int w = (int)(i - r * W);
  G[i] = (scale ? scale[r] : 1.f) * v[w];
}

// grid (ceil(N/32), row_slices); block (32, 8)
__global__ void colsum_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int N,
                              const float* __restrict__ d, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per;
  const int64_t r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8)
      s = fmaf(X[r * ldx + n], d ? d[r] : 1.f, s);
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

__global__ void groupsum_kernel(const float* __restrict__ X, int64_t B, int64_t n, int W,
                                float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W) return;
  const int64_t b = i / W;
  const int w = (int)(i - b * W);
  const float* x = X + b * n * W + w;
  float s = 0.f;
  for (int64_t k = 0; k < n; ++k) s += x[k * W];
  out[i] = s;
}

__global__ void sum_kernel(const float* __restrict__ d, int64_t R, float* __restrict__ out) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R;
       i += (int64_t)gridDim.x * blockDim.x)
    s += d[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// C = A . B + bias for a short K (the observation half of the first Dense layer: K = obs_dim,
// once per observation): one output per thread, A's row broadcast within the warp, B read
// coalesced.  The 64 x 64 tiles of gemm_kernel put a 512 x 128 output on 16 CTAs (9 us for
// 2.6 MFLOP); this is a full wave of short threads.
__global__ void __launch_bounds__(256) gemm_short_k_kernel(GemmP p) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (int64_t)p.M * p.N) return;
  const int m = (int)(o / p.N), n = (int)(o - (int64_t)m * p.N);
  const float* a = p.A + (int64_t)m * p.lda;
  const float* b = p.B + n;
  float acc = 0.f;
  for (int k = 0; k < p.K; ++k) acc = fmaf(a[k], b[(int64_t)k * p.ldb], acc);
  if (p.bias) acc += p.bias[n];
  p.C[(int64_t)m * p.ldc + n] = acc;
}

}  // namespace

int gemm(ibc_handle* h, const GemmP& p, bool ta, bool tb, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0) return IBC_OK;
  if (p.tc && h && h->precision == IBC_PRECISION_TF32 && p.M >= 32 && p.K >= 16 &&
      gemm_tc_supported(p))
    return gemm_tc(h, p, ta, tb, s);
  if (!ta && !tb && !p.tc && p.K > 0 && p.K <= 64 && p.k_chunk == 0 && !p.atomic && !p.gateA && !p.gateC &&
      !p.rowbias && !p.res && !p.relu_out && p.aop == AOP_ID && p.alpha == 1.f && p.conv_C == 0 &&
      (int64_t)p.M * p.N <= (1 << 20)) {
    const int64_t tot = (int64_t)p.M * p.N;
    gemm_short_k_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(p);
    IBC_LAUNCH_CHECK(h);
    return IBC_OK;
  }
  const int kc = p.k_chunk > 0 ? p.k_chunk : p.K;
  const int splits = p.K > 0 ? (p.K + kc - 1) / kc : 1;
  if (splits > 1 && !p.atomic)
    IBC_FAIL(h, IBC_ERR_INVALID, "gemm: split-K requires atomic accumulation");
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
  if (grid.y > 65535u) IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "gemm: M=%d too large", p.M);
  if (ta && tb) gemm_kernel<true, true><<<grid, 256, 0, s>>>(p);
  else if (ta) gemm_kernel<true, false><<<grid, 256, 0, s>>>(p);
  else if (tb) gemm_kernel<false, true><<<grid, 256, 0, s>>>(p);
  else gemm_kernel<false, false><<<grid, 256, 0, s>>>(p);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int rowdot(ibc_handle* h, const float* X, int64_t ldx, const float* v, const float* b,
           int64_t R, int W, float* E, cudaStream_t s) {
  if (R <= 0) return IBC_OK;
  const int wpb = 8;
  rowdot_kernel<<<(unsigned)((R + wpb - 1) / wpb), wpb * 32, 0, s>>>(X, ldx, v, b, R, W, E);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int rowfill(ibc_handle* h, const float* v, const float* scale, int64_t R, int W, float* G,
            cudaStream_t s) {
  if (R <= 0) return IBC_OK;
  const int64_t tot = R * W;
  if ((W & 3) == 0 && ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    rowfill4_kernel<<<(unsigned)((tot / 4 + 255) / 256), 256, 0, s>>>(v, scale, R, W / 4, G);
  } else {
    rowfill_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(v, scale, R, W, G);
  }
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int act_rows(ibc_handle* h, const float* act, const float* Wa, const float* rowbias, int64_t R,
             int64_t n, int A, int W, float* X, cudaStream_t s) {
  if (R <= 0) return IBC_OK;
  if (W & 3) IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "act_rows: width must be a multiple of 4");
  const int64_t tot = R * (W / 4);
  act_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(act, Wa, rowbias, R, n, A, W, X);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int act_rows_bwd(ibc_handle* h, const float* G, const float* act, int64_t B, int64_t n, int A,
                 int W, float* gsum, float* dW, cudaStream_t s) {
  if (B <= 0) return IBC_OK;
  if (A > kMaxActRows || W % 128)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "act_rows_bwd: needs act_dim <= %d and width %% 128 == 0",
             kMaxActRows);
  act_rows_bwd_kernel<<<dim3((unsigned)B, (unsigned)(W / 128)), 256, 0, s>>>(G, act, n, A, W, gsum, dW);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

static int colsum_launch(ibc_handle* h, const float* X, int64_t ldx, int64_t M, int N,
                         const float* d, float* out, cudaStream_t s) {
  if (M <= 0 || N <= 0) return IBC_OK;
  int slices = (int)((M + 511) / 512);
  if (slices < 1) slices = 1;
  if (slices > 1024) slices = 1024;
  dim3 grid((N + 31) / 32, slices), block(32, 8);
  colsum_kernel<<<grid, block, 0, s>>>(X, ldx, M, N, d, out);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int colsum_add(ibc_handle* h, const float* X, int64_t ldx, int64_t M, int N, float* out,
               cudaStream_t s) {
  return colsum_launch(h, X, ldx, M, N, nullptr, out, s);
}

int wcolsum_add(ibc_handle* h, const float* X, int64_t ldx, const float* d, int64_t R, int W,
                float* out, cudaStream_t s) {
  return colsum_launch(h, X, ldx, R, W, d, out, s);
}

int groupsum(ibc_handle* h, const float* X, int64_t B, int64_t n, int W, float* out,
             cudaStream_t s) {
  if (B <= 0) return IBC_OK;
  const int64_t tot = B * W;
  groupsum_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(X, B, n, W, out);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int sum_add(ibc_handle* h, const float* d, int64_t R, float* out, cudaStream_t s) {
  if (R <= 0) return IBC_OK;
  int blocks = (int)((R + 255) / 256);
  if (blocks > 256) blocks = 256;
  sum_kernel<<<blocks, 256, 0, s>>>(d, R, out);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

}  // namespace ibc
