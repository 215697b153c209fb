// Late-fusion PixelEBM (fp32 path): image preprocessing, ConvMaxpool encoder,
// DenseResnetValue head, their backward passes and DFO inference.
//
//   networks/utils/image_prepro.py:20-43     u8 -> [0,1], RAW reshape, bilinear resize
//   networks/layers/conv_maxpool.py:20-48    4 x (conv3x3 same + ReLU + maxpool2) + GAP
//   networks/layers/dense_resnet_value.py:22-69
//   networks/pixel_ebm.py:79-125             encode / call with observation_encoding
//   ibc/agents/ibc_agent.py:191-213          late-fusion loss (encode once, tile N+1)
//
// Convolutions are im2col (per batch chunk) + a GEMM with fused bias/ReLU: the
// tcgen05 tf32 kernel of gemm_tc.cu (IBC_PRECISION_TF32, the default) or the
// fp32 FMA kernel (IBC_PRECISION_FP32).
// Flat parameter layout: conv{1..4} kernel [3,3,cin,cout] + bias, value dense0
// [256+A, Wv] + b, per block d0 [Wv,Wv/4], d1 [Wv/4,Wv/4], d2 [Wv/4,Wv] (+ biases),
// dense1 [Wv] + b.
#include <functional>

#include "gemm_simt.cuh"
#include "ops.cuh"
#include "umma_pack.cuh"

struct ibc_pixel_handle {
  ibc_handle core;          // workspaces, error string, launch accounting
  ibc_pixel_arch arch;
  const float* params = nullptr;
  // ibc_pixel_train_begin .. ibc_pixel_train_finish: the encoder pass of a training step
  // whose saved activations wait in core.ws (ibc::PendingTrain; null when none is open)
  void* pending = nullptr;
};

namespace ibc {
namespace {

constexpr int kConvCh[4] = {32, 64, 128, 256};
constexpr int kEnc = 256;

struct PixelLayout {
  int CT, Wv, Wq, nblk, A;
  int64_t conv_k[4], conv_b[4];
  int64_t d0_w, d0_b, d1_w, d1_b, count;
  int64_t blk(int j) const { return blk0 + (int64_t)j * blk_size; }
  int64_t blk0, blk_size;
  // inside a block: d0 W [Wv,Wq], b [Wq], d1 W [Wq,Wq], b [Wq], d2 W [Wq,Wv], b [Wv]
  int64_t b_d0w(int j) const { return blk(j); }
  int64_t b_d0b(int j) const { return blk(j) + (int64_t)Wv * Wq; }
  int64_t b_d1w(int j) const { return b_d0b(j) + Wq; }
  int64_t b_d1b(int j) const { return b_d1w(j) + (int64_t)Wq * Wq; }
  int64_t b_d2w(int j) const { return b_d1b(j) + Wq; }
  int64_t b_d2b(int j) const { return b_d2w(j) + (int64_t)Wq * Wv; }
};

PixelLayout make_pixel_layout(const ibc_pixel_arch& a) {
  PixelLayout L;
  L.CT = a.channels * a.seq_len; L.Wv = a.value_width; L.Wq = a.value_width / 4;
  L.nblk = a.value_blocks; L.A = a.act_dim;
  int64_t off = 0;
  int cin = L.CT;
  for (int i = 0; i < 4; ++i) {
    L.conv_k[i] = off; off += (int64_t)9 * cin * kConvCh[i];
    L.conv_b[i] = off; off += kConvCh[i];
    cin = kConvCh[i];
  }
  L.d0_w = off; off += (int64_t)(kEnc + L.A) * L.Wv;
  L.d0_b = off; off += L.Wv;
  L.blk0 = off;
  L.blk_size = (int64_t)L.Wv * L.Wq + L.Wq + (int64_t)L.Wq * L.Wq + L.Wq + (int64_t)L.Wq * L.Wv + L.Wv;
  off += L.blk_size * L.nblk;
  L.d1_w = off; off += L.Wv;
  L.d1_b = off; off += 1;
  L.count = off;
  return L;
}

// Every contraction of the pixel path may run on the tcgen05 tf32 kernel (gemm_tc.cu);
// gemm() picks it when the handle's precision is IBC_PRECISION_TF32.
int pgemm(ibc_handle* h, GemmP g, bool ta, bool tb, cudaStream_t s) {
  g.tc = 1;
  return gemm(h, g, ta, tb, s);
}

// ---- elementwise / data-movement kernels ----------------------------------------------
// out[b, y, x, c] (NHWC, c < CT): bilinear sample (half-pixel centres, no
// antialias) of the input viewed as [B, H, W, CT] -- the reference's raw
// reshape of [B, T, H, W, C] (image_prepro.py:28).
// The output has Cp >= CT channels per pixel; channels CT .. Cp-1 are zero (the tf32 path
// pads to a multiple of 4 so that the convolution's loader moves 16-byte pieces).
__global__ void preprocess_kernel(const void* __restrict__ img, int is_u8, int64_t B, int H, int W,
                                  int CT, int Cp, int th, int tw, int round,
                                  float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = B * th * tw * Cp;
  if (i >= tot) return;
  const int c = (int)(i % Cp);
  const int x = (int)((i / Cp) % tw);
  const int y = (int)((i / ((int64_t)Cp * tw)) % th);
  const int64_t b = i / ((int64_t)Cp * tw * th);
  if (c >= CT) { out[i] = 0.f; return; }
  const float sy = (float)H / (float)th, sx = (float)W / (float)tw;
  float fy = sy * ((float)y + 0.5f) - 0.5f, fx = sx * ((float)x + 0.5f) - 0.5f;
  fy = fmaxf(fy, 0.f); fx = fmaxf(fx, 0.f);
  const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const int64_t base = b * H * W * CT;
  auto at = [&](int yy, int xx) -> float {
    const int64_t idx = base + ((int64_t)yy * W + xx) * CT + c;
    return is_u8 ? (float)reinterpret_cast<const unsigned char*>(img)[idx] * (1.0f / 255.0f)
                 : reinterpret_cast<const float*>(img)[idx];
  };
  const float top = at(y0, x0) + lx * (at(y0, x1) - at(y0, x0));
  const float bot = at(y1, x0) + lx * (at(y1, x1) - at(y1, x0));
  const float v = top + ly * (bot - top);
  out[i] = round ? tf32_rn(v) : v;
}

// col[(b,y,x), (ky,kx,c)] = X[b, y+ky-1, x+kx-1, c] (zero padded), X is NHWC.
__global__ void im2col3x3_kernel(const float* __restrict__ X, int64_t nb, int H, int W, int C,
                                 float* __restrict__ col) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t K = 9 * C, tot = nb * H * W * K;
  if (i >= tot) return;
  const int j = (int)(i % K);
  const int64_t m = i / K;
  const int c = j % C, kk = j / C, ky = kk / 3, kx = kk % 3;
  const int x = (int)(m % W), y = (int)((m / W) % H);
  const int64_t b = m / ((int64_t)W * H);
  const int yy = y + ky - 1, xx = x + kx - 1;
  float v = 0.f;
  if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = X[((b * H + yy) * W + xx) * C + c];
  col[i] = v;
}

// dX[b, y, x, c] = sum_{ky,kx} dcol[(b, y-ky+1, x-kx+1), (ky,kx,c)]
__global__ void col2im3x3_kernel(const float* __restrict__ dcol, int64_t nb, int H, int W, int C,
                                 float* __restrict__ dX) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = nb * H * W * C;
  if (i >= tot) return;
  const int c = (int)(i % C);
  const int x = (int)((i / C) % W), y = (int)((i / ((int64_t)C * W)) % H);
  const int64_t b = i / ((int64_t)C * W * H);
  const int64_t K = 9 * C;
  float s = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yo = y - ky + 1, xo = x - kx + 1;
      if (yo >= 0 && yo < H && xo >= 0 && xo < W)
        s += dcol[((b * H + yo) * W + xo) * K + (ky * 3 + kx) * C + c];
    }
  dX[i] = s;
}

// Kd[(8 - tap), co, c] = tf32(K[tap, c, co]): the data gradient of a 3x3 'same' convolution is
// the convolution of dY with the taps flipped and the channel roles exchanged.
__global__ void flip_kernel3x3_kernel(const float* __restrict__ K, int Ci, int Co,
                                      float* __restrict__ Kd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * Ci * Co) return;
  const int co = i % Co, c = (i / Co) % Ci, tap = i / (Co * Ci);
  Kd[((8 - tap) * Co + co) * Ci + c] = tf32_rn(K[i]);
}

__global__ void add_rows_kernel(float* __restrict__ a, const float* __restrict__ b, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

// dEda[r, a] = sum_c G[r, c] * W[a, c]   (dE/dact from the gradient at the value head's first
// layer: N = act_dim columns, a bandwidth-bound pass over G; one warp per row, 16-byte loads)
__global__ void act_rows_dact_kernel(const float* __restrict__ G, const float* __restrict__ W,
                                     int64_t R, int A, int Wv, float* __restrict__ dEda) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float acc[kMaxActRows];
#pragma unroll
  for (int a = 0; a < kMaxActRows; ++a) acc[a] = 0.f;
  for (int c4 = lane; c4 < Wv / 4; c4 += 32) {
    const float4 g = *reinterpret_cast<const float4*>(G + r * Wv + c4 * 4);
#pragma unroll
    for (int a = 0; a < kMaxActRows; ++a) {
      if (a < A) {
        const float4 w = *reinterpret_cast<const float4*>(W + (int64_t)a * Wv + c4 * 4);
        acc[a] = fmaf(g.x, w.x, fmaf(g.y, w.y, fmaf(g.z, w.z, fmaf(g.w, w.w, acc[a]))));
      }
    }
  }
#pragma unroll
  for (int a = 0; a < kMaxActRows; ++a) {
    if (a < A) {
      float v = acc[a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) dEda[r * A + a] = v;
    }
  }
}

// E[r] = x[r, :] . w1 + relu(y1[r, :]) . c[0:Wq] + c[Wq]   (one warp per row, 16-byte loads;
// Wv % 4 == 0, Wq % 4 == 0)
__global__ void energy_top_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                  const float* __restrict__ y1, const float* __restrict__ c,
                                  int64_t R, int Wv, int Wq, float* __restrict__ E) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float acc = 0.f;
  for (int c4 = lane; c4 < Wv / 4; c4 += 32) {
    const float4 a = *reinterpret_cast<const float4*>(x + r * Wv + c4 * 4);
    const float4 w = *reinterpret_cast<const float4*>(w1 + c4 * 4);
    acc = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc))));
  }
  for (int q4 = lane; q4 < Wq / 4; q4 += 32) {
    const float4 a = *reinterpret_cast<const float4*>(y1 + r * Wq + q4 * 4);
    const float4 w = *reinterpret_cast<const float4*>(c + q4 * 4);
    acc = fmaf(fmaxf(a.x, 0.f), w.x, fmaf(fmaxf(a.y, 0.f), w.y,
          fmaf(fmaxf(a.z, 0.f), w.z, fmaf(fmaxf(a.w, 0.f), w.w, acc))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) E[r] = acc + c[Wq];
}

// out[r, q] = gate[r, q] > 0 ? c[q] : 0   (a constant row through a ReLU mask; Wq % 4 == 0)
__global__ void gate_const_rows_kernel(const float* __restrict__ c, const float* __restrict__ gate,
                                       int64_t R, int Wq, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int W4 = Wq / 4;
  if (i >= R * W4) return;
  const int q4 = (int)(i % W4);
  const float4 g = *reinterpret_cast<const float4*>(gate + i * 4);
  const float4 v = *reinterpret_cast<const float4*>(c + q4 * 4);
  float4 o;
  o.x = g.x > 0.f ? v.x : 0.f; o.y = g.y > 0.f ? v.y : 0.f;
  o.z = g.z > 0.f ? v.z : 0.f; o.w = g.w > 0.f ? v.w : 0.f;
  *reinterpret_cast<float4*>(out + i * 4) = o;
}

// out[i] = tf32(in[i]): the weights as MMA operands (rounded once per call instead of
// in every block of every product that reads them)
__global__ void round_copy_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tf32_rn(in[i]);
}

// Kp[tap, c, co] (c < Cp) = c < C ? tf32(K[tap, c, co]) : 0
__global__ void pad_kernel_channels_kernel(const float* __restrict__ K, int C, int Cp, int Co,
                                           float* __restrict__ Kp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * Cp * Co) return;
  const int co = i % Co, c = (i / Co) % Cp, tap = i / (Co * Cp);
  Kp[i] = c < C ? tf32_rn(K[(tap * C + c) * Co + co]) : 0.f;
}
// dK[tap, c, co] += dKp[tap, c, co] for c < C
__global__ void unpad_kernel_channels_add_kernel(const float* __restrict__ dKp, int C, int Cp,
                                                 int Co, float* __restrict__ dK) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C * Co) return;
  const int co = i % Co, c = (i / Co) % C, tap = i / (Co * C);
  dK[i] += dKp[(tap * Cp + c) * Co + co];
}

// round: the pooled map is only ever a tf32 MMA operand (the next convolution and its
// weight gradient), so the tf32 path rounds it here instead of in every consumer.
__global__ void maxpool2_fwd_kernel(const float* __restrict__ Y, int64_t n, int H, int W, int C,
                                    int round, float* __restrict__ P) {
  // one thread = one pooled pixel x 4 channels (C is 32 .. 256): four 16-byte loads, one store
  const int Hp = H / 2, Wp = W / 2, C4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * Hp * Wp * C4) return;
  const int c4 = (int)(i % C4);
  const int x = (int)((i / C4) % Wp), y = (int)((i / ((int64_t)C4 * Wp)) % Hp);
  const int64_t b = i / ((int64_t)C4 * Wp * Hp);
  const float* p = Y + ((b * H + 2 * y) * W + 2 * x) * C + c4 * 4;
  const int64_t row = (int64_t)W * C;
  const float4 a00 = *reinterpret_cast<const float4*>(p);
  const float4 a01 = *reinterpret_cast<const float4*>(p + C);
  const float4 a10 = *reinterpret_cast<const float4*>(p + row);
  const float4 a11 = *reinterpret_cast<const float4*>(p + row + C);
  float4 v = make_float4(fmaxf(fmaxf(a00.x, a01.x), fmaxf(a10.x, a11.x)),
                         fmaxf(fmaxf(a00.y, a01.y), fmaxf(a10.y, a11.y)),
                         fmaxf(fmaxf(a00.z, a01.z), fmaxf(a10.z, a11.z)),
                         fmaxf(fmaxf(a00.w, a01.w), fmaxf(a10.w, a11.w)));
  if (round) v = tf32_rn4(v);
  *reinterpret_cast<float4*>(P + i * 4) = v;
}

// dY = (first maximum of its 2x2 window ? dP : 0) * (Y > 0)   (MaxPool + ReLU backward)
// One thread = one 2x2 window (or unpooled edge pixel pair) x 4 channels: four 16-byte
// loads of Y, one of dP, four 16-byte stores.  C must be a multiple of 4 (32 .. 256).
__global__ void maxpool2_relu_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ dP,
                                         int64_t n, int H, int W, int C, int round,
                                         float* __restrict__ dY) {
  const int Hp = H / 2, Wp = W / 2;
  const int Hc = (H + 1) / 2, Wc = (W + 1) / 2, C4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * Hc * Wc * C4) return;
  const int c4 = (int)(i % C4);
  const int xp = (int)((i / C4) % Wc), yp = (int)((i / ((int64_t)C4 * Wc)) % Hc);
  const int64_t b = i / ((int64_t)C4 * Wc * Hc);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t base = ((b * H + 2 * yp) * W + 2 * xp) * C + c4 * 4;
  const int64_t row = (int64_t)W * C;
  const bool has_x = 2 * xp + 1 < W, has_y = 2 * yp + 1 < H;
  float4* o00 = reinterpret_cast<float4*>(dY + base);
  if (yp >= Hp || xp >= Wp) {          // pixels outside every 'valid' window get no gradient
    *o00 = zero;
    if (has_x) *reinterpret_cast<float4*>(dY + base + C) = zero;
    if (has_y) *reinterpret_cast<float4*>(dY + base + row) = zero;
    if (has_x && has_y) *reinterpret_cast<float4*>(dY + base + row + C) = zero;
    return;
  }
  const float4 w00 = *reinterpret_cast<const float4*>(Y + base);
  const float4 w01 = *reinterpret_cast<const float4*>(Y + base + C);
  const float4 w10 = *reinterpret_cast<const float4*>(Y + base + row);
  const float4 w11 = *reinterpret_cast<const float4*>(Y + base + row + C);
  float4 g = *reinterpret_cast<const float4*>(dP + ((b * Hp + yp) * Wp + xp) * C + c4 * 4);
  if (round) g = tf32_rn4(g);   // dY is a tf32 MMA operand of both gradient products
  float4 r00 = zero, r01 = zero, r10 = zero, r11 = zero;
  auto one = [](float a00, float a01, float a10, float a11, float gv, float& q00, float& q01,
                float& q10, float& q11) {
    int arg = 0; float mx = a00;
    if (a01 > mx) { mx = a01; arg = 1; }
    if (a10 > mx) { mx = a10; arg = 2; }
    if (a11 > mx) { mx = a11; arg = 3; }
    if (!(mx > 0.f)) return;            // ReLU gate of the selected pixel
    if (arg == 0) q00 = gv; else if (arg == 1) q01 = gv; else if (arg == 2) q10 = gv; else q11 = gv;
  };
  one(w00.x, w01.x, w10.x, w11.x, g.x, r00.x, r01.x, r10.x, r11.x);
  one(w00.y, w01.y, w10.y, w11.y, g.y, r00.y, r01.y, r10.y, r11.y);
  one(w00.z, w01.z, w10.z, w11.z, g.z, r00.z, r01.z, r10.z, r11.z);
  one(w00.w, w01.w, w10.w, w11.w, g.w, r00.w, r01.w, r10.w, r11.w);
  *o00 = r00;
  *reinterpret_cast<float4*>(dY + base + C) = r01;
  *reinterpret_cast<float4*>(dY + base + row) = r10;
  *reinterpret_cast<float4*>(dY + base + row + C) = r11;
}

__global__ void gap_fwd_kernel(const float* __restrict__ P, int64_t B, int HW, int C,
                               float* __restrict__ enc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = (int)(i % C);
  const int64_t b = i / C;
  float s = 0.f;
  for (int k = 0; k < HW; ++k) s += P[(b * HW + k) * C + c];
  enc[i] = s / (float)HW;
}

__global__ void gap_bwd_kernel(const float* __restrict__ dEnc, int64_t B, int HW, int C,
                               float* __restrict__ dP) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW * C) return;
  const int c = (int)(i % C);
  const int64_t b = i / ((int64_t)C * HW);
  dP[i] = dEnc[b * C + c] / (float)HW;
}

inline unsigned nblocks(int64_t n) { return (unsigned)((n + 255) / 256); }

struct EncDims { int H[5], W[5], C[5]; };   // [0] = preprocessed input, [l] = conv l output (pre-pool)

// pad = true: the preprocessed frames carry C*T rounded up to a multiple of 4 channels
EncDims enc_dims(const ibc_pixel_arch& a, bool pad = false) {
  EncDims d;
  d.H[0] = a.target_height; d.W[0] = a.target_width; d.C[0] = a.channels * a.seq_len;
  if (pad) d.C[0] = (d.C[0] + 3) / 4 * 4;
  int h = d.H[0], w = d.W[0];
  for (int l = 1; l <= 4; ++l) {
    d.H[l] = h; d.W[l] = w; d.C[l] = kConvCh[l - 1];
    h /= 2; w /= 2;
  }
  return d;
}

struct EncSave {
  float* X0 = nullptr;
  float* Y[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // conv + relu outputs
  float* P[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // pooled
  float* col = nullptr;     // im2col rows of one batch chunk (fp32 path only)
  float* wflip = nullptr;   // flipped kernel of the data-gradient convolution (tf32 path)
  float* w1pad = nullptr;   // first kernel with zero rows for the padding channels (tf32 path)
  float* dw1pad = nullptr;  // its gradient
  float* wr = nullptr;      // tf32-rounded copy of the conv parameters, same offsets (tf32 path)
  int chunk = 1;
  bool tc = false;          // implicit-GEMM convolutions on the tcgen05 kernel
};

size_t enc_bytes(const ibc_pixel_arch& a, int64_t B, int* chunk_out, bool tc) {
  const EncDims d = enc_dims(a, tc);
  size_t bytes = Workspace::rounded((size_t)B * d.H[0] * d.W[0] * d.C[0] * 4);
  size_t col_per_img = 0;
  for (int l = 1; l <= 4; ++l) {
    bytes += Workspace::rounded((size_t)B * d.H[l] * d.W[l] * d.C[l] * 4);
    bytes += Workspace::rounded((size_t)B * (d.H[l] / 2) * (d.W[l] / 2) * d.C[l] * 4);
    const size_t c = (size_t)d.H[l] * d.W[l] * 9 * d.C[l - 1] * 4;
    if (c > col_per_img) col_per_img = c;
  }
  int chunk = (int)((size_t)(256u << 20) / col_per_img);
  if (chunk < 1) chunk = 1;
  if (chunk > B) chunk = (int)B;
  // keep the GEMM's grid.y (M / 64) under 65535
  while (chunk > 1 && (int64_t)chunk * d.H[1] * d.W[1] / 64 > 60000) --chunk;
  *chunk_out = chunk;
  if (!tc) bytes += Workspace::rounded((size_t)chunk * col_per_img);
  bytes += Workspace::rounded((size_t)9 * kConvCh[3] * kConvCh[2] * 4);
  bytes += 2 * Workspace::rounded((size_t)9 * d.C[0] * kConvCh[0] * 4);
  if (tc) bytes += Workspace::rounded((size_t)make_pixel_layout(a).d0_w * 4);
  return bytes;
}

void carve_enc(ibc_handle* h, const ibc_pixel_arch& a, int64_t B, int chunk, bool tc, EncSave* s) {
  const EncDims d = enc_dims(a, tc);
  s->chunk = chunk;
  s->tc = tc;
  s->X0 = h->ws.take<float>((size_t)B * d.H[0] * d.W[0] * d.C[0]);
  size_t col_per_img = 0;
  for (int l = 1; l <= 4; ++l) {
    s->Y[l] = h->ws.take<float>((size_t)B * d.H[l] * d.W[l] * d.C[l]);
    s->P[l] = h->ws.take<float>((size_t)B * (d.H[l] / 2) * (d.W[l] / 2) * d.C[l]);
    const size_t c = (size_t)d.H[l] * d.W[l] * 9 * d.C[l - 1];
    if (c > col_per_img) col_per_img = c;
  }
  if (!tc) s->col = h->ws.take<float>((size_t)chunk * col_per_img);
  s->wflip = h->ws.take<float>((size_t)9 * kConvCh[3] * kConvCh[2]);
  s->w1pad = h->ws.take<float>((size_t)9 * d.C[0] * kConvCh[0]);
  s->dw1pad = h->ws.take<float>((size_t)9 * d.C[0] * kConvCh[0]);
  if (tc) s->wr = h->ws.take<float>((size_t)make_pixel_layout(a).d0_w);
}

int encode_fwd(ibc_pixel_handle* ph, const PixelLayout& L, const void* images, int is_u8,
               int64_t B, const EncSave& sv, float* enc, cudaStream_t s) {
  ibc_handle* h = &ph->core;
  const ibc_pixel_arch& a = ph->arch;
  const EncDims d = enc_dims(a, sv.tc);
  const float* P = ph->params;
  const int CT = a.channels * a.seq_len;
  preprocess_kernel<<<nblocks((int64_t)B * d.H[0] * d.W[0] * d.C[0]), 256, 0, s>>>(
      images, is_u8, B, a.height, a.width, CT, d.C[0], d.H[0], d.W[0], sv.tc ? 1 : 0, sv.X0);
  IBC_LAUNCH_CHECK(h);
  if (sv.tc) {
    round_copy_kernel<<<nblocks(L.d0_w), 256, 0, s>>>(P, L.d0_w, sv.wr);
    IBC_LAUNCH_CHECK(h);
    pad_kernel_channels_kernel<<<nblocks(9 * d.C[0] * kConvCh[0]), 256, 0, s>>>(
        P + L.conv_k[0], CT, d.C[0], kConvCh[0], sv.w1pad);
    IBC_LAUNCH_CHECK(h);
  }
  const float* in = sv.X0;
  for (int l = 1; l <= 4; ++l) {
    const int H = d.H[l], W = d.W[l], Ci = d.C[l - 1], Co = d.C[l];
    if (sv.tc) {   // implicit GEMM: the loader gathers the im2col rows from the NHWC input
      GemmP g;     // Y = relu(im2col(in) . K + b)
      g.A = in; g.conv_H = H; g.conv_W = W; g.conv_C = Ci; g.lda = 9 * Ci;
      g.B = (l == 1) ? sv.w1pad : sv.wr + L.conv_k[l - 1]; g.ldb = Co; g.C = sv.Y[l]; g.ldc = Co;
      g.bias = P + L.conv_b[l - 1]; g.M = (int)(B * H * W); g.N = Co; g.K = 9 * Ci;
      g.relu_out = 1; g.a_tf32 = 1; g.b_tf32 = 1;   // X0 / P[l-1] / weights are rounded
      IBC_TRY(gemm_tc(h, g, false, false, s));
    }
    for (int64_t b0 = 0; !sv.tc && b0 < B; b0 += sv.chunk) {
      const int64_t nb = std::min<int64_t>(sv.chunk, B - b0);
      const int64_t M = nb * H * W;
      im2col3x3_kernel<<<nblocks(M * 9 * Ci), 256, 0, s>>>(in + b0 * H * W * Ci, nb, H, W, Ci, sv.col);
      IBC_LAUNCH_CHECK(h);
      GemmP g;   // Y = relu(col . K + b)
      g.A = sv.col; g.lda = 9 * Ci; g.B = P + L.conv_k[l - 1]; g.ldb = Co;
      g.C = sv.Y[l] + b0 * H * W * Co; g.ldc = Co; g.bias = P + L.conv_b[l - 1];
      g.M = (int)M; g.N = Co; g.K = 9 * Ci; g.relu_out = 1;
      IBC_TRY(pgemm(h, g, false, false, s));
    }
    maxpool2_fwd_kernel<<<nblocks((int64_t)B * (H / 2) * (W / 2) * (Co / 4)), 256, 0, s>>>(
        sv.Y[l], B, H, W, Co, (sv.tc && l < 4) ? 1 : 0, sv.P[l]);
    IBC_LAUNCH_CHECK(h);
    in = sv.P[l];
  }
  gap_fwd_kernel<<<nblocks(B * kEnc), 256, 0, s>>>(sv.P[4], B, (d.H[4] / 2) * (d.W[4] / 2), kEnc, enc);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// Encoder backward given dEnc [B, 256]; accumulates conv kernel / bias gradients.
int encode_bwd(ibc_pixel_handle* ph, const PixelLayout& L, int64_t B, const EncSave& sv,
               const float* dEnc, float* grads, float* dbufA, float* dbufB, cudaStream_t s) {
  ibc_handle* h = &ph->core;
  const EncDims d = enc_dims(ph->arch, sv.tc);
  const float* P = ph->params;
  const int CT = ph->arch.channels * ph->arch.seq_len;
  // dP4
  float* dP = dbufA;
  gap_bwd_kernel<<<nblocks(B * (d.H[4] / 2) * (d.W[4] / 2) * kEnc), 256, 0, s>>>(
      dEnc, B, (d.H[4] / 2) * (d.W[4] / 2), kEnc, dP);
  IBC_LAUNCH_CHECK(h);
  float* dY = dbufB;
  for (int l = 4; l >= 1; --l) {
    const int H = d.H[l], W = d.W[l], Ci = d.C[l - 1], Co = d.C[l];
    maxpool2_relu_bwd_kernel<<<nblocks((int64_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (Co / 4)),
                               256, 0, s>>>(sv.Y[l], dP, B, H, W, Co, sv.tc ? 1 : 0, dY);
    IBC_LAUNCH_CHECK(h);
    IBC_TRY(colsum_add(h, dY, Co, (int64_t)B * H * W, Co, grads + L.conv_b[l - 1], s));
    const float* in = (l == 1) ? sv.X0 : sv.P[l - 1];
    float* dIn = dP;   // overwritten with the gradient of this layer's input (l > 1)
    if (sv.tc) {
      GemmP gk;   // dK += im2col(in)^T . dY   (rows split over CTAs, atomics)
      gk.A = in; gk.conv_H = H; gk.conv_W = W; gk.conv_C = Ci; gk.lda = 9 * Ci;
      gk.B = dY; gk.ldb = Co; gk.C = grads + L.conv_k[l - 1]; gk.ldc = Co;
      gk.M = 9 * Ci; gk.N = Co; gk.K = (int)(B * H * W); gk.atomic = 1;
      gk.a_tf32 = 1; gk.b_tf32 = 1;
      if (l == 1 && Ci != CT) {
        gk.C = sv.dw1pad;
        IBC_CUDA(h, cudaMemsetAsync(sv.dw1pad, 0, (size_t)9 * Ci * Co * 4, s));
      }
      IBC_TRY(gemm_tc(h, gk, true, false, s));
      if (l == 1 && Ci != CT) {
        unpad_kernel_channels_add_kernel<<<nblocks(9 * CT * Co), 256, 0, s>>>(
            sv.dw1pad, CT, Ci, Co, grads + L.conv_k[0]);
        IBC_LAUNCH_CHECK(h);
      }
      if (l > 1) {
        flip_kernel3x3_kernel<<<nblocks(9 * Ci * Co), 256, 0, s>>>(P + L.conv_k[l - 1], Ci, Co,
                                                                  sv.wflip);
        IBC_LAUNCH_CHECK(h);
        GemmP gd;   // dIn = im2col(dY) . Kflip
        gd.A = dY; gd.conv_H = H; gd.conv_W = W; gd.conv_C = Co; gd.lda = 9 * Co;
        gd.B = sv.wflip; gd.ldb = Ci; gd.C = dIn; gd.ldc = Ci;
        gd.M = (int)(B * H * W); gd.N = Ci; gd.K = 9 * Co; gd.a_tf32 = 1; gd.b_tf32 = 1;
        IBC_TRY(gemm_tc(h, gd, false, false, s));
      }
    }
    for (int64_t b0 = 0; !sv.tc && b0 < B; b0 += sv.chunk) {
      const int64_t nb = std::min<int64_t>(sv.chunk, B - b0);
      const int64_t M = nb * H * W;
      im2col3x3_kernel<<<nblocks(M * 9 * Ci), 256, 0, s>>>(in + b0 * H * W * Ci, nb, H, W, Ci, sv.col);
      IBC_LAUNCH_CHECK(h);
      GemmP gk;   // dK += col^T . dY
      gk.A = sv.col; gk.lda = 9 * Ci; gk.B = dY + b0 * H * W * Co; gk.ldb = Co;
      gk.C = grads + L.conv_k[l - 1]; gk.ldc = Co; gk.M = 9 * Ci; gk.N = Co; gk.K = (int)M;
      gk.k_chunk = 2048; gk.atomic = 1;
      IBC_TRY(pgemm(h, gk, true, false, s));
      if (l > 1) {
        GemmP gd;   // dcol = dY . K^T
        gd.A = dY + b0 * H * W * Co; gd.lda = Co; gd.B = P + L.conv_k[l - 1]; gd.ldb = Co;
        gd.C = sv.col; gd.ldc = 9 * Ci; gd.M = (int)M; gd.N = 9 * Ci; gd.K = Co;
        IBC_TRY(pgemm(h, gd, false, true, s));
        col2im3x3_kernel<<<nblocks(M * Ci), 256, 0, s>>>(sv.col, nb, H, W, Ci,
                                                          dIn + b0 * H * W * Ci);
        IBC_LAUNCH_CHECK(h);
      }
    }
  }
  return IBC_OK;
}

// ---- value head --------------------------------------------------------------------------------
struct ValSave {
  float* encp = nullptr;   // [B, Wv]
  float* X = nullptr;      // [(nblk+1), R, Wv]
  float* Y0 = nullptr;     // [nblk, R, Wq]
  float* Y1 = nullptr;     // [nblk, R, Wq]
  float* wr = nullptr;     // [param count] tf32-rounded copy of the value parameters at the flat
                           // vector's offsets (filled by value_fwd on the tf32 path)
  // Energy-only evaluation (samplers): E = x_L . w1 + b1 with x_L = x + relu(y1) . D2 + b2 is
  // x . w1 + relu(y1) . (D2 w1) + (b2 . w1 + b1) -- the last block's widest product collapses
  // into the vector ctop[0:Wq] = D2_last . w1 and the scalar ctop[Wq] = b2_last . w1 + b1
  // (fp32, value_prepare).  `top` says whether this evaluation uses them (x_L is then NOT
  // formed; the training pass needs x_L for dw1 and keeps the product).
  float* ctop = nullptr;   // [Wq + 4]
  bool top = false;
  int64_t R = 0;
};

inline bool top_shortcut_ok(const PixelLayout& L) { return L.nblk >= 1 && (L.Wq & 3) == 0; }

size_t val_bytes(const PixelLayout& L, int64_t B, int64_t R) {
  const int nb = L.nblk > 0 ? L.nblk : 1;
  return Workspace::rounded((size_t)B * L.Wv * 4) +
         Workspace::rounded((size_t)(L.nblk + 1) * R * L.Wv * 4) +
         2 * Workspace::rounded((size_t)nb * R * L.Wq * 4) +
         Workspace::rounded((size_t)L.count * 4) + Workspace::rounded((size_t)(L.Wq + 4) * 4);
}

// The sampler entry points (energy, DFO, dE/dact, Langevin) carve from core.ws_outer, the
// training pass from core.ws: a chain may run between ibc_pixel_train_begin and
// ibc_pixel_train_finish without touching the encoder activations saved in core.ws.
void carve_val(Workspace& w, const PixelLayout& L, int64_t B, int64_t R, ValSave* v) {
  const int nb = L.nblk > 0 ? L.nblk : 1;
  v->R = R;
  v->encp = w.take<float>((size_t)B * L.Wv);
  v->X = w.take<float>((size_t)(L.nblk + 1) * R * L.Wv);
  v->Y0 = w.take<float>((size_t)nb * R * L.Wq);
  v->Y1 = w.take<float>((size_t)nb * R * L.Wq);
  v->wr = w.take<float>((size_t)L.count);
  v->ctop = w.take<float>((size_t)L.Wq + 4);
}

// The part of the value head that does not depend on the actions: the tf32-rounded copy of
// the block weights and encp = enc . W0[:256] + b0 for enc [B, 256] (un-tiled).  A Langevin
// chain on a fixed encoding runs it once for all of its iterations.
int value_prepare(ibc_pixel_handle* ph, const PixelLayout& L, const float* enc, int64_t B,
                  const ValSave& v, cudaStream_t s) {
  ibc_handle* h = &ph->core;
  const float* P = ph->params;
  const int Wv = L.Wv;
  if (h->precision == IBC_PRECISION_TF32) {
    round_copy_kernel<<<nblocks(L.count - L.d0_w), 256, 0, s>>>(P + L.d0_w, L.count - L.d0_w,
                                                                 v.wr + L.d0_w);
    IBC_LAUNCH_CHECK(h);
  }
  if (v.top) {
    const int jl = L.nblk - 1;
    IBC_TRY(rowdot(h, P + L.b_d2w(jl), Wv, P + L.d1_w, nullptr, L.Wq, Wv, v.ctop, s));
    IBC_TRY(rowdot(h, P + L.b_d2b(jl), Wv, P + L.d1_w, P + L.d1_b, 1, Wv, v.ctop + L.Wq, s));
  }
  GemmP g;  // encp = enc . W0[:256] + b0
  g.A = enc; g.lda = kEnc; g.B = P + L.d0_w; g.ldb = Wv; g.C = v.encp; g.ldc = Wv;
  g.bias = P + L.d0_b; g.M = (int)B; g.N = Wv; g.K = kEnc;
  return pgemm(h, g, false, false, s);
}

// E[r] for act [B*n, A] on the encoding value_prepare has projected into `v`
// (pixel_ebm.py:113-125)
int value_rows(ibc_pixel_handle* ph, const PixelLayout& L, int64_t B, const float* act, int64_t n,
               const ValSave& v, float* E, cudaStream_t s) {
  ibc_handle* h = &ph->core;
  const float* P = ph->params;
  const int Wv = L.Wv, Wq = L.Wq;
  const int64_t R = B * n;
  const bool tc = h->precision == IBC_PRECISION_TF32;
  const float* PB = tc ? v.wr : P;          // B operands of the block products
  if (tc) {   // x0 = act . W0[256:] + encp[row / n]
    IBC_TRY(act_rows(h, act, P + L.d0_w + (int64_t)kEnc * Wv, v.encp, R, n, L.A, Wv, v.X, s));
  } else {
    GemmP g;
    g.A = act; g.lda = L.A; g.B = P + L.d0_w + (int64_t)kEnc * Wv; g.ldb = Wv;
    g.C = v.X; g.ldc = Wv; g.rowbias = v.encp; g.ldrb = Wv; g.rows_per_group = n;
    g.M = (int)R; g.N = Wv; g.K = L.A;
    IBC_TRY(pgemm(h, g, false, false, s));
  }
  for (int j = 0; j < L.nblk; ++j) {
    const float* x = v.X + (int64_t)j * R * Wv;
    float* xn = v.X + (int64_t)(j + 1) * R * Wv;
    float* y0 = v.Y0 + (int64_t)j * R * Wq;
    float* y1 = v.Y1 + (int64_t)j * R * Wq;
    GemmP g0;  // y0 = relu(x) . D0 + b
    g0.A = x; g0.lda = Wv; g0.aop = AOP_RELU; g0.B = PB + L.b_d0w(j); g0.ldb = Wq; g0.b_tf32 = tc;
    g0.C = y0; g0.ldc = Wq; g0.bias = P + L.b_d0b(j); g0.M = (int)R; g0.N = Wq; g0.K = Wv;
    IBC_TRY(pgemm(h, g0, false, false, s));
    GemmP g1;  // y1 = relu(y0) . D1 + b
    g1.A = y0; g1.lda = Wq; g1.aop = AOP_RELU; g1.B = PB + L.b_d1w(j); g1.ldb = Wq; g1.b_tf32 = tc;
    g1.C = y1; g1.ldc = Wq; g1.bias = P + L.b_d1b(j); g1.M = (int)R; g1.N = Wq; g1.K = Wq;
    IBC_TRY(pgemm(h, g1, false, false, s));
    if (v.top && j == L.nblk - 1) {   // E = x . w1 + relu(y1) . ctop + ctop[Wq]
      energy_top_kernel<<<(unsigned)((R + 7) / 8), 256, 0, s>>>(x, P + L.d1_w, y1, v.ctop, R, Wv,
                                                                Wq, E);
      IBC_LAUNCH_CHECK(h);
      return IBC_OK;
    }
    GemmP g2;  // x' = x + relu(y1) . D2 + b
    g2.A = y1; g2.lda = Wq; g2.aop = AOP_RELU; g2.B = PB + L.b_d2w(j); g2.ldb = Wv; g2.b_tf32 = tc;
    g2.C = xn; g2.ldc = Wv; g2.bias = P + L.b_d2b(j); g2.res = x; g2.ldres = Wv;
    g2.M = (int)R; g2.N = Wv; g2.K = Wq;
    IBC_TRY(pgemm(h, g2, false, false, s));
  }
  IBC_TRY(rowdot(h, v.X + (int64_t)L.nblk * R * Wv, Wv, P + L.d1_w, P + L.d1_b, R, Wv, E, s));
  return IBC_OK;
}

// E[r] for enc [B, 256] (un-tiled) and act [B*n, A]   (pixel_ebm.py:113-125)
int value_fwd(ibc_pixel_handle* ph, const PixelLayout& L, const float* enc, int64_t B,
              const float* act, int64_t n, const ValSave& v, float* E, cudaStream_t s) {
  IBC_TRY(value_prepare(ph, L, enc, B, v, s));
  return value_rows(ph, L, B, act, n, v, E, s);
}

static int dwg(ibc_handle* h, const float* X, int64_t ldx, int aop, const float* Y, int64_t ldy,
               int64_t R, int Kin, int Nout, float* dW, cudaStream_t s) {
  GemmP g;
  g.A = X; g.lda = ldx; g.aop = aop; g.B = Y; g.ldb = ldy; g.C = dW; g.ldc = Nout;
  g.M = Kin; g.N = Nout; g.K = (int)R; g.k_chunk = 1024; g.atomic = 1;
  return pgemm(h, g, true, false, s);
}

// Value-head backward: parameter gradients (accumulated) and dEnc [B, 256].
int value_bwd(ibc_pixel_handle* ph, const PixelLayout& L, const float* enc, int64_t B,
              const float* act, int64_t n, const ValSave& v, const float* dE, float* grads,
              float* dEnc, cudaStream_t s) {
  ibc_handle* h = &ph->core;
  const float* P = ph->params;
  const int Wv = L.Wv, Wq = L.Wq;
  const int64_t R = B * n;
  float* Ga = h->ws.take<float>((size_t)R * Wv);
  float* Gb = h->ws.take<float>((size_t)R * Wv);
  float* g1b = h->ws.take<float>((size_t)R * Wq);
  float* g0b = h->ws.take<float>((size_t)R * Wq);
  float* gsum = h->ws.take<float>((size_t)B * Wv);
  const bool tc = h->precision == IBC_PRECISION_TF32;
  const float* PB = tc ? v.wr : P;   // rounded by value_fwd earlier in the same call
  const float* xl = v.X + (int64_t)L.nblk * R * Wv;
  IBC_TRY(wcolsum_add(h, xl, Wv, dE, R, Wv, grads + L.d1_w, s));
  IBC_TRY(sum_add(h, dE, R, grads + L.d1_b, s));
  IBC_TRY(rowfill(h, P + L.d1_w, dE, R, Wv, Ga, s));
  float* G = Ga;
  for (int j = L.nblk - 1; j >= 0; --j) {
    float* Gn = (G == Ga) ? Gb : Ga;
    const float* x = v.X + (int64_t)j * R * Wv;
    const float* y0 = v.Y0 + (int64_t)j * R * Wq;
    const float* y1 = v.Y1 + (int64_t)j * R * Wq;
    IBC_TRY(dwg(h, y1, Wq, AOP_RELU, G, Wv, R, Wq, Wv, grads + L.b_d2w(j), s));
    IBC_TRY(colsum_add(h, G, Wv, R, Wv, grads + L.b_d2b(j), s));
    GemmP a1;  // g1 = (G . D2^T) (.) (y1 > 0)
    a1.A = G; a1.lda = Wv; a1.B = PB + L.b_d2w(j); a1.ldb = Wv; a1.C = g1b; a1.ldc = Wq;
    a1.b_tf32 = tc;
    a1.gateC = y1; a1.ldgc = Wq; a1.M = (int)R; a1.N = Wq; a1.K = Wv;
    IBC_TRY(pgemm(h, a1, false, true, s));
    IBC_TRY(dwg(h, y0, Wq, AOP_RELU, g1b, Wq, R, Wq, Wq, grads + L.b_d1w(j), s));
    IBC_TRY(colsum_add(h, g1b, Wq, R, Wq, grads + L.b_d1b(j), s));
    GemmP a0;  // g0 = (g1 . D1^T) (.) (y0 > 0)
    a0.A = g1b; a0.lda = Wq; a0.B = PB + L.b_d1w(j); a0.ldb = Wq; a0.C = g0b; a0.ldc = Wq;
    a0.b_tf32 = tc;
    a0.gateC = y0; a0.ldgc = Wq; a0.M = (int)R; a0.N = Wq; a0.K = Wq;
    IBC_TRY(pgemm(h, a0, false, true, s));
    IBC_TRY(dwg(h, x, Wv, AOP_RELU, g0b, Wq, R, Wv, Wq, grads + L.b_d0w(j), s));
    IBC_TRY(colsum_add(h, g0b, Wq, R, Wq, grads + L.b_d0b(j), s));
    GemmP ax;  // G' = G + (g0 . D0^T) (.) (x > 0)
    ax.A = g0b; ax.lda = Wq; ax.B = PB + L.b_d0w(j); ax.ldb = Wq; ax.C = Gn; ax.ldc = Wv;
    ax.b_tf32 = tc;
    ax.gateC = x; ax.ldgc = Wv; ax.res = G; ax.ldres = Wv; ax.M = (int)R; ax.N = Wv; ax.K = Wq;
    IBC_TRY(pgemm(h, ax, false, true, s));
    G = Gn;
  }
  // dense0 on x = [enc tiled, act]
  if (tc && L.A <= kMaxActRows && Wv % 128 == 0) {   // one pass over G for both reductions
    IBC_TRY(act_rows_bwd(h, G, act, B, n, L.A, Wv, gsum, grads + L.d0_w + (int64_t)kEnc * Wv, s));
  } else {
    IBC_TRY(dwg(h, act, L.A, AOP_ID, G, Wv, R, L.A, Wv, grads + L.d0_w + (int64_t)kEnc * Wv, s));
    IBC_TRY(groupsum(h, G, B, n, Wv, gsum, s));
  }
  IBC_TRY(dwg(h, enc, kEnc, AOP_ID, gsum, Wv, B, kEnc, Wv, grads + L.d0_w, s));
  IBC_TRY(colsum_add(h, gsum, Wv, B, Wv, grads + L.d0_b, s));
  if (dEnc) {  // dEnc = gsum . W0[:256]^T
    GemmP g;
    g.A = gsum; g.lda = Wv; g.B = P + L.d0_w; g.ldb = Wv; g.C = dEnc; g.ldc = kEnc;
    g.M = (int)B; g.N = kEnc; g.K = Wv;
    IBC_TRY(pgemm(h, g, false, true, s));
  }
  return IBC_OK;
}

// ---- value head: dE/dact and the gradient penalty ---------------------------------------------
// The ones-seeded reverse chain of the value head (mcmc.gradient_wrt_act on a late-fusion
// PixelEBM, mcmc.py:161-191): G_L = w1 in every row, then per block (last to first)
//   g1 = (G . D2^T) (.) (y1 > 0),  g0 = (g1 . D1^T) (.) (y0 > 0),  G' = G + (g0 . D0^T) (.) (x > 0)
// and dE/dact = G_0 . W0[256:]^T.  With `keep` every level is kept for the tangent pass of
// the gradient penalty; without it two levels ping-pong.
struct ValRev {
  float* G = nullptr;    // [levels, R, Wv]
  float* g0 = nullptr;   // [blocks kept, R, Wq]
  float* g1 = nullptr;
  int64_t R = 0;
  int Wv = 0, Wq = 0;
  bool keep = false;
  float* Gl(int j) const { return G + (int64_t)(keep ? j : (j & 1)) * R * Wv; }
  float* q0(int j) const { return g0 + (int64_t)(keep ? j : 0) * R * Wq; }
  float* q1(int j) const { return g1 + (int64_t)(keep ? j : 0) * R * Wq; }
};

size_t val_rev_bytes(const PixelLayout& L, int64_t R, bool keep) {
  const int lv = keep ? L.nblk + 1 : 2, lq = keep ? (L.nblk > 0 ? L.nblk : 1) : 1;
  return Workspace::rounded((size_t)lv * R * L.Wv * 4) +
         2 * Workspace::rounded((size_t)lq * R * L.Wq * 4);
}

void carve_val_rev(Workspace& w, const PixelLayout& L, int64_t R, bool keep, ValRev* r) {
  const int lv = keep ? L.nblk + 1 : 2, lq = keep ? (L.nblk > 0 ? L.nblk : 1) : 1;
  r->R = R; r->Wv = L.Wv; r->Wq = L.Wq; r->keep = keep;
  r->G = w.take<float>((size_t)lv * R * L.Wv);
  r->g0 = w.take<float>((size_t)lq * R * L.Wq);
  r->g1 = w.take<float>((size_t)lq * R * L.Wq);
}

// dEda [R, A] = +dE/dact for the rows value_fwd has just evaluated into `v`.
int value_dact(ibc_pixel_handle* ph, const PixelLayout& L, const ValSave& v, const ValRev& rv,
               float* dEda, cudaStream_t s) {
  ibc_handle* h = &ph->core;
  const float* P = ph->params;
  const int Wv = L.Wv, Wq = L.Wq;
  const int64_t R = v.R;
  const bool tc = h->precision == IBC_PRECISION_TF32;
  const float* PB = tc ? v.wr : P;   // rounded by value_fwd earlier in the same call
  // Every row of G_L is w1, so the last block's G . D2^T is the constant row ctop = D2 w1 of
  // value_prepare: no product and no materialised G_L (the residual reads w1 with a zero row
  // stride).  The tangent pass of the gradient penalty reads G_L, so `keep` fills it.
  const bool const_top = v.top && !rv.keep;
  if (!const_top) IBC_TRY(rowfill(h, P + L.d1_w, nullptr, R, Wv, rv.Gl(L.nblk), s));
  for (int j = L.nblk - 1; j >= 0; --j) {
    const bool top = const_top && j == L.nblk - 1;
    const float* G = top ? P + L.d1_w : rv.Gl(j + 1);
    const int64_t ldG = top ? 0 : Wv;
    if (top) {  // g1 = c (.) (y1 > 0)
      gate_const_rows_kernel<<<nblocks(R * (Wq / 4)), 256, 0, s>>>(
          v.ctop, v.Y1 + (int64_t)j * R * Wq, R, Wq, rv.q1(j));
      IBC_LAUNCH_CHECK(h);
    } else {
      GemmP a1;  // g1 = (G . D2^T) (.) (y1 > 0)
      a1.A = G; a1.lda = Wv; a1.B = PB + L.b_d2w(j); a1.ldb = Wv; a1.b_tf32 = tc;
      a1.C = rv.q1(j); a1.ldc = Wq; a1.gateC = v.Y1 + (int64_t)j * R * Wq; a1.ldgc = Wq;
      a1.M = (int)R; a1.N = Wq; a1.K = Wv;
      IBC_TRY(pgemm(h, a1, false, true, s));
    }
    GemmP a0;  // g0 = (g1 . D1^T) (.) (y0 > 0)
    a0.A = rv.q1(j); a0.lda = Wq; a0.B = PB + L.b_d1w(j); a0.ldb = Wq; a0.b_tf32 = tc;
    a0.C = rv.q0(j); a0.ldc = Wq; a0.gateC = v.Y0 + (int64_t)j * R * Wq; a0.ldgc = Wq;
    a0.M = (int)R; a0.N = Wq; a0.K = Wq;
    IBC_TRY(pgemm(h, a0, false, true, s));
    GemmP ax;  // G' = G + (g0 . D0^T) (.) (x > 0)
    ax.A = rv.q0(j); ax.lda = Wq; ax.B = PB + L.b_d0w(j); ax.ldb = Wq; ax.b_tf32 = tc;
    ax.C = rv.Gl(j); ax.ldc = Wv; ax.gateC = v.X + (int64_t)j * R * Wv; ax.ldgc = Wv;
    ax.res = G; ax.ldres = ldG; ax.M = (int)R; ax.N = Wv; ax.K = Wq;
    IBC_TRY(pgemm(h, ax, false, true, s));
  }
  if (L.A <= kMaxActRows) {   // dE/dact = G_0 . W0[256:]^T
    act_rows_dact_kernel<<<(unsigned)((R + 7) / 8), 256, 0, s>>>(
        rv.Gl(0), P + L.d0_w + (int64_t)kEnc * Wv, R, L.A, Wv, dEda);
    IBC_LAUNCH_CHECK(h);
    return IBC_OK;
  }
  GemmP g;
  g.A = rv.Gl(0); g.lda = Wv; g.B = P + L.d0_w + (int64_t)kEnc * Wv; g.ldb = Wv;
  g.C = dEda; g.ldc = L.A; g.M = (int)R; g.N = L.A; g.K = Wv;
  return gemm(h, g, false, true, s);
}

// gradient_loss.grad_penalty on the late-fusion PixelEBM (gradient_loss.py:35-72 through
// PixelEBM.call without a prior encoding, pixel_ebm.py:107-114): dE/dact at the combined
// actions, the penalty per observation (added to per_example) and its parameter gradients
// (accumulated into grads).  dE/dact depends on the parameters of the value head only --
// on the encoding, the biases and the encoder only through the ReLU masks -- so the
// gradient is a fixed-mask tangent pass T through the blocks (first to last):
//   T_0 = cot . W0[256:],                      dW0[256:] += cot^T . G_0
//   S0 = ((T (.) mx) . D0) (.) m0,             dD0 += (T (.) mx)^T . g0
//   S1 = (S0 . D1) (.) m1,                     dD1 += S0^T . g1
//   T' = T + S1 . D2,                          dD2 += S1^T . G_{j+1}
//   dw1 += sum_r T_L
int value_grad_penalty(ibc_pixel_handle* ph, const PixelLayout& L, int64_t B, int64_t n1,
                       const ValSave& v, const ibc_loss_cfg& c, float gscale, float* per_example,
                       float* grad_loss_out, float* grads, cudaStream_t s) {
  ibc_handle* h = &ph->core;
  const float* P = ph->params;
  const int Wv = L.Wv, Wq = L.Wq, A = L.A;
  const int64_t R = B * n1;
  const bool tc = h->precision == IBC_PRECISION_TF32;
  const float* PB = tc ? v.wr : P;
  ValRev rv;
  carve_val_rev(h->ws, L, R, true, &rv);
  float* dEda = h->ws.take<float>((size_t)R * A);
  float* cot = h->ws.take<float>((size_t)R * A);
  float* gl = h->ws.take<float>((size_t)B);
  float* Ta = h->ws.take<float>((size_t)R * Wv);
  float* Tb = h->ws.take<float>((size_t)R * Wv);
  float* S0 = h->ws.take<float>((size_t)R * Wq);
  float* S1 = h->ws.take<float>((size_t)R * Wq);
  IBC_TRY(value_dact(ph, L, v, rv, dEda, s));
  IBC_TRY(grad_penalty_from_dact(h, dEda, B, n1, A, c, gscale, gl, cot, s));
  add_rows_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(per_example, gl, B);
  IBC_LAUNCH_CHECK(h);
  if (grad_loss_out)
    IBC_CUDA(h, cudaMemcpyAsync(grad_loss_out, gl, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
  const float* Wa = P + L.d0_w + (int64_t)kEnc * Wv;
  {  // T_0 = cot . Wa
    GemmP g;
    g.A = cot; g.lda = A; g.B = Wa; g.ldb = Wv; g.C = Ta; g.ldc = Wv;
    g.M = (int)R; g.N = Wv; g.K = A;
    IBC_TRY(pgemm(h, g, false, false, s));
  }
  IBC_TRY(dwg(h, cot, A, AOP_ID, rv.Gl(0), Wv, R, A, Wv, grads + L.d0_w + (int64_t)kEnc * Wv, s));
  float* T = Ta;
  for (int j = 0; j < L.nblk; ++j) {
    float* Tn = (T == Ta) ? Tb : Ta;
    const float* x = v.X + (int64_t)j * R * Wv;
    const float* y0 = v.Y0 + (int64_t)j * R * Wq;
    const float* y1 = v.Y1 + (int64_t)j * R * Wq;
    GemmP s0;  // S0 = ((T (.) mx) . D0) (.) m0
    s0.A = T; s0.lda = Wv; s0.aop = AOP_GATE; s0.gateA = x; s0.B = PB + L.b_d0w(j); s0.ldb = Wq;
    s0.b_tf32 = tc; s0.C = S0; s0.ldc = Wq; s0.gateC = y0; s0.ldgc = Wq;
    s0.M = (int)R; s0.N = Wq; s0.K = Wv;
    IBC_TRY(pgemm(h, s0, false, false, s));
    {  // dD0 += (T (.) mx)^T . g0
      GemmP g;
      g.A = T; g.lda = Wv; g.aop = AOP_GATE; g.gateA = x; g.B = rv.q0(j); g.ldb = Wq;
      g.C = grads + L.b_d0w(j); g.ldc = Wq; g.M = Wv; g.N = Wq; g.K = (int)R;
      g.k_chunk = 1024; g.atomic = 1;
      IBC_TRY(pgemm(h, g, true, false, s));
    }
    GemmP s1;  // S1 = (S0 . D1) (.) m1
    s1.A = S0; s1.lda = Wq; s1.B = PB + L.b_d1w(j); s1.ldb = Wq; s1.b_tf32 = tc;
    s1.C = S1; s1.ldc = Wq; s1.gateC = y1; s1.ldgc = Wq; s1.M = (int)R; s1.N = Wq; s1.K = Wq;
    IBC_TRY(pgemm(h, s1, false, false, s));
    IBC_TRY(dwg(h, S0, Wq, AOP_ID, rv.q1(j), Wq, R, Wq, Wq, grads + L.b_d1w(j), s));
    IBC_TRY(dwg(h, S1, Wq, AOP_ID, rv.Gl(j + 1), Wv, R, Wq, Wv, grads + L.b_d2w(j), s));
    GemmP t;  // T' = T + S1 . D2
    t.A = S1; t.lda = Wq; t.B = PB + L.b_d2w(j); t.ldb = Wv; t.b_tf32 = tc;
    t.C = Tn; t.ldc = Wv; t.res = T; t.ldres = Wv; t.M = (int)R; t.N = Wv; t.K = Wq;
    IBC_TRY(pgemm(h, t, false, false, s));
    T = Tn;
  }
  return colsum_add(h, T, Wv, R, Wv, grads + L.d1_w, s);
}

size_t val_gp_bytes(const PixelLayout& L, int64_t B, int64_t R) {
  return val_rev_bytes(L, R, true) + 2 * Workspace::rounded((size_t)R * L.A * 4) +
         Workspace::rounded((size_t)B * 4) + 2 * Workspace::rounded((size_t)R * L.Wv * 4) +
         2 * Workspace::rounded((size_t)R * L.Wq * 4);
}

// An open training step (ibc_pixel_train_begin): everything carved from core.ws.
struct PendingTrain {
  int64_t B = 0, N = 0;
  int precision = 0;
  const float* params = nullptr;
  EncSave sv;
  ValSave v;
  float *dA = nullptr, *dB = nullptr, *enc = nullptr, *dEnc = nullptr, *E = nullptr, *dE = nullptr;
};

void drop_pending(ibc_pixel_handle* ph) {
  delete static_cast<PendingTrain*>(ph->pending);
  ph->pending = nullptr;
}

int check_arch(ibc_handle* h, const ibc_pixel_arch& a) {
  if (a.seq_len < 1 || a.height < 1 || a.width < 1 || a.channels < 1 || a.act_dim < 1 ||
      a.value_width < 4 || a.value_width % 4 || a.value_blocks < 0)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_create: bad architecture");
  if (a.target_height < 16 || a.target_width < 16)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_create: target size must be >= 16 (four 2x2 pools)");
  return IBC_OK;
}

}  // namespace
}  // namespace ibc

using namespace ibc;

#define PIX_ENTER(ph)                                                                  \
  if (!(ph)) { strncpy(ibc::g_create_error, "null handle", 511); return IBC_ERR_INVALID; } \
  ibc_handle* h = &(ph)->core;                                                         \
  int _prev = -1; cudaGetDevice(&_prev);                                               \
  if (_prev != h->device) cudaSetDevice(h->device);                                    \
  struct _Restore { int p, d; ~_Restore() { if (p >= 0 && p != d) cudaSetDevice(p); } } \
      _restore{_prev, h->device}

extern "C" {

int64_t ibc_pixel_param_count(const ibc_pixel_arch* arch) {
  if (!arch) return -1;
  return make_pixel_layout(*arch).count;
}

int ibc_pixel_create(const ibc_pixel_arch* arch, int device, ibc_pixel_handle** out) {
  ibc_handle* h = nullptr;
  if (!arch || !out) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_create: null argument");
  IBC_TRY(check_arch(h, *arch));
  int ndev = 0;
  IBC_CUDA(h, cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_create: device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  IBC_CUDA(h, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "ibc_b200 is built for sm_100a only; device %d is sm_%d%d",
             device, prop.major, prop.minor);
  ibc_pixel_handle* ph = new ibc_pixel_handle();
  ph->arch = *arch;
  ph->core.device = device;
  ph->core.sm_count = prop.multiProcessorCount;
  ph->core.precision = IBC_PRECISION_TF32;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  const cudaError_t me = cudaMalloc((void**)&ph->core.dev_status, sizeof(int));
  if (me == cudaSuccess) cudaMemset(ph->core.dev_status, 0, sizeof(int));
  if (prev >= 0) cudaSetDevice(prev);
  if (me != cudaSuccess) {
    delete ph;
    IBC_FAIL(h, IBC_ERR_CUDA, "ibc_pixel_create: cudaMalloc failed: %s", cudaGetErrorString(me));
  }
  *out = ph;
  return IBC_OK;
}

int ibc_pixel_destroy(ibc_pixel_handle* ph) {
  if (!ph) return IBC_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ph->core.device);
  cudaDeviceSynchronize();
  if (ph->core.ws.base) cudaFree(ph->core.ws.base);
  if (ph->core.ws_outer.base) cudaFree(ph->core.ws_outer.base);
  if (ph->core.dev_status) cudaFree(ph->core.dev_status);
  if (prev >= 0) cudaSetDevice(prev);
  drop_pending(ph);
  delete ph;
  return IBC_OK;
}

const char* ibc_pixel_last_error(const ibc_pixel_handle* ph) {
  return ph ? ph->core.err.c_str() : ibc::g_create_error;
}

int ibc_pixel_set_precision(ibc_pixel_handle* ph, int precision) {
  PIX_ENTER(ph);
  if (precision != IBC_PRECISION_FP32 && precision != IBC_PRECISION_TF32)
    IBC_FAIL(h, IBC_ERR_INVALID, "unknown precision %d", precision);
  drop_pending(ph);
  h->precision = precision;
  return IBC_OK;
}

int ibc_pixel_get_precision(const ibc_pixel_handle* ph) {
  return ph ? ph->core.precision : IBC_ERR_INVALID;
}

int ibc_pixel_kernel_status(ibc_pixel_handle* ph, int32_t* status_host) {
  PIX_ENTER(ph);
  if (!status_host) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_kernel_status: null argument");
  IBC_CUDA(h, cudaDeviceSynchronize());
  int st = 0;
  IBC_CUDA(h, cudaMemcpy(&st, h->dev_status, sizeof(int), cudaMemcpyDeviceToHost));
  IBC_CUDA(h, cudaMemset(h->dev_status, 0, sizeof(int)));
  *status_host = st;
  return IBC_OK;
}

int ibc_pixel_bind_params(ibc_pixel_handle* ph, const float* params) {
  PIX_ENTER(ph);
  if (!params) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_bind_params: null params");
  drop_pending(ph);
  ph->params = params;
  return IBC_OK;
}

int ibc_pixel_encode(ibc_pixel_handle* ph, const void* images, int32_t is_uint8, int64_t B,
                     float* enc, void* stream) {
  PIX_ENTER(ph);
  drop_pending(ph);   // core.ws is re-carved below
  if (!images || !enc || B < 1) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_encode: bad argument");
  if (!ph->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_pixel_bind_params)");
  const PixelLayout L = make_pixel_layout(ph->arch);
  int chunk = 1;
  const bool tc = h->precision == IBC_PRECISION_TF32;
  const size_t bytes = enc_bytes(ph->arch, B, &chunk, tc);
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  EncSave sv;
  carve_enc(h, ph->arch, B, chunk, tc, &sv);
  return encode_fwd(ph, L, images, is_uint8, B, sv, enc, (cudaStream_t)stream);
}

int ibc_pixel_energy(ibc_pixel_handle* ph, const float* enc, int64_t B, const float* act, int64_t n,
                     float* energy, void* stream) {
  PIX_ENTER(ph);
  if (!enc || !act || !energy || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_energy: bad argument");
  if (!ph->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_pixel_bind_params)");
  const PixelLayout L = make_pixel_layout(ph->arch);
  IBC_CUDA(h, h->ws_outer.reserve(val_bytes(L, B, B * n)));
  h->ws_outer.reset();
  ValSave v;
  carve_val(h->ws_outer, L, B, B * n, &v);
  v.top = top_shortcut_ok(L);
  return value_fwd(ph, L, enc, B, act, n, v, energy, (cudaStream_t)stream);
}

int ibc_pixel_dfo(ibc_pixel_handle* ph, const float* enc, int64_t B, float* actions, int64_t n,
                  const float* min_actions, const float* max_actions, float temperature,
                  int32_t num_iterations, float iteration_std, const double* uniforms,
                  const float* normals, uint64_t seed, float* probs, void* stream) {
  PIX_ENTER(ph);
  if (!enc || !actions || !min_actions || !max_actions || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_dfo: bad argument");
  if (!ph->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_pixel_bind_params)");
  const PixelLayout L = make_pixel_layout(ph->arch);
  cudaStream_t s = (cudaStream_t)stream;
  // dfo_generic keeps its own buffers at the start of core.ws_outer (reserved here together
  // with the value head's, so that its reserve is a no-op); the energy pass carves behind them
  const int64_t R = B * n;
  IBC_CUDA(h, h->ws_outer.reserve(2 * Workspace::rounded((size_t)R * 4) +
                                  Workspace::rounded((size_t)R * 8) +
                                  Workspace::rounded((size_t)R * ph->arch.act_dim * 4) +
                                  val_bytes(L, B, R)));
  auto energy_fn = [&](const float* acts, float* E) -> int {
    const size_t mark = h->ws_outer.used;
    ValSave v;
    carve_val(h->ws_outer, L, B, R, &v);
    v.top = top_shortcut_ok(L);
    const int rc = value_fwd(ph, L, enc, B, acts, n, v, E, s);
    h->ws_outer.used = mark;
    return rc;
  };
  return dfo_generic(h, ph->arch.act_dim, B, actions, n, min_actions, max_actions, temperature,
                     num_iterations, iteration_std, uniforms, normals, seed, probs, energy_fn, s);
}

int ibc_pixel_energy_dact(ibc_pixel_handle* ph, const float* enc, int64_t B, const float* act,
                          int64_t n, float* energy, float* dE_dact, void* stream) {
  PIX_ENTER(ph);
  if (!enc || !act || !energy || !dE_dact || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_energy_dact: bad argument");
  if (!ph->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_pixel_bind_params)");
  const PixelLayout L = make_pixel_layout(ph->arch);
  const int64_t R = B * n;
  IBC_CUDA(h, h->ws_outer.reserve(val_bytes(L, B, R) + val_rev_bytes(L, R, false)));
  h->ws_outer.reset();
  ValSave v;
  carve_val(h->ws_outer, L, B, R, &v);
  v.top = top_shortcut_ok(L);
  ValRev rv;
  carve_val_rev(h->ws_outer, L, R, false, &rv);
  IBC_TRY(value_fwd(ph, L, enc, B, act, n, v, energy, (cudaStream_t)stream));
  return value_dact(ph, L, v, rv, dE_dact, (cudaStream_t)stream);
}

int ibc_pixel_langevin(ibc_pixel_handle* ph, const float* enc, int64_t B, float* actions, int64_t n,
                       const float* min_actions, const float* max_actions,
                       const ibc_langevin_cfg* cfg, const float* normals, uint64_t seed,
                       float* chain_actions, float* chain_energies, float* chain_grad_norms,
                       void* stream) {
  PIX_ENTER(ph);
  if (!enc || !actions || !min_actions || !max_actions || !cfg || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_langevin: bad argument");
  if (!ph->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_pixel_bind_params)");
  if (cfg->num_iterations < 0) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_langevin: num_iterations < 0");
  if (cfg->grad_norm_type < IBC_NORM_NONE || cfg->grad_norm_type > IBC_NORM_INF)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_langevin: unknown grad_norm_type %d",
             cfg->grad_norm_type);
  if (cfg->apply_exp)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "ibc_pixel_langevin: apply_exp is not built (no gin sets it)");
  cudaStream_t s = (cudaStream_t)stream;
  const PixelLayout L = make_pixel_layout(ph->arch);
  const int A = L.A;
  const int64_t R = B * n;
  std::vector<float> steps;
  langevin_stepsizes(*cfg, &steps);
  IBC_CUDA(h, h->ws_outer.reserve(val_bytes(L, B, R) + val_rev_bytes(L, R, false) +
                                  2 * Workspace::rounded((size_t)R * 4) +
                                  Workspace::rounded((size_t)R * A * 4)));
  h->ws_outer.reset();
  ValSave v;
  carve_val(h->ws_outer, L, B, R, &v);
  ValRev rv;
  carve_val_rev(h->ws_outer, L, R, false, &rv);
  float* E = h->ws_outer.take<float>(R);
  float* nrm = h->ws_outer.take<float>(R);
  float* dEda = h->ws_outer.take<float>((size_t)R * A);
  // the encoding is fixed along the chain: its projection and the rounded weights once
  v.top = top_shortcut_ok(L);
  if (cfg->num_iterations > 0) IBC_TRY(value_prepare(ph, L, enc, B, v, s));
  for (int k = 0; k < cfg->num_iterations; ++k) {
    float* Ek = chain_energies ? chain_energies + (int64_t)k * R : E;
    float* Nk = chain_grad_norms ? chain_grad_norms + (int64_t)k * R : nrm;
    IBC_TRY(value_rows(ph, L, B, actions, n, v, Ek, s));
    IBC_TRY(value_dact(ph, L, v, rv, dEda, s));
    IBC_TRY(langevin_update(h, actions, dEda, R, A, normals ? normals + (int64_t)k * R * A : nullptr,
                            seed, (uint64_t)k + 1, cfg->noise_scale, cfg->grad_clip,
                            cfg->delta_action_clip, steps[k], min_actions, max_actions,
                            cfg->grad_norm_type, Nk, s));
    if (chain_actions)   // post-step actions (update_chain_data, mcmc.py:297-327)
      IBC_CUDA(h, cudaMemcpyAsync(chain_actions + (int64_t)k * R * A, actions, (size_t)R * A * 4,
                                  cudaMemcpyDeviceToDevice, s));
  }
  return IBC_OK;
}

int ibc_pixel_train_begin(ibc_pixel_handle* ph, const void* images, int32_t is_uint8, int64_t B,
                          int64_t N, float* enc_out, void* stream) {
  PIX_ENTER(ph);
  drop_pending(ph);
  if (!images || B < 1 || N < 0) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_train_begin: bad argument");
  if (!ph->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_pixel_bind_params)");
  cudaStream_t s = (cudaStream_t)stream;
  const PixelLayout L = make_pixel_layout(ph->arch);
  const EncDims d = enc_dims(ph->arch);
  const int64_t n1 = N + 1, R = B * n1;
  int chunk = 1;
  size_t dmax = 0;   // largest activation tensor (gradient ping-pong buffers)
  for (int l = 1; l <= 4; ++l)
    dmax = std::max(dmax, (size_t)B * d.H[l] * d.W[l] * d.C[l]);
  const bool tc = h->precision == IBC_PRECISION_TF32;
  size_t bytes = enc_bytes(ph->arch, B, &chunk, tc) + val_bytes(L, B, R) +
                 2 * Workspace::rounded(dmax * 4) + 2 * Workspace::rounded((size_t)R * L.Wv * 4) +
                 2 * Workspace::rounded((size_t)R * L.Wq * 4) +
                 Workspace::rounded((size_t)B * L.Wv * 4) + 2 * Workspace::rounded((size_t)B * kEnc * 4) +
                 2 * Workspace::rounded((size_t)R * 4) + val_gp_bytes(L, B, R) + (1 << 16);
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  PendingTrain* pt = new PendingTrain;
  pt->B = B; pt->N = N; pt->precision = h->precision; pt->params = ph->params;
  carve_enc(h, ph->arch, B, chunk, tc, &pt->sv);
  carve_val(h->ws, L, B, R, &pt->v);
  pt->dA = h->ws.take<float>(dmax);
  pt->dB = h->ws.take<float>(dmax);
  pt->enc = h->ws.take<float>((size_t)B * kEnc);
  pt->dEnc = h->ws.take<float>((size_t)B * kEnc);
  pt->E = h->ws.take<float>(R);
  pt->dE = h->ws.take<float>(R);
  const int rc = encode_fwd(ph, L, images, is_uint8, B, pt->sv, pt->enc, s);
  if (rc != IBC_OK) { delete pt; return rc; }
  if (enc_out)
    IBC_CUDA(h, cudaMemcpyAsync(enc_out, pt->enc, (size_t)B * kEnc * 4, cudaMemcpyDeviceToDevice, s));
  ph->pending = pt;
  return IBC_OK;
}

int ibc_pixel_train_finish(ibc_pixel_handle* ph, const float* actions, const ibc_loss_cfg* cfg,
                           int64_t global_batch, float* per_example_loss, float* grad_loss,
                           float* energies, float* grads, void* stream) {
  PIX_ENTER(ph);
  PendingTrain* pt = static_cast<PendingTrain*>(ph->pending);
  if (!pt)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_train_finish: no open ibc_pixel_train_begin (any "
                                 "encode / parameter / precision call on the handle closes it)");
  struct Close { ibc_pixel_handle* p; ~Close() { drop_pending(p); } } close{ph};
  if (!actions || !cfg || !per_example_loss || !grads || global_batch < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_train_finish: bad argument");
  if (cfg->grad_norm_type < IBC_NORM_NONE || cfg->grad_norm_type > IBC_NORM_INF)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_train_finish: unknown grad_norm_type %d",
             cfg->grad_norm_type);
  const bool gp = cfg->add_grad_penalty != 0 && cfg->grad_norm_type != IBC_NORM_NONE;
  cudaStream_t s = (cudaStream_t)stream;
  const PixelLayout L = make_pixel_layout(ph->arch);
  const int64_t B = pt->B, n1 = pt->N + 1, R = B * n1;
  IBC_CUDA(h, cudaMemsetAsync(grads, 0, (size_t)L.count * 4, s));
  IBC_TRY(value_fwd(ph, L, pt->enc, B, actions, n1, pt->v, pt->E, s));
  if (energies)
    IBC_CUDA(h, cudaMemcpyAsync(energies, pt->E, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
  IBC_TRY(ebm_loss(h, *cfg, pt->E, actions, B, n1, ph->arch.act_dim, 1.0f / (float)global_batch,
                   per_example_loss, pt->dE, s));
  IBC_TRY(value_bwd(ph, L, pt->enc, B, actions, n1, pt->v, pt->dE, grads, pt->dEnc, s));
  IBC_TRY(encode_bwd(ph, L, B, pt->sv, pt->dEnc, grads, pt->dA, pt->dB, s));
  if (gp) {
    IBC_TRY(value_grad_penalty(ph, L, B, n1, pt->v, *cfg, 1.0f / (float)global_batch,
                               per_example_loss, grad_loss, grads, s));
  } else if (grad_loss) {
    IBC_CUDA(h, cudaMemsetAsync(grad_loss, 0, (size_t)B * 4, s));
  }
  return IBC_OK;
}

int ibc_pixel_train_grads(ibc_pixel_handle* ph, const void* images, int32_t is_uint8, int64_t B,
                          const float* actions, int64_t N, const ibc_loss_cfg* cfg,
                          int64_t global_batch, float* per_example_loss, float* grad_loss,
                          float* energies, float* grads, void* stream) {
  if (ph && (!actions || !cfg || !per_example_loss || !grads || global_batch < 1)) {
    ibc_handle* h = &ph->core;
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_pixel_train_grads: bad argument");
  }
  const int rc = ibc_pixel_train_begin(ph, images, is_uint8, B, N, nullptr, stream);
  if (rc != IBC_OK) return rc;
  return ibc_pixel_train_finish(ph, actions, cfg, global_batch, per_example_loss, grad_loss,
                                energies, grads, stream);
}

int ibc_conv3x3_tf32(const float* X, int64_t nb, int32_t H, int32_t W, int32_t C, const float* K,
                     int32_t Co, const float* bias, int32_t relu, int32_t mode, const float* dY,
                     float* out, int32_t* status_host, void* stream) {
  ibc_handle* none = nullptr;
  if (!K || !out || nb < 1 || H < 1 || W < 1 || C < 1 || Co < 1 || mode < 0 || mode > 2 ||
      (mode != 2 && !X) || (mode != 0 && !dY) || nb * H * W > 0x7fffffffLL)
    IBC_FAIL(none, IBC_ERR_INVALID, "ibc_conv3x3_tf32: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, major = 0, sms = 0;
  IBC_CUDA(none, cudaGetDevice(&dev));
  IBC_CUDA(none, cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  IBC_CUDA(none, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10) IBC_FAIL(none, IBC_ERR_UNSUPPORTED, "ibc_b200 is built for sm_100a only");
  ibc_handle tmp;                       // error string, SM count and status word
  tmp.device = dev;
  tmp.sm_count = sms;
  float* flipped = nullptr;
  IBC_CUDA(none, cudaMalloc((void**)&tmp.dev_status, sizeof(int)));
  cudaMemsetAsync(tmp.dev_status, 0, sizeof(int), s);
  const int rows = (int)(nb * H * W);
  GemmP g;
  int rc = IBC_OK;
  if (mode == 0) {          // Y = f(im2col(X) . K + b)
    g.A = X; g.conv_H = H; g.conv_W = W; g.conv_C = C; g.B = K; g.ldb = Co; g.C = out; g.ldc = Co;
    g.bias = bias; g.M = rows; g.N = Co; g.K = 9 * C; g.relu_out = relu;
    rc = gemm_tc(&tmp, g, false, false, s);
  } else if (mode == 1) {   // dK += im2col(X)^T . dY
    g.A = X; g.conv_H = H; g.conv_W = W; g.conv_C = C; g.B = dY; g.ldb = Co; g.C = out; g.ldc = Co;
    g.M = 9 * C; g.N = Co; g.K = rows; g.atomic = 1;
    rc = gemm_tc(&tmp, g, true, false, s);
  } else {                  // dX = im2col(dY) . flip(K)
    if (cudaMalloc((void**)&flipped, (size_t)9 * C * Co * 4) != cudaSuccess) rc = IBC_ERR_CUDA;
    if (rc == IBC_OK) {
      flip_kernel3x3_kernel<<<nblocks(9 * C * Co), 256, 0, s>>>(K, C, Co, flipped);
      g.A = dY; g.conv_H = H; g.conv_W = W; g.conv_C = Co; g.B = flipped; g.ldb = C; g.C = out;
      g.ldc = C; g.M = rows; g.N = C; g.K = 9 * Co; g.b_tf32 = 1;
      rc = gemm_tc(&tmp, g, false, false, s);
    }
  }
  int st = 0;
  if (cudaStreamSynchronize(s) != cudaSuccess && rc == IBC_OK) rc = IBC_ERR_CUDA;
  if (rc == IBC_OK) cudaMemcpy(&st, tmp.dev_status, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(tmp.dev_status);
  if (flipped) cudaFree(flipped);
  if (rc != IBC_OK)
    strncpy(ibc::g_create_error, tmp.err.empty() ? "ibc_conv3x3_tf32 failed" : tmp.err.c_str(), 511);
  if (status_host) *status_host = st;
  return rc;
}

}  // extern "C"
