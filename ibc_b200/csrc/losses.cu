// InfoNCE, gradient penalty, the training backward pass and Adam.
//
//   ebm_loss.info_nce            ibc/losses/ebm_loss.py:21-48
//   gradient_loss.grad_penalty   ibc/losses/gradient_loss.py:23-72
//   ImplicitBCAgent._loss        ibc/agents/ibc_agent.py:132-300
//   tape.gradient + Adam         ibc/agents/base_agent.py:43-59,
//                                ibc/train/get_agent.py:64-68
#include <math.h>

#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "ops.cuh"

namespace ibc {
namespace {

constexpr float kKerasEps = 1e-7f;

__device__ __forceinline__ float block_reduce(float v, float* s_red, bool is_max) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = s_red[0];
  for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, s_red[i]) : r + s_red[i];
  return r;
}

// One block (128 threads) per example.  Keras KLDivergence(NONE) on
// softmax(pred / T) against one-hot labels at the LAST column, both clipped to
// [1e-7, 1].
__global__ void __launch_bounds__(128)
info_nce_kernel(const float* __restrict__ pred, int n1, float temperature, float grad_scale,
                float* __restrict__ loss, float* __restrict__ dpred,
                unsigned int* __restrict__ dpred_absmax) {
  __shared__ float s_red[4];
  const int tid = threadIdx.x;
  const float* row = pred + (int64_t)blockIdx.x * n1;
  float mx = -INFINITY;
  for (int j = tid; j < n1; j += 128) mx = fmaxf(mx, __fdiv_rn(row[j], temperature));
  mx = block_reduce(mx, s_red, true);
  float sum = 0.f;
  for (int j = tid; j < n1; j += 128) sum += expf(__fdiv_rn(row[j], temperature) - mx);
  sum = block_reduce(sum, s_red, false);
  float l = 0.f, qsum = 0.f;
  for (int j = tid; j < n1; j += 128) {
    const float p = __fdiv_rn(expf(__fdiv_rn(row[j], temperature) - mx), sum);
    const float yt = (j == n1 - 1) ? 1.0f : kKerasEps;
    // tf.clip_by_value propagates NaN (fminf / fmaxf would swallow it and hide a
    // diverged network from check_numerics, base_agent.py:46)
    const float yp = (p != p) ? p : fminf(fmaxf(p, kKerasEps), 1.0f);
    l += yt * logf(__fdiv_rn(yt, yp));
    // d loss / d p_j = -yt / p_j where the clip passes gradient
    const bool inside = (p >= kKerasEps) && (p <= 1.0f);
    if (inside) qsum += yt;
  }
  l = block_reduce(l, s_red, false);
  qsum = block_reduce(qsum, s_red, false);
  if (tid == 0) loss[blockIdx.x] = l;
  if (dpred) {
    float* drow = dpred + (int64_t)blockIdx.x * n1;
    float dmax = 0.f;
    for (int j = tid; j < n1; j += 128) {
      const float p = __fdiv_rn(expf(__fdiv_rn(row[j], temperature) - mx), sum);
      const float yt = (j == n1 - 1) ? 1.0f : kKerasEps;
      const bool inside = (p >= kKerasEps) && (p <= 1.0f);
      const float q = inside ? yt : 0.f;
      const float d = grad_scale * (p * qsum - q) / temperature;
      drow[j] = d;
      dmax = fmaxf(dmax, fabsf(d));
    }
    if (dpred_absmax) {     // max |dpred| over the batch (non-negative floats order like their bits)
      dmax = block_reduce(dmax, s_red, true);
      if (tid == 0) atomicMax(dpred_absmax, __float_as_uint(dmax));
    }
  }
}

// One block (128 threads) per example; threads stride over its n1 rows.
__global__ void __launch_bounds__(128)
grad_penalty_kernel(const float* __restrict__ dEda, int n1, int A, int norm_type, float margin,
                    int square, float weight, float grad_scale, float* __restrict__ grad_loss,
                    float* __restrict__ cot) {
  __shared__ float s_red[4];
  const int tid = threadIdx.x;
  float acc = 0.f;
  for (int i = tid; i < n1; i += 128) {
    const int64_t r = (int64_t)blockIdx.x * n1 + i;
    const float* d = dEda + r * A;
    float nrm = 0.f;
    for (int a = 0; a < A; ++a) {
      const float g = d[a];
      if (norm_type == IBC_NORM_INF) nrm = fmaxf(nrm, fabsf(g));
      else if (norm_type == IBC_NORM_L1) nrm += fabsf(g);
      else if (norm_type == IBC_NORM_L2) nrm = fmaf(g, g, nrm);
    }
    if (norm_type == IBC_NORM_L2) nrm = sqrtf(nrm);
    float x = nrm, dx = 1.f;
    if (margin >= 0.f) {
      x = nrm - margin;
      dx = (x >= 0.f && x <= 1e10f) ? 1.f : 0.f;
      x = (x != x) ? x : fminf(fmaxf(x, 0.f), 1e10f);
    }
    float pen = x;
    if (square) { pen = x * x; dx *= 2.f * x; }
    acc += pen;
    if (cot) {
      const float coef = grad_scale * weight * dx / (float)n1;
      int ties = 0;
      if (norm_type == IBC_NORM_INF)
        for (int a = 0; a < A; ++a) ties += (fabsf(d[a]) == nrm) ? 1 : 0;
      for (int a = 0; a < A; ++a) {
        const float g = d[a];
        const float sg = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);
        float dn = 0.f;
        if (norm_type == IBC_NORM_INF) dn = (fabsf(g) == nrm) ? sg / (float)ties : 0.f;
        else if (norm_type == IBC_NORM_L1) dn = sg;
        else if (norm_type == IBC_NORM_L2) dn = (nrm > 0.f) ? g / nrm : 0.f;
        cot[r * A + a] = coef * dn;
      }
    }
  }
  acc = block_reduce(acc, s_red, false);
  if (tid == 0) grad_loss[blockIdx.x] = acc / (float)n1 * weight;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, int64_t n, float lr_t,
                            float b1, float b2, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void add_vec_kernel(float* __restrict__ a, const float* __restrict__ b, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

// ebm_loss.cd / clipped_cd (ibc/losses/ebm_loss.py:51-66, 109-148): one block per example.
// err_j = |counter_j - true|^2 over the action dimensions (tf.norm(...)**2), threshold 0.2;
// d|x|/dx = sign(x) with 0 at 0 (tf.abs); reduce_mean over no counter examples is 0/0 = NaN
// as upstream.
__global__ void __launch_bounds__(128)
cd_loss_kernel(const float* __restrict__ pred, const float* __restrict__ actions, int n1, int A,
               int type, float grad_scale, float* __restrict__ loss, float* __restrict__ dpred,
               unsigned int* __restrict__ dpred_absmax) {
  __shared__ float s_red[4];
  const int tid = threadIdx.x;
  const int n = n1 - 1;
  const float* row = pred + (int64_t)blockIdx.x * n1;
  const float* arow = actions ? actions + (int64_t)blockIdx.x * n1 * A : nullptr;
  const float e_data = row[n];
  const float inv_n = 1.0f / (float)n;
  float acc = 0.f, dlast = 0.f, dmax = 0.f;
  float* drow = dpred ? dpred + (int64_t)blockIdx.x * n1 : nullptr;
  for (int j = tid; j < n; j += 128) {
    float term, dj, dn;
    if (type == IBC_LOSS_CD) {
      term = row[j];           // mean(samp) - data: the data term is subtracted once below
      dj = 1.f; dn = 0.f;
    } else {
      float err = 0.f;
      for (int a = 0; a < A; ++a) {
        const float d = arow[(int64_t)j * A + a] - arow[(int64_t)n * A + a];
        err = fmaf(d, d, err);
      }
      // tf.norm(...)**2: the square of the rounded square root
      const float nrm = sqrtf(err);
      err = nrm * nrm;
      const bool outside = err > 0.2f;
      term = outside ? row[j] - e_data : 0.f;
      dj = outside ? 1.f : 0.f;
      dn = outside ? -1.f : 0.f;
      if (type == IBC_LOSS_SOFT_CLIPPED_CD && !(err > 0.2f)) {
        const float res = (e_data - row[j]) - err;
        term += fabsf(res);
        const float sg = res > 0.f ? 1.f : (res < 0.f ? -1.f : 0.f);
        dj -= sg;
        dn += sg;
      }
    }
    acc += term;
    dlast += dn;
    if (drow) {
      const float d = grad_scale * dj * inv_n;
      drow[j] = d;
      dmax = fmaxf(dmax, fabsf(d));
    }
  }
  acc = block_reduce(acc, s_red, false);
  dlast = block_reduce(dlast, s_red, false);
  if (tid == 0) {
    float l = acc * inv_n, dl = grad_scale * dlast * inv_n;
    if (type == IBC_LOSS_CD) { l -= e_data; dl = -grad_scale; }
    loss[blockIdx.x] = l;
    if (drow) {
      drow[n] = dl;
      dmax = fmaxf(dmax, fabsf(dl));
    }
  }
  if (drow && dpred_absmax) {
    dmax = block_reduce(dmax, s_red, true);
    if (tid == 0) atomicMax(dpred_absmax, __float_as_uint(dmax));
  }
}


}  // namespace

int info_nce(ibc_handle* h, const float* pred, int64_t B, int64_t n1, float temperature,
             float grad_scale, float* loss, float* dpred, cudaStream_t s,
             unsigned int* dpred_absmax) {
  if (B <= 0 || n1 <= 0) IBC_FAIL(h, IBC_ERR_INVALID, "info_nce: bad shape");
  if (temperature <= 0.f) IBC_FAIL(h, IBC_ERR_INVALID, "info_nce: temperature must be > 0");
  // (one warp per example with the row in registers was tried: 10 us against this kernel's 7 for
  // 512 rows of 257 -- nine serial exp / IEEE divisions / log per lane cost more than the barriers)
  info_nce_kernel<<<(unsigned)B, 128, 0, s>>>(pred, (int)n1, temperature, grad_scale, loss, dpred,
                                              dpred_absmax);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int cd_loss(ibc_handle* h, const float* pred, const float* actions, int64_t B, int64_t n1, int A,
            int type, float grad_scale, float* loss, float* dpred, cudaStream_t s,
            unsigned int* dpred_absmax) {
  if (B <= 0 || n1 <= 0) IBC_FAIL(h, IBC_ERR_INVALID, "cd_loss: bad shape");
  if (type != IBC_LOSS_CD && type != IBC_LOSS_CLIPPED_CD && type != IBC_LOSS_SOFT_CLIPPED_CD)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "cd_loss: ebm_loss_type %d is not built (cd_kl needs a second network)", type);
  if (type != IBC_LOSS_CD && (!actions || A < 1))
    IBC_FAIL(h, IBC_ERR_INVALID, "cd_loss: the clipped forms read the combined actions");
  cd_loss_kernel<<<(unsigned)B, 128, 0, s>>>(pred, actions, (int)n1, A, type, grad_scale, loss, dpred,
                                             dpred_absmax);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// ImplicitBCAgent._compute_ebm_loss (ibc_agent.py:303-335) on the step's energies
int ebm_loss(ibc_handle* h, const ibc_loss_cfg& c, const float* pred, const float* actions, int64_t B,
             int64_t n1, int A, float grad_scale, float* loss, float* dpred, cudaStream_t s,
             unsigned int* dpred_absmax) {
  if (c.ebm_loss_type == IBC_LOSS_INFO_NCE)
    return info_nce(h, pred, B, n1, c.softmax_temperature, grad_scale, loss, dpred, s, dpred_absmax);
  return cd_loss(h, pred, actions, B, n1, A, c.ebm_loss_type, grad_scale, loss, dpred, s, dpred_absmax);
}

int grad_penalty_from_dact(ibc_handle* h, const float* dEda, int64_t B, int64_t n1, int A,
                           const ibc_loss_cfg& c, float grad_scale, float* grad_loss, float* cot,
                           cudaStream_t s) {
  if (B <= 0 || n1 <= 0 || A <= 0) IBC_FAIL(h, IBC_ERR_INVALID, "grad_penalty: bad shape");
  if (c.grad_norm_type < IBC_NORM_NONE || c.grad_norm_type > IBC_NORM_INF)
    IBC_FAIL(h, IBC_ERR_INVALID, "grad_penalty: unknown grad_norm_type %d", c.grad_norm_type);
  grad_penalty_kernel<<<(unsigned)B, 128, 0, s>>>(dEda, (int)n1, A, c.grad_norm_type,
                                                  c.grad_margin, c.square_grad_penalty,
                                                  c.grad_loss_weight, grad_scale, grad_loss, cot);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int adam_step(ibc_handle* h, float* p, const float* g, float* m, float* v, int64_t n, float lr_t,
              float b1, float b2, float eps, cudaStream_t s) {
  if (n <= 0) return IBC_OK;
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, g, m, v, n, lr_t, b1, b2, eps);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// dW (+)= f(X)^T . Y with split-K over the R rows.
static int dw_gemm(ibc_handle* h, const float* X, int64_t ldx, int aop, const float* gate,
                   const float* Y, int64_t ldy, int64_t R, int Kin, int Nout, float* dW,
                   cudaStream_t s, bool tc = true) {
  GemmP g;
  g.A = X; g.lda = ldx; g.aop = aop; g.gateA = gate; g.B = Y; g.ldb = ldy;
  g.C = dW; g.ldc = Nout; g.M = Kin; g.N = Nout; g.K = (int)R;
  g.k_chunk = R <= 1024 ? 64 : 1024;   // few rows (per-observation sums): still 8+ K slices
  g.atomic = 1; g.tc = tc ? 1 : 0;
  return gemm(h, g, true, false, s);
}

// ImplicitBCAgent._loss + tape.gradient on the fp32 path.
int train_grads_fp32(ibc_handle* h, const float* obs, int64_t B, const float* actions, int64_t N,
                     const ibc_loss_cfg& c, int64_t global_batch, float* per_example,
                     float* grad_loss_out, float* energies_out, float* grads, cudaStream_t s) {
  const Layout& L = h->L;
  const float* P = h->params;
  const int W = L.W, A = L.A, nb = L.nb;
  const int64_t n1 = N + 1, R = B * n1;
  const float gscale = 1.0f / (float)global_batch;
  const bool gp = c.add_grad_penalty != 0 && c.grad_norm_type != IBC_NORM_NONE;

  const size_t RW = Workspace::rounded((size_t)R * W * 4);
  size_t bytes = fwd_save_bytes(L, B, R) + 3 * RW + Workspace::rounded((size_t)B * W * 4) +
                 4 * Workspace::rounded((size_t)R * 4) + 2 * Workspace::rounded((size_t)R * A * 4) +
                 2 * Workspace::rounded((size_t)B * 4);
  if (gp) bytes += (size_t)(2 * nb + 2) * RW + 3 * RW;
  bytes += RW;  // slack for per-take rounding
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  FwdSave f;
  carve_fwd_save(h, B, R, &f);
  float* E = h->ws.take<float>(R);
  float* dE = h->ws.take<float>(R);
  float* lnce = h->ws.take<float>(B);
  float* gl = h->ws.take<float>(B);
  float* gsum = h->ws.take<float>((size_t)B * W);

  IBC_CUDA(h, cudaMemsetAsync(grads, 0, (size_t)L.count * 4, s));
  IBC_TRY(mlp_forward_fp32(h, obs, B, actions, n1, E, &f, s));
  if (energies_out)
    IBC_CUDA(h, cudaMemcpyAsync(energies_out, E, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
  IBC_TRY(ebm_loss(h, c, E, actions, B, n1, L.A, gscale, lnce, dE, s, nullptr));
  IBC_CUDA(h, cudaMemcpyAsync(per_example, lnce, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));

  if (gp) {
    // ---- gradient penalty: reverse chain, penalty, fixed-mask tangent pass
    // (hard part 2) -------------------------------------
    RevSave rv;
    rv.R = R;
    rv.G = h->ws.take<float>((size_t)(nb + 1) * R * W);
    rv.U = h->ws.take<float>((size_t)(nb > 0 ? nb : 1) * R * W);
    float* dEda = h->ws.take<float>((size_t)R * A);
    float* cot = h->ws.take<float>((size_t)R * A);
    float* Ta = h->ws.take<float>((size_t)R * W);
    float* Tb = h->ws.take<float>((size_t)R * W);
    float* Sj = h->ws.take<float>((size_t)R * W);
    IBC_TRY(mlp_reverse_fp32(h, f, nullptr, dEda, &rv, s));
    IBC_TRY(grad_penalty_from_dact(h, dEda, B, n1, A, c, gscale, gl, cot, s));
    add_vec_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(per_example, gl, B);
    IBC_LAUNCH_CHECK(h);
    const float* WpA = P + L.wp + (int64_t)L.O * W;
    {  // T0 = cot . Wp[O:]
      GemmP g;
      g.A = cot; g.lda = A; g.B = WpA; g.ldb = W; g.C = Ta; g.ldc = W;
      g.M = (int)R; g.N = W; g.K = A;
      g.tc = 1;
      IBC_TRY(gemm(h, g, false, false, s));
    }
    // dWp[O:] += cot^T . G0
    IBC_TRY(dw_gemm(h, cot, A, AOP_ID, nullptr, rv.Gj(0, W), W, R, A, W,
                    grads + L.wp + (int64_t)L.O * W, s));
    float* Tcur = Ta;
    for (int j = 0; j < nb; ++j) {
      float* Tnext = (Tcur == Ta) ? Tb : Ta;
      GemmP g1;  // S = ((T (.) mh) . W1) (.) mv
      g1.A = Tcur; g1.lda = W; g1.aop = AOP_GATE; g1.gateA = f.Hj(j, W);
      g1.B = P + L.w1(j); g1.ldb = W; g1.C = Sj; g1.ldc = W;
      g1.gateC = f.Vj(j, W); g1.ldgc = W; g1.M = (int)R; g1.N = W; g1.K = W;
      g1.tc = 1;
      IBC_TRY(gemm(h, g1, false, false, s));
      // dW1 += (T (.) mh)^T . U_j ;  dW2 += S^T . G_{j+1}
      IBC_TRY(dw_gemm(h, Tcur, W, AOP_GATE, f.Hj(j, W), rv.Uj(j, W), W, R, W, W,
                      grads + L.w1(j), s));
      IBC_TRY(dw_gemm(h, Sj, W, AOP_ID, nullptr, rv.Gj(j + 1, W), W, R, W, W, grads + L.w2(j), s));
      GemmP g2;  // T' = T + S . W2
      g2.A = Sj; g2.lda = W; g2.B = P + L.w2(j); g2.ldb = W; g2.C = Tnext; g2.ldc = W;
      g2.res = Tcur; g2.ldres = W; g2.M = (int)R; g2.N = W; g2.K = W;
      g2.tc = 1;
      IBC_TRY(gemm(h, g2, false, false, s));
      Tcur = Tnext;
    }
    IBC_TRY(colsum_add(h, Tcur, W, R, W, grads + L.we, s));   // dwe += sum_r T_nb
  } else {
    IBC_CUDA(h, cudaMemsetAsync(gl, 0, (size_t)B * 4, s));
  }
  if (grad_loss_out)
    IBC_CUDA(h, cudaMemcpyAsync(grad_loss_out, gl, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));

  // ---- InfoNCE backward ------------------------------------------------------
  float* Ga = h->ws.take<float>((size_t)R * W);
  float* Gb = h->ws.take<float>((size_t)R * W);
  float* GU = h->ws.take<float>((size_t)R * W);
  IBC_TRY(wcolsum_add(h, f.Hj(nb, W), W, dE, R, W, grads + L.we, s));   // dwe += H_nb^T dE
  IBC_TRY(sum_add(h, dE, R, grads + L.be, s));                          // dbe
  IBC_TRY(rowfill(h, P + L.we, dE, R, W, Ga, s));                       // g = dE (x) we
  float* Gcur = Ga;
  for (int j = nb - 1; j >= 0; --j) {
    float* Gnext = (Gcur == Ga) ? Gb : Ga;
    IBC_TRY(dw_gemm(h, f.Vj(j, W), W, AOP_RELU, nullptr, Gcur, W, R, W, W, grads + L.w2(j), s));
    IBC_TRY(colsum_add(h, Gcur, W, R, W, grads + L.b2(j), s));
    GemmP g1;  // gu = (g . W2^T) (.) (V > 0)
    g1.A = Gcur; g1.lda = W; g1.B = P + L.w2(j); g1.ldb = W; g1.C = GU; g1.ldc = W;
    g1.gateC = f.Vj(j, W); g1.ldgc = W; g1.M = (int)R; g1.N = W; g1.K = W;
    g1.tc = 1;
    IBC_TRY(gemm(h, g1, false, true, s));
    IBC_TRY(dw_gemm(h, f.Hj(j, W), W, AOP_RELU, nullptr, GU, W, R, W, W, grads + L.w1(j), s));
    IBC_TRY(colsum_add(h, GU, W, R, W, grads + L.b1(j), s));
    GemmP g2;  // g' = g + (gu . W1^T) (.) (H > 0)
    g2.A = GU; g2.lda = W; g2.B = P + L.w1(j); g2.ldb = W; g2.C = Gnext; g2.ldc = W;
    g2.gateC = f.Hj(j, W); g2.ldgc = W; g2.res = Gcur; g2.ldres = W;
    g2.M = (int)R; g2.N = W; g2.K = W;
    g2.tc = 1;
    IBC_TRY(gemm(h, g2, false, true, s));
    Gcur = Gnext;
  }
  // first layer: x = [obs tiled, act]
  if (h->precision == IBC_PRECISION_TF32 && A <= kMaxActRows && W % 128 == 0) {
    // one pass over G for both reductions (the K = R FMA product for two rows of dW took
    // 187 us on C1's 135 MB of G, the group sum another 28)
    IBC_TRY(act_rows_bwd(h, Gcur, actions, B, n1, A, W, gsum, grads + L.wp + (int64_t)L.O * W, s));
  } else {
    IBC_TRY(dw_gemm(h, actions, A, AOP_ID, nullptr, Gcur, W, R, A, W,
                    grads + L.wp + (int64_t)L.O * W, s));
    IBC_TRY(groupsum(h, Gcur, B, n1, W, gsum, s));
  }
  IBC_TRY(dw_gemm(h, obs, L.O, AOP_ID, nullptr, gsum, W, B, L.O, W, grads + L.wp, s, false));
  IBC_TRY(colsum_add(h, gsum, W, B, W, grads + L.bp, s));
  return IBC_OK;
}

}  // namespace ibc
