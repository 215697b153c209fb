// Per-layer dense-stack passes (GEMMs with fused prologue/epilogue): fp32 FMA kernels on
// IBC_PRECISION_FP32 handles; on IBC_PRECISION_TF32 handles the same products run on the
// general tcgen05 tf32 GEMM (gemm_tc.cu) -- this is the training path of the widths and
// options the layer-fused kernels do not cover (gradient penalty, W = 256 / 512 / 2048).
//
//   forward   networks/mlp_ebm.py:72-96, networks/layers/resnet.py:135-153
//   reverse   dE/dact of mcmc.gradient_wrt_act (ibc/agents/mcmc.py:161-191)
//
// The observation half of the first Dense layer is hoisted: concat([obs, act])
// . Wp == obs . Wp[:O] + act . Wp[O:], and obs is shared by the n rows of an
// observation (tile_batch), so obs . Wp[:O] + bp is computed once per
// observation and added as a per-group row bias.
#include "gemm_simt.cuh"
#include "mlp.cuh"

namespace ibc {

size_t fwd_save_bytes(const Layout& L, int64_t B, int64_t R) {
  return Workspace::rounded((size_t)B * L.W * 4) +
         Workspace::rounded((size_t)(L.nb + 1) * R * L.W * 4) +
         Workspace::rounded((size_t)(L.nb > 0 ? L.nb : 1) * R * L.W * 4);
}

void carve_fwd_save(ibc_handle* h, int64_t B, int64_t R, FwdSave* out) {
  const Layout& L = h->L;
  out->R = R;
  out->hp = h->ws.take<float>((size_t)B * L.W);
  out->H = h->ws.take<float>((size_t)(L.nb + 1) * R * L.W);
  out->V = h->ws.take<float>((size_t)(L.nb > 0 ? L.nb : 1) * R * L.W);
}

int mlp_forward_fp32(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                     float* E, FwdSave* save, cudaStream_t s) {
  const Layout& L = h->L;
  const float* P = h->params;
  const int64_t R = B * n;
  const int W = L.W;
  FwdSave local;
  float *hp, *Ha, *Hb, *Vt;
  if (save) {
    hp = save->hp;
  } else {
    // three rotating [R, W] buffers instead of the full per-block history
    hp = h->ws.take<float>((size_t)B * W);
    Ha = h->ws.take<float>((size_t)R * W);
    Hb = h->ws.take<float>((size_t)R * W);
    Vt = h->ws.take<float>((size_t)R * W);
  }
  {  // hp = obs . Wp[:O] + bp
    GemmP g;
    g.A = obs; g.lda = L.O; g.B = P + L.wp; g.ldb = W; g.C = hp; g.ldc = W;
    g.bias = P + L.bp; g.M = (int)B; g.N = W; g.K = L.O;
    IBC_TRY(gemm(h, g, false, false, s));   // always fp32 (once per observation)
  }
  float* Hcur = save ? save->Hj(0, W) : Ha;
  if (h->precision == IBC_PRECISION_TF32 && (W & 3) == 0) {
    // H0 = act . Wp[O:] + hp[row / n] as a bandwidth-bound rank-A pass (the FMA GEMM with
    // K = act_dim took 145 us for the 135 MB of C1's H0; the exact fp32 mode keeps the GEMM
    // and its summation order)
    IBC_TRY(act_rows(h, act, P + L.wp + (int64_t)L.O * W, hp, R, n, L.A, W, Hcur, s));
  } else {  // H0 = act . Wp[O:] + hp[row / n]
    GemmP g;
    g.A = act; g.lda = L.A; g.B = P + L.wp + (int64_t)L.O * W; g.ldb = W;
    g.C = Hcur; g.ldc = W; g.rowbias = hp; g.ldrb = W; g.rows_per_group = n;
    g.M = (int)R; g.N = W; g.K = L.A;
    g.tc = 1;
    IBC_TRY(gemm(h, g, false, false, s));
  }
  for (int j = 0; j < L.nb; ++j) {
    float* Vj = save ? save->Vj(j, W) : Vt;
    float* Hnext = save ? save->Hj(j + 1, W) : (Hcur == Ha ? Hb : Ha);
    GemmP g1;  // V = relu(H) . W1 + b1
    g1.A = Hcur; g1.lda = W; g1.aop = AOP_RELU; g1.B = P + L.w1(j); g1.ldb = W;
    g1.C = Vj; g1.ldc = W; g1.bias = P + L.b1(j); g1.M = (int)R; g1.N = W; g1.K = W;
    g1.tc = 1;
    IBC_TRY(gemm(h, g1, false, false, s));
    GemmP g2;  // H' = H + relu(V) . W2 + b2
    g2.A = Vj; g2.lda = W; g2.aop = AOP_RELU; g2.B = P + L.w2(j); g2.ldb = W;
    g2.C = Hnext; g2.ldc = W; g2.bias = P + L.b2(j); g2.res = Hcur; g2.ldres = W;
    g2.M = (int)R; g2.N = W; g2.K = W;
    g2.tc = 1;
    IBC_TRY(gemm(h, g2, false, false, s));
    Hcur = Hnext;
  }
  IBC_TRY(rowdot(h, Hcur, W, P + L.we, P + L.be, R, W, E, s));
  return IBC_OK;
}

int mlp_reverse_fp32(ibc_handle* h, const FwdSave& f, const float* row_scale, float* dEda,
                     RevSave* rev, cudaStream_t s) {
  const Layout& L = h->L;
  const float* P = h->params;
  const int64_t R = f.R;
  const int W = L.W;
  float *Ga, *Gb, *Ut;
  if (!rev) {
    Ga = h->ws.take<float>((size_t)R * W);
    Gb = h->ws.take<float>((size_t)R * W);
    Ut = h->ws.take<float>((size_t)R * W);
  }
  float* Gcur = rev ? rev->Gj(L.nb, W) : Ga;
  IBC_TRY(rowfill(h, P + L.we, row_scale, R, W, Gcur, s));
  for (int j = L.nb - 1; j >= 0; --j) {
    float* Uj = rev ? rev->Uj(j, W) : Ut;
    float* Gnext = rev ? rev->Gj(j, W) : (Gcur == Ga ? Gb : Ga);
    GemmP g1;  // U = (G . W2^T) (.) (V > 0)
    g1.A = Gcur; g1.lda = W; g1.B = P + L.w2(j); g1.ldb = W; g1.C = Uj; g1.ldc = W;
    g1.gateC = f.Vj(j, W); g1.ldgc = W; g1.M = (int)R; g1.N = W; g1.K = W;
    g1.tc = 1;
    IBC_TRY(gemm(h, g1, false, true, s));
    GemmP g2;  // G' = G + (U . W1^T) (.) (H > 0)
    g2.A = Uj; g2.lda = W; g2.B = P + L.w1(j); g2.ldb = W; g2.C = Gnext; g2.ldc = W;
    g2.gateC = f.Hj(j, W); g2.ldgc = W; g2.res = Gcur; g2.ldres = W;
    g2.M = (int)R; g2.N = W; g2.K = W;
    g2.tc = 1;
    IBC_TRY(gemm(h, g2, false, true, s));
    Gcur = Gnext;
  }
  if (dEda) {  // dE/dact = G0 . Wp[O:]^T
    GemmP g;
    g.A = Gcur; g.lda = W; g.B = P + L.wp + (int64_t)L.O * W; g.ldb = W;
    g.C = dEda; g.ldc = L.A; g.M = (int)R; g.N = L.A; g.K = W;
    g.tc = 1;
    IBC_TRY(gemm(h, g, false, true, s));
  }
  return IBC_OK;
}

__global__ void exp_inplace_kernel(float* e, int64_t R) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) e[i] = expf(e[i]);
}

int mlp_energy_dact_fp32(ibc_handle* h, const float* obs, int64_t B, const float* act,
                         int64_t n, float* E, float* dEda, bool apply_exp, cudaStream_t s) {
  const Layout& L = h->L;
  const int64_t R = B * n;
  const size_t bytes = fwd_save_bytes(L, B, R) + 3 * Workspace::rounded((size_t)R * L.W * 4);
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  FwdSave f;
  carve_fwd_save(h, B, R, &f);
  IBC_TRY(mlp_forward_fp32(h, obs, B, act, n, E, &f, s));
  if (apply_exp) {  // gradient of exp(E): cotangent exp(E) per row (mcmc.py:185-187)
    exp_inplace_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(E, R);
    IBC_LAUNCH_CHECK(h);
  }
  IBC_TRY(mlp_reverse_fp32(h, f, apply_exp ? E : nullptr, dEda, nullptr, s));
  return IBC_OK;
}

int mlp_energy(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
               float* E, cudaStream_t s) {
  if (!h->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_bind_params)");
  if (h->precision == IBC_PRECISION_TF32)
    return mlp_energy_umma(h, obs, B, act, n, E, nullptr, false, s);
  const Layout& L = h->L;
  const int64_t R = B * n;
  const size_t bytes = Workspace::rounded((size_t)B * L.W * 4) +
                       3 * Workspace::rounded((size_t)R * L.W * 4);
  IBC_CUDA(h, h->ws.reserve(bytes));
  h->ws.reset();
  return mlp_forward_fp32(h, obs, B, act, n, E, nullptr, s);
}

int mlp_energy_dact(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                    float* E, float* dEda, bool apply_exp, cudaStream_t s) {
  if (!h->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_bind_params)");
  if (h->precision == IBC_PRECISION_TF32)
    return mlp_energy_umma(h, obs, B, act, n, E, dEda, apply_exp, s);
  return mlp_energy_dact_fp32(h, obs, B, act, n, E, dEda, apply_exp, s);
}

}  // namespace ibc
