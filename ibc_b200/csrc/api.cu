// extern "C" surface of libibc_b200.so (include/ibc_b200.h).
#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "ops.cuh"

namespace ibc {
std::atomic<long long> g_launch_count{0};
thread_local char g_create_error[512] = {0};
}  // namespace ibc

using namespace ibc;

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define IBC_ENTER(h)                                                           \
  if (!(h)) { strncpy(ibc::g_create_error, "null handle", 511); return IBC_ERR_INVALID; } \
  DeviceGuard _guard((h)->device);                                             \
  if (!_guard.ok) IBC_FAIL(h, IBC_ERR_CUDA, "cannot select device %d", (h)->device)

}  // namespace

extern "C" {

const char* ibc_version(void) { return "ibc_b200 0.1 (sm_100a, tcgen05/tf32 + fp32)"; }

int64_t ibc_param_count(const ibc_mlp_arch* arch) {
  if (!arch) return -1;
  return make_layout(*arch).count;
}

int ibc_create(const ibc_mlp_arch* arch, int device, ibc_handle** out) {
  ibc_handle* h = nullptr;
  if (!arch || !out) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_create: null argument");
  if (arch->obs_dim < 1 || arch->act_dim < 1 || arch->width < 1 || arch->depth < 0)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_create: bad arch (O=%d A=%d W=%d D=%d)", arch->obs_dim,
             arch->act_dim, arch->width, arch->depth);
  if (arch->depth % 2 != 0)   // resnet.py:42
    IBC_FAIL(h, IBC_ERR_INVALID, "ResNet wants layers to be even numbers (depth=%d)", arch->depth);
  int ndev = 0;
  IBC_CUDA(h, cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_create: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) IBC_FAIL(h, IBC_ERR_CUDA, "ibc_create: cannot select device %d", device);
  cudaDeviceProp prop;
  IBC_CUDA(h, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED,
             "ibc_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major,
             prop.minor, prop.name);
  ibc_handle* nh = new ibc_handle();
  nh->arch = *arch;
  nh->L = make_layout(*arch);
  nh->device = device;
  nh->sm_count = prop.multiProcessorCount;
  if (cudaMalloc((void**)&nh->dev_status, sizeof(int)) != cudaSuccess) {
    delete nh;
    IBC_FAIL(h, IBC_ERR_CUDA, "ibc_create: cudaMalloc failed");
  }
  cudaMemset(nh->dev_status, 0, sizeof(int));
  nh->precision = (umma_supported(nh->L) || wide_supported(nh->L)) ? IBC_PRECISION_TF32
                                                                    : IBC_PRECISION_FP32;
  *out = nh;
  return IBC_OK;
}

int ibc_destroy(ibc_handle* h) {
  if (!h) return IBC_OK;
  {
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->packed) cudaFree(h->packed);
    if (h->packed16) cudaFree(h->packed16);
    if (h->loss_event) cudaEventDestroy(h->loss_event);
    if (h->dev_status) cudaFree(h->dev_status);
    for (int i = 0; i < 6; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (auto& g : h->langevin_graphs) g.release();
    if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    if (h->ws.base) cudaFree(h->ws.base);
    if (h->ws_outer.base) cudaFree(h->ws_outer.base);
  }
  delete h;
  return IBC_OK;
}

const char* ibc_last_error(const ibc_handle* h) {
  return h ? h->err.c_str() : ibc::g_create_error;
}

int ibc_set_precision(ibc_handle* h, int precision) {
  IBC_ENTER(h);
  if (precision != IBC_PRECISION_FP32 && precision != IBC_PRECISION_TF32)
    IBC_FAIL(h, IBC_ERR_INVALID, "unknown precision %d", precision);
  if (precision == IBC_PRECISION_TF32 && !umma_supported(h->L) && !wide_supported(h->L))
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED,
             "tcgen05 kernels cover width 128/256 (act_dim <= 32) and multiples of 128 above, "
             "depth >= 2 (got W=%d D=%d A=%d)", h->L.W, h->L.D, h->L.A);
  h->precision = precision;
  return IBC_OK;
}

int ibc_get_precision(const ibc_handle* h) { return h ? h->precision : IBC_ERR_INVALID; }

int ibc_bind_params(ibc_handle* h, const float* params, void* stream) {
  IBC_ENTER(h);
  (void)stream;
  if (!params) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_bind_params: null params");
  h->params = params;
  h->packed_stale = true;
  h->packed_rest_stale = true;
  return IBC_OK;
}

int ibc_params_changed(ibc_handle* h) {
  IBC_ENTER(h);
  h->packed_stale = true;
  h->packed_rest_stale = true;
  return IBC_OK;
}

int ibc_mlp_energy(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                   float* energy, void* stream) {
  IBC_ENTER(h);
  if (!obs || !act || !energy || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_mlp_energy: bad argument (B=%lld n=%lld)", (long long)B,
             (long long)n);
  return mlp_energy(h, obs, B, act, n, energy, (cudaStream_t)stream);
}

int ibc_mlp_energy_dact(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                        float* energy, float* dE_dact, void* stream) {
  IBC_ENTER(h);
  if (!obs || !act || !energy || !dE_dact || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_mlp_energy_dact: bad argument");
  return mlp_energy_dact(h, obs, B, act, n, energy, dE_dact, false, (cudaStream_t)stream);
}

int ibc_resample(const float* logits, const double* uniforms, int64_t B, int64_t n, int64_t count,
                 int32_t* draws, int32_t* counts, int64_t* rows, void* stream) {
  ibc_handle* h = nullptr;
  if (!logits || !uniforms) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_resample: null input");
  return resample(h, logits, uniforms, B, n, count, 0, 0, draws, counts, rows,
                  (cudaStream_t)stream);
}

int ibc_dfo(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
            const float* min_actions, const float* max_actions, float temperature,
            int32_t num_iterations, float iteration_std, const double* uniforms,
            const float* normals, uint64_t seed, float* probs, void* stream) {
  IBC_ENTER(h);
  if (!obs || !actions || !min_actions || !max_actions || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_dfo: bad argument");
  if (!h->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_bind_params)");
  return dfo(h, obs, B, actions, n, min_actions, max_actions, temperature, num_iterations,
             iteration_std, uniforms, normals, seed, probs, (cudaStream_t)stream);
}

int ibc_langevin(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
                 const float* min_actions, const float* max_actions, const ibc_langevin_cfg* cfg,
                 const float* normals, uint64_t seed, float* chain_actions, float* chain_energies,
                 float* chain_grad_norms, void* stream) {
  IBC_ENTER(h);
  if (!obs || !actions || !min_actions || !max_actions || !cfg || B < 1 || n < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_langevin: bad argument");
  if (!h->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_bind_params)");
  return langevin(h, obs, B, actions, n, min_actions, max_actions, *cfg, normals, seed,
                  chain_actions, chain_energies, chain_grad_norms, (cudaStream_t)stream);
}

int ibc_dfo_gather_perturb(const float* src, const int64_t* rows, int64_t R, int32_t A,
                           const float* normals, uint64_t seed, float std,
                           const float* min_actions, const float* max_actions, float* out,
                           void* stream) {
  ibc_handle* h = nullptr;
  if (!src || !rows || !out || (std != 0.f && (!min_actions || !max_actions)))
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_dfo_gather_perturb: null argument");
  return gather_perturb(h, src, rows, R, A, normals, seed, 2, std, min_actions, max_actions, out,
                        (cudaStream_t)stream);
}

int ibc_langevin_update(float* actions, const float* dE_dact, int64_t R, int32_t A,
                        const float* normals, uint64_t seed, float noise_scale, float grad_clip,
                        float delta_action_clip, float stepsize, const float* min_actions,
                        const float* max_actions, int32_t grad_norm_type, float* grad_norms,
                        void* stream) {
  ibc_handle* h = nullptr;
  if (!actions || !dE_dact || !min_actions || !max_actions)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_langevin_update: null argument");
  if (grad_norm_type < IBC_NORM_NONE || grad_norm_type > IBC_NORM_INF)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_langevin_update: unknown grad_norm_type %d", grad_norm_type);
  return langevin_update(h, actions, dE_dact, R, A, normals, seed, 1, noise_scale, grad_clip,
                         delta_action_clip, stepsize, min_actions, max_actions, grad_norm_type,
                         grad_norms, (cudaStream_t)stream);
}

int ibc_langevin_stepsizes(const ibc_langevin_cfg* cfg, float* out_host) {
  ibc_handle* h = nullptr;
  if (!cfg || !out_host) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_langevin_stepsizes: null argument");
  std::vector<float> v;
  langevin_stepsizes(*cfg, &v);
  for (size_t i = 0; i < v.size(); ++i) out_host[i] = v[i];
  return IBC_OK;
}

int ibc_softmax_rows(const float* logits, int64_t B, int64_t n, float temperature, float* probs,
                     void* stream) {
  ibc_handle* h = nullptr;
  if (!logits || !probs) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_softmax_rows: null argument");
  if (temperature <= 0.f) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_softmax_rows: temperature must be > 0");
  return softmax_rows(h, logits, B, n, temperature, probs, (cudaStream_t)stream);
}

int ibc_uniform_actions(const float* u, int64_t rows, int32_t A, const float* min_actions,
                        const float* max_actions, uint64_t seed, float* out, void* stream) {
  ibc_handle* h = nullptr;
  if (!min_actions || !max_actions || !out || rows < 0 || A < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_uniform_actions: bad argument");
  return uniform_actions(h, u, rows, A, min_actions, max_actions, seed, out, (cudaStream_t)stream);
}

int ibc_peer_alloc(int64_t bytes, void** ptr, void* handle64) {
  ibc_handle* h = nullptr;
  return peer_alloc(h, bytes, ptr, handle64);
}

int ibc_peer_open(const void* handle64, void** ptr) {
  ibc_handle* h = nullptr;
  return peer_open(h, handle64, ptr);
}

int ibc_peer_close(void* ptr, int32_t owned) {
  ibc_handle* h = nullptr;
  return peer_close(h, ptr, owned);
}

int ibc_peer_allreduce_adam(const float* const* peer_grads, int32_t* const* peer_flags, int32_t rank,
                            int32_t world, int32_t epoch, int64_t count, float* params, float* m,
                            float* v, float lr_t, float beta1, float beta2, float epsilon,
                            float* tail_out, int32_t* status, void* stream) {
  ibc_handle* h = nullptr;
  if (!peer_grads || !peer_flags || !params || !m || !v)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_peer_allreduce_adam: null argument");
  return peer_allreduce_adam(h, peer_grads, reinterpret_cast<int* const*>(peer_flags), rank, world, epoch,
                             count, params, m, v, lr_t, beta1, beta2, epsilon, tail_out,
                             reinterpret_cast<int*>(status), (cudaStream_t)stream);
}

int ibc_greedy_actions(const float* probs, const float* values, int64_t B, int64_t n, int32_t A,
                       const float* min_actions, const float* max_actions, float* out, void* stream) {
  ibc_handle* h = nullptr;
  if (!probs || !values || !out || (min_actions == nullptr) != (max_actions == nullptr))
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_greedy_actions: bad argument");
  return greedy_actions(h, probs, values, B, n, A, min_actions, max_actions, out, (cudaStream_t)stream);
}

int ibc_counter_examples_uniform(const float* u, int64_t B, int64_t N, int32_t A,
                                 const float* min_actions, const float* max_actions, uint64_t seed,
                                 const float* actions, float* combined, void* stream) {
  ibc_handle* h = nullptr;
  if (!min_actions || !max_actions || !actions || !combined || B < 0 || N < 0 || A < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_counter_examples_uniform: bad argument");
  return uniform_combined(h, u, B, N, A, min_actions, max_actions, seed, actions, combined,
                          (cudaStream_t)stream);
}

int ibc_info_nce(const float* predictions, int64_t B, int64_t n1, float temperature,
                 float grad_scale, float* loss, float* dpred, void* stream) {
  ibc_handle* h = nullptr;
  if (!predictions || !loss) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_info_nce: null argument");
  return info_nce(h, predictions, B, n1, temperature, grad_scale, loss, dpred,
                  (cudaStream_t)stream);
}

int ibc_cd_loss(const float* predictions, const float* actions, int64_t B, int64_t n1, int32_t A,
                int32_t loss_type, float grad_scale, float* loss, float* dpred, void* stream) {
  ibc_handle* h = nullptr;
  if (!predictions || !loss) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_cd_loss: null argument");
  return cd_loss(h, predictions, actions, B, n1, A, loss_type, grad_scale, loss, dpred,
                 (cudaStream_t)stream);
}

int ibc_grad_penalty_from_dact(const float* dE_dact, int64_t B, int64_t n1, int32_t A,
                               const ibc_loss_cfg* cfg, float grad_scale, float* grad_loss,
                               float* cotangent, void* stream) {
  ibc_handle* h = nullptr;
  if (!dE_dact || !cfg || !grad_loss)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_grad_penalty_from_dact: null argument");
  return grad_penalty_from_dact(h, dE_dact, B, n1, A, *cfg, grad_scale, grad_loss, cotangent,
                                (cudaStream_t)stream);
}

int ibc_train_grads(ibc_handle* h, const float* obs, int64_t B, const float* actions, int64_t N,
                    const ibc_loss_cfg* cfg, int64_t global_batch, float* per_example_loss,
                    float* grad_loss, float* energies, float* grads, void* stream) {
  IBC_ENTER(h);
  if (!obs || !actions || !cfg || !per_example_loss || !grads || B < 1 || N < 0 ||
      global_batch < 1)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_train_grads: bad argument");
  if (!h->params) IBC_FAIL(h, IBC_ERR_INVALID, "no parameters bound (ibc_bind_params)");
  if (cfg->softmax_temperature <= 0.f)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_train_grads: softmax_temperature must be > 0");
  const bool gp = cfg->add_grad_penalty != 0 && cfg->grad_norm_type != IBC_NORM_NONE;
  h->loss_event_recorded = false;
  if (h->precision == IBC_PRECISION_TF32 && umma_train_supported(h->L) && !gp)
    return train_grads_umma(h, obs, B, actions, N, *cfg, global_batch, per_example_loss,
                            grad_loss, energies, grads, (cudaStream_t)stream);
  // gradient penalty (second order) and other widths: fp32 path
  return train_grads_fp32(h, obs, B, actions, N, *cfg, global_batch, per_example_loss, grad_loss,
                          energies, grads, (cudaStream_t)stream);
}

int ibc_stream_wait_loss(ibc_handle* h, void* stream, int32_t* recorded) {
  IBC_ENTER(h);
  if (!recorded) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_stream_wait_loss: null argument");
  *recorded = (h->loss_event && h->loss_event_recorded) ? 1 : 0;
  if (*recorded) IBC_CUDA(h, cudaStreamWaitEvent((cudaStream_t)stream, h->loss_event, 0));
  return IBC_OK;
}

int ibc_adam_step(float* params, const float* grads, float* m, float* v, int64_t count, float lr_t,
                  float beta1, float beta2, float epsilon, void* stream) {
  ibc_handle* h = nullptr;
  if (!params || !grads || !m || !v || count < 0)
    IBC_FAIL(h, IBC_ERR_INVALID, "ibc_adam_step: bad argument");
  return adam_step(h, params, grads, m, v, count, lr_t, beta1, beta2, epsilon,
                   (cudaStream_t)stream);
}

int ibc_kernel_status(ibc_handle* h, int32_t* status_host) {
  IBC_ENTER(h);
  if (!status_host) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_kernel_status: null argument");
  int st = 0;
  IBC_CUDA(h, cudaDeviceSynchronize());
  IBC_CUDA(h, cudaMemcpy(&st, h->dev_status, sizeof(int), cudaMemcpyDeviceToHost));
  IBC_CUDA(h, cudaMemset(h->dev_status, 0, sizeof(int)));
  *status_host = st;
  return IBC_OK;
}

int ibc_set_profiling(ibc_handle* h, int32_t on) {
  IBC_ENTER(h);
  if (on && !h->ev[0])
    for (int i = 0; i < 6; ++i) IBC_CUDA(h, cudaEventCreate(&h->ev[i]));
  h->profiling = on != 0;
  h->ev_valid = false;
  return IBC_OK;
}

int ibc_get_train_timings(ibc_handle* h, float* ms_host) {
  IBC_ENTER(h);
  if (!ms_host) IBC_FAIL(h, IBC_ERR_INVALID, "ibc_get_train_timings: null argument");
  if (!h->profiling || !h->ev_valid)
    IBC_FAIL(h, IBC_ERR_INVALID, "no profiled tensor-core training step has run");
  IBC_CUDA(h, cudaEventSynchronize(h->ev[5]));
  for (int i = 0; i < 5; ++i) IBC_CUDA(h, cudaEventElapsedTime(&ms_host[i], h->ev[i], h->ev[i + 1]));
  return IBC_OK;
}

int64_t ibc_launch_count(void) { return (int64_t)ibc::g_launch_count.load(); }
void ibc_reset_launch_count(void) { ibc::g_launch_count.store(0); }

int ibc_umma_selftest(const float* A, const float* Bt, int32_t N, int32_t K, int32_t mode, float* D,
                      int32_t* status_host, void* stream) {
  return umma_selftest(A, Bt, N, K, mode, D, status_host, (cudaStream_t)stream);
}

int ibc_umma_f16_selftest(const float* A, const float* B, int32_t N, int32_t K, uint32_t lbo_a,
                          uint32_t lbo_b, uint32_t sbo, uint32_t layout_type, uint32_t kstep_bytes,
                          float* D, int32_t* status_host, void* stream) {
  return umma_f16_selftest(A, B, N, K, lbo_a, lbo_b, sbo, layout_type, kstep_bytes, D, status_host,
                           (cudaStream_t)stream);
}

int ibc_gemm_tf32(const float* A, int64_t lda, int32_t trans_a, const float* B, int64_t ldb,
                  int32_t trans_b, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                  const float* bias, const float* gate_c, const float* res, int32_t relu_a,
                  int32_t relu_out, int32_t accumulate, int32_t* status_host, void* stream) {
  ibc_handle* none = nullptr;
  if (!A || !B || !C || M < 1 || N < 1 || K < 1)
    IBC_FAIL(none, IBC_ERR_INVALID, "ibc_gemm_tf32: bad argument");
  // per-device state of this handle-less entry point: SM count and a status word
  constexpr int kMaxDev = 64;
  static int sm_count[kMaxDev] = {0};
  static int* dev_status[kMaxDev] = {nullptr};
  int dev = 0;
  IBC_CUDA(none, cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) IBC_FAIL(none, IBC_ERR_INVALID, "ibc_gemm_tf32: device index");
  if (!dev_status[dev]) {
    int major = 0, sms = 0;
    IBC_CUDA(none, cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    IBC_CUDA(none, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) IBC_FAIL(none, IBC_ERR_UNSUPPORTED, "ibc_b200 is built for sm_100a only");
    IBC_CUDA(none, cudaMalloc((void**)&dev_status[dev], sizeof(int)));
    IBC_CUDA(none, cudaMemset(dev_status[dev], 0, sizeof(int)));
    sm_count[dev] = sms;
  }
  ibc_handle tmp;                       // carries the error string, SM count and status word
  tmp.device = dev;
  tmp.sm_count = sm_count[dev];
  tmp.dev_status = dev_status[dev];
  ibc::GemmP g;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.bias = bias;
  g.gateC = gate_c; g.ldgc = ldc; g.res = res; g.ldres = ldc;
  g.aop = relu_a ? ibc::AOP_RELU : ibc::AOP_ID; g.relu_out = relu_out; g.atomic = accumulate;
  const int rc = ibc::gemm_tc(&tmp, g, trans_a != 0, trans_b != 0, (cudaStream_t)stream);
  if (rc != IBC_OK) {
    strncpy(ibc::g_create_error, tmp.err.empty() ? "ibc_gemm_tf32 failed" : tmp.err.c_str(), 511);
    return rc;
  }
  if (status_host) {                    // optional: waits for the product
    int st = 0;
    IBC_CUDA(none, cudaStreamSynchronize((cudaStream_t)stream));
    IBC_CUDA(none, cudaMemcpy(&st, dev_status[dev], sizeof(int), cudaMemcpyDeviceToHost));
    if (st) IBC_CUDA(none, cudaMemset(dev_status[dev], 0, sizeof(int)));
    *status_host = st;
  }
  return IBC_OK;
}

}  // extern "C"
