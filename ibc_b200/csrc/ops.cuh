// Internal C++ entry points behind the C-ABI (api.cu).
#pragma once

#include <functional>

#include "common.cuh"

namespace ibc {

int resample(ibc_handle* h, const float* logits, const double* uniforms, int64_t B, int64_t n,
             int64_t count, uint64_t seed, uint64_t rng_stream, int32_t* draws,
             int32_t* counts, int64_t* rows, cudaStream_t s);
int softmax_rows(ibc_handle* h, const float* logits, int64_t B, int64_t n, float temperature,
                 float* probs, cudaStream_t s);
int uniform_actions(ibc_handle* h, const float* u, int64_t rows, int A, const float* mn,
                    const float* mx, uint64_t seed, float* out, cudaStream_t s);
int peer_alloc(ibc_handle* h, int64_t bytes, void** ptr, void* handle64);
int peer_open(ibc_handle* h, const void* handle64, void** ptr);
int peer_close(ibc_handle* h, void* ptr, int owned);
int peer_allreduce_adam(ibc_handle* h, const float* const* peer_grads, int* const* peer_flags, int rank,
                        int world, int epoch, int64_t n, float* params, float* m, float* v, float lr_t,
                        float b1, float b2, float eps, float* tail_out, int* status, cudaStream_t s);
int greedy_actions(ibc_handle* h, const float* probs, const float* values, int64_t B, int64_t n, int A,
                   const float* mn, const float* mx, float* out, cudaStream_t s);
int uniform_combined(ibc_handle* h, const float* u, int64_t B, int64_t N, int A, const float* mn,
                     const float* mx, uint64_t seed, const float* actions, float* out,
                     cudaStream_t s);
int dfo(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n, const float* mn,
        const float* mx, float temperature, int num_iterations, float iteration_std,
        const double* uniforms, const float* normals, uint64_t seed, float* probs,
        cudaStream_t s);
int dfo_generic(ibc_handle* h, int A, int64_t B, float* actions, int64_t n, const float* mn,
                const float* mx, float temperature, int num_iterations, float iteration_std,
                const double* uniforms, const float* normals, uint64_t seed, float* probs,
                const std::function<int(const float*, float*)>& energy_fn, cudaStream_t s);
int gather_perturb(ibc_handle* h, const float* src, const int64_t* rows, int64_t R, int A,
                   const float* normals, uint64_t seed, uint64_t rng_stream, float std,
                   const float* mn, const float* mx, float* out, cudaStream_t s);
int langevin_update(ibc_handle* h, float* actions, const float* dEda, int64_t R, int A,
                    const float* normals, uint64_t seed, uint64_t rng_stream, float noise_scale,
                    float grad_clip, float delta_action_clip, float stepsize, const float* mn,
                    const float* mx, int norm_type, float* grad_norms, cudaStream_t s);
void langevin_stepsizes(const ibc_langevin_cfg& c, std::vector<float>* out);
int langevin(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
             const float* mn, const float* mx, const ibc_langevin_cfg& c, const float* normals,
             uint64_t seed, float* chain_actions, float* chain_energies, float* chain_grad_norms,
             cudaStream_t s);
int info_nce(ibc_handle* h, const float* pred, int64_t B, int64_t n1, float temperature,
             float grad_scale, float* loss, float* dpred, cudaStream_t s,
             unsigned int* dpred_absmax = nullptr);
int cd_loss(ibc_handle* h, const float* pred, const float* actions, int64_t B, int64_t n1, int A,
            int type, float grad_scale, float* loss, float* dpred, cudaStream_t s,
            unsigned int* dpred_absmax = nullptr);
int ebm_loss(ibc_handle* h, const ibc_loss_cfg& c, const float* pred, const float* actions, int64_t B,
             int64_t n1, int A, float grad_scale, float* loss, float* dpred, cudaStream_t s,
             unsigned int* dpred_absmax = nullptr);
int grad_penalty_from_dact(ibc_handle* h, const float* dEda, int64_t B, int64_t n1, int A,
                           const ibc_loss_cfg& c, float grad_scale, float* grad_loss, float* cot,
                           cudaStream_t s);
int adam_step(ibc_handle* h, float* p, const float* g, float* m, float* v, int64_t n, float lr_t,
              float b1, float b2, float eps, cudaStream_t s);
int train_grads_fp32(ibc_handle* h, const float* obs, int64_t B, const float* actions, int64_t N,
                     const ibc_loss_cfg& c, int64_t global_batch, float* per_example,
                     float* grad_loss_out, float* energies_out, float* grads, cudaStream_t s);
int umma_selftest(const float* A, const float* Bt, int N, int K, int mode, float* D,
                  int* status_host, cudaStream_t s);
int umma_f16_selftest(const float* A, const float* B, int N, int K, uint32_t lbo_a, uint32_t lbo_b,
                      uint32_t sbo, uint32_t layout_type, uint32_t kstep_bytes, float* D,
                      int* status_host, cudaStream_t s);

}  // namespace ibc
