// Dense-stack passes of the MLPEBM (networks/mlp_ebm.py:72-96,
// networks/layers/resnet.py:135-153), composed per layer for the fp32 path and
// dispatched to the fused tcgen05 kernels for IBC_PRECISION_TF32.
#pragma once

#include "common.cuh"

namespace ibc {

// Saved state of one forward over R rows (all device pointers into the
// handle's workspace): H[j] (j <= nb) is the residual stream entering block j
// (H[nb] feeds Dense(1)); V[j] is block j's hidden pre-activation.
struct FwdSave {
  float* hp = nullptr;   // [B, W]  obs . Wp[:O] + bp (hoisted per observation)
  float* H = nullptr;    // [(nb+1), R, W]
  float* V = nullptr;    // [nb, R, W]
  int64_t R = 0;
  float* Hj(int j, int W) const { return H + (int64_t)j * R * W; }
  float* Vj(int j, int W) const { return V + (int64_t)j * R * W; }
};

// Reverse-chain state: G[j] = dE/dH[j] (j <= nb), U[j] = masked dE/dV[j].
struct RevSave {
  float* G = nullptr;    // [(nb+1), R, W]
  float* U = nullptr;    // [nb, R, W]
  int64_t R = 0;
  float* Gj(int j, int W) const { return G + (int64_t)j * R * W; }
  float* Uj(int j, int W) const { return U + (int64_t)j * R * W; }
};

size_t fwd_save_bytes(const Layout& L, int64_t B, int64_t R);
void carve_fwd_save(ibc_handle* h, int64_t B, int64_t R, FwdSave* out);

// fp32 path --------------------------------------------------------------
// E[R]; when `save` is non-null its buffers are filled, otherwise scratch
// buffers are carved from the workspace (which the caller must have reserved).
int mlp_forward_fp32(ibc_handle* h, const float* obs, int64_t B, const float* act,
                     int64_t n, float* E, FwdSave* save, cudaStream_t s);
// dE/dact [R, A] from a saved forward; row_scale (nullable) multiplies the
// output cotangent per row (apply_exp).  rev (nullable) keeps G/U.
int mlp_reverse_fp32(ibc_handle* h, const FwdSave& f, const float* row_scale,
                     float* dEda, RevSave* rev, cudaStream_t s);
// E and dE/dact in one call (workspace reserved and carved inside).
int mlp_energy_dact_fp32(ibc_handle* h, const float* obs, int64_t B, const float* act,
                         int64_t n, float* E, float* dEda, bool apply_exp, cudaStream_t s);

// Dispatchers honouring h->precision ---------------------------------------
int mlp_energy(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
               float* E, cudaStream_t s);
int mlp_energy_dact(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                    float* E, float* dEda, bool apply_exp, cudaStream_t s);

// tcgen05 path (mlp_umma.cu) ---------------------------------------------
bool umma_supported(const Layout& L);
int umma_pack_params(ibc_handle* h, cudaStream_t s);   // every operand copy up to date
bool train_forward_is_f16(const Layout& L);            // the saving forward runs in kind::f16 (mlp_fwd_halves.cu)
int umma_pack_core(ibc_handle* h, cudaStream_t s);     // ... only those of the layer-fused training step
// Saved state for the reverse chain (both nullable): xa_save = the A operands of
// every layer (umma_save_floats() floats; the weight-gradient kernel's input),
// mask_save = their ReLU masks as bits (umma_mask_words() words; all the
// reverse chain itself reads).
size_t umma_save_floats(const Layout& L, int64_t R);
size_t umma_mask_words(const Layout& L, int64_t R);
int umma_forward(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                 float* hp, float* E, float* xa_save, uint32_t* mask_save, cudaStream_t s);
int umma_forward_hp(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp,
                    float* E, float* xa_save, uint32_t* mask_save, cudaStream_t s);
int langevin_umma(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
                  const float* mn, const float* mx, const ibc_langevin_cfg& c,
                  const float* normals, uint64_t seed, const std::vector<float>& steps,
                  float* chain_actions, float* chain_energies, float* chain_grad_norms,
                  cudaStream_t s);
// widths that are a multiple of 128 above 256: one tcgen05 GEMM kernel per layer (mlp_wide.cu)
bool wide_supported(const Layout& L);
int wide_pack_if_stale(ibc_handle* h, cudaStream_t s);
int mlp_energy_wide(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                    float* E, float* dEda, bool apply_exp, cudaStream_t s);
// the whole Langevin chain of a width-128 stack as one persistent kernel (langevin_tmem.cu)
// width 128, act_dim <= 8: the same chain with kind::f16 GEMMs and sixteen epilogue warps
// (langevin_f16.cu); pack16_params refreshes its fp16 weight images
bool langevin_f16_supported(const Layout& L);
int pack16_params(ibc_handle* h, cudaStream_t s, bool core);     // core: the transposed layers; else the rest
int langevin_f16(ibc_handle* h, const float* hp, int64_t B, float* actions, int64_t n,
                 const float* mn, const float* mx, const ibc_langevin_cfg& c, const float* normals,
                 uint64_t seed, const uint64_t* seed_ptr, const float* steps_dev,
                 float* chain_actions, float* chain_energies, float* chain_grad_norms,
                 cudaStream_t s);
bool langevin_fused_supported(const Layout& L);
int langevin_fused(ibc_handle* h, const float* hp, int64_t B, float* actions, int64_t n,
                   const float* mn, const float* mx, const ibc_langevin_cfg& c,
                   const float* normals, uint64_t seed, const uint64_t* seed_ptr,
                   const float* steps_dev, float* chain_actions, float* chain_energies,
                   float* chain_grad_norms, cudaStream_t s);
// the same for width 256, depth 2 (langevin256.cu); cvec: W + 1 floats of scratch
bool langevin256_supported(const Layout& L);
int langevin256_fused(ibc_handle* h, const float* hp, int64_t B, float* actions, int64_t n,
                      const float* mn, const float* mx, const ibc_langevin_cfg& c,
                      const float* normals, uint64_t seed, const uint64_t* seed_ptr,
                      const float* steps_dev, float* cvec, float* chain_actions,
                      float* chain_energies, float* chain_grad_norms, cudaStream_t s);
// width-128 kernels with the activations in tensor memory (mlp_umma_tmem.cu); tile
// ids and the saved-tile layouts are the same as the shared-memory-operand kernels'
bool tmem_path_supported(const Layout& L);
int tmem_forward(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp, float* E,
                 float* xa_save, uint32_t* mask_save, cudaStream_t s);
int tmem_forward_f16(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp, float* E,
                     float* xa_save, void* xa16, float* hn_save, uint32_t* mask_save,
                     cudaStream_t s);
// energies only (xa16 == null) or the fp16-saving forward, eight epilogue warps per tile
// (mlp_fwd_halves.cu)
int tmem_forward_halves(ibc_handle* h, const float* act, int64_t R, int64_t n, const float* hp,
                        float* E, void* xa16, float* hn_save, uint32_t* mask_save, cudaStream_t s);
int tmem_reverse(ibc_handle* h, const float* dE, int64_t R, const float* xa_save,
                 const uint32_t* mask_save, float* gy_save, float* g0, float* dwe,
                 cudaStream_t s);
// reverse chain + weight / bias gradients of the W x W layers in one block-major
// kernel (mlp_bwd_wgrad.cu); scratch sizes in floats
size_t bwd_wgrad_scratch_floats(const ibc_handle* h, int64_t R, size_t* park, size_t* accw);
int tmem_reverse_wgrad(ibc_handle* h, const float* dE, const uint32_t* dE_absmax, int64_t R,
                       const void* xa16, const float* hn, const uint32_t* mask_save, float* gpark,
                       float* acc_w, float* grads, cudaStream_t s);
// tcgen05 training path (mlp_umma_train.cu): W == 128, no gradient penalty.
bool umma_train_supported(const Layout& L);
int train_grads_umma(ibc_handle* h, const float* obs, int64_t B, const float* actions, int64_t N,
                     const ibc_loss_cfg& c, int64_t global_batch, float* per_example,
                     float* grad_loss_out, float* energies_out, float* grads, cudaStream_t s);
int umma_reverse(ibc_handle* h, const float* dE, int64_t R, const float* xa_save,
                 const uint32_t* mask_save, float* gy_save,
                 float* g0, float* dwe, cudaStream_t s);
int mlp_energy_dact_umma(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                         float* E, float* dEda, bool apply_exp, cudaStream_t s);
int mlp_energy_umma(ibc_handle* h, const float* obs, int64_t B, const float* act, int64_t n,
                    float* E, float* dEda, bool apply_exp, cudaStream_t s);

}  // namespace ibc
