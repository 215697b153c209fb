/* ibc_b200 -- C-ABI of the B200-native IBC energy-based-policy hot path.
 *
 * The reference (google-research/ibc) is pure Python on TensorFlow and has no
 * FFI layer; its "operator API" for this path is the Python call protocol of
 *   ibc/agents/mcmc.py, ibc/losses/ebm_loss.py, ibc/losses/gradient_loss.py,
 *   networks/mlp_ebm.py, networks/pixel_ebm.py, ibc/agents/base_agent.py.
 * Each entry point below names the reference function (file:line) whose tensor
 * arithmetic it replaces.  The Python package `ibc_b200.ibc` re-hosts those
 * signatures on top of this library through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *     the library borrows caller memory and never frees it;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     no call synchronises the host unless documented;
 *   - all functions return 0 on success or a negative IBC_ERR_* code; the
 *     message is available from ibc_last_error();
 *   - float = IEEE fp32, double = fp64, row-major dense tensors;
 *   - B observations are tiled against n candidate actions per observation the
 *     way tf-agents nest_utils.tile_batch does: row r of an [R = B*n, .] tensor
 *     belongs to observation r / n.  Observations are passed
 *     UN-tiled ([B, O]); the tiling is implicit.
 *
 * Flat parameter vector of an MLPEBM (ResNetPreActivation body, relu, no
 * normaliser, dropout 0; networks/mlp_ebm.py:29-96, layers/resnet.py:32-153),
 * Keras kernels are [in, out] row-major:
 *   Wp [O + A, W] (observation rows first), bp [W],
 *   for each of depth/2 blocks: W1 [W, W], b1 [W], W2 [W, W], b2 [W],
 *   we [W] (Dense(1) kernel), be [1].
 */
#ifndef IBC_B200_H_
#define IBC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBC_OK 0
#define IBC_ERR_INVALID (-1)     /* bad argument (ValueError / assert upstream) */
#define IBC_ERR_UNSUPPORTED (-2) /* shape or option this build has no kernel for */
#define IBC_ERR_CUDA (-3)        /* CUDA runtime error, message has the detail */
#define IBC_ERR_NOT_FINITE (-4)  /* loss is inf/nan (base_agent.py:46)         */

/* Arithmetic used by the dense layers. */
#define IBC_PRECISION_FP32 0 /* fp32 FMA kernels (exact-order-independent)     */
#define IBC_PRECISION_TF32 1 /* tcgen05 kind::tf32 MMA, fp32 accumulate in TMEM */

#define IBC_NORM_NONE 0 /* grad_norm_type=None -> zeros (mcmc.py:203-206) */
#define IBC_NORM_L1 1
#define IBC_NORM_L2 2
#define IBC_NORM_INF 3

typedef struct ibc_handle ibc_handle;

typedef struct {
  int32_t obs_dim; /* O = T * sum of observation feature dims, flattened       */
  int32_t act_dim; /* A                                                        */
  int32_t width;   /* W, MLPEBM.width                                          */
  int32_t depth;   /* D, MLPEBM.depth (even; resnet.py:42)                     */
} ibc_mlp_arch;

/* langevin_actions_given_obs keyword arguments (mcmc.py:332-355). */
typedef struct {
  int32_t num_iterations;        /* default 25                                 */
  float sampler_stepsize_init;   /* 1e-1                                       */
  float sampler_stepsize_decay;  /* 0.8 (ExponentialSchedule, mcmc.py:267-278) */
  float noise_scale;             /* 1.0                                        */
  float grad_clip;               /* <= 0 means None                            */
  float delta_action_clip;       /* 0.1                                        */
  int32_t use_polynomial_rate;   /* 1                                          */
  float sampler_stepsize_final;  /* 1e-5                                       */
  float sampler_stepsize_power;  /* 2.0                                        */
  int32_t grad_norm_type;        /* IBC_NORM_*, default INF                    */
  int32_t apply_exp;             /* 0; 1 differentiates exp(E) (mcmc.py:186)   */
} ibc_langevin_cfg;

/* ImplicitBCAgent._loss options (ibc_agent.py:43-66, gradient_loss.py:24-34). */
typedef struct {
  float softmax_temperature;   /* 1.0                                          */
  int32_t add_grad_penalty;    /* 1                                            */
  int32_t grad_norm_type;      /* IBC_NORM_INF                                 */
  float grad_margin;           /* 1.0; < 0 means None                          */
  int32_t square_grad_penalty; /* 1                                            */
  float grad_loss_weight;      /* 1.0                                          */
  int32_t ebm_loss_type;       /* IBC_LOSS_*: ImplicitBCAgent(ebm_loss_type=...) (ibc_agent.py:303-335) */
} ibc_loss_cfg;

/* ibc/losses/ebm_loss.py: info_nce (:21-48), cd (:51-66), clipped_cd with soft = False / True
 * (:109-148; ibc_agent.py 'clipped_cd' / 'soft_clipped_cd').  'cd_kl' (:69-106) needs a second
 * network and gradients through the sampler: not built (IBC_ERR_UNSUPPORTED). */
#define IBC_LOSS_INFO_NCE 0
#define IBC_LOSS_CD 1
#define IBC_LOSS_CLIPPED_CD 2
#define IBC_LOSS_SOFT_CLIPPED_CD 3

/* ---- lifetime ---------------------------------------------------------- */

/* Creates the per-GPU state that owns workspaces and packed weight copies for
 * one MLPEBM (networks/mlp_ebm.py:29-70).  depth must be even. */
int ibc_create(const ibc_mlp_arch* arch, int device, ibc_handle** out);
int ibc_destroy(ibc_handle* h);
/* Last error message of `h` (or of the last failed ibc_create when h==NULL). */
const char* ibc_last_error(const ibc_handle* h);
/* Number of fp32 parameters in the flat vector described above. */
int64_t ibc_param_count(const ibc_mlp_arch* arch);
/* Library build tag, e.g. "ibc_b200 0.1 sm_100a". */
const char* ibc_version(void);

/* IBC_PRECISION_*; TF32 is refused (IBC_ERR_UNSUPPORTED) for widths the
 * tcgen05 kernels do not cover. */
int ibc_set_precision(ibc_handle* h, int precision);
int ibc_get_precision(const ibc_handle* h);

/* Borrow the caller's flat parameter vector (device, fp32).  Must be called
 * again (or ibc_params_changed) whenever the values change so packed copies
 * are refreshed.  Replaces the Keras variable ownership of
 * networks/mlp_ebm.py:41-70. */
int ibc_bind_params(ibc_handle* h, const float* params, void* stream);
int ibc_params_changed(ibc_handle* h);

/* ---- energy network ---------------------------------------------------- */

/* MLPEBM.call (networks/mlp_ebm.py:72-96 + layers/resnet.py:135-153):
 * energy[r] = E(obs[r / n], act[r]) for r in [0, B*n). */
int ibc_mlp_energy(ibc_handle* h, const float* obs, int64_t B,
                   const float* act, int64_t n, float* energy, void* stream);

/* mcmc.gradient_wrt_act (ibc/agents/mcmc.py:161-191) without the sign flip:
 * energy[r] and dE_dact[r, :] = +dE/dact (the reference returns -dE/dact). */
int ibc_mlp_energy_dact(ibc_handle* h, const float* obs, int64_t B,
                        const float* act, int64_t n, float* energy,
                        float* dE_dact, void* stream);

/* ---- samplers ---------------------------------------------------------- */

/* categorical_bincount + tf.repeat (mcmc.py:32-55, 136-137), bit-exact with
 * TF's CPU multinomial given the same uniforms: sequential fp64 CDF of
 * exp(logit - max), to_find = u * total, upper_bound.
 *   logits [B, n] f32; uniforms [B, count] f64 in [0, 1);
 *   draws  [B, count] i32 raw class ids        (nullable)
 *   counts [B, n] i32                          (nullable)
 *   rows   [B * count] i64 global row ids b*n + class, ascending (nullable) */
int ibc_resample(const float* logits, const double* uniforms, int64_t B,
                 int64_t n, int64_t count, int32_t* draws, int32_t* counts,
                 int64_t* rows, void* stream);

/* mcmc.iterative_dfo (mcmc.py:58-158).  actions [B*n, A] is updated in place;
 * probs [B*n] receives softmax(last logits) (NOT aligned with the resampled
 * actions, as upstream).  uniforms [num_iterations, B, n] f64 and normals
 * [num_iterations-1, B*n, A] f32 are the explicit draws; when NULL they are
 * generated on device from `seed` (Philox4x32-10). */
int ibc_dfo(ibc_handle* h, const float* obs, int64_t B, float* actions,
            int64_t n, const float* min_actions, const float* max_actions,
            float temperature, int32_t num_iterations, float iteration_std,
            const double* uniforms, const float* normals, uint64_t seed,
            float* probs, void* stream);

/* mcmc.langevin_actions_given_obs (mcmc.py:330-425) with stop_chain_grad.
 * actions [B*n, A] updated in place.  normals [K, B*n, A] or NULL (-> seed).
 * Optional chain outputs (update_chain_data, mcmc.py:297-327):
 *   chain_actions [K, R, A] post-step, chain_energies [K, R] and
 *   chain_grad_norms [K, R] pre-step. */
int ibc_langevin(ibc_handle* h, const float* obs, int64_t B, float* actions,
                 int64_t n, const float* min_actions, const float* max_actions,
                 const ibc_langevin_cfg* cfg, const float* normals,
                 uint64_t seed, float* chain_actions, float* chain_energies,
                 float* chain_grad_norms, void* stream);

/* Pieces of the two samplers for energy functions that are not an MLPEBM owned
 * by a handle (e.g. the mock network of ibc/agents/mcmc_test.py:67-81), so the
 * host can interleave its own energy evaluation:
 *   out[r, :] = clip(src[rows[r], :] + normal * std, min, max)   (std == 0:
 *   plain gather)                                   mcmc.py:139-150
 *   one langevin_step update after dE/dact is known mcmc.py:238-262
 * normals NULL -> generated from seed. */
int ibc_dfo_gather_perturb(const float* src, const int64_t* rows, int64_t R,
                           int32_t A, const float* normals, uint64_t seed,
                           float std, const float* min_actions,
                           const float* max_actions, float* out, void* stream);
int ibc_langevin_update(float* actions, const float* dE_dact, int64_t R,
                        int32_t A, const float* normals, uint64_t seed,
                        float noise_scale, float grad_clip,
                        float delta_action_clip, float stepsize,
                        const float* min_actions, const float* max_actions,
                        int32_t grad_norm_type, float* grad_norms, void* stream);

/* PolynomialSchedule / ExponentialSchedule (mcmc.py:267-294): step size used
 * AT each of the K steps, written to out_host[K] (host memory). */
int ibc_langevin_stepsizes(const ibc_langevin_cfg* cfg, float* out_host);

/* mcmc.get_probabilities (mcmc.py:428-441) given the logits: row softmax of
 * logits / temperature over n, flattened. */
int ibc_softmax_rows(const float* logits, int64_t B, int64_t n,
                     float temperature, float* probs, void* stream);

/* tensor_spec.sample_spec_nest on a bounded float spec (ibc_agent.py:350-352,
 * ibc_policy.py:263-265): out = u * (max - min) + min with u [rows, A] given
 * or (NULL) generated from seed. */
int ibc_uniform_actions(const float* u, int64_t rows, int32_t A,
                        const float* min_actions, const float* max_actions,
                        uint64_t seed, float* out, void* stream);

/* Data-parallel training over NVLink peer memory: the MirroredStrategy gradient sum
 * (ibc/agents/base_agent.py:50-58, ibc/train_eval.py:163) and the Adam update
 * (ibc/train/get_agent.py:76-84) in one launch.  peer_grads[r] / peer_flags[r] are rank r's
 * gradient buffer of this step's parity and its flag array (>= world ints, zero at start) as
 * mapped into THIS process (CUDA IPC); epoch increases by one per call on every rank.  The kernel
 * publishes the epoch to every peer, waits (bounded: status 6001) for all peers, sums the
 * gradients in rank order and applies Adam to the local replica; tail_out (optional; device or
 * pinned host memory) receives the reduced last entry.
 * ibc_peer_alloc makes one zeroed device allocation on the current device and writes its 64-byte
 * CUDA IPC handle to handle64; ibc_peer_open maps a peer process's handle into the current
 * device's address space (peer access over NVLink); ibc_peer_close unmaps (owned = 0) or frees
 * (owned = 1).  The handles travel between the ranks by any host channel (the process group). */
int ibc_peer_alloc(int64_t bytes, void** ptr, void* handle64);
int ibc_peer_open(const void* handle64, void** ptr);
int ibc_peer_close(void* ptr, int32_t owned);
int ibc_peer_allreduce_adam(const float* const* peer_grads, int32_t* const* peer_flags,
                            int32_t rank, int32_t world, int32_t epoch, int64_t count,
                            float* params, float* m, float* v, float lr_t, float beta1,
                            float beta2, float epsilon, float* tail_out, int32_t* status,
                            void* stream);

/* greedy_policy.GreedyPolicy over the MappedCategorical of ibc/agents/ibc_policy.py:96-117:
 * out[b, :] = values[b, argmax_j probs[b, j], :] (first index of the maximum; NaN counts as the
 * maximum, as in tf.argmax), clipped to [min_actions, max_actions] when given (both or neither;
 * ibc_policy.py clip=True). */
int ibc_greedy_actions(const float* probs, const float* values, int64_t B, int64_t n, int32_t A,
                       const float* min_actions, const float* max_actions, float* out,
                       void* stream);

/* The counter-example block of a training step with uniform negatives only
 * (ibc_agent.py:340-352 and 461-473 in one pass): combined [B, N + 1, A] with
 * combined[b, j, :] = the sample ibc_uniform_actions draws for row b * N + j of a
 * [B * N, A] call with the same u / seed (j < N) and combined[b, N, :] = actions[b, :]
 * (the expert action last). */
int ibc_counter_examples_uniform(const float* u, int64_t B, int64_t N, int32_t A,
                                 const float* min_actions, const float* max_actions,
                                 uint64_t seed, const float* actions, float* combined,
                                 void* stream);

/* ---- losses and training ---------------------------------------------- */

/* ebm_loss.info_nce (ibc/losses/ebm_loss.py:21-48) in the Keras KLDivergence
 * epsilon-clipped form.  predictions [B, n1] with the true action in the LAST
 * column.  loss [B]; dpred [B, n1] (nullable) = grad_scale * dloss/dpred. */
int ibc_info_nce(const float* predictions, int64_t B, int64_t n1,
                 float temperature, float grad_scale, float* loss,
                 float* dpred, void* stream);

/* The contrastive-divergence losses of ibc/losses/ebm_loss.py on predictions [B, n1] (true action
 * in the LAST column): loss_type IBC_LOSS_CD (cd, :51-66: mean of the counter examples'
 * predictions minus the true one), IBC_LOSS_CLIPPED_CD / IBC_LOSS_SOFT_CLIPPED_CD (clipped_cd,
 * :109-148: counter examples within squared distance 0.2 of the true action drop out of the
 * difference; soft adds mean |(pred_true - pred_j) - err_j| over those inside).  actions
 * [B*n1, A] (the combined counter-example + true actions; only read by the clipped forms).
 * loss [B]; dpred [B, n1] (nullable) = grad_scale * dloss/dpred. */
int ibc_cd_loss(const float* predictions, const float* actions, int64_t B, int64_t n1, int32_t A,
                int32_t loss_type, float grad_scale, float* loss, float* dpred, void* stream);

/* gradient_loss.grad_penalty given dE/dact (gradient_loss.py:49-72):
 * grad_loss [B]; cotangent [B*n1, A] (nullable) = grad_scale * d(sum_b
 * grad_loss[b]) / d(dE_dact). */
int ibc_grad_penalty_from_dact(const float* dE_dact, int64_t B, int64_t n1,
                               int32_t A, const ibc_loss_cfg* cfg,
                               float grad_scale, float* grad_loss,
                               float* cotangent, void* stream);

/* ImplicitBCAgent._loss + tape.gradient (ibc_agent.py:132-300,
 * base_agent.py:43-48) for given combined actions [B*(N+1), A] (true action
 * last per observation):
 *   per_example_loss [B]  = info_nce | cd | clipped_cd (cfg->ebm_loss_type) (+ grad_penalty)
 *   grad_loss        [B]  (nullable; zeros when the penalty is off)
 *   energies   [B*(N+1)]  (nullable)
 *   grads [param_count]   = d( sum_b per_example_loss[b] / global_batch )/dtheta
 * (common.aggregate_losses scaling, ibc_agent.py:246-250).  grads is
 * overwritten. */
int ibc_train_grads(ibc_handle* h, const float* obs, int64_t B,
                    const float* actions, int64_t N, const ibc_loss_cfg* cfg,
                    int64_t global_batch, float* per_example_loss,
                    float* grad_loss, float* energies, float* grads,
                    void* stream);

/* Reading the loss of a step does not have to wait for the step (in TF eager, reading
 * loss_info.loss waits for the op that produced it, not for the optimizer update that follows:
 * ibc/agents/base_agent.py:43-59).  On the layer-fused path the last ibc_train_grads records an
 * event on its stream right behind its loss kernel, when per_example_loss is final and the
 * backward is still to run.  ibc_stream_wait_loss sets *recorded (0: the path taken recorded
 * none) and, if 1, makes `stream` wait for that event, so that the caller can sum and copy the
 * loss on a second stream. */
int ibc_stream_wait_loss(ibc_handle* h, void* stream, int32_t* recorded);

/* Keras-2.6 Adam (get_agent.py:64-68): m, v, params updated in place with the
 * bias-corrected step size lr_t computed by the caller. */
int ibc_adam_step(float* params, const float* grads, float* m, float* v,
                  int64_t count, float lr_t, float beta1, float beta2,
                  float epsilon, void* stream);

/* ---- late-fusion PixelEBM ---------------------------------------------- */

/* networks/pixel_ebm.py:46-125 with encoder_network='ConvMaxpoolEncoder'
 * (networks/layers/conv_maxpool.py:20-48) and value_network='DenseResnetValue'
 * (networks/layers/dense_resnet_value.py:22-69).  Flat parameter vector:
 * conv{1..4}: kernel [3,3,cin,cout] (Keras layout) + bias [cout], channels
 * C*T -> 32 -> 64 -> 128 -> 256; value dense0 [256 + A, Wv] + b [Wv]; per block
 * d0 [Wv, Wv/4] + b, d1 [Wv/4, Wv/4] + b, d2 [Wv/4, Wv] + b; dense1 [Wv] + b [1]. */
typedef struct ibc_pixel_handle ibc_pixel_handle;

typedef struct {
  int32_t seq_len;       /* T: stacked frames (obs_spec['rgb'].shape[0])           */
  int32_t height, width; /* raw frame size (240 x 320)                             */
  int32_t channels;      /* 3                                                      */
  int32_t target_height; /* PixelEBM.target_height (180)                           */
  int32_t target_width;  /* PixelEBM.target_width (240)                            */
  int32_t act_dim;       /* A                                                      */
  int32_t value_width;   /* DenseResnetValue.width (multiple of 4)                 */
  int32_t value_blocks;  /* DenseResnetValue.num_blocks                            */
} ibc_pixel_arch;

int ibc_pixel_create(const ibc_pixel_arch* arch, int device, ibc_pixel_handle** out);
int ibc_pixel_destroy(ibc_pixel_handle* h);
const char* ibc_pixel_last_error(const ibc_pixel_handle* h);
int64_t ibc_pixel_param_count(const ibc_pixel_arch* arch);
int ibc_pixel_bind_params(ibc_pixel_handle* h, const float* params);
/* IBC_PRECISION_TF32 (default): convolutions (as im2col rows), value head and all
 * their data / weight gradients on the tcgen05 kind::tf32 GEMM; IBC_PRECISION_FP32:
 * the same contractions on fp32 FMA kernels. */
int ibc_pixel_set_precision(ibc_pixel_handle* h, int precision);
int ibc_pixel_get_precision(const ibc_pixel_handle* h);
/* As ibc_kernel_status, for the kernels launched through a pixel handle. */
int ibc_pixel_kernel_status(ibc_pixel_handle* h, int32_t* status_host);

/* PixelEBM.encode (pixel_ebm.py:79-96): images [B, T, H, W, C] uint8 (is_uint8)
 * or float32 in [0,1] -> u8/255, RAW reshape to [B, H, W, C*T]
 * (image_prepro.py:28), bilinear resize (half-pixel centres), 4 x (conv3x3 same
 * + ReLU + maxpool2), global average pool -> enc [B, 256]. */
int ibc_pixel_encode(ibc_pixel_handle* h, const void* images, int32_t is_uint8,
                     int64_t B, float* enc, void* stream);

/* PixelEBM.call with observation_encoding (pixel_ebm.py:113-125): energy[r] of
 * concat(enc[r / n], act[r]) through DenseResnetValue; enc is un-tiled. */
int ibc_pixel_energy(ibc_pixel_handle* h, const float* enc, int64_t B,
                     const float* act, int64_t n, float* energy, void* stream);

/* mcmc.iterative_dfo with late_fusion=True (mcmc.py:98-110): the observation
 * is encoded once by the caller (ibc_pixel_encode); same draws contract as
 * ibc_dfo. */
int ibc_pixel_dfo(ibc_pixel_handle* h, const float* enc, int64_t B, float* actions,
                  int64_t n, const float* min_actions, const float* max_actions,
                  float temperature, int32_t num_iterations, float iteration_std,
                  const double* uniforms, const float* normals, uint64_t seed,
                  float* probs, void* stream);

/* mcmc.gradient_wrt_act on the late-fusion network (mcmc.py:161-191 with
 * PixelEBM.call's observation_encoding, pixel_ebm.py:113-125): energy [B*n] and
 * dE_dact [B*n, A] = +dE/dact (the caller negates, as gradient_wrt_act returns
 * -dE/dact) for enc [B, 256] un-tiled and act [B*n, A]. */
int ibc_pixel_energy_dact(ibc_pixel_handle* h, const float* enc, int64_t B,
                          const float* act, int64_t n, float* energy,
                          float* dE_dact, void* stream);

/* mcmc.langevin_actions_given_obs with late_fusion=True (mcmc.py:384-388: the
 * observation is encoded once -- here by the caller, ibc_pixel_encode -- and
 * every step evaluates the value head only); same contract as ibc_langevin
 * (in-place actions, normals [K, B*n, A] or NULL -> seed, optional chain
 * outputs).  cfg->apply_exp must be 0. */
int ibc_pixel_langevin(ibc_pixel_handle* h, const float* enc, int64_t B,
                       float* actions, int64_t n, const float* min_actions,
                       const float* max_actions, const ibc_langevin_cfg* cfg,
                       const float* normals, uint64_t seed, float* chain_actions,
                       float* chain_energies, float* chain_grad_norms,
                       void* stream);

/* Late-fusion ImplicitBCAgent._loss + tape.gradient (ibc_agent.py:191-250):
 * encode (training=True), tile the encoding N+1 times, value head, InfoNCE,
 * backward through the value head and the conv encoder.  With
 * cfg->add_grad_penalty (pushing_pixels/pixel_ebm_langevin.gin) the gradient
 * penalty of gradient_loss.py:35-72 on dE/dact at the combined actions is added
 * to per_example_loss and its parameter gradients (value head only: dE/dact
 * depends on the encoder through ReLU masks alone) to grads; grad_loss [B]
 * (nullable) receives the penalty term.  grads [param_count] is overwritten. */
int ibc_pixel_train_grads(ibc_pixel_handle* h, const void* images, int32_t is_uint8,
                          int64_t B, const float* actions, int64_t N,
                          const ibc_loss_cfg* cfg, int64_t global_batch,
                          float* per_example_loss, float* grad_loss,
                          float* energies, float* grads, void* stream);

/* ibc_pixel_train_grads in two calls, so that the counter-example sampler can run on the
 * encoding of the SAME encoder pass the backward uses (the reference encodes twice:
 * once inside langevin_actions_given_obs, mcmc.py:384-388, once in _loss,
 * ibc_agent.py:191-197; the encoder has no dropout / batch norm, so both passes
 * give the same [B, 256]).  begin runs encode (training=True) keeping its
 * activations in the handle and writes enc_out [B, 256] (nullable); finish takes
 * the combined actions [B*(N+1), A] and does the rest of ibc_pixel_train_grads.
 * Between the two only the sampler entry points (ibc_pixel_energy / _energy_dact /
 * _dfo / _langevin) may be called on the handle; ibc_pixel_encode, _bind_params,
 * _set_precision or another begin discard the open step, and finish then fails
 * with IBC_ERR_INVALID. */
int ibc_pixel_train_begin(ibc_pixel_handle* h, const void* images, int32_t is_uint8,
                          int64_t B, int64_t N, float* enc_out, void* stream);
int ibc_pixel_train_finish(ibc_pixel_handle* h, const float* actions,
                           const ibc_loss_cfg* cfg, int64_t global_batch,
                           float* per_example_loss, float* grad_loss,
                           float* energies, float* grads, void* stream);

/* ---- diagnostics ------------------------------------------------------- */

/* Number of kernels this library has launched since the handle was created
 * (bench.py's gpu_launches) and a reset. */
int64_t ibc_launch_count(void);
void ibc_reset_launch_count(void);

/* Per-kernel timing of the tensor-core training step (ibc_train_grads on the
 * TF32 path): with profiling on, CUDA events are recorded on the call's stream
 * around its phases; ibc_get_train_timings waits for the last step and writes
 * ms_host[5] = {fused forward, InfoNCE, fused reverse chain, weight-gradient
 * GEMM, first/last-layer fp32 pieces} in milliseconds. */
int ibc_set_profiling(ibc_handle* h, int32_t on);
int ibc_get_train_timings(ibc_handle* h, float* ms_host);

/* Synchronises the device and returns (then clears) the sticky status word the
 * fused kernels write when one of their internal barriers timed out (0 = ok;
 * otherwise the code of the first wait that gave up -- results are invalid). */
int ibc_kernel_status(ibc_handle* h, int32_t* status_host);

/* tcgen05 self test: D[128, N] = A[128, K] * Bt[N, K]^T through the same
 * shared-memory operand layout, descriptors and TMEM epilogue the fused
 * kernels use.  mode 0 = A from shared memory, 1 = A from tensor memory.
 * status_host[0] receives 0, or a non-zero stage code if a barrier timed out. */
int ibc_umma_selftest(const float* A, const float* Bt, int32_t N, int32_t K,
                      int32_t mode, float* D, int32_t* status_host,
                      void* stream);

/* tcgen05 kind::f16 self test of the MN-major operand image the weight-gradient GEMM
 * uses (tape.gradient of ibc/agents/base_agent.py:43-48: dW = X^T dY with the rows as
 * the K dimension): D[128, N] = A[K, 128]^T . B[K, N], operands rounded to fp16, fp32
 * accumulate.  lbo / sbo / layout_type / kstep_bytes are the shared-memory descriptor
 * fields (0 = the library's own: umma_f16.cuh).
 * layout_type >= 100 tests the K-major forms the kind::f16 chain GEMMs use (the dependent
 * layers of mcmc.py:161-191 in langevin_f16.cu and mlp_bwd_wgrad.cu): D[128, N] =
 * A[128, K] . B[N, K]^T with A, B given row-major (N <= 128), B a no-swizzle K-major image in
 * shared memory; 100: A the same kind of image (SS form); 101: A in tensor memory, two
 * K-elements per 32-bit column, even k in the low half (TS form, the one in use); 102: odd k
 * low (must fail). */
int ibc_umma_f16_selftest(const float* A, const float* B, int32_t N, int32_t K,
                          uint32_t lbo_a, uint32_t lbo_b, uint32_t sbo,
                          uint32_t layout_type, uint32_t kstep_bytes, float* D,
                          int32_t* status_host, void* stream);

/* The general tcgen05 tf32 GEMM behind the Keras Dense / Conv2D contractions that
 * are not layer-fused (networks/layers/dense_resnet_value.py:31-69,
 * networks/layers/conv_maxpool.py:20-48, the second-order passes of
 * ibc/losses/gradient_loss.py:23-72), exposed for parity tests and benchmarks:
 *   C[M, N] (+)= f((op(A) . op(B)) + bias) (.) (gate_c > 0) + res),  f = id or relu
 *   op(A)(m, k) = trans_a ? A[k * lda + m] : A[m * lda + k], ReLU'd first when relu_a
 *   op(B)(k, n) = trans_b ? B[n * ldb + k] : B[k * ldb + n]
 * bias [N], gate_c / res [M, N] with leading dimension ldc (each may be NULL).
 * Operands are rounded to nearest tf32, accumulation is fp32.  accumulate != 0
 * adds into C with atomics over a K split (weight-gradient form).
 * status_host, when not NULL, makes the call wait for the product and receives 0,
 * or the code of a barrier that timed out. */
int ibc_gemm_tf32(const float* A, int64_t lda, int32_t trans_a, const float* B,
                  int64_t ldb, int32_t trans_b, float* C, int64_t ldc, int32_t M,
                  int32_t N, int32_t K, const float* bias, const float* gate_c,
                  const float* res, int32_t relu_a, int32_t relu_out,
                  int32_t accumulate, int32_t* status_host, void* stream);

/* The implicit-GEMM 3x3 'same' convolution of the ConvMaxpool encoder (Keras Conv2D,
 * networks/layers/conv_maxpool.py:20-48) on its own, for parity tests: NHWC tensors,
 * kernel K [3, 3, C, Co] in the Keras layout, tf32 operands, fp32 accumulation.
 *   mode 0: out [nb, H, W, Co] = f(conv(X, K) + bias),  f = relu when `relu`
 *   mode 1: out [3, 3, C, Co] += dK = im2col(X)^T . dY   (accumulates into out)
 *   mode 2: out [nb, H, W, C]  = dX = conv(dY [nb, H, W, Co], K flipped and transposed)
 * Waits for the result; status_host as in ibc_gemm_tf32. */
int ibc_conv3x3_tf32(const float* X, int64_t nb, int32_t H, int32_t W, int32_t C,
                     const float* K, int32_t Co, const float* bias, int32_t relu,
                     int32_t mode, const float* dY, float* out, int32_t* status_host,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IBC_B200_H_ */
