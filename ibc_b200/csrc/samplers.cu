// Samplers of ibc/agents/mcmc.py: the categorical resample of iterative_dfo
// (bit-exact with the CPU multinomial), its gather/perturb, the Langevin
// update, row softmax and uniform action init.
#include <math.h>

#include <functional>

#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "ops.cuh"
#include "philox.cuh"

namespace ibc {
namespace {

// ---------------------------------------------------------------------------
// Portable fp64 exp: mul/add/rint only, no FMA contraction, identical bits to
// oracle/resample.c and oracle/mcmc.py::exp64.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double exp64_portable(double x) {
  const double kLog2e = 1.4426950408889634;
  const double kLn2Hi = 6.93147180369123816490e-01;
  const double kLn2Lo = 1.90821492927058770002e-10;
  const double c[14] = {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720,
                        1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800,
                        1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
  if (x < -708.0) return 0.0;
  const double k = rint(__dmul_rn(x, kLog2e));
  const double r = __dadd_rn(__dadd_rn(x, -__dmul_rn(k, kLn2Hi)), -__dmul_rn(k, kLn2Lo));
  double p = c[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = __dadd_rn(__dmul_rn(p, r), c[i]);
  const long long bits = ((long long)k + 1023LL) << 52;
  return __dmul_rn(p, __longlong_as_double(bits));
}

// One block per observation row.  Dynamic smem: double cdf[n]; int cnt[n];
// int part[blockDim.x].
__global__ void __launch_bounds__(1024)
resample_kernel(const float* __restrict__ logits, const double* __restrict__ uniforms,
                int n, int count, uint64_t seed, uint64_t rng_stream,
                int32_t* __restrict__ draws, int32_t* __restrict__ counts,
                int64_t* __restrict__ rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* cdf = reinterpret_cast<double*>(smem_raw);
  int* cnt = reinterpret_cast<int*>(cdf + n);
  int* part = cnt + n;
  __shared__ float s_red[32];
  __shared__ float s_max;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* row = logits + (int64_t)b * n;

  // 1. max over finite logits (lowest() when there is none)
  float mx = -3.402823466e+38f;
  for (int j = tid; j < n; j += nt) {
    const float l = row[j];
    if (isfinite(l) && l > mx) mx = l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  if (tid < 32) {
    float v = (tid < (nt + 31) / 32) ? s_red[tid] : -3.402823466e+38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (tid == 0) s_max = v;
  }
  __syncthreads();
  const double dmx = (double)s_max;

  // 2. weights (parallel).  3. CDF.  The specification is the strictly sequential
  // fp64 accumulation of TF's CPU multinomial kernel.  16 384 dependent fp64 adds are
  // ~150 us of a 190 us kernel, so the CDF is first built by a parallel scan (same
  // terms, another association): both sums carry a relative error <= n * 2^-53 of the
  // running total, hence a draw whose to_find is farther than eps = 4 n 2^-53 * total
  // from both neighbouring CDF entries selects the same index under either CDF (the
  // code uses twice that margin).  Draws
  // are searched in the scanned CDF and checked against that margin; only if one of
  // them is ambiguous (probability ~2 n count 2^-51, about 1e-3 per call at n = count
  // = 16 384) does the block fall back to the sequential chain and repeat its draws.
  // The result is bit-identical to the sequential specification in every case.
  __shared__ int s_ambiguous;
  __shared__ double s_warp_tot[32];
  const double eps_rel = 8.0 * (double)n * 1.1102230246251565e-16;   // 2x margin
  for (int pass = 0; pass < 2; ++pass) {
    for (int j = tid; j < n; j += nt) {
      const float l = row[j];
      cdf[j] = isfinite(l) ? exp64_portable(__dadd_rn((double)l, -dmx)) : 0.0;
      cnt[j] = 0;
    }
    if (tid == 0) s_ambiguous = 0;
    __syncthreads();
    if (pass == 0) {
      // blocked inclusive scan: contiguous segment per thread, then across threads
      const int seg = (n + nt - 1) / nt;
      const int j0 = min(n, tid * seg), j1 = min(n, j0 + seg);
      double local = 0.0;
      for (int j = j0; j < j1; ++j) local += cdf[j];
      double incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += up;
      }
      if ((tid & 31) == 31) s_warp_tot[tid >> 5] = incl;
      __syncthreads();
      if (tid < 32) {
        double w = (tid < (nt + 31) / 32) ? s_warp_tot[tid] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, w, o);
          if (tid >= o) w += up;
        }
        s_warp_tot[tid] = w;        // inclusive totals of the warps
      }
      __syncthreads();
      double run = (incl - local) + ((tid >> 5) > 0 ? s_warp_tot[(tid >> 5) - 1] : 0.0);
      for (int j = j0; j < j1; ++j) { run += cdf[j]; cdf[j] = run; }
    } else if (tid == 0) {
      // same left-to-right order as the CPU loop; loads are batched ahead of the
      // dependent add chain so only the fp64 add latency is serial
      double total = 0.0;
      int j = 0;
      for (; j + 16 <= n; j += 16) {
        double w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] = cdf[j + q];
#pragma unroll
        for (int q = 0; q < 16; ++q) { total = __dadd_rn(total, w[q]); w[q] = total; }
#pragma unroll
        for (int q = 0; q < 16; ++q) cdf[j + q] = w[q];
      }
      for (; j < n; ++j) {
        total = __dadd_rn(total, cdf[j]);
        cdf[j] = total;
      }
    }
    __syncthreads();
    const double total = n > 0 ? cdf[n - 1] : 0.0;
    const double eps = eps_rel * total;

    // 4. draws: to_find = u * total; std::upper_bound
    bool ambiguous = false;
    for (int s = tid; s < count; s += nt) {
      double u;
      if (uniforms) {
        u = uniforms[(int64_t)b * count + s];
      } else {
        u = philox_uniform_double(seed, rng_stream, (uint64_t)b * (uint64_t)count + s);
      }
      const double to_find = __dmul_rn(u, total);
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (cdf[mid] > to_find) hi = mid; else lo = mid + 1;
      }
      if (pass == 0) {
        // cdf[lo - 1] <= to_find < cdf[lo]: too close to either neighbour?
        if (lo > 0 && to_find - cdf[lo - 1] <= eps) ambiguous = true;
        if (lo < n && cdf[lo] - to_find <= eps) ambiguous = true;
      }
      if (lo > n - 1) lo = n - 1;
      if (draws) draws[(int64_t)b * count + s] = lo;
      atomicAdd(&cnt[lo], 1);
    }
    if (pass == 0 && ambiguous) s_ambiguous = 1;
    __syncthreads();
    if (pass == 1 || !s_ambiguous) break;
    __syncthreads();
  }
  __syncthreads();
  if (counts)
    for (int j = tid; j < n; j += nt) counts[(int64_t)b * n + j] = cnt[j];
  if (!rows) return;

  // 5. rows = repeat(arange, counts): inclusive scan of cnt, then each output
  // position looks its class up.
  const int chunk = (n + nt - 1) / nt;
  const int j0 = min(n, tid * chunk), j1 = min(n, j0 + chunk);
  int local = 0;
  for (int j = j0; j < j1; ++j) local += cnt[j];
  part[tid] = local;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < nt; ++t) { const int v = part[t]; part[t] = run; run += v; }
  }
  __syncthreads();
  int run = part[tid];
  for (int j = j0; j < j1; ++j) { run += cnt[j]; cnt[j] = run; }   // inclusive
  __syncthreads();
  for (int p = tid; p < count; p += nt) {
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = lo + ((hi - lo) >> 1);
      if (cnt[mid] > p) hi = mid; else lo = mid + 1;
    }
    rows[(int64_t)b * count + p] = (int64_t)b * n + lo;
  }
}

// out[r, :] = clip(src[rows[r], :] + normal * std, min, max)   (std == 0: copy)
__global__ void gather_perturb_kernel(const float* __restrict__ src,
                                      const int64_t* __restrict__ rows, int64_t R, int A,
                                      const float* __restrict__ normals, uint64_t seed,
                                      uint64_t rng_stream, float std,
                                      const float* __restrict__ mn, const float* __restrict__ mx,
                                      float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * A) return;
  const int64_t r = i / A;
  const int a = (int)(i - r * A);
  float v = src[rows[r] * A + a];
  if (std != 0.f) {
    const float z = normals ? normals[i] : philox_normal(seed, rng_stream, (uint64_t)i);
    v = __fadd_rn(v, __fmul_rn(z, std));
    v = fminf(fmaxf(v, mn[a]), mx[a]);
  }
  out[i] = v;
}

__global__ void scale_div_kernel(const float* __restrict__ x, float t, int64_t n,
                                 float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __fdiv_rn(x[i], t);
}

// One block per row: probs = softmax(logits / temperature).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ logits, int n, float temperature,
                    float* __restrict__ probs) {
  __shared__ float s_red[8];
  __shared__ float s_b;
  const int tid = threadIdx.x;
  const float* row = logits + (int64_t)blockIdx.x * n;
  float* out = probs + (int64_t)blockIdx.x * n;
  float mx = -INFINITY;
  for (int j = tid; j < n; j += 256) mx = fmaxf(mx, __fdiv_rn(row[j], temperature));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) { float v = s_red[0]; for (int i = 1; i < 8; ++i) v = fmaxf(v, s_red[i]); s_b = v; }
  __syncthreads();
  mx = s_b;
  float sum = 0.f;
  for (int j = tid; j < n; j += 256) {
    const float e = expf(__fdiv_rn(row[j], temperature) - mx);
    out[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int i = 0; i < 8; ++i) v += s_red[i]; s_b = v; }
  __syncthreads();
  const float inv = s_b;
  for (int j = tid; j < n; j += 256) out[j] = __fdiv_rn(out[j], inv);
}

__global__ void uniform_actions_kernel(const float* __restrict__ u, int64_t rows, int A,
                                       const float* __restrict__ mn, const float* __restrict__ mx,
                                       uint64_t seed, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * A) return;
  const int a = (int)(i % A);
  const float x = u ? u[i] : philox_uniform_float(seed, 0x0u, (uint64_t)i);
  out[i] = __fadd_rn(__fmul_rn(x, __fadd_rn(mx[a], -mn[a])), mn[a]);
}

// One thread per row: mcmc.langevin_step (mcmc.py:209-264) after dE/dact.
__global__ void langevin_update_kernel(float* __restrict__ actions,
                                       const float* __restrict__ dEda, int64_t R, int A,
                                       const float* __restrict__ normals, uint64_t seed,
                                       uint64_t rng_stream, float noise_scale, float grad_clip,
                                       float dclip_scale, float stepsize,
                                       const float* __restrict__ mn, const float* __restrict__ mx,
                                       int norm_type, float* __restrict__ grad_norms,
                                       float* __restrict__ chain_actions) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float nrm = 0.f;
  for (int a = 0; a < A; ++a) {
    const float g = -dEda[r * A + a];          // de_dact = -dE/dact (mcmc.py:190)
    if (norm_type == IBC_NORM_INF) nrm = fmaxf(nrm, fabsf(g));
    else if (norm_type == IBC_NORM_L1) nrm += fabsf(g);
    else if (norm_type == IBC_NORM_L2) nrm = fmaf(g, g, nrm);
  }
  if (norm_type == IBC_NORM_L2) nrm = sqrtf(nrm);
  if (grad_norms) grad_norms[r] = nrm;
  for (int a = 0; a < A; ++a) {
    const int64_t i = r * A + a;
    float g = -dEda[i];
    if (grad_clip > 0.f) g = fminf(fmaxf(g, -grad_clip), grad_clip);
    const float z = normals ? normals[i] : philox_normal(seed, rng_stream, (uint64_t)i);
    g = __fadd_rn(__fmul_rn(0.5f, g), __fmul_rn(z, noise_scale));
    float d = __fmul_rn(stepsize, g);
    const float dc = __fmul_rn(dclip_scale, __fadd_rn(mx[a], -mn[a]));
    d = fminf(fmaxf(d, -dc), dc);
    float v = __fadd_rn(actions[i], -d);
    v = fminf(fmaxf(v, mn[a]), mx[a]);
    actions[i] = v;
    if (chain_actions) chain_actions[i] = v;
  }
}

}  // namespace

int resample(ibc_handle* h, const float* logits, const double* uniforms, int64_t B, int64_t n,
             int64_t count, uint64_t seed, uint64_t rng_stream, int32_t* draws,
             int32_t* counts, int64_t* rows, cudaStream_t s) {
  if (B <= 0 || n <= 0 || count < 0)
    IBC_FAIL(h, IBC_ERR_INVALID, "resample: bad shape B=%lld n=%lld count=%lld", (long long)B,
             (long long)n, (long long)count);
  int threads = 1024;
  while (threads > 64 && threads / 2 >= n) threads >>= 1;
  const size_t smem = (size_t)n * 12 + (size_t)threads * 4;
  if (smem > 220 * 1024)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "resample: n=%lld exceeds the shared-memory CDF (max %d)",
             (long long)n, (int)((220 * 1024 - 4096) / 12));
  static bool attr_set = false;
  if (!attr_set) {
    IBC_CUDA(h, cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     220 * 1024));
    attr_set = true;
  }
  resample_kernel<<<(unsigned)B, threads, smem, s>>>(logits, uniforms, (int)n, (int)count, seed,
                                                     rng_stream, draws, counts, rows);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int softmax_rows(ibc_handle* h, const float* logits, int64_t B, int64_t n, float temperature,
                 float* probs, cudaStream_t s) {
  if (B <= 0 || n <= 0) IBC_FAIL(h, IBC_ERR_INVALID, "softmax_rows: bad shape");
  softmax_rows_kernel<<<(unsigned)B, 256, 0, s>>>(logits, (int)n, temperature, probs);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int uniform_actions(ibc_handle* h, const float* u, int64_t rows, int A, const float* mn,
                    const float* mx, uint64_t seed, float* out, cudaStream_t s) {
  const int64_t tot = rows * A;
  if (tot <= 0) return IBC_OK;
  uniform_actions_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(u, rows, A, mn, mx, seed,
                                                                      out);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int gather_perturb(ibc_handle* h, const float* src, const int64_t* rows, int64_t R, int A,
                   const float* normals, uint64_t seed, uint64_t rng_stream, float std,
                   const float* mn, const float* mx, float* out, cudaStream_t s) {
  if (R <= 0 || A <= 0) return IBC_OK;
  gather_perturb_kernel<<<(unsigned)((R * A + 255) / 256), 256, 0, s>>>(
      src, rows, R, A, normals, seed, rng_stream, std, mn, mx, out);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int langevin_update(ibc_handle* h, float* actions, const float* dEda, int64_t R, int A,
                    const float* normals, uint64_t seed, uint64_t rng_stream, float noise_scale,
                    float grad_clip, float delta_action_clip, float stepsize, const float* mn,
                    const float* mx, int norm_type, float* grad_norms, cudaStream_t s) {
  if (R <= 0 || A <= 0) return IBC_OK;
  const float dclip_scale = (float)((double)delta_action_clip * 0.5);
  langevin_update_kernel<<<(unsigned)((R + 127) / 128), 128, 0, s>>>(
      actions, dEda, R, A, normals, seed, rng_stream, noise_scale, grad_clip, dclip_scale,
      stepsize, mn, mx, norm_type, grad_norms, nullptr);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

// mcmc.iterative_dfo (ibc/agents/mcmc.py:58-158) around any energy evaluator
// energy_fn(actions [B*n, A]) -> E [B*n] (MLPEBM, or the PixelEBM value head on
// a fixed observation encoding for late fusion, mcmc.py:98-110).
int dfo_generic(ibc_handle* h, int A, int64_t B, float* actions, int64_t n, const float* mn,
                const float* mx, float temperature, int num_iterations, float iteration_std,
                const double* uniforms, const float* normals, uint64_t seed, float* probs,
                const std::function<int(const float*, float*)>& energy_fn, cudaStream_t s) {
  const int64_t R = B * n;
  if (num_iterations < 1) IBC_FAIL(h, IBC_ERR_INVALID, "dfo: num_iterations must be >= 1");
  if (temperature <= 0.f) IBC_FAIL(h, IBC_ERR_INVALID, "dfo: temperature must be > 0");
  const size_t bytes = 2 * Workspace::rounded((size_t)R * 4) + Workspace::rounded((size_t)R * 8) +
                       Workspace::rounded((size_t)R * A * 4);
  IBC_CUDA(h, h->ws_outer.reserve(bytes));
  h->ws_outer.reset();
  float* E = h->ws_outer.take<float>(R);
  float* logits_buf = h->ws_outer.take<float>(R);
  int64_t* rows = h->ws_outer.take<int64_t>(R);
  float* tmp = h->ws_outer.take<float>((size_t)R * A);
  float std = iteration_std;
  const float* logits = E;
  for (int it = 0; it < num_iterations; ++it) {
    IBC_TRY(energy_fn(actions, E));
    if (temperature != 1.0f) {
      scale_div_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(E, temperature, R, logits_buf);
      IBC_LAUNCH_CHECK(h);
      logits = logits_buf;
    }
    IBC_TRY(resample(h, logits, uniforms ? uniforms + (int64_t)it * R : nullptr, B, n, n, seed,
                     2ull * it + 1, nullptr, nullptr, rows, s));
    const bool last = (it == num_iterations - 1);
    const float this_std = last ? 0.f : std;
    gather_perturb_kernel<<<(unsigned)((R * A + 255) / 256), 256, 0, s>>>(
        actions, rows, R, A, (normals && !last) ? normals + (int64_t)it * R * A : nullptr, seed,
        2ull * it + 2, this_std, mn, mx, tmp);
    IBC_LAUNCH_CHECK(h);
    IBC_CUDA(h, cudaMemcpyAsync(actions, tmp, (size_t)R * A * 4, cudaMemcpyDeviceToDevice, s));
    if (!last) std *= 0.5f;
  }
  if (probs) IBC_TRY(softmax_rows(h, logits, B, n, 1.0f, probs, s));
  return IBC_OK;
}

int dfo(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n, const float* mn,
        const float* mx, float temperature, int num_iterations, float iteration_std,
        const double* uniforms, const float* normals, uint64_t seed, float* probs,
        cudaStream_t s) {
  auto energy_fn = [&](const float* acts, float* E) -> int {
    return mlp_energy(h, obs, B, acts, n, E, s);
  };
  return dfo_generic(h, h->L.A, B, actions, n, mn, mx, temperature, num_iterations, iteration_std,
                     uniforms, normals, seed, probs, energy_fn, s);
}

void langevin_stepsizes(const ibc_langevin_cfg& c, std::vector<float>* out) {
  const int K = c.num_iterations;
  out->resize(K > 0 ? K : 0);
  if (K <= 0) return;
  (*out)[0] = c.sampler_stepsize_init;
  double latest = (double)c.sampler_stepsize_init;
  for (int i = 1; i < K; ++i) {
    double r;
    if (c.use_polynomial_rate) {
      // PolynomialSchedule.get_rate(i), mcmc.py:290-294
      const double init = (double)c.sampler_stepsize_init, fin = (double)c.sampler_stepsize_final;
      r = (init - fin) * pow(1.0 - (double)i / (double)(K - 1), (double)c.sampler_stepsize_power) +
          fin;
    } else {
      latest *= (double)c.sampler_stepsize_decay;   // ExponentialSchedule, mcmc.py:275-278
      r = latest;
    }
    (*out)[i] = (float)r;
  }
}

// mcmc.langevin_actions_given_obs (ibc/agents/mcmc.py:330-425).
int langevin(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
             const float* mn, const float* mx, const ibc_langevin_cfg& c, const float* normals,
             uint64_t seed, float* chain_actions, float* chain_energies, float* chain_grad_norms,
             cudaStream_t s) {
  const int A = h->L.A;
  const int64_t R = B * n;
  if (c.num_iterations < 0) IBC_FAIL(h, IBC_ERR_INVALID, "langevin: num_iterations < 0");
  if (c.grad_norm_type < IBC_NORM_NONE || c.grad_norm_type > IBC_NORM_INF)
    IBC_FAIL(h, IBC_ERR_INVALID, "langevin: unknown grad_norm_type %d", c.grad_norm_type);
  std::vector<float> steps;
  langevin_stepsizes(c, &steps);
  if (h->precision == IBC_PRECISION_TF32 && umma_supported(h->L))
    return langevin_umma(h, obs, B, actions, n, mn, mx, c, normals, seed, steps, chain_actions,
                         chain_energies, chain_grad_norms, s);
  const size_t bytes = 2 * Workspace::rounded((size_t)R * 4) + Workspace::rounded((size_t)R * A * 4);
  IBC_CUDA(h, h->ws_outer.reserve(bytes));
  h->ws_outer.reset();
  float* E = h->ws_outer.take<float>(R);
  float* nrm = h->ws_outer.take<float>(R);
  float* dEda = h->ws_outer.take<float>((size_t)R * A);
  // delta_action_clip * 0.5 evaluated in double like the Python scalar product
  const float dclip_scale = (float)((double)c.delta_action_clip * 0.5);
  for (int k = 0; k < c.num_iterations; ++k) {
    float* Ek = chain_energies ? chain_energies + (int64_t)k * R : E;
    float* Nk = chain_grad_norms ? chain_grad_norms + (int64_t)k * R : nrm;
    IBC_TRY(mlp_energy_dact(h, obs, B, actions, n, Ek, dEda, c.apply_exp != 0, s));
    langevin_update_kernel<<<(unsigned)((R + 127) / 128), 128, 0, s>>>(
        actions, dEda, R, A, normals ? normals + (int64_t)k * R * A : nullptr, seed,
        (uint64_t)k + 1, c.noise_scale, c.grad_clip, dclip_scale, steps[k], mn, mx,
        c.grad_norm_type, Nk, chain_actions ? chain_actions + (int64_t)k * R * A : nullptr);
    IBC_LAUNCH_CHECK(h);
  }
  return IBC_OK;
}

}  // namespace ibc
