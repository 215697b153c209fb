// Samplers of ibc/agents/mcmc.py: the categorical resample of iterative_dfo
// (bit-exact with the CPU multinomial), its gather/perturb, the Langevin
// update, row softmax and uniform action init.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <functional>

#include "gemm_simt.cuh"
#include "mlp.cuh"
#include "chain_graph.cuh"
#include "ops.cuh"
#include "philox.cuh"

#include <cooperative_groups.h>

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

constexpr int kResampleBuckets = 2048;    // value-space index over the CDF (power of two)
constexpr int kResampleHeavy = 512;       // classes expanded by the whole block

// The fp64 scan of the CDF, with a contiguous chunk per WARP: the warps run their chunks independently
// (a 32-element shuffle scan per step with a running carry) and meet once for the chunk
// offsets, instead of three block barriers and a serial warp-0 pass for every blockDim.x
// elements.  Another association of the same terms: the bound n 2^-53 on the relative error
// of a running total holds for every order of summation.
__device__ __forceinline__ void block_scan_chunked(double* a, int n, double* s_w) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nw = (nt + 31) >> 5;
  const int chunk = (((n + nw - 1) / nw) + 31) & ~31;
  const int lo = min(n, warp * chunk), hi = min(n, lo + chunk);
  double carry = 0.0;
  for (int base = lo; base < hi; base += 32) {
    const int j = base + lane;
    double v = (j < hi) ? a[j] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += up;
    }
    v += carry;
    if (j < hi) a[j] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) s_w[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    double w = (lane < nw) ? s_w[lane] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += up;
    }
    s_w[lane] = w;
  }
  __syncthreads();
  if (warp > 0) {
    const double off = s_w[warp - 1];
    for (int j = lo + lane; j < hi; j += 32) a[j] += off;
  }
  __syncthreads();
}

// Block-wide inclusive scan over shared memory, element j = base + tid (no bank conflicts; a
// contiguous segment per thread is a 32-way conflict on every access).
__device__ __forceinline__ void block_scan_inplace(int* a, int n, int* s_w) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nw = (nt + 31) >> 5;
  int carry = 0;
  for (int base = 0; base < n; base += nt) {
    const int j = base + tid;
    int v = (j < n) ? a[j] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += up;
    }
    if (lane == 31) s_w[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = (lane < nw) ? s_w[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += up;
      }
      s_w[lane] = w;
    }
    __syncthreads();
    if (j < n) a[j] = v + carry + (warp > 0 ? s_w[warp - 1] : 0);
    carry += s_w[nw - 1];
    __syncthreads();
  }
}

// std::upper_bound over a[lo, hi): first index with a[j] > x, hi if none
__device__ __forceinline__ int upper_bound_range(const double* a, int lo, int hi, double x) {
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (a[mid] > x) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// max over the finite logits of a row (lowest() when there is none), block-wide
__device__ __forceinline__ float block_row_max(const float* __restrict__ row, int n, float* s_red,
                                               float* s_max) {
  const int tid = threadIdx.x, nt = blockDim.x;
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
    if (tid == 0) *s_max = v;
  }
  __syncthreads();
  return *s_max;
}

// The strictly sequential fp64 accumulation of TF's CPU multinomial kernel, in place (one
// thread): loads are batched ahead of the dependent add chain so only the fp64 add latency is
// serial.
__device__ __forceinline__ void sequential_cdf(double* cdf, int n) {
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

// One draw: to_find = u * total; std::upper_bound over the CDF, confined to the entries of its
// value-space bucket when the index is in use.  `check`: flag a to_find closer than eps to either
// neighbouring CDF entry.
__device__ __forceinline__ int draw_index(const double* cdf, const int* tab, bool use_tab, int n,
                                          double u, double total, double eps, bool check,
                                          bool* ambiguous) {
  constexpr double kInvBuckets = 1.0 / kResampleBuckets;
  const double to_find = __dmul_rn(u, total);
  int lo = 0, hi = n;
  if (use_tab) {
    int k = (int)(u * (double)kResampleBuckets);
    k = max(0, min(kResampleBuckets - 1, k));
    while (k > 0 && to_find < __dmul_rn((double)k * kInvBuckets, total)) --k;
    while (k < kResampleBuckets - 1 &&
           to_find >= __dmul_rn((double)(k + 1) * kInvBuckets, total)) ++k;
    // u >= 1 (not produced by the samplers) or a NaN falls outside every bucket
    if (to_find >= __dmul_rn((double)k * kInvBuckets, total) &&
        to_find < __dmul_rn((double)(k + 1) * kInvBuckets, total)) {
      lo = tab[k];
      hi = tab[k + 1];
    }
  }
  lo = upper_bound_range(cdf, lo, hi, to_find);
  if (check) {
    // cdf[lo - 1] <= to_find < cdf[lo]: too close to either neighbour?
    if (lo > 0 && to_find - cdf[lo - 1] <= eps) *ambiguous = true;
    if (lo < n && cdf[lo] - to_find <= eps) *ambiguous = true;
  }
  return lo > n - 1 ? n - 1 : lo;
}

// rows = repeat(arange, counts) for the classes [j0, j0 + m) whose inclusive count scan is
// cnt[0, m) (+ base): every class writes its own run; classes with long runs are left to the
// whole block.
__device__ __forceinline__ void write_runs(const int* cnt, int m, int j0, int base, int64_t row0,
                                           int64_t* __restrict__ out, int* heavy, int* s_nheavy) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = tid; j < m; j += nt) {
    const int end = cnt[j], start = (j > 0) ? cnt[j - 1] : 0;
    const int c = end - start;
    if (c >= 64) {
      const int slot = atomicAdd(s_nheavy, 1);
      if (slot < kResampleHeavy) { heavy[slot] = j; continue; }
    }
    for (int q = start; q < end; ++q) out[base + q] = row0 + j0 + j;
  }
  __syncthreads();
  const int nheavy = min(*s_nheavy, kResampleHeavy);
  for (int hslot = 0; hslot < nheavy; ++hslot) {
    const int j = heavy[hslot];
    const int end = cnt[j], start = (j > 0) ? cnt[j - 1] : 0;
    for (int q = start + tid; q < end; q += nt) out[base + q] = row0 + j0 + j;
  }
}

// 2. weights (parallel).  3. CDF.  The specification is the strictly sequential
// fp64 accumulation of TF's CPU multinomial kernel.  16 384 dependent fp64 adds are
// ~80 us on one thread, so the CDF is first built by a parallel scan (same terms,
// another association): both sums carry a relative error <= n * 2^-53 of the running
// total, hence a draw whose to_find is farther than 4 n 2^-53 * total from both
// neighbouring CDF entries selects the same index under either CDF (the code uses
// twice that margin).  Draws are searched in the scanned CDF and checked against the
// margin; only if one of them is ambiguous (probability ~4 n count 2^-51, about
// 2e-3 per call at n = count = 16 384) does the block fall back to the sequential
// chain and repeat its draws.  The result is bit-identical to the sequential
// specification in every case.
// 4. Draws: to_find = u * total; std::upper_bound.  A value-space index (tab[k] =
// upper_bound(cdf, k/K * total)) confines each search to the entries of its bucket
// (v_k <= to_find < v_k+1 implies tab[k] <= answer <= tab[k+1] by monotonicity), so a
// draw costs a couple of shared-memory probes instead of log2(n) conflicting ones.
//
// One block per observation row.  Dynamic smem: double cdf[n]; int cnt[n];
// int tab[kResampleBuckets + 1]; int heavy[kResampleHeavy].
__global__ void __launch_bounds__(1024)
resample_kernel(const float* __restrict__ logits, const double* __restrict__ uniforms,
                int n, int count, uint64_t seed, uint64_t rng_stream,
                int32_t* __restrict__ draws, int32_t* __restrict__ counts,
                int64_t* __restrict__ rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* cdf = reinterpret_cast<double*>(smem_raw);
  int* cnt = reinterpret_cast<int*>(cdf + n);
  int* tab = cnt + n;
  int* heavy = tab + kResampleBuckets + 1;
  __shared__ float s_red[32];
  __shared__ float s_max;
  __shared__ double s_wd[32];
  __shared__ int s_wi[32];
  __shared__ int s_ambiguous;
  __shared__ int s_nheavy;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* row = logits + (int64_t)b * n;

  // 1. max over finite logits (lowest() when there is none)
  const double dmx = (double)block_row_max(row, n, s_red, &s_max);

  const double eps_rel = 8.0 * (double)n * 1.1102230246251565e-16;
  const bool use_tab = n >= 4 * kResampleBuckets;
  constexpr double kInvBuckets = 1.0 / kResampleBuckets;
  for (int pass = 0; pass < 2; ++pass) {
    for (int j = tid; j < n; j += nt) {
      const float l = row[j];
      cdf[j] = isfinite(l) ? exp64_portable(__dadd_rn((double)l, -dmx)) : 0.0;
      cnt[j] = 0;
    }
    if (tid == 0) s_ambiguous = 0;
    __syncthreads();
    if (pass == 0) {
      block_scan_chunked(cdf, n, s_wd);
    } else if (tid == 0) {
      sequential_cdf(cdf, n);     // same left-to-right order as the CPU loop
    }
    __syncthreads();
    const double total = n > 0 ? cdf[n - 1] : 0.0;
    const double eps = eps_rel * total;
    if (use_tab) {
      for (int k = tid; k <= kResampleBuckets; k += nt)
        tab[k] = upper_bound_range(cdf, 0, n, __dmul_rn((double)k * kInvBuckets, total));
      __syncthreads();
    }

    bool ambiguous = false;
    for (int s = tid; s < count; s += nt) {
      double u;
      if (uniforms) {
        u = uniforms[(int64_t)b * count + s];
      } else {
        u = philox_uniform_double(seed, rng_stream, (uint64_t)b * (uint64_t)count + s);
      }
      const int lo = draw_index(cdf, tab, use_tab, n, u, total, eps, pass == 0, &ambiguous);
      if (draws) draws[(int64_t)b * count + s] = lo;
      atomicAdd(&cnt[lo], 1);
    }
    if (pass == 0 && ambiguous) s_ambiguous = 1;
    __syncthreads();
    if (pass == 1 || !s_ambiguous) break;
    __syncthreads();
  }
  if (counts)
    for (int j = tid; j < n; j += nt) counts[(int64_t)b * n + j] = cnt[j];
  if (!rows) return;

  // 5. rows = repeat(arange, counts): inclusive scan of cnt, then the runs
  if (tid == 0) s_nheavy = 0;
  block_scan_inplace(cnt, n, s_wi);        // ends with __syncthreads()
  write_runs(cnt, n, 0, 0, (int64_t)b * n, rows + (int64_t)b * count, heavy, &s_nheavy);
}

// The same resample on a thread-block cluster per observation row.  One SM is instruction
// bound on a row of 16 384 classes and draws (237 k warp instructions, 62 us, ncu): with one
// or a few rows -- policy inference -- the other 147 SMs idle.  Every CTA of the cluster
// builds the whole CDF (cheap, and each needs all of it to search), takes every C-th block of
// draws, and bins a draw into the CTA that OWNS the class (slice c of ceil(n / C) classes)
// with an atomic on distributed shared memory; after a cluster barrier each CTA scans its
// slice of the counts, learns its output offset from the lower slices' totals and writes its
// part of `rows`.  Same draws, same counts, same rows as resample_kernel: integer arithmetic
// after the (identical) CDF.  Dynamic smem: double cdf[n]; int cnt[slice]; tab; heavy.
__device__ long long g_resample_dbg[16];    // IBC_DBG_TIMELINE: phase stamps of cluster 0, CTA 0

__global__ void __launch_bounds__(1024)
resample_cluster_kernel(const float* __restrict__ logits, const double* __restrict__ uniforms,
                        int n, int count, uint64_t seed, uint64_t rng_stream,
                        int32_t* __restrict__ draws, int32_t* __restrict__ counts,
                        int64_t* __restrict__ rows) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), c = (int)cluster.block_rank();
  const int slice = (n + C - 1) / C;
  const int j0 = c * slice, m = max(0, min(n, j0 + slice) - j0);      // classes this CTA owns
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* cdf = reinterpret_cast<double*>(smem_raw);
  int* cnt = reinterpret_cast<int*>(cdf + n);       // [slice]
  int* tab = cnt + slice;
  int* heavy = tab + kResampleBuckets + 1;
  __shared__ float s_red[32];
  __shared__ float s_max;
  __shared__ double s_wd[32];
  __shared__ int s_wi[32];
  __shared__ int s_ambiguous;
  __shared__ int s_nheavy;
  __shared__ int s_total;
  __shared__ double s_slice_total;
  const int b = blockIdx.x / C;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* row = logits + (int64_t)b * n;

  const bool stamp = blockIdx.x == 0 && tid == 0;
  if (stamp) g_resample_dbg[0] = clock64();
  const double dmx = (double)block_row_max(row, n, s_red, &s_max);
  const double eps_rel = 8.0 * (double)n * 1.1102230246251565e-16;
  const bool use_tab = n >= 4 * kResampleBuckets;
  constexpr double kInvBuckets = 1.0 / kResampleBuckets;
  if (stamp) g_resample_dbg[1] = clock64();
  if (stamp) g_resample_dbg[2] = clock64();
  for (int pass = 0; pass < 2; ++pass) {
    for (int j = tid; j < slice; j += nt) cnt[j] = 0;
    if (tid == 0) s_ambiguous = 0;
    if (pass == 0) {
      // The fp64 pipe is the bottleneck of the CDF (exponentials, scan and index all run on
      // it: 13 k cycles for a 16 384-entry scan on one SM), so each CTA exponentiates and scans
      // only the classes it owns, learns its offset from the lower slices' totals (summed in
      // slice order, so the CDF stays monotone across slices) and stores its part of the CDF
      // into every CTA's array.  One more association of the same positive terms: the margin
      // below covers it.
      for (int j = j0 + tid; j < j0 + m; j += nt) {
        const float l = row[j];
        cdf[j] = isfinite(l) ? exp64_portable(__dadd_rn((double)l, -dmx)) : 0.0;
      }
      __syncthreads();
      if (stamp) g_resample_dbg[3] = clock64();
      block_scan_chunked(cdf + j0, m, s_wd);
      if (tid == 0) s_slice_total = m > 0 ? cdf[j0 + m - 1] : 0.0;
      cluster.sync();       // slice totals; every CTA is running and has zeroed its counts
      if (stamp) g_resample_dbg[4] = clock64();
      double offset = 0.0;
      for (int r = 0; r < c; ++r) offset = __dadd_rn(offset, *cluster.map_shared_rank(&s_slice_total, r));
      // (pulling the other slices with 16-byte loads instead was slower: 10.5 k against 8.4 k cycles)
      for (int j = j0 + tid; j < j0 + m; j += nt) {
        const double v = __dadd_rn(offset, cdf[j]);
        for (int r = 0; r < C; ++r) cluster.map_shared_rank(cdf, r)[j] = v;
      }
      cluster.sync();       // the whole CDF has landed everywhere
    } else {
      // sequential fallback: weights everywhere, then the CPU kernel's left-to-right chain
      for (int j = j0 + tid; j < j0 + m; j += nt) {
        const float l = row[j];
        const double w = isfinite(l) ? exp64_portable(__dadd_rn((double)l, -dmx)) : 0.0;
        for (int r = 0; r < C; ++r) cluster.map_shared_rank(cdf, r)[j] = w;
      }
      cluster.sync();
      if (tid == 0) sequential_cdf(cdf, n);
      __syncthreads();
    }
    if (stamp) g_resample_dbg[5] = clock64();
    const double total = n > 0 ? cdf[n - 1] : 0.0;
    const double eps = eps_rel * total;
    if (use_tab) {
      // the value-space index, one share of the buckets per CTA, stored into every CTA's table
      const int per = (kResampleBuckets + 1 + C - 1) / C;
      for (int k = c * per + tid; k < min(kResampleBuckets + 1, (c + 1) * per); k += nt) {
        const int t = upper_bound_range(cdf, 0, n, __dmul_rn((double)k * kInvBuckets, total));
        for (int r = 0; r < C; ++r) cluster.map_shared_rank(tab, r)[k] = t;
      }
      cluster.sync();
    }
    if (stamp) g_resample_dbg[6] = clock64();

    bool ambiguous = false;
    for (int s = c * nt + tid; s < count; s += C * nt) {
      double u;
      if (uniforms) {
        u = uniforms[(int64_t)b * count + s];
      } else {
        u = philox_uniform_double(seed, rng_stream, (uint64_t)b * (uint64_t)count + s);
      }
      const int lo = draw_index(cdf, tab, use_tab, n, u, total, eps, pass == 0, &ambiguous);
      if (draws) draws[(int64_t)b * count + s] = lo;
      const int owner = lo / slice;
      atomicAdd(cluster.map_shared_rank(cnt, owner) + (lo - owner * slice), 1);
    }
    if (pass == 0 && ambiguous) s_ambiguous = 1;
    if (stamp) g_resample_dbg[7] = clock64();
    cluster.sync();
    if (stamp) g_resample_dbg[8] = clock64();
    int any = 0;          // one ambiguous draw anywhere sends the whole cluster to the sequential CDF
    for (int r = 0; r < C; ++r) any |= *cluster.map_shared_rank(&s_ambiguous, r);
    if (pass == 1 || !any) break;
    cluster.sync();       // every CTA has read the flags before they are reset
  }
  if (counts)
    for (int j = tid; j < m; j += nt) counts[(int64_t)b * n + j0 + j] = cnt[j];
  if (rows) {
    if (tid == 0) s_nheavy = 0;
    block_scan_inplace(cnt, m, s_wi);        // ends with __syncthreads()
    if (tid == 0) s_total = m > 0 ? cnt[m - 1] : 0;
    if (stamp) g_resample_dbg[9] = clock64();
    cluster.sync();
    if (stamp) g_resample_dbg[10] = clock64();
    int base = 0;
    for (int r = 0; r < c; ++r) base += *cluster.map_shared_rank(&s_total, r);
    write_runs(cnt, m, j0, base, (int64_t)b * n, rows + (int64_t)b * count, heavy, &s_nheavy);
  }
  if (stamp) g_resample_dbg[11] = clock64();
  cluster.sync();         // no CTA leaves while a peer may still read its shared memory
  if (stamp) g_resample_dbg[12] = clock64();
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
template <int NT>
__global__ void __launch_bounds__(NT)
softmax_rows_kernel(const float* __restrict__ logits, int n, float temperature,
                    float* __restrict__ probs) {
  __shared__ float s_red[32];
  __shared__ float s_b;
  const int tid = threadIdx.x;
  const float* row = logits + (int64_t)blockIdx.x * n;
  float* out = probs + (int64_t)blockIdx.x * n;
  float mx = -INFINITY;
  for (int j = tid; j < n; j += NT) mx = fmaxf(mx, __fdiv_rn(row[j], temperature));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) { float v = s_red[0]; for (int i = 1; i < NT / 32; ++i) v = fmaxf(v, s_red[i]); s_b = v; }
  __syncthreads();
  mx = s_b;
  float sum = 0.f;
  for (int j = tid; j < n; j += NT) {
    const float e = expf(__fdiv_rn(row[j], temperature) - mx);
    out[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int i = 0; i < NT / 32; ++i) v += s_red[i]; s_b = v; }
  __syncthreads();
  const float inv = s_b;
  for (int j = tid; j < n; j += NT) out[j] = __fdiv_rn(out[j], inv);
}

// The same softmax on a thread-block cluster per row (policy inference: one row of 16 384
// logits kept a single SM busy for 14 us, most of it IEEE divisions and expf in sequence): every
// CTA takes a slice; the row maximum and the normaliser are exchanged over distributed shared
// memory (slice sums added in slice order).
__global__ void __launch_bounds__(1024)
softmax_rows_cluster_kernel(const float* __restrict__ logits, int n, float temperature,
                            float* __restrict__ probs) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), c = (int)cluster.block_rank();
  const int slice = (n + C - 1) / C;
  const int j0 = c * slice, j1 = min(n, j0 + slice);
  __shared__ float s_red[32];
  __shared__ float s_max, s_sum;
  const int tid = threadIdx.x, nt = blockDim.x, nw = (nt + 31) >> 5;
  const float* row = logits + (int64_t)(blockIdx.x / C) * n;
  float* out = probs + (int64_t)(blockIdx.x / C) * n;
  float mx = -INFINITY;
  for (int j = j0 + tid; j < j1; j += nt) mx = fmaxf(mx, __fdiv_rn(row[j], temperature));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) { float v = s_red[0]; for (int i = 1; i < nw; ++i) v = fmaxf(v, s_red[i]); s_max = v; }
  cluster.sync();
  mx = -INFINITY;
  for (int r = 0; r < C; ++r) mx = fmaxf(mx, *cluster.map_shared_rank(&s_max, r));
  float sum = 0.f;
  for (int j = j0 + tid; j < j1; j += nt) {
    const float e = expf(__fdiv_rn(row[j], temperature) - mx);
    out[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int i = 0; i < nw; ++i) v += s_red[i]; s_sum = v; }
  cluster.sync();
  float total = 0.f;
  for (int r = 0; r < C; ++r) total += *cluster.map_shared_rank(&s_sum, r);
  for (int j = j0 + tid; j < j1; j += nt) out[j] = __fdiv_rn(out[j], total);
  cluster.sync();         // no CTA leaves while a peer may still read its shared memory
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

// The counter-example block of a training step in one pass: combined[b, j, :] = the same uniform
// sample uniform_actions_kernel draws for row b * N + j (j < N), combined[b, N, :] = the true
// action (ibc_agent.py:461-473: the expert action LAST).
__global__ void uniform_combined_kernel(const float* __restrict__ u, int64_t B, int64_t N, int A,
                                        const float* __restrict__ mn, const float* __restrict__ mx,
                                        uint64_t seed, const float* __restrict__ actions,
                                        float* __restrict__ out) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // index into [B, N + 1, A]
  if (o >= B * (N + 1) * A) return;
  const int a = (int)(o % A);
  const int64_t bj = o / A, b = bj / (N + 1), j = bj - b * (N + 1);
  if (j == N) { out[o] = actions[b * A + a]; return; }
  const int64_t i = (b * N + j) * A + a;                                  // index into [B * N, A]
  const float x = u ? u[i] : philox_uniform_float(seed, 0x0u, (uint64_t)i);
  out[o] = __fadd_rn(__fmul_rn(x, __fadd_rn(mx[a], -mn[a])), mn[a]);
}

// GreedyPolicy over a MappedCategorical (ibc_policy.py:96-117, greedy_policy.py): per observation
// the FIRST index of the largest probability and the action mapped to it.  One block per
// observation; replaces an arg-max, an arange, a gather and (in the caller) two clips.
__global__ void __launch_bounds__(256)
greedy_actions_kernel(const float* __restrict__ probs, const float* __restrict__ values, int n, int A,
                      const float* __restrict__ mn, const float* __restrict__ mx,
                      float* __restrict__ out) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const int tid = threadIdx.x;
  const float* row = probs + (int64_t)blockIdx.x * n;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = tid; j < n; j += 256) {
    const float p = row[j];
    if (p > best || (p == best && j < bi) || (p != p && !(best != best))) { best = p; bi = j; }   // NaN wins, like arg max
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const bool take = (ob != ob && !(best != best)) || (!(best != best) && ob > best) ||
                      ((ob == best || (ob != ob && best != best)) && oi < bi);
    if (take) { best = ob; bi = oi; }
  }
  if ((tid & 31) == 0) { s_v[tid >> 5] = best; s_i[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      const float ob = s_v[w];
      const int oi = s_i[w];
      const bool take = (ob != ob && !(best != best)) || (!(best != best) && ob > best) ||
                        ((ob == best || (ob != ob && best != best)) && oi < bi);
      if (take) { best = ob; bi = oi; }
    }
    s_i[0] = bi;
  }
  __syncthreads();
  bi = min(s_i[0], n - 1);
  for (int a = tid; a < A; a += 256) {
    float v = values[((int64_t)blockIdx.x * n + bi) * A + a];
    if (mn) v = fminf(fmaxf(v, mn[a]), mx[a]);
    out[(int64_t)blockIdx.x * A + a] = v;
  }
}

// One thread per row: mcmc.langevin_step (mcmc.py:209-264) after dE/dact.
__global__ void langevin_update_kernel(float* __restrict__ actions,
                                       const float* __restrict__ dEda, int64_t R, int A,
                                       const float* __restrict__ normals, uint64_t seed,
                                       uint64_t rng_stream, float noise_scale, float grad_clip,
                                       float dclip_scale, float stepsize,
                                       const float* __restrict__ mn, const float* __restrict__ mx,
                                       int norm_type, float* __restrict__ grad_norms,
                                       float* __restrict__ chain_actions,
                                       const uint64_t* __restrict__ seed_ptr) {
  if (seed_ptr) seed = *seed_ptr;     // captured chains read the call's seed from memory
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
  // Few rows of many classes (policy inference): a cluster of CTAs per row, as many as keep
  // the launch within one wave.  IBC_RESAMPLE_CLUSTER=1|2|4|8 overrides.
  int C = 1;
  if (n >= 4096 && count >= 4096) {
    int sms = h ? h->sm_count : 0;
    if (!h) {      // the stateless entry point (ibc_resample without a handle)
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    C = 8;
    while (C > 1 && B * C > sms) C >>= 1;
  }
  if (const char* e = getenv("IBC_RESAMPLE_CLUSTER")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) C = v;
  }
  if (C > 1) {
    const int64_t slice = (n + C - 1) / C;
    const size_t smem = (size_t)n * 8 + (size_t)slice * 4 +
                        (size_t)(kResampleBuckets + 1 + kResampleHeavy) * 4 + 16;
    if (smem <= 220 * 1024) {
      IBC_RAISE_SMEM(h, resample_cluster_kernel, 220 * 1024);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(B * C));
      cfg.blockDim = dim3((unsigned)threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)C;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      IBC_CUDA(h, cudaLaunchKernelEx(&cfg, resample_cluster_kernel, logits, uniforms, (int)n,
                                     (int)count, seed, rng_stream, draws, counts, rows));
      count_launch();
      static const bool want_dbg = getenv("IBC_DBG_TIMELINE") != nullptr;
      if (want_dbg) {
        long long t[16];
        cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(t, g_resample_dbg, sizeof(t));
        fprintf(stderr, "[ibc timeline resample C=%d] max %lld | - %lld | weights %lld | scan+sync %lld | offsets+sync %lld | tab+sync %lld | "
                "draws %lld | sync %lld | counts+scan %lld | sync %lld | rows %lld | sync %lld\n", C,
                t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6],
                t[8] - t[7], t[9] - t[8], t[10] - t[9], t[11] - t[10], t[12] - t[11]);
      }
      return IBC_OK;
    }
  }
  const size_t smem = (size_t)n * 12 + (size_t)(kResampleBuckets + 1 + kResampleHeavy) * 4 + 16;
  if (smem > 220 * 1024)
    IBC_FAIL(h, IBC_ERR_UNSUPPORTED, "resample: n=%lld exceeds the shared-memory CDF (max %d)",
             (long long)n, (int)((220 * 1024 - 16384) / 12));
  IBC_RAISE_SMEM(h, resample_kernel, 220 * 1024);
  resample_kernel<<<(unsigned)B, threads, smem, s>>>(logits, uniforms, (int)n, (int)count, seed,
                                                     rng_stream, draws, counts, rows);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int softmax_rows(ibc_handle* h, const float* logits, int64_t B, int64_t n, float temperature,
                 float* probs, cudaStream_t s) {
  if (B <= 0 || n <= 0) IBC_FAIL(h, IBC_ERR_INVALID, "softmax_rows: bad shape");
  // few long rows (policy inference: one row of 16 384): a cluster of eight CTAs per row
  int sms = h ? h->sm_count : 0;
  if (!h) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (n >= 8192 && B * 8 <= sms) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * 8));
    cfg.blockDim = dim3(1024);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IBC_CUDA(h, cudaLaunchKernelEx(&cfg, softmax_rows_cluster_kernel, logits, (int)n, temperature, probs));
    count_launch();
    return IBC_OK;
  }
  if (n >= 4096) softmax_rows_kernel<1024><<<(unsigned)B, 1024, 0, s>>>(logits, (int)n, temperature, probs);
  else softmax_rows_kernel<256><<<(unsigned)B, 256, 0, s>>>(logits, (int)n, temperature, probs);
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

int uniform_combined(ibc_handle* h, const float* u, int64_t B, int64_t N, int A, const float* mn,
                     const float* mx, uint64_t seed, const float* actions, float* out,
                     cudaStream_t s) {
  const int64_t tot = B * (N + 1) * A;
  if (tot <= 0) return IBC_OK;
  uniform_combined_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(u, B, N, A, mn, mx, seed,
                                                                        actions, out);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

int greedy_actions(ibc_handle* h, const float* probs, const float* values, int64_t B, int64_t n, int A,
                   const float* mn, const float* mx, float* out, cudaStream_t s) {
  if (B <= 0 || n <= 0 || A <= 0) IBC_FAIL(h, IBC_ERR_INVALID, "greedy_actions: bad shape");
  greedy_actions_kernel<<<(unsigned)B, 256, 0, s>>>(probs, values, (int)n, A, mn, mx, out);
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
      stepsize, mn, mx, norm_type, grad_norms, nullptr, nullptr);
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
  float* cur = actions;      // the candidates ping-pong between `actions` and `tmp`
  float* nxt = tmp;
  for (int it = 0; it < num_iterations; ++it) {
    IBC_TRY(energy_fn(cur, E));
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
        cur, rows, R, A, (normals && !last) ? normals + (int64_t)it * R * A : nullptr, seed,
        2ull * it + 2, this_std, mn, mx, nxt);
    IBC_LAUNCH_CHECK(h);
    float* t = cur; cur = nxt; nxt = t;
    if (!last) std *= 0.5f;
  }
  if (cur != actions)
    IBC_CUDA(h, cudaMemcpyAsync(actions, cur, (size_t)R * A * 4, cudaMemcpyDeviceToDevice, s));
  if (probs) IBC_TRY(softmax_rows(h, logits, B, n, 1.0f, probs, s));
  return IBC_OK;
}

int dfo(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n, const float* mn,
        const float* mx, float temperature, int num_iterations, float iteration_std,
        const double* uniforms, const float* normals, uint64_t seed, float* probs,
        cudaStream_t s) {
  // The observation does not change between the iterations: on the layer-fused tensor path its
  // projection hp = obs . Wp[:O] + bp is computed once per call (a 7 us launch for one row).
  const Layout& L = h->L;
  const bool hoist = h->params && h->precision == IBC_PRECISION_TF32 && umma_supported(L) &&
                     !wide_supported(L);
  float* hp = nullptr;
  int it = 0;
  auto energy_fn = [&](const float* acts, float* E) -> int {
    if (!hoist) return mlp_energy(h, obs, B, acts, n, E, s);
    if (it++ == 0) {
      IBC_CUDA(h, h->ws.reserve(Workspace::rounded((size_t)B * L.W * 4)));
      h->ws.reset();
      hp = h->ws.take<float>((size_t)B * L.W);
      return umma_forward(h, obs, B, acts, n, hp, E, nullptr, nullptr, s);
    }
    return umma_forward_hp(h, acts, B * n, n, hp, E, nullptr, nullptr, s);
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
  // delta_action_clip * 0.5 evaluated in double like the Python scalar product
  const float dclip_scale = (float)((double)c.delta_action_clip * 0.5);
  auto chain = [&](const float* o, float* a, const float* nrmls, const uint64_t* seed_ptr,
                   float* ca, float* ce, float* cn, cudaStream_t cs) -> int {
    h->ws_outer.reset();
    float* E = h->ws_outer.take<float>(R);
    float* nrm = h->ws_outer.take<float>(R);
    float* dEda = h->ws_outer.take<float>((size_t)R * A);
    for (int k = 0; k < c.num_iterations; ++k) {
      float* Ek = ce ? ce + (int64_t)k * R : E;
      float* Nk = cn ? cn + (int64_t)k * R : nrm;
      IBC_TRY(mlp_energy_dact(h, o, B, a, n, Ek, dEda, c.apply_exp != 0, cs));
      langevin_update_kernel<<<(unsigned)((R + 127) / 128), 128, 0, cs>>>(
          a, dEda, R, A, nrmls ? nrmls + (int64_t)k * R * A : nullptr, seed, (uint64_t)k + 1,
          c.noise_scale, c.grad_clip, dclip_scale, steps[k], mn, mx, c.grad_norm_type, Nk,
          ca ? ca + (int64_t)k * R * A : nullptr, seed_ptr);
      IBC_LAUNCH_CHECK(h);
    }
    return IBC_OK;
  };
  // per-layer tensor-core stacks: ~20 small launches per iteration -> replay from a graph
  if (h->precision == IBC_PRECISION_TF32 && wide_supported(h->L) && !normals &&
      chain_graphs_enabled() && c.num_iterations >= 4) {
    // the operand copies of the current parameters, outside the capture: a replayed graph
    // never repacks (ibc_params_changed / ibc_bind_params after the capture)
    IBC_TRY(wide_pack_if_stale(h, s));
    return run_chain_graphed(
        h, obs, B, actions, n, mn, mx, c, seed, steps, chain_actions, chain_energies,
        chain_grad_norms, s,
        [&](const float* o, float* a, const uint64_t* seed_ptr, float* ca, float* ce, float* cn,
            cudaStream_t cs) { return chain(o, a, nullptr, seed_ptr, ca, ce, cn, cs); });
  }
  return chain(obs, actions, normals, nullptr, chain_actions, chain_energies, chain_grad_norms, s);
}

}  // namespace ibc
