/* Oracle (TEST INFRASTRUCTURE, not product code): plain-C restatement of the
 * categorical draw + bincount the reference runs on CPU:0
 *   ibc/agents/mcmc.py:32-55   categorical_bincount
 *   ibc/agents/mcmc.py:136-140 tf.repeat(range(B*n), counts)
 * whose arithmetic lives in TensorFlow 2.6.0
 * tensorflow/core/kernels/multinomial_op.cc (MultinomialFunctor<CPUDevice>):
 * max over finite logits; cdf accumulated sequentially in fp64 over
 * exp(double(logit) - double(max)), non-finite logits skipped;
 * to_find = u * total; index = upper_bound(cdf, to_find).
 * The exp is the portable mul/add-only exp64 (see oracle/mcmc.py) so that
 * NumPy, this file and the CUDA kernel agree bit-for-bit.  Compile with
 * -ffp-contract=off (oracle/Makefile).  Parity vs TF bits: unpinned.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const double kLog2e = 1.4426950408889634;
static const double kLn2Hi = 6.93147180369123816490e-01;
static const double kLn2Lo = 1.90821492927058770002e-10;
static const double kCoef[14] = {
    1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040,
    1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,
    1.0 / 479001600, 1.0 / 6227020800.0};

double ibc_oracle_exp64(double x) {
  if (x < -708.0) return 0.0;
  double k = rint(x * kLog2e);
  double r = (x - k * kLn2Hi) - k * kLn2Lo;
  double p = kCoef[13];
  for (int i = 12; i >= 0; --i) p = p * r + kCoef[i];
  uint64_t bits = (uint64_t)((int64_t)k + 1023) << 52;
  double scale;
  memcpy(&scale, &bits, sizeof(scale));
  return p * scale;
}

/* logits [B, n] f32; uniforms [B, count] f64 in [0,1).
 * draws [B, count] raw class ids (may be NULL); counts [B, n] (may be NULL);
 * sorted_rows [B * count] global row ids b*n + class, ascending (may be NULL).
 */
int ibc_oracle_resample(const float* logits, const double* uniforms, int B,
                        int n, int count, int32_t* draws, int32_t* counts,
                        int64_t* sorted_rows) {
  double* cdf = (double*)malloc(sizeof(double) * (size_t)n);
  int32_t* cnt = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  if (!cdf || !cnt) return -1;
  for (int b = 0; b < B; ++b) {
    const float* row = logits + (size_t)b * n;
    float mx = -3.402823466e+38f; /* numeric_limits<float>::lowest() */
    for (int j = 0; j < n; ++j)
      if (isfinite(row[j]) && row[j] > mx) mx = row[j];
    double total = 0.0;
    for (int j = 0; j < n; ++j) {
      if (isfinite(row[j])) total += ibc_oracle_exp64((double)row[j] - (double)mx);
      cdf[j] = total;
    }
    memset(cnt, 0, sizeof(int32_t) * (size_t)n);
    for (int s = 0; s < count; ++s) {
      double to_find = uniforms[(size_t)b * count + s] * total;
      int lo = 0, hi = n; /* std::upper_bound: first j with cdf[j] > to_find */
      while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (cdf[mid] > to_find) hi = mid; else lo = mid + 1;
      }
      if (lo > n - 1) lo = n - 1;
      if (draws) draws[(size_t)b * count + s] = lo;
      cnt[lo]++;
    }
    if (counts) memcpy(counts + (size_t)b * n, cnt, sizeof(int32_t) * (size_t)n);
    if (sorted_rows) {
      size_t o = (size_t)b * count;
      for (int j = 0; j < n; ++j)
        for (int c = 0; c < cnt[j]; ++c) sorted_rows[o++] = (int64_t)b * n + j;
    }
  }
  free(cdf);
  free(cnt);
  return 0;
}
