// fp32 FMA GEMM with fused prologue/epilogue.  This is the IBC_PRECISION_FP32
// path of the dense layers (exact fp32 arithmetic; the tcgen05 path lives in
// mlp_umma.cu) and the fallback for widths the fused kernels do not cover.
#pragma once

#include "common.cuh"

namespace ibc {

enum { AOP_ID = 0, AOP_RELU = 1, AOP_GATE = 2 };

// C[M,N] (+)= f(((op(A) * B) * alpha + bias + rowbias) (.) (gateC > 0) + res)
//   f = identity, or max(., 0) when relu_out
//   op(A)(m,k): A(m,k), relu(A(m,k)) or A(m,k) * (gateA(m,k) > 0)
//   A(m,k) = TA ? A[k*lda + m] : A[m*lda + k]
//   B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
struct GemmP {
  const float* A = nullptr; int64_t lda = 0;
  const float* gateA = nullptr;
  const float* B = nullptr; int64_t ldb = 0;
  float* C = nullptr; int64_t ldc = 0;
  const float* bias = nullptr;
  const float* rowbias = nullptr; int64_t ldrb = 0; int64_t rows_per_group = 1;
  const float* res = nullptr; int64_t ldres = 0;
  const float* gateC = nullptr; int64_t ldgc = 0;
  int M = 0, N = 0, K = 0;
  int aop = AOP_ID;
  int k_chunk = 0;   // K elements per grid.z slice (0 = all of K)
  int atomic = 0;    // accumulate into C with atomicAdd
  int relu_out = 0;  // C = max(C, 0) (not with split-K)
  float alpha = 1.f;
};

int gemm(ibc_handle* h, const GemmP& p, bool ta, bool tb, cudaStream_t s);

// E[r] = sum_w X[r*ldx + w] * v[w] + (b ? b[0] : 0)
int rowdot(ibc_handle* h, const float* X, int64_t ldx, const float* v,
           const float* b, int64_t R, int W, float* E, cudaStream_t s);
// G[r, w] = (scale ? scale[r] : 1) * v[w]
int rowfill(ibc_handle* h, const float* v, const float* scale, int64_t R, int W,
            float* G, cudaStream_t s);
// out[n] += sum_m X[m*ldx + n]   (atomic)
int colsum_add(ibc_handle* h, const float* X, int64_t ldx, int64_t M, int N,
               float* out, cudaStream_t s);
// out[b, w] = sum_{i<n} X[(b*n+i)*W + w]
int groupsum(ibc_handle* h, const float* X, int64_t B, int64_t n, int W,
             float* out, cudaStream_t s);
// out[w] += sum_r X[r*ldx + w] * d[r]   (atomic)   (dwe = H^T dE)
int wcolsum_add(ibc_handle* h, const float* X, int64_t ldx, const float* d,
                int64_t R, int W, float* out, cudaStream_t s);
// out[0] += sum_r d[r] (atomic)
int sum_add(ibc_handle* h, const float* d, int64_t R, float* out, cudaStream_t s);

}  // namespace ibc
