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
  int tc = 0;        // tcgen05 tf32 kernel (gemm_tc.cu) when the handle's precision is TF32
  // gemm_tc only: the operand's values are already tf32 (rounded by their producer), so
  // the kernel's in-place rounding pass over that operand is skipped
  int a_tf32 = 0, b_tf32 = 0;
  // Implicit 3x3 'same' convolution (gemm_tc only): when conv_C > 0, A points at an NHWC
  // tensor [*, conv_H, conv_W, conv_C] and the logical A operand is its im2col matrix
  //   A[(b, y, x), (ky, kx, c)] = X[b, y + ky - 1, x + kx - 1, c]   (zero outside the frame)
  // of shape [rows, 9 * conv_C]; lda is ignored.  With `ta` the logical operand is its
  // transpose (weight-gradient form: M = 9 * conv_C, K = rows).
  int conv_H = 0, conv_W = 0, conv_C = 0;
};

// fp32 FMA kernel, or gemm_tc when p.tc is set, the handle computes in tf32 and the
// block is large enough to fill an MMA tile.
int gemm(ibc_handle* h, const GemmP& p, bool ta, bool tb, cudaStream_t s);
// tcgen05 kind::tf32 kernel (operands rounded to nearest tf32, fp32 accumulation);
// with p.atomic the K split is chosen by the kernel's own occupancy rule.
int gemm_tc(ibc_handle* h, const GemmP& p, bool ta, bool tb, cudaStream_t s);
bool gemm_tc_supported(const GemmP& p);

// E[r] = sum_w X[r*ldx + w] * v[w] + (b ? b[0] : 0)
int rowdot(ibc_handle* h, const float* X, int64_t ldx, const float* v,
           const float* b, int64_t R, int W, float* E, cudaStream_t s);
// G[r, w] = (scale ? scale[r] : 1) * v[w]
int rowfill(ibc_handle* h, const float* v, const float* scale, int64_t R, int W,
            float* G, cudaStream_t s);
// X[r, :] = rowbias[r / n, :] + sum_a act[r, a] * Wa[a, :]   (the action half of a first Dense
// layer on tiled observations: K = act_dim is far too short for an MMA tile; a bandwidth-bound
// rank-A update with 16-byte stores; W % 4 == 0)
int act_rows(ibc_handle* h, const float* act, const float* Wa, const float* rowbias, int64_t R,
             int64_t n, int A, int W, float* X, cudaStream_t s);
// Its reverse in one pass over G [B*n, W]: gsum[b, :] = sum of the group's rows and
// dW[a, :] += sum_r act[r, a] * G[r, :]   (atomic).  A <= kMaxActRows, W % 128 == 0.
constexpr int kMaxActRows = 8;
int act_rows_bwd(ibc_handle* h, const float* G, const float* act, int64_t B, int64_t n, int A,
                 int W, float* gsum, float* dW, cudaStream_t s);
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
