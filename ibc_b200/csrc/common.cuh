// Shared host/device helpers for the ibc_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/ibc_b200.h"

namespace ibc {

extern std::atomic<long long> g_launch_count;
extern thread_local char g_create_error[512];

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// Parameter offsets in the flat vector (see include/ibc_b200.h).
struct Layout {
  int O, A, W, D, nb;
  int64_t wp, bp, we, be, count;
  __host__ __device__ int64_t w1(int j) const { return blk(j); }
  __host__ __device__ int64_t b1(int j) const { return blk(j) + (int64_t)W * W; }
  __host__ __device__ int64_t w2(int j) const { return blk(j) + (int64_t)W * W + W; }
  __host__ __device__ int64_t b2(int j) const { return blk(j) + 2 * (int64_t)W * W + W; }
  __host__ __device__ int64_t blk(int j) const {
    return bp + W + (int64_t)j * (2 * (int64_t)W * W + 2 * W);
  }
};

inline Layout make_layout(const ibc_mlp_arch& a) {
  Layout L;
  L.O = a.obs_dim; L.A = a.act_dim; L.W = a.width; L.D = a.depth; L.nb = a.depth / 2;
  L.wp = 0;
  L.bp = (int64_t)(L.O + L.A) * L.W;
  L.we = L.blk(L.nb);
  L.be = L.we + L.W;
  L.count = L.be + 1;
  return L;
}

// Growable device scratch owned by a handle.  Pointers stay valid until the
// next grow; callers carve it per call.
struct Workspace {
  char* base = nullptr;
  size_t cap = 0;
  size_t used = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (base) {
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) return e;
      cudaFree(base);
      base = nullptr; cap = 0;
    }
    size_t want = bytes + (bytes >> 3) + (1 << 20);
    cudaError_t e = cudaMalloc((void**)&base, want);
    if (e != cudaSuccess) return e;
    cap = want;
    return cudaSuccess;
  }
  void reset() { used = 0; }
  template <typename T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    T* p = (T*)(base + used);
    used += bytes;
    return p;
  }
  static size_t rounded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};

// One captured Langevin chain (mlp_umma_train.cu::langevin_umma): the chain is
// launch-bound (three ~15 us kernels per iteration, 100 iterations), so it is
// recorded once into a CUDA graph that runs on handle-owned staging buffers.
struct LangevinGraph {
  std::vector<unsigned char> key;    // shapes, config, step sizes, workspace / parameter identity
  bool warm = false;                 // the chain has run once un-captured (lazy kernel attributes)
  cudaGraphExec_t exec = nullptr;
  float* obs = nullptr;              // staging: [B, O]
  float* act = nullptr;              //          [R, A]
  float* chain_actions = nullptr;    //          [K, R, A] when requested
  float* chain_energies = nullptr;   //          [K, R]
  float* chain_norms = nullptr;      //          [K, R]
  uint64_t* seed = nullptr;          // device copy of the call's Philox seed
  uint64_t last_use = 0;
  void release() {
    if (exec) cudaGraphExecDestroy(exec);
    cudaFree(obs); cudaFree(act); cudaFree(chain_actions); cudaFree(chain_energies);
    cudaFree(chain_norms); cudaFree(seed);
    *this = LangevinGraph();
  }
};

}  // namespace ibc

struct ibc_handle {
  ibc_mlp_arch arch;
  ibc::Layout L;
  int device = 0;
  int precision = IBC_PRECISION_FP32;
  int sm_count = 148;
  const float* params = nullptr;
  bool packed_stale = true;       // the operand copies every layer-fused step needs (umma_pack_core)
  bool packed_rest_stale = true;  // ... the remaining ones (reverse chains in tf32, shared-memory-operand
                                  // kernels, fp16 forward layers): umma_pack_params brings both up to date
  float* packed = nullptr;        // tcgen05 operand copies of the weights
  size_t packed_bytes = 0;
  void* packed16 = nullptr;       // fp16 operand copies (width 128, langevin_f16.cu)
  // Recorded on the stream of the last ibc_train_grads right behind its loss kernel (the
  // per-example losses are final; the backward follows): ibc_stream_wait_loss
  cudaEvent_t loss_event = nullptr;
  bool loss_event_recorded = false;
  int* dev_status = nullptr;      // kernel-side barrier-timeout flag
  bool profiling = false;         // record CUDA events around the training kernels
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;          // a full set of events has been recorded
  ibc::Workspace ws;        // dense-stack passes (reset by each pass)
  ibc::Workspace ws_outer;  // sampler / trainer scratch that outlives a pass
  std::vector<ibc::LangevinGraph> langevin_graphs;
  cudaStream_t capture_stream = nullptr;   // graphs are recorded here (the caller's stream may
                                           // be the legacy default stream, which cannot capture)
  uint64_t graph_clock = 0;
  std::string err;
};

#define IBC_FAIL(h, code, ...)                                  \
  do {                                                          \
    char _buf[512];                                             \
    snprintf(_buf, sizeof(_buf), __VA_ARGS__);                  \
    if (h) (h)->err = _buf;                                     \
    else strncpy(ibc::g_create_error, _buf, 511);               \
    return (code);                                              \
  } while (0)

#define IBC_CUDA(h, expr)                                                     \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess)                                                    \
      IBC_FAIL(h, IBC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
               cudaGetErrorString(_e), __FILE__, __LINE__);                   \
  } while (0)

#define IBC_LAUNCH_CHECK(h)                                                   \
  do {                                                                        \
    ibc::count_launch();                                                      \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess)                                                    \
      IBC_FAIL(h, IBC_ERR_CUDA, "kernel launch failed: %s (%s:%d)",           \
               cudaGetErrorString(_e), __FILE__, __LINE__);                   \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember which devices a call
// site has raised it on (a process-wide flag left the second device of a multi-GPU process
// at the 48 KB default).
#define IBC_RAISE_SMEM(h, kernel, bytes)                                                   \
  do {                                                                                     \
    static std::atomic<unsigned long long> _done{0};                                       \
    int _d = 0;                                                                            \
    if (h) _d = (h)->device; else cudaGetDevice(&_d);                                      \
    if (_d < 0 || _d >= 64 || !((_done.load(std::memory_order_relaxed) >> _d) & 1ull)) {   \
      IBC_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)(bytes)));                                     \
      if (_d >= 0 && _d < 64) _done.fetch_or(1ull << _d, std::memory_order_relaxed);        \
    }                                                                                      \
  } while (0)

#define IBC_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != IBC_OK) return _rc; \
  } while (0)
