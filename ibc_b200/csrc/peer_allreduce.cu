// Data-parallel gradient all-reduce + Adam in ONE kernel over NVLink peer memory
// (the MirroredStrategy sum of ibc/agents/base_agent.py:50-58 followed by the Keras Adam update of
// ibc/train/get_agent.py:76-84).
//
// Why: the flat gradient is 1.07 MB.  NCCL's all-reduce of it costs ~25-40 us of latency per
// step (kernel launch on its own stream, two event hand-overs, a ring / tree protocol meant for
// large buffers) against a 0.5 ms step -- 4 % at two GPUs, 8 % at four -- and the Adam kernel and
// the re-pack wait behind it.  With every rank's gradient buffer mapped into every other rank's
// address space (CUDA IPC; NVSwitch gives each GPU full bandwidth to each peer) a rank simply
// READS its peers' gradients: one launch that (1) publishes "my gradient of step t is complete" to
// every peer's flag array and waits until all peers have published, (2) sums the world's
// gradients element by element in rank order (the same order on every rank: the replicas stay
// bit-identical) and (3) applies Adam to its own replica of the parameters.  Gradients are
// double-buffered by step parity: a rank can only overwrite a buffer two steps later, after the
// barrier of the step in between, which no peer passes before it has finished reading.
// The wait is bounded (about a second): a missing peer raises status 6001 instead of hanging.
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace ibc {

namespace {

constexpr int kMaxPeers = 8;

struct PeerParams {
  const float* g[kMaxPeers];   // every rank's gradient buffer of this step's parity
  int* flags[kMaxPeers];       // every rank's flag array: flags[p][r] = last step rank r published to p
  int rank, world;
  int epoch;
  int64_t n;
  float* p;
  float* m;
  float* v;
  float lr_t, b1, b2, eps;
  float* tail_out;             // optional: the reduced last entry (finiteness probe of the caller)
  int* status;
};

__device__ __forceinline__ int ld_sys(const int* ptr) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(int* ptr, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256) peer_allreduce_adam_kernel(PeerParams q) {
  __shared__ int s_abort;
  if (threadIdx.x == 0) {
    s_abort = 0;
    if (blockIdx.x == 0) {
      // this rank's gradient was written by earlier kernels of the stream: complete and visible
      __threadfence_system();
      for (int p = 0; p < q.world; ++p) st_sys(q.flags[p] + q.rank, q.epoch);
    }
    const long long t0 = clock64();
    for (int p = 0; p < q.world && !s_abort; ++p) {
      while (ld_sys(q.flags[q.rank] + p) - q.epoch < 0) {
        if (clock64() - t0 > 3000000000LL) {       // ~1.5 s: a peer is gone
          s_abort = 1;
          if (q.status) atomicCAS(q.status, 0, 6001);
          break;
        }
        __nanosleep(200);
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (s_abort) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gi = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < q.world) gi += __ldcg(q.g[p] + i);      // rank order: identical sums on every rank
    const float mi = q.b1 * q.m[i] + (1.f - q.b1) * gi;
    const float vi = q.b2 * q.v[i] + (1.f - q.b2) * gi * gi;
    q.m[i] = mi;
    q.v[i] = vi;
    q.p[i] = q.p[i] - q.lr_t * mi / (sqrtf(vi) + q.eps);
    if (q.tail_out && i == q.n - 1) *q.tail_out = gi;
  }
}

}  // namespace

// Peer-visible allocations are made and mapped by the library itself: one cudaMalloc per rank
// (its IPC handle names exactly this allocation, offset 0), opened by every peer in the context of
// ITS OWN device with cudaIpcMemLazyEnablePeerAccess -- the mapping a kernel on that device can
// dereference.  (A handle opened in a context of the exporting device, which is what a framework's
// tensor sharing does, is not addressable from the opening rank's own GPU.)
int peer_alloc(ibc_handle* h, int64_t bytes, void** ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  if (bytes <= 0 || !ptr || !handle64) IBC_FAIL(h, IBC_ERR_INVALID, "peer_alloc: bad argument");
  void* p = nullptr;
  IBC_CUDA(h, cudaMalloc(&p, (size_t)bytes));
  IBC_CUDA(h, cudaMemset(p, 0, (size_t)bytes));
  IBC_CUDA(h, cudaDeviceSynchronize());
  cudaIpcMemHandle_t hd;
  const cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    IBC_FAIL(h, IBC_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &hd, sizeof(hd));
  *ptr = p;
  return IBC_OK;
}

int peer_open(ibc_handle* h, const void* handle64, void** ptr) {
  if (!ptr || !handle64) IBC_FAIL(h, IBC_ERR_INVALID, "peer_open: bad argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, sizeof(hd));
  IBC_CUDA(h, cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return IBC_OK;
}

int peer_close(ibc_handle* h, void* ptr, int owned) {
  if (!ptr) return IBC_OK;
  if (owned) IBC_CUDA(h, cudaFree(ptr));
  else IBC_CUDA(h, cudaIpcCloseMemHandle(ptr));
  return IBC_OK;
}

int peer_allreduce_adam(ibc_handle* h, const float* const* peer_grads, int* const* peer_flags, int rank,
                        int world, int epoch, int64_t n, float* params, float* m, float* v, float lr_t,
                        float b1, float b2, float eps, float* tail_out, int* status, cudaStream_t s) {
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world || n <= 0)
    IBC_FAIL(h, IBC_ERR_INVALID, "peer_allreduce_adam: bad world / rank / count");
  PeerParams q;
  for (int p = 0; p < kMaxPeers; ++p) {
    q.g[p] = p < world ? peer_grads[p] : nullptr;
    q.flags[p] = p < world ? peer_flags[p] : nullptr;
  }
  q.rank = rank; q.world = world; q.epoch = epoch; q.n = n; q.p = params; q.m = m; q.v = v;
  q.lr_t = lr_t; q.b1 = b1; q.b2 = b2; q.eps = eps; q.tail_out = tail_out; q.status = status;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 296) blocks = 296;
  peer_allreduce_adam_kernel<<<blocks, 256, 0, s>>>(q);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

}  // namespace ibc
