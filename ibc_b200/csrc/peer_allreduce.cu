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

__device__ __forceinline__ float adam_one(float g, float& m, float& v, float p, const PeerParams& q) {
  m = q.b1 * m + (1.f - q.b1) * g;
  v = q.b2 * v + (1.f - q.b2) * g * g;
  return p - q.lr_t * m / (sqrtf(v) + q.eps);
}

// One float4 per thread and ONE round of remote loads: a peer read over NVLink has microseconds
// of latency, so the kernel is shaped by how many dependent remote round trips a thread makes
// (a grid-stride loop made four: 25 us), not by the 1 MB it moves.  Everything local (p, m, v,
// this rank's gradient) is loaded BEFORE the wait on the peers' flags, the world - 1 remote loads
// are issued together after it.
__global__ void __launch_bounds__(256) peer_allreduce_adam_kernel(PeerParams q) {
  __shared__ int s_abort;
  const int64_t n4 = q.n >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool body = i < n4;                       // one float4
  const int64_t it = (n4 << 2) + (i - n4);        // or one scalar of the tail (n % 4 entries)
  const bool tail = !body && it < q.n;
  if (threadIdx.x == 0) {
    s_abort = 0;
    // this rank's gradient was written by earlier kernels of the stream: complete, and in the
    // memory its peers read; the release store orders it before the flag
    if (blockIdx.x == 0)
      for (int p = 0; p < q.world; ++p) st_sys(q.flags[p] + q.rank, q.epoch);
  }
  float4 g[kMaxPeers], p4, m4, v4;
  float gt[kMaxPeers], pt = 0.f, mt = 0.f, vt = 0.f;
  if (body) {
    p4 = reinterpret_cast<const float4*>(q.p)[i];
    m4 = reinterpret_cast<const float4*>(q.m)[i];
    v4 = reinterpret_cast<const float4*>(q.v)[i];
    g[0] = __ldcg(reinterpret_cast<const float4*>(q.g[q.rank]) + i);     // parked in slot 0
  } else if (tail) {
    pt = q.p[it]; mt = q.m[it]; vt = q.v[it];
    gt[0] = __ldcg(q.g[q.rank] + it);
  }
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int p = 0; p < q.world && !s_abort; ++p) {
      while (ld_sys(q.flags[q.rank] + p) - q.epoch < 0) {
        if (clock64() - t0 > 3000000000LL) {       // ~1.5 s: a peer is gone
          s_abort = 1;
          if (q.status) atomicCAS(q.status, 0, 6001);
          break;
        }
      }
    }
  }
  __syncthreads();
  if (s_abort) return;
  if (body) {
    const float4 mine = g[0];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < q.world && p != q.rank) g[p] = __ldcg(reinterpret_cast<const float4*>(q.g[p]) + i);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < q.world) {                           // rank order: identical sums on every rank
        const float4 x = p == q.rank ? mine : g[p];
        sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
      }
    p4.x = adam_one(sum.x, m4.x, v4.x, p4.x, q);
    p4.y = adam_one(sum.y, m4.y, v4.y, p4.y, q);
    p4.z = adam_one(sum.z, m4.z, v4.z, p4.z, q);
    p4.w = adam_one(sum.w, m4.w, v4.w, p4.w, q);
    reinterpret_cast<float4*>(q.m)[i] = m4;
    reinterpret_cast<float4*>(q.v)[i] = v4;
    reinterpret_cast<float4*>(q.p)[i] = p4;
    if (q.tail_out && (i << 2) + 3 == q.n - 1) *q.tail_out = sum.w;
  } else if (tail) {
    const float mine = gt[0];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < q.world && p != q.rank) gt[p] = __ldcg(q.g[p] + it);
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < q.world) sum += p == q.rank ? mine : gt[p];
    q.p[it] = adam_one(sum, mt, vt, pt, q);
    q.m[it] = mt;
    q.v[it] = vt;
    if (q.tail_out && it == q.n - 1) *q.tail_out = sum;
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
  if (((uintptr_t)params | (uintptr_t)m | (uintptr_t)v) & 15)
    IBC_FAIL(h, IBC_ERR_INVALID, "peer_allreduce_adam: params / m / v must be 16-byte aligned");
  for (int p = 0; p < world; ++p)
    if (!peer_grads[p] || !peer_flags[p] || ((uintptr_t)peer_grads[p] & 15))
      IBC_FAIL(h, IBC_ERR_INVALID, "peer_allreduce_adam: null or misaligned peer buffer %d", p);
  // one thread per float4 plus n % 4 tail threads (8 x 148 blocks are resident at once: 1.2 M
  // parameters go in one wave, larger nets in several -- blocks only wait on peers, never on
  // each other)
  const int64_t threads = (n >> 2) + (n & 3);
  const int blocks = (int)((threads + 255) / 256);
  peer_allreduce_adam_kernel<<<blocks, 256, 0, s>>>(q);
  IBC_LAUNCH_CHECK(h);
  return IBC_OK;
}

}  // namespace ibc
