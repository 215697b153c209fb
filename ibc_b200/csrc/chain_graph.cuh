// CUDA-graph replay of a Langevin chain (launch-bound: ~3 to ~20 small kernels per
// iteration, 100 iterations).  The second call with the same key records the chain on a
// handle-owned stream; later calls stage their inputs into the graph's buffers, write
// the seed to device memory and launch it.  Used by the width-256 fused path
// (mlp_umma_train.cu) and by the per-layer wide path (samplers.cu).
#pragma once

#include <vector>

#include "common.cuh"

namespace ibc {

// chain(obs, actions, seed_ptr, chain_actions, chain_energies, chain_norms, stream)
template <class Chain>
int run_chain_graphed(ibc_handle* h, const float* obs, int64_t B, float* actions, int64_t n,
                      const float* mn, const float* mx, const ibc_langevin_cfg& c, uint64_t seed,
                      const std::vector<float>& steps, float* chain_actions,
                      float* chain_energies, float* chain_grad_norms, cudaStream_t s,
                      Chain chain) {
  const Layout& L = h->L;
  const int64_t R = B * n;
  const int K = c.num_iterations;
  std::vector<unsigned char> key;
  auto put = [&](const void* p, size_t nbytes) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    key.insert(key.end(), b, b + nbytes);
  };
  const void* idents[6] = {h->params, h->packed, h->ws.base, h->ws_outer.base, mn, mx};
  const int flags[3] = {chain_actions != nullptr, chain_energies != nullptr,
                        chain_grad_norms != nullptr};
  put(&B, sizeof(B)); put(&n, sizeof(n)); put(&c, sizeof(c)); put(idents, sizeof(idents));
  put(flags, sizeof(flags)); put(&h->precision, sizeof(int));
  put(steps.data(), steps.size() * sizeof(float));
  LangevinGraph* g = nullptr;
  for (auto& e : h->langevin_graphs)
    if (e.key == key) g = &e;
  if (!g) {
    if (h->langevin_graphs.size() >= 4) {        // evict the least recently used
      size_t lru = 0;
      for (size_t i = 1; i < h->langevin_graphs.size(); ++i)
        if (h->langevin_graphs[i].last_use < h->langevin_graphs[lru].last_use) lru = i;
      IBC_CUDA(h, cudaStreamSynchronize(s));
      h->langevin_graphs[lru].release();
      h->langevin_graphs.erase(h->langevin_graphs.begin() + lru);
    }
    h->langevin_graphs.emplace_back();
    g = &h->langevin_graphs.back();
    g->key = key;
    IBC_CUDA(h, cudaMalloc((void**)&g->obs, (size_t)B * L.O * 4));
    IBC_CUDA(h, cudaMalloc((void**)&g->act, (size_t)R * L.A * 4));
    IBC_CUDA(h, cudaMalloc((void**)&g->seed, sizeof(uint64_t)));
    if (chain_actions) IBC_CUDA(h, cudaMalloc((void**)&g->chain_actions, (size_t)K * R * L.A * 4));
    if (chain_energies) IBC_CUDA(h, cudaMalloc((void**)&g->chain_energies, (size_t)K * R * 4));
    if (chain_grad_norms) IBC_CUDA(h, cudaMalloc((void**)&g->chain_norms, (size_t)K * R * 4));
  }
  g->last_use = ++h->graph_clock;
  IBC_CUDA(h, cudaMemcpyAsync(g->obs, obs, (size_t)B * L.O * 4, cudaMemcpyDeviceToDevice, s));
  IBC_CUDA(h, cudaMemcpyAsync(g->act, actions, (size_t)R * L.A * 4, cudaMemcpyDeviceToDevice, s));
  IBC_CUDA(h, cudaMemcpyAsync(g->seed, &seed, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  if (!g->exec && g->warm) {
    cudaGraph_t graph = nullptr;
    if (!h->capture_stream)
      IBC_CUDA(h, cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
    cudaStream_t cs = h->capture_stream;
    IBC_CUDA(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const int rc = chain(g->obs, g->act, g->seed, g->chain_actions, g->chain_energies,
                         g->chain_norms, cs);
    const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (rc != IBC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    IBC_CUDA(h, ce);
    const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    IBC_CUDA(h, ie);
  }
  if (g->exec) {
    IBC_CUDA(h, cudaGraphLaunch(g->exec, s));
    count_launch();
  } else {
    IBC_TRY(chain(g->obs, g->act, g->seed, g->chain_actions, g->chain_energies, g->chain_norms, s));
    g->warm = true;
  }
  IBC_CUDA(h, cudaMemcpyAsync(actions, g->act, (size_t)R * L.A * 4, cudaMemcpyDeviceToDevice, s));
  if (chain_actions)
    IBC_CUDA(h, cudaMemcpyAsync(chain_actions, g->chain_actions, (size_t)K * R * L.A * 4,
                                cudaMemcpyDeviceToDevice, s));
  if (chain_energies)
    IBC_CUDA(h, cudaMemcpyAsync(chain_energies, g->chain_energies, (size_t)K * R * 4,
                                cudaMemcpyDeviceToDevice, s));
  if (chain_grad_norms)
    IBC_CUDA(h, cudaMemcpyAsync(chain_grad_norms, g->chain_norms, (size_t)K * R * 4,
                                cudaMemcpyDeviceToDevice, s));
  return IBC_OK;
}

// graphs are off with IBC_NO_GRAPHS=1 and while the timeline instrumentation is on
inline bool chain_graphs_enabled() {
  static const bool off =
      getenv("IBC_NO_GRAPHS") != nullptr || getenv("IBC_DBG_TIMELINE") != nullptr;
  return !off;
}

}  // namespace ibc
