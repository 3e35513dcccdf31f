// Host-buffer entry points (the e2e path): the same kernels driven from HOST arrays, with the
// host<->device copies pipelined against the kernel in env chunks (double-buffered scratch).
#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

namespace dpt {
constexpr int HOST_CHUNK_ENVS = 8192;

struct ChunkLayout {
  size_t means, s, a, ns, r, total;  // float offsets
};
static ChunkLayout chunk_layout(int C, int H, int d) {
  auto up = [](size_t x) { return (x + 63) & ~size_t(63); };  // keep every sub-buffer 256 B aligned
  ChunkLayout L;
  size_t o = 0;
  L.means = o, o += up((size_t)C * d);
  L.s = o, o += up((size_t)C * H);
  L.a = o, o += up((size_t)C * H * d);
  L.ns = o, o += up((size_t)C * H);
  L.r = o, o += up((size_t)C * H);
  L.total = o;
  return L;
}
}  // namespace dpt

using namespace dpt;

extern "C" uint64_t dpt_bandit_rollin_host_scratch_bytes(int N, int H, int d) {
  if (N <= 0 || H <= 0 || d <= 0) return 0;
  const int C = N < HOST_CHUNK_ENVS ? N : HOST_CHUNK_ENVS;
  return 2 * chunk_layout(C, H, d).total * sizeof(float);
}

extern "C" int dpt_bandit_rollin_host(const float* means_host, float var, uint64_t seed, uint64_t env_id0, int N,
                                      int H, int d, float* ctx_states_host, float* ctx_actions_host,
                                      float* ctx_next_states_host, float* ctx_rewards_host, void* scratch,
                                      uint64_t scratch_bytes, void* stream) {
  DPT_CHECK_ARG(N >= 0 && H >= 0 && d >= 1, "dpt_bandit_rollin_host: bad sizes N=%d H=%d d=%d", N, H, d);
  if (N == 0 || H == 0) return DPT_OK;
  DPT_CHECK_ARG(means_host && ctx_states_host && ctx_actions_host && ctx_next_states_host && ctx_rewards_host,
                "dpt_bandit_rollin_host: null host pointer");
  DPT_CHECK_ARG(scratch && scratch_bytes >= dpt_bandit_rollin_host_scratch_bytes(N, H, d),
                "dpt_bandit_rollin_host: scratch too small (%llu < %llu bytes)", (unsigned long long)scratch_bytes,
                (unsigned long long)dpt_bandit_rollin_host_scratch_bytes(N, H, d));
  const int C = N < HOST_CHUNK_ENVS ? N : HOST_CHUNK_ENVS;
  const ChunkLayout L = chunk_layout(C, H, d);
  cudaStream_t cs = (cudaStream_t)stream;
  cudaStream_t copy;  // D2H stream, so chunk k's copies overlap chunk k+1's kernel
  DPT_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
  cudaEvent_t done[2], freed[2];
  for (int i = 0; i < 2; ++i) {
    DPT_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    DPT_CUDA(cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming));
  }
  // The bandit state is the constant [1] (envs/bandit_env.py:38), so context_states / context_next_states
  // are filled on the host by a few threads while the GPU pipeline runs, instead of crossing PCIe (25 % of
  // the bytes of a collection).
  const size_t total_rows = (size_t)N * H;
  const int n_fill = (int)std::min<size_t>(4, std::max<size_t>(1, total_rows >> 20));
  std::vector<std::thread> fillers;
  for (int t = 0; t < n_fill; ++t)
    fillers.emplace_back([=] {
      const size_t lo = total_rows * t / n_fill, hi = total_rows * (t + 1) / n_fill;
      std::fill(ctx_states_host + lo, ctx_states_host + hi, 1.0f);
      std::fill(ctx_next_states_host + lo, ctx_next_states_host + hi, 1.0f);
    });
  int rc = DPT_OK;
  int k = 0;
  for (int e0 = 0; e0 < N && rc == DPT_OK; e0 += C, ++k) {
    const int n = (N - e0) < C ? (N - e0) : C;
    const int b = k & 1;
    float* buf = reinterpret_cast<float*>(scratch) + (size_t)b * L.total;
    if (k >= 2) cudaStreamWaitEvent(cs, freed[b], 0);  // buffer b's previous D2H has drained
    cudaMemcpyAsync(buf + L.means, means_host + (size_t)e0 * d, sizeof(float) * n * d, cudaMemcpyHostToDevice, cs);
    rc = dpt_bandit_rollin(buf + L.means, var, DPT_REWARD_GAUSSIAN, seed, env_id0 + (uint64_t)e0, n, H, d, buf + L.s, buf + L.a,
                           buf + L.ns, buf + L.r, nullptr, nullptr, nullptr, cs);
    if (rc != DPT_OK) break;
    cudaEventRecord(done[b], cs);
    cudaStreamWaitEvent(copy, done[b], 0);
    const size_t row = (size_t)e0 * H, nrow = (size_t)n * H;
    cudaMemcpyAsync(ctx_actions_host + row * d, buf + L.a, sizeof(float) * nrow * d, cudaMemcpyDeviceToHost, copy);
    cudaMemcpyAsync(ctx_rewards_host + row, buf + L.r, sizeof(float) * nrow, cudaMemcpyDeviceToHost, copy);
    cudaEventRecord(freed[b], copy);
  }
  for (auto& th : fillers) th.join();
  cudaError_t e1 = cudaStreamSynchronize(copy);
  cudaError_t e2 = cudaStreamSynchronize(cs);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(done[i]), cudaEventDestroy(freed[i]);
  cudaStreamDestroy(copy);
  if (rc != DPT_OK) return rc;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("dpt_bandit_rollin_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return DPT_ERR_CUDA;
  }
  return DPT_OK;
}
