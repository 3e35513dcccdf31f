// Host-buffer entry points (the e2e path): the same kernels driven from HOST arrays, with the
// host<->device copies pipelined against the kernel in env chunks (double-buffered scratch).
#include <emmintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "common.cuh"

namespace dpt {
static int host_chunk_envs() {
  static const int v = [] {
    const char* e = getenv("DPT_HOST_CHUNK_ENVS");
    const int n = e ? atoi(e) : 0;
    return n >= 256 ? n : 8192;
  }();
  return v;
}

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

static thread_local uint64_t t_last_d2h_bytes = 0;
extern "C" uint64_t dpt_bandit_rollin_host_last_d2h_bytes(void) { return t_last_d2h_bytes; }

extern "C" uint64_t dpt_bandit_rollin_host_scratch_bytes(int N, int H, int d) {
  if (N <= 0 || H <= 0 || d <= 0) return 0;
  const int C = N < host_chunk_envs() ? N : host_chunk_envs();
  return 2 * chunk_layout(C, H, d).total * sizeof(float);
}

// ---- hybrid host pipeline --------------------------------------------------------------------
// PCIe (~56 GB/s) carries 24 B per env-step when the one-hot actions and the rewards come back as fp32, and the host
// cores are otherwise idle.  A fraction of the chunks is therefore produced in COMPACT form (arm index 1 B + reward 4 B
// per step: 5 B over PCIe) and expanded to the one-hot fp32 layout by host threads, while the other chunks keep
// coming back fully formed by DMA.  The compact arm indices are staged in the chunk's own (not yet written)
// next_states region of the caller's buffer, so no extra pinned memory is needed.  Both kinds of chunks give
// bit-identical results (the same Philox counters; tests/test_rollout_gpu.py::test_bandit_rollin_host_path).
namespace {
constexpr int HOST_PARTS = 8;   // host jobs per chunk and phase

// DPT_HOST_COMPACT: "auto" (default: a chunk goes compact whenever the host workers are about to run dry, otherwise by
// DMA -- self-balancing between PCIe and the host cores), "0" (all chunks by DMA) or "1" (all compact).
int compact_policy() {
  const char* e = getenv("DPT_HOST_COMPACT");
  if (!e || !strcmp(e, "auto")) return -1;
  return atoi(e) != 0;
}

// Host-side writers use non-temporal stores: the output arrays are written once and never read here, and a normal
// store would first pull every cache line in from DRAM (read-for-ownership), doubling the memory traffic that
// already competes with the DMA engine.
void fill_ones_nt(float* lo, float* hi) {
  while (lo < hi && (reinterpret_cast<uintptr_t>(lo) & 15)) *lo++ = 1.0f;
  const __m128 one = _mm_set1_ps(1.0f);
  for (; lo + 4 <= hi; lo += 4) _mm_stream_ps(lo, one);
  while (lo < hi) *lo++ = 1.0f;
}

// one-hot rows [lo, hi) of width d from arm indices; 4 rows = d aligned 16 B vectors when lo % 4 == 0
void expand_onehot_nt(float* a, const uint8_t* acts, size_t lo, size_t hi, int d) {
  auto scalar = [&](size_t row) {
    float* o = a + row * d;
    const int arm = acts[row];
    for (int j = 0; j < d; ++j) o[j] = (j == arm) ? 1.0f : 0.0f;
  };
  if ((reinterpret_cast<uintptr_t>(a) & 15) || d > 64) {
    for (size_t row = lo; row < hi; ++row) scalar(row);
    return;
  }
  while (lo < hi && (lo & 3)) scalar(lo++);
  alignas(16) float tmp[4 * 64];
  for (; lo + 4 <= hi; lo += 4) {
    for (int j = 0; j < 4 * d; ++j) tmp[j] = 0.0f;
    tmp[acts[lo]] = 1.0f, tmp[d + acts[lo + 1]] = 1.0f, tmp[2 * d + acts[lo + 2]] = 1.0f, tmp[3 * d + acts[lo + 3]] = 1.0f;
    float* o = a + lo * d;
    for (int v = 0; v < d; ++v) _mm_stream_ps(o + 4 * v, _mm_load_ps(tmp + 4 * v));
  }
  while (lo < hi) scalar(lo++);
}

struct HostJobs {
  int K, C, N, H, d, device;
  std::vector<char> compact;                 // per chunk
  std::vector<cudaEvent_t> d2h_done;         // per chunk, recorded on the copy stream after its D2H
  std::unique_ptr<std::atomic<int>[]> enqueued, phase1_left;
  std::atomic<int> next{0}, compact_parts_pending{0};
  std::atomic<bool> failed{false};
  float *s, *a, *ns, *r;

  void rows_of(int k, int part, size_t& lo, size_t& hi) const {
    const int e0 = k * C, n = std::min(C, N - e0);
    const size_t row0 = (size_t)e0 * H, nrow = (size_t)n * H;
    lo = row0 + nrow * part / HOST_PARTS, hi = row0 + nrow * (part + 1) / HOST_PARTS;
  }
  // job order: chunk k -> HOST_PARTS phase-1 jobs, then (compact chunks only) HOST_PARTS phase-2 jobs
  void run_worker() {
    cudaSetDevice(device);   // a new thread starts on device 0; the events below belong to the caller's device
    const int per_chunk = 2 * HOST_PARTS, total = K * per_chunk;
    for (;;) {
      const int j = next.fetch_add(1);
      if (j >= total) return;
      const int k = j / per_chunk, sub = j - k * per_chunk, part = sub % HOST_PARTS;
      const bool phase2 = sub >= HOST_PARTS;
      size_t lo, hi;
      rows_of(k, part, lo, hi);
      while (!enqueued[k].load(std::memory_order_acquire) && !failed.load()) std::this_thread::yield();   // kind decided
      if (failed.load()) {
        if (!phase2) phase1_left[k].fetch_sub(1);
        continue;
      }
      if (!compact[k]) {                      // DMA chunk: only the constant state columns are host work
        if (!phase2) {
          fill_ones_nt(s + lo, s + hi);
          fill_ones_nt(ns + lo, ns + hi);
        }
        continue;
      }
      if (!phase2) {
        if (cudaEventSynchronize(d2h_done[k]) != cudaSuccess) {
          failed.store(true);
          phase1_left[k].fetch_sub(1);
          compact_parts_pending.fetch_sub(1);
          continue;
        }
        const uint8_t* acts = reinterpret_cast<const uint8_t*>(ns + (size_t)k * C * H) - (size_t)k * C * H;  // acts[row]
        expand_onehot_nt(a, acts, lo, hi, d);
        fill_ones_nt(s + lo, s + hi);
        _mm_sfence();
        phase1_left[k].fetch_sub(1, std::memory_order_release);
        compact_parts_pending.fetch_sub(1);
      } else {                                // the staged arm indices of this chunk are consumed: overwrite them
        while (phase1_left[k].load(std::memory_order_acquire) > 0) std::this_thread::yield();
        fill_ones_nt(ns + lo, ns + hi);
      }
    }
    _mm_sfence();
  }
};
}  // namespace

// ---- host memory write peak (measurement aid for the e2e roofline) ----------------------------------------
// The host half of the e2e path is a pure write stream into the caller's arrays.  This measures what the box's
// cores can stream with non-temporal stores: `n_threads` threads (0 = every core this process may run on) each
// write their own contiguous slice of `dst` (NULL: an internal malloc'ed, first-touched buffer) of `bytes` bytes,
// three passes after one warm-up pass, best pass reported in GB/s.  Debug / bench export, not on the product path.
static int usable_cores() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int n = CPU_COUNT(&set);
    if (n > 0) return n;
  }
  return (int)std::max(1u, std::thread::hardware_concurrency());
}

extern "C" double dpt_host_write_peak(void* dst, uint64_t bytes, int n_threads) {
  if (bytes < (1u << 20)) bytes = 1ull << 30;
  const int T = n_threads > 0 ? n_threads : usable_cores();
  void* own = nullptr;
  if (!dst) {
    if (posix_memalign(&own, 4096, bytes) != 0) return 0.0;
    dst = own;
  }
  float* base = reinterpret_cast<float*>(dst);
  const size_t n = bytes / sizeof(float);
  double best = 0.0;
  for (int pass = 0; pass < 4; ++pass) {
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        float* lo = base + ((n * t / T) & ~size_t(15));
        float* hi = (t == T - 1) ? base + n : base + ((n * (t + 1) / T) & ~size_t(15));
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
        fill_ones_nt(lo, hi);
        _mm_sfence();
      });
    while (ready.load() < T) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& x : th) x.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (pass > 0) best = std::max(best, (double)bytes / s / 1e9);   // pass 0 first-touches the pages
  }
  free(own);
  return best;
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
  const int C = N < host_chunk_envs() ? N : host_chunk_envs();
  const ChunkLayout L = chunk_layout(C, H, d);
  cudaStream_t cs = (cudaStream_t)stream;
  cudaStream_t copy;  // D2H stream, so chunk k's copies overlap chunk k+1's kernel
  DPT_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
  cudaEvent_t done[2], freed[2];
  for (int i = 0; i < 2; ++i) {
    DPT_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    DPT_CUDA(cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming));
  }
  // The bandit state is the constant [1] (envs/bandit_env.py:38): context_states / context_next_states never cross
  // PCIe, host threads write them (and expand the compact chunks) while the GPU pipeline runs.
  HostJobs jobs;
  jobs.K = (N + C - 1) / C, jobs.C = C, jobs.N = N, jobs.H = H, jobs.d = d;
  DPT_CUDA(cudaGetDevice(&jobs.device));
  jobs.s = ctx_states_host, jobs.a = ctx_actions_host, jobs.ns = ctx_next_states_host, jobs.r = ctx_rewards_host;
  jobs.compact.assign(jobs.K, 0);
  jobs.enqueued.reset(new std::atomic<int>[jobs.K]);
  jobs.phase1_left.reset(new std::atomic<int>[jobs.K]);
  const int policy = (jobs.K >= 4 && d <= 255) ? compact_policy() : 0;
  for (int kk = 0; kk < jobs.K; ++kk) {
    jobs.enqueued[kk].store(0);
    jobs.phase1_left[kk].store(HOST_PARTS);
  }
  jobs.d2h_done.resize(jobs.K);
  for (int kk = 0; kk < jobs.K; ++kk) DPT_CUDA(cudaEventCreateWithFlags(&jobs.d2h_done[kk], cudaEventDisableTiming));
  // one process per GPU shares the host cores: torchrun exports LOCAL_WORLD_SIZE
  unsigned procs = 1;
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) procs = (unsigned)std::max(1, atoi(lw));
  const unsigned hw = std::max(2u, std::thread::hardware_concurrency() / procs);
  const size_t total_rows = (size_t)N * H;
  const int n_workers = (int)std::min<size_t>(std::min(16u, hw), std::max<size_t>(1, total_rows >> 18));
  std::vector<std::thread> fillers;
  for (int t = 0; t < n_workers; ++t) fillers.emplace_back([&jobs] { jobs.run_worker(); });
  int rc = DPT_OK;
  int k = 0;
  uint64_t d2h_bytes = 0;
  for (int e0 = 0; e0 < N && rc == DPT_OK; e0 += C, ++k) {
    const int n = (N - e0) < C ? (N - e0) : C;
    const int b = k & 1;
    float* buf = reinterpret_cast<float*>(scratch) + (size_t)b * L.total;
    if (k >= 2) cudaEventSynchronize(freed[b]);        // buffer b's previous D2H has drained (host-side pacing)
    // kind of this chunk: compact when the host workers would otherwise run out of expansion work
    jobs.compact[k] = policy < 0 ? (jobs.compact_parts_pending.load() <= n_workers) : (char)policy;
    if (jobs.compact[k]) jobs.compact_parts_pending.fetch_add(HOST_PARTS);
    cudaMemcpyAsync(buf + L.means, means_host + (size_t)e0 * d, sizeof(float) * n * d, cudaMemcpyHostToDevice, cs);
    const size_t row = (size_t)e0 * H, nrow = (size_t)n * H;
    if (jobs.compact[k]) {
      uint8_t* acts_dev = reinterpret_cast<uint8_t*>(buf + L.a);
      rc = bandit_rollin_compact(buf + L.means, var, seed, env_id0 + (uint64_t)e0, n, H, d, acts_dev, buf + L.r, cs);
      if (rc != DPT_OK) break;
      cudaEventRecord(done[b], cs);
      cudaStreamWaitEvent(copy, done[b], 0);
      cudaMemcpyAsync(ctx_next_states_host + row, acts_dev, nrow, cudaMemcpyDeviceToHost, copy);   // staged arm indices
      cudaMemcpyAsync(ctx_rewards_host + row, buf + L.r, sizeof(float) * nrow, cudaMemcpyDeviceToHost, copy);
    } else {
      rc = dpt_bandit_rollin(buf + L.means, var, DPT_REWARD_GAUSSIAN, seed, env_id0 + (uint64_t)e0, n, H, d, buf + L.s, buf + L.a,
                             buf + L.ns, buf + L.r, nullptr, nullptr, nullptr, cs);
      if (rc != DPT_OK) break;
      cudaEventRecord(done[b], cs);
      cudaStreamWaitEvent(copy, done[b], 0);
      cudaMemcpyAsync(ctx_actions_host + row * d, buf + L.a, sizeof(float) * nrow * d, cudaMemcpyDeviceToHost, copy);
      cudaMemcpyAsync(ctx_rewards_host + row, buf + L.r, sizeof(float) * nrow, cudaMemcpyDeviceToHost, copy);
    }
    d2h_bytes += jobs.compact[k] ? nrow * 5 : nrow * 4 * (uint64_t)(d + 1);
    cudaEventRecord(freed[b], copy);
    cudaEventRecord(jobs.d2h_done[k], copy);
    jobs.enqueued[k].store(1, std::memory_order_release);
  }
  if (rc != DPT_OK) jobs.failed.store(true);
  t_last_d2h_bytes = d2h_bytes;
  for (auto& th : fillers) th.join();
  cudaError_t e1 = cudaStreamSynchronize(copy);
  cudaError_t e2 = cudaStreamSynchronize(cs);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(done[i]), cudaEventDestroy(freed[i]);
  for (auto& ev : jobs.d2h_done) cudaEventDestroy(ev);
  cudaStreamDestroy(copy);
  if (jobs.failed.load() && rc == DPT_OK) {
    set_error("dpt_bandit_rollin_host: a host worker failed while waiting for its chunk");
    return DPT_ERR_CUDA;
  }
  if (rc != DPT_OK) return rc;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("dpt_bandit_rollin_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return DPT_ERR_CUDA;
  }
  return DPT_OK;
}
