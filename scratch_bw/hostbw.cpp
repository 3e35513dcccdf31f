#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
  const size_t bytes = (size_t)2 << 30;
  float* buf = (float*)aligned_alloc(4096, bytes);
  memset(buf, 0, bytes);
  for (int nt : {1, 2, 4, 8, 12, 16, 24, 32}) {
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t)
        th.emplace_back([=] {
          size_t n = bytes / 4, lo = n * t / nt, hi = n * (t + 1) / nt;
          for (size_t i = lo; i < hi; ++i) buf[i] = (i % 5 == 2) ? 1.0f : 0.0f;   // one-hot-like expand
        });
      for (auto& x : th) x.join();
      double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      best = std::max(best, bytes / s / 1e9);
    }
    printf("threads %2d: %.1f GB/s write\n", nt, best);
  }
  printf("hw threads %u\n", std::thread::hardware_concurrency());
  return 0;
}
