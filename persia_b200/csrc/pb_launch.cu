// pb_launch.cu — launch bookkeeping: the launch counter and the optional per-kernel-family event timing.
#include "pb_device.cuh"

#include <vector>

namespace pb {

std::atomic<uint64_t> g_launches{0};

// Optional per-kernel-family timing with CUDA events on the launching stream (bench.py's roofline leg).
// Off by default: the hot path then pays one relaxed atomic increment per launch and nothing else.
struct Profiler {
  uint32_t mask = 0;  // families being timed
  static constexpr int MAX_EV = 1 << 15;
  std::vector<cudaEvent_t> ev;  // pairs
  std::vector<int> fam;
  int used = 0;
  int cur = -1;  // family of the launch being bracketed, -1 = not timed
};
Profiler g_prof;

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void prof_begin(int family, cudaStream_t st) {
  if (!((g_prof.mask >> family) & 1u)) {
    g_prof.cur = -1;
    return;
  }
  g_prof.cur = family;
  if ((int)g_prof.ev.size() < 2 * (g_prof.used + 1)) {
    if (g_prof.used >= Profiler::MAX_EV) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    g_prof.ev.push_back(a);
    g_prof.ev.push_back(b);
    g_prof.fam.push_back(family);
  }
  g_prof.fam[g_prof.used] = family;
  cudaEventRecord(g_prof.ev[2 * g_prof.used], st);
}
void prof_end(cudaStream_t st) {
  if (g_prof.cur < 0 || g_prof.used >= Profiler::MAX_EV || (int)g_prof.ev.size() < 2 * (g_prof.used + 1)) return;
  cudaEventRecord(g_prof.ev[2 * g_prof.used + 1], st);
  g_prof.used++;
}
bool profiling() { return g_prof.mask != 0; }

void profile_enable(uint32_t family_mask) {
  g_prof.mask = family_mask;
  g_prof.used = 0;
  g_prof.cur = -1;
}
// sums elapsed ms and launch counts per family; call after synchronising the stream(s)
void profile_read(double* ms, uint64_t* count, int n_families) {
  for (int i = 0; i < n_families; ++i) {
    ms[i] = 0;
    count[i] = 0;
  }
  for (int i = 0; i < g_prof.used; ++i) {
    float t = 0;
    if (cudaEventElapsedTime(&t, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]) == cudaSuccess && g_prof.fam[i] < n_families) {
      ms[g_prof.fam[i]] += t;
      count[g_prof.fam[i]]++;
    }
  }
  g_prof.used = 0;
}


uint64_t launch_count() { return g_launches.load(); }

}  // namespace pb
