// common.cu — small shared host utilities of libalgorithm.so.
#include "common.cuh"

#include <atomic>

namespace aresb {

static std::atomic<unsigned long long> g_launches{0};
void noteLaunches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int smCount() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int &c = cached[dev & 63];
  if (c == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    c = n;
  }
  return c;
}

uint64_t testHash64Mask() {
  static const uint64_t mask = [] {
    const char *e = getenv("ARESDB_B200_TEST_HASH64_MASK");
    return e && *e ? (uint64_t)strtoull(e, nullptr, 16) : ~0ull;
  }();
  return mask;
}

void Scratch::reset(size_t n, cudaStream_t s) {
  release();
  stream = s;
  bytes = n;
  ptr = aresbPoolAllocAsync(n, s);
  if (!ptr) throw EngineError("out of device memory for engine scratch (" + std::to_string(n) + " bytes)");
}

void Scratch::release() {
  if (ptr) aresbPoolFreeAsync(ptr, stream);
  ptr = nullptr;
  bytes = 0;
}

}  // namespace aresb

// Additive (not part of the reference ABI): number of engine kernels launched by this process so far.
extern "C" unsigned long long AresKernelLaunchCount() { return aresb::g_launches.load(); }
