// red_lanes.cu — what a fire-and-forget L2 reduction costs as a function of the ACTIVE LANES of the warp instruction that
// issues it (design-time measurement for the dense HLL form: 58 % of the lanes survive the filter, and most updates do not
// raise their register).  Random 4-byte slots in an L2-resident array; every iteration a lane is active with probability
// p.  Reports warp instructions/s and lane updates/s for red.max / red.add, and for a plain L2 load in place of the
// reduction.  Build & run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/red_lanes.cu -o /tmp/rl && /tmp/rl
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// mode 0: red.max.u32, 1: red.add.u32, 2: ld.global.cg (result folded into a sink), 3: load, then red.max only when it raises
template <int MODE>
__global__ void lanes(uint32_t *tab, uint32_t slots, int iters, uint32_t threshold, unsigned long long *sink,
                      unsigned long long *active, int scatter) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  uint32_t acc = 0, cnt = 0;
  for (int k = 0; k < iters; k++) {
    s = mix(s + k);
    uint32_t slot = s % slots;
    if (scatter) slot = ((slot >> 14) * 2654435761u >> 19) * 16384u + (slot & 16383u);   // the 64 KB chunk of a group: somewhere in 8192 chunks
    const bool on = (mix(s ^ 0x9E3779B9u) >> 8) < threshold;   // 24-bit threshold
    const uint32_t v = (s >> 7) & 0x3Fu;
    cnt += on;
    if (MODE == 0) asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p red.global.max.u32 [%1], %2; }" ::"r"((uint32_t)on), "l"(tab + slot), "r"(v) : "memory");
    if (MODE == 1) asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p red.global.add.u32 [%1], %2; }" ::"r"((uint32_t)on), "l"(tab + slot), "r"(v) : "memory");
    if (MODE == 2) { if (on) acc += __ldcg(tab + slot); }
    if (MODE == 3) {
      uint32_t cur = 0xFFFFFFFFu;
      if (on) cur = __ldcg(tab + slot);
      asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %2; @p red.global.max.u32 [%1], %2; }" ::"r"(cur), "l"(tab + slot), "r"(v) : "memory");
    }
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
  atomicAdd(active, (unsigned long long)cnt);
}

template <int MODE>
static int run(const char *name, uint32_t *tab, uint32_t slots, int ctasPerSm, int threads, float p, unsigned long long *sink,
               unsigned long long *active, int scatter = 0) {
  const int iters = 1024;
  const int ctas = 148 * ctasPerSm;
  const uint32_t threshold = (uint32_t)(p * 16777216.0f);
  CK(cudaMemset(tab, 0, scatter ? (size_t)8192 * 65536 : (size_t)slots * 4));
  CK(cudaMemset(active, 0, 8));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  lanes<MODE><<<ctas, threads>>>(tab, slots, 32, threshold, sink, active, scatter);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(active, 0, 8));
  if (MODE != 3) CK(cudaMemset(tab, 0, scatter ? (size_t)8192 * 65536 : (size_t)slots * 4));   // (mode 3 keeps the warmed-up maxima: most updates do not raise)
  CK(cudaEventRecord(a));
  lanes<MODE><<<ctas, threads>>>(tab, slots, iters, threshold, sink, active, scatter);
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  unsigned long long act; CK(cudaMemcpy(&act, active, 8, cudaMemcpyDeviceToHost));
  const double winst = (double)ctas * threads / 32 * iters;
  printf("%-8s %s %dx%4d thr/SM  p=%.2f (%.1f lanes)  %.3f ms  %.2f G warp-inst/s  %.1f G lane-ops/s\n", name, scatter ? "scattered 64K chunks in 512 MB" : "contiguous", ctasPerSm, threads, p,
         (double)act / winst, ms, winst / ms * 1e-6, (double)act / ms * 1e-6);
  return 0;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("device %s, %d SMs\n", pr.name, pr.multiProcessorCount);
  const uint32_t slots = 101u * 16384u;   // the registers of one cfg4-HLL day-batch
  uint32_t *tab; unsigned long long *sink, *active;
  CK(cudaMalloc(&tab, (size_t)8192 * 65536)); CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&active, 8));
  CK(cudaMemset(sink, 0, 8));
  const float ps[] = {1.0f, 0.58f, 0.25f, 0.09f, 0.03f};
  for (float p : ps) {
    if (run<0>("red.max", tab, slots, 1, 1024, p, sink, active)) return 1;
    if (run<1>("red.add", tab, slots, 1, 1024, p, sink, active)) return 1;
    if (run<2>("ld.cg", tab, slots, 1, 1024, p, sink, active)) return 1;
  }
  if (run<0>("red.max", tab, slots, 8, 256, 1.0f, sink, active)) return 1;
  if (run<0>("red.max", tab, slots, 8, 256, 0.58f, sink, active)) return 1;
  if (run<3>("ld+red", tab, slots, 1, 1024, 1.0f, sink, active)) return 1;
  if (run<3>("ld+red", tab, slots, 1, 1024, 0.58f, sink, active)) return 1;
  // the same registers as 101 chunks of 64 KB scattered over a 512 MB allocation (what a directory addressed by hash slot does)
  if (run<0>("red.max", tab, slots, 1, 1024, 1.0f, sink, active, 1)) return 1;
  if (run<0>("red.max", tab, slots, 1, 1024, 0.58f, sink, active, 1)) return 1;
  if (run<2>("ld.cg", tab, slots, 1, 1024, 0.58f, sink, active, 1)) return 1;
  if (run<3>("ld+red", tab, slots, 1, 1024, 0.58f, sink, active, 1)) return 1;
  return 0;
}
