// agg_microbench.cu — design-time measurements behind the aggregation strategy of the fused
// kernel (DESIGN.md "Aggregation"): streaming-read ceiling, shared-memory atomic throughput
// (CTA-private tables) and L2 atomic throughput (global tables) on B200, by element type and
// table size.  Build & run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/agg_microbench.cu -o /tmp/mb && /tmp/mb
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__global__ void streamRead(const uint4 *__restrict__ p, size_t n, unsigned long long *out) {
  uint32_t acc = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint4 v = __ldg(p + i);
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

template <typename T>
__global__ void smemAtomics(int slots, int iters, unsigned long long *sink) {
  extern __shared__ unsigned char raw[];
  T *tab = reinterpret_cast<T *>(raw);
  for (int i = threadIdx.x; i < slots; i += blockDim.x) tab[i] = T(0);
  __syncthreads();
  uint32_t s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
  for (int k = 0; k < iters; k++) {
    uint32_t slot = lcg(s) % (uint32_t)slots;
    atomicAdd(&tab[slot], T(1));
  }
  __syncthreads();
  if (threadIdx.x == 0 && tab[0] == T(123456789)) atomicAdd(sink, 1ull);
}

template <typename T>
__global__ void globalAtomics(T *tab, uint32_t slots, int iters) {
  uint32_t s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
  for (int k = 0; k < iters; k++) {
    uint32_t slot = lcg(s) % slots;
    atomicAdd(&tab[slot], T(1));
  }
}

// hash-table shaped update: read the 8-byte key of a slot (plain load), then atomicAdd its accumulator
template <typename T>
__global__ void globalProbeAdd(const unsigned long long *keys, T *acc, uint32_t slots, int iters, unsigned long long *sink) {
  uint32_t s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
  unsigned long long miss = 0;
  for (int k = 0; k < iters; k++) {
    uint32_t slot = lcg(s) % slots;
    if (keys[slot] != (unsigned long long)slot) miss++;
    atomicAdd(&acc[slot], T(1));
  }
  if (miss) atomicAdd(sink, miss);
}

__global__ void matchAny(int iters, unsigned long long *sink) {
  uint32_t s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
  uint32_t acc = 0;
  for (int k = 0; k < iters; k++) acc += __popc(__match_any_sync(0xffffffffu, lcg(s) & 255u));
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <typename F>
float timeIt(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  unsigned long long *sink;
  CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));

  {  // A: streaming read
    size_t bytes = 4ull << 30;
    uint4 *p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 1, bytes));
    for (int bps : {4, 8, 16}) {
      float ms = timeIt([&] { streamRead<<<sms * bps, 256>>>(p, bytes / 16, sink); });
      printf("A stream-read 4GiB  blocks/SM=%2d  %.3f ms  %.1f GB/s\n", bps, ms, bytes / ms / 1e6);
    }
    cudaFree(p);
  }
  const int iters = 2048;
  const int bps = 4, threads = 256;
  double opsS = (double)sms * bps * threads * iters;
  for (int slots : {128, 2048, 4096}) {
    float ms;
    ms = timeIt([&] { smemAtomics<unsigned int><<<sms * bps, threads, slots * 8>>>(slots, iters, sink); });
    printf("B smem atomicAdd u32  slots=%5d  %.3f ms  %.1f Gop/s\n", slots, ms, opsS / ms / 1e6);
    ms = timeIt([&] { smemAtomics<unsigned long long><<<sms * bps, threads, slots * 8>>>(slots, iters, sink); });
    printf("B smem atomicAdd u64  slots=%5d  %.3f ms  %.1f Gop/s\n", slots, ms, opsS / ms / 1e6);
    ms = timeIt([&] { smemAtomics<float><<<sms * bps, threads, slots * 8>>>(slots, iters, sink); });
    printf("B smem atomicAdd f32  slots=%5d  %.3f ms  %.1f Gop/s\n", slots, ms, opsS / ms / 1e6);
    ms = timeIt([&] { smemAtomics<double><<<sms * bps, threads, slots * 8>>>(slots, iters, sink); });
    printf("B smem atomicAdd f64  slots=%5d  %.3f ms  %.1f Gop/s\n", slots, ms, opsS / ms / 1e6);
  }
  {
    void *tab; size_t maxSlots = 1u << 26;
    CK(cudaMalloc(&tab, maxSlots * 8)); CK(cudaMemset(tab, 0, maxSlots * 8));
    unsigned long long *keys; CK(cudaMalloc(&keys, maxSlots * 8));
    {
      // keys[i] = i
      unsigned long long *h = (unsigned long long *)malloc(maxSlots * 8);
      for (size_t i = 0; i < maxSlots; i++) h[i] = i;
      CK(cudaMemcpy(keys, h, maxSlots * 8, cudaMemcpyHostToDevice));
      free(h);
    }
    const int it2 = 512;
    const int bps2 = 8;
    double ops = (double)sms * bps2 * threads * it2;
    for (uint32_t slots : {100u, 2400u, 19200u, 1u << 20, 1u << 23, 1u << 26}) {
      float ms;
      ms = timeIt([&] { globalAtomics<unsigned int><<<sms * bps2, threads>>>((unsigned int *)tab, slots, it2); });
      printf("C global atomicAdd u32  slots=%9u  %.3f ms  %.1f Gop/s\n", slots, ms, ops / ms / 1e6);
      ms = timeIt([&] { globalAtomics<unsigned long long><<<sms * bps2, threads>>>((unsigned long long *)tab, slots, it2); });
      printf("C global atomicAdd u64  slots=%9u  %.3f ms  %.1f Gop/s\n", slots, ms, ops / ms / 1e6);
      ms = timeIt([&] { globalAtomics<float><<<sms * bps2, threads>>>((float *)tab, slots, it2); });
      printf("C global atomicAdd f32  slots=%9u  %.3f ms  %.1f Gop/s\n", slots, ms, ops / ms / 1e6);
      ms = timeIt([&] { globalAtomics<double><<<sms * bps2, threads>>>((double *)tab, slots, it2); });
      printf("C global atomicAdd f64  slots=%9u  %.3f ms  %.1f Gop/s\n", slots, ms, ops / ms / 1e6);
      ms = timeIt([&] { globalProbeAdd<double><<<sms * bps2, threads>>>(keys, (double *)tab, slots, it2, sink); });
      printf("D global key-load + atomicAdd f64  slots=%9u  %.3f ms  %.1f Gop/s\n", slots, ms, ops / ms / 1e6);
    }
    cudaFree(tab); cudaFree(keys);
  }
  {
    float ms = timeIt([&] { matchAny<<<sms * 4, 256>>>(2048, sink); });
    printf("E match_any.sync  %.3f ms  %.1f G lane-ops/s\n", ms, (double)sms * 4 * 256 * 2048 / ms / 1e6);
  }
  return 0;
}
