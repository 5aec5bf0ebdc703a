// dsmem_atomics.cu — design-time measurement for a cluster-distributed slot array (cfg4: 145,541 slots do not fit one
// CTA's shared memory but fit the shared memory of a cluster of 8): throughput of fire-and-forget reductions into the
// shared memory of the CTAs of a thread-block cluster (`red.shared::cluster`), by element type, target spread and cluster
// size, next to the L2 atomics they would replace, and of a mix of both.  Build & run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/dsmem_atomics.cu -o /tmp/dsm && /tmp/dsm
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__device__ __forceinline__ uint32_t mapa(uint32_t laddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
  return r;
}

// mode 0: u32 add, 1: u64 add, 2: f64 add, 3: f32 add.  spread 0: own CTA only (through the cluster window), 1: random CTA
// of the cluster, 2: random OTHER CTA.  globalEvery: every k-th update goes to the L2 array instead (0: never).
template <int MODE>
__global__ void clusterReds(int slots, int iters, int spread, int globalEvery, unsigned long long *gtab, uint32_t gslots,
                            unsigned long long *sink) {
  extern __shared__ __align__(16) unsigned char raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t csize = cluster.num_blocks(), me = cluster.block_rank();
  unsigned long long *tab = reinterpret_cast<unsigned long long *>(raw);
  for (int i = threadIdx.x; i < slots; i += blockDim.x) tab[i] = 0ull;
  cluster.sync();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
  uint32_t s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
  for (int k = 0; k < iters; k++) {
    const uint32_t r = lcg(s);
    const uint32_t slot = r % (uint32_t)slots;
    if (globalEvery && (k % globalEvery) == 0) {
      const uint32_t g = (r * 2654435761u) % gslots;
      if (MODE == 2) atomicAdd(reinterpret_cast<double *>(gtab) + g, 1.0);
      else atomicAdd(gtab + g, 1ull);
      continue;
    }
    uint32_t rank = me;
    if (spread == 1) rank = (r >> 16) % csize;
    if (spread == 2) rank = (me + 1u + (r >> 16) % (csize - 1u)) % csize;
    const uint32_t addr = mapa(base + slot * 8u, rank);
    if (MODE == 0) asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(1u) : "memory");
    if (MODE == 1) asm volatile("red.relaxed.cluster.shared::cluster.add.u64 [%0], %1;" ::"r"(addr), "l"(1ull) : "memory");
    if (MODE == 2) asm volatile("red.relaxed.cluster.shared::cluster.add.f64 [%0], %1;" ::"r"(addr), "d"(1.0) : "memory");
    if (MODE == 3) asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(addr), "f"(1.0f) : "memory");
  }
  cluster.sync();
  if (threadIdx.x == 0 && tab[0] == 123456789ull) atomicAdd(sink, 1ull);
}

template <int MODE>
static int run(const char *name, int csize, int ctas, int slots, int spread, int globalEvery, unsigned long long *gtab, uint32_t gslots,
               unsigned long long *sink) {
  const int threads = 1024, iters = 2048;
  const size_t smem = (size_t)slots * 8;
  auto kern = clusterReds<MODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (csize > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int maxClusters = 0;
  cudaError_t oe = cudaOccupancyMaxActiveClusters(&maxClusters, kern, &cfg);
  if (oe != cudaSuccess) { printf("%s cluster=%d: occupancy query failed: %s\n", name, csize, cudaGetErrorString(oe)); cudaGetLastError(); return 0; }
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, slots, 64, spread, globalEvery, gtab, gslots, sink);   // warm-up
  if (le != cudaSuccess) { printf("%s cluster=%d: launch failed: %s\n", name, csize, cudaGetErrorString(le)); cudaGetLastError(); return 0; }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  CK(cudaLaunchKernelEx(&cfg, kern, slots, iters, spread, globalEvery, gtab, gslots, sink));
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  const double ops = (double)ctas * threads * iters;
  printf("%-4s cluster=%2d ctas=%3d (max active clusters %2d) slots/CTA=%6d spread=%d globalEvery=%d  %.3f ms  %.1f Gop/s\n", name, csize,
         ctas, maxClusters, slots, spread, globalEvery, ms, ops / ms * 1e-6);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  unsigned long long *gtab, *sink;
  const uint32_t gslots = 1u << 20;
  CK(cudaMalloc(&gtab, (size_t)gslots * 8)); CK(cudaMemset(gtab, 0, (size_t)gslots * 8));
  CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
  const int slots = 18200;   // 145,541 slots over a cluster of 8
  for (int csize : {1, 2, 4, 8, 16}) {
    const int ctas = (148 / csize) * csize;
    for (int spread : {0, 1, 2}) {
      if (csize == 1 && spread) continue;
      if (run<0>("u32", csize, ctas, slots, spread, 0, gtab, gslots, sink)) return 1;
      if (run<1>("u64", csize, ctas, slots, spread, 0, gtab, gslots, sink)) return 1;
      if (run<2>("f64", csize, ctas, slots, spread, 0, gtab, gslots, sink)) return 1;
    }
  }
  if (run<3>("f32", 8, 144, slots, 1, 0, gtab, gslots, sink)) return 1;
  // a mix: every k-th update goes to the L2 array (1M slots) instead
  for (int every : {2, 3, 4}) {
    if (run<1>("u64", 8, 144, slots, 1, every, gtab, gslots, sink)) return 1;
    if (run<2>("f64", 8, 144, slots, 1, every, gtab, gslots, sink)) return 1;
  }
  // all global, same loop shape (reference point)
  if (run<1>("u64", 1, 148, slots, 0, 1, gtab, gslots, sink)) return 1;
  if (run<2>("f64", 1, 148, slots, 0, 1, gtab, gslots, sink)) return 1;
  return 0;
}
