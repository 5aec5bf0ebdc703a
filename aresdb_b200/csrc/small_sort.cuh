// small_sort.cuh — one-CTA sort of up to kSmallSortMax (key, index) pairs with pairwise distinct indices, as a device
// function so that the single-launch finalize (batch_plan.cu) can run it between its gather and merge phases.
// Bucket by the top kSmallBits key bits (unordered scatter into the tmp arrays), then every element finds its rank
// inside its bucket by comparing (key, index) with the bucket's other members.  With hash keys and 4096 buckets a
// bucket of the largest input (32768 keys) holds 8 elements on average, so the rank loop is a handful of L1-resident
// loads; a degenerate bucket only costs time, never correctness.
#pragma once
#include "scan.cuh"

namespace aresb {

constexpr int kSmallBits = 12;
constexpr int kSmallBuckets = 1 << kSmallBits;

// cnt[kSmallBuckets], off[kSmallBuckets + 1], sWarp[33]: shared memory of the calling CTA (1024 threads).
__device__ __forceinline__ void smallSortBody(uint64_t *__restrict__ keys, uint32_t *__restrict__ index, uint64_t *__restrict__ tmpK,
                                              uint32_t *__restrict__ tmpI, int n, int shift, uint32_t *cnt, uint32_t *off,
                                              uint32_t *sWarp) {
  constexpr int kPer = kSmallBuckets / 1024;   // buckets per thread in the scan
  for (int b = threadIdx.x; b < kSmallBuckets; b += 1024) cnt[b] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 1024) atomicAdd(&cnt[(uint32_t)(keys[i] >> shift) & (kSmallBuckets - 1)], 1u);
  __syncthreads();
  {
    uint32_t c[kPer], sum = 0, total;
#pragma unroll
    for (int j = 0; j < kPer; j++) { c[j] = cnt[threadIdx.x * kPer + j]; sum += c[j]; }
    uint32_t excl = blockExclusiveScan<1024>(sum, sWarp, &total);
#pragma unroll
    for (int j = 0; j < kPer; j++) { off[threadIdx.x * kPer + j] = excl; excl += c[j]; }
    if (threadIdx.x == 0) off[kSmallBuckets] = total;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSmallBuckets; b += 1024) cnt[b] = 0;   // reused as the scatter cursors
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 1024) {
    const uint64_t k = keys[i];
    const uint32_t d = (uint32_t)(k >> shift) & (kSmallBuckets - 1);
    const uint32_t p = off[d] + atomicAdd(&cnt[d], 1u);
    tmpK[p] = k;
    tmpI[p] = index[i];
  }
  __syncthreads();   // the CTA's own global writes are visible to it after the barrier
  for (int i = threadIdx.x; i < n; i += 1024) {
    const uint64_t k = tmpK[i];
    const uint32_t v = tmpI[i];
    const uint32_t d = (uint32_t)(k >> shift) & (kSmallBuckets - 1);
    const uint32_t lo = off[d], hi = off[d + 1];
    uint32_t rank = 0;
    for (uint32_t j = lo; j < hi; j++) {
      const uint64_t kj = tmpK[j];
      rank += (kj < k) || (kj == k && tmpI[j] < v);
    }
    keys[lo + rank] = k;
    index[lo + rank] = v;
  }
  __syncthreads();
}

}  // namespace aresb
