// radix_sort.cu — see radix_sort.cuh.
#include "radix_sort.cuh"
#include "scan.cuh"
#include "small_sort.cuh"

namespace aresb {

struct SortGrid {
  int blocks;
  int chunk;  // elements per block, multiple of kSortTile
};

static SortGrid planGrid(int n) {
  SortGrid g;
  int tiles = divUp(n, kSortTile);
  int maxBlocks = smCount() * 4;
  int blocks = tiles < maxBlocks ? tiles : maxBlocks;
  if (blocks < 1) blocks = 1;
  int tilesPerBlock = divUp(tiles, blocks);
  g.chunk = tilesPerBlock * kSortTile;
  g.blocks = divUp(n, g.chunk);
  if (g.blocks < 1) g.blocks = 1;
  return g;
}

size_t radixSortScratchBytes(int n) {
  SortGrid g = planGrid(n);
  return sizeof(uint32_t) * (size_t)kRadix * g.blocks;
}

__global__ void __launch_bounds__(kSortThreads)
digitHistogram(const uint64_t *__restrict__ keys, int n, int shift, int chunk, uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[kRadix];
  sh[threadIdx.x] = 0;  // kSortThreads == kRadix
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * chunk;
  int64_t end = begin + chunk;
  if (end > n) end = n;
  for (int64_t i = begin + threadIdx.x; i < end; i += kSortThreads) {
    uint32_t d = (uint32_t)(keys[i] >> shift) & (kRadix - 1);
    atomicAdd(&sh[d], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = sh[threadIdx.x];
}

// Exclusive scan, in place, of the digit-major kRadix x G counter table (one block).  A warp owns digits w, w + 32, ...:
// it first sums each of its digits over the G blocks (coalesced rows), the block scans the kRadix digit totals, and the
// warp then walks its rows again turning counts into exclusive prefixes (32 blocks per step, warp scan + carry).
// (The first version gave every thread a contiguous slice of the flattened table: uncoalesced, 137 us for 1.5e5 counters.)
__global__ void __launch_bounds__(1024) scanHistograms(uint32_t *hist, int G) {
  __shared__ uint32_t sWarp[1024 / 32 + 1];
  __shared__ uint32_t digitBase[kRadix];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = warp; d < kRadix; d += 32) {
    const uint32_t *row = hist + (size_t)d * G;
    uint32_t s = 0;
    for (int b = lane; b < G; b += 32) s += row[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) digitBase[d] = s;
  }
  __syncthreads();
  {
    uint32_t total;
    const uint32_t v = threadIdx.x < kRadix ? digitBase[threadIdx.x] : 0u;
    const uint32_t excl = blockExclusiveScan<1024>(v, sWarp, &total);
    __syncthreads();
    if (threadIdx.x < kRadix) digitBase[threadIdx.x] = excl;
  }
  __syncthreads();
  for (int d = warp; d < kRadix; d += 32) {
    uint32_t *row = hist + (size_t)d * G;
    uint32_t carry = digitBase[d];
    for (int b0 = 0; b0 < G; b0 += 32) {
      const int b = b0 + lane;
      const uint32_t c = b < G ? row[b] : 0u;
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
      }
      if (b < G) row[b] = carry + incl - c;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

template <typename V>
__global__ void __launch_bounds__(kSortThreads)
scatterByDigit(const uint64_t *__restrict__ keysIn, const V *__restrict__ valsIn, uint64_t *__restrict__ keysOut,
               V *__restrict__ valsOut, int n, int shift, int chunk, const uint32_t *__restrict__ histScanned) {
  constexpr int W = kSortThreads / 32;
  __shared__ uint64_t sKeys[kSortTile];
  __shared__ V sVals[kSortTile];
  __shared__ uint32_t warpCount[W][kRadix];
  __shared__ uint32_t digitBase[kRadix];
  __shared__ uint32_t tileStart[kRadix];
  __shared__ uint32_t tileCount[kRadix];
  __shared__ uint32_t sScan[W + 1];

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ltMask = (1u << lane) - 1u;
  digitBase[threadIdx.x] = histScanned[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
  const int64_t begin = (int64_t)blockIdx.x * chunk;
  int64_t end = begin + chunk;
  if (end > n) end = n;

  for (int64_t tileBase = begin; tileBase < end; tileBase += kSortTile) {
#pragma unroll
    for (int w = 0; w < W; w++) warpCount[w][threadIdx.x] = 0;
    __syncthreads();

    uint64_t key[kSortItems];
    V val[kSortItems];
    uint32_t rank[kSortItems];
    const int64_t warpBase = tileBase + (int64_t)warp * 32 * kSortItems;
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
      int64_t p = warpBase + j * 32 + lane;
      bool ok = p < end;
      key[j] = ok ? keysIn[p] : 0;
      val[j] = ok ? valsIn[p] : V(0);
    }
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
      int64_t p = warpBase + j * 32 + lane;
      bool ok = p < end;
      uint32_t d = ok ? ((uint32_t)(key[j] >> shift) & (kRadix - 1)) : 0xFFFFFFFFu;
      uint32_t peers = __match_any_sync(0xffffffffu, d);
      uint32_t leader = __ffs(peers) - 1;
      uint32_t prev = 0;
      if (ok && lane == leader) {
        prev = warpCount[warp][d];
        warpCount[warp][d] = prev + __popc(peers);
      }
      prev = __shfl_sync(0xffffffffu, prev, leader);
      rank[j] = prev + __popc(peers & ltMask);
      __syncwarp();
    }
    __syncthreads();
    {  // cross-warp exclusive prefix per digit, then exclusive scan over digits
      uint32_t running = 0;
#pragma unroll
      for (int w = 0; w < W; w++) {
        uint32_t c = warpCount[w][threadIdx.x];
        warpCount[w][threadIdx.x] = running;
        running += c;
      }
      tileCount[threadIdx.x] = running;
      uint32_t total;
      tileStart[threadIdx.x] = blockExclusiveScan<kSortThreads>(running, sScan, &total);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
      int64_t p = warpBase + j * 32 + lane;
      if (p < end) {
        uint32_t d = (uint32_t)(key[j] >> shift) & (kRadix - 1);
        uint32_t pos = tileStart[d] + warpCount[warp][d] + rank[j];
        sKeys[pos] = key[j];
        sVals[pos] = val[j];
      }
    }
    __syncthreads();
    const int valid = (int)((end - tileBase) < kSortTile ? (end - tileBase) : kSortTile);
    for (int t = threadIdx.x; t < valid; t += kSortThreads) {
      uint64_t k = sKeys[t];
      uint32_t d = (uint32_t)(k >> shift) & (kRadix - 1);
      uint32_t g = digitBase[d] + ((uint32_t)t - tileStart[d]);
      keysOut[g] = k;
      valsOut[g] = sVals[t];
    }
    __syncthreads();
    digitBase[threadIdx.x] += tileCount[threadIdx.x];
    // next iteration's first __syncthreads orders this update before digitBase is read again
  }
}

template <typename V>
void radixSortPairs(uint64_t *keys, V *vals, uint64_t *keysTmp, V *valsTmp, int n, int beginBit, int endBit,
                    cudaStream_t s) {
  static_assert(kSortThreads == kRadix, "one thread per digit");
  if (n <= 1) return;
  SortGrid g = planGrid(n);
  Scratch hist(radixSortScratchBytes(n), s);
  int passes = (endBit - beginBit + kRadixBits - 1) / kRadixBits;
  uint64_t *kin = keys, *kout = keysTmp;
  V *vin = vals, *vout = valsTmp;
  for (int p = 0; p < passes; p++) {
    int shift = beginBit + p * kRadixBits;
    digitHistogram<<<g.blocks, kSortThreads, 0, s>>>(kin, n, shift, g.chunk, hist.as<uint32_t>());
    scanHistograms<<<1, 1024, 0, s>>>(hist.as<uint32_t>(), g.blocks);
    scatterByDigit<V><<<g.blocks, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, shift, g.chunk, hist.as<uint32_t>());
    noteLaunches(2);
    checkLastError("radixSortPairs");
    uint64_t *tk = kin; kin = kout; kout = tk;
    V *tv = vin; vin = vout; vout = tv;
  }
  if (kin != keys) {  // odd number of passes: bring the result home
    ARES_CUDA(cudaMemcpyAsync(keys, kin, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    ARES_CUDA(cudaMemcpyAsync(vals, vin, sizeof(V) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  }
}

__global__ void __launch_bounds__(1024)
smallSortKernel(uint64_t *__restrict__ keys, uint32_t *__restrict__ index, uint64_t *__restrict__ tmpK,
                uint32_t *__restrict__ tmpI, int n, int shift) {
  __shared__ uint32_t cnt[kSmallBuckets], off[kSmallBuckets + 1];
  __shared__ uint32_t sWarp[1024 / 32 + 1];
  smallSortBody(keys, index, tmpK, tmpI, n, shift, cnt, off, sWarp);
}

void sortKeyIndexPairs(uint64_t *keys, uint32_t *index, uint64_t *keysTmp, uint32_t *indexTmp, int n, int keyBits,
                       cudaStream_t s) {
  if (n <= 1) return;
  if (n > kSmallSortMax || keyBits < kSmallBits) {
    radixSortPairs<uint32_t>(keys, index, keysTmp, indexTmp, n, 0, keyBits, s);
    return;
  }
  smallSortKernel<<<1, 1024, 0, s>>>(keys, index, keysTmp, indexTmp, n, keyBits - kSmallBits);
  checkLastError("sortKeyIndexPairs");
}

template void radixSortPairs<uint32_t>(uint64_t *, uint32_t *, uint64_t *, uint32_t *, int, int, int, cudaStream_t);
template void radixSortPairs<uint64_t>(uint64_t *, uint64_t *, uint64_t *, uint64_t *, int, int, int, cudaStream_t);

}  // namespace aresb
