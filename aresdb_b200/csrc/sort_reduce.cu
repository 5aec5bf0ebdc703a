// sort_reduce.cu — Sort and Reduce entry points (reference: query/sort_reduce.cu:27-80, 118-249).
//
//   Sort    hashRows (murmur3_x64_128 low word of each packed dim row) + stable LSD radix sort of
//           (hash u64, index u32) — replaces thrust::copy(DimensionHashIterator) +
//           thrust::stable_sort_by_key.
//   Reduce  segmentHeads (single-pass decoupled look-back compaction of run starts) +
//           segmentReduce (one warp per run, fixed shuffle tree; very long runs go to one block
//           each) + gatherDims — replaces thrust::reduce_by_key + the byte-granular dim copy.
// Runs are delimited by equal HASHES (not equal rows), and the first row of a run supplies the
// dimension values, exactly as the reference does.
#include "agg.cuh"
#include "dimrow.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace aresb {

__global__ void __launch_bounds__(256)
hashRowsKernel(const uint8_t *__restrict__ block, DimLayout L, const uint32_t *__restrict__ index, int n,
               uint64_t *__restrict__ hashOut, uint64_t mask) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += stride) {
    uint64_t w[4];
    packRow(block, L, index[i], w);
    hashOut[i] = murmur3_128_lo(w, L.rowBytes, 0) & mask;
  }
}

constexpr int kHeadThreads = 256;
constexpr int kHeadItems = 8;
constexpr int kHeadTile = kHeadThreads * kHeadItems;

// segStart[s] = position of the first element of run s; segStart[g] = n; *outCount = g.
__global__ void __launch_bounds__(kHeadThreads)
segmentHeadsKernel(const uint64_t *__restrict__ hash, int n, ScanTileState st, uint32_t *__restrict__ segStart,
                   uint32_t *__restrict__ outCount) {
  __shared__ uint32_t sTile, sPrefix;
  __shared__ uint32_t sWarp[kHeadThreads / 32 + 1];
  if (threadIdx.x == 0) sTile = atomicAdd(st.ticket, 1u);
  __syncthreads();
  const uint32_t tile = sTile;
  const uint32_t base = tile * kHeadTile + threadIdx.x * kHeadItems;
  uint32_t mask = 0;
  uint64_t prev = (base > 0 && base <= (uint32_t)n) ? hash[base - 1] : 0;
#pragma unroll
  for (int k = 0; k < kHeadItems; k++) {
    uint32_t i = base + k;
    if (i < (uint32_t)n) {
      uint64_t h = hash[i];
      if (i == 0 || h != prev) mask |= 1u << k;
      prev = h;
    }
  }
  uint32_t blockTotal;
  const uint32_t excl = blockExclusiveScan<kHeadThreads>(__popc(mask), sWarp, &blockTotal);
  if (threadIdx.x < 32) {
    uint32_t p = decoupledLookback(st, tile, blockTotal);
    if (threadIdx.x == 0) {
      sPrefix = p;
      if ((uint64_t)(tile + 1) * kHeadTile >= (uint64_t)n) {
        *outCount = p + blockTotal;
        segStart[p + blockTotal] = (uint32_t)n;
      }
    }
  }
  __syncthreads();
  uint32_t pos = sPrefix + excl;
#pragma unroll
  for (int k = 0; k < kHeadItems; k++)
    if (mask & (1u << k)) segStart[pos++] = base + k;
}

constexpr uint32_t kLongRun = 1u << 15;

// One warp per run; lanes stride the run, then a fixed shuffle tree.  Runs longer than
// kLongRun are queued for longRunKernel.
__global__ void __launch_bounds__(256)
segmentReduceKernel(const uint32_t *__restrict__ index, const uint8_t *__restrict__ measures, int width, AggOp op,
                    const uint32_t *__restrict__ segStart, uint32_t g, uint32_t *__restrict__ outIndex,
                    uint8_t *__restrict__ outValues, uint32_t *__restrict__ longList, uint32_t *__restrict__ longCount,
                    const uint64_t *__restrict__ hash, uint64_t *__restrict__ outHash,
                    const uint32_t *__restrict__ runList = nullptr, const uint32_t *__restrict__ runCount = nullptr) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  if (runList != nullptr) g = *runCount;   // only the runs segmentReduceShortKernel left over
  for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < g; q += warpsPerGrid) {
    const uint32_t s = runList != nullptr ? runList[q] : q;
    const uint32_t begin = segStart[s], end = segStart[s + 1];
    if (outHash != nullptr && lane == 0) outHash[s] = hash[begin];
    if (end - begin > kLongRun) {
      if (lane == 0) longList[atomicAdd(longCount, 1u)] = s;
      continue;
    }
    uint64_t acc = 0;
    bool has = false;
    for (uint32_t j = begin + lane; j < end; j += 32) {
      uint64_t v = loadMeasure(measures, index[j], width);
      acc = has ? aggCombine(op, acc, v) : v;
      has = true;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {  // lane l absorbs lane l+d: keeps run order left-to-right
      uint64_t other = __shfl_down_sync(0xffffffffu, acc, d);
      bool otherHas = __shfl_down_sync(0xffffffffu, has ? 1 : 0, d) != 0;
      if ((lane & (2 * d - 1)) == 0 && lane + d < 32 && otherHas) {
        acc = has ? aggCombine(op, acc, other) : other;
        has = true;
      }
    }
    if (lane == 0) {
      outIndex[s] = index[begin];
      storeMeasure(outValues, s, width, acc);
    }
  }
}

// Mostly-distinct hashes (g close to n, e.g. the 1.16e6 groups of cfg4 at finalize): one THREAD per run.  Runs of one or
// two elements are finished here (for two elements the shuffle tree above is combine(first, second) as well); longer
// runs are listed for the warp kernel.  (One warp per one-element run cost 404 us for 1.16e6 runs.)
__global__ void __launch_bounds__(256)
segmentReduceShortKernel(const uint32_t *__restrict__ index, const uint8_t *__restrict__ measures, int width, AggOp op,
                         const uint32_t *__restrict__ segStart, uint32_t g, uint32_t *__restrict__ outIndex,
                         uint8_t *__restrict__ outValues, uint32_t *__restrict__ runList, uint32_t *__restrict__ runCount,
                         const uint64_t *__restrict__ hash, uint64_t *__restrict__ outHash) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < g; s += stride) {
    const uint32_t begin = segStart[s], end = segStart[s + 1];
    if (end - begin > 2) { runList[atomicAdd(runCount, 1u)] = s; continue; }
    if (outHash != nullptr) outHash[s] = hash[begin];
    const uint32_t first = index[begin];
    uint64_t acc = loadMeasure(measures, first, width);
    if (end - begin == 2) acc = aggCombine(op, acc, loadMeasure(measures, index[begin + 1], width));
    outIndex[s] = first;
    storeMeasure(outValues, s, width, acc);
  }
}

__global__ void __launch_bounds__(512)
longRunKernel(const uint32_t *__restrict__ index, const uint8_t *__restrict__ measures, int width, AggOp op,
              const uint32_t *__restrict__ segStart, const uint32_t *__restrict__ longList,
              const uint32_t *__restrict__ longCount, uint32_t *__restrict__ outIndex, uint8_t *__restrict__ outValues) {
  __shared__ uint64_t sAcc[512];
  __shared__ uint8_t sHas[512];
  for (uint32_t q = blockIdx.x; q < *longCount; q += gridDim.x) {
    const uint32_t s = longList[q];
    const uint32_t begin = segStart[s], end = segStart[s + 1];
    uint64_t acc = 0;
    bool has = false;
    for (uint32_t j = begin + threadIdx.x; j < end; j += blockDim.x) {
      uint64_t v = loadMeasure(measures, index[j], width);
      acc = has ? aggCombine(op, acc, v) : v;
      has = true;
    }
    sAcc[threadIdx.x] = acc;
    sHas[threadIdx.x] = has;
    __syncthreads();
    for (int d = 1; d < 512; d <<= 1) {
      if ((threadIdx.x & (2 * d - 1)) == 0 && threadIdx.x + d < 512 && sHas[threadIdx.x + d]) {
        sAcc[threadIdx.x] = sHas[threadIdx.x] ? aggCombine(op, sAcc[threadIdx.x], sAcc[threadIdx.x + d])
                                              : sAcc[threadIdx.x + d];
        sHas[threadIdx.x] = 1;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      outIndex[s] = index[begin];
      storeMeasure(outValues, s, width, sAcc[0]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
gatherDimsKernel(const uint8_t *__restrict__ in, DimLayout Lin, const uint32_t *__restrict__ rows, uint32_t g,
                 uint8_t *__restrict__ out, DimLayout Lout) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < g; s += stride) copyRow(in, Lin, rows[s], out, Lout, s);
}

static int gridFor(int64_t n, int perBlock, int maxPerSm = 8) {
  int64_t blocks = (n + perBlock - 1) / perBlock;
  int64_t cap = (int64_t)smCount() * maxPerSm;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

void hashRows(const uint8_t *block, const DimLayout &L, const uint32_t *index, int n, uint64_t *hashOut,
              cudaStream_t s) {
  if (n <= 0) return;
  hashRowsKernel<<<gridFor(n, 256), 256, 0, s>>>(block, L, index, n, hashOut, testHash64Mask());
  checkLastError("hashRows");
}

void gatherDims(const uint8_t *in, const DimLayout &Lin, const uint32_t *rows, int g, uint8_t *out,
                const DimLayout &Lout, cudaStream_t s) {
  if (g <= 0) return;
  gatherDimsKernel<<<gridFor(g, 256), 256, 0, s>>>(in, Lin, rows, (uint32_t)g, out, Lout);
  checkLastError("gatherDims");
}

// Device-side reduce_by_key over equal consecutive hashes.  Returns the number of runs.
int reduceByHash(const uint64_t *hash, const uint32_t *index, const uint8_t *measures, int width, AggOp op,
                 int n, uint32_t *outIndex, uint8_t *outValues, cudaStream_t s, uint64_t *outHash = nullptr) {
  if (n <= 0) return 0;
  const int tiles = divUp(n, kHeadTile);
  Scratch state(scanStateBytes(tiles) + 4 * sizeof(uint32_t), s);
  ARES_CUDA(cudaMemsetAsync(state.ptr, 0, state.bytes, s));
  ScanTileState st = makeScanState(state.ptr, tiles);
  uint32_t *dCount = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(state.ptr) + scanStateBytes(tiles));
  uint32_t *dLongCount = dCount + 1;
  Scratch segStart(sizeof(uint32_t) * ((size_t)n + 1), s);
  segmentHeadsKernel<<<tiles, kHeadThreads, 0, s>>>(hash, n, st, segStart.as<uint32_t>(), dCount);
  checkLastError("segmentHeads");
  uint32_t g = 0;
  ARES_CUDA(cudaMemcpyAsync(&g, dCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  const uint32_t maxLong = (uint32_t)(n / kLongRun) + 1;
  Scratch longList(sizeof(uint32_t) * maxLong, s);
  if ((int64_t)g * 4 > n) {   // average run shorter than 4: thread per run, the few longer runs go to the warp kernel
    uint32_t *dRunCount = dCount + 2;
    Scratch runList(sizeof(uint32_t) * ((size_t)n / 3 + 1), s);   // runs of >= 3 elements
    segmentReduceShortKernel<<<gridFor(g, 256), 256, 0, s>>>(index, measures, width, op, segStart.as<uint32_t>(), g, outIndex,
                                                            outValues, runList.as<uint32_t>(), dRunCount, hash, outHash);
    checkLastError("segmentReduceShort");
    segmentReduceKernel<<<gridFor((int64_t)(n / 3 + 1) * 32, 256, 2), 256, 0, s>>>(index, measures, width, op, segStart.as<uint32_t>(), g,
                                                                                  outIndex, outValues, longList.as<uint32_t>(), dLongCount,
                                                                                  hash, outHash, runList.as<uint32_t>(), dRunCount);
    checkLastError("segmentReduce");
  } else {
    segmentReduceKernel<<<gridFor((int64_t)g * 32, 256), 256, 0, s>>>(index, measures, width, op, segStart.as<uint32_t>(), g,
                                                                     outIndex, outValues, longList.as<uint32_t>(), dLongCount,
                                                                     hash, outHash);
    checkLastError("segmentReduce");
  }
  if ((uint32_t)n > kLongRun) {
    int blocks = (int)(maxLong < (uint32_t)smCount() * 2 ? maxLong : (uint32_t)smCount() * 2);
    longRunKernel<<<blocks, 512, 0, s>>>(index, measures, width, op, segStart.as<uint32_t>(), longList.as<uint32_t>(),
                                         dLongCount, outIndex, outValues);
    checkLastError("longRun");
  }
  return (int)g;
}

}  // namespace aresb

using namespace aresb;

extern "C" {

CGoCallResHandle Sort(DimensionVector keys, int length, void *cudaStream, int device) {
  return guarded("Sort", device, [&]() -> int64_t {
    if (length <= 0) return 0;
    cudaStream_t s = (cudaStream_t)cudaStream;
    DimLayout L = makeDimLayout(keys.NumDimsPerDimWidth, keys.VectorCapacity);
    hashRows(keys.DimValues, L, keys.IndexVector, length, keys.HashValues, s);
    Scratch tmpK(sizeof(uint64_t) * (size_t)length, s), tmpV(sizeof(uint32_t) * (size_t)length, s);
    radixSortPairs<uint32_t>(keys.HashValues, keys.IndexVector, tmpK.as<uint64_t>(), tmpV.as<uint32_t>(), length, 0, 64, s);
    return 0;
  });
}

CGoCallResHandle Reduce(DimensionVector inputKeys, uint8_t *inputValues, DimensionVector outputKeys,
                        uint8_t *outputValues, int valueBytes, int length, enum AggregateFunction aggFunc,
                        void *cudaStream, int device) {
  return guarded("Reduce", device, [&]() -> int64_t {
    if (length <= 0) return 0;
    cudaStream_t s = (cudaStream_t)cudaStream;
    int width;
    AggOp op = aggOpOf(aggFunc, valueBytes, &width);
    int g = reduceByHash(inputKeys.HashValues, inputKeys.IndexVector, inputValues, width, op, length,
                         outputKeys.IndexVector, outputValues, s);
    // output block uses the INPUT capacity (reference query/sort_reduce.cu:231-236)
    DimLayout L = makeDimLayout(inputKeys.NumDimsPerDimWidth, inputKeys.VectorCapacity);
    gatherDims(inputKeys.DimValues, L, outputKeys.IndexVector, g, outputKeys.DimValues, L, s);
    ARES_CUDA(cudaStreamSynchronize(s));
    return g;
  });
}

}  // extern "C"
