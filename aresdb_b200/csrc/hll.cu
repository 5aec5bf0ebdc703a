// hll.cu — HyperLogLog entry point (reference: query/hll.cu:21-290; iterators
// query/iterator.hpp:1168-1258; functors query/functor.hpp:1299-1374).
//
// One call per batch, same five steps and same observable buffers as the reference:
//   1. key[i] = (murmur3_x64_128(dim row).lo & ~0xFFFF) | regID(value[i]); stable radix sort of
//      (key, (index, value))                                  [replaces transform + stable_sort_by_key]
//   2. per run of equal keys keep the first index and the MAX value (max rho), appended after the
//      carried rows                                            [replaces reduce_by_key]
//   3. stable merge of carried and new rows by (key asc, value desc) via per-element binary
//      search (merge path)                                     [replaces merge_by_key]
//   4. last batch: dimension heads (key >> 16 changes), register heads (key changes), per-dim
//      register counts, sparse (< 4096 regs: 4 B each, (rho+1) << 16 | reg) or dense (16384 B,
//      rho+1 per register) vector offsets, scatter            [replaces 3 scans + transform_if]
//   5. gather the dim columns of the surviving rows            [replaces the byte-granular copy]
// The two output vectors are allocated with libmem's deviceMalloc and adopted by the caller,
// exactly like the reference (query/time_series_aggregate.go:661-681).
#include "agg.cuh"
#include "dimrow.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace aresb {

int reduceByHash(const uint64_t *hash, const uint32_t *index, const uint8_t *measures, int width, AggOp op, int n,
                 uint32_t *outIndex, uint8_t *outValues, cudaStream_t s, uint64_t *outHash = nullptr);
void gatherDims(const uint8_t *in, const DimLayout &Lin, const uint32_t *rows, int g, uint8_t *out,
                const DimLayout &Lout, cudaStream_t s);

__global__ void __launch_bounds__(256)
hllKeysKernel(const uint8_t *__restrict__ block, DimLayout L, const uint32_t *__restrict__ index,
              const uint32_t *__restrict__ values, int n, uint64_t *__restrict__ keys, uint64_t *__restrict__ payload) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += stride) {
    uint64_t w[4];
    const uint32_t idx = index[i], v = values[i];
    packRow(block, L, idx, w);
    keys[i] = (murmur3_128_lo(w, L.rowBytes, 0) & 0xFFFFFFFFFFFF0000ull) | (v & 0x3FFFu);
    payload[i] = ((uint64_t)v << 32) | idx;
  }
}

__global__ void unpackPayloadKernel(const uint64_t *__restrict__ payload, int n, uint32_t *__restrict__ idx,
                                    uint32_t *__restrict__ vals, uint32_t *__restrict__ pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    idx[i] = (uint32_t)payload[i];
    vals[i] = (uint32_t)(payload[i] >> 32);
    pos[i] = (uint32_t)i;
  }
}

// run s of the sorted batch -> carried arrays at prevResultSize + s
__global__ void appendRunsKernel(const uint32_t *__restrict__ runStart, const uint32_t *__restrict__ runMax,
                                 const uint64_t *__restrict__ runKey, const uint32_t *__restrict__ sortedIdx, int m,
                                 uint64_t *__restrict__ outHash, uint32_t *__restrict__ outIdx, uint32_t *__restrict__ outVal) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < m) {
    outHash[s] = runKey[s];
    outIdx[s] = sortedIdx[runStart[s]];
    outVal[s] = runMax[s];
  }
}

// comp(x, y): x strictly before y in (key asc, value desc) order (HLLMergeComparator)
__device__ __forceinline__ bool hllBefore(uint64_t hx, uint32_t vx, uint64_t hy, uint32_t vy) {
  return hx == hy ? vx > vy : hx < hy;
}

// Stable merge of A = [0, p) and B = [p, p + m) of (hash, value, index); ties keep A first.
__global__ void __launch_bounds__(256)
hllMergeKernel(const uint64_t *__restrict__ hash, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ idx,
               int p, int m, uint64_t *__restrict__ outHash, uint32_t *__restrict__ outVals, uint32_t *__restrict__ outIdx) {
  const int total = p + m;
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const uint64_t h = hash[e];
    const uint32_t v = vals[e];
    int lo, hi, dst;
    if (e < p) {  // element of A: count B elements strictly before it
      lo = 0; hi = m;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (hllBefore(hash[p + mid], vals[p + mid], h, v)) lo = mid + 1; else hi = mid;
      }
      dst = e + lo;
    } else {      // element of B: count A elements not after it (A wins ties)
      lo = 0; hi = p;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (!hllBefore(h, v, hash[mid], vals[mid])) lo = mid + 1; else hi = mid;
      }
      dst = (e - p) + lo;
    }
    outHash[dst] = h;
    outVals[dst] = v;
    outIdx[dst] = idx[e];
  }
}

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// out[r] = number of heads in [0, r] where head(r) = r == 0 || (hash[r] >> shift) != (hash[r-1] >> shift)
__global__ void __launch_bounds__(kScanThreads)
headCountScanKernel(const uint64_t *__restrict__ hash, int n, int shift, ScanTileState st, uint32_t *__restrict__ out) {
  __shared__ uint32_t sTile, sPrefix;
  __shared__ uint32_t sWarp[kScanThreads / 32 + 1];
  if (threadIdx.x == 0) sTile = atomicAdd(st.ticket, 1u);
  __syncthreads();
  const uint32_t tile = sTile;
  const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
  uint32_t flags = 0;
  uint64_t prev = (base > 0 && base <= (uint32_t)n) ? hash[base - 1] >> shift : 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    uint32_t i = base + k;
    if (i < (uint32_t)n) {
      uint64_t h = hash[i] >> shift;
      if (i == 0 || h != prev) flags |= 1u << k;
      prev = h;
    }
  }
  uint32_t blockTotal;
  const uint32_t excl = blockExclusiveScan<kScanThreads>(__popc(flags), sWarp, &blockTotal);
  if (threadIdx.x < 32) {
    uint32_t p = decoupledLookback(st, tile, blockTotal);
    if (threadIdx.x == 0) sPrefix = p;
  }
  __syncthreads();
  uint32_t run = sPrefix + excl;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    uint32_t i = base + k;
    if (i < (uint32_t)n) {
      run += (flags >> k) & 1;
      out[i] = run;
    }
  }
}

// For every dim head r: firstRow[dim] = index[r], regBefore[dim] = regCum[r] - 1 (registers of earlier dims).
__global__ void dimHeadsKernel(const uint32_t *__restrict__ dimOrd, const uint32_t *__restrict__ regCum,
                               const uint32_t *__restrict__ index, int n, uint32_t *__restrict__ firstRow,
                               uint32_t *__restrict__ regBefore) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (r == 0 || dimOrd[r] != dimOrd[r - 1]) {
    uint32_t d = dimOrd[r] - 1;
    firstRow[d] = index[r];
    regBefore[d] = regCum[r] - 1;
  }
}

// regCount[d] (u16) and byte offsets (exclusive scan of the per-dim vector sizes), one block.
__global__ void __launch_bounds__(1024)
dimSizesKernel(const uint32_t *__restrict__ regBefore, uint32_t dims, uint32_t totalRegs, uint16_t *__restrict__ regCount,
               unsigned long long *__restrict__ offsets) {
  __shared__ unsigned long long sSum[1024];
  const uint32_t per = (dims + 1023) / 1024;
  const uint32_t begin = threadIdx.x * per;
  uint32_t end = begin + per;
  if (end > dims) end = dims;
  unsigned long long sum = 0;
  for (uint32_t d = begin; d < end; d++) {
    uint32_t c = (d + 1 < dims ? regBefore[d + 1] : totalRegs) - regBefore[d];
    regCount[d] = (uint16_t)c;
    sum += c < HLL_DENSE_THRESHOLD ? (unsigned long long)c * 4 : HLL_DENSE_SIZE;
  }
  sSum[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; t++) { unsigned long long c = sSum[t]; sSum[t] = run; run += c; }
    offsets[dims] = run;
  }
  __syncthreads();
  unsigned long long run = sSum[threadIdx.x];
  for (uint32_t d = begin; d < end; d++) {
    uint32_t c = (d + 1 < dims ? regBefore[d + 1] : totalRegs) - regBefore[d];
    offsets[d] = run;
    run += c < HLL_DENSE_THRESHOLD ? (unsigned long long)c * 4 : HLL_DENSE_SIZE;
  }
}

__global__ void __launch_bounds__(256)
hllScatterKernel(const uint32_t *__restrict__ dimOrd, const uint32_t *__restrict__ regCum, const uint32_t *__restrict__ vals,
                 int n, const uint32_t *__restrict__ regBefore, const uint16_t *__restrict__ regCount,
                 const unsigned long long *__restrict__ offsets, uint8_t *__restrict__ hll) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (r != 0 && regCum[r] == regCum[r - 1]) return;  // not a register head: a smaller rho of the same register
  const uint32_t d = dimOrd[r] - 1, v = vals[r];
  const uint32_t reg = v & 0x3FFFu;
  const uint8_t rho = (uint8_t)(((v >> 16) & 0xFF) + 1);
  if (regCount[d] < HLL_DENSE_THRESHOLD) {
    const uint32_t within = regCum[r] - 1 - regBefore[d];
    *reinterpret_cast<uint32_t *>(hll + offsets[d] + (size_t)within * 4) = ((uint32_t)rho << 16) | reg;
  } else {
    hll[offsets[d] + reg] = rho;
  }
}

static int gridOf(int64_t n, int per = 256) {
  int64_t b = (n + per - 1) / per;
  int64_t cap = (int64_t)smCount() * 16;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

static void headCountScan(const uint64_t *hash, int n, int shift, uint32_t *out, cudaStream_t s) {
  const int tiles = divUp(n, kScanTile);
  Scratch state(scanStateBytes(tiles), s);
  ARES_CUDA(cudaMemsetAsync(state.ptr, 0, state.bytes, s));
  headCountScanKernel<<<tiles, kScanThreads, 0, s>>>(hash, n, shift, makeScanState(state.ptr, tiles), out);
  checkLastError("headCountScan");
}

// Step 4 on merged rows sorted by (key asc, value desc): per-dim register counts, sparse / dense
// vector offsets and the scatter.  `index[r]` is the dim-row index of entry r; on return
// index[0..dims) holds the row of each dim's first entry.  Allocates the two outputs with
// deviceMalloc (the caller adopts them) and returns the number of dims.
int hllRegisterVectors(const uint64_t *hash, const uint32_t *values, uint32_t *index, int R, uint8_t **hllVectorPtr,
                       size_t *hllVectorSizePtr, uint16_t **hllDimRegIDCountPtr, cudaStream_t s) {
  Scratch dimOrd(sizeof(uint32_t) * (size_t)R, s), regCum(sizeof(uint32_t) * (size_t)R, s);
  headCountScan(hash, R, 16, dimOrd.as<uint32_t>(), s);
  headCountScan(hash, R, 0, regCum.as<uint32_t>(), s);
  uint32_t totals[2];
  ARES_CUDA(cudaMemcpyAsync(&totals[0], dimOrd.as<uint32_t>() + (R - 1), 4, cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaMemcpyAsync(&totals[1], regCum.as<uint32_t>() + (R - 1), 4, cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  const uint32_t dims = totals[0], regs = totals[1];
  Scratch firstRow(sizeof(uint32_t) * (size_t)dims, s), regBefore(sizeof(uint32_t) * (size_t)dims, s);
  Scratch offsets(sizeof(unsigned long long) * ((size_t)dims + 1), s);
  dimHeadsKernel<<<divUp(R, 256), 256, 0, s>>>(dimOrd.as<uint32_t>(), regCum.as<uint32_t>(), index, R,
                                               firstRow.as<uint32_t>(), regBefore.as<uint32_t>());
  checkLastError("dimHeads");
  void *regCountDev = nullptr;
  CGoCallResHandle h = deviceMalloc(&regCountDev, sizeof(uint16_t) * (size_t)dims);
  if (h.pStrErr) { std::string msg(h.pStrErr); free((void *)h.pStrErr); throw EngineError(msg); }
  dimSizesKernel<<<1, 1024, 0, s>>>(regBefore.as<uint32_t>(), dims, regs, static_cast<uint16_t *>(regCountDev),
                                    offsets.as<unsigned long long>());
  checkLastError("dimSizes");
  unsigned long long totalBytes = 0;
  ARES_CUDA(cudaMemcpyAsync(&totalBytes, offsets.as<unsigned long long>() + dims, 8, cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  void *hllDev = nullptr;
  h = deviceMalloc(&hllDev, totalBytes ? (size_t)totalBytes : 1);
  if (h.pStrErr) { std::string msg(h.pStrErr); free((void *)h.pStrErr); deviceFree(regCountDev); throw EngineError(msg); }
  ARES_CUDA(cudaMemsetAsync(hllDev, 0, (size_t)totalBytes, s));
  hllScatterKernel<<<divUp(R, 256), 256, 0, s>>>(dimOrd.as<uint32_t>(), regCum.as<uint32_t>(), values, R,
                                                 regBefore.as<uint32_t>(), static_cast<uint16_t *>(regCountDev),
                                                 offsets.as<unsigned long long>(), static_cast<uint8_t *>(hllDev));
  checkLastError("hllScatter");
  // surviving rows: the first row of every dim, in order
  ARES_CUDA(cudaMemcpyAsync(index, firstRow.ptr, sizeof(uint32_t) * (size_t)dims, cudaMemcpyDeviceToDevice, s));
  *hllVectorPtr = static_cast<uint8_t *>(hllDev);
  *hllVectorSizePtr = (size_t)totalBytes;
  *hllDimRegIDCountPtr = static_cast<uint16_t *>(regCountDev);
  return (int)dims;
}

static int64_t hyperloglog(DimensionVector prev, DimensionVector cur, uint32_t *prevValues, uint32_t *curValues,
                           int prevResultSize, int curBatchSize, bool isLastBatch, uint8_t **hllVectorPtr,
                           size_t *hllVectorSizePtr, uint16_t **hllDimRegIDCountPtr, cudaStream_t s) {
  const int P = prevResultSize, n = curBatchSize;
  // the batch's dim values live in prev.DimValues and are addressed through cur.IndexVector
  DimLayout L = makeDimLayout(cur.NumDimsPerDimWidth, cur.VectorCapacity);
  int m = 0;
  if (n > 0) {
    // 1. keys + stable sort
    Scratch payload(sizeof(uint64_t) * (size_t)n, s), tmpK(sizeof(uint64_t) * (size_t)n, s), tmpV(sizeof(uint64_t) * (size_t)n, s);
    hllKeysKernel<<<gridOf(n), 256, 0, s>>>(prev.DimValues, L, cur.IndexVector, curValues, n, cur.HashValues, payload.as<uint64_t>());
    checkLastError("hllKeys");
    radixSortPairs<uint64_t>(cur.HashValues, payload.as<uint64_t>(), tmpK.as<uint64_t>(), tmpV.as<uint64_t>(), n, 0, 64, s);
    Scratch pos(sizeof(uint32_t) * (size_t)n, s);
    unpackPayloadKernel<<<divUp(n, 256), 256, 0, s>>>(payload.as<uint64_t>(), n, cur.IndexVector, curValues, pos.as<uint32_t>());
    checkLastError("unpackPayload");
    // 2. per key: first index, max value -> appended after the carried rows
    Scratch runStart(sizeof(uint32_t) * (size_t)n, s), runMax(sizeof(uint32_t) * (size_t)n, s), runKey(sizeof(uint64_t) * (size_t)n, s);
    m = reduceByHash(cur.HashValues, pos.as<uint32_t>(), reinterpret_cast<const uint8_t *>(curValues), 4, OP_MAX_U32, n,
                     runStart.as<uint32_t>(), runMax.as<uint8_t>(), s, runKey.as<uint64_t>());
    appendRunsKernel<<<divUp(m, 256), 256, 0, s>>>(runStart.as<uint32_t>(), runMax.as<uint32_t>(), runKey.as<uint64_t>(),
                                                   cur.IndexVector, m, prev.HashValues + P, prev.IndexVector + P, prevValues + P);
    checkLastError("appendRuns");
  }
  // 3. merge carried [0, P) with new [P, P + m) into cur.*
  int resSize = P + m;
  if (resSize > 0) {
    hllMergeKernel<<<gridOf(resSize), 256, 0, s>>>(prev.HashValues, prevValues, prev.IndexVector, P, m, cur.HashValues,
                                                   curValues, cur.IndexVector);
    checkLastError("hllMerge");
  }
  if (isLastBatch && resSize > 0)  // 4. register vectors
    resSize = hllRegisterVectors(cur.HashValues, curValues, cur.IndexVector, resSize, hllVectorPtr, hllVectorSizePtr,
                                 hllDimRegIDCountPtr, s);
  // 5. dims of the surviving rows (output block uses prev's capacity, query/hll.cu:169-187)
  DimLayout Lc = makeDimLayout(prev.NumDimsPerDimWidth, prev.VectorCapacity);
  gatherDims(prev.DimValues, Lc, cur.IndexVector, resSize, cur.DimValues, Lc, s);
  ARES_CUDA(cudaStreamSynchronize(s));
  return resSize;
}

}  // namespace aresb

using namespace aresb;

extern "C" CGoCallResHandle HyperLogLog(DimensionVector prevDimOut, DimensionVector curDimOut, uint32_t *prevValuesOut,
                                        uint32_t *curValuesOut, int prevResultSize, int curBatchSize, bool isLastBatch,
                                        uint8_t **hllVectorPtr, size_t *hllVectorSizePtr, uint16_t **hllDimRegIDCountPtr,
                                        void *cudaStream, int device) {
  return guarded("HyperLogLog", device, [&]() -> int64_t {
    return hyperloglog(prevDimOut, curDimOut, prevValuesOut, curValuesOut, prevResultSize, curBatchSize, isLastBatch,
                       hllVectorPtr, hllVectorSizePtr, hllDimRegIDCountPtr, (cudaStream_t)cudaStream);
  });
}
