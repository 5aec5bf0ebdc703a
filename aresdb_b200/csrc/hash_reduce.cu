// hash_reduce.cu — HashReduce entry point (reference: query/hash_reduction.cu:183-391, map:
// thirdparty/cudf .../concurrent_unordered_map.cuh:295-355).
//
// Same semantics, own table: group identity is murmur3_x86_32 of the packed dim row (rows whose
// hashes collide are merged, as in the reference), open addressing with linear probing over a
// power-of-two table of >= 2 x length slots.  A slot is one 64-bit word (hash << 32 | row):
// claimed with atomicCAS, then atomicMin keeps the SMALLEST row index as the group's
// representative, which makes the output deterministic (the reference keeps whichever thread
// won the race; its HOST build keeps the first row — we match the HOST build).  Measures are
// folded with native atomics into a parallel accumulator array.  Extraction is a stable
// compaction in table order (decoupled look-back) instead of an atomic cursor.
#include "agg.cuh"
#include "dimrow.cuh"
#include "scan.cuh"

namespace aresb {

constexpr unsigned long long kEmptySlot = ~0ull;

__global__ void __launch_bounds__(256)
fillIdentityKernel(uint8_t *acc, size_t cap, int width, uint64_t identity) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) storeMeasure(acc, i, width, identity);
}

__global__ void __launch_bounds__(256)
hashInsertKernel(const uint8_t *__restrict__ block, DimLayout L, int n, const uint8_t *__restrict__ measures, int width,
                 AggOp op, unsigned long long *__restrict__ slots, uint8_t *__restrict__ acc, uint32_t mask) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += stride) {
    uint64_t w[4];
    packRow(block, L, i, w);
    const uint32_t h = murmur3_32(w, L.rowBytes, 0);
    const unsigned long long key = ((unsigned long long)h << 32) | i;
    uint32_t p = h & mask;
    while (true) {
      unsigned long long cur = slots[p];
      if (cur == kEmptySlot) {
        cur = atomicCAS(&slots[p], kEmptySlot, key);
        if (cur == kEmptySlot) break;
      }
      if ((uint32_t)(cur >> 32) == h) {
        if (key < cur) atomicMin(&slots[p], key);
        break;
      }
      p = (p + 1) & mask;
    }
    aggAtomic(op, acc + (size_t)p * width, loadMeasure(measures, i, width));
  }
}

constexpr int kExtThreads = 256;
constexpr int kExtItems = 8;
constexpr int kExtTile = kExtThreads * kExtItems;

__global__ void __launch_bounds__(kExtThreads)
extractSlotsKernel(const unsigned long long *__restrict__ slots, const uint8_t *__restrict__ acc, int width, size_t cap,
                   ScanTileState st, uint32_t *__restrict__ outRows, uint8_t *__restrict__ outValues,
                   uint32_t *__restrict__ outCount) {
  __shared__ uint32_t sTile, sPrefix;
  __shared__ uint32_t sWarp[kExtThreads / 32 + 1];
  if (threadIdx.x == 0) sTile = atomicAdd(st.ticket, 1u);
  __syncthreads();
  const uint32_t tile = sTile;
  const size_t base = (size_t)tile * kExtTile + (size_t)threadIdx.x * kExtItems;
  uint32_t mask = 0;
  uint32_t rows[kExtItems];
#pragma unroll
  for (int k = 0; k < kExtItems; k++) {
    size_t i = base + k;
    if (i < cap) {
      unsigned long long sl = slots[i];
      if (sl != kEmptySlot) { mask |= 1u << k; rows[k] = (uint32_t)sl; }
    }
  }
  uint32_t blockTotal;
  const uint32_t excl = blockExclusiveScan<kExtThreads>(__popc(mask), sWarp, &blockTotal);
  if (threadIdx.x < 32) {
    uint32_t p = decoupledLookback(st, tile, blockTotal);
    if (threadIdx.x == 0) {
      sPrefix = p;
      if (((size_t)tile + 1) * kExtTile >= cap) *outCount = p + blockTotal;
    }
  }
  __syncthreads();
  uint32_t pos = sPrefix + excl;
#pragma unroll
  for (int k = 0; k < kExtItems; k++) {
    if (mask & (1u << k)) {
      outRows[pos] = rows[k];
      storeMeasure(outValues, pos, width, loadMeasure(acc, base + k, width));
      pos++;
    }
  }
}

void gatherDims(const uint8_t *in, const DimLayout &Lin, const uint32_t *rows, int g, uint8_t *out,
                const DimLayout &Lout, cudaStream_t s);

static uint64_t identityBits(int aggFunc, AggOp op) {
  switch (op) {
    case OP_MIN_U32: case OP_MIN_I32: case OP_MAX_I32: case OP_MAX_U32: return aggIdentity(aggFunc, aggFunc == AGGR_MIN_UNSIGNED || aggFunc == AGGR_MAX_UNSIGNED ? VC_U32 : VC_I32);
    case OP_MIN_F32: case OP_MAX_F32: return aggIdentity(aggFunc, VC_F32);
    default: return 0;
  }
}

}  // namespace aresb

using namespace aresb;

extern "C" CGoCallResHandle HashReduce(DimensionVector inputKeys, uint8_t *inputValues, DimensionVector outputKeys,
                                       uint8_t *outputValues, int valueBytes, int length,
                                       enum AggregateFunction aggFunc, void *cudaStream, int device) {
  return guarded("HashReduce", device, [&]() -> int64_t {
    if (length <= 0) return 0;
    cudaStream_t s = (cudaStream_t)cudaStream;
    int width;
    AggOp op = aggOpOf(aggFunc, valueBytes, &width);
    size_t cap = 64;
    while (cap < 2 * (size_t)length) cap <<= 1;
    DimLayout L = makeDimLayout(inputKeys.NumDimsPerDimWidth, inputKeys.VectorCapacity);
    Scratch slots(sizeof(unsigned long long) * cap, s), acc((size_t)width * cap, s);
    ARES_CUDA(cudaMemsetAsync(slots.ptr, 0xFF, slots.bytes, s));
    const uint64_t ident = identityBits(aggFunc, op);
    if (ident == 0) ARES_CUDA(cudaMemsetAsync(acc.ptr, 0, acc.bytes, s));
    else fillIdentityKernel<<<smCount() * 8, 256, 0, s>>>(acc.as<uint8_t>(), cap, width, ident);
    int blocks = divUp(length, 256);
    if (blocks > smCount() * 16) blocks = smCount() * 16;
    hashInsertKernel<<<blocks, 256, 0, s>>>(inputKeys.DimValues, L, length, inputValues, width, op,
                                            slots.as<unsigned long long>(), acc.as<uint8_t>(), (uint32_t)(cap - 1));
    checkLastError("hashInsert");
    const int tiles = divUp((int64_t)cap, kExtTile);
    Scratch state(scanStateBytes(tiles) + sizeof(uint32_t), s);
    ARES_CUDA(cudaMemsetAsync(state.ptr, 0, state.bytes, s));
    ScanTileState st = makeScanState(state.ptr, tiles);
    uint32_t *dCount = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(state.ptr) + scanStateBytes(tiles));
    Scratch rows(sizeof(uint32_t) * (size_t)length, s);
    extractSlotsKernel<<<tiles, kExtThreads, 0, s>>>(slots.as<unsigned long long>(), acc.as<uint8_t>(), width, cap, st,
                                                     rows.as<uint32_t>(), outputValues, dCount);
    checkLastError("extractSlots");
    uint32_t g = 0;
    ARES_CUDA(cudaMemcpyAsync(&g, dCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    ARES_CUDA(cudaStreamSynchronize(s));
    gatherDims(inputKeys.DimValues, L, rows.as<uint32_t>(), (int)g, outputKeys.DimValues, L, s);
    ARES_CUDA(cudaStreamSynchronize(s));
    return g;
  });
}
