// scan.cuh — block-level exclusive scan and the single-pass "decoupled look-back" chained
// scan across thread blocks (Merrill & Garland's formulation, written from the published
// algorithm) used for stable stream compaction and radix-sort scatter offsets.  No Thrust/CUB.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aresb {

// One 64-bit word per tile: high 32 bits = status, low 32 bits = value.
enum : uint32_t { SCAN_EMPTY = 0, SCAN_AGGREGATE = 1, SCAN_PREFIX = 2 };

struct ScanTileState {
  unsigned long long *words;  // [tiles], zero-initialised before launch
  uint32_t *ticket;           // dynamic tile id (launch order != blockIdx order is not assumed)
};

inline size_t scanStateBytes(int tiles) { return sizeof(unsigned long long) * (size_t)(tiles + 1); }
inline ScanTileState makeScanState(void *mem, int tiles) {
  ScanTileState s;
  s.words = reinterpret_cast<unsigned long long *>(mem);
  s.ticket = reinterpret_cast<uint32_t *>(s.words + tiles);
  return s;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t x) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= (uint32_t)d) x += y;
  }
  return x;
}

__device__ __forceinline__ uint32_t warpReduceSum(uint32_t x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
  return x;
}

// Exclusive prefix of x over the block (thread order); *total = block sum.  sWarp needs
// THREADS/32 + 1 words.  Contains a __syncthreads(): every global load issued by the block
// before this call is ordered before anything a thread does after it.
template <int THREADS>
__device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t x, uint32_t *sWarp, uint32_t *total) {
  constexpr int W = THREADS / 32;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = warpInclusiveScan(x);
  if (lane == 31) sWarp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t v = lane < W ? sWarp[lane] : 0;
    uint32_t s = warpInclusiveScan(v);
    if (lane < W) sWarp[lane] = s - v;  // exclusive warp offsets
    if (lane == W - 1) sWarp[W] = s;
  }
  __syncthreads();
  *total = sWarp[W];
  return sWarp[warp] + inc - x;
}

__device__ __forceinline__ unsigned long long loadWord(const unsigned long long *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ void storeWord(unsigned long long *p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// Called by ALL 32 lanes of one warp of the block owning `tile`, after the block-wide
// barrier that follows the tile's loads.  Publishes the tile aggregate, waits for the
// predecessors and returns the exclusive prefix of the tile (same value in every lane).
__device__ __forceinline__ uint32_t decoupledLookback(ScanTileState st, uint32_t tile, uint32_t aggregate) {
  const uint32_t lane = threadIdx.x & 31;
  __threadfence();  // order the tile's prior loads before the publication (in-place compaction)
  if (tile == 0) {
    if (lane == 0) storeWord(&st.words[0], ((unsigned long long)SCAN_PREFIX << 32) | aggregate);
    return 0;
  }
  if (lane == 0) storeWord(&st.words[tile], ((unsigned long long)SCAN_AGGREGATE << 32) | aggregate);
  uint32_t exclusive = 0;
  int64_t look = (int64_t)tile - 1;
  while (true) {
    int64_t idx = look - lane;
    unsigned long long w;
    do {
      w = idx >= 0 ? loadWord(&st.words[idx]) : ((unsigned long long)SCAN_PREFIX << 32);
    } while (__any_sync(0xffffffffu, (uint32_t)(w >> 32) == SCAN_EMPTY));
    const uint32_t status = (uint32_t)(w >> 32), value = (uint32_t)w;
    const uint32_t prefMask = __ballot_sync(0xffffffffu, status == SCAN_PREFIX);
    const uint32_t firstPref = prefMask ? (uint32_t)(__ffs(prefMask) - 1) : 32u;
    exclusive += warpReduceSum(lane <= firstPref ? value : 0u);
    if (prefMask) break;
    look -= 32;
  }
  if (lane == 0)
    storeWord(&st.words[tile], ((unsigned long long)SCAN_PREFIX << 32) | (exclusive + aggregate));
  __threadfence();  // acquire side: stores below must not pass the status reads above
  return exclusive;
}
#endif  // __CUDACC__

}  // namespace aresb
