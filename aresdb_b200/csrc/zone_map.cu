// zone_map.cu — ComputeColumnRanges: the engine produces its own zone maps.
//
// BatchPlan.Ranges (min / max of the VALID values of a column in one batch) is what lets the fused kernel address
// its accumulators directly by dimension value and accumulate bounded float sums as exact integers.  The reference
// keeps such a pair only for live Uint32 vector parties (memstore/live_vector_party.go:74-75, read by
// query/aql_processor.go:1505-1509 for batch skipping); here one kernel computes it for every fixed-width column of
// any batch — live or archive, 1 / 2 / 4-byte integers, bools and Float32 — when the batch becomes device resident
// (once per batch, amortised over every query that reads it).  One launch covers all columns of the batch: every
// column is a streaming read of its values + null bitmap (16-byte loads, 8 rows x 4 per thread), block-reduced and
// folded into the column's slot with one atomicMin / atomicMax pair per CTA.
#include "column.cuh"
#include "common.cuh"

namespace aresb {

constexpr int kZmThreads = 256;
constexpr int kZmMaxCols = 16;

struct ZmColumn {
  const uint8_t *values;
  const uint8_t *nulls;    // null: every row valid
  uint32_t length;         // stored values (rows; runs for RLE columns)
  uint8_t width;           // 0: bit-packed bool
  uint8_t isSigned, isFloat, startBit;
};
struct ZmArgs {
  ZmColumn cols[kZmMaxCols];
  int ncols;
};
// per column: [0] min, [1] max (as ordered uint32: signed values biased by 2^31), [2] count of valid values,
// [3] flag: a valid value the range cannot describe (negative / non-finite float)
struct ZmOut { uint32_t lo, hi, any, bad; };

__device__ __forceinline__ uint32_t zmLoad(const ZmColumn &c, uint32_t i) {
  switch (c.width) {
    case 0: return bitAt(c.values, i + c.startBit) ? 1u : 0u;
    case 1: return c.isSigned ? (uint32_t)(int32_t)reinterpret_cast<const int8_t *>(c.values)[i] : c.values[i];
    case 2: return c.isSigned ? (uint32_t)(int32_t)reinterpret_cast<const int16_t *>(c.values)[i]
                              : reinterpret_cast<const uint16_t *>(c.values)[i];
    default: return reinterpret_cast<const uint32_t *>(c.values)[i];
  }
}

__global__ void __launch_bounds__(kZmThreads) columnRangesKernel(const __grid_constant__ ZmArgs A, ZmOut *__restrict__ out) {
  __shared__ uint32_t sLo[kZmThreads / 32], sHi[kZmThreads / 32], sAny[kZmThreads / 32], sBad[kZmThreads / 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ci = 0; ci < A.ncols; ci++) {
    const ZmColumn &c = A.cols[ci];
    uint32_t lo = 0xFFFFFFFFu, hi = 0u, any = 0u, bad = 0u;
    const uint32_t bias = c.isSigned ? 0x80000000u : 0u;
    // groups of 32 rows: one null word per lane-group when the bitmap is byte aligned, else bit by bit
    const uint32_t groups = (c.length + 31) / 32;
    for (uint32_t g = blockIdx.x * kZmThreads + threadIdx.x; g < groups; g += gridDim.x * kZmThreads) {
      const uint32_t base = g * 32, n = c.length - base < 32 ? c.length - base : 32;
      uint32_t valid = 0xFFFFFFFFu;
      if (c.nulls) {
        if (c.startBit == 0 && n == 32) {
          valid = (uint32_t)c.nulls[base / 8] | ((uint32_t)c.nulls[base / 8 + 1] << 8) | ((uint32_t)c.nulls[base / 8 + 2] << 16) |
                  ((uint32_t)c.nulls[base / 8 + 3] << 24);
        } else {
          valid = 0;
          for (uint32_t k = 0; k < n; k++) valid |= (bitAt(c.nulls, base + k + c.startBit) ? 1u : 0u) << k;
        }
      }
      if (n < 32) valid &= (1u << n) - 1u;
      if (valid == 0) continue;
      any = 1;
      if (c.width == 4 && n == 32) {   // 8 x 16-byte loads (values are 64-byte aligned by the memstore layout)
        const uint4 *p = reinterpret_cast<const uint4 *>(c.values + (size_t)base * 4);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint4 x = p[k];
          const uint32_t v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int r = 0; r < 4; r++) {
            if (!((valid >> (4 * k + r)) & 1)) continue;
            if (c.isFloat && v[r] >= 0x7F800000u) { bad = 1; continue; }   // negative (sign bit), inf or NaN
            const uint32_t o = v[r] + bias;
            lo = o < lo ? o : lo; hi = o > hi ? o : hi;
          }
        }
      } else {
        for (uint32_t k = 0; k < n; k++) {
          if (!((valid >> k) & 1)) continue;
          const uint32_t v = zmLoad(c, base + k);
          if (c.isFloat && v >= 0x7F800000u) { bad = 1; continue; }
          const uint32_t o = v + bias;
          lo = o < lo ? o : lo; hi = o > hi ? o : hi;
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t l2 = __shfl_xor_sync(0xFFFFFFFFu, lo, d), h2 = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
      lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
      any |= __shfl_xor_sync(0xFFFFFFFFu, any, d); bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, d);
    }
    if (lane == 0) { sLo[warp] = lo; sHi[warp] = hi; sAny[warp] = any; sBad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kZmThreads / 32; w++) {
        lo = sLo[w] < lo ? sLo[w] : lo; hi = sHi[w] > hi ? sHi[w] : hi; any |= sAny[w]; bad |= sBad[w];
      }
      if (any) { atomicMin(&out[ci].lo, lo); atomicMax(&out[ci].hi, hi); atomicOr(&out[ci].any, 1u); }
      if (bad) atomicOr(&out[ci].bad, 1u);
    }
    __syncthreads();
  }
}

}  // namespace aresb

using namespace aresb;

extern "C" CGoCallResHandle ComputeColumnRanges(const VectorPartySlice *columns, int numColumns, ColumnRange *out, void *cudaStream,
                                                int device) {
  return guarded("ComputeColumnRanges", device, [&]() -> int64_t {
    if (!columns || !out || numColumns < 0 || numColumns > kZmMaxCols) throw EngineError("invalid column list (at most 16 columns per call)");
    cudaStream_t s = (cudaStream_t)cudaStream;
    ZmArgs A;
    memset(&A, 0, sizeof(A));
    int slotOf[kZmMaxCols];
    uint32_t longest = 0;
    for (int i = 0; i < numColumns; i++) {
      out[i].Known = 0; out[i].Min = 0; out[i].Max = 0;
      slotOf[i] = -1;
      const VectorPartySlice &vp = columns[i];
      const int dt = vp.DataType;
      const int width = dt == Bool ? 0 : (dt == Int8 || dt == Uint8) ? 1 : (dt == Int16 || dt == Uint16) ? 2
                      : (dt == Int32 || dt == Uint32 || dt == Float32) ? 4 : -1;
      if (width < 0) continue;                       // 8 / 16-byte columns carry no zone map
      InputDesc d = makeColumnDesc(vp, /*allowWide=*/true);
      if (d.mode == 0) {                             // constant column: the default value is the range
        const uint32_t v = (uint32_t)d.constLo;
        const bool ok = d.constValid && (dt == Float32 ? v < 0x7F800000u : v < 0x80000000u);
        if (ok) { out[i].Known = 1; out[i].Min = v; out[i].Max = v; }
        continue;
      }
      if (d.length == 0) continue;
      ZmColumn &c = A.cols[A.ncols];
      c.values = d.base + d.valuesOff;
      c.nulls = d.mode >= 2 ? d.base + d.nullsOff : nullptr;
      c.length = d.length;
      c.width = (uint8_t)width;
      c.isSigned = dt == Int8 || dt == Int16 || dt == Int32;
      c.isFloat = dt == Float32;
      c.startBit = d.startBit;
      longest = d.length > longest ? d.length : longest;
      slotOf[i] = A.ncols++;
    }
    if (A.ncols == 0) return 0;
    Scratch res(sizeof(ZmOut) * kZmMaxCols, s);
    ZmOut init[kZmMaxCols];
    for (int i = 0; i < kZmMaxCols; i++) init[i] = ZmOut{0xFFFFFFFFu, 0u, 0u, 0u};
    ARES_CUDA(cudaMemcpyAsync(res.ptr, init, sizeof(init), cudaMemcpyHostToDevice, s));
    int blocks = divUp((int64_t)(longest + 31) / 32, kZmThreads);
    if (blocks > smCount() * 8) blocks = smCount() * 8;
    columnRangesKernel<<<blocks, kZmThreads, 0, s>>>(A, res.as<ZmOut>());
    checkLastError("ComputeColumnRanges");
    ZmOut host[kZmMaxCols];
    ARES_CUDA(cudaMemcpyAsync(host, res.ptr, sizeof(host), cudaMemcpyDeviceToHost, s));
    ARES_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < numColumns; i++) {
      if (slotOf[i] < 0) continue;
      const ZmOut &z = host[slotOf[i]];
      const ZmColumn &c = A.cols[slotOf[i]];
      if (!z.any || z.bad) continue;
      const uint32_t bias = c.isSigned ? 0x80000000u : 0u;
      const uint32_t lo = z.lo - bias, hi = z.hi - bias;
      // the engine's contract: non-negative integers below 2^31 (float: bit patterns of non-negative finite values)
      if (c.isSigned && (int32_t)lo < 0) continue;
      if (!c.isFloat && hi >= 0x80000000u) continue;
      out[i].Known = 1; out[i].Min = lo; out[i].Max = hi;
    }
    return A.ncols;
  });
}
