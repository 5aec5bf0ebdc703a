// column.cuh — VectorParty column-slice decode.
//
// What a slice looks like is documented in include/aresdb_b200/aql_abi.h (VectorPartySlice);
// the access rules restated here follow the reference's VectorPartyIterator
// (query/iterator.hpp:62-289): 1/2-byte integers widen to (u)int32, Bool values and the
// null bitmap are bit-packed LSB-first with StartingIndex as bit offset of row 0, a
// mode-3 (RLE) column is addressed by searching its cumulative-count vector for the run
// that contains the row number, and the raw stored value is returned even when the row is
// NULL (values under nulls are NOT canonicalised).
#pragma once
#include "cell.cuh"

namespace aresb {

enum InputKind : uint8_t { IN_COLUMN = 0, IN_SCRATCH = 1, IN_CONST = 2, IN_FOREIGN = 3 /* ForeignColumnInput: see join.cuh */ };

struct InputDesc {
  const uint8_t *base;     // column: BasePtr; scratch: Values
  uint32_t nullsOff;       // column: NullsOffset; scratch: byte offset of the bool vector
  uint32_t valuesOff;
  uint32_t length;         // column: number of stored values (runs for mode 3)
  uint64_t constLo;        // const / mode-0 default value bits (UUID: p1)
  uint64_t constHi;        //                                   (UUID: p2)
  uint8_t kind;            // InputKind
  uint8_t mode;            // column mode 0..3
  uint8_t startBit;
  uint8_t dtype;           // enum DataType of the stored values
  uint8_t vclass;          // ValClass the reference's iterator would yield
  uint8_t constValid;
};

#ifndef __CUDACC_RTC__
// Host: ABI struct -> descriptor; throws EngineError for inputs outside the hot path.
InputDesc makeInputDesc(const InputVector &in, bool allowWide);
InputDesc makeColumnDesc(const VectorPartySlice &vp, bool allowWide);
#endif

#ifdef __CUDACC__
__device__ __forceinline__ bool bitAt(const uint8_t *p, uint32_t bit) {
  return (p[bit >> 3] >> (bit & 7)) & 1;
}

// Position of `row` in an RLE column: last p in [0, length) with counts[p] <= row.
__device__ __forceinline__ uint32_t rlePosition(const uint32_t *counts, uint32_t length, uint32_t row) {
  uint32_t lo = 0, hi = length;
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (counts[mid] > row) hi = mid; else lo = mid + 1;
  }
  return lo == 0 ? 0 : lo - 1;
}

// Reads element `i` of the logical input vector.  For columns the row is index[i]
// (mode 3: the row number is baseCounts[index[i]] or startCount + index[i]); scratch and
// constants are positional.  hi receives the upper half of a UUID.
__device__ __forceinline__ Cell loadInput(const InputDesc &d, uint32_t i, const uint32_t *index,
                                          const uint32_t *baseCounts, uint32_t startCount,
                                          uint64_t *hi) {
  Cell c;
  if (d.kind == IN_CONST || (d.kind == IN_COLUMN && d.mode == 0)) {
    c.v = d.constLo; c.valid = d.constValid; if (hi) *hi = d.constHi; return c;
  }
  if (d.kind == IN_SCRATCH) {
    switch (d.dtype) {
      case Int64: case Uint64: c.v = reinterpret_cast<const uint64_t *>(d.base)[i]; break;
      case UUID: c.v = reinterpret_cast<const uint64_t *>(d.base)[2 * i];
                 if (hi) *hi = reinterpret_cast<const uint64_t *>(d.base)[2 * i + 1]; break;
      default: c.v = reinterpret_cast<const uint32_t *>(d.base)[i]; break;
    }
    // validity bytes sit NullsOffset bytes after the value vector start, one byte per row
    c.valid = d.base[d.nullsOff + i] != 0;
    return c;
  }
  uint32_t idx = index ? index[i] : i;
  uint32_t p = idx;
  if (d.mode == 3) {
    uint32_t row = baseCounts ? baseCounts[idx] : startCount + idx;
    p = rlePosition(reinterpret_cast<const uint32_t *>(d.base), d.length, row);
  }
  const uint8_t *vals = d.base + d.valuesOff;
  switch (d.dtype) {
    case Bool: c.v = bitAt(vals, p + d.startBit) ? 1 : 0; break;
    case Int8: c.v = (uint32_t)(int32_t)reinterpret_cast<const int8_t *>(vals)[p]; break;
    case Uint8: c.v = vals[p]; break;
    case Int16: c.v = (uint32_t)(int32_t)reinterpret_cast<const int16_t *>(vals)[p]; break;
    case Uint16: c.v = reinterpret_cast<const uint16_t *>(vals)[p]; break;
    case Int64: case Uint64: c.v = reinterpret_cast<const uint64_t *>(vals)[p]; break;
    case UUID: c.v = reinterpret_cast<const uint64_t *>(vals)[2 * (size_t)p];
               if (hi) *hi = reinterpret_cast<const uint64_t *>(vals)[2 * (size_t)p + 1]; break;
    default: c.v = reinterpret_cast<const uint32_t *>(vals)[p]; break;  // Int32 / Uint32 / Float32
  }
  c.valid = d.mode >= 2 ? bitAt(d.base + d.nullsOff, p + d.startBit) : true;
  return c;
}
#endif  // __CUDACC__

}  // namespace aresb
