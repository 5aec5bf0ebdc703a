// dimrow.cuh — the packed dimension row and the column-major DimensionVector block.
//
// Group identity in the reference is a hash of the *packed row*: the values of all dimensions
// in layout order (16-, 8-, 4-, 2-, 1-byte columns) followed by one validity byte per
// dimension, in a zero-initialised 32-byte buffer, hashed over rowBytes = sum(widths) + numDims
// bytes with seed 0 (reference DimensionHashIterator, query/iterator.hpp:934-1025).  Here the
// row lives in four 64-bit registers.
#pragma once
#include "common.cuh"
#include "murmur.cuh"

namespace aresb {

constexpr int kMaxDims = 16;  // a row is <= 32 bytes including validity, so <= 16 dims

struct DimLayout {
  uint64_t valueOff[kMaxDims];  // byte offset of the dim's value column in the block
  uint64_t nullOff[kMaxDims];   // byte offset of the dim's validity column
  uint8_t width[kMaxDims];
  uint8_t rowOff[kMaxDims];     // byte offset of the value inside the packed row
  int32_t numDims;
  int32_t valueBytes;           // sum of widths
  int32_t rowBytes;             // valueBytes + numDims
  int32_t capacity;
};

inline DimLayout makeDimLayout(const uint8_t numDimsPerDimWidth[NUM_DIM_WIDTH], int capacity) {
  DimLayout L;
  memset(&L, 0, sizeof(L));
  L.capacity = capacity;
  uint64_t pos = 0;
  int row = 0, n = 0;
  for (int i = 0; i < NUM_DIM_WIDTH; i++) {
    int w = 1 << (NUM_DIM_WIDTH - 1 - i);
    for (int j = 0; j < numDimsPerDimWidth[i]; j++) {
      if (n >= kMaxDims) throw EngineError("too many dimensions");
      L.valueOff[n] = pos;
      L.width[n] = (uint8_t)w;
      L.rowOff[n] = (uint8_t)row;
      pos += (uint64_t)w * capacity;
      row += w;
      n++;
    }
  }
  L.numDims = n;
  L.valueBytes = row;
  L.rowBytes = row + n;
  if (L.rowBytes > MAX_DIMENSION_BYTES) throw EngineError("dimension row exceeds MAX_DIMENSION_BYTES");
  for (int d = 0; d < n; d++) {
    L.nullOff[d] = pos;
    pos += capacity;
  }
  return L;
}

inline size_t dimBlockBytes(const DimLayout &L) { return (size_t)L.rowBytes * (size_t)L.capacity; }

#ifdef __CUDACC__
__device__ __forceinline__ void rowInsert(uint64_t w[4], int byteOff, uint64_t v) {
  const int word = byteOff >> 3, sh = (byteOff & 7) * 8;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (k == word) w[k] |= v << sh;
}

__device__ __forceinline__ uint64_t rowExtract(const uint64_t w[4], int byteOff, int width) {
  const int word = byteOff >> 3, sh = (byteOff & 7) * 8;
  uint64_t x = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (k == word) x = w[k];
  x >>= sh;
  return width >= 8 ? x : (x & ((1ull << (width * 8)) - 1ull));
}

// Packs row `idx` of a DimensionVector block into w[0..3].
__device__ __forceinline__ void packRow(const uint8_t *__restrict__ block, const DimLayout &L, uint32_t idx,
                                        uint64_t w[4]) {
  w[0] = w[1] = w[2] = w[3] = 0;
  for (int d = 0; d < L.numDims; d++) {
    const uint8_t *p = block + L.valueOff[d] + (size_t)idx * L.width[d];
    switch (L.width[d]) {
      case 16:
        rowInsert(w, L.rowOff[d], reinterpret_cast<const uint64_t *>(p)[0]);
        rowInsert(w, L.rowOff[d] + 8, reinterpret_cast<const uint64_t *>(p)[1]);
        break;
      case 8: rowInsert(w, L.rowOff[d], *reinterpret_cast<const uint64_t *>(p)); break;
      case 4: rowInsert(w, L.rowOff[d], *reinterpret_cast<const uint32_t *>(p)); break;
      case 2: rowInsert(w, L.rowOff[d], *reinterpret_cast<const uint16_t *>(p)); break;
      default: rowInsert(w, L.rowOff[d], *p); break;
    }
    rowInsert(w, L.valueBytes + d, block[L.nullOff[d] + idx]);
  }
}

// Writes a packed row to row `to` of a block laid out by `L`.
__device__ __forceinline__ void unpackRow(uint8_t *__restrict__ block, const DimLayout &L, uint32_t to,
                                          const uint64_t w[4]) {
  for (int d = 0; d < L.numDims; d++) {
    uint8_t *p = block + L.valueOff[d] + (size_t)to * L.width[d];
    switch (L.width[d]) {
      case 16:
        reinterpret_cast<uint64_t *>(p)[0] = rowExtract(w, L.rowOff[d], 8);
        reinterpret_cast<uint64_t *>(p)[1] = rowExtract(w, L.rowOff[d] + 8, 8);
        break;
      case 8: *reinterpret_cast<uint64_t *>(p) = rowExtract(w, L.rowOff[d], 8); break;
      case 4: *reinterpret_cast<uint32_t *>(p) = (uint32_t)rowExtract(w, L.rowOff[d], 4); break;
      case 2: *reinterpret_cast<uint16_t *>(p) = (uint16_t)rowExtract(w, L.rowOff[d], 2); break;
      default: *p = (uint8_t)rowExtract(w, L.rowOff[d], 1); break;
    }
    block[L.nullOff[d] + to] = (uint8_t)rowExtract(w, L.valueBytes + d, 1);
  }
}

// Column-wise copy of one row between two blocks that share dim widths (capacities may differ).
__device__ __forceinline__ void copyRow(const uint8_t *__restrict__ in, const DimLayout &Lin, uint32_t from,
                                        uint8_t *__restrict__ out, const DimLayout &Lout, uint32_t to) {
  for (int d = 0; d < Lin.numDims; d++) {
    const uint8_t *p = in + Lin.valueOff[d] + (size_t)from * Lin.width[d];
    uint8_t *q = out + Lout.valueOff[d] + (size_t)to * Lout.width[d];
    switch (Lin.width[d]) {
      case 16:
        reinterpret_cast<uint64_t *>(q)[0] = reinterpret_cast<const uint64_t *>(p)[0];
        reinterpret_cast<uint64_t *>(q)[1] = reinterpret_cast<const uint64_t *>(p)[1];
        break;
      case 8: *reinterpret_cast<uint64_t *>(q) = *reinterpret_cast<const uint64_t *>(p); break;
      case 4: *reinterpret_cast<uint32_t *>(q) = *reinterpret_cast<const uint32_t *>(p); break;
      case 2: *reinterpret_cast<uint16_t *>(q) = *reinterpret_cast<const uint16_t *>(p); break;
      default: *q = *p; break;
    }
    out[Lout.nullOff[d] + to] = in[Lin.nullOff[d] + from];
  }
}
#endif

}  // namespace aresb
