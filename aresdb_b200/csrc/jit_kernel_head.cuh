// jit_kernel_head.cuh — first half of the NVRTC-specialised fused kernel's source text: runtime
// parameter block and the quad loaders.  Compiled ONLY by NVRTC (see jit.cu); the generated
// rowEval() for one plan shape is pasted between this file and jit_kernel_tail.cuh, after a
// block of #defines / constexpr tables describing the shape (tile size, stage layout, key words,
// aggregate op).  Everything shape-dependent is a compile-time constant; only pointers, literal
// operands and row counts are runtime values, so one compiled kernel serves every batch and every
// query with the same structure.
namespace aresb {

constexpr int kJitMaxParts = 32;
constexpr int kJitMaxConsts = 64;
constexpr int kJitMaxWide = 8;
constexpr int kJitMaxMagic = 16;
constexpr int kJitMaxDenseDims = 8;
constexpr int kJitMaxRle = 4;
// A mode-3 column read by the kernel straight from its runs: cumulative counts (length + 1 entries), null bitmap and
// values of the RUNS, and the per-tile run hint computed by rleTileRunsKernel.
struct RleColumn {
  const uint32_t *counts;
  const uint8_t *nulls, *values;
  const uint32_t *tileRun;
  uint32_t length, startBit;
};

// mirrors plan_device.cuh
constexpr int kMaxForeignTables = 4;
constexpr int kMaxForeignCols = 8;
struct DevJoin {
  CuckooDesc tables[kMaxForeignTables];
  ForeignDesc cols[kMaxForeignCols];
};

struct JitParams {
  const uint8_t *partSrc[kJitMaxParts];   // global base address of every staged part
  const uint8_t *wideValues[kJitMaxWide]; // 8/16-byte dimension columns read straight from global
  const uint8_t *wideNulls[kJitMaxWide];
  uint32_t consts[kJitMaxConsts];         // literal operands / mode-0 defaults (raw 32-bit cells)
  unsigned long long magic[kJitMaxMagic]; // 2^64/d + 1 for literal divisors d > 0 (0: use the generic path)
  unsigned long long measureIdentity;
  unsigned long long accNeutral;
  DevTable G;
  unsigned long long *ctaAcc;             // [grid][JIT_SMEM_SLOTS] accumulator slices in global memory
  uint32_t numFullTiles;
  uint32_t numRows;                       // rows of the batch (tail = numRows - numFullTiles * JIT_TILE_ROWS)
  // direct-indexed aggregation (JIT_DENSE): dimension k of a row has index (value or quotient) - dLo[k], valid when
  // below dCnt[k]; index dCnt[k] is the dimension's NULL; slot = sum_k index_k * dStride[k]; value = (dLo + index) * dStep
  uint32_t dLo[kJitMaxDenseDims], dCnt[kJitMaxDenseDims], dStride[kJitMaxDenseDims], dStep[kJitMaxDenseDims];
  uint32_t dBase[kJitMaxDenseDims], dSpan[kJitMaxDenseDims], dMagic32[kJitMaxDenseDims];   // span division (see plan_device.cuh)
  uint32_t dStrideB[kJitMaxDenseDims];   // dStride in the unit the fast path addresses slots in (bytes for the integer form: x 12)
  uint32_t dTotal, dReps, dRepStride;     // slots of one copy; lane-private copies (power of two), dRepStride slots apart
  unsigned long long *gAcc;               // JIT_DENSE == 2: the state's global accumulator array (dTotal slots in use)
  double fxInv;                           // JIT_DENSE_ACC == 4: 2^-S
  float fxScale;                          //                     2^S (a float sum's rows are added as integers x * 2^S)
  uint32_t fxPad;
  const DevJoin *join;                    // joined dimension tables (join.cuh), null without joins
  uint32_t resume;
  uint32_t startCount;                    // row number of index position 0 when the batch has no base counts
  uint32_t partShift;                     // JIT_PARTITION: partition = home slot >> partShift
  uint4 *partBuf;                         //   entries (key, measure) of this batch, tile by tile, sorted by partition inside a tile
  uint32_t *partDir;                      //   per tile: segment offsets + span base
  uint32_t *partCursor;                   //   entries appended so far
  RleColumn rle[kJitMaxRle];              // run-length encoded columns decoded in place (see ldrle)                        // 1: second launch of the same batch after the table grew (progress[] says where)
};


// rows 4q .. 4q+3 of a staged value column of W bytes per value
template <int W, bool SIGNED>
__device__ __forceinline__ void ldq(const uint8_t *vals, uint32_t q, uint32_t (&v)[4]) {
  if (W == 4) {
    uint4 x = *reinterpret_cast<const uint4 *>(vals + 16 * q);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  } else if (W == 2) {
    uint2 x = *reinterpret_cast<const uint2 *>(vals + 8 * q);
    if (SIGNED) {
      v[0] = (uint32_t)(int32_t)(int16_t)(x.x & 0xffff); v[1] = (uint32_t)((int32_t)x.x >> 16);
      v[2] = (uint32_t)(int32_t)(int16_t)(x.y & 0xffff); v[3] = (uint32_t)((int32_t)x.y >> 16);
    } else {
      v[0] = x.x & 0xffff; v[1] = x.x >> 16; v[2] = x.y & 0xffff; v[3] = x.y >> 16;
    }
  } else {
    uint32_t x = *reinterpret_cast<const uint32_t *>(vals + 4 * q);
#pragma unroll
    for (int r = 0; r < 4; r++)
      v[r] = SIGNED ? (uint32_t)(int32_t)(int8_t)((x >> (8 * r)) & 0xff) : ((x >> (8 * r)) & 0xff);
  }
}

// Gather forms (compacted survivors: four arbitrary rows of the tile): same results as ldq / ldbits for rows[0..3].
template <int W, bool SIGNED>
__device__ __forceinline__ void ldqg(const uint8_t *vals, const uint32_t (&rows)[4], uint32_t (&v)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    if (W == 4) v[r] = reinterpret_cast<const uint32_t *>(vals)[rows[r]];
    else if (W == 2) v[r] = SIGNED ? (uint32_t)(int32_t)reinterpret_cast<const int16_t *>(vals)[rows[r]] : reinterpret_cast<const uint16_t *>(vals)[rows[r]];
    else v[r] = SIGNED ? (uint32_t)(int32_t)reinterpret_cast<const int8_t *>(vals)[rows[r]] : vals[rows[r]];
  }
}
template <int START_BIT>
__device__ __forceinline__ uint32_t ldbitsg(const uint8_t *bits, const uint32_t (&rows)[4]) {
  uint32_t nib = 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t bit = rows[r] + START_BIT;
    nib |= ((bits[bit >> 3] >> (bit & 7)) & 1u) << r;
  }
  return nib;
}

// 4 consecutive bits (rows 4q..4q+3) of a bit-packed vector whose row 0 sits at bit START_BIT
template <int START_BIT>
__device__ __forceinline__ uint32_t ldbits(const uint8_t *bits, uint32_t q) {
  if (START_BIT == 0) return (bits[q >> 1] >> ((q & 1) * 4)) & 0xF;
  const uint32_t bit = 4 * q + START_BIT;
  const uint32_t w = bits[bit >> 3] | ((uint32_t)bits[(bit >> 3) + 1] << 8);
  return w >> (bit & 7);  // callers test bits 0..3
}

// rows 4q .. 4q+3 of a run-length encoded column, decoded from its RUNS (the column is never expanded): the row numbers of
// the quad's index positions are increasing, and the run of every position of tile `tile` lies in
// [tileRun[tile], tileRun[tile + 1]] — usually one or two runs — so the quad's first row is found by a search over that
// window and the next three by stepping forward.  Neighbouring threads read the same few counts / values (broadcast hits).
template <int W, bool SIGNED>
__device__ __forceinline__ void ldrle(const RleColumn &R, uint32_t tile, const uint32_t (&rows)[4], uint32_t (&v)[4], uint32_t &validNibble) {
  uint32_t lo = R.tileRun[tile], hi = R.tileRun[tile + 1];
  while (lo < hi) {   // last run in [lo, hi] that starts at or before rows[0]
    const uint32_t mid = lo + ((hi - lo + 1) >> 1);
    if (R.counts[mid] <= rows[0]) lo = mid; else hi = mid - 1;
  }
  uint32_t p = lo;
  const uint32_t last = R.length - 1;
  validNibble = 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    while (p < last && R.counts[p + 1] <= rows[r]) p++;
    if (W == 0) v[r] = (R.values[(p + R.startBit) >> 3] >> ((p + R.startBit) & 7)) & 1u;
    else if (W == 1) v[r] = SIGNED ? (uint32_t)(int32_t)reinterpret_cast<const int8_t *>(R.values)[p] : R.values[p];
    else if (W == 2) v[r] = SIGNED ? (uint32_t)(int32_t)reinterpret_cast<const int16_t *>(R.values)[p] : reinterpret_cast<const uint16_t *>(R.values)[p];
    else v[r] = reinterpret_cast<const uint32_t *>(R.values)[p];
    validNibble |= ((R.nulls[(p + R.startBit) >> 3] >> ((p + R.startBit) & 7)) & 1u) << r;
  }
}

// x / d for a runtime-constant divisor as the high word of a 64x32-bit product (Lemire's fastdiv):
// M = floor((2^64 - 1) / d) + 1; exact for every 32-bit x and d >= 1.  Two IMAD.WIDE.
__device__ __forceinline__ uint32_t fastDivU32(uint32_t x, unsigned long long M) {
  const uint32_t carry = __umulhi((uint32_t)M, x);
  const unsigned long long hi = (unsigned long long)(uint32_t)(M >> 32) * x + carry;
  return (uint32_t)(hi >> 32);
}
// Quotient / remainder / floor-to-multiple with C semantics (truncation, sign follows the dividend)
// for a signed or unsigned 32-bit x and a positive literal d.  WHAT: 0 = x / d, 1 = x % d, 2 = x - x % d.
template <int WHAT, bool SIGNED>
__device__ __forceinline__ uint32_t fastDivOp(uint32_t x, uint32_t d, unsigned long long M) {
  const bool neg = SIGNED && (int32_t)x < 0;
  const uint32_t ux = neg ? 0u - x : x;
  const uint32_t q = fastDivU32(ux, M);
  const uint32_t res = WHAT == 0 ? q : WHAT == 1 ? ux - q * d : q * d;
  return neg ? 0u - res : res;
}

__device__ __forceinline__ bool bitOf(const uint8_t *p, uint32_t bit) { return (p[bit >> 3] >> (bit & 7)) & 1; }

}  // namespace aresb
