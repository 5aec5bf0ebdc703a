// batch_plan.cu — the fused whole-batch engine behind include/aresdb_b200/batch_plan.h.
//
// ExecuteBatchPlan = ONE persistent kernel per batch (fusedBatchKernel):
//   * column slices are staged tile by tile into shared memory by the TMA engine
//     (cp.async.bulk global->shared, completion on an mbarrier, kStages-deep ring), so HBM is
//     read exactly once, fully coalesced, with no LSU issue slots or registers spent on it;
//   * every thread owns quads of 4 consecutive rows (128-bit LDS for 4-byte columns, 64-bit for
//     2-byte, 32-bit for 1-byte, nibbles of the null bitmaps) and interprets the plan's
//     instruction list in registers: filters clear bits of an alive mask, dimension roots pack
//     the row key, the measure root yields the value (NULL -> identity, x RLE count);
//   * surviving rows are aggregated into a CTA-private open-addressing table in SHARED memory
//     (native shared atomics: measured 0.5-2 Tops/s on B200, profiles/r01_agg_microbench.txt),
//     which is flushed into the L2-resident global group table when the CTA retires; rows that
//     do not fit the shared table (high cardinality) go to the global table directly.
// The global table lives in an AggState across batches.  AggStateFinalize compacts it, hashes
// each group's packed dimension row with the reference's murmur3, sorts the g groups by hash,
// merges equal hashes and writes the reference's output layout — the observable result of the
// reference's Sort+Reduce (ARES_REDUCE_SORT) or HashReduce (ARES_REDUCE_HASH) over all batches.
#include <cooperative_groups.h>
#include <memory>
#include <vector>

#include "agg.cuh"
#include "column.cuh"
#include "dimrow.cuh"
#include "fused_device.cuh"
#include "plan_device.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "small_sort.cuh"

namespace cg = cooperative_groups;

namespace aresb {

// from legacy_nodes.cu
ForeignDesc makeForeignDesc(const ForeignColumnVector &f);
// from sort_reduce.cu
void gatherDims(const uint8_t *in, const DimLayout &Lin, const uint32_t *rows, int g, uint8_t *out,
                const DimLayout &Lout, cudaStream_t s);
int hllRegisterVectors(const uint64_t *hash, const uint32_t *values, uint32_t *index, int R, uint8_t **hllVectorPtr,
                       size_t *hllVectorSizePtr, uint16_t **hllDimRegIDCountPtr, cudaStream_t s);
int reduceByHash(const uint64_t *hash, const uint32_t *index, const uint8_t *measures, int width, AggOp op, int n,
                 uint32_t *outIndex, uint8_t *outValues, cudaStream_t s, uint64_t *outHash = nullptr);


// ---------------------------------------------------------------------------------------
// 4-row vector evaluation
// ---------------------------------------------------------------------------------------
constexpr int R = 4;  // rows per quad
// status word of the single-launch finalize / of an exchange part
enum SmallFinalizeStatus : uint32_t { SF_OK = 0, SF_TOO_MANY = 1, SF_TABLE_OVERFLOW = 2, SF_OUTPUT_TOO_SMALL = 3, SF_PART_TRUNCATED = 4, SF_UNSETTLED = 5, SF_PEER_LATE = 6 };

__device__ __forceinline__ void cvtVec(uint32_t (&v)[R], ValClass from, ValClass to) {
  if (from == to) return;
  const bool fi = from == VC_I32 || from == VC_U32 || from == VC_BOOL;
  const bool ti = to == VC_I32 || to == VC_U32;
  if (fi && ti) return;  // bit-identical (bool is stored as 0/1)
#pragma unroll
  for (int r = 0; r < R; r++) v[r] = (uint32_t)cvt(v[r], from, to);
}

// Binary functor on R rows; NULL results carry value 0 like the reference's (0, false).
__device__ __forceinline__ void binVec(int fn, ValClass tc, const uint32_t (&a)[R], uint32_t av, const uint32_t (&b)[R],
                                       uint32_t bv, uint32_t (&out)[R], uint32_t &ov) {
  const uint32_t both = av & bv;
  ov = 0;
#define ARES_ROWS(expr_f, expr_i, expr_u)                                                                 \
  {                                                                                                       \
    ov = both;                                                                                            \
    if (tc == VC_F32) {                                                                                   \
      _Pragma("unroll") for (int r = 0; r < R; r++) {                                                     \
        float x = __uint_as_float(a[r]), y = __uint_as_float(b[r]); (void)x; (void)y;                     \
        out[r] = (both >> r) & 1 ? (uint32_t)(expr_f) : 0u;                                               \
      }                                                                                                   \
    } else if (tc == VC_I32) {                                                                            \
      _Pragma("unroll") for (int r = 0; r < R; r++) {                                                     \
        int32_t x = (int32_t)a[r], y = (int32_t)b[r]; (void)x; (void)y;                                   \
        out[r] = (both >> r) & 1 ? (uint32_t)(expr_i) : 0u;                                               \
      }                                                                                                   \
    } else {                                                                                              \
      _Pragma("unroll") for (int r = 0; r < R; r++) {                                                     \
        uint32_t x = a[r], y = b[r]; (void)x; (void)y;                                                    \
        out[r] = (both >> r) & 1 ? (uint32_t)(expr_u) : 0u;                                               \
      }                                                                                                   \
    }                                                                                                     \
  }
  switch (fn) {
    case Equal: ARES_ROWS(x == y, x == y, x == y) return;
    case NotEqual: ARES_ROWS(x != y, x != y, x != y) return;
    case LessThan: ARES_ROWS(x < y, x < y, x < y) return;
    case LessThanOrEqual: ARES_ROWS(x <= y, x <= y, x <= y) return;
    case GreaterThan: ARES_ROWS(x > y, x > y, x > y) return;
    case GreaterThanOrEqual: ARES_ROWS(x >= y, x >= y, x >= y) return;
    case Plus: ARES_ROWS(__float_as_uint(__fadd_rn(x, y)), (uint32_t)x + (uint32_t)y, x + y) return;
    case Minus: ARES_ROWS(__float_as_uint(__fsub_rn(x, y)), (uint32_t)x - (uint32_t)y, x - y) return;
    case Multiply: ARES_ROWS(__float_as_uint(__fmul_rn(x, y)), (uint32_t)x * (uint32_t)y, x * y) return;
    default: break;
  }
#undef ARES_ROWS
  // everything else (And/Or/Divide/Mod/bitwise/Floor): scalar functor per row
#pragma unroll
  for (int r = 0; r < R; r++) {
    Cell ca, cb;
    ca.v = a[r]; ca.valid = (av >> r) & 1;
    cb.v = b[r]; cb.valid = (bv >> r) & 1;
    ValClass rc;
    Cell cr = evalBinary(fn, ca, cb, tc, &rc);
    out[r] = (uint32_t)cr.v;
    ov = (ov & ~(1u << r)) | ((cr.valid ? 1u : 0u) << r);
  }
}

__device__ __forceinline__ void unVec(int fn, ValClass ic, const uint32_t (&a)[R], uint32_t av, uint32_t (&out)[R],
                                      uint32_t &ov) {
  if (fn == Noop) {
#pragma unroll
    for (int r = 0; r < R; r++) out[r] = a[r];
    ov = av;
    return;
  }
  ov = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    Cell ca;
    ca.v = a[r]; ca.valid = (av >> r) & 1;
    ValClass rc;
    Cell cr = evalUnary(fn, ca, ic, &rc);
    out[r] = (uint32_t)cr.v;
    ov = (ov & ~(1u << r)) | ((cr.valid ? 1u : 0u) << r);
  }
}

// ---------------------------------------------------------------------------------------
// operand fetch
// ---------------------------------------------------------------------------------------
// Staged: `stage` is the shared-memory copy of the tile; q is the quad index inside the tile.
__device__ __forceinline__ void fetchStaged(const DevColumn &c, const uint8_t *stage, uint32_t q, uint32_t (&v)[R],
                                            uint32_t &valid) {
  const uint8_t *vals = stage + c.smemValues;
  switch (c.width) {
    case 4: {
      uint4 x = *reinterpret_cast<const uint4 *>(vals + 16 * q);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
      break;
    }
    case 2: {
      uint2 x = *reinterpret_cast<const uint2 *>(vals + 8 * q);
      if (c.in.dtype == Int16) {
        v[0] = (uint32_t)(int32_t)(int16_t)(x.x & 0xffff); v[1] = (uint32_t)(int32_t)(int16_t)(x.x >> 16);
        v[2] = (uint32_t)(int32_t)(int16_t)(x.y & 0xffff); v[3] = (uint32_t)(int32_t)(int16_t)(x.y >> 16);
      } else {
        v[0] = x.x & 0xffff; v[1] = x.x >> 16; v[2] = x.y & 0xffff; v[3] = x.y >> 16;
      }
      break;
    }
    case 1: {
      uint32_t x = *reinterpret_cast<const uint32_t *>(vals + 4 * q);
      if (c.in.dtype == Int8) {
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = (uint32_t)(int32_t)(int8_t)((x >> (8 * r)) & 0xff);
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = (x >> (8 * r)) & 0xff;
      }
      break;
    }
    default: {  // bit-packed bool
      uint32_t bit = 4 * q + c.in.startBit;
      uint32_t w = vals[bit >> 3] | ((uint32_t)vals[(bit >> 3) + 1] << 8);
      w >>= (bit & 7);
#pragma unroll
      for (int r = 0; r < R; r++) v[r] = (w >> r) & 1;
      break;
    }
  }
  if (c.hasNulls) {
    const uint8_t *nb = stage + c.smemNulls;
    uint32_t bit = 4 * q + c.in.startBit;
    uint32_t w = nb[bit >> 3] | ((uint32_t)nb[(bit >> 3) + 1] << 8);
    valid = (w >> (bit & 7)) & 0xF;
  } else {
    valid = 0xF;
  }
}

// Direct: straight from global memory, any column mode, row by row (tail tiles, unaligned or
// RLE columns).
__device__ __forceinline__ void fetchDirect(const DevPlan &P, const DevColumn &c, uint32_t row0, uint32_t nrows,
                                            uint32_t (&v)[R], uint32_t &valid) {
  valid = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    v[r] = 0;
    if ((uint32_t)r < nrows) {
      Cell x = loadInput(c.in, row0 + r, nullptr, P.baseCounts, P.startCount, nullptr);
      v[r] = (uint32_t)x.v;
      valid |= (x.valid ? 1u : 0u) << r;
    }
  }
}

struct QuadState {
  uint32_t st[ARES_PLAN_STACK_DEPTH][R];
  uint32_t stv[ARES_PLAN_STACK_DEPTH];
};

template <bool STAGED>
__device__ __forceinline__ void getOperand(const DevPlan &P, const DevInst &I, bool second, const uint8_t *stage,
                                           uint32_t q, uint32_t row0, uint32_t nrows, QuadState &S, int &sp,
                                           uint32_t (&v)[R], uint32_t &valid) {
  const uint8_t kind = second ? I.bkind : I.akind;
  if (kind == OPK_COLUMN) {
    const DevColumn &c = P.cols[second ? I.bcol : I.acol];
    if (c.in.mode == 0) {
#pragma unroll
      for (int r = 0; r < R; r++) v[r] = (uint32_t)c.in.constLo;
      valid = c.in.constValid ? 0xF : 0;
    } else if (STAGED && !c.rle) {
      fetchStaged(c, stage, q, v, valid);
    } else {
      fetchDirect(P, c, row0, nrows, v, valid);
    }
  } else if (kind == OPK_CONST) {
    const uint32_t k = second ? I.bconst : I.aconst;
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = k;
    valid = (second ? I.bvalid : I.avalid) ? 0xF : 0;
  } else if (kind == OPK_FOREIGN) {
    // joined dimension table: probe its index with the row's join key, read the foreign column at the RecordID
    const uint8_t fc = second ? I.bcol : I.acol;
    const uint8_t t = P.foreignTableOf[fc];
    const DevColumn &jc = P.cols[P.joinCol[t]];
    uint32_t key[R], kvalid;
    if (jc.in.mode == 0) {
#pragma unroll
      for (int r = 0; r < R; r++) key[r] = (uint32_t)jc.in.constLo;
      kvalid = jc.in.constValid ? 0xF : 0;
    } else if (STAGED) {
      fetchStaged(jc, stage, q, key, kvalid);
    } else {
      fetchDirect(P, jc, row0, nrows, key, kvalid);
    }
    valid = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const unsigned long long rid = ((kvalid >> r) & 1) && (uint32_t)r < nrows ? cuckooLookup(P.join->tables[t], key[r], 0) : 0ull;
      const Cell f = foreignLoad(P.join->cols[fc], rid, nullptr);
      v[r] = (uint32_t)f.v;
      valid |= (f.valid ? 1u : 0u) << r;
    }
  } else {  // stack pop (static unrolled select keeps the stack in registers)
    sp--;
#pragma unroll
    for (int d = 0; d < ARES_PLAN_STACK_DEPTH; d++) {
      if (d == sp) {
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = S.st[d][r];
        valid = S.stv[d];
      }
    }
  }
}

// Processes one quad (rows row0 .. row0+nrows-1 of the batch).
template <int KW>
__device__ __forceinline__ void keyInsert(uint64_t (&w)[KW], int byteOff, uint64_t v) {
  if (KW == 1) {
    w[0] |= v << ((byteOff & 7) * 8);
  } else {
    const int word = byteOff >> 3, sh = (byteOff & 7) * 8;
#pragma unroll
    for (int k = 0; k < KW; k++)
      if (k == word) w[k] |= v << sh;
  }
}

template <bool STAGED, bool WIDEKEY>
__device__ __forceinline__ void processQuad(const DevPlan &P, const DevTable &G, const SmemTable &T, bool useSmem,
                                            bool allowClaim, const uint8_t *stage, uint32_t q, uint32_t row0,
                                            uint32_t nrows) {
  QuadState S;
  int sp = 0;
  uint32_t alive = (1u << nrows) - 1u;
  constexpr int KW = WIDEKEY ? 4 : 1;
  uint64_t kw[R][KW];
#pragma unroll
  for (int r = 0; r < R; r++) {
#pragma unroll
    for (int k = 0; k < KW; k++) kw[r][k] = 0;
  }
  uint64_t meas[R];
#pragma unroll
  for (int r = 0; r < R; r++) meas[r] = P.measureIdentity;

  for (int pc = 0; pc < P.ninsts; pc++) {
    const DevInst &I = P.insts[pc];
    if (I.wide) {  // 8/16-byte column copied verbatim into a dimension (Int64 / UUID dims)
      const DevColumn &c = P.cols[I.acol];
#pragma unroll
      for (int r = 0; r < R; r++) {
        if ((uint32_t)r < nrows) {
          uint64_t hi = 0;
          Cell x = loadInput(c.in, row0 + r, nullptr, P.baseCounts, P.startCount, &hi);
          keyInsert<KW>(kw[r], I.rowOff, x.v);
          if (I.width == 16) keyInsert<KW>(kw[r], I.rowOff + 8, hi);
          keyInsert<KW>(kw[r], I.nullOff, x.valid ? 1 : 0);
        }
      }
      continue;
    }
    uint32_t a[R], b[R], res[R];
    uint32_t av = 0, bv = 0, rv = 0;
    if (I.nops == 2) {
      // both on the stack: rhs was pushed last
      if (I.bkind == OPK_STACK) getOperand<STAGED>(P, I, true, stage, q, row0, nrows, S, sp, b, bv);
      getOperand<STAGED>(P, I, false, stage, q, row0, nrows, S, sp, a, av);
      if (I.bkind != OPK_STACK) getOperand<STAGED>(P, I, true, stage, q, row0, nrows, S, sp, b, bv);
      cvtVec(a, (ValClass)I.aclass, (ValClass)I.tclass);
      cvtVec(b, (ValClass)I.bclass, (ValClass)I.tclass);
      binVec(I.fn, (ValClass)I.tclass, a, av, b, bv, res, rv);
    } else {
      getOperand<STAGED>(P, I, false, stage, q, row0, nrows, S, sp, a, av);
      unVec(I.fn, (ValClass)I.aclass, a, av, res, rv);
    }
    switch (I.sink) {
      case PLAN_SINK_STACK: {
        cvtVec(res, (ValClass)I.rclass, (ValClass)I.oclass);
#pragma unroll
        for (int d = 0; d < ARES_PLAN_STACK_DEPTH; d++) {
          if (d == sp) {
#pragma unroll
            for (int r = 0; r < R; r++) S.st[d][r] = res[r];
            S.stv[d] = rv;
          }
        }
        sp++;
        break;
      }
      case PLAN_SINK_FILTER: {
        uint32_t keep = 0;
        if (I.rclass == VC_F32) {
#pragma unroll
          for (int r = 0; r < R; r++) keep |= (__uint_as_float(res[r]) != 0.0f ? 1u : 0u) << r;
        } else {
#pragma unroll
          for (int r = 0; r < R; r++) keep |= (res[r] != 0 ? 1u : 0u) << r;
        }
        alive &= keep;
        if (pc == P.lastFilter && !__any_sync(__activemask(), alive != 0)) return;
        break;
      }
      case PLAN_SINK_DIMENSION: {
#pragma unroll
        for (int r = 0; r < R; r++) {
          uint64_t o = cvt(res[r], (ValClass)I.rclass, (ValClass)I.oclass);
          keyInsert<KW>(kw[r], I.rowOff, o);
          keyInsert<KW>(kw[r], I.nullOff, (rv >> r) & 1);
        }
        break;
      }
      default: {  // PLAN_SINK_MEASURE
#pragma unroll
        for (int r = 0; r < R; r++) {
          if ((rv >> r) & 1) {
            uint64_t o = cvt(res[r], (ValClass)I.rclass, (ValClass)I.oclass);
            if (P.aggOp == OP_AVG) {  // (float average, count) pair, as assignAvg packs it (query/iterator.hpp:636-645)
              uint32_t cnt = 1;
              if (P.baseCounts != nullptr && (uint32_t)r < nrows) cnt = P.baseCounts[row0 + r + 1] - P.baseCounts[row0 + r];
              meas[r] = ((uint64_t)cnt << 32) | (uint32_t)cvt(o, (ValClass)I.oclass, VC_F32);
              continue;
            }
            if (!P.skipCount && P.baseCounts != nullptr && (uint32_t)r < nrows) {
              uint32_t cnt = P.baseCounts[row0 + r + 1] - P.baseCounts[row0 + r];
              o = mulCount(o, (ValClass)I.oclass, cnt);
            }
            meas[r] = o;
          }
        }
        break;
      }
    }
  }

  if (alive == 0) return;
  const AggOp op = (AggOp)P.aggOp;
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (!((alive >> r) & 1)) continue;
    unsigned long long key;
    const uint64_t *roww = nullptr;
    if constexpr (WIDEKEY) {
      key = P.hashBits == 64 ? murmur3_128_lo(kw[r], P.rowBytes, 0) : (unsigned long long)murmur3_32(kw[r], P.rowBytes, 0);
      if (P.hll == 1) key = (key & 0xFFFFFFFFFFFF0000ull) | (meas[r] & 0x3FFFu);
      roww = kw[r];
    } else {
      key = kw[r][0];
    }
    if (P.hll == 2) {  // dense registers (no shared mirror on this generic path)
      hllDenseUpdate(G, nullptr, key, roww, (uint32_t)meas[r]);
      continue;
    }
    if (!useSmem || !smemUpdate(T, G, op, key, roww, meas[r], allowClaim)) globalUpdate(G, op, key, roww, meas[r]);
  }
}

// ---------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void issueTile(const DevPlan &P, uint32_t tile, uint8_t *stage, uint64_t *bar) {
  // one elected thread: arm the barrier with the byte count, then one bulk copy per column part
  uint32_t total = 0;
  for (int c = 0; c < P.ncols; c++) {
    const DevColumn &col = P.cols[c];
    if (col.staged) total += col.tileValueBytes;
    if (col.hasNulls) total += col.tileNullBytes;
  }
  mbarExpectTx(bar, total);
  const size_t row0 = (size_t)tile * P.tileRows;
  for (int c = 0; c < P.ncols; c++) {
    const DevColumn &col = P.cols[c];
    if (col.staged) {
      const uint8_t *src = col.in.base + col.in.valuesOff + (col.width ? row0 * col.width : row0 / 8);
      tmaLoad1D(stage + col.smemValues, src, col.tileValueBytes, bar);
    }
    if (col.hasNulls) {
      const uint8_t *src = col.in.base + col.in.nullsOff + row0 / 8;
      tmaLoad1D(stage + col.smemNulls, src, col.tileNullBytes, bar);
    }
  }
}

template <bool WIDEKEY>
__global__ void __launch_bounds__(kFusedThreads, 1)
fusedBatchKernel(const __grid_constant__ DevPlan P, const DevTable G) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);                       // kStages barriers
  uint32_t *claims = reinterpret_cast<uint32_t *>(smem + 64);
  unsigned long long *tKeys = reinterpret_cast<unsigned long long *>(smem + 128);
  // accumulators of the CTA's table live in an L2-resident private slice of global memory:
  // fire-and-forget RED instead of shared-memory CAS loops, and 64 KB of shared memory back for the ring
  unsigned long long *tAcc = P.ctaAcc + (size_t)blockIdx.x * P.smemSlots;
  uint8_t *stages = reinterpret_cast<uint8_t *>(tKeys + P.smemSlots);

  SmemTable T;
  T.keys = tKeys; T.acc = tAcc; T.claims = claims; T.mask = P.smemSlots - 1;
  const bool useSmem = P.smemSlots > 0;
  for (uint32_t i = threadIdx.x; i < P.smemSlots; i += blockDim.x) {
    tKeys[i] = kEmptyKey;
    tAcc[i] = P.accNeutral;
  }
  if (threadIdx.x == 0) {
    *claims = 0;
    for (int s = 0; s < kMaxStages; s++) mbarInit(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t quadsPerTile = P.tileRows / R;
  // Growth of the group table (see jit_kernel_tail.cuh): this kernel stops at CTA granularity — thread 0 looks at the
  // STOP flag before every tile — and records the iterations it has folded; a resumed launch starts there.
  volatile uint32_t &sStop = *reinterpret_cast<volatile uint32_t *>(smem + 72);   // header word (no static shared memory)
  const uint32_t progIdx = blockIdx.x * kProgressWarps;
  uint32_t foldedUntil = 0xFFFFFFFFu;
  // ---- staged full tiles: tile t handled by CTA (t mod gridDim), ring of kStages buffers ----
  if (P.staged && P.numFullTiles > 0) {
    const uint32_t first = blockIdx.x, step = gridDim.x;
    const uint32_t startIt = P.resume ? G.progress[progIdx] : 0u;
    if (threadIdx.x == 0 && startIt != 0xFFFFFFFFu) {
      for (uint32_t s = 0; s < P.numStages; s++) {
        uint32_t t = first + (startIt + s) * step;
        if (t < P.numFullTiles) issueTile(P, t, stages + (size_t)s * P.stageBytes, &bars[s]);
      }
    }
    uint32_t it = 0, drainEnd = 0xFFFFFFFFu;
    bool draining = false;
    for (uint32_t t = first + startIt * step; startIt != 0xFFFFFFFFu && t < P.numFullTiles && it < drainEnd; t += step, it++) {
      const uint32_t s = it % P.numStages, parity = (it / P.numStages) & 1;
      if (!draining) {
        if (threadIdx.x == 0) sStop = *reinterpret_cast<volatile uint32_t *>(&G.counters[3]);
        __syncthreads();
        if (sStop != 0u && P.hll != 2) {   // stop folding; the tiles already in flight are still waited for
          draining = true;
          foldedUntil = startIt + it;
          drainEnd = it + P.numStages;
        }
      }
      mbarWait(&bars[s], parity);
      if (draining) continue;
      const uint8_t *stage = stages + (size_t)s * P.stageBytes;
      // shared-table admission is decided per tile (uniform in the CTA)
      const bool allowClaim = *reinterpret_cast<volatile uint32_t *>(claims) < (P.smemSlots / 4) * 3;
      for (uint32_t q = threadIdx.x; q < quadsPerTile; q += blockDim.x)
        processQuad<true, WIDEKEY>(P, G, T, useSmem, allowClaim, stage, q, t * P.tileRows + q * R, R);
      __syncthreads();  // everyone is done reading stage s
      if (threadIdx.x == 0) {
        uint32_t nt = t + P.numStages * step;
        if (nt < P.numFullTiles) issueTile(P, nt, stages + (size_t)s * P.stageBytes, &bars[s]);
      }
    }
  }
  if (threadIdx.x == 0) G.progress[progIdx] = foldedUntil;
  // ---- rows not covered by staged tiles (tail, or the whole batch on the direct path) --------
  // (every CTA folds its share of them once: progress[progIdx + 1] says so to a resumed launch; a launch that found the
  // table at its threshold leaves its share to the next one.  The host makes room for direct-path batches up front.)
  if (threadIdx.x == 0) {
    if (!P.resume) G.progress[progIdx + 1] = 0u;
    sStop = (P.hll != 2 && *reinterpret_cast<volatile uint32_t *>(&G.counters[3]) != 0u) || G.progress[progIdx + 1] != 0u;
    if (!sStop) G.progress[progIdx + 1] = 1u;
  }
  __syncthreads();
  if (!sStop) {
    const uint32_t begin = P.staged ? P.numFullTiles * P.tileRows : P.tailBegin;
    const uint32_t quads = (P.numRows - begin + R - 1) / R;
    const bool allowClaim = true;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
      uint32_t row0 = begin + q * R;
      uint32_t nrows = P.numRows - row0 < R ? P.numRows - row0 : R;
      processQuad<false, WIDEKEY>(P, G, T, useSmem, allowClaim, nullptr, q, row0, nrows);
    }
  }
  // ---- flush the shared table into the global one ---------------------------------------------
  __syncthreads();
  if (P.hll == 2) return;
  const AggOp op = (AggOp)P.aggOp;
  for (uint32_t i = threadIdx.x; i < P.smemSlots; i += blockDim.x) {
    unsigned long long k = tKeys[i];
    if (k != kEmptyKey) globalUpdate(G, op, k, nullptr, __ldcg(&tAcc[i]), /*spillWhenStopped=*/true);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *G.occPublish = *reinterpret_cast<volatile uint32_t *>(&G.counters[0]);
}

// ---------------------------------------------------------------------------------------
// merge of already-reduced rows, finalize
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mergeRowsKernel(const uint8_t *__restrict__ block, DimLayout L, const uint8_t *__restrict__ measures, int width,
                AggOp op, int n, uint8_t keyMode, uint8_t hashBits, int hll, DevTable G) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += stride) {
    uint64_t w[4];
    packRow(block, L, i, w);
    unsigned long long key = keyMode == KEY_PACKED ? w[0]
                           : (hashBits == 64 ? murmur3_128_lo(w, L.rowBytes, 0) : (unsigned long long)murmur3_32(w, L.rowBytes, 0));
    const uint64_t v = loadMeasure(measures, i, width);
    if (hll == 2) { hllDenseUpdate(G, nullptr, key, keyMode == KEY_HASHED ? w : nullptr, (uint32_t)v); continue; }
    if (hll == 1) key = (key & 0xFFFFFFFFFFFF0000ull) | (v & 0x3FFFu);
    globalUpdate(G, op, key, keyMode == KEY_HASHED ? w : nullptr, v);
  }
}

// Exchange step of a sharded query, receiving side: all N gathered parts ([groups, status, rows | dimension block of
// `L.capacity` rows | measures]) are folded by ONE launch; the row counts are read from the parts' headers on the
// device, so the host never waits for them.  A part whose sender could not fit its rows raises counters[2].
// `flags` != nullptr (exchange over peer memory, AggStateMergePartsWhenFlagged): the parts are written into this GPU's
// memory by the PEERS' export kernels, each of which then stores `epoch` into flags[its rank] with release semantics at
// system scope; every CTA waits for all numParts flags (acquire, system scope) before it reads a part.  The wait is
// bounded (a peer that never arrives: counters[2], reported by finalize), so the kernel cannot hang the GPU.
__global__ void __launch_bounds__(256)
mergePartsKernel(const uint8_t *parts, int numParts, size_t partStride, size_t dimOff, size_t valOff, DimLayout L,
                 int width, AggOp op, uint8_t keyMode, uint8_t hashBits, int hll, DevTable G, const uint32_t *flags, uint32_t epoch) {
  const uint32_t stride = gridDim.x * blockDim.x;
  if (flags != nullptr) {
    __shared__ uint32_t sLate;
    if (threadIdx.x == 0) {
      uint32_t late = 0;
      const long long t0 = clock64();
      for (int p = 0; p < numParts && !late; p++) {
        for (;;) {
          uint32_t seen;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + p) : "memory");
          if ((int32_t)(seen - epoch) >= 0) break;
          if (clock64() - t0 > 4000000000ll) { late = 1; break; }   // ~2 s at 2 GHz
          __nanosleep(100);
        }
      }
      sLate = late;
      if (late) atomicExch(&G.counters[2], 2u);
    }
    __syncthreads();
    if (sLate) return;
  }
  for (int p = 0; p < numParts; p++) {
    const uint8_t *part = parts + (size_t)p * partStride;
    const uint32_t *hdr = reinterpret_cast<const uint32_t *>(part);
    if (hdr[1] != SF_OK) {
      if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(&G.counters[2], 1u);
      continue;
    }
    const uint32_t n = hdr[0];
    const uint8_t *block = part + dimOff, *measures = part + valOff;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      uint64_t w[4];
      packRow(block, L, i, w);
      unsigned long long key = keyMode == KEY_PACKED ? w[0]
                             : (hashBits == 64 ? murmur3_128_lo(w, L.rowBytes, 0) : (unsigned long long)murmur3_32(w, L.rowBytes, 0));
      const uint64_t v = loadMeasure(measures, i, width);
      if (hll == 2) { hllDenseUpdate(G, nullptr, key, keyMode == KEY_HASHED ? w : nullptr, (uint32_t)v); continue; }
      if (hll == 1) key = (key & 0xFFFFFFFFFFFF0000ull) | (v & 0x3FFFu);
      globalUpdate(G, op, key, keyMode == KEY_HASHED ? w : nullptr, v);
    }
  }
}

// ---------------------------------------------------------------------------------------
// dense HLL mode: registers -> the carried (key, value) rows of query/hll.cu, already in key order
// ---------------------------------------------------------------------------------------
// one CTA per group (d-th in hash order): number of registers that were hit
__global__ void __launch_bounds__(256)
hllDenseCountKernel(const uint32_t *__restrict__ regs, const unsigned long long *__restrict__ acc, const uint32_t *__restrict__ slotOf, const uint32_t *__restrict__ order,
                    uint32_t *__restrict__ counts) {
  __shared__ uint32_t sWarp[256 / 32 + 1];
  const uint32_t *r = regs + (size_t)((uint32_t)acc[slotOf[order[blockIdx.x]]] - 1u) * kHllRegisters;   // (claim ordinal: hllRegArray)
  uint32_t c = 0;
  for (uint32_t i = threadIdx.x; i < kHllRegisters; i += 256) c += r[i] != 0;
  uint32_t total;
  blockExclusiveScan<256>(c, sWarp, &total);
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

// exclusive prefix of n <= 8192 counts by one CTA; offsets[n] = total
__global__ void __launch_bounds__(1024) hllDenseOffsetsKernel(const uint32_t *__restrict__ counts, int n, uint32_t *__restrict__ offsets) {
  __shared__ uint32_t sWarp[1024 / 32 + 1];
  __shared__ uint32_t sCarry;
  if (threadIdx.x == 0) sCarry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t x = i < n ? counts[i] : 0;
    uint32_t total;
    const uint32_t excl = blockExclusiveScan<1024>(x, sWarp, &total);
    if (i < n) offsets[i] = sCarry + excl;
    __syncthreads();
    if (threadIdx.x == 0) sCarry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = sCarry;
}

// one CTA per group: its hit registers in ascending register order -> (key, value, group ordinal)
__global__ void __launch_bounds__(256)
hllDenseEmitKernel(const uint32_t *__restrict__ regs, const unsigned long long *__restrict__ acc, const uint32_t *__restrict__ slotOf, const uint32_t *__restrict__ order,
                   const uint64_t *__restrict__ sortedHash, const uint32_t *__restrict__ offsets,
                   uint64_t *__restrict__ outHash, uint32_t *__restrict__ outVals, uint32_t *__restrict__ outIndex) {
  __shared__ uint32_t sWarp[256 / 32 + 1];
  const uint32_t d = blockIdx.x;
  const uint32_t *r = regs + (size_t)((uint32_t)acc[slotOf[order[d]]] - 1u) * kHllRegisters;
  const uint64_t hi = sortedHash[d] & 0xFFFFFFFFFFFF0000ull;
  uint32_t pos = offsets[d];
  for (uint32_t base = 0; base < kHllRegisters; base += 256) {
    const uint32_t reg = base + threadIdx.x;
    const uint32_t v = r[reg];
    uint32_t total;
    const uint32_t excl = blockExclusiveScan<256>(v != 0, sWarp, &total);
    if (v != 0) {
      outHash[pos + excl] = hi | reg;
      outVals[pos + excl] = v - 1u;
      outIndex[pos + excl] = d;
    }
    pos += total;
    __syncthreads();
  }
}

// Dense registers -> the reference's register vectors directly (AggStateFinalizeHLL; query/hll.cu:262-290 semantics without
// the detour through (key, value) entries: 13M entries = 212 MB for 808 groups).  Layout pass, one CTA: per group d in hash
// order with counts[d] hit registers — its output ordinal among the groups that have any, the byte offset of its vector
// (4 bytes per hit register below HLL_DENSE_THRESHOLD, else HLL_DENSE_SIZE), totals[0] = dims, totals[1] = bytes.
__global__ void __launch_bounds__(1024)
hllVectorLayoutKernel(const uint32_t *__restrict__ counts, int n, uint32_t *__restrict__ outIdx, uint32_t *__restrict__ byteOff,
                      uint32_t *__restrict__ rowsOf, unsigned long long *__restrict__ totals) {
  __shared__ uint32_t sWarp[1024 / 32 + 1];
  __shared__ uint32_t sCarryP, sCarryB;
  if (threadIdx.x == 0) { sCarryP = 0; sCarryB = 0; }
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t c = i < n ? counts[i] : 0;
    const uint32_t p = c != 0, sz = c == 0 ? 0u : (c < (uint32_t)HLL_DENSE_THRESHOLD ? c * 4u : (uint32_t)HLL_DENSE_SIZE);
    uint32_t totP, totB;
    const uint32_t exP = blockExclusiveScan<1024>(p, sWarp, &totP);
    __syncthreads();
    const uint32_t exB = blockExclusiveScan<1024>(sz, sWarp, &totB);
    if (i < n) {
      outIdx[i] = sCarryP + exP;
      byteOff[i] = sCarryB + exB;
      if (p) rowsOf[sCarryP + exP] = (uint32_t)i;
    }
    __syncthreads();
    if (threadIdx.x == 0) { sCarryP += totP; sCarryB += totB; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = sCarryP; totals[1] = sCarryB; }
}

// one CTA per group: its vector (sparse: (rho << 16 | register) entries in ascending register order; dense: one rho byte
// per register) and its register count
__global__ void __launch_bounds__(256)
hllVectorEmitKernel(const uint32_t *__restrict__ regs, const unsigned long long *__restrict__ acc, const uint32_t *__restrict__ slotOf,
                    const uint32_t *__restrict__ order, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ outIdx,
                    const uint32_t *__restrict__ byteOff, uint8_t *__restrict__ vec, uint16_t *__restrict__ regCount) {
  __shared__ uint32_t sWarp[256 / 32 + 1];
  const uint32_t d = blockIdx.x, c = counts[d];
  if (c == 0) return;
  if (threadIdx.x == 0) regCount[outIdx[d]] = (uint16_t)c;
  const uint32_t *r = regs + (size_t)((uint32_t)acc[slotOf[order[d]]] - 1u) * kHllRegisters;
  uint8_t *dst = vec + byteOff[d];
  // a register holds value + 1, value = rho << 16 | register (0: never hit); the vectors carry rho + 1
  if (c < (uint32_t)HLL_DENSE_THRESHOLD) {
    uint32_t pos = 0;
    for (uint32_t base = 0; base < kHllRegisters; base += 256) {
      const uint32_t reg = base + threadIdx.x, v = r[reg];
      uint32_t total;
      const uint32_t excl = blockExclusiveScan<256>(v != 0, sWarp, &total);
      if (v != 0) reinterpret_cast<uint32_t *>(dst)[pos + excl] = (((((v - 1u) >> 16) & 0xFFu) + 1u) << 16) | reg;
      pos += total;
      __syncthreads();
    }
  } else {
    for (uint32_t q = threadIdx.x; q < kHllRegisters / 4; q += 256) {
      const uint4 v = reinterpret_cast<const uint4 *>(r)[q];
      auto rho = [](uint32_t x) { return x ? ((((x - 1u) >> 16) & 0xFFu) + 1u) : 0u; };
      reinterpret_cast<uint32_t *>(dst)[q] = rho(v.x) | (rho(v.y) << 8) | (rho(v.z) << 16) | (rho(v.w) << 24);
    }
  }
}


constexpr int kCmpThreads = 256;
constexpr int kCmpItems = 8;
constexpr int kCmpTile = kCmpThreads * kCmpItems;

// Compacts occupied slots: slotOf[g] = slot index, hash[g] = reference hash of the group's row,
// vals[g] = accumulator (tight `width`-byte elements).
__global__ void __launch_bounds__(kCmpThreads)
compactGroupsKernel(DevTable G, size_t cap, uint8_t keyMode, uint8_t hashBits, uint64_t hashMask, int rowBytes, int width, ScanTileState st,
                    uint32_t *__restrict__ slotOf, uint64_t *__restrict__ hash, uint8_t *__restrict__ vals,
                    uint32_t *__restrict__ outCount) {
  __shared__ uint32_t sTile, sPrefix;
  __shared__ uint32_t sWarp[kCmpThreads / 32 + 1];
  if (threadIdx.x == 0) sTile = atomicAdd(st.ticket, 1u);
  __syncthreads();
  const uint32_t tile = sTile;
  const size_t base = (size_t)tile * kCmpTile + (size_t)threadIdx.x * kCmpItems;
  uint32_t mask = 0;
#pragma unroll
  for (int k = 0; k < kCmpItems; k++) {
    size_t i = base + k;
    if (i < cap && G.keys[i] != kEmptyKey) mask |= 1u << k;
  }
  uint32_t blockTotal;
  const uint32_t excl = blockExclusiveScan<kCmpThreads>(__popc(mask), sWarp, &blockTotal);
  if (threadIdx.x < 32) {
    uint32_t p = decoupledLookback(st, tile, blockTotal);
    if (threadIdx.x == 0) {
      sPrefix = p;
      if (((size_t)tile + 1) * kCmpTile >= cap) *outCount = p + blockTotal;
    }
  }
  __syncthreads();
  uint32_t pos = sPrefix + excl;
#pragma unroll
  for (int k = 0; k < kCmpItems; k++) {
    if (!(mask & (1u << k))) continue;
    size_t i = base + k;
    unsigned long long key = G.keys[i];
    uint64_t h;
    if (keyMode == KEY_PACKED) {
      uint64_t w[4] = {key, 0, 0, 0};
      h = hashBits == 64 ? murmur3_128_lo(w, rowBytes, 0) : (uint64_t)murmur3_32(w, rowBytes, 0);
    } else {
      h = key;
    }
    slotOf[pos] = (uint32_t)i;
    hash[pos] = hashBits == 64 ? h & hashMask : h;
    storeMeasure(vals, pos, width, G.acc[i]);
    pos++;
  }
}

// Large results: the claim list replaces the table scan — slotOf / hash / vals of the n claimed groups, claim order.
__global__ void __launch_bounds__(256)
gatherClaimedKernel(DevTable G, uint32_t n, uint8_t keyMode, uint8_t hashBits, uint64_t hashMask, int rowBytes, int width,
                    uint32_t *__restrict__ slotOf, uint64_t *__restrict__ hash, uint8_t *__restrict__ vals) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t slot = G.claimed[i];
    const unsigned long long key = G.keys[slot];
    uint64_t h;
    if (keyMode == KEY_PACKED) {
      uint64_t w[4] = {key, 0, 0, 0};
      h = hashBits == 64 ? murmur3_128_lo(w, rowBytes, 0) : (uint64_t)murmur3_32(w, rowBytes, 0);
    } else {
      h = key;
    }
    slotOf[i] = slot;
    hash[i] = hashBits == 64 ? h & hashMask : h;
    storeMeasure(vals, i, width, G.acc[slot]);
  }
}

__global__ void __launch_bounds__(256)
emitGroupsKernel(DevTable G, uint8_t keyMode, const uint32_t *__restrict__ slotOf, const uint32_t *__restrict__ repIndex,
                 uint32_t g, uint8_t *__restrict__ outBlock, DimLayout L, uint32_t *__restrict__ outIndex) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < g; s += stride) {
    const uint32_t slot = slotOf[repIndex[s]];
    uint64_t w[4];
    if (keyMode == KEY_PACKED) {
      w[0] = G.keys[slot]; w[1] = w[2] = w[3] = 0;
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) w[i] = G.rows[(size_t)slot * 4 + i];
    }
    unpackRow(outBlock, L, s, w);
    if (outIndex) outIndex[s] = s;
  }
}

__global__ void iotaKernel(uint32_t *p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
fillTableKernel(unsigned long long *keys, unsigned long long *acc, size_t cap, unsigned long long neutral) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
    keys[i] = kEmptyKey;
    acc[i] = neutral;
  }
}

// out[i] = in[i] + (+0.0): float sums of the hash-reduce mode (see finalize())
__global__ void __launch_bounds__(256) copyPlusZeroKernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int n, int width) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (width == 8) reinterpret_cast<double *>(out)[i] = reinterpret_cast<const double *>(in)[i] + 0.0;
  else reinterpret_cast<float *>(out)[i] = reinterpret_cast<const float *>(in)[i] + 0.0f;
}

// Global dense slots (DevPlan::denseGlobal): after the batch, every reached slot of the state's accumulator array is
// folded into the group table under the packed dimension row its index decodes to, and reset to the neutral element.
struct DenseFold {
  uint32_t lo[8], cnt[8], step[8], stride[8];
  uint8_t rowOff[8], width[8], nullOff[8];
  uint32_t nd, total, reps;
  uint8_t keyMode, hashBits, rowBytes, op;
  unsigned long long neutral;
};

__global__ void __launch_bounds__(256)
denseFoldKernel(unsigned long long *__restrict__ acc, DenseFold F, DevTable G) {
  const uint32_t strideAll = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < F.total; i += strideAll) {
    unsigned long long v = F.neutral;
    for (uint32_t c = 0; c < F.reps; c++) {   // the copies the CTA groups accumulated into
      const unsigned long long x = acc[(size_t)c * F.total + i];
      if (x == F.neutral) continue;
      acc[(size_t)c * F.total + i] = F.neutral;
      v = v == F.neutral ? x : aggCombine((AggOp)F.op, v, x);
    }
    if (v == F.neutral) continue;
    uint64_t row[4] = {0, 0, 0, 0};
    uint8_t *rb = reinterpret_cast<uint8_t *>(row);
    uint32_t rem = i;
    for (int k = (int)F.nd - 1; k >= 0; k--) {
      const uint32_t ix = rem / F.stride[k];
      rem -= ix * F.stride[k];
      const bool valid = ix != F.cnt[k];
      const uint32_t val = valid ? (F.lo[k] + ix) * F.step[k] : 0u;
      for (int b = 0; b < F.width[k]; b++) rb[F.rowOff[k] + b] = (uint8_t)(val >> (8 * b));
      rb[F.nullOff[k]] = valid ? 1 : 0;
    }
    unsigned long long key;
    if (F.keyMode == KEY_PACKED) key = row[0];
    else key = F.hashBits == 64 ? murmur3_128_lo(row, F.rowBytes, 0) : (unsigned long long)murmur3_32(row, F.rowBytes, 0);
    globalUpdate(G, (AggOp)F.op, key, F.keyMode == KEY_PACKED ? nullptr : row, v);
  }
}

// ---------------------------------------------------------------------------------------
// single-launch finalize for results of up to kSmallFinalizeMax groups
// ---------------------------------------------------------------------------------------
// One thread-block cluster of kFinCtas CTAs walks the claim list (no table scan, no compaction), hashes each group's
// packed row with the reference's murmur3, sorts the (hash, claim ordinal) pairs — one bucketing pass over the top 12 hash
// bits (histogram in L2, every CTA scans it into its own shared memory), then every element ranks itself inside its
// bucket —, merges runs of equal hashes with the aggregate's rule — the member claimed first names the run — and writes
// the reference's DimensionVector block + measure vector.  The phases are chains of two or three dependent loads per
// element: 8192 threads give every thread at most four elements, so a phase costs a few memory latencies, and the
// phases are separated by the hardware cluster barrier (one CTA of 1024 threads took 225 us for 19,200 groups, all of
// it serialised latency: profiles/r02_launches_cfg3.csv).  The group count goes to a mapped pinned host word, so the
// host's only interaction is one stream synchronise.  `ordered == 0` is the exchange form (AggStateExport /
// AggStateExportPart): the claimed slots as they are, no sort, no merge.
constexpr int kSmallFinalizeMax = kSmallSortMax;
constexpr int kFinCtas = 8;                          // portable cluster maximum
constexpr uint32_t kFinThreads = kFinCtas * 1024u;
constexpr size_t kFinHistWords = 2 * kSmallBuckets + 64;   // bucket counts | scatter cursors | per-CTA run-head totals

struct SmallFinalizeArgs {
  DevTable G;
  DimLayout L;              // output block layout (capacity = outputKeys.VectorCapacity)
  uint64_t hashMask;
  uint64_t *hashA, *tmpK;   // scratch: kSmallFinalizeMax entries each
  uint32_t *idxA, *tmpI;
  uint32_t *hist;           // scratch: kFinHistWords
  uint8_t *outBlock, *outValues;
  uint64_t *outHash;
  uint32_t *outIndex;
  uint32_t *resultDev;      // [0] groups, [1] status, [2] claimed slots
  volatile uint32_t *resultHost;
  int32_t rowBytes, width, outCapacity;
  uint8_t keyMode, hashBits, op, plusZero, ordered;
};

__global__ void __cluster_dims__(kFinCtas, 1, 1) __launch_bounds__(1024) finalizeSmallKernel(const __grid_constant__ SmallFinalizeArgs A) {
  __shared__ uint32_t off[kSmallBuckets + 1];
  __shared__ uint32_t sWarp[1024 / 32 + 1];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t ctaRank = cluster.block_rank(), gtid = ctaRank * 1024u + threadIdx.x;
  const DevTable &G = A.G;
  const uint32_t n = G.counters[0];
  uint32_t status = SF_OK;
  if (G.counters[1]) status = SF_TABLE_OVERFLOW;
  else if (G.counters[2] == 2u) status = SF_PEER_LATE;   // AggStateMergePartsWhenFlagged gave up waiting for a peer's part
  else if (G.counters[2]) status = SF_PART_TRUNCATED;   // AggStateMergeParts met a part that did not hold all its rows
  else if (G.counters[3] || G.counters[4]) status = SF_UNSETTLED;   // stopped at the growth threshold / rows parked: the host settles first
  else if (n > (uint32_t)kSmallFinalizeMax) status = SF_TOO_MANY;
  else if (!A.ordered && n > (uint32_t)A.outCapacity) status = SF_OUTPUT_TOO_SMALL;
  auto publish = [&](uint32_t groups, uint32_t st, uint32_t claimedSlots) {
    A.resultDev[0] = groups; A.resultDev[1] = st; A.resultDev[2] = claimedSlots;
    A.resultHost[0] = groups; A.resultHost[1] = st; A.resultHost[2] = claimedSlots;
  };
  if (status != SF_OK || n == 0) {   // (uniform over the cluster: nobody waits at a barrier below)
    if (gtid == 0) publish(0, status, n);
    return;
  }
  auto rowOf = [&](uint32_t slot, uint64_t (&w)[4]) {
    if (A.keyMode == KEY_PACKED) { w[0] = G.keys[slot]; w[1] = w[2] = w[3] = 0; }
    else {
#pragma unroll
      for (int i = 0; i < 4; i++) w[i] = G.rows[(size_t)slot * 4 + i];
    }
  };
  if (!A.ordered) {
    for (uint32_t i = gtid; i < n; i += kFinThreads) {
      const uint32_t slot = G.claimed[i];
      uint64_t w[4];
      rowOf(slot, w);
      unpackRow(A.outBlock, A.L, i, w);
      storeMeasure(A.outValues, i, A.width, G.acc[slot]);
      if (A.outIndex) A.outIndex[i] = i;
    }
    if (gtid == 0) publish(n, SF_OK, n);
    return;
  }
  uint32_t *cnt = A.hist, *cur = A.hist + kSmallBuckets, *ctaHeads = A.hist + 2 * kSmallBuckets;
  const int shift = A.hashBits - kSmallBits;
  for (uint32_t b = gtid; b < (uint32_t)kSmallBuckets; b += kFinThreads) { cnt[b] = 0; cur[b] = 0; }
  cluster.sync();
  // 1. reference hash of every group's row; histogram of the top hash bits
  for (uint32_t i = gtid; i < n; i += kFinThreads) {
    const unsigned long long key = G.keys[G.claimed[i]];
    uint64_t h;
    if (A.keyMode == KEY_PACKED) {
      uint64_t w[4] = {key, 0, 0, 0};
      h = A.hashBits == 64 ? murmur3_128_lo(w, A.rowBytes, 0) & A.hashMask : (uint64_t)murmur3_32(w, A.rowBytes, 0);
    } else {
      h = A.hashBits == 64 ? key & A.hashMask : key;
    }
    A.hashA[i] = h;
    atomicAdd(&cnt[(uint32_t)(h >> shift) & (kSmallBuckets - 1)], 1u);
  }
  cluster.sync();
  // 2. bucket offsets: every CTA scans the histogram for itself (no barrier, and the offsets are in shared memory)
  {
    constexpr int kPer = kSmallBuckets / 1024;
    uint32_t c[kPer], sum = 0, total;
#pragma unroll
    for (int j = 0; j < kPer; j++) { c[j] = cnt[threadIdx.x * kPer + j]; sum += c[j]; }
    uint32_t excl = blockExclusiveScan<1024>(sum, sWarp, &total);
#pragma unroll
    for (int j = 0; j < kPer; j++) { off[threadIdx.x * kPer + j] = excl; excl += c[j]; }
    if (threadIdx.x == 0) off[kSmallBuckets] = total;
  }
  __syncthreads();
  // 3. unordered scatter into the buckets
  for (uint32_t i = gtid; i < n; i += kFinThreads) {
    const uint64_t h = A.hashA[i];
    const uint32_t d = (uint32_t)(h >> shift) & (kSmallBuckets - 1);
    const uint32_t p = off[d] + atomicAdd(&cur[d], 1u);
    A.tmpK[p] = h;
    A.tmpI[p] = i;
  }
  cluster.sync();
  // 4. every element finds its rank among its bucket's members by (hash, claim ordinal): 4.7 members on average for
  // 19,200 groups; a degenerate bucket only costs time
  for (uint32_t i = gtid; i < n; i += kFinThreads) {
    const uint64_t k = A.tmpK[i];
    const uint32_t v = A.tmpI[i];
    const uint32_t d = (uint32_t)(k >> shift) & (kSmallBuckets - 1);
    const uint32_t lo = off[d], hi = off[d + 1];
    uint32_t rank = 0;
    for (uint32_t j = lo; j < hi; j++) {
      const uint64_t kj = A.tmpK[j];
      rank += (kj < k) || (kj == k && A.tmpI[j] < v);
    }
    A.hashA[lo + rank] = k;
    A.idxA[lo + rank] = v;
  }
  cluster.sync();
  // 5. runs of equal hashes: every thread owns a contiguous chunk, counts the run heads in it ...
  const uint32_t per = (n + kFinThreads - 1) / kFinThreads;
  const uint32_t begin = gtid * per < n ? gtid * per : n, end = begin + per < n ? begin + per : n;
  uint32_t heads = 0;
  for (uint32_t j = begin; j < end; j++) heads += j == 0 || A.hashA[j] != A.hashA[j - 1];
  uint32_t ctaTotal;
  uint32_t pos = blockExclusiveScan<1024>(heads, sWarp, &ctaTotal);
  if (threadIdx.x == 0) ctaHeads[ctaRank] = ctaTotal;
  cluster.sync();
  uint32_t g = 0;
#pragma unroll
  for (int r = 0; r < kFinCtas; r++) {
    const uint32_t t = ctaHeads[r];
    if ((uint32_t)r < ctaRank) pos += t;
    g += t;
  }
  if (g > (uint32_t)A.outCapacity) {
    if (gtid == 0) publish(0, SF_OUTPUT_TOO_SMALL, g);
    return;
  }
  // ... 6. and folds + emits the runs that start in its chunk (a run may extend into the next chunks)
  for (uint32_t j = begin; j < end; j++) {
    const uint64_t h = A.hashA[j];
    if (!(j == 0 || h != A.hashA[j - 1])) continue;
    const uint32_t slot = G.claimed[A.idxA[j]];
    uint64_t acc = G.acc[slot];
    for (uint32_t k = j + 1; k < n && A.hashA[k] == h; k++) acc = aggCombine((AggOp)A.op, acc, G.acc[G.claimed[A.idxA[k]]]);
    if (A.plusZero) {   // hash-reduce float sums start from +0.0 (see finalize())
      if (A.width == 8) acc = (uint64_t)__double_as_longlong(__longlong_as_double((long long)acc) + 0.0);
      else acc = __float_as_uint(__uint_as_float((uint32_t)acc) + 0.0f);
    }
    uint64_t w[4];
    rowOf(slot, w);
    unpackRow(A.outBlock, A.L, pos, w);
    storeMeasure(A.outValues, pos, A.width, acc);
    if (A.outHash) A.outHash[pos] = h;
    if (A.outIndex) A.outIndex[pos] = pos;
    pos++;
  }
  if (gtid == 0) publish(g, SF_OK, n);
}

// Exchange over peer memory, sending side: ONE kernel (a cluster of kFinCtas CTAs) writes this rank's rows as a part
// ([rows, status, claimed | dimension block | measures], the layout of AggStateExportPart) into its own receive buffer,
// copies the part into the same slot of every peer's receive buffer with 16-byte stores over NVLink, and then stores
// `epoch` into flags[myRank] on every peer (release, system scope) — export, all-gather and the "it has arrived" signal
// in one launch, no collective library and no host in between.
struct PeerExportArgs {
  SmallFinalizeArgs F;            // the export itself (ordered = 0); F.outBlock / F.outValues / F.resultDev lie in the local slot
  uint8_t *peerSlot[16];          // this rank's slot in every peer's receive buffer (peerSlot[myRank]: the local one)
  uint32_t *peerFlag[16];         // &flags[myRank] on every peer
  size_t partBytes;               // multiple of 16
  uint32_t numPeers, myRank, epoch;
};

__global__ void __cluster_dims__(kFinCtas, 1, 1) __launch_bounds__(1024) exportToPeersKernel(const __grid_constant__ PeerExportArgs E) {
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t gtid = cluster.block_rank() * 1024u + threadIdx.x;
  const SmallFinalizeArgs &A = E.F;
  const DevTable &G = A.G;
  const uint32_t n = G.counters[0];
  uint32_t status = SF_OK;
  if (G.counters[1]) status = SF_TABLE_OVERFLOW;
  else if (G.counters[3] || G.counters[4]) status = SF_UNSETTLED;
  else if (n > (uint32_t)A.outCapacity) status = SF_OUTPUT_TOO_SMALL;
  if (status == SF_OK) {
    for (uint32_t i = gtid; i < n; i += kFinThreads) {
      const uint32_t slot = G.claimed[i];
      uint64_t w[4];
      if (A.keyMode == KEY_PACKED) { w[0] = G.keys[slot]; w[1] = w[2] = w[3] = 0; }
      else {
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = G.rows[(size_t)slot * 4 + k];
      }
      unpackRow(A.outBlock, A.L, i, w);
      storeMeasure(A.outValues, i, A.width, G.acc[slot]);
    }
  }
  if (gtid == 0) { A.resultDev[0] = status == SF_OK ? n : 0u; A.resultDev[1] = status; A.resultDev[2] = n; }
  cluster.sync();   // the local part is complete (and visible to the cluster)
  // the part travels as it lies: header, the used prefix of every section would save bytes, but a part is 0.5 MB and the
  // copy is a few microseconds of NVLink time
  const uint4 *src = reinterpret_cast<const uint4 *>(E.peerSlot[E.myRank]);
  const uint32_t words = (uint32_t)(E.partBytes / 16);
  for (uint32_t p = 0; p < E.numPeers; p++) {
    if (p == E.myRank) continue;
    uint4 *dst = reinterpret_cast<uint4 *>(E.peerSlot[p]);
    for (uint32_t i = gtid; i < words; i += kFinThreads) dst[i] = src[i];
  }
  __threadfence_system();
  cluster.sync();   // every thread's stores are ordered before the flags
  if (gtid < E.numPeers) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(E.peerFlag[gtid]), "r"(E.epoch) : "memory");
}

// AggStateReset: only the claimed slots are emptied (and, for dense HLL states, only their register arrays).
__global__ void __launch_bounds__(256)
resetClaimedKernel(DevTable G, unsigned long long neutral) {
  const uint32_t n = G.counters[0];
  if (G.regs != nullptr) {   // one CTA per claimed group at a time: 16384 registers
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
      if (i >= kHllDenseMaxGroups) break;   // (claims past the register arrays were turned away)
      uint4 *r = reinterpret_cast<uint4 *>(G.regs + (size_t)i * kHllRegisters);   // register arrays go by claim ordinal
      for (uint32_t k = threadIdx.x; k < kHllRegisters / 4; k += blockDim.x) r[k] = make_uint4(0, 0, 0, 0);
    }
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t slot = G.claimed[i];
    G.keys[slot] = kEmptyKey;
    G.acc[slot] = neutral;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct AggState {
  AggSpec spec;
  int device;
  DimLayout rowLayout;     // capacity-independent parts (rowOff / width / numDims)
  KeyMode keyMode;
  int hashBits;
  AggOp op;
  int measWidth;
  ValClass measClass;
  uint64_t accNeutral;
  bool hll;
  bool hllDense;           // HLL with one dense register array per group (few groups) instead of (group, register) entries
  size_t capacity;
  DevTable table;
  void *mem;               // single allocation behind the table
  unsigned long long *ctaAcc;  // [kMaxGridCtas][8192] private accumulator slices of the fused kernel's CTAs
  unsigned long long *denseAcc = nullptr;  // [kGlobalDenseMaxSlots] shared accumulators of the global dense form (lazy)
  uint64_t occUpper = 0;                   // host-side upper bound of the occupied slots (exact after a synchronise)
  uint32_t unchecked = 0;                  // hash-table batches launched since the last check of the STOP flag
  bool everChecked = false;                // a batch of this query has been waited for (its occupancy is known)
  uint8_t *smallScratch = nullptr;         // single-launch finalize: hash / index ping-pong arrays (in `mem`)
  uint32_t *resultDev = nullptr;           // [0] groups, [1] status, [2] claimed slots of the last single-launch finalize
  uint32_t *resultHost = nullptr;          // the same three words in mapped pinned host memory
  uint32_t *resultHostDev = nullptr;       // device alias of resultHost
};

constexpr size_t kHllDenseSlots = 8192;

static uint64_t neutralOf(AggOp op) {
  switch (op) {
    case OP_SUM_F32: return 0x80000000ull;                // -0.0f: (-0) + x == x for every x
    case OP_SUM_F64: return 0x8000000000000000ull;        // -0.0
    case OP_MIN_U32: return 0xFFFFFFFFull;
    case OP_MIN_I32: return 0x7FFFFFFFull;
    case OP_MAX_I32: return 0x80000000ull;
    case OP_MIN_F32: return 0x7F800000ull;                // +inf
    case OP_MAX_F32: return 0xFF800000ull;                // -inf
    default: return 0;                                     // integer sums, unsigned max
  }
}

static ValClass measureClassOf(int dt) {
  switch (dt) {
    case Int32: return VC_I32;
    case Uint32: return VC_U32;
    case Float32: return VC_F32;
    case Int64: return VC_I64;
    case Float64: return VC_F64;
    default: throw EngineError("Unsupported data type for MeasureOutput");
  }
}

static void allocTable(AggState *st, size_t cap, cudaStream_t s) {
  const bool rows = st->keyMode == KEY_HASHED;
  const size_t ctaAccBytes = (size_t)kMaxGridCtas * 8192 * sizeof(unsigned long long);
  const size_t regBytes = st->hllDense ? (size_t)kHllDenseMaxGroups * kHllRegisters * sizeof(uint32_t) : 0;   // by claim ordinal
  const size_t claimedBytes = (cap * 4 + 255) / 256 * 256;
  const size_t smallBytes = (size_t)kSmallFinalizeMax * (8 + 8 + 4 + 4) + kFinHistWords * 4;
  const size_t progressBytes = ((size_t)(kProgressTail + 1) * 4 + 255) / 256 * 256;
  const size_t spillBytes = (size_t)kSpillCap * sizeof(SpillEntry);
  size_t bytes = cap * 16 + (rows ? cap * 32 : 0) + 256 + ctaAccBytes + regBytes + claimedBytes + smallBytes + progressBytes + spillBytes;
  void *mem = nullptr;
  CGoCallResHandle h = deviceMalloc(&mem, bytes);
  if (h.pStrErr) { std::string m(h.pStrErr); free((void *)h.pStrErr); throw EngineError(m); }
  st->mem = mem;
  st->capacity = cap;
  uint8_t *p = static_cast<uint8_t *>(mem);
  st->table.counters = reinterpret_cast<uint32_t *>(p);
  st->table.keys = reinterpret_cast<unsigned long long *>(p + 256);
  st->table.acc = st->table.keys + cap;
  st->table.rows = rows ? reinterpret_cast<uint64_t *>(st->table.acc + cap) : nullptr;
  st->ctaAcc = reinterpret_cast<unsigned long long *>(p + 256 + cap * 16 + (rows ? cap * 32 : 0));
  st->table.mask = (uint32_t)(cap - 1);
  st->table.regs = st->hllDense ? reinterpret_cast<uint32_t *>(p + 256 + cap * 16 + (rows ? cap * 32 : 0) + ctaAccBytes) : nullptr;
  uint8_t *tail = p + 256 + cap * 16 + (rows ? cap * 32 : 0) + ctaAccBytes + regBytes;
  st->table.claimed = reinterpret_cast<uint32_t *>(tail);
  st->smallScratch = tail + claimedBytes;
  st->table.progress = reinterpret_cast<uint32_t *>(tail + claimedBytes + smallBytes);
  st->table.spill = reinterpret_cast<SpillEntry *>(tail + claimedBytes + smallBytes + progressBytes);
  // half full = time to grow (the dense HLL directory does not grow: its register arrays are sized with it)
  st->table.growAt = st->hllDense ? 0xFFFFFFFFu : (uint32_t)(cap / 2);
  st->resultDev = st->table.counters + 8;   // inside the 256-byte header
  if (!st->resultHost) {
    ARES_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&st->resultHost), 64, cudaHostAllocMapped));
    ARES_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&st->resultHostDev), st->resultHost, 0));
    memset(st->resultHost, 0, 64);
  }
  st->table.occPublish = st->resultHostDev + 8;
  if (st->hllDense) ARES_CUDA(cudaMemsetAsync(st->table.regs, 0, regBytes, s));
  ARES_CUDA(cudaMemsetAsync(p, 0, 256, s));
  fillTableKernel<<<smCount() * 8, 256, 0, s>>>(st->table.keys, st->table.acc, cap, st->accNeutral);
  checkLastError("fillTable");
}

static void describeState(AggState *st, const AggSpec &spec) {
  st->spec = spec;
  st->rowLayout = makeDimLayout(spec.NumDimsPerDimWidth, 1);
  if (spec.ReduceMode != ARES_REDUCE_SORT && spec.ReduceMode != ARES_REDUCE_HASH) throw EngineError("unknown ReduceMode");
  st->hashBits = spec.ReduceMode == ARES_REDUCE_SORT ? 64 : 32;
  st->keyMode = st->rowLayout.rowBytes <= 8 ? KEY_PACKED : KEY_HASHED;
  st->measClass = measureClassOf(spec.MeasureDataType);
  int bytes = (st->measClass == VC_I64 || st->measClass == VC_F64) ? 8 : 4;
  if (spec.AggFunc == AGGR_AVG_FLOAT && bytes != 8) throw EngineError("an AVG measure is 8 bytes (average, count)");
  st->hll = spec.AggFunc == AGGR_HLL;
  st->hllDense = st->hll && spec.ExpectedGroups <= kHllDenseMaxGroups;
  if (st->hll) {
    if (st->measClass != VC_U32) throw EngineError("an HLL measure is Uint32");
    st->hashBits = 64;
    st->op = OP_MAX_U32;
    st->measWidth = 4;
    // entry mode: group identity = the reference's HLL key (dim-row hash with the register in its low
    // 16 bits, query/functor.hpp:1299-1305), the table keeps the max value (rho << 16 | reg) per key.
    // dense mode (ExpectedGroups <= 4096 groups): the table is the directory of dimension rows and
    // every group owns 16384 registers in DevTable::regs.
    if (!st->hllDense) st->keyMode = KEY_HASHED;
  } else {
    st->op = aggOpOf(spec.AggFunc, bytes, &st->measWidth);
  }
  st->accNeutral = neutralOf(st->op);
}

static AggState *createState(const AggSpec &spec, cudaStream_t s, int device) {
  AggState *st = new AggState();
  try {
    describeState(st, spec);
    st->device = device;
    // global table: 2^21 slots (32 MB, L2-resident) unless the caller expects more groups
    size_t want = (size_t)spec.ExpectedGroups * 2;
    size_t cap = (size_t)1 << 21;
    while (cap < want) cap <<= 1;
    if (const char *e = getenv("ARESDB_B200_TABLE_SLOTS")) {   // tests: start small so that the table has to grow
      const long v = atol(e);
      if (v >= 1024 && (v & (v - 1)) == 0) cap = (size_t)v;
    }
    if (st->hllDense) cap = kHllDenseSlots;
    allocTable(st, cap, s);
  } catch (...) {
    delete st;
    throw;
  }
  return st;
}

static AggState *asState(void *p) {
  if (!p) throw EngineError("null AggState handle");
  return static_cast<AggState *>(p);
}

static uint8_t operandClassOf(const PlanOperand &o, const BatchPlan &bp, const std::vector<uint8_t> &stackClasses) {
  switch (o.Kind) {
    case PLAN_OPERAND_COLUMN: {
      if (o.Column >= bp.NumColumns) throw EngineError("plan operand references a column outside BatchPlan.Columns");
      switch (bp.Columns[o.Column].DataType) {
        case Bool: return VC_BOOL;
        case Int8: case Int16: case Int32: return VC_I32;
        case Uint8: case Uint16: case Uint32: return VC_U32;
        case Float32: return VC_F32;
        case Int64: return VC_I64;
        case UUID: return VC_UUID;
        default: throw EngineError("Unsupported data type for VectorPartyInput");
      }
    }
    case PLAN_OPERAND_CONST:
      if (o.ConstType == ConstInt) return VC_I32;
      if (o.ConstType == ConstFloat) return VC_F32;
      throw EngineError("Unsupported constant type in plan");
    case PLAN_OPERAND_STACK:
      if (stackClasses.empty()) throw EngineError("plan pops an empty evaluation stack");
      return stackClasses.back();
    case PLAN_OPERAND_FOREIGN: {
      if (o.Column >= bp.NumForeignColumns) throw EngineError("plan operand references a column outside BatchPlan.ForeignColumns");
      switch (bp.ForeignColumns[o.Column].Column.DataType) {
        case Bool: return VC_BOOL;
        case Int8: case Int16: case Int32: return VC_I32;
        case Uint8: case Uint16: case Uint32: return VC_U32;
        case Float32: return VC_F32;
        default: throw EngineError("foreign columns of the fused path are Bool / 1-, 2-, 4-byte integers / Float32");
      }
    }
    default: throw EngineError("plan instruction has a missing operand");
  }
}

static ValClass sinkClassOf(int dt, bool dim) {
  switch (dt) {
    case Bool: if (dim) return VC_BOOL; break;
    case Int8: if (dim) return VC_I8; break;
    case Uint8: if (dim) return VC_U8; break;
    case Int16: if (dim) return VC_I16; break;
    case Uint16: if (dim) return VC_U16; break;
    case Int32: return VC_I32;
    case Uint32: return VC_U32;
    case Float32: return VC_F32;
    case Int64: return VC_I64;
    case Float64: if (!dim) return VC_F64; break;
    case UUID: if (dim) return VC_UUID; break;
    default: break;
  }
  throw EngineError("Unsupported sink data type in plan");
}

static int classWidth(ValClass c) {
  switch (c) {
    case VC_BOOL: case VC_I8: case VC_U8: return 1;
    case VC_I16: case VC_U16: return 2;
    case VC_I32: case VC_U32: case VC_F32: return 4;
    case VC_I64: case VC_F64: return 8;
    default: return 16;
  }
}

// Translates the ABI plan into the device form, resolving value classes by the reference's rules.
static void compilePlan(const AggState *st, const BatchPlan &bp, DevPlan &P) {
  memset(&P, 0, sizeof(P));
  if (bp.NumColumns < 0 || bp.NumColumns > kMaxPlanCols)
    throw EngineError("the fused path stages at most 16 distinct columns per batch");
  if (bp.NumInsts <= 0 || bp.NumInsts > ARES_MAX_PLAN_INSTS) throw EngineError("invalid plan instruction count");
  P.ncols = bp.NumColumns;
  P.ninsts = bp.NumInsts;
  P.baseCounts = bp.BaseCounts;
  P.startCount = bp.StartCount;
  P.numRows = bp.NumRows;
  for (int c = 0; c < bp.NumColumns; c++) {
    P.cols[c].in = makeColumnDesc(bp.Columns[c], /*allowWide=*/true);
    int dt = bp.Columns[c].DataType;
    P.cols[c].width = dt == Bool ? 0 : (dt == Int8 || dt == Uint8) ? 1 : (dt == Int16 || dt == Uint16) ? 2
                    : (dt == Int64 || dt == Uint64) ? 8 : dt == UUID ? 16 : 4;
    P.cols[c].used = 0;      // set below by the instructions that read the column: only those are staged
    const ColumnRange &cr = bp.Ranges[c];
    P.cols[c].rangeKnown = cr.Known && cr.Min <= cr.Max && cr.Max < 0x80000000u && P.cols[c].width <= 4 ? 1 : 0;
    P.cols[c].rangeLo = cr.Min;
    P.cols[c].rangeHi = cr.Max;
    P.cols[c].staged = 0;
    P.cols[c].hasNulls = 0;
  }
  // joined dimension tables
  if (bp.NumForeignTables < 0 || bp.NumForeignTables > kMaxForeignTables || bp.NumForeignColumns < 0 ||
      bp.NumForeignColumns > kMaxForeignCols)
    throw EngineError("invalid number of foreign tables / columns");
  P.numForeignTables = (uint8_t)bp.NumForeignTables;
  P.numForeignCols = (uint8_t)bp.NumForeignColumns;
  for (int t = 0; t < bp.NumForeignTables; t++) {
    const int jc = bp.ForeignTables[t].JoinColumn;
    if (jc < 0 || jc >= bp.NumColumns) throw EngineError("foreign table joins on a column outside BatchPlan.Columns");
    if (P.cols[jc].width > 4) throw EngineError("join keys of the fused path are 1-, 2- or 4-byte columns");
    P.joinCol[t] = (uint8_t)jc;
  }
  for (int k = 0; k < bp.NumForeignColumns; k++) {
    const int t = bp.ForeignColumns[k].Table;
    if (t < 0 || t >= bp.NumForeignTables) throw EngineError("foreign column of a table outside BatchPlan.ForeignTables");
    P.foreignTableOf[k] = (uint8_t)t;
  }
  const DimLayout &RL = st->rowLayout;
  std::vector<uint8_t> stack;
  std::vector<bool> dimSeen(RL.numDims, false);
  bool measureSeen = false;
  P.lastFilter = -1;
  for (int i = 0; i < bp.NumInsts; i++) {
    const PlanInst &pi = bp.Insts[i];
    DevInst &I = P.insts[i];
    if (pi.NumOperands != 1 && pi.NumOperands != 2) throw EngineError("plan instruction must have 1 or 2 operands");
    I.nops = pi.NumOperands; I.fn = pi.Functor; I.sink = pi.Sink; I.sinkArg = pi.SinkArg;
    // operand classes (rhs popped first when both come from the stack)
    uint8_t bcls = VC_NONE, acls;
    if (pi.NumOperands == 2) {
      if (pi.B.Kind == PLAN_OPERAND_STACK) { bcls = operandClassOf(pi.B, bp, stack); stack.pop_back(); }
    }
    acls = operandClassOf(pi.A, bp, stack);
    if (pi.A.Kind == PLAN_OPERAND_STACK) stack.pop_back();
    if (pi.NumOperands == 2 && pi.B.Kind != PLAN_OPERAND_STACK) bcls = operandClassOf(pi.B, bp, stack);
    auto fill = [&](const PlanOperand &o, uint8_t &kind, uint8_t &col, uint8_t &valid, uint32_t &k) {
      kind = o.Kind; col = o.Column; valid = o.ConstValid;
      if (o.Kind == PLAN_OPERAND_COLUMN) P.cols[o.Column].used = 1;
      if (o.Kind == PLAN_OPERAND_FOREIGN) P.cols[P.joinCol[P.foreignTableOf[o.Column]]].used = 1;   // the join key is staged
      if (o.Kind == PLAN_OPERAND_CONST) {
        if (o.ConstType == ConstFloat) memcpy(&k, &o.Const.FloatVal, 4); else k = (uint32_t)o.Const.IntVal;
      }
    };
    fill(pi.A, I.akind, I.acol, I.avalid, I.aconst);
    if (pi.NumOperands == 2) fill(pi.B, I.bkind, I.bcol, I.bvalid, I.bconst);
    I.aclass = acls; I.bclass = bcls;

    const bool wideIn = acls == VC_I64 || acls == VC_UUID || bcls == VC_I64 || bcls == VC_UUID;
    if (wideIn) {
      // only "dimension = 8/16-byte column" is meaningful (what UnaryTransform Noop to a
      // DimensionOutput of the same type does)
      ValClass oc = pi.Sink == PLAN_SINK_DIMENSION ? sinkClassOf(pi.SinkDataType, true) : VC_NONE;
      if (pi.NumOperands != 1 || pi.Functor != Noop || pi.Sink != PLAN_SINK_DIMENSION || pi.A.Kind != PLAN_OPERAND_COLUMN ||
          oc != (ValClass)acls)
        throw EngineError("int64/UUID columns are only supported as verbatim dimensions on the fused path");
      I.wide = 1;
    } else if (pi.NumOperands == 2) {
      I.tclass = commonClass((ValClass)acls, (ValClass)bcls);
      Cell z; z.v = 0; z.valid = true;
      ValClass rc;
      evalBinary(pi.Functor, z, z, (ValClass)I.tclass, &rc);
      I.rclass = rc;
    } else {
      I.tclass = acls;
      Cell z; z.v = 0; z.valid = true;
      ValClass rc;
      evalUnary(pi.Functor, z, (ValClass)acls, &rc);
      I.rclass = rc;
    }
    switch (pi.Sink) {
      case PLAN_SINK_STACK: {
        ValClass oc = sinkClassOf(pi.SinkDataType, false);
        if (oc != VC_I32 && oc != VC_U32 && oc != VC_F32) throw EngineError("stack temporaries are Int32/Uint32/Float32");
        if ((int)stack.size() >= ARES_PLAN_STACK_DEPTH) throw EngineError("plan exceeds the evaluation stack depth");
        I.oclass = oc;
        stack.push_back(oc);
        break;
      }
      case PLAN_SINK_FILTER:
        I.oclass = VC_BOOL;
        P.lastFilter = i;
        break;
      case PLAN_SINK_DIMENSION: {
        if (pi.SinkArg >= RL.numDims) throw EngineError("dimension ordinal outside AggSpec.NumDimsPerDimWidth");
        ValClass oc = sinkClassOf(pi.SinkDataType, true);
        if (classWidth(oc) != RL.width[pi.SinkArg]) throw EngineError("dimension data type does not match its layout width");
        if (dimSeen[pi.SinkArg]) throw EngineError("dimension written twice");
        dimSeen[pi.SinkArg] = true;
        I.oclass = oc;
        I.rowOff = RL.rowOff[pi.SinkArg];
        I.width = RL.width[pi.SinkArg];
        I.nullOff = (uint8_t)(RL.valueBytes + pi.SinkArg);
        break;
      }
      case PLAN_SINK_MEASURE: {
        if (measureSeen) throw EngineError("only one measure per plan");
        measureSeen = true;
        ValClass oc = sinkClassOf(pi.SinkDataType, false);
        if (oc != st->measClass) throw EngineError("measure data type differs from AggSpec.MeasureDataType");
        I.oclass = oc;
        break;
      }
      default: throw EngineError("unknown plan sink");
    }
  }
  if (!stack.empty()) throw EngineError("plan leaves values on the evaluation stack");
  for (int d = 0; d < RL.numDims; d++)
    if (!dimSeen[d]) throw EngineError("plan does not produce every dimension of AggSpec");
  if (!measureSeen) throw EngineError("plan has no measure instruction");
  P.hasMeasure = 1;
  P.keyMode = st->keyMode;
  P.rowBytes = (uint8_t)RL.rowBytes;
  P.valueBytes = (uint8_t)RL.valueBytes;
  P.hashBits = (uint8_t)st->hashBits;
  P.aggOp = st->op;
  {  // can a reached accumulator return to the neutral element?  (global dense slots carry no "reached" flags)
    bool safe = st->op != OP_SUM_I32 && st->op != OP_SUM_I64;   // float sums (-0.0), min / max (extreme), AVG (count 0)
    if (!safe)   // integer sums: only of a positive literal (count(*)): never 0 again below 2^32 rows
      for (int i = 0; i < P.ninsts; i++) {
        const DevInst &I = P.insts[i];
        if (I.sink == PLAN_SINK_MEASURE && !I.wide && I.nops == 1 && I.fn == Noop && I.akind == OPK_CONST && I.avalid &&
            (I.aclass == VC_I32 || I.aclass == VC_U32) && (int32_t)I.aconst > 0)
          safe = true;
      }
    P.neutralSafe = safe && !st->hll;
  }
  P.measWidth = (uint8_t)st->measWidth;
  P.measClass = st->measClass;
  P.hll = st->hll ? (st->hllDense ? 2 : 1) : 0;
  P.denseSlots = st->hllDense ? (uint32_t)st->capacity : 0;
  const int agg = st->spec.AggFunc;
  P.skipCount = !((agg >= AGGR_SUM_UNSIGNED && agg <= AGGR_SUM_FLOAT) || agg == AGGR_AVG_FLOAT);
  P.measureIdentity = aggIdentity(agg, st->measClass);
  P.accNeutral = st->accNeutral;
}

// ---------------------------------------------------------------------------------------
// archive batches: run-length encoded (mode-3) columns are expanded once per batch into plain
// mode-2 scratch columns (one value + one validity bit per index position), so that the staged,
// specialised kernel can run on them instead of a positional run search per access per row.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
expandRleKernel(InputDesc d, const uint32_t *__restrict__ baseCounts, uint32_t startCount, uint32_t n, int width, bool direct,
                uint8_t *__restrict__ outValues, uint32_t *__restrict__ outNulls) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warpsPerGrid = gridDim.x * (blockDim.x >> 5);
  const uint32_t *counts = reinterpret_cast<const uint32_t *>(d.base);
  const uint8_t *vals = d.base + d.valuesOff, *nulls = d.base + d.nullsOff;
  for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w * 32 < n; w += warpsPerGrid) {
    const uint32_t i = w * 32 + lane;
    bool valid = false, bit = false;
    if (i < n) {
      // index position i stands for row baseCounts[i] (or startCount + i); `direct`: this column's own
      // count vector IS the batch's base counts, so its run number is i
      const uint32_t p = direct ? i : rlePosition(counts, d.length, baseCounts ? baseCounts[i] : startCount + i);
      valid = bitAt(nulls, p + d.startBit);
      switch (width) {
        case 0: bit = bitAt(vals, p + d.startBit); break;
        case 1: outValues[i] = vals[p]; break;
        case 2: reinterpret_cast<uint16_t *>(outValues)[i] = reinterpret_cast<const uint16_t *>(vals)[p]; break;
        default: reinterpret_cast<uint32_t *>(outValues)[i] = reinterpret_cast<const uint32_t *>(vals)[p]; break;
      }
    }
    const uint32_t vword = __ballot_sync(0xFFFFFFFFu, valid), bword = __ballot_sync(0xFFFFFFFFu, bit);
    if (lane == 0) {
      outNulls[w] = vword;
      if (width == 0) reinterpret_cast<uint32_t *>(outValues)[w] = bword;
    }
  }
}

// First-class RLE columns: run that holds the first index position of every tile (and of the batch's last position), so
// that the kernel searches a window of a few runs per quad instead of the whole count vector per row.
__global__ void __launch_bounds__(256)
rleTileRunsKernel(const uint32_t *__restrict__ counts, uint32_t length, const uint32_t *__restrict__ baseCounts, uint32_t startCount,
                  uint32_t numRows, uint32_t tileRows, uint32_t entries, uint32_t *__restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= entries) return;
  uint64_t pos = (uint64_t)t * tileRows;
  if (pos > numRows - 1) pos = numRows - 1;
  const uint32_t row = baseCounts ? baseCounts[pos] : startCount + (uint32_t)pos;
  out[t] = rlePosition(counts, length, row);
}

// Decides staged vs direct, the tile size, the stage layout, the TMA ring depth and the shared table
// size.  The shared table gets what the workload needs first (a table that overflows sends rows to
// contended L2 atomics, profiles/r01_agg_microbench.txt), the ring takes the rest.
static size_t layoutStages(DevPlan &P, uint32_t expectedGroups, bool allowDense = true) {
  bool canStage = P.numRows >= 1024;
  bool anyRle = false;
  uint32_t rowBits = 0;
  for (int c = 0; c < P.ncols; c++) {
    DevColumn &col = P.cols[c];
    if (col.in.mode == 0 || !col.used) continue;
    if (col.rle) { anyRle = true; continue; }            // RLE decoded in the kernel from its runs: nothing to stage
    if (col.in.mode == 3) { canStage = false; break; }  // RLE without the specialised kernel: positional search, direct path
    if (col.width > 4) continue;                          // wide dims are read directly
    const uintptr_t v = reinterpret_cast<uintptr_t>(col.in.base + col.in.valuesOff);
    if (v & 15) canStage = false;
    if (col.in.mode == 2 && (reinterpret_cast<uintptr_t>(col.in.base + col.in.nullsOff) & 15)) canStage = false;
    rowBits += col.width ? col.width * 8 : 1;
    if (col.in.mode == 2) rowBits += 1;
  }
  // The shared table takes 8192 slots (128 KB) whenever a ring of >= 2 stages still fits beside
  // it: measured on cfg3 (2,400 groups per batch) 8192 slots beat 4096 by 1.5x because fewer
  // probe iterations are paid per warp; a plan with very wide rows falls back to fewer slots.
  uint32_t slots = 8192;
  // The caller expects far more groups than the shared table holds (or HLL, whose entries are
  // (group, register) pairs): compile the kernel with the direct-to-global mode it can switch to.
  P.bypassOk = (P.hll || expectedGroups > 4 * slots) ? 1 : 0;
  // zone map known for every dimension: no key table, slots addressed by dimension value (jit.cu)
  P.denseGlobal = 0;
  if (allowDense) jitAnalyzeDense(P, P.bypassOk != 0);
  else P.denseNd = 0;
  if (const char *e = getenv("ARESDB_B200_SMEM_SLOTS")) {  // tuning / experiments
    uint32_t v = (uint32_t)atoi(e);
    if (v >= 256 && v <= 8192 && (v & (v - 1)) == 0) slots = v;
  }
  auto stageBytesFor = [&](uint32_t tr) {
    size_t stage = 0;
    for (int c = 0; c < P.ncols; c++) {
      const DevColumn &col = P.cols[c];
      if (col.in.mode == 0 || !col.used || col.width > 4 || col.rle) continue;
      stage += ((col.width ? (size_t)tr * col.width : tr / 8 + 16) + 15) / 16 * 16;
      if (col.in.mode == 2) stage += (tr / 8 + 16 + 15) / 16 * 16;
    }
    // base counts of an RLE batch ride along when a SUM / AVG measure needs the run lengths, or an RLE column the row numbers
    if (P.baseCounts != nullptr && (!P.skipCount || anyRle) && (reinterpret_cast<uintptr_t>(P.baseCounts) & 15) == 0) stage += ((size_t)tr + 4) * 4;
    return stage;
  };
  uint32_t tileRows = 0, stages = 0;
  if (canStage && rowBits > 0) {
    uint32_t forceTile = 0;
    if (const char *e = getenv("ARESDB_B200_TILE_ROWS")) forceTile = (uint32_t)atoi(e);  // tuning / experiments
    if (P.denseNd != 0) {
      // Dense slots cost 9 bytes each (flag + 8-byte accumulator).  Give the ring as many stages as possible
      // and the slots the rest: the capacity (part of the kernel text) then depends on the stage layout only,
      // not on the batch's ranges.
      for (uint32_t tr : {3968u, 1920u, 896u}) {
        if (forceTile && tr != forceTile) continue;
        for (uint32_t n = kMaxStages; n >= 2 && !tileRows; n--) {
          const size_t need = 128 + n * stageBytesFor(tr);
          if (need >= (size_t)kSmemBudget) continue;
          // three 32-bit piece counters, or flag + 8-byte accumulator; HLL: one 32-bit map entry (slot -> group's registers)
          const size_t slotBytes = P.hll ? 4 : P.denseFx ? 12 : 9;
          uint32_t cap = (uint32_t)(((size_t)kSmemBudget - need) / slotBytes / 16 * 16);
          if (cap > kDenseMaxSlots) cap = kDenseMaxSlots;
          if (cap >= P.denseTotal) { tileRows = tr; stages = n; slots = cap; }
        }
        if (tileRows) break;
      }
      if (!tileRows && !P.hll && P.neutralSafe && P.denseTotal <= kGlobalDenseMaxSlots) {
        // more slots than a CTA holds: one accumulator array in global memory for the whole grid; shared memory is all ring
        for (uint32_t tr : {3968u, 1920u, 896u}) {
          if (forceTile && tr != forceTile) continue;
          const uint32_t n = (uint32_t)(((size_t)kSmemBudget - 128 - 256) / stageBytesFor(tr));
          if (n >= 2) {
            tileRows = tr; stages = n > (uint32_t)kMaxStages ? kMaxStages : n; slots = 16; P.denseGlobal = 1;
            // L2 atomics saturate (~190 G/s) at >= ~1M distinct addresses and contend below (profiles/r01_agg_microbench.txt):
            // replicate the slot array until it has about that many
            uint32_t reps = 1;
            while (reps < 8 && (uint64_t)P.denseTotal * reps * 2 <= kGlobalDenseMaxSlots && (uint64_t)P.denseTotal * reps < (1u << 20)) reps *= 2;
            if (const char *e = getenv("ARESDB_B200_GLOBAL_REPS")) { const int v = atoi(e); if (v >= 1 && v <= 8 && (v & (v - 1)) == 0 && (uint64_t)P.denseTotal * v <= kGlobalDenseMaxSlots) reps = (uint32_t)v; }
            P.denseGlobalReps = (uint8_t)reps;
            break;
          }
        }
      }
      if (!tileRows) P.denseNd = 0;   // no layout holds the slots: hash table
    }
    if (!tileRows) {
      // compacted-index form (survivors of the filters gathered per tile before the expensive part)
      P.compact = 0;
      if (allowDense && P.denseNd == 0 && jitAvailable() && planCompactable(P)) {
        // (measured on cfg4 HLL without zone maps: 1.40 ms against 1.25 ms for the plain form — the kernel is bound by the
        // latency of its directory / register accesses, and the barrier of the compaction costs more than the instructions
        // it saves, profiles/r02_fused_hll_v2_compact_rejected.summary.txt: opt-in only)
        const char *e = getenv("ARESDB_B200_COMPACT");
        P.compact = e ? (e[0] == '1') : 0;
      }
      if (P.partition) P.compact = 0;
      for (uint32_t sl : {slots, slots / 2, slots / 4}) {
        if (P.partition && sl != 8192) break;      // the tile buffer of the partitioned form IS the 64 KB table region
        for (uint32_t tr : {3968u, 1920u, 896u}) {  // 128 rows x (31 | 15 | 7) consumer warps
          if (forceTile && tr != forceTile) continue;
          size_t avail = (size_t)kSmemBudget - 128 - (size_t)sl * 8 - (P.compact ? kCompactListBytes : 0u) - (P.partition ? kPartitionExtraBytes : 0u);
          uint32_t n = (uint32_t)(avail / stageBytesFor(tr));
          if (n >= 2) { tileRows = tr; stages = n > (uint32_t)kMaxStages ? kMaxStages : n; break; }
        }
        if (tileRows) { slots = sl; break; }
      }
    }
  } else {
    P.denseNd = 0;
  }
  size_t stageBytes = 0;
  bool anyStaged = false;
  if (tileRows) {
    for (int c = 0; c < P.ncols; c++) {
      DevColumn &col = P.cols[c];
      if (col.in.mode == 0 || !col.used || col.width > 4 || col.rle) continue;
      col.staged = 1;
      anyStaged = true;
      col.smemValues = (uint32_t)stageBytes;
      // bit-packed bools and bitmaps copy one 16-byte chunk beyond the tile so that a non-zero
      // StartingIndex can read across the tile's last byte; full tiles always have those bytes.
      col.tileValueBytes = col.width ? tileRows * col.width : tileRows / 8 + 16;
      stageBytes += (col.tileValueBytes + 15) / 16 * 16;
      if (col.in.mode == 2) {
        col.hasNulls = 1;
        col.smemNulls = (uint32_t)stageBytes;
        col.tileNullBytes = tileRows / 8 + 16;
        stageBytes += (col.tileNullBytes + 15) / 16 * 16;
      }
    }
  }
  P.smemBc = 0;
  P.tileBcBytes = 0;
  if (anyStaged && P.baseCounts != nullptr && (!P.skipCount || anyRle) && (reinterpret_cast<uintptr_t>(P.baseCounts) & 15) == 0) {
    P.smemBc = (uint32_t)stageBytes;
    P.tileBcBytes = (tileRows + 4) * 4;   // one count more than rows (run length = difference), padded to 16 bytes
    stageBytes += P.tileBcBytes;
  }
  P.staged = anyStaged;
  P.tileRows = anyStaged ? tileRows : 4 * kFusedThreads;
  P.numStages = anyStaged ? stages : 0;
  // keep >= 128 rows for the direct tail so that the last staged tile's 16-byte bitmap over-read
  // stays inside the column
  P.numFullTiles = anyStaged && P.numRows > 128 ? (P.numRows - 128) / tileRows : 0;
  if (P.numFullTiles == 0) { P.staged = 0; stageBytes = 0; P.numStages = 0; P.denseNd = 0; P.denseGlobal = 0; if (slots > 8192 || slots < 256) slots = 8192; }
  if (P.denseNd == 0) P.denseGlobal = 0;
  P.stageBytes = (uint32_t)stageBytes;
  P.smemSlots = slots;
  P.tableBytes = P.denseNd != 0 ? (slots * (P.hll ? 4 : P.denseFx ? 12 : 9) + 127) / 128 * 128 : slots * 8;
  if (P.denseNd != 0 || !P.staged) { P.compact = 0; P.partition = 0; }
  if (P.partition && (slots != 8192 || P.tileRows > 4096)) P.partition = 0;
  return 128 + (size_t)P.tableBytes + (P.compact ? kCompactListBytes : 0u) + (P.partition ? kPartitionExtraBytes : 0u) + stageBytes * P.numStages;
}

// ---------------------------------------------------------------------------------------
// growth of the group table
// ---------------------------------------------------------------------------------------
// The reference sizes its hash map at 2 x rows for every batch and so never runs out (query/hash_reduction.cu:211-292);
// here the table starts L2-sized and doubles when it is half full.  Kernels notice on the device (DevTable::growAt
// raises the STOP flag at claim time): tile kernels whose launch the host waits for drain and are resumed after the
// growth; kernels the host does not wait for (direct-indexed ones) park new groups in the spill list; launches that
// cannot be resumed (merges, folds, the direct path) get their room up front.
__global__ void __launch_bounds__(256) rehashKernel(DevTable oldG, uint32_t n, DevTable newG) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t slot = oldG.claimed[i];
    const uint32_t to = globalFindOrClaim(newG, oldG.keys[slot], oldG.rows ? &oldG.rows[(size_t)slot * 4] : nullptr);
    if (to != 0xFFFFFFFFu) newG.acc[to] = oldG.acc[slot];
  }
}

__global__ void __launch_bounds__(256) mergeSpillKernel(DevTable G, uint32_t n, AggOp op) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const SpillEntry e = G.spill[i];
    globalUpdate(G, op, e.key, G.rows ? e.row : nullptr, e.val);
  }
}

// Radix-partitioned aggregation, pass 2: the entries the fused kernel appended tile by tile (each tile's span sorted by
// partition, its directory line giving the 64 segment offsets) are folded PARTITION-MAJOR: warp w takes pair
// (partition, tile) number w, w + W, ... in that order, so that at any moment the whole grid updates the same
// partition = one contiguous 1/64 of the table's slots, which stays in L2 while it is being worked on.
__global__ void __launch_bounds__(256)
partitionAggregateKernel(const uint4 *__restrict__ entries, const uint32_t *__restrict__ dir, uint32_t numTiles, AggOp op, DevTable G) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5), pairs = (uint64_t)kPartitions * numTiles;
  for (uint64_t idx = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); idx < pairs; idx += warps) {
    const uint32_t p = (uint32_t)(idx / numTiles), tile = (uint32_t)(idx % numTiles);
    const uint32_t *line = dir + (size_t)tile * kPartDirWords;
    const uint32_t begin = line[kPartitions + 1] + line[p], end = line[kPartitions + 1] + line[p + 1];
    for (uint32_t i = begin + lane; i < end; i += 32) {
      const uint4 e = entries[i];
      globalUpdate(G, op, (unsigned long long)e.x | ((unsigned long long)e.y << 32), nullptr,
                   (uint64_t)e.z | ((uint64_t)e.w << 32), /*spillWhenStopped=*/true);
    }
  }
}

struct TableCounters { uint32_t occupied, overflow, truncated, stop, spilled, spillOverflow; };
static void checkOverflow(AggState *st, const uint32_t counters[2]);

static TableCounters readCounters(AggState *st, cudaStream_t s) {
  TableCounters c;
  uint32_t *pinned = st->resultHost + 10;   // (pinned: a pageable target costs a staging copy and ~100 us)
  ARES_CUDA(cudaMemcpyAsync(pinned, st->table.counters, sizeof(c), cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  memcpy(&c, pinned, sizeof(c));
  st->occUpper = c.occupied;
  st->resultHost[8] = c.occupied;
  return c;
}

// New table of `newCap` slots holding the groups of the old one; STOP flag cleared.
static void growTable(AggState *st, size_t newCap, cudaStream_t s) {
  if (st->hllDense) throw EngineError("dense HLL state: more than " + std::to_string(st->capacity) + " dimension groups");
  if (newCap > ((size_t)1 << 31)) throw EngineError("group table cannot grow beyond 2^31 slots");
  void *oldMem = st->mem;
  const DevTable oldG = st->table;
  uint32_t header[64];
  ARES_CUDA(cudaMemcpyAsync(header, oldG.counters, sizeof(header), cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  const uint32_t n = header[0];
  allocTable(st, newCap, s);   // fresh table (keys empty, counters 0), new claim list / progress / spill areas
  if (n) {
    int blocks = divUp(n, 256);
    if (blocks > smCount() * 8) blocks = smCount() * 8;
    rehashKernel<<<blocks, 256, 0, s>>>(oldG, n, st->table);
    checkLastError("rehash");
  }
  // progress of a stopped launch and parked rows survive the move
  ARES_CUDA(cudaMemcpyAsync(st->table.progress, oldG.progress, (size_t)(kProgressTail + 1) * 4, cudaMemcpyDeviceToDevice, s));
  const uint32_t spilled = header[4] < kSpillCap ? header[4] : kSpillCap;
  if (spilled) ARES_CUDA(cudaMemcpyAsync(st->table.spill, oldG.spill, (size_t)spilled * sizeof(SpillEntry), cudaMemcpyDeviceToDevice, s));
  uint32_t keep[2] = {header[4], header[5]};
  ARES_CUDA(cudaMemcpyAsync(st->table.counters + 4, keep, sizeof(keep), cudaMemcpyHostToDevice, s));
  ARES_CUDA(cudaMemcpyAsync(st->table.counters + 1, header + 1, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));   // overflow / truncated
  ARES_CUDA(cudaStreamSynchronize(s));
  deviceFree(oldMem);
  st->occUpper = n;
}

// Folds parked rows (growing first when needed); loud when more rows were parked than the list holds.
static void settleTable(AggState *st, cudaStream_t s) {
  TableCounters c = readCounters(st, s);
  if (c.spillOverflow)
    throw EngineError("group table: more than " + std::to_string(kSpillCap) + " rows of new groups arrived from direct-indexed batches while the "
                      "table was full (a zone map far off the data); recreate the AggState with a larger AggSpec.ExpectedGroups and replay");
  if (c.overflow) checkOverflow(st, &c.occupied);
  if (c.stop && st->unchecked > 0) {
    const uint32_t n = st->unchecked;
    st->unchecked = 0;
    throw EngineError("the group table reached its growth threshold (" + std::to_string(c.occupied) + " groups) during " + std::to_string(n) +
                      " batch(es) whose launches were not waited for — the number of groups jumped from under an eighth of that threshold; "
                      "rows were NOT folded: recreate the AggState with AggSpec.ExpectedGroups >= " + std::to_string((uint64_t)c.occupied * 4) +
                      " (or set ARESDB_B200_CHECK_EVERY_BATCH=1) and replay the batches");
  }
  if (!c.stop && !c.spilled) return;
  size_t cap = st->capacity;
  while ((uint64_t)c.occupied + c.spilled > cap / 4) cap <<= 1;   // settle at a quarter full at most
  if (cap != st->capacity || c.stop) growTable(st, cap == st->capacity ? cap << 1 : cap, s);
  if (c.spilled) {
    int blocks = divUp(c.spilled, 256);
    mergeSpillKernel<<<blocks, 256, 0, s>>>(st->table, c.spilled, st->op);
    checkLastError("mergeSpill");
    ARES_CUDA(cudaMemsetAsync(st->table.counters + 4, 0, 8, s));
    readCounters(st, s);
  }
}

// Room for a launch that cannot be resumed and inserts at most `bound` new groups.
static void ensureRoom(AggState *st, uint64_t bound, cudaStream_t s) {
  if (st->hllDense) return;
  if (st->occUpper + bound < st->table.growAt) { st->occUpper += bound; return; }
  settleTable(st, s);   // exact occupancy (and nothing parked)
  size_t cap = st->capacity;
  while (st->occUpper + bound >= cap / 2) cap <<= 1;
  if (cap != st->capacity) growTable(st, cap, s);
  st->occUpper += bound;
}

static void executePlan(AggState *st, const BatchPlan &bp, cudaStream_t s) {
  if (bp.NumRows == 0) return;
  if (bp.NumRows > 0x7FFFFFFFu) throw EngineError("a batch holds at most 2^31-1 rows");
  static thread_local DevPlan P;  // ~3 KB; passed by value as a __grid_constant__ parameter
  compilePlan(st, bp, P);
  P.tailBegin = 0;
  P.resume = 0;
  // Archive batches: run-length encoded (mode 3) columns.
  //  * the column whose count vector IS the batch's base counts has one value per index position: it is read like an
  //    uncompressed column (values / null bitmap of its runs), no copy;
  //  * every other RLE column is a FIRST-CLASS input of the specialised kernel: decoded from its runs inside the tile loop
  //    (jit_kernel_head.cuh ldrle) with a per-tile run hint computed below — HBM sees the runs, not the rows;
  //  * without NVRTC (or with ARESDB_B200_EXPAND_RLE=1) the column is expanded once per batch for the interpreter.
  std::vector<std::unique_ptr<Scratch>> expanded;
  static const bool forceExpand = [] { const char *e = getenv("ARESDB_B200_EXPAND_RLE"); return e && e[0] == '1'; }();
  static const bool keepPositional = [] { const char *e = getenv("ARESDB_B200_EXPAND_RLE"); return e && e[0] == '0'; }();
  if (bp.NumRows >= 1024 && !keepPositional) {
    const uint32_t n = bp.NumRows;
    // (the tile loop is driven by the TMA ring: at least one plain column must be staged for the RLE columns to ride along)
    int stageable = 0;
    for (int c = 0; c < P.ncols; c++) {
      const DevColumn &col = P.cols[c];
      if (col.used && col.width <= 4 && (col.in.mode == 1 || col.in.mode == 2)) stageable++;
      if (col.used && col.in.mode == 3 && col.width <= 4 && bp.BaseCounts != nullptr &&
          reinterpret_cast<const uint32_t *>(col.in.base) == bp.BaseCounts && col.in.length >= n) stageable++;
    }
    const bool firstClass = jitAvailable() && !forceExpand && stageable > 0;
    int nrle = 0;
    for (int c = 0; c < P.ncols; c++) {
      DevColumn &col = P.cols[c];
      if (!col.used || col.in.mode != 3 || col.width > 4) continue;
      const bool direct = bp.BaseCounts != nullptr && reinterpret_cast<const uint32_t *>(col.in.base) == bp.BaseCounts;
      if (direct && col.in.length >= n) {   // run number == index position
        col.in.mode = 2;
        continue;
      }
      if (firstClass && nrle < 4 && col.in.length > 0) {
        col.rle = 1;
        nrle++;
        continue;
      }
      const size_t nullBytes = ((size_t)(n + 31) / 32 * 4 + 16 + 63) / 64 * 64;
      const size_t valueBytes = (col.width ? (size_t)n * col.width : (size_t)(n + 31) / 32 * 4) + 64;
      expanded.emplace_back(new Scratch(nullBytes + valueBytes, s));
      uint8_t *buf = expanded.back()->as<uint8_t>();
      int blocks = divUp((int64_t)(n + 31) / 32, 8);
      if (blocks > smCount() * 16) blocks = smCount() * 16;
      expandRleKernel<<<blocks, 256, 0, s>>>(col.in, bp.BaseCounts, bp.StartCount, n, col.width, direct, buf + nullBytes,
                                             reinterpret_cast<uint32_t *>(buf));
      checkLastError("expandRle");
      col.in.base = buf;
      col.in.nullsOff = 0;
      col.in.valuesOff = (uint32_t)nullBytes;
      col.in.length = n;
      col.in.mode = 2;
      col.in.startBit = 0;
    }
  }
  // joined dimension tables: indexes + foreign-column batches go to device memory for the kernel's lifetime
  std::unique_ptr<Scratch> joinMem;
  P.join = nullptr;
  if (P.numForeignCols > 0 || P.numForeignTables > 0) {
    static thread_local DevJoin J;
    memset(&J, 0, sizeof(J));
    for (int t = 0; t < bp.NumForeignTables; t++) {
      const CuckooHashIndex &h = bp.ForeignTables[t].Index;
      if (h.buckets == nullptr || h.numBuckets <= 0 || h.keyBytes <= 0 || h.keyBytes > 4 || h.numHashes < 0 || h.numHashes > 4)
        throw EngineError("invalid CuckooHashIndex in BatchPlan.ForeignTables (fused path: keys of at most 4 bytes)");
      J.tables[t].buckets = h.buckets;
      for (int i = 0; i < 4; i++) J.tables[t].seeds[i] = h.seeds[i];
      J.tables[t].keyBytes = h.keyBytes; J.tables[t].numHashes = h.numHashes; J.tables[t].numBuckets = h.numBuckets;
    }
    for (int k = 0; k < bp.NumForeignColumns; k++) J.cols[k] = makeForeignDesc(bp.ForeignColumns[k].Column);
    joinMem.reset(new Scratch(sizeof(DevJoin), s));
    ARES_CUDA(cudaMemcpyAsync(joinMem->ptr, &J, sizeof(DevJoin), cudaMemcpyHostToDevice, s));
    ARES_CUDA(cudaStreamSynchronize(s));   // J is reused by the next call of this thread
    P.join = joinMem->as<DevJoin>();
  }
  // radix-partitioned aggregation when the table is far beyond L2 (or ARESDB_B200_PARTITION=1): packed keys, plain sums /
  // min / max, specialised kernel only
  {
    const char *e = getenv("ARESDB_B200_PARTITION");
    const bool want = e ? e[0] == '1' : st->capacity >= kPartitionMinSlots;
    P.partition = want && jitAvailable() && st->keyMode == KEY_PACKED && !st->hll && P.numForeignCols == 0 && st->capacity >= 4096;
    int lg = 0;
    while (((size_t)1 << lg) < st->capacity) lg++;
    P.partShift = (uint8_t)(lg > 6 ? lg - 6 : 0);
  }
  size_t smemBytes = layoutStages(P, st->spec.ExpectedGroups);
  std::unique_ptr<Scratch> partMem;
  if (P.partition) {
    const size_t dirBytes = ((size_t)P.numFullTiles * kPartDirWords * 4 + 255) / 256 * 256;
    partMem.reset(new Scratch(256 + dirBytes + (size_t)P.numRows * sizeof(uint4), s));
    uint8_t *pm = partMem->as<uint8_t>();
    P.partCursor = reinterpret_cast<uint32_t *>(pm);
    P.partDir = reinterpret_cast<uint32_t *>(pm + 256);
    P.partBuf = reinterpret_cast<uint4 *>(pm + 256 + dirBytes);
    ARES_CUDA(cudaMemsetAsync(pm, 0, 256, s));
  }
  // per-tile run hints of the first-class RLE columns (the tile size is known now)
  for (int c = 0; c < P.ncols; c++) {
    DevColumn &col = P.cols[c];
    if (!col.rle) continue;
    if (!P.staged) { col.rle = 0; continue; }   // no tile loop (tiny batch): the generic positional read
    const uint32_t entries = (P.numRows + P.tileRows - 1) / P.tileRows + 2;
    expanded.emplace_back(new Scratch(sizeof(uint32_t) * entries, s));
    rleTileRunsKernel<<<divUp(entries, 256), 256, 0, s>>>(reinterpret_cast<const uint32_t *>(col.in.base), col.in.length, bp.BaseCounts,
                                                         bp.StartCount, P.numRows, P.tileRows, entries, expanded.back()->as<uint32_t>());
    checkLastError("rleTileRuns");
    col.tileRun = expanded.back()->as<uint32_t>();
  }
  // room in the group table (see "growth of the group table"): the direct path and the direct-indexed kernels are not
  // waited for, so what they may insert is reserved up front (flush of the CTA slots / fold of the global slot array;
  // out-of-range rows park); hash-table tile kernels are checked after the launch and resumed when they stopped.
  const bool resumable = P.staged && P.denseNd == 0 && !st->hllDense && !P.partition;
  if (!resumable && !st->hllDense) ensureRoom(st, P.staged ? (uint64_t)P.denseTotal : (uint64_t)P.numRows, s);
  P.ctaAcc = st->ctaAcc;   // (after a possible growth: the slices live in the table's allocation)
  static bool attrSet[64] = {false};
  if (!attrSet[st->device & 63]) {
    ARES_CUDA(cudaFuncSetAttribute(fusedBatchKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    ARES_CUDA(cudaFuncSetAttribute(fusedBatchKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attrSet[st->device & 63] = true;
  }
  auto gridFor = [&]() {
    int g = smCount() < kMaxGridCtas ? smCount() : kMaxGridCtas;
    const uint32_t work = P.staged ? P.numFullTiles : (P.numRows + 4 * kFusedThreads - 1) / (4 * kFusedThreads);
    if ((uint32_t)g > work) g = work ? (int)work : 1;
    return g;
  };
  int grid = gridFor();
  if (P.denseGlobal) {
    if (!st->denseAcc) {   // first use: 16 MB of accumulators at the neutral element (denseFoldKernel leaves them so)
      void *mem = nullptr;
      CGoCallResHandle h = deviceMalloc(&mem, (size_t)kGlobalDenseMaxSlots * sizeof(unsigned long long));
      if (h.pStrErr) { std::string m(h.pStrErr); free((void *)h.pStrErr); throw EngineError(m); }
      st->denseAcc = static_cast<unsigned long long *>(mem);
      fillTableKernel<<<smCount() * 8, 256, 0, s>>>(st->denseAcc, st->denseAcc, kGlobalDenseMaxSlots, st->accNeutral);
      checkLastError("denseAcc fill");
    }
    P.denseAcc = st->denseAcc;
  }
  // the specialised kernel covers the staged tiles AND the tail; the interpreter below is the
  // generic fallback (unaligned / RLE columns, NVRTC unavailable or disabled)
  for (;;) {
    bool checkAfter = false;   // a hash-table tile kernel ran: wait for it and resume it if the table stopped it
    if (P.staged && jitLaunchStaged(P, st->table, smemBytes, grid, s)) {
      checkAfter = P.denseNd == 0 && !st->hllDense;
      if (P.partition) {   // pass 2: fold the entries partition by partition; the table is checked (and grown) after every batch
        partitionAggregateKernel<<<smCount() * 8, 256, 0, s>>>(P.partBuf, P.partDir, P.numFullTiles, (AggOp)P.aggOp, st->table);
        checkLastError("partitionAggregate");
        const TableCounters c = readCounters(st, s);
        st->everChecked = true;
        if (c.overflow) checkOverflow(st, &c.occupied);
        if (c.stop || c.spilled) settleTable(st, s);
        return;
      }
      if (P.denseGlobal) {
        DenseFold F;
        memset(&F, 0, sizeof(F));
        uint32_t stride = 1;
        for (int k = 0; k < P.denseNd; k++) {
          const DevInst &I = P.insts[P.denseInst[k]];
          F.lo[k] = P.denseLo[k]; F.cnt[k] = P.denseCnt[k]; F.step[k] = P.denseStep[k]; F.stride[k] = stride;
          F.rowOff[k] = I.rowOff; F.width[k] = I.width; F.nullOff[k] = I.nullOff;
          stride *= P.denseCnt[k] + 1;
        }
        F.nd = P.denseNd; F.total = P.denseTotal; F.reps = P.denseGlobalReps ? P.denseGlobalReps : 1;
        F.keyMode = P.keyMode; F.hashBits = P.hashBits; F.rowBytes = P.rowBytes; F.op = P.aggOp;
        F.neutral = P.accNeutral;
        int blocks = divUp((int64_t)P.denseTotal, 256);
        if (blocks > smCount() * 8) blocks = smCount() * 8;
        denseFoldKernel<<<blocks, 256, 0, s>>>(st->denseAcc, F, st->table);
        checkLastError("denseFold");
      }
    } else {
      if (P.denseNd != 0) {  // the interpreter needs the key-table layout
        smemBytes = layoutStages(P, st->spec.ExpectedGroups, /*allowDense=*/false);
        grid = gridFor();
      }
      checkAfter = P.staged && !st->hllDense;
      if (st->keyMode == KEY_HASHED) fusedBatchKernel<true><<<grid, kFusedThreads, smemBytes, s>>>(P, st->table);
      else fusedBatchKernel<false><<<grid, kFusedThreads, smemBytes, s>>>(P, st->table);
      checkLastError("ExecuteBatchPlan");
    }
    if (!checkAfter) return;
    // Waiting for every hash-table batch would cost a launch gap per batch (the host cannot prepare the next one while it
    // waits).  The wait is therefore adaptive: always for the first batch of a query, and from then on whenever the
    // occupancy last seen — read back then, or published by a finishing kernel into mapped pinned memory — is above an
    // eighth of the threshold.  A table that fills from under an eighth of its threshold within unwatched batches is
    // reported loudly at the next synchronising call (settleTable), never folded wrongly.
    static const bool everyBatch = [] { const char *e = getenv("ARESDB_B200_CHECK_EVERY_BATCH"); return e && e[0] == '1'; }();
    const uint64_t seen = st->resultHost[8] > st->occUpper ? st->resultHost[8] : st->occUpper;
    if (P.resume == 0 && st->everChecked && !everyBatch && st->unchecked < 64 && seen * 8 < st->table.growAt) {
      st->unchecked++;
      return;
    }
    const TableCounters c = readCounters(st, s);
    st->everChecked = true;
    if (c.overflow) checkOverflow(st, &c.occupied);
    if (!c.stop && !c.spilled) { st->unchecked = 0; return; }
    if (c.stop && st->unchecked > 0) settleTable(st, s);   // raises: earlier unwatched batches stopped as well
    const bool stopped = c.stop != 0;
    settleTable(st, s);   // grows (a quarter full at most afterwards) and folds the parked groups
    if (!stopped) return;
    P.resume = 1;
    P.ctaAcc = st->ctaAcc;
  }
}

static void mergeRows(AggState *st, const DimensionVector &in, const uint8_t *values, int length, cudaStream_t s) {
  if (length <= 0) return;
  for (int i = 0; i < NUM_DIM_WIDTH; i++)
    if (in.NumDimsPerDimWidth[i] != st->spec.NumDimsPerDimWidth[i]) throw EngineError("dimension layout differs from AggSpec");
  DimLayout L = makeDimLayout(in.NumDimsPerDimWidth, in.VectorCapacity);
  ensureRoom(st, (uint64_t)length, s);
  int blocks = divUp(length, 256);
  if (blocks > smCount() * 8) blocks = smCount() * 8;
  mergeRowsKernel<<<blocks, 256, 0, s>>>(in.DimValues, L, values, st->measWidth, st->op, length, st->keyMode,
                                        (uint8_t)st->hashBits, st->hll ? (st->hllDense ? 2 : 1) : 0, st->table);
  checkLastError("AggStateMerge");
}

static void checkOverflow(AggState *st, const uint32_t counters[2]) {
  if (counters[1] && st->hllDense)
    throw EngineError("dense HLL state: more than " + std::to_string(kHllDenseMaxGroups) +
                      " dimension groups; recreate the AggState with AggSpec.ExpectedGroups > 4096 (entry mode) and replay the batches");
  if (counters[1])
    throw EngineError("group table overflow: more than " + std::to_string(st->capacity) +
                      " slots needed; recreate the AggState with a larger AggSpec.ExpectedGroups and replay the batches");
}

static int64_t groupCount(AggState *st, cudaStream_t s) {
  TableCounters c = readCounters(st, s);
  if (c.stop || c.spilled || c.spillOverflow) {
    settleTable(st, s);
    c = readCounters(st, s);
  }
  checkOverflow(st, &c.occupied);
  return c.occupied;
}

// Dense HLL state -> carried rows.  Fills `block` (one dim row per group, capacity = groups, hash
// order) and, per hit register in (group, register) order, key / value / group ordinal.
struct DenseCarried {
  int groups = 0;
  int64_t entries = 0;
  Scratch block, hash, values, index;
};

static void denseCarried(AggState *st, cudaStream_t s, DenseCarried &out, bool countOnly) {
  uint32_t c[2];
  ARES_CUDA(cudaMemcpyAsync(c, st->table.counters, sizeof(c), cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  checkOverflow(st, c);
  const int n = (int)c[0];
  out.groups = n;
  out.entries = 0;
  if (n == 0) return;
  // groups in the order of the reference hash of their dim row
  const int tiles = divUp((int64_t)st->capacity, kCmpTile);
  Scratch state(scanStateBytes(tiles) + sizeof(uint32_t), s);
  ARES_CUDA(cudaMemsetAsync(state.ptr, 0, state.bytes, s));
  ScanTileState sst = makeScanState(state.ptr, tiles);
  uint32_t *dCount = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(state.ptr) + scanStateBytes(tiles));
  Scratch slotOf(sizeof(uint32_t) * (size_t)n, s), hash(sizeof(uint64_t) * (size_t)n, s), vals(sizeof(uint32_t) * (size_t)n, s);
  compactGroupsKernel<<<tiles, kCmpThreads, 0, s>>>(st->table, st->capacity, st->keyMode, 64, ~0ull, st->rowLayout.rowBytes, 4, sst,
                                                   slotOf.as<uint32_t>(), hash.as<uint64_t>(), vals.as<uint8_t>(), dCount);
  checkLastError("compactGroups");
  Scratch order(sizeof(uint32_t) * (size_t)n, s), tmpK(sizeof(uint64_t) * (size_t)n, s), tmpV(sizeof(uint32_t) * (size_t)n, s);
  iotaKernel<<<divUp(n, 256), 256, 0, s>>>(order.as<uint32_t>(), n);
  sortKeyIndexPairs(hash.as<uint64_t>(), order.as<uint32_t>(), tmpK.as<uint64_t>(), tmpV.as<uint32_t>(), n, 64, s);
  Scratch counts(sizeof(uint32_t) * (size_t)n, s), offsets(sizeof(uint32_t) * ((size_t)n + 1), s);
  hllDenseCountKernel<<<n, 256, 0, s>>>(st->table.regs, st->table.acc, slotOf.as<uint32_t>(), order.as<uint32_t>(), counts.as<uint32_t>());
  checkLastError("hllDenseCount");
  hllDenseOffsetsKernel<<<1, 1024, 0, s>>>(counts.as<uint32_t>(), n, offsets.as<uint32_t>());
  checkLastError("hllDenseOffsets");
  uint32_t total = 0;
  ARES_CUDA(cudaMemcpyAsync(&total, offsets.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  out.entries = total;
  if (countOnly || total == 0) return;
  out.block.reset((size_t)st->rowLayout.rowBytes * n, s);
  out.hash.reset(sizeof(uint64_t) * (size_t)total, s);
  out.values.reset(sizeof(uint32_t) * (size_t)total, s);
  out.index.reset(sizeof(uint32_t) * (size_t)total, s);
  DimLayout L = makeDimLayout(st->spec.NumDimsPerDimWidth, n);
  emitGroupsKernel<<<divUp(n, 256), 256, 0, s>>>(st->table, st->keyMode, slotOf.as<uint32_t>(), order.as<uint32_t>(), (uint32_t)n,
                                                 out.block.as<uint8_t>(), L, nullptr);
  checkLastError("emitGroups");
  hllDenseEmitKernel<<<n, 256, 0, s>>>(st->table.regs, st->table.acc, slotOf.as<uint32_t>(), order.as<uint32_t>(), hash.as<uint64_t>(),
                                       offsets.as<uint32_t>(), out.hash.as<uint64_t>(), out.values.as<uint32_t>(),
                                       out.index.as<uint32_t>());
  checkLastError("hllDenseEmit");
  ARES_CUDA(cudaStreamSynchronize(s));  // the locals above are released in stream order; keep it simple
}

// Dense HLL state -> the final outputs of AggStateFinalizeHLL, straight from the register arrays.
static int64_t denseVectors(AggState *st, cudaStream_t s, uint8_t **dimValuesPtr, uint8_t **hllVectorPtr, size_t *hllVectorSizePtr,
                            uint16_t **hllDimRegIDCountPtr) {
  uint32_t c[2];
  ARES_CUDA(cudaMemcpyAsync(c, st->table.counters, sizeof(c), cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  checkOverflow(st, c);
  const int n = (int)c[0];
  if (n == 0) return 0;
  // groups in the order of the reference hash of their dim row
  const int tiles = divUp((int64_t)st->capacity, kCmpTile);
  Scratch state(scanStateBytes(tiles) + sizeof(uint32_t), s);
  ARES_CUDA(cudaMemsetAsync(state.ptr, 0, state.bytes, s));
  ScanTileState sst = makeScanState(state.ptr, tiles);
  uint32_t *dCount = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(state.ptr) + scanStateBytes(tiles));
  Scratch slotOf(sizeof(uint32_t) * (size_t)n, s), hash(sizeof(uint64_t) * (size_t)n, s), vals(sizeof(uint32_t) * (size_t)n, s);
  compactGroupsKernel<<<tiles, kCmpThreads, 0, s>>>(st->table, st->capacity, st->keyMode, 64, ~0ull, st->rowLayout.rowBytes, 4, sst,
                                                   slotOf.as<uint32_t>(), hash.as<uint64_t>(), vals.as<uint8_t>(), dCount);
  checkLastError("compactGroups");
  Scratch order(sizeof(uint32_t) * (size_t)n, s), tmpK(sizeof(uint64_t) * (size_t)n, s), tmpV(sizeof(uint32_t) * (size_t)n, s);
  iotaKernel<<<divUp(n, 256), 256, 0, s>>>(order.as<uint32_t>(), n);
  sortKeyIndexPairs(hash.as<uint64_t>(), order.as<uint32_t>(), tmpK.as<uint64_t>(), tmpV.as<uint32_t>(), n, 64, s);
  Scratch counts(sizeof(uint32_t) * (size_t)n, s), outIdx(sizeof(uint32_t) * (size_t)n, s), byteOff(sizeof(uint32_t) * (size_t)n, s);
  Scratch rowsOf(sizeof(uint32_t) * (size_t)n, s), totalsDev(sizeof(unsigned long long) * 2, s);
  hllDenseCountKernel<<<n, 256, 0, s>>>(st->table.regs, st->table.acc, slotOf.as<uint32_t>(), order.as<uint32_t>(), counts.as<uint32_t>());
  checkLastError("hllDenseCount");
  hllVectorLayoutKernel<<<1, 1024, 0, s>>>(counts.as<uint32_t>(), n, outIdx.as<uint32_t>(), byteOff.as<uint32_t>(), rowsOf.as<uint32_t>(),
                                          totalsDev.as<unsigned long long>());
  checkLastError("hllVectorLayout");
  // the groups' dim rows in hash order (the groups with registers are picked out of it below)
  Scratch blockAll((size_t)st->rowLayout.rowBytes * n, s);
  DimLayout Lin = makeDimLayout(st->spec.NumDimsPerDimWidth, n);
  emitGroupsKernel<<<divUp(n, 256), 256, 0, s>>>(st->table, st->keyMode, slotOf.as<uint32_t>(), order.as<uint32_t>(), (uint32_t)n,
                                                 blockAll.as<uint8_t>(), Lin, nullptr);
  checkLastError("emitGroups");
  unsigned long long totals[2] = {0, 0};
  ARES_CUDA(cudaMemcpyAsync(totals, totalsDev.ptr, sizeof(totals), cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  const int dims = (int)totals[0];
  if (dims == 0) return 0;
  void *vec = nullptr, *cnt = nullptr, *out = nullptr;
  auto fail = [&](CGoCallResHandle h) {
    std::string m(h.pStrErr); free((void *)h.pStrErr);
    if (vec) deviceFree(vec);
    if (cnt) deviceFree(cnt);
    throw EngineError(m);
  };
  CGoCallResHandle h = deviceMalloc(&vec, (size_t)totals[1]);
  if (h.pStrErr) fail(h);
  h = deviceMalloc(&cnt, sizeof(uint16_t) * (size_t)dims);
  if (h.pStrErr) fail(h);
  h = deviceMalloc(&out, (size_t)st->rowLayout.rowBytes * dims);
  if (h.pStrErr) fail(h);
  hllVectorEmitKernel<<<n, 256, 0, s>>>(st->table.regs, st->table.acc, slotOf.as<uint32_t>(), order.as<uint32_t>(), counts.as<uint32_t>(),
                                        outIdx.as<uint32_t>(), byteOff.as<uint32_t>(), static_cast<uint8_t *>(vec),
                                        static_cast<uint16_t *>(cnt));
  checkLastError("hllVectorEmit");
  DimLayout Lout = makeDimLayout(st->spec.NumDimsPerDimWidth, dims);
  gatherDims(blockAll.as<uint8_t>(), Lin, rowsOf.as<uint32_t>(), dims, static_cast<uint8_t *>(out), Lout, s);
  ARES_CUDA(cudaStreamSynchronize(s));
  *dimValuesPtr = static_cast<uint8_t *>(out);
  *hllVectorPtr = static_cast<uint8_t *>(vec);
  *hllVectorSizePtr = (size_t)totals[1];
  *hllDimRegIDCountPtr = static_cast<uint16_t *>(cnt);
  return dims;
}

static int64_t finalize(AggState *st, const DimensionVector &out, uint8_t *outValues, cudaStream_t s, bool ordered = true) {
  for (int i = 0; i < NUM_DIM_WIDTH; i++)
    if (out.NumDimsPerDimWidth[i] != st->spec.NumDimsPerDimWidth[i]) throw EngineError("dimension layout differs from AggSpec");
  if (st->hllDense) {  // carried rows straight from the register arrays (already in key order)
    DenseCarried dc;
    denseCarried(st, s, dc, false);
    if (dc.entries == 0) return 0;
    if (dc.entries > out.VectorCapacity) throw EngineError("output DimensionVector capacity is smaller than the number of rows");
    DimLayout Lin = makeDimLayout(out.NumDimsPerDimWidth, dc.groups), Lout = makeDimLayout(out.NumDimsPerDimWidth, out.VectorCapacity);
    gatherDims(dc.block.as<uint8_t>(), Lin, dc.index.as<uint32_t>(), (int)dc.entries, out.DimValues, Lout, s);
    ARES_CUDA(cudaMemcpyAsync(outValues, dc.values.ptr, sizeof(uint32_t) * (size_t)dc.entries, cudaMemcpyDeviceToDevice, s));
    if (out.HashValues)
      ARES_CUDA(cudaMemcpyAsync(out.HashValues, dc.hash.ptr, sizeof(uint64_t) * (size_t)dc.entries, cudaMemcpyDeviceToDevice, s));
    if (out.IndexVector) iotaKernel<<<divUp(dc.entries, 256), 256, 0, s>>>(out.IndexVector, (int)dc.entries);
    ARES_CUDA(cudaStreamSynchronize(s));
    return dc.entries;
  }
  const int width = st->measWidth;
  const bool plusZero = st->spec.ReduceMode == ARES_REDUCE_HASH && (st->op == OP_SUM_F64 || st->op == OP_SUM_F32);
  if (st->spec.ExpectedGroups <= (uint32_t)kSmallFinalizeMax && out.VectorCapacity > 0) {
    // results of up to 32768 groups: ONE launch (claim list -> hash -> sort -> merge -> emit) and ONE synchronise;
    // the group count comes back through mapped pinned memory
    SmallFinalizeArgs A;
    memset(&A, 0, sizeof(A));
    A.G = st->table;
    A.L = makeDimLayout(out.NumDimsPerDimWidth, out.VectorCapacity);
    A.hashMask = testHash64Mask();
    A.hashA = reinterpret_cast<uint64_t *>(st->smallScratch);
    A.tmpK = A.hashA + kSmallFinalizeMax;
    A.idxA = reinterpret_cast<uint32_t *>(A.tmpK + kSmallFinalizeMax);
    A.tmpI = A.idxA + kSmallFinalizeMax;
    A.hist = A.tmpI + kSmallFinalizeMax;
    A.outBlock = out.DimValues; A.outValues = outValues;
    A.outHash = ordered ? out.HashValues : nullptr; A.outIndex = out.IndexVector;
    A.resultDev = st->resultDev; A.resultHost = st->resultHostDev;
    A.rowBytes = st->rowLayout.rowBytes; A.width = width; A.outCapacity = out.VectorCapacity;
    A.keyMode = st->keyMode; A.hashBits = (uint8_t)st->hashBits; A.op = st->op; A.plusZero = plusZero; A.ordered = ordered;
    finalizeSmallKernel<<<kFinCtas, 1024, 0, s>>>(A);
    checkLastError("finalizeSmall");
    ARES_CUDA(cudaStreamSynchronize(s));
    uint32_t status = st->resultHost[1];
    if (status == SF_UNSETTLED) {   // grow / fold the parked rows, then once more (the table moved: new pointers)
      settleTable(st, s);
      return finalize(st, out, outValues, s, ordered);
    }
    if (status == SF_OK) return st->resultHost[0];
    if (status == SF_TABLE_OVERFLOW) { const uint32_t c[2] = {st->resultHost[2], 1}; checkOverflow(st, c); }
    if (status == SF_OUTPUT_TOO_SMALL) throw EngineError("output DimensionVector capacity is smaller than the number of groups");
    if (status == SF_PEER_LATE) throw EngineError("exchange over peer memory: a peer's part did not arrive within the wait bound");
    if (status == SF_PART_TRUNCATED) throw EngineError("exchange part truncated: a rank held more rows than the fixed part carries; repeat the exchange with exact sizes");
    // SF_TOO_MANY: more groups than one CTA sorts — the multi-launch path below
  }
  const int64_t occupied = groupCount(st, s);
  if (occupied == 0) return 0;
  const int n = (int)occupied;
  // 1. the claimed slots (claim order) with their reference hash and accumulator
  Scratch slotOf(sizeof(uint32_t) * (size_t)n, s), hash(sizeof(uint64_t) * (size_t)n, s), vals((size_t)width * n, s);
  {
    int blocks = divUp(n, 256);
    if (blocks > smCount() * 8) blocks = smCount() * 8;
    gatherClaimedKernel<<<blocks, 256, 0, s>>>(st->table, (uint32_t)n, st->keyMode, (uint8_t)st->hashBits, testHash64Mask(),
                                               st->rowLayout.rowBytes, width, slotOf.as<uint32_t>(), hash.as<uint64_t>(),
                                               vals.as<uint8_t>());
  }
  checkLastError("gatherClaimed");
  // 2. sort the groups by hash (stable), 3. merge equal hashes (reference group identity)
  Scratch order(sizeof(uint32_t) * (size_t)n, s), tmpK(sizeof(uint64_t) * (size_t)n, s), tmpV(sizeof(uint32_t) * (size_t)n, s);
  iotaKernel<<<divUp(n, 256), 256, 0, s>>>(order.as<uint32_t>(), n);
  if (!ordered) {
    // exchange form: the occupied slots as they are (no hash order, no merge of colliding hashes —
    // the receiving state's own finalize does both)
    if (n > out.VectorCapacity) throw EngineError("output DimensionVector capacity is smaller than the number of groups");
    DimLayout Lx = makeDimLayout(out.NumDimsPerDimWidth, out.VectorCapacity);
    emitGroupsKernel<<<divUp(n, 256), 256, 0, s>>>(st->table, st->keyMode, slotOf.as<uint32_t>(), order.as<uint32_t>(), (uint32_t)n,
                                                   out.DimValues, Lx, out.IndexVector);
    checkLastError("emitGroups");
    ARES_CUDA(cudaMemcpyAsync(outValues, vals.ptr, (size_t)width * n, cudaMemcpyDeviceToDevice, s));
    ARES_CUDA(cudaStreamSynchronize(s));
    return n;
  }
  sortKeyIndexPairs(hash.as<uint64_t>(), order.as<uint32_t>(), tmpK.as<uint64_t>(), tmpV.as<uint32_t>(), n, st->hashBits, s);
  Scratch rep(sizeof(uint32_t) * (size_t)n, s);
  Scratch mergedVals((size_t)width * n, s);
  const int g = reduceByHash(hash.as<uint64_t>(), order.as<uint32_t>(), vals.as<uint8_t>(), width, st->op, n,
                             rep.as<uint32_t>(), mergedVals.as<uint8_t>(), s,
                             n <= out.VectorCapacity ? out.HashValues : nullptr);
  if (g > out.VectorCapacity) throw EngineError("output DimensionVector capacity is smaller than the number of groups");
  // 4. emit in the reference's layout
  DimLayout L = makeDimLayout(out.NumDimsPerDimWidth, out.VectorCapacity);
  int blocks = divUp(g, 256);
  emitGroupsKernel<<<blocks, 256, 0, s>>>(st->table, st->keyMode, slotOf.as<uint32_t>(), rep.as<uint32_t>(), (uint32_t)g,
                                          out.DimValues, L, out.IndexVector);
  checkLastError("emitGroups");
  if (plusZero) {
    // The reference's hash map folds every value into a slot that starts at the identity +0.0
    // (query/hash_reduction.cu:246-249 + the map's unused element), so a group whose values are all -0.0 ends
    // at +0.0 there, while its sort-reduce (and this table, whose neutral element is -0.0) keeps -0.0.
    copyPlusZeroKernel<<<divUp(g, 256), 256, 0, s>>>(mergedVals.as<uint8_t>(), outValues, g, width);
    checkLastError("copyPlusZero");
  } else {
    ARES_CUDA(cudaMemcpyAsync(outValues, mergedVals.ptr, (size_t)width * g, cudaMemcpyDeviceToDevice, s));
  }
  ARES_CUDA(cudaStreamSynchronize(s));
  return g;
}

// HLL state -> the reference's final outputs.  AggStateFinalize on such a state already yields the
// carried form (one row per (group, register) entry, key-ascending, value = max rho << 16 | reg);
// this runs it into scratch buffers and then the shared register-vector stage of hll.cu.  The dim
// block (VectorCapacity == number of groups) and the two vectors are allocated with deviceMalloc.
static int64_t finalizeHLL(AggState *st, uint8_t **dimValuesPtr, uint8_t **hllVectorPtr, size_t *hllVectorSizePtr,
                           uint16_t **hllDimRegIDCountPtr, cudaStream_t s) {
  if (!st->hll) throw EngineError("AggStateFinalizeHLL needs a state created with AGGR_HLL");
  if (!dimValuesPtr || !hllVectorPtr || !hllVectorSizePtr || !hllDimRegIDCountPtr) throw EngineError("null output pointer");
  *dimValuesPtr = nullptr; *hllVectorPtr = nullptr; *hllVectorSizePtr = 0; *hllDimRegIDCountPtr = nullptr;
  if (st->hllDense) return denseVectors(st, s, dimValuesPtr, hllVectorPtr, hllVectorSizePtr, hllDimRegIDCountPtr);
  const int64_t entries = groupCount(st, s);
  if (entries == 0) return 0;
  const int n = (int)entries;
  const int rowBytes = st->rowLayout.rowBytes;
  Scratch block((size_t)rowBytes * n, s), hash(sizeof(uint64_t) * (size_t)n, s), index(sizeof(uint32_t) * (size_t)n, s);
  Scratch values(sizeof(uint32_t) * (size_t)n, s);
  DimensionVector carried;
  carried.DimValues = block.as<uint8_t>();
  carried.HashValues = hash.as<uint64_t>();
  carried.IndexVector = index.as<uint32_t>();
  carried.VectorCapacity = n;
  for (int i = 0; i < NUM_DIM_WIDTH; i++) carried.NumDimsPerDimWidth[i] = st->spec.NumDimsPerDimWidth[i];
  const int64_t g = finalize(st, carried, values.as<uint8_t>(), s);
  if (g != entries) throw EngineError("HLL state: duplicate keys in the group table");
  const int dims = hllRegisterVectors(hash.as<uint64_t>(), values.as<uint32_t>(), index.as<uint32_t>(), n, hllVectorPtr,
                                      hllVectorSizePtr, hllDimRegIDCountPtr, s);
  void *out = nullptr;
  CGoCallResHandle h = deviceMalloc(&out, (size_t)rowBytes * dims);
  if (h.pStrErr) {
    std::string m(h.pStrErr); free((void *)h.pStrErr);
    deviceFree(*hllVectorPtr); deviceFree(*hllDimRegIDCountPtr);
    *hllVectorPtr = nullptr; *hllDimRegIDCountPtr = nullptr;
    throw EngineError(m);
  }
  DimLayout Lin = makeDimLayout(carried.NumDimsPerDimWidth, n), Lout = makeDimLayout(carried.NumDimsPerDimWidth, dims);
  gatherDims(block.as<uint8_t>(), Lin, index.as<uint32_t>(), dims, static_cast<uint8_t *>(out), Lout, s);
  ARES_CUDA(cudaStreamSynchronize(s));
  *dimValuesPtr = static_cast<uint8_t *>(out);
  return dims;
}

}  // namespace aresb

using namespace aresb;

extern "C" {

CGoCallResHandle AggStateCreate(AggSpec spec, void *cudaStream, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  try {
    ARES_CUDA(cudaSetDevice(device));
    h.res = createState(spec, (cudaStream_t)cudaStream, device);
  } catch (const std::exception &e) {
    h.pStrErr = strdup((std::string("AggStateCreate: ") + e.what()).c_str());
  }
  return h;
}

CGoCallResHandle ExecuteBatchPlan(void *state, const BatchPlan *plan, void *cudaStream, int device) {
  return guarded("ExecuteBatchPlan", device, [&]() -> int64_t {
    if (!plan) throw EngineError("null plan");
    executePlan(asState(state), *plan, (cudaStream_t)cudaStream);
    return 0;
  });
}

CGoCallResHandle AggStateMerge(void *state, DimensionVector inputKeys, uint8_t *inputValues, int length,
                               void *cudaStream, int device) {
  return guarded("AggStateMerge", device, [&]() -> int64_t {
    mergeRows(asState(state), inputKeys, inputValues, length, (cudaStream_t)cudaStream);
    return 0;
  });
}

CGoCallResHandle AggStateGroupCount(void *state, void *cudaStream, int device) {
  return guarded("AggStateGroupCount", device, [&]() -> int64_t {
    AggState *st = asState(state);
    if (st->hllDense) {  // rows of the carried form = registers that were hit
      DenseCarried dc;
      denseCarried(st, (cudaStream_t)cudaStream, dc, true);
      return dc.entries;
    }
    return groupCount(st, (cudaStream_t)cudaStream);
  });
}

CGoCallResHandle AggStateFinalize(void *state, DimensionVector outputKeys, uint8_t *outputValues, void *cudaStream,
                                  int device) {
  return guarded("AggStateFinalize", device, [&]() -> int64_t {
    return finalize(asState(state), outputKeys, outputValues, (cudaStream_t)cudaStream);
  });
}

CGoCallResHandle AggStateExport(void *state, DimensionVector outputKeys, uint8_t *outputValues, void *cudaStream,
                                int device) {
  return guarded("AggStateExport", device, [&]() -> int64_t {
    return finalize(asState(state), outputKeys, outputValues, (cudaStream_t)cudaStream, false);
  });
}

// Exchange form without host involvement (sharded queries): the claimed slots go to `part` = [uint32 rows, status,
// claimed | ... | dimension block of capRows rows at dimOffset | measures at valuesOffset]; nothing is synchronised and
// the row count stays on the device.  More rows than capRows (or than one CTA handles): status != 0, no rows.
CGoCallResHandle AggStateExportPart(void *state, uint8_t *part, int capRows, size_t dimOffset, size_t valuesOffset,
                                    void *cudaStream, int device) {
  return guarded("AggStateExportPart", device, [&]() -> int64_t {
    AggState *st = asState(state);
    if (st->hll) throw EngineError("AggStateExportPart: HLL states exchange through AggStateExport");
    if (capRows <= 0 || capRows > kSmallFinalizeMax) throw EngineError("AggStateExportPart: capRows must be in [1, 32768]");
    SmallFinalizeArgs A;
    memset(&A, 0, sizeof(A));
    A.G = st->table;
    A.L = makeDimLayout(st->spec.NumDimsPerDimWidth, capRows);
    A.hashMask = ~0ull;
    A.outBlock = part + dimOffset; A.outValues = part + valuesOffset;
    A.resultDev = reinterpret_cast<uint32_t *>(part); A.resultHost = st->resultHostDev + 4;   // host copy unused
    A.rowBytes = st->rowLayout.rowBytes; A.width = st->measWidth; A.outCapacity = capRows;
    A.keyMode = st->keyMode; A.hashBits = (uint8_t)st->hashBits; A.op = st->op; A.plusZero = 0; A.ordered = 0;
    finalizeSmallKernel<<<kFinCtas, 1024, 0, (cudaStream_t)cudaStream>>>(A);
    checkLastError("AggStateExportPart");
    return 0;
  });
}

// Receiving side: folds `numParts` parts laid out `partStride` bytes apart (as all_gather leaves them) into `state` with
// one launch; asynchronous.  A truncated part is reported by the next AggStateFinalize ("exchange part truncated").
CGoCallResHandle AggStateMergeParts(void *state, const uint8_t *parts, int numParts, size_t partStride, int capRows,
                                    size_t dimOffset, size_t valuesOffset, void *cudaStream, int device) {
  return guarded("AggStateMergeParts", device, [&]() -> int64_t {
    AggState *st = asState(state);
    if (numParts <= 0) return 0;
    ensureRoom(st, (uint64_t)numParts * (uint64_t)capRows, (cudaStream_t)cudaStream);
    DimLayout L = makeDimLayout(st->spec.NumDimsPerDimWidth, capRows);
    mergePartsKernel<<<smCount() * 2, 256, 0, (cudaStream_t)cudaStream>>>(parts, numParts, partStride, dimOffset, valuesOffset, L,
                                                                         st->measWidth, st->op, st->keyMode, (uint8_t)st->hashBits,
                                                                         st->hll ? (st->hllDense ? 2 : 1) : 0, st->table, nullptr, 0u);
    checkLastError("AggStateMergeParts");
    return 0;
  });
}

// Exchange over peer memory (sharded queries on one NVLink / NVSwitch node).  `peerSlots[r]` = the address, in THIS
// process, of this rank's part slot inside rank r's receive buffer; `peerFlags[r]` = the address of flags[myRank] on rank
// r (the host maps the peers' buffers: CUDA IPC / fabric handles, e.g. torch symmetric memory).  One launch; asynchronous.
CGoCallResHandle AggStateExportPartToPeers(void *state, uint8_t *const *peerSlots, uint32_t *const *peerFlags, int numPeers, int myRank,
                                           size_t partBytes, int capRows, size_t dimOffset, size_t valuesOffset, uint32_t epoch,
                                           void *cudaStream, int device) {
  return guarded("AggStateExportPartToPeers", device, [&]() -> int64_t {
    AggState *st = asState(state);
    if (st->hll) throw EngineError("AggStateExportPartToPeers: HLL states exchange through AggStateExport");
    if (capRows <= 0 || capRows > kSmallFinalizeMax) throw EngineError("AggStateExportPartToPeers: capRows must be in [1, 32768]");
    if (numPeers <= 0 || numPeers > 16 || myRank < 0 || myRank >= numPeers) throw EngineError("AggStateExportPartToPeers: 1..16 peers");
    if (partBytes % 16 != 0) throw EngineError("AggStateExportPartToPeers: partBytes must be a multiple of 16");
    PeerExportArgs E;
    memset(&E, 0, sizeof(E));
    for (int r = 0; r < numPeers; r++) { E.peerSlot[r] = peerSlots[r]; E.peerFlag[r] = peerFlags[r]; }
    uint8_t *part = peerSlots[myRank];
    E.F.G = st->table;
    E.F.L = makeDimLayout(st->spec.NumDimsPerDimWidth, capRows);
    E.F.outBlock = part + dimOffset; E.F.outValues = part + valuesOffset;
    E.F.resultDev = reinterpret_cast<uint32_t *>(part);
    E.F.rowBytes = st->rowLayout.rowBytes; E.F.width = st->measWidth; E.F.outCapacity = capRows;
    E.F.keyMode = st->keyMode; E.F.hashBits = (uint8_t)st->hashBits; E.F.op = st->op;
    E.partBytes = partBytes; E.numPeers = (uint32_t)numPeers; E.myRank = (uint32_t)myRank; E.epoch = epoch;
    exportToPeersKernel<<<kFinCtas, 1024, 0, (cudaStream_t)cudaStream>>>(E);
    checkLastError("AggStateExportPartToPeers");
    return 0;
  });
}

// Receiving side: as AggStateMergeParts, but the merge kernel itself waits (bounded) until every peer has stored `epoch`
// into flags[peer].  A peer that does not arrive is reported by the next AggStateFinalize.
CGoCallResHandle AggStateMergePartsWhenFlagged(void *state, const uint8_t *parts, int numParts, size_t partStride, int capRows,
                                               size_t dimOffset, size_t valuesOffset, const uint32_t *flags, uint32_t epoch,
                                               void *cudaStream, int device) {
  return guarded("AggStateMergePartsWhenFlagged", device, [&]() -> int64_t {
    AggState *st = asState(state);
    if (numParts <= 0) return 0;
    if (flags == nullptr) throw EngineError("AggStateMergePartsWhenFlagged: flags is null");
    ensureRoom(st, (uint64_t)numParts * (uint64_t)capRows, (cudaStream_t)cudaStream);
    DimLayout L = makeDimLayout(st->spec.NumDimsPerDimWidth, capRows);
    mergePartsKernel<<<smCount() * 2, 256, 0, (cudaStream_t)cudaStream>>>(parts, numParts, partStride, dimOffset, valuesOffset, L,
                                                                         st->measWidth, st->op, st->keyMode, (uint8_t)st->hashBits,
                                                                         0, st->table, flags, epoch);
    checkLastError("AggStateMergePartsWhenFlagged");
    return 0;
  });
}

CGoCallResHandle AggStateFinalizeHLL(void *state, uint8_t **dimValuesPtr, uint8_t **hllVectorPtr, size_t *hllVectorSizePtr,
                                     uint16_t **hllDimRegIDCountPtr, void *cudaStream, int device) {
  return guarded("AggStateFinalizeHLL", device, [&]() -> int64_t {
    return finalizeHLL(asState(state), dimValuesPtr, hllVectorPtr, hllVectorSizePtr, hllDimRegIDCountPtr,
                       (cudaStream_t)cudaStream);
  });
}

CGoCallResHandle AggStateReset(void *state, void *cudaStream, int device) {
  return guarded("AggStateReset", device, [&]() -> int64_t {
    AggState *st = asState(state);
    cudaStream_t s = (cudaStream_t)cudaStream;
    // only what was claimed is emptied (claim list): a 2^21-slot table with 19,200 groups resets 0.3 MB, not 32 MB,
    // and a dense HLL state clears the register arrays of its groups, not all 512 MB
    resetClaimedKernel<<<smCount() * 4, 256, 0, s>>>(st->table, st->accNeutral);
    checkLastError("AggStateReset");
    ARES_CUDA(cudaMemsetAsync(st->table.counters, 0, 256, s));
    st->occUpper = 0;
    st->unchecked = 0;
    st->everChecked = false;
    st->resultHost[8] = 0;
    return 0;
  });
}

// Additive diagnostics, usable without a GPU: generates + NVRTC-compiles the specialised kernel of
// (spec, plan) and returns the cubin size in res (0: plan not eligible); *sourceOut (optional) gets a
// malloc'd copy of the generated shape-specific source.
CGoCallResHandle AresJitDryRun(AggSpec spec, const BatchPlan *plan, char **sourceOut) {
  CGoCallResHandle h = {nullptr, nullptr};
  try {
    AggState st;
    memset(&st.table, 0, sizeof(st.table));
    describeState(&st, spec);
    st.capacity = st.hllDense ? kHllDenseSlots : 0;
    static thread_local DevPlan P;
    compilePlan(&st, *plan, P);
    P.tailBegin = 0;
    if (const char *e = getenv("ARESDB_B200_PARTITION")) {   // the partitioned form of the kernel (normally chosen by table size)
      P.partition = e[0] == '1' && st.keyMode == KEY_PACKED && !st.hll && P.numForeignCols == 0;
      P.partShift = 15;
    }
    layoutStages(P, spec.ExpectedGroups);
    std::string src;
    size_t n = P.staged ? jitCompileOnly(P, &src) : 0;
    if (sourceOut) *sourceOut = strdup(src.c_str());
    h.res = reinterpret_cast<void *>(n);
  } catch (const std::exception &e) {
    h.pStrErr = strdup((std::string("AresJitDryRun: ") + e.what()).c_str());
  }
  return h;
}

CGoCallResHandle AggStateDestroy(void *state, int device) {
  return guarded("AggStateDestroy", device, [&]() -> int64_t {
    AggState *st = asState(state);
    if (st->mem) deviceFree(st->mem);
    if (st->resultHost) cudaFreeHost(st->resultHost);
    if (st->denseAcc) deviceFree(st->denseAcc);
    delete st;
    return 0;
  });
}

}  // extern "C"
