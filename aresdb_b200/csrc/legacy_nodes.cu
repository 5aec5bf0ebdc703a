// legacy_nodes.cu — the per-AST-node entry points the unchanged Go driver calls:
// InitIndexVector, UnaryTransform, BinaryTransform, UnaryFilter, BinaryFilter (reference:
// query/algorithm.cu:22-41, query/transform.cu:21-86, query/filter.cu:130-253) and the
// out-of-scope symbols that must still link.  One hand-written kernel per call; the
// filter is a single-pass stable in-place compaction (decoupled look-back scan) instead of
// the reference's transform + remove_if pair.  The whole-batch fused path is batch_plan.cu.
#include <vector>

#include "column.cuh"
#include "common.cuh"
#include "join.cuh"
#include "scan.cuh"

namespace aresb {

// ---------------------------------------------------------------------------------------
// host: ABI structs -> device descriptors
// ---------------------------------------------------------------------------------------
static ValClass columnClass(int dtype) {
  switch (dtype) {
    case Bool: return VC_BOOL;
    case Int8: case Int16: case Int32: return VC_I32;
    case Uint8: case Uint16: case Uint32: return VC_U32;
    case Float32: return VC_F32;
    case Int64: return VC_I64;
    case UUID: return VC_UUID;
    default: throw EngineError("Unsupported data type for VectorPartyInput");
  }
}

InputDesc makeColumnDesc(const VectorPartySlice &vp, bool allowWide) {
  InputDesc d;
  memset(&d, 0, sizeof(d));
  d.kind = IN_COLUMN;
  d.dtype = (uint8_t)vp.DataType;
  d.vclass = columnClass(vp.DataType);
  if ((d.vclass == VC_I64 || d.vclass == VC_UUID) && !allowWide)
    throw EngineError("int64/UUID data types are only supported in UnaryTransform");
  d.base = vp.BasePtr; d.nullsOff = vp.NullsOffset; d.valuesOff = vp.ValuesOffset;
  d.length = vp.Length; d.startBit = vp.StartingIndex;
  if (vp.BasePtr == nullptr) {
    d.mode = 0;
    d.constValid = vp.DefaultValue.HasDefault;
    switch (d.vclass) {
      case VC_BOOL: d.constLo = vp.DefaultValue.Value.BoolVal ? 1 : 0; break;
      case VC_I64: d.constLo = (uint64_t)vp.DefaultValue.Value.Int64Val; break;
      case VC_UUID: d.constLo = vp.DefaultValue.Value.UUIDVal.p1; d.constHi = vp.DefaultValue.Value.UUIDVal.p2; break;
      default: d.constLo = vp.DefaultValue.Value.Uint32Val; break;  // int32 / uint32 / float share the bits
    }
  } else {
    d.mode = vp.ValuesOffset == 0 ? 1 : (vp.NullsOffset == 0 ? 2 : 3);
  }
  return d;
}

InputDesc makeInputDesc(const InputVector &in, bool allowWide) {
  InputDesc d;
  memset(&d, 0, sizeof(d));
  switch (in.Type) {
    case ConstantInput: {
      const ConstantVector &c = in.Vector.Constant;
      d.kind = IN_CONST;
      d.constValid = c.IsValid;
      if (c.DataType == ConstInt) { d.vclass = VC_I32; d.constLo = (uint32_t)c.Value.IntVal; }
      else if (c.DataType == ConstFloat) { d.vclass = VC_F32; uint32_t b; memcpy(&b, &c.Value.FloatVal, 4); d.constLo = b; }
      else throw EngineError("Unsupported constant type (GeoPoint/UUID constants are outside the hot path)");
      return d;
    }
    case ScratchSpaceInput: {
      const ScratchSpaceVector &s = in.Vector.ScratchSpace;
      d.kind = IN_SCRATCH;
      d.base = s.Values; d.nullsOff = s.NullsOffset; d.dtype = (uint8_t)s.DataType;
      switch (s.DataType) {
        case Int32: d.vclass = VC_I32; break;
        case Uint32: d.vclass = VC_U32; break;
        case Float32: d.vclass = VC_F32; break;
        case UUID: d.vclass = VC_UUID; break;
        default: throw EngineError("Unsupported data type for ScratchSpaceInput");
      }
      if (d.vclass == VC_UUID && !allowWide) throw EngineError("UUID operand is only supported as a unary root input");
      return d;
    }
    case VectorPartyInput:
      return makeColumnDesc(in.Vector.VP, allowWide);
    case ForeignColumnInput: {   // the descriptor proper (batches, RecordIDs, timezone table) is a ForeignDesc: makeForeignDesc
      const ForeignColumnVector &f = in.Vector.ForeignVP;
      d.kind = IN_FOREIGN;
      d.dtype = (uint8_t)f.DataType;
      d.vclass = columnClass(f.DataType);
      if ((d.vclass == VC_I64 || d.vclass == VC_UUID) && !allowWide)
        throw EngineError("int64/UUID data types are only supported in UnaryTransform");
      return d;
    }
    default:
      throw EngineError("ArrayVectorPartyInput (array columns) is outside the B200 hot path");
  }
}

// ForeignColumnInput -> device descriptor.  `Batches` points to HOST memory valid for the duration of the call (Go heap,
// query/time_series_aggregate.go:100-112): the slices are copied by value into the kernel parameters.
ForeignDesc makeForeignDesc(const ForeignColumnVector &f) {
  ForeignDesc F;
  memset(&F, 0, sizeof(F));
  if (f.NumBatches < 0 || f.NumBatches > kMaxForeignBatches)
    throw EngineError("a joined dimension table may hold at most " + std::to_string(kMaxForeignBatches) + " batches");
  if (f.NumBatches > 0 && f.Batches == nullptr) throw EngineError("ForeignColumnVector.Batches is NULL");
  F.recordIDs = reinterpret_cast<const unsigned long long *>(f.RecordIDs);
  F.tzLookup = f.TimezoneLookup;
  F.tzSize = f.TimezoneLookupSize;
  F.numBatches = f.NumBatches; F.baseBatchID = f.BaseBatchID; F.numRecordsInLastBatch = f.NumRecordsInLastBatch;
  for (int b = 0; b < f.NumBatches; b++) {
    VectorPartySlice vp = f.Batches[b];
    if (vp.BasePtr == nullptr) {   // mode 0: the COLUMN's default (prepareForeignTableIterators passes hasDefault / defaultValue)
      vp.DefaultValue = f.DefaultValue;
    }
    vp.DataType = f.DataType;
    F.batches[b] = makeColumnDesc(vp, /*allowWide=*/true);
    if (F.batches[b].mode == 3) throw EngineError("dimension-table columns are not compressed (mode 3 foreign batch)");
  }
  return F;
}

enum SinkKind : uint8_t { SINK_SCRATCH, SINK_DIMENSION, SINK_MEASURE, SINK_PREDICATE };

struct SinkDesc {
  uint8_t *values;
  uint8_t *nulls;          // scratch / dimension validity bytes
  uint8_t kind;
  uint8_t oclass;          // ValClass of one output element
  uint8_t width;           // bytes per output element
  uint8_t isAvg, skipCount;
  int32_t aggFunc;
  uint64_t identity;
};

static SinkDesc makeSink(const OutputVector &out) {
  SinkDesc s;
  memset(&s, 0, sizeof(s));
  auto cls = [&](int dt, bool dim) -> ValClass {
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
      case UUID: return VC_UUID;
      default: break;
    }
    return VC_NONE;
  };
  auto widthOf = [](ValClass c) -> int {
    switch (c) {
      case VC_BOOL: case VC_I8: case VC_U8: return 1;
      case VC_I16: case VC_U16: return 2;
      case VC_I32: case VC_U32: case VC_F32: return 4;
      case VC_I64: case VC_F64: return 8;
      case VC_UUID: return 16;
      default: return 0;
    }
  };
  switch (out.Type) {
    case ScratchSpaceOutput: {
      const ScratchSpaceVector &v = out.Vector.ScratchSpace;
      ValClass c = cls(v.DataType, false);
      if (c == VC_NONE || c == VC_F64) throw EngineError("Unsupported data type for ScratchSpaceOutput");
      s.kind = SINK_SCRATCH; s.values = v.Values; s.nulls = v.Values + v.NullsOffset;
      s.oclass = c; s.width = widthOf(c);
      break;
    }
    case DimensionOutput: {
      const DimensionOutputVector &v = out.Vector.Dimension;
      ValClass c = cls(v.DataType, true);
      if (c == VC_NONE) throw EngineError("Unsupported data type for DimensionOutput");
      s.kind = SINK_DIMENSION; s.values = v.DimValues; s.nulls = v.DimNulls;
      s.oclass = c; s.width = widthOf(c);
      break;
    }
    case MeasureOutput: {
      const MeasureOutputVector &v = out.Vector.Measure;
      ValClass c = cls(v.DataType, false);
      if (c == VC_NONE || c == VC_UUID) throw EngineError("Unsupported data type for MeasureOutput");
      s.kind = SINK_MEASURE; s.values = reinterpret_cast<uint8_t *>(v.Values);
      s.oclass = c; s.width = widthOf(c);
      s.aggFunc = v.AggFunc;
      s.isAvg = v.AggFunc == AGGR_AVG_FLOAT;
      s.skipCount = !((v.AggFunc >= AGGR_SUM_UNSIGNED && v.AggFunc <= AGGR_SUM_FLOAT) || s.isAvg);
      s.identity = aggIdentity(v.AggFunc, c);
      break;
    }
    default: throw EngineError("Unsupported output vector type");
  }
  return s;
}

struct NodeDesc {
  InputDesc in[2];
  ForeignDesc fd[2];       // in[k].kind == IN_FOREIGN: its batches / RecordIDs
  const uint32_t *index;
  const uint32_t *baseCounts;
  uint32_t startCount;
  int32_t n;
  int32_t fn;
  uint8_t nin;
  uint8_t tclass;  // binary: common class; unary: class of the operand
};

// ---------------------------------------------------------------------------------------
// device: evaluate one node for one row
// ---------------------------------------------------------------------------------------
// Unary functors on an int64 operand (UnaryFunctor<O, int64_t>, generic template).
__device__ __forceinline__ Cell evalUnaryI64(int fn, Cell a, ValClass *rc) {
  Cell r; r.v = 0; r.valid = false;
  switch (fn) {
    case Not: *rc = VC_BOOL; if (!a.valid) return r; r.v = a.v ? 0 : 1; r.valid = true; return r;
    case IsNull: *rc = VC_BOOL; r.v = a.valid ? 0 : 1; r.valid = true; return r;
    case IsNotNull: *rc = VC_BOOL; r.v = a.valid ? 1 : 0; r.valid = true; return r;
    case Negate: *rc = VC_I64; if (!a.valid) return r; r.v = 0 - a.v; r.valid = true; return r;
    case BitwiseNot: *rc = VC_I64; if (!a.valid) return r; r.v = ~a.v; r.valid = true; return r;
    case GetHLLValue: {
      *rc = VC_U32; if (!a.valid) return r;
      uint64_t w[4] = {a.v, 0, 0, 0};
      r.v = hllValueOfHash(murmur3_128_lo(w, 8, 0)); r.valid = true; return r;
    }
    case GetWeekStart: case GetMonthStart: case GetQuarterStart: case GetYearStart:
    case GetDayOfMonth: case GetDayOfYear: case GetMonthOfYear: case GetQuarterOfYear: {
      Cell t; t.v = (uint32_t)a.v; t.valid = a.valid;
      return evalUnary(fn, t, VC_U32, rc);
    }
    default: *rc = VC_I64; return a;
  }
}

// Operand k of row position i: a foreign column is read at the RecordID the join left for THIS position
// (RecordIDJoinIterator is positional over the RecordID vector, zipped with the index vector: query/binder.hpp:147-176).
__device__ __forceinline__ Cell loadNodeInput(const NodeDesc &nd, int k, uint32_t i, uint64_t *hi) {
  if (nd.in[k].kind == IN_FOREIGN) return foreignLoad(nd.fd[k], nd.fd[k].recordIDs[i], hi);
  return loadInput(nd.in[k], i, nd.index, nd.baseCounts, nd.startCount, hi);
}

__device__ __forceinline__ Cell evalNode(const NodeDesc &nd, uint32_t i, ValClass *rc, uint64_t *hi) {
  if (nd.nin == 1) {
    Cell a = loadNodeInput(nd, 0, i, hi);
    ValClass ic = (ValClass)nd.in[0].vclass;
    if (ic == VC_I64) return evalUnaryI64(nd.fn, a, rc);
    if (ic == VC_UUID) {
      if (nd.fn == GetHLLValue) {  // hll_hash(UUID) = p1 ^ p2 (reference query/functor.hpp:439-442)
        Cell r; r.valid = a.valid; r.v = a.valid ? hllValueOfHash(a.v ^ *hi) : 0; *rc = VC_U32; return r;
      }
      *rc = VC_UUID; return a;  // UUID -> UUID sinks copy; every other sink yields NULL (handled by caller)
    }
    return evalUnary(nd.fn, a, ic, rc);
  }
  Cell a = loadNodeInput(nd, 0, i, nullptr);
  Cell b = loadNodeInput(nd, 1, i, nullptr);
  ValClass tc = (ValClass)nd.tclass;
  a.v = cvt(a.v, (ValClass)nd.in[0].vclass, tc);
  b.v = cvt(b.v, (ValClass)nd.in[1].vclass, tc);
  return evalBinary(nd.fn, a, b, tc, rc);
}

__device__ __forceinline__ void storeSized(uint8_t *p, uint64_t v, int width) {
  switch (width) {
    case 1: *p = (uint8_t)v; break;
    case 2: *reinterpret_cast<uint16_t *>(p) = (uint16_t)v; break;
    case 4: *reinterpret_cast<uint32_t *>(p) = (uint32_t)v; break;
    default: *reinterpret_cast<uint64_t *>(p) = v; break;
  }
}

__global__ void __launch_bounds__(256) transformKernel(NodeDesc nd, SinkDesc sk) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)nd.n; i += stride) {
    ValClass rc; uint64_t hi = 0;
    Cell r = evalNode(nd, i, &rc, &hi);
    ValClass oc = (ValClass)sk.oclass;
    if (rc == VC_UUID || oc == VC_UUID) {
      // Only UUID -> UUID is a copy; any other pairing with a UUID is "(garbage, false)".
      if (rc == VC_UUID && oc == VC_UUID && sk.kind != SINK_MEASURE) {
        reinterpret_cast<uint64_t *>(sk.values)[2 * (size_t)i] = r.v;
        reinterpret_cast<uint64_t *>(sk.values)[2 * (size_t)i + 1] = hi;
        sk.nulls[i] = r.valid;
        continue;
      }
      r.v = 0; r.valid = false; rc = oc;
      if (oc == VC_UUID) {
        reinterpret_cast<uint64_t *>(sk.values)[2 * (size_t)i] = 0;
        reinterpret_cast<uint64_t *>(sk.values)[2 * (size_t)i + 1] = 0;
        sk.nulls[i] = 0;
        continue;
      }
    }
    uint64_t v = cvt(r.v, rc, oc);
    if (sk.kind == SINK_MEASURE) {
      // NULL -> identity of the aggregate; SUM/AVG rows of an RLE batch count `count` times
      // (reference query/iterator.hpp:616-727).
      uint8_t *out = sk.values + (size_t)i * sk.width;
      if (!r.valid) { storeSized(out, sk.identity, sk.width); continue; }
      uint32_t count = 1;
      if (!sk.skipCount && nd.baseCounts != nullptr) {
        uint32_t idx = nd.index[i];
        count = nd.baseCounts[idx + 1] - nd.baseCounts[idx];
      }
      if (sk.isAvg) {
        float f = asF32(cvt(v, oc, VC_F32));
        reinterpret_cast<uint32_t *>(out)[0] = (uint32_t)fromF32(f);
        reinterpret_cast<uint32_t *>(out)[1] = count;
        continue;
      }
      switch (oc) {
        case VC_I32: case VC_U32: v = (uint32_t)v * count; break;
        case VC_F32: v = fromF32(asF32(v) * (float)count); break;
        case VC_I64: v = (uint64_t)((int64_t)v * (int64_t)(uint64_t)count); break;
        default: v = fromF64(asF64(v) * (double)count); break;
      }
      storeSized(out, v, sk.width);
    } else {
      storeSized(sk.values + (size_t)i * sk.width, v, sk.width);
      sk.nulls[i] = r.valid ? 1 : 0;
    }
  }
}

// Filter: predicate = value part of f(...) converted to bool (validity is NOT consulted —
// comparisons already yield (false, false) on NULL; Noop on a NULL bool passes its stored
// bit, as the reference does, query/functor.hpp:905-921), then stable in-place compaction.
constexpr int kFilterThreads = 256;
constexpr int kFilterItems = 4;
constexpr int kFilterTile = kFilterThreads * kFilterItems;

__global__ void __launch_bounds__(kFilterThreads)
filterKernel(NodeDesc nd, uint32_t *index, uint8_t *predicate, ScanTileState st, uint32_t *outCount,
             RecordID **recIn, RecordID **recOut, int numForeign) {
  __shared__ uint32_t sTile;
  __shared__ uint32_t sWarp[kFilterThreads / 32 + 1];
  __shared__ uint32_t sPrefix;
  if (threadIdx.x == 0) sTile = atomicAdd(st.ticket, 1u);
  __syncthreads();
  const uint32_t tile = sTile;
  const uint32_t base = tile * kFilterTile + threadIdx.x * kFilterItems;
  uint32_t vals[kFilterItems];
  uint32_t keepMask = 0;
#pragma unroll
  for (int k = 0; k < kFilterItems; k++) {
    uint32_t i = base + k;
    if (i < (uint32_t)nd.n) {
      ValClass rc; uint64_t hi;
      Cell r = evalNode(nd, i, &rc, &hi);
      bool keep = (rc == VC_UUID) ? false : (cvt(r.v, rc, VC_BOOL) != 0);
      vals[k] = index[i];
      if (predicate) predicate[i] = keep;
      if (keep) keepMask |= 1u << k;
    }
  }
  const uint32_t mine = __popc(keepMask);
  uint32_t blockTotal;
  const uint32_t total = blockExclusiveScan<kFilterThreads>(mine, sWarp, &blockTotal);
  // all loads of this tile are complete (barrier inside the scan) before anything is published
  if (threadIdx.x < 32) {
    uint32_t p = decoupledLookback(st, tile, blockTotal);
    if (threadIdx.x == 0) {
      sPrefix = p;
      if ((tile + 1) * (uint64_t)kFilterTile >= (uint64_t)nd.n) *outCount = p + blockTotal;
    }
  }
  __syncthreads();
  uint32_t pos = sPrefix + total;
#pragma unroll
  for (int k = 0; k < kFilterItems; k++) {
    if (keepMask & (1u << k)) {
      index[pos] = vals[k];
      for (int f = 0; f < numForeign; f++) recOut[f][pos] = recIn[f][base + k];
      pos++;
    }
  }
}

// HashLookup (reference query/hash_lookup.cu:70-157): RecordID of the dimension-table row whose primary key equals the
// main-table join column at index position i; {0, 0} for NULL keys and keys the table does not hold.
__global__ void __launch_bounds__(256)
hashLookupKernel(InputDesc in, const uint32_t *__restrict__ index, const uint32_t *__restrict__ baseCounts, uint32_t startCount,
                 int n, CuckooDesc H, unsigned long long *__restrict__ out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += stride) {
    uint64_t hi = 0;
    const Cell c = loadInput(in, i, index, baseCounts, startCount, &hi);
    out[i] = c.valid ? cuckooLookup(H, c.v, hi) : 0ull;
  }
}

__global__ void initIndexKernel(uint32_t *index, uint32_t start, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) index[i] = start + (uint32_t)i;
}

// ---------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------
static int gridFor(int n, int perBlock) {
  int blocks = divUp(n, perBlock);
  int cap = smCount() * 16;
  return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

static NodeDesc makeNode(const InputVector *ins, int nin, int fn, uint32_t *index, int n,
                         uint32_t *baseCounts, uint32_t startCount) {
  NodeDesc nd;
  memset(&nd, 0, sizeof(nd));
  nd.nin = (uint8_t)nin; nd.fn = fn; nd.index = index; nd.n = n;
  nd.baseCounts = baseCounts; nd.startCount = startCount;
  for (int k = 0; k < nin; k++) {
    nd.in[k] = makeInputDesc(ins[k], /*allowWide=*/nin == 1);
    if (nd.in[k].kind == IN_FOREIGN) nd.fd[k] = makeForeignDesc(ins[k].Vector.ForeignVP);
  }
  if (nin == 2) nd.tclass = commonClass((ValClass)nd.in[0].vclass, (ValClass)nd.in[1].vclass);
  else nd.tclass = nd.in[0].vclass;
  for (int k = 0; k < nin; k++)
    if (nd.in[k].kind == IN_COLUMN && nd.in[k].mode != 0 && index == nullptr)
      throw EngineError("indexVector must not be NULL for a column input");
  return nd;
}

static int64_t runTransform(const InputVector *ins, int nin, const OutputVector &out, uint32_t *index,
                            int n, uint32_t *baseCounts, uint32_t startCount, int fn, cudaStream_t s) {
  if (n <= 0) return 0;
  NodeDesc nd = makeNode(ins, nin, fn, index, n, baseCounts, startCount);
  SinkDesc sk = makeSink(out);
  if (sk.kind == SINK_MEASURE && !sk.skipCount && baseCounts != nullptr && index == nullptr)
    throw EngineError("indexVector must not be NULL for a measure output over a compressed batch");
  transformKernel<<<gridFor(n, 256), 256, 0, s>>>(nd, sk);
  checkLastError("transform");
  return n;
}

static int64_t runFilter(const InputVector *ins, int nin, uint32_t *index, uint8_t *predicate, int n,
                         RecordID **recordIDVectors, int numForeign, uint32_t *baseCounts,
                         uint32_t startCount, int fn, cudaStream_t s) {
  if (n <= 0) return 0;
  if (numForeign < 0 || numForeign > 8) throw EngineError("only support up to 8 foreign tables");
  NodeDesc nd = makeNode(ins, nin, fn, index, n, baseCounts, startCount);
  const int tiles = divUp(n, kFilterTile);
  Scratch state(scanStateBytes(tiles) + sizeof(uint32_t), s);
  ARES_CUDA(cudaMemsetAsync(state.ptr, 0, state.bytes, s));
  ScanTileState st = makeScanState(state.ptr, tiles);
  uint32_t *dCount = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(state.ptr) + scanStateBytes(tiles));

  // RecordID vectors of joined tables are compacted through a scratch copy (rare path).
  Scratch recPtrs, recCopy;
  RecordID **dIn = nullptr, **dOut = nullptr;
  std::vector<RecordID *> hOut;
  if (numForeign > 0) {
    size_t each = (size_t)n * sizeof(RecordID);
    recCopy.reset(each * numForeign, s);
    recPtrs.reset(sizeof(RecordID *) * 2 * numForeign, s);
    std::vector<RecordID *> h(2 * numForeign);
    for (int f = 0; f < numForeign; f++) {
      h[f] = reinterpret_cast<RecordID *>(static_cast<uint8_t *>(recCopy.ptr) + each * f);      // in (copy)
      h[numForeign + f] = recordIDVectors[f];                                                   // out
      ARES_CUDA(cudaMemcpyAsync(h[f], recordIDVectors[f], each, cudaMemcpyDeviceToDevice, s));
    }
    ARES_CUDA(cudaMemcpyAsync(recPtrs.ptr, h.data(), h.size() * sizeof(RecordID *), cudaMemcpyHostToDevice, s));
    ARES_CUDA(cudaStreamSynchronize(s));  // h is a stack vector
    dIn = recPtrs.as<RecordID *>();
    dOut = dIn + numForeign;
  }
  filterKernel<<<tiles, kFilterThreads, 0, s>>>(nd, index, predicate, st, dCount, dIn, dOut, numForeign);
  checkLastError("filter");
  uint32_t hCount = 0;
  ARES_CUDA(cudaMemcpyAsync(&hCount, dCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  ARES_CUDA(cudaStreamSynchronize(s));
  return hCount;
}

}  // namespace aresb

using namespace aresb;

extern "C" {

CGoCallResHandle InitIndexVector(uint32_t *indexVector, uint32_t start, int indexVectorLength,
                                 void *cudaStream, int device) {
  return guarded("InitIndexVector", device, [&]() -> int64_t {
    if (indexVectorLength <= 0) return 0;
    initIndexKernel<<<divUp(indexVectorLength, 256), 256, 0, (cudaStream_t)cudaStream>>>(
        indexVector, start, indexVectorLength);
    checkLastError("InitIndexVector");
    return 0;
  });
}

CGoCallResHandle UnaryTransform(InputVector input, OutputVector output, uint32_t *indexVector,
                                int indexVectorLength, uint32_t *baseCounts, uint32_t startCount,
                                enum UnaryFunctorType functorType, void *cudaStream, int device) {
  return guarded("UnaryTransform", device, [&]() -> int64_t {
    return runTransform(&input, 1, output, indexVector, indexVectorLength, baseCounts, startCount,
                        (int)functorType, (cudaStream_t)cudaStream);
  });
}

CGoCallResHandle BinaryTransform(InputVector lhs, InputVector rhs, OutputVector output,
                                 uint32_t *indexVector, int indexVectorLength, uint32_t *baseCounts,
                                 uint32_t startCount, enum BinaryFunctorType functorType,
                                 void *cudaStream, int device) {
  return guarded("BinaryTransform", device, [&]() -> int64_t {
    InputVector ins[2] = {lhs, rhs};
    return runTransform(ins, 2, output, indexVector, indexVectorLength, baseCounts, startCount,
                        (int)functorType, (cudaStream_t)cudaStream);
  });
}

CGoCallResHandle UnaryFilter(InputVector input, uint32_t *indexVector, uint8_t *predicateVector,
                             int indexVectorLength, RecordID **recordIDVectors, int numForeignTables,
                             uint32_t *baseCounts, uint32_t startCount,
                             enum UnaryFunctorType functorType, void *cudaStream, int device) {
  return guarded("UnaryFilter", device, [&]() -> int64_t {
    return runFilter(&input, 1, indexVector, predicateVector, indexVectorLength, recordIDVectors,
                     numForeignTables, baseCounts, startCount, (int)functorType, (cudaStream_t)cudaStream);
  });
}

CGoCallResHandle BinaryFilter(InputVector lhs, InputVector rhs, uint32_t *indexVector,
                              uint8_t *predicateVector, int indexVectorLength,
                              RecordID **recordIDVectors, int numForeignTables, uint32_t *baseCounts,
                              uint32_t startCount, enum BinaryFunctorType functorType,
                              void *cudaStream, int device) {
  return guarded("BinaryFilter", device, [&]() -> int64_t {
    InputVector ins[2] = {lhs, rhs};
    return runFilter(ins, 2, indexVector, predicateVector, indexVectorLength, recordIDVectors,
                     numForeignTables, baseCounts, startCount, (int)functorType, (cudaStream_t)cudaStream);
  });
}

// ---- symbols outside the hot path (must exist, may return an error) -------
CGoCallResHandle HashLookup(InputVector input, RecordID *output, uint32_t *indexVector, int indexVectorLength,
                            uint32_t *baseCounts, uint32_t startCount, CuckooHashIndex hashIndex, void *cudaStream,
                            int device) {
  return guarded("HashLookup", device, [&]() -> int64_t {
    if (indexVectorLength <= 0) return 0;
    if (input.Type != VectorPartyInput) throw EngineError("HashLookup joins on a main-table column (VectorPartyInput)");
    InputDesc d = makeInputDesc(input, /*allowWide=*/true);
    if (d.mode != 0 && indexVector == nullptr) throw EngineError("indexVector must not be NULL for a column input");
    if (hashIndex.numBuckets <= 0 || hashIndex.keyBytes <= 0 || hashIndex.keyBytes > 16 || hashIndex.numHashes < 0 || hashIndex.numHashes > 4)
      throw EngineError("invalid CuckooHashIndex");
    CuckooDesc H;
    H.buckets = hashIndex.buckets;
    for (int i = 0; i < 4; i++) H.seeds[i] = hashIndex.seeds[i];
    H.keyBytes = hashIndex.keyBytes; H.numHashes = hashIndex.numHashes; H.numBuckets = hashIndex.numBuckets;
    hashLookupKernel<<<gridFor(indexVectorLength, 256), 256, 0, (cudaStream_t)cudaStream>>>(
        d, indexVector, baseCounts, startCount, indexVectorLength, H, reinterpret_cast<unsigned long long *>(output));
    checkLastError("HashLookup");
    return indexVectorLength;   // thrust::transform's end - begin (query/hash_lookup.cu:147-156)
  });
}
CGoCallResHandle Expand(DimensionVector, DimensionVector, uint32_t *, uint32_t *, int, int, void *, int) {
  return unsupported("Expand", "non-aggregate queries are out of scope");
}
CGoCallResHandle GeoBatchIntersects(GeoShapeBatch, InputVector, uint32_t *, int, uint32_t, RecordID **,
                                    int, uint32_t *, bool, void *, int) {
  return unsupported("GeoBatchIntersects", "geofence joins are out of scope");
}
CGoCallResHandle WriteGeoShapeDim(int, DimensionOutputVector, int, uint32_t *, void *, int) {
  return unsupported("WriteGeoShapeDim", "geofence joins are out of scope");
}

// Calendar tables are compile-time constants here, so there is nothing to upload; the call
// still validates that every visible device is usable (reference query/utils.cu:63-85).
CGoCallResHandle BootstrapDevice() {
  CGoCallResHandle h = {nullptr, nullptr};
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    h.pStrErr = strdup((std::string("BootstrapDevice: ") + cudaGetErrorString(e)).c_str());
    return h;
  }
  for (int d = 0; d < n; d++) {
    e = cudaSetDevice(d);
    if (e == cudaSuccess) e = cudaFree(0);
    if (e != cudaSuccess) {
      h.pStrErr = strdup((std::string("BootstrapDevice: ") + cudaGetErrorString(e)).c_str());
      return h;
    }
  }
  return h;
}

}  // extern "C"
