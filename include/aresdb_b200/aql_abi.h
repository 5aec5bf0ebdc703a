/*
 * aql_abi.h — C ABI of the B200-native AQL batch-execution engine (libalgorithm.so).
 *
 * This is the drop-in boundary for the hot path that AresDB's Go query processor drives
 * through cgo.  Every type and entry point below is byte- and name-compatible with what
 * the reference's cgo preamble binds (reference: query/time_series_aggregate.h, cited
 * per item as "ref: <line range>"), so `query/time_series_aggregate.go:17`
 * (`#cgo LDFLAGS: -lalgorithm`) links against this library unchanged.  The additive
 * whole-batch API that replaces the per-AST-node call sequence lives in batch_plan.h.
 *
 * Layout facts relied upon (verified with ctypes against the reference build,
 * tests/test_abi_vs_reference_headers.py): natural x86-64 alignment, enums are 4-byte ints,
 * sizeof(DefaultValue)=24, VectorPartySlice=56, ScratchSpaceVector=16, ConstantVector=24,
 * ForeignColumnVector=72, ArrayVectorPartySlice=24, InputVector=80, OutputVector=32,
 * DimensionVector=40.  All structs are passed BY VALUE.
 */
#ifndef ARESDB_B200_AQL_ABI_H_
#define ARESDB_B200_AQL_ABI_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "cgo_result.h"

/* ---- limits (ref: time_series_aggregate.h:33-47) ------------------------------------ */
enum {
  MAX_FOREIGN_TABLES = 7,
  MAX_COLUMNS_OF_A_TABLE = 32,
  MAX_DIMENSIONS = 8,
  MAX_DIMENSION_BYTES = 32,
  MAX_MEASURES = 32,
  MAX_INSTRUCTIONS = 1024,
  HASH_BUCKET_SIZE = 8,
  HASH_STASH_SIZE = 4,
  HLL_BITS = 14,
  HLL_DENSE_SIZE = 1 << HLL_BITS,
  HLL_DENSE_THRESHOLD = HLL_DENSE_SIZE / 4,
  NUM_DIM_WIDTH = 5 /* dimension widths 16, 8, 4, 2, 1 bytes, in this order */
};

/* ---- enums: numeric values are ABI (ref: :50-62, :65-79, :82-87, :90-107, :110-130) -- */
enum AggregateFunction {
  AGGR_SUM_UNSIGNED = 1, AGGR_SUM_SIGNED = 2, AGGR_SUM_FLOAT = 3,
  AGGR_MIN_UNSIGNED = 4, AGGR_MIN_SIGNED = 5, AGGR_MIN_FLOAT = 6,
  AGGR_MAX_UNSIGNED = 7, AGGR_MAX_SIGNED = 8, AGGR_MAX_FLOAT = 9,
  AGGR_HLL = 10, AGGR_AVG_FLOAT = 11
};

enum DataType {
  Bool, Int8, Uint8, Int16, Uint16, Int32, Uint32, Float32, Int64, Uint64, Float64,
  GeoPoint, UUID
};

enum ConstDataType { ConstInt, ConstFloat, ConstGeoPoint, ConstUUID };

enum UnaryFunctorType {
  Negate, Not, BitwiseNot, IsNull, IsNotNull, Noop,
  GetWeekStart, GetMonthStart, GetQuarterStart, GetYearStart,
  GetDayOfMonth, GetDayOfYear, GetMonthOfYear, GetQuarterOfYear,
  GetHLLValue, ArrayLength
};

enum BinaryFunctorType {
  And, Or, Equal, NotEqual, LessThan, LessThanOrEqual, GreaterThan, GreaterThanOrEqual,
  Plus, Minus, Multiply, Divide, Mod, BitwiseAnd, BitwiseOr, BitwiseXor, Floor,
  ArrayContains, ArrayElementAt
};

/* ---- small value types (ref: :133-174) ---------------------------------------------- */
typedef struct { int32_t batchID; uint32_t index; } RecordID;

typedef struct {
  uint8_t *buckets;
  uint32_t seeds[4];
  int keyBytes;
  int numHashes;
  int numBuckets;
} CuckooHashIndex;

typedef struct { float Lat; float Long; } GeoPointT;
typedef struct { uint64_t p1; uint64_t p2; } UUIDT;

typedef struct {
  bool HasDefault;
  union {
    bool BoolVal;
    int32_t Int32Val;
    uint32_t Uint32Val;
    float FloatVal;
    int64_t Int64Val;
    GeoPointT GeoPointVal;
    UUIDT UUIDVal;
  } Value;
} DefaultValue;

/*
 * VectorPartySlice (ref: :177-198; built by makeVectorPartySlice,
 * query/time_series_aggregate.go:166-206).  One device allocation holds
 * [counts u32 x (Length+1)] [null bitmap] [values]; BasePtr addresses the first part that
 * exists and the two offsets are relative to it:
 *   mode 0  BasePtr == NULL                      -> constant DefaultValue
 *   mode 1  ValuesOffset == 0                    -> values only, all valid
 *   mode 2  ValuesOffset != 0, NullsOffset == 0  -> bitmap + values
 *   mode 3  both != 0                            -> RLE counts + bitmap + values
 * StartingIndex (0..7) is the bit offset of row 0 in the bitmap (and in bit-packed Bool
 * values).
 */
typedef struct {
  uint8_t *BasePtr;
  uint32_t NullsOffset;
  uint32_t ValuesOffset;
  uint8_t StartingIndex;
  enum DataType DataType;
  DefaultValue DefaultValue;
  uint32_t Length;
} VectorPartySlice;

/* Intermediate of the per-node path: T Values[n] then bool valid[n] at NullsOffset (ref: :202-206). */
typedef struct {
  uint8_t *Values;
  uint32_t NullsOffset;
  enum DataType DataType;
} ScratchSpaceVector;

typedef struct { /* ref: :209-221 */
  union {
    int32_t IntVal;
    float FloatVal;
    GeoPointT GeoPointVal;
    UUIDT UUIDVal;
  } Value;
  bool IsValid;
  enum ConstDataType DataType;
} ConstantVector;

typedef struct { /* ref: :227-237 — dimension-table join input (out of scope, kept for layout) */
  RecordID *RecordIDs;
  VectorPartySlice *Batches;
  int32_t BaseBatchID;
  int32_t NumBatches;
  int32_t NumRecordsInLastBatch;
  int16_t *const TimezoneLookup;
  int16_t TimezoneLookupSize;
  enum DataType DataType;
  DefaultValue DefaultValue;
} ForeignColumnVector;

typedef struct { /* ref: :240-247 — array columns (out of scope, kept for layout) */
  uint8_t *OffsetLengthVector;
  uint32_t ValueOffsetAdj;
  enum DataType DataType;
  uint32_t Length;
} ArrayVectorPartySlice;

enum InputVectorType {
  VectorPartyInput, ScratchSpaceInput, ConstantInput, ForeignColumnInput, ArrayVectorPartyInput
};

typedef struct { /* ref: :260-269 */
  union {
    ConstantVector Constant;
    VectorPartySlice VP;
    ScratchSpaceVector ScratchSpace;
    ForeignColumnVector ForeignVP;
    ArrayVectorPartySlice ArrayVP;
  } Vector;
  enum InputVectorType Type;
} InputVector;

/*
 * DimensionVector (ref: :277-283; offsets query/common/dimval.go:122-145).  DimValues is a
 * column-major block with row capacity C = VectorCapacity: for each width in
 * {16,8,4,2,1} the value columns of the dims of that width (width*C bytes each), then one
 * validity byte-column (C bytes) per dim in the same dim order.
 */
typedef struct {
  uint8_t *DimValues;
  uint64_t *HashValues;
  uint32_t *IndexVector;
  int VectorCapacity;
  uint8_t NumDimsPerDimWidth[NUM_DIM_WIDTH];
} DimensionVector;

typedef struct { /* ref: :287-291 */
  uint8_t *DimValues;
  uint8_t *DimNulls;
  enum DataType DataType;
} DimensionOutputVector;

typedef struct { /* ref: :296-302 */
  uint32_t *Values;
  enum DataType DataType;
  enum AggregateFunction AggFunc;
} MeasureOutputVector;

enum OutputVectorType { ScratchSpaceOutput, MeasureOutput, DimensionOutput };

typedef struct { /* ref: :312-319 */
  union {
    ScratchSpaceVector ScratchSpace;
    DimensionOutputVector Dimension;
    MeasureOutputVector Measure;
  } Vector;
  enum OutputVectorType Type;
} OutputVector;

typedef struct { /* ref: :402-413 (geofence joins: out of scope) */
  float *Lats;
  float *Longs;
  uint16_t NumPoints;
} GeoShape;

typedef struct { /* ref: :416-425 */
  uint8_t *LatLongs;
  int32_t TotalNumPoints;
  uint8_t TotalWords;
} GeoShapeBatch;

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Error contract (ref: cgoutils/utils.h:20-23, cgoutils/utils.go:25-33): every entry point
 * returns {res, pStrErr}.  On success pStrErr is NULL and res carries an integer cast to a
 * pointer (new length / group count) where documented.  On failure pStrErr is a malloc'd
 * C string that the CALLER frees; nothing is thrown across the boundary.  Every entry
 * point selects `device` itself (cgo calls hop OS threads), enqueues all work on
 * `cudaStream`, and synchronises that stream only when it must return a count.
 */

/* index[i] = start + i                                         (ref: :438-442, algorithm.cu:22-41) */
CGoCallResHandle InitIndexVector(uint32_t *indexVector, uint32_t start, int indexVectorLength,
                                 void *cudaStream, int device);

/* Dimension-table join probe — out of scope: always returns an error string (ref: :446-454). */
CGoCallResHandle HashLookup(InputVector input, RecordID *output, uint32_t *indexVector,
                            int indexVectorLength, uint32_t *baseCounts, uint32_t startCount,
                            CuckooHashIndex hashIndex, void *cudaStream, int device);

/* out[i] = f(in[index[i]]) into a scratch / dimension / measure sink (ref: :461-471, transform.cu:21-53) */
CGoCallResHandle UnaryTransform(InputVector input, OutputVector output, uint32_t *indexVector,
                                int indexVectorLength, uint32_t *baseCounts, uint32_t startCount,
                                enum UnaryFunctorType functorType, void *cudaStream, int device);

/* Stable in-place compaction of index (and RecordID vectors) by f(in); res = new length (ref: :478-490, filter.cu:130-166) */
CGoCallResHandle UnaryFilter(InputVector input, uint32_t *indexVector, uint8_t *predicateVector,
                             int indexVectorLength, RecordID **recordIDVectors,
                             int numForeignTables, uint32_t *baseCounts, uint32_t startCount,
                             enum UnaryFunctorType functorType, void *cudaStream, int device);

/* (ref: :495-506, transform.cu:55-86) */
CGoCallResHandle BinaryTransform(InputVector lhs, InputVector rhs, OutputVector output,
                                 uint32_t *indexVector, int indexVectorLength,
                                 uint32_t *baseCounts, uint32_t startCount,
                                 enum BinaryFunctorType functorType, void *cudaStream, int device);

/* (ref: :509-522, filter.cu:168-204) */
CGoCallResHandle BinaryFilter(InputVector lhs, InputVector rhs, uint32_t *indexVector,
                              uint8_t *predicateVector, int indexVectorLength,
                              RecordID **recordIDVectors, int numForeignTables,
                              uint32_t *baseCounts, uint32_t startCount,
                              enum BinaryFunctorType functorType, void *cudaStream, int device);

/*
 * HashValues[i] = murmur3_x64_128(packed dim row IndexVector[i], seed 0).lo, then a STABLE
 * ascending sort of (HashValues, IndexVector)             (ref: :528-531, sort_reduce.cu:118-133)
 */
CGoCallResHandle Sort(DimensionVector keys, int length, void *cudaStream, int device);

/*
 * Segmented reduce over runs of equal HashValues; first row of a run supplies the dims;
 * res = number of groups                                   (ref: :537-545, sort_reduce.cu:135-249)
 */
CGoCallResHandle Reduce(DimensionVector inputKeys, uint8_t *inputValues, DimensionVector outputKeys,
                        uint8_t *outputValues, int valueBytes, int length,
                        enum AggregateFunction aggFunc, void *cudaStream, int device);

/*
 * Group by murmur3_32(packed dim row) without sorting; output order unspecified;
 * res = number of groups                                   (ref: :552-560, hash_reduction.cu:346-391)
 */
CGoCallResHandle HashReduce(DimensionVector inputKeys, uint8_t *inputValues,
                            DimensionVector outputKeys, uint8_t *outputValues, int valueBytes,
                            int length, enum AggregateFunction aggFunc, void *cudaStream, int device);

/* RLE expansion for non-aggregate queries — out of scope: returns an error string (ref: :572-579). */
CGoCallResHandle Expand(DimensionVector inputKeys, DimensionVector outputKeys, uint32_t *baseCounts,
                        uint32_t *indexVector, int indexVectorLen, int outputOccupiedLen,
                        void *cudaStream, int device);

/* One batch of a HyperLogLog distinct-count aggregation (ref: :590-601, hll.cu:262-290). */
CGoCallResHandle HyperLogLog(DimensionVector prevDimOut, DimensionVector curDimOut,
                             uint32_t *prevValuesOut, uint32_t *curValuesOut, int prevResultSize,
                             int curBatchSize, bool isLastBatch, uint8_t **hllVectorPtr,
                             size_t *hllVectorSizePtr, uint16_t **hllDimRegIDCountPtr,
                             void *cudaStream, int device);

/* Geofence joins — out of scope: return an error string (ref: :608-618). */
CGoCallResHandle GeoBatchIntersects(GeoShapeBatch geoShapeBatch, InputVector points,
                                    uint32_t *indexVector, int indexVectorLength,
                                    uint32_t startCount, RecordID **recordIDVectors,
                                    int numForeignTables, uint32_t *outputPredicate, bool inOrOut,
                                    void *cudaStream, int device);
CGoCallResHandle WriteGeoShapeDim(int shapeTotalWords, DimensionOutputVector dimOut,
                                  int indexVectorLengthBeforeGeo, uint32_t *outputPredicate,
                                  void *cudaStream, int device);

/* Uploads the calendar tables to every device (ref: :621, utils.cu:63-85). */
CGoCallResHandle BootstrapDevice();

#ifdef __cplusplus
}
#endif

#endif /* ARESDB_B200_AQL_ABI_H_ */
