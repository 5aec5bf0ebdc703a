/*
 * batch_plan.h — ADDITIVE whole-batch API of the B200 engine (no counterpart symbol in the
 * reference; SURVEY.md §0.1-0.2 / §8b explain why it must exist).
 *
 * The reference executes one cgo call (= 1-2 Thrust launches, a 5 B/row scratch vector and a
 * stream sync for every filter) per AST node: query/time_series_aggregate.go:493-593.  Here
 * the Go batch executor (query/aql_batchexecutor.go:102-273: preExec + filter + project +
 * reduce) hands the WHOLE batch to one call.  The plan is the same post-order walk
 * processExpression() performs, flattened: every PlanInst is exactly one non-leaf AST node
 * with the meaning of the legacy Unary/Binary Transform/Filter call it replaces — leaf
 * operands (VarRef columns, literals) are referenced in place, sub-expression operands come
 * from an evaluation stack that lives in registers instead of scratch vectors, and the root
 * of each expression is routed to a filter / dimension / measure sink.
 *
 * Aggregation state lives across batches in an AggState (a device hash table keyed by the
 * packed dimension row), replacing the "carry previous results in the input buffers and
 * re-sort / re-insert them every batch" protocol (query/aql_processor.go:743-776,
 * query/time_series_aggregate.go:683-716).  AggStateFinalize emits exactly what the last
 * Reduce / HashReduce of the reference would have left in the output DimensionVector and
 * measure vector: groups identified by the murmur3 hash of the packed row (64-bit low word
 * for ARES_REDUCE_SORT, 32-bit for ARES_REDUCE_HASH), ascending-hash order for SORT.
 */
#ifndef ARESDB_B200_BATCH_PLAN_H_
#define ARESDB_B200_BATCH_PLAN_H_

#include "aql_abi.h"

enum {
  ARES_MAX_PLAN_COLUMNS = MAX_COLUMNS_OF_A_TABLE,
  ARES_MAX_PLAN_INSTS = 64,
  ARES_PLAN_STACK_DEPTH = 4,
  ARES_MAX_FOREIGN_TABLES = 4,
  ARES_MAX_FOREIGN_COLUMNS = 8
};

enum PlanOperandKind {
  PLAN_OPERAND_NONE = 0,
  PLAN_OPERAND_COLUMN = 1, /* BatchPlan.Columns[Column]: what makeVectorPartySliceInput passes   */
  PLAN_OPERAND_CONST = 2,  /* what makeConstantInput passes (ConstInt / ConstFloat + IsValid)   */
  PLAN_OPERAND_STACK = 3,  /* result of an earlier PLAN_SINK_STACK instruction (LIFO)            */
  PLAN_OPERAND_FOREIGN = 4 /* BatchPlan.ForeignColumns[Column]: what makeForeignColumnInput passes — a column of a
                            * joined dimension table, read at the RecordID the join finds for the row       */
};

typedef struct {
  uint8_t Kind;       /* enum PlanOperandKind */
  uint8_t Column;     /* PLAN_OPERAND_COLUMN: index into BatchPlan.Columns */
  uint8_t ConstType;  /* PLAN_OPERAND_CONST: ConstInt or ConstFloat */
  uint8_t ConstValid; /* PLAN_OPERAND_CONST: ConstantVector.IsValid */
  union {
    int32_t IntVal;
    float FloatVal;
  } Const;
} PlanOperand;

enum PlanSink {
  PLAN_SINK_STACK = 0,     /* non-root node: ScratchSpaceOutput of DataType SinkDataType       */
  PLAN_SINK_FILTER = 1,    /* root of a filter: row survives iff bool(value) (filterAction)      */
  PLAN_SINK_DIMENSION = 2, /* root of dimension #SinkArg (layout order), DimensionOutput         */
  PLAN_SINK_MEASURE = 3    /* root of the measure, MeasureOutput of SinkDataType / AggSpec.AggFunc */
};

typedef struct {
  uint8_t NumOperands;  /* 1: Functor is a UnaryFunctorType; 2: a BinaryFunctorType */
  uint8_t Functor;
  uint8_t Sink;         /* enum PlanSink */
  uint8_t SinkArg;      /* dimension ordinal for PLAN_SINK_DIMENSION */
  uint8_t SinkDataType; /* enum DataType of the sink element */
  uint8_t Reserved[3];
  PlanOperand A;
  PlanOperand B;
} PlanInst;

/* Optional zone-map entry of one column of one batch: every VALID (non-NULL) value of the column in this
 * batch lies in [Min, Max] as a non-negative integer below 2^31 (Bool / Uint8 / Uint16 / Uint32 columns and
 * non-negative signed ones; for a Float32 column whose valid values are all >= 0: the IEEE bit patterns of the
 * smallest and largest value, which order like the values — a float SUM then accumulates the rows that lie on
 * the 2^-S grid the maximum allows as exact integers).  The reference keeps exactly this for the time column of live batches
 * (LiveVectorParty.GetMinMaxValue, memstore/common/vector_party.go:184-186, used for batch skipping in
 * query/aql_processor.go:1509) and knows it by construction for archive batches (batch ID = day) and for
 * enum columns (dictionary size).  It is a HINT: when every dimension of the query has a small known range
 * the fused kernel addresses its CTA-private accumulators directly by (dimension value - Min) instead of
 * probing a hash table; every row is checked against the range and rows outside it take the hash path, so a
 * stale or wrong entry costs speed, never correctness.  Known = 0: no information. */
typedef struct {
  uint8_t Known;
  uint8_t Reserved[3];
  uint32_t Min;
  uint32_t Max;
} ColumnRange;

/* Dimension-table joins on the fused path (what BatchExecutorImpl.join() prepares per batch with HashLookup,
 * query/aql_batchexecutor.go:115-147, and what makeForeignColumnInput passes per expression leaf).  The lookup is a gather
 * stage INSIDE the fused kernel: the RecordID of a surviving row is found by probing the dimension table's cuckoo index
 * with the row's join-column value when the first instruction that reads the table is reached, and the foreign column is
 * read at it; no RecordID vector is materialised.  Join keys are 1- / 2- / 4-byte main-table columns. */
typedef struct {
  int32_t JoinColumn;    /* index into BatchPlan.Columns: the main-table column matched with the table's primary key */
  CuckooHashIndex Index; /* the dimension table's primary-key index (device memory)                                  */
} PlanForeignTable;

typedef struct {
  int32_t Table;              /* index into BatchPlan.ForeignTables                                                  */
  ForeignColumnVector Column; /* as makeForeignColumnInput fills it; RecordIDs is ignored; Batches points to HOST memory
                               * valid for the duration of the call (at most 8 batches); TimezoneLookup to device memory */
} PlanForeignColumn;

/* One batch of one table shard: column slices already resident on the device. */
typedef struct {
  VectorPartySlice Columns[ARES_MAX_PLAN_COLUMNS];
  int32_t NumColumns; /* at most 16 per call: list the columns the query reads (what transferBatch copies), not the table */
  PlanInst Insts[ARES_MAX_PLAN_INSTS];
  int32_t NumInsts;
  /* Index space of the batch = rows of the first column (as in the reference).  BaseCounts
   * is that column's cumulative count vector for RLE (archive, sorted) batches, or NULL;
   * StartCount is the row number of index 0 when BaseCounts is NULL
   * (oopkBatchContext.baseCountD / startRow, query/aql_context.go:151-235). */
  uint32_t *BaseCounts;
  uint32_t StartCount;
  uint32_t NumRows;
  ColumnRange Ranges[ARES_MAX_PLAN_COLUMNS]; /* zone map per entry of Columns (all zero: none) */
  PlanForeignTable ForeignTables[ARES_MAX_FOREIGN_TABLES];
  int32_t NumForeignTables;
  PlanForeignColumn ForeignColumns[ARES_MAX_FOREIGN_COLUMNS];
  int32_t NumForeignColumns;
} BatchPlan;

enum AresReduceMode {
  ARES_REDUCE_SORT = 0, /* semantics of Sort + Reduce (64-bit hash identity, hash-ascending output) */
  ARES_REDUCE_HASH = 1  /* semantics of HashReduce (32-bit hash identity, unordered output)        */
};

typedef struct {
  uint8_t NumDimsPerDimWidth[NUM_DIM_WIDTH]; /* as DimensionVector */
  uint8_t Reserved[3];
  int32_t AggFunc;         /* enum AggregateFunction: SUM/MIN/MAX families, AGGR_AVG_FLOAT (8-byte (float average, count)
                            * pairs combined with the reference's rolling average), or AGGR_HLL (measure = Uint32
                            * rho << 16 | reg values; group identity and results as HyperLogLog's,
                            * query/hll.cu:21-290; read the result with AggStateFinalizeHLL) */
  int32_t MeasureDataType; /* enum DataType of one measure element: Int32/Uint32/Float32/Int64/Float64 */
  int32_t ReduceMode;      /* enum AresReduceMode */
  uint32_t ExpectedGroups; /* capacity hint: the group table holds max(2^21, 2 x ExpectedGroups) slots; exceeding it
                            * is reported as an error by AggStateGroupCount / AggStateFinalize, never silently */
} AggSpec;

#ifdef __cplusplus
extern "C" {
#endif

/* res = opaque state handle.  Allocates the group table on `device`. */
CGoCallResHandle AggStateCreate(AggSpec spec, void *cudaStream, int device);

/* Fused preExec+filter+project+reduce of one batch into `state`.  Asynchronous on
 * cudaStream: nothing is returned to the host, nothing is synchronised (res = 0).  A state is used from
 * one stream at a time (the batches of a query follow each other, as in the reference); different states
 * run concurrently on different streams / devices. */
CGoCallResHandle ExecuteBatchPlan(void *state, const BatchPlan *plan, void *cudaStream, int device);

/* Folds already-reduced rows (a DimensionVector block + measure vector, e.g. the carried
 * result of the legacy protocol, or the all-gathered results of other GPUs) into `state`
 * with the aggregate's combine rule (broker/result_merge.go:80-105 semantics). */
CGoCallResHandle AggStateMerge(void *state, DimensionVector inputKeys, uint8_t *inputValues,
                               int length, void *cudaStream, int device);

/* res = number of occupied group slots (>= number of output groups); synchronises. */
CGoCallResHandle AggStateGroupCount(void *state, void *cudaStream, int device);

/* Writes the groups into outputKeys (capacity outputKeys.VectorCapacity; DimValues required,
 * HashValues / IndexVector filled when non-NULL) and outputValues; res = number of groups.
 * Synchronises cudaStream.  The state stays valid (it can be finalized again or reset). */
CGoCallResHandle AggStateFinalize(void *state, DimensionVector outputKeys, uint8_t *outputValues,
                                  void *cudaStream, int device);

/* Exchange form of AggStateFinalize: the occupied table slots as (dim row, partial measure) pairs in
 * no particular order and without merging equal hashes — what another state's AggStateMerge consumes.
 * Cheaper than AggStateFinalize (no sort); res = number of rows (== AggStateGroupCount). Synchronises. */
CGoCallResHandle AggStateExport(void *state, DimensionVector outputKeys, uint8_t *outputValues,
                                void *cudaStream, int device);

/* The exchange step of a sharded query without host involvement.  AggStateExportPart writes this state's rows as ONE
 * fixed-capacity part — [uint32 rows, uint32 status, uint32 claimed, pad | DimensionVector block of capRows rows at
 * dimOffset | measures at valuesOffset] — with one launch and no synchronisation (the row count stays on the device);
 * the parts of all ranks are all-gathered; AggStateMergeParts folds every gathered part into the receiving state with one
 * launch, reading the counts from the part headers.  capRows <= 32768.  A state with more rows marks its part
 * (status != 0, no rows) and the receiver's next AggStateFinalize fails with "exchange part truncated": repeat the step
 * with AggStateGroupCount / AggStateExport / AggStateMerge (exact sizes).  Not for AGGR_HLL states. */
CGoCallResHandle AggStateExportPart(void *state, uint8_t *part, int capRows, size_t dimOffset, size_t valuesOffset,
                                    void *cudaStream, int device);
CGoCallResHandle AggStateMergeParts(void *state, const uint8_t *parts, int numParts, size_t partStride, int capRows,
                                    size_t dimOffset, size_t valuesOffset, void *cudaStream, int device);

/* The same exchange over PEER MEMORY (one NVLink / NVSwitch node), without a collective library: the host maps every
 * rank's receive buffer into every process (CUDA IPC / fabric handles — torch symmetric memory does it) and hands over
 * peerSlots[r] = the address of THIS rank's part slot inside rank r's receive buffer and peerFlags[r] = the address of
 * flags[myRank] on rank r.  AggStateExportPartToPeers is ONE launch: it writes the part into the local slot, copies it
 * into every peer's slot with 16-byte stores over NVLink and then stores `epoch` into the flag on every peer (release,
 * system scope).  AggStateMergePartsWhenFlagged is AggStateMergeParts whose kernel first waits (bounded: ~2 s, then the
 * next AggStateFinalize fails) until all numParts flags have reached `epoch`.  Use two receive buffers alternately and a
 * growing epoch: a rank can be at most one exchange ahead of its peers.  numPeers <= 16, partBytes % 16 == 0. */
CGoCallResHandle AggStateExportPartToPeers(void *state, uint8_t *const *peerSlots, uint32_t *const *peerFlags, int numPeers, int myRank,
                                           size_t partBytes, int capRows, size_t dimOffset, size_t valuesOffset, uint32_t epoch,
                                           void *cudaStream, int device);
CGoCallResHandle AggStateMergePartsWhenFlagged(void *state, const uint8_t *parts, int numParts, size_t partStride, int capRows,
                                               size_t dimOffset, size_t valuesOffset, const uint32_t *flags, uint32_t epoch,
                                               void *cudaStream, int device);

/* AGGR_HLL states: the final outputs of the reference's last-batch HyperLogLog call
 * (query/hll.cu:262-290, adopted by query/time_series_aggregate.go:661-681).  res = number of dimension
 * groups g.  *dimValuesPtr = a DimensionVector block of VectorCapacity g (groups in key order),
 * *hllDimRegIDCountPtr = g register counts, *hllVectorPtr / *hllVectorSizePtr = per group either
 * count x 4 bytes ((rho+1) << 16 | reg, count < 4096) or 16384 dense bytes.  All three are allocated
 * with deviceMalloc; the caller frees them with DeviceFree.  On such a state AggStateGroupCount
 * counts (group, register) entries and AggStateFinalize / AggStateMerge exchange the carried form
 * (one row per entry, Uint32 value) — which is how several GPUs combine HLL states.  Synchronises. */
CGoCallResHandle AggStateFinalizeHLL(void *state, uint8_t **dimValuesPtr, uint8_t **hllVectorPtr,
                                     size_t *hllVectorSizePtr, uint16_t **hllDimRegIDCountPtr,
                                     void *cudaStream, int device);

/* Zone-map production: out[i] = min / max of the VALID values of columns[i] (at most 16 per call; Bool / 1- / 2- /
 * 4-byte integer and Float32 columns in modes 0-3; everything else, columns without a valid value, and columns whose
 * values the ColumnRange contract cannot describe — negative, >= 2^31, negative or non-finite floats — get Known = 0).
 * One kernel over all columns, called once when a batch becomes device resident; the result is what BatchPlan.Ranges
 * takes.  The reference keeps this pair only for live Uint32 vector parties (memstore/live_vector_party.go:74-75).
 * res = number of columns scanned.  Synchronises cudaStream. */
CGoCallResHandle ComputeColumnRanges(const VectorPartySlice *columns, int numColumns, ColumnRange *out,
                                     void *cudaStream, int device);

/* Empties the table, keeping its memory. */
CGoCallResHandle AggStateReset(void *state, void *cudaStream, int device);

CGoCallResHandle AggStateDestroy(void *state, int device);

/* Diagnostics: engine kernels launched by this process so far (bench.py reports the delta over the
 * timed region as `gpu_launches`). */
unsigned long long AresKernelLaunchCount();

#ifdef __cplusplus
}
#endif

#endif /* ARESDB_B200_BATCH_PLAN_H_ */
