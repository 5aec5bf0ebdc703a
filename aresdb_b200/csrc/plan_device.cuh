// plan_device.cuh — the device form of a BatchPlan: what compilePlan() (batch_plan.cu) produces and
// both executors of the staged path consume (the interpreter kernel and the NVRTC generator).
#pragma once
#include "column.cuh"
#include "fused_device.cuh"
#include "join.cuh"

namespace aresb {

constexpr int kFusedThreads = 512;
constexpr int kMaxPlanCols = 16;
constexpr int kSmemBudget = 232448;           // 227 KB: the most dynamic shared memory a CTA may opt into on sm_100
constexpr int kJitThreads = 1024;
constexpr int kMaxGridCtas = 160;             // persistent grid: one CTA per SM (148 on B200)

enum KeyMode : uint8_t { KEY_PACKED = 0, KEY_HASHED = 1 };
enum OperandKind : uint8_t { OPK_NONE = 0, OPK_COLUMN = 1, OPK_CONST = 2, OPK_STACK = 3, OPK_FOREIGN = 4 };

// Joined dimension tables of a plan, in device memory for the duration of the batch's kernel (uploaded by executePlan).
constexpr int kMaxForeignTables = ARES_MAX_FOREIGN_TABLES;
constexpr int kMaxForeignCols = ARES_MAX_FOREIGN_COLUMNS;
struct DevJoin {
  CuckooDesc tables[kMaxForeignTables];
  ForeignDesc cols[kMaxForeignCols];
};

struct DevColumn {
  InputDesc in;            // how to read it straight from global memory (any mode)
  uint32_t smemValues;     // byte offsets inside one stage (staged path)
  uint32_t smemNulls;
  uint32_t tileValueBytes; // bytes of one full tile
  uint32_t tileNullBytes;
  uint8_t width;           // bytes per value, 0 for bit-packed bool
  uint8_t staged;          // values staged
  uint8_t hasNulls;        // mode 2 bitmap staged
  uint8_t used;            // some instruction reads the column (unreferenced columns of the batch cost nothing)
  uint32_t rangeLo, rangeHi;  // zone map of the batch (BatchPlan.Ranges): valid values lie in [rangeLo, rangeHi]
  uint8_t rangeKnown;
  uint8_t rle;             // run-length encoded (mode 3) column decoded in the kernel from its runs — never expanded, never staged
  uint8_t pad[2];
  const uint32_t *tileRun; // rle: run that holds the first index position of every tile (+ one entry for the last position)
};

struct DevInst {
  uint32_t aconst, bconst;
  uint8_t nops, fn, sink, sinkArg;
  uint8_t akind, acol, aclass, avalid;
  uint8_t bkind, bcol, bclass, bvalid;
  uint8_t tclass;   // class the functor runs in
  uint8_t rclass;   // class of the functor result
  uint8_t oclass;   // class of the sink element
  uint8_t wide;     // 1: 8/16-byte column copied verbatim into a dimension
  uint8_t rowOff, width, nullOff, pad;
};

struct DevPlan {
  DevColumn cols[kMaxPlanCols];
  DevInst insts[ARES_MAX_PLAN_INSTS];
  const uint32_t *baseCounts;
  unsigned long long *ctaAcc;   // [grid][smemSlots] accumulator slices (global, L2-resident)
  uint32_t startCount;
  uint32_t numRows;
  uint32_t tileRows;
  uint32_t numFullTiles;   // staged tiles; the tail goes through the direct path
  uint32_t stageBytes;
  uint32_t smemSlots;      // shared table slots (power of two)
  uint32_t tailBegin;      // first row of the direct (non-staged) pass
  uint32_t numStages;      // depth of the TMA ring (2..kMaxStages)
  uint32_t denseSlots;     // dense HLL mode: slots of the group directory (0 otherwise)
  uint32_t smemBc;         // staged base counts (RLE batches, SUM / AVG): byte offset inside a stage, tileBcBytes bytes
  uint32_t tileBcBytes;    // (tileRows + 4) * 4, or 0 when the base counts are not staged
  int32_t ncols, ninsts, lastFilter;
  uint64_t measureIdentity;  // NULL measure -> this (sink class bits)
  uint64_t accNeutral;       // neutral element of the combine op
  uint8_t keyMode, rowBytes, valueBytes, hashBits;
  uint8_t aggOp, measWidth, measClass, skipCount;
  uint8_t hasMeasure, staged, hll, bypassOk;
  // direct-indexed aggregation (jitAnalyzeDense; denseNd == 0: hash table).  Dimension k is produced by instruction
  // denseInst[k]; its slot index is (quotient or value) - denseLo[k], below denseCnt[k]; index denseCnt[k] = NULL.
  uint8_t denseNd;
  uint8_t denseInst[8], denseViaQuot[8], denseNullCanon[8];
  uint32_t denseLo[8], denseCnt[8], denseStep[8];
  // quotient dimensions whose dividend range is small (span * step <= 2^32): index = umulhi(x - denseBase, 2^32 / step + 1),
  // in range iff x - denseBase < denseSpan (one subtraction, one multiply, one compare per row)
  uint8_t denseSpan32[8];
  uint32_t denseBase[8], denseSpan[8];
  uint32_t denseTotal;     // slots of one copy = prod (denseCnt[k] + 1)
  uint32_t tableBytes;     // shared memory between the header and the first stage (keys, or flags + accumulators)
  // more slots than a CTA holds: ONE array of accumulators in global memory (L2-resident) shared by all CTAs, folded
  // into the group table by denseFoldKernel after the batch; a slot was reached iff it differs from accNeutral
  unsigned long long *denseAcc;
  uint8_t denseGlobal;
  uint8_t denseGlobalReps;   // copies of the global slot array (power of two; CTA b uses copy b mod reps): spreads the L2 atomics
  uint8_t neutralSafe;     // no sequence of row values can bring a reached accumulator back to accNeutral (set by compilePlan)
  uint8_t denseFx;         // float sum accumulated as exact integers (three 32-bit pieces per slot), see jitAnalyzeDense
  int8_t fxMeasureInst;    // the measure instruction (a verbatim Float32 column with a zone map)
  int8_t fxShift;          // S: a row adds x * 2^S
  uint8_t numForeignTables, numForeignCols;
  uint8_t joinCol[kMaxForeignTables];       // main-table column matched with table t's primary key
  uint8_t foreignTableOf[kMaxForeignCols];  // table of foreign column k (operand kind OPK_FOREIGN, acol / bcol = k)
  uint8_t foreignClass[kMaxForeignCols];    // ValClass a read of foreign column k yields
  uint8_t pad2[1];
  const DevJoin *join;     // device copy of the tables' indexes and the foreign columns' batches
  uint8_t compact;         // survivors of the filters are compacted per tile (warp ballot + prefix sum into a CTA-wide index
                           // list) before dimensions / measure are evaluated — chosen when that work dominates (HLL)
  uint8_t partition;       // radix-partitioned aggregation (tables beyond L2): the kernel emits (key, measure) entries sorted by
                           // table partition, a second kernel folds them partition by partition
  uint8_t partShift;       // partition of a key = its home slot >> partShift (64 partitions)
  uint4 *partBuf;          // [numRows] entries of this batch
  uint32_t *partDir;       // [numFullTiles][kPartDirWords]: per tile, offsets of the 64 partition segments + span base
  uint32_t *partCursor;    // entries appended so far
  uint32_t resume;         // 1: relaunch of the same batch after the group table grew (DevTable::progress holds the resume points)
};

constexpr uint32_t kDenseMaxSlots = 8192;   // = slots of a CTA's accumulator slice in AggState::ctaAcc
constexpr uint32_t kFxMaxRowsPerCta = 1u << 21;   // each 32-bit piece accumulator takes 2^21 adds of an 11-bit piece
constexpr uint32_t kGlobalDenseMaxSlots = 1u << 21;   // 16 MB of accumulators per state, allocated on first use

// jit.cu: runs the staged tiles of `P` with a kernel specialised for the plan's shape.  Returns false
// when NVRTC is unavailable or disabled (ARESDB_B200_JIT=0) so that the caller falls back to the
// interpreter kernel; throws EngineError when code generation / compilation fails.
bool jitAvailable();
bool planCompactable(const DevPlan &P);   // filters first, no RLE / base counts / joins: the compacted-index form applies
constexpr uint32_t kPartitions = 64;
constexpr uint32_t kPartDirWords = kPartitions + 2;   // off[0..64] (off[64] = entries of the tile), span base
constexpr uint32_t kPartitionExtraBytes = 2048;   // histogram / offsets / fill cursors behind the 64 KB tile buffer
constexpr size_t kPartitionMinSlots = (size_t)1 << 23;   // tables from 8M slots (>= 128 MB of keys + accumulators: beyond L2)
constexpr uint32_t kCompactListBytes = 16384;  // two lists (double-buffered per tile) of the survivors' tile row numbers (u16)
void jitAnalyzeDense(DevPlan &P, bool bypass);
size_t jitCompileOnly(const DevPlan &P, std::string *sourceOut);
bool jitLaunchStaged(const DevPlan &P, const DevTable &G, size_t smemBytes, int grid, cudaStream_t s);

}  // namespace aresb
