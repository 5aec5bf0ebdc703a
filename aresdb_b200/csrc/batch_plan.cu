// batch_plan.cu — additive whole-batch API (include/aresdb_b200/batch_plan.h).  Placeholder.
#include "common.cuh"
using namespace aresb;
extern "C" {
CGoCallResHandle AggStateCreate(AggSpec, void *, int) { return unsupported("AggStateCreate", "not implemented yet"); }
CGoCallResHandle ExecuteBatchPlan(void *, const BatchPlan *, void *, int) { return unsupported("ExecuteBatchPlan", "not implemented yet"); }
CGoCallResHandle AggStateMerge(void *, DimensionVector, uint8_t *, int, void *, int) { return unsupported("AggStateMerge", "not implemented yet"); }
CGoCallResHandle AggStateGroupCount(void *, void *, int) { return unsupported("AggStateGroupCount", "not implemented yet"); }
CGoCallResHandle AggStateFinalize(void *, DimensionVector, uint8_t *, void *, int) { return unsupported("AggStateFinalize", "not implemented yet"); }
CGoCallResHandle AggStateReset(void *, void *, int) { return unsupported("AggStateReset", "not implemented yet"); }
CGoCallResHandle AggStateDestroy(void *, int) { return unsupported("AggStateDestroy", "not implemented yet"); }
}
