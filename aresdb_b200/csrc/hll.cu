// hll.cu — HyperLogLog entry point (reference: query/hll.cu:21-290).  Placeholder until the
// kernels land: fails loudly instead of computing anything on the CPU.
#include "common.cuh"
using namespace aresb;
extern "C" CGoCallResHandle HyperLogLog(DimensionVector, DimensionVector, uint32_t *, uint32_t *, int, int, bool,
                                        uint8_t **, size_t *, uint16_t **, void *, int) {
  return unsupported("HyperLogLog", "not implemented yet");
}
