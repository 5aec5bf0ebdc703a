// common.cuh — error plumbing and launch helpers shared by every translation unit of
// libalgorithm.so (B200 / sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <exception>
#include <stdexcept>
#include <string>

#include "aresdb_b200/aql_abi.h"
#include "aresdb_b200/batch_plan.h"
#include "aresdb_b200/device_memory.h"

namespace aresb {

// Thrown inside an entry point, turned into CGoCallResHandle::pStrErr at the boundary
// (contract: reference query/filter.cu:142-164 — try/catch -> strdup(e.what())).
struct EngineError : public std::runtime_error {
  explicit EngineError(const std::string &m) : std::runtime_error(m) {}
};

inline void cudaCheck(cudaError_t e, const char *what) {
  if (e != cudaSuccess) {
    throw EngineError(std::string("ERROR: ") + what + ": " + cudaGetErrorString(e));
  }
}
#define ARES_CUDA(expr) ::aresb::cudaCheck((expr), #expr)

// Every kernel launch site is followed by checkLastError (or noteLaunches for grouped launches),
// which also feeds the launch counter reported by bench.py as `gpu_launches`.
void noteLaunches(int n);
inline void checkLastError(const char *what) { noteLaunches(1); cudaCheck(cudaGetLastError(), what); }

// Runs `body` (returns the integer result) with the device selected; never lets an
// exception cross the C boundary.
template <typename F>
inline CGoCallResHandle guarded(const char *fn, int device, F &&body) {
  CGoCallResHandle h = {nullptr, nullptr};
  try {
    ARES_CUDA(cudaSetDevice(device));
    int64_t r = body();
    h.res = reinterpret_cast<void *>(static_cast<intptr_t>(r));
  } catch (const std::exception &e) {
    std::string m = std::string(fn) + ": " + e.what();
    h.pStrErr = strdup(m.c_str());
  } catch (...) {
    h.pStrErr = strdup((std::string(fn) + ": unknown error").c_str());
  }
  return h;
}

inline CGoCallResHandle unsupported(const char *fn, const char *why) {
  CGoCallResHandle h = {nullptr, nullptr};
  h.pStrErr = strdup((std::string(fn) + ": not supported by the B200 engine: " + why).c_str());
  return h;
}

// Number of SMs of the current device (cached per device).
int smCount();

// TEST SEAM: ARESDB_B200_TEST_HASH64_MASK=<hex> ANDs the 64-bit group-identity hash of a dimension row (legacy Sort and
// the finalize of the fused path), so that the merge rule for colliding hashes can be exercised
// (tests/test_hash_collisions.py).  ~0 unless the variable is set.
uint64_t testHash64Mask();

// Stream-ordered scratch from libmem's pool.  Freed when the object dies; the pool defers
// reuse until the recorded stream has passed this point.
struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;
  Scratch() = default;
  Scratch(size_t n, cudaStream_t s) { reset(n, s); }
  ~Scratch() { release(); }
  void reset(size_t n, cudaStream_t s);  // (re)allocate n bytes ordered on stream s
  void release();
  Scratch(const Scratch &) = delete;
  Scratch &operator=(const Scratch &) = delete;
  template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

inline int divUp(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace aresb

// libmem-internal pool hooks used by Scratch (defined in memory_pool.cu; not part of the Go ABI).
extern "C" void *aresbPoolAllocAsync(size_t bytes, void *stream);
extern "C" void aresbPoolFreeAsync(void *p, void *stream);
