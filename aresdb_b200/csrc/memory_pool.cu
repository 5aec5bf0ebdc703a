// memory_pool.cu — libmem.so: the memory / stream half of the cgo boundary
// (reference: cgoutils/memory.h:51-99, CUDA backend cgoutils/memory/cuda_malloc.cu).
//
// The reference pays one cudaMalloc + cudaMemset per DeviceAllocate and one device-wide
// synchronising cudaFree per DeviceFree; the Go driver issues ~10 such pairs per batch
// (scratch "stack frames", index / predicate vectors, result buffers).  Here every device
// has a size-class pool: blocks are cached on free and handed back on the next request of
// the same class, so steady-state batches never call the driver.  Two kinds of free:
//   * DeviceFree / deviceFree (caller unknown-stream): the block is quarantined behind an
//     event recorded on the legacy default stream — it becomes reusable once all work
//     submitted to blocking streams before the free has finished (what cudaFree's implicit
//     synchronisation guaranteed), but the host never blocks.
//   * aresbPoolFreeAsync (engine-internal scratch): stream-ordered; reusable immediately on
//     the same stream, or after its event on any other.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "aresdb_b200/device_memory.h"

namespace {

constexpr int kMaxDevices = 64;
constexpr size_t kMinBlock = 512;
constexpr size_t kLargeStep = 2u << 20;  // classes above 1 MiB are multiples of 2 MiB

size_t classOf(size_t bytes) {
  if (bytes <= kMinBlock) return kMinBlock;
  if (bytes <= (1u << 20)) {
    size_t c = kMinBlock;
    while (c < bytes) c <<= 1;
    return c;
  }
  return (bytes + kLargeStep - 1) / kLargeStep * kLargeStep;
}

struct Pending {
  void *ptr;
  size_t cls;
  cudaStream_t stream;  // nullptr + legacy=true: quarantined behind the legacy stream
  cudaEvent_t event;
};

struct DevicePool {
  std::mutex mu;
  std::multimap<size_t, void *> freeBlocks;
  std::vector<Pending> pending;
  std::unordered_map<void *, size_t> live;  // ptr -> class, for every block handed out
  std::vector<cudaEvent_t> eventCache;
  cudaStream_t zeroStream = nullptr;
  size_t cachedBytes = 0;
};

DevicePool g_pools[kMaxDevices];

char *cudaErr(const char *what, cudaError_t e) {
  char *buf = static_cast<char *>(malloc(160));
  snprintf(buf, 160, "ERROR when calling CUDA functions: %s: %s\n", what, cudaGetErrorString(e));
  return buf;
}

char *lastErr(const char *what) {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaErr(what, e);
}

int currentDevice() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

cudaEvent_t takeEvent(DevicePool &p) {
  if (!p.eventCache.empty()) {
    cudaEvent_t e = p.eventCache.back();
    p.eventCache.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}

// Moves every pending block whose event has completed to the free lists.  mu held.
void reapLocked(DevicePool &p, bool block) {
  size_t w = 0;
  for (size_t i = 0; i < p.pending.size(); i++) {
    Pending &q = p.pending[i];
    cudaError_t st = block ? cudaEventSynchronize(q.event) : cudaEventQuery(q.event);
    if (st == cudaSuccess) {
      p.freeBlocks.emplace(q.cls, q.ptr);
      p.eventCache.push_back(q.event);
    } else {
      if (st != cudaErrorNotReady) cudaGetLastError();
      p.pending[w++] = q;
    }
  }
  p.pending.resize(w);
}

void trimLocked(DevicePool &p) {
  reapLocked(p, true);
  for (auto &kv : p.freeBlocks) cudaFree(kv.second);
  p.cachedBytes = 0;
  p.freeBlocks.clear();
}

// Device must be current.  Returns nullptr and sets *err on failure.
void *poolAlloc(size_t bytes, cudaStream_t stream, bool streamOrdered, cudaError_t *err) {
  *err = cudaSuccess;
  if (bytes == 0) bytes = 1;
  const size_t cls = classOf(bytes);
  DevicePool &p = g_pools[currentDevice() % kMaxDevices];
  std::lock_guard<std::mutex> lock(p.mu);
  void *ptr = nullptr;
  if (streamOrdered) {
    // a block freed earlier on the same stream can be reused without waiting
    for (size_t i = 0; i < p.pending.size(); i++) {
      if (p.pending[i].cls == cls && p.pending[i].stream == stream && p.pending[i].stream != nullptr) {
        ptr = p.pending[i].ptr;
        p.cachedBytes -= cls;
        p.eventCache.push_back(p.pending[i].event);
        p.pending[i] = p.pending.back();
        p.pending.pop_back();
        break;
      }
    }
  }
  if (!ptr) {
    auto it = p.freeBlocks.find(cls);
    if (it == p.freeBlocks.end() && !p.pending.empty()) {
      reapLocked(p, false);
      it = p.freeBlocks.find(cls);
    }
    if (it != p.freeBlocks.end()) {
      ptr = it->second;
      p.freeBlocks.erase(it);
      p.cachedBytes -= cls;
    }
  }
  if (!ptr) {
    cudaError_t e = cudaMalloc(&ptr, cls);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trimLocked(p);  // give cached blocks back and retry once
      e = cudaMalloc(&ptr, cls);
      if (e != cudaSuccess) { cudaGetLastError(); *err = e; return nullptr; }
    }
  }
  p.live[ptr] = cls;
  return ptr;
}

void poolFree(void *ptr, cudaStream_t stream, bool streamOrdered) {
  if (!ptr) return;
  DevicePool &p = g_pools[currentDevice() % kMaxDevices];
  std::lock_guard<std::mutex> lock(p.mu);
  auto it = p.live.find(ptr);
  if (it == p.live.end()) {  // not ours (e.g. allocated before a trim by another library): hand to the driver
    cudaFree(ptr);
    return;
  }
  Pending q;
  q.ptr = ptr; q.cls = it->second;
  p.live.erase(it);
  q.stream = streamOrdered ? stream : nullptr;
  q.event = takeEvent(p);
  // legacy stream 0 waits for all earlier work on blocking streams of this device
  cudaEventRecord(q.event, streamOrdered ? stream : (cudaStream_t)0);
  p.cachedBytes += q.cls;
  p.pending.push_back(q);
}

}  // namespace

extern "C" {

void *aresbPoolAllocAsync(size_t bytes, void *stream) {
  cudaError_t e;
  return poolAlloc(bytes, (cudaStream_t)stream, true, &e);
}

void aresbPoolFreeAsync(void *p, void *stream) { poolFree(p, (cudaStream_t)stream, true); }

DeviceMemoryFlags GetFlags() { return DEVICE_MEMORY_IMPLEMENTATION_FLAG | HASH_REDUCTION_SUPPORT; }

CGoCallResHandle HostAlloc(size_t bytes) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaError_t e = cudaHostAlloc(&h.res, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) { cudaGetLastError(); h.res = nullptr; h.pStrErr = cudaErr("Allocate", e); return h; }
  memset(h.res, 0, bytes);
  return h;
}

CGoCallResHandle HostFree(void *p) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaFreeHost(p);
  h.pStrErr = lastErr("Free");
  return h;
}

CGoCallResHandle HostMemCpy(void *dst, const void *src, size_t bytes) {
  CGoCallResHandle h = {nullptr, nullptr};
  memcpy(dst, src, bytes);
  return h;
}

CGoCallResHandle CreateCudaStream(int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  cudaStream_t s = nullptr;
  cudaStreamCreate(&s);  // blocking w.r.t. the legacy stream, like the reference's
  h.res = s;
  h.pStrErr = lastErr("CreateCudaStream");
  return h;
}

CGoCallResHandle WaitForCudaStream(void *s, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  cudaStreamSynchronize((cudaStream_t)s);
  h.pStrErr = lastErr("WaitForCudaStream");
  return h;
}

CGoCallResHandle DestroyCudaStream(void *s, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  cudaStreamDestroy((cudaStream_t)s);
  h.pStrErr = lastErr("DestroyCudaStream");
  return h;
}

CGoCallResHandle DeviceAllocate(size_t bytes, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { cudaGetLastError(); h.pStrErr = cudaErr("DeviceAllocate", e); return h; }
  void *p = poolAlloc(bytes, nullptr, false, &e);
  if (!p) { h.pStrErr = cudaErr("DeviceAllocate", e); return h; }
  DevicePool &pool = g_pools[device % kMaxDevices];
  {
    std::lock_guard<std::mutex> lock(pool.mu);
    if (!pool.zeroStream) cudaStreamCreateWithFlags(&pool.zeroStream, cudaStreamNonBlocking);
  }
  // zero fill is complete on return (reference: cudaMalloc + cudaMemset, cuda_malloc.cu:97-104)
  cudaMemsetAsync(p, 0, bytes, pool.zeroStream);
  cudaStreamSynchronize(pool.zeroStream);
  h.res = p;
  h.pStrErr = lastErr("DeviceAllocate");
  return h;
}

CGoCallResHandle DeviceFree(void *p, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  poolFree(p, nullptr, false);
  h.pStrErr = lastErr("DeviceFree");
  return h;
}

CGoCallResHandle AsyncCopyHostToDevice(void *dst, void *src, size_t bytes, void *stream, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  h.pStrErr = lastErr("AsyncCopyHostToDevice");
  return h;
}

CGoCallResHandle AsyncCopyDeviceToDevice(void *dst, void *src, size_t bytes, void *stream, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  h.pStrErr = lastErr("AsyncCopyDeviceToDevice");
  return h;
}

CGoCallResHandle AsyncCopyDeviceToHost(void *dst, void *src, size_t bytes, void *stream, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  h.pStrErr = lastErr("AsyncCopyDeviceToHost");
  return h;
}

CGoCallResHandle GetDeviceCount() {
  CGoCallResHandle h = {nullptr, nullptr};
  int n = 0;
  cudaGetDeviceCount(&n);
  h.res = reinterpret_cast<void *>(static_cast<intptr_t>(n));
  h.pStrErr = lastErr("GetDeviceCount");
  return h;
}

CGoCallResHandle GetDeviceGlobalMemoryInMB(int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  h.res = reinterpret_cast<void *>(static_cast<intptr_t>(prop.totalGlobalMem / (1024 * 1024)));
  h.pStrErr = lastErr("GetDeviceGlobalMemoryInMB");
  return h;
}

CGoCallResHandle CudaProfilerStart() {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaProfilerStart();
  h.pStrErr = lastErr("cudaProfilerStart");
  return h;
}

CGoCallResHandle CudaProfilerStop() {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaDeviceSynchronize();
  cudaProfilerStop();
  h.pStrErr = lastErr("cudaProfilerStop");
  return h;
}

CGoCallResHandle GetDeviceMemoryInfo(size_t *freeSize, size_t *totalSize, int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaSetDevice(device);
  size_t f = 0, t = 0;
  cudaMemGetInfo(&f, &t);
  DevicePool &pool = g_pools[device % kMaxDevices];
  {
    std::lock_guard<std::mutex> lock(pool.mu);
    f += pool.cachedBytes;  // cached blocks are available to the next DeviceAllocate
  }
  if (freeSize) *freeSize = f;
  if (totalSize) *totalSize = t;
  h.pStrErr = lastErr("GetDeviceMemoryInfo");
  return h;
}

CGoCallResHandle deviceMalloc(void **devPtr, size_t size) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaError_t e;
  *devPtr = poolAlloc(size, nullptr, false, &e);
  if (!*devPtr) h.pStrErr = cudaErr("deviceMalloc", e);
  return h;
}

CGoCallResHandle deviceFree(void *devPtr) {
  CGoCallResHandle h = {nullptr, nullptr};
  poolFree(devPtr, nullptr, false);
  h.pStrErr = lastErr("deviceFree");
  return h;
}

CGoCallResHandle deviceMemset(void *devPtr, int value, size_t count) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaMemset(devPtr, value, count);
  h.pStrErr = lastErr("deviceMemset");
  return h;
}

CGoCallResHandle asyncCopyHostToDevice(void *dst, const void *src, size_t count, void *stream) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaMemcpyAsync(dst, src, count, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  h.pStrErr = lastErr("asyncCopyHostToDevice");
  return h;
}

CGoCallResHandle asyncCopyDeviceToHost(void *dst, const void *src, size_t count, void *stream) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaMemcpyAsync(dst, src, count, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  h.pStrErr = lastErr("asyncCopyDeviceToHost");
  return h;
}

CGoCallResHandle waitForCudaStream(void *stream) {
  CGoCallResHandle h = {nullptr, nullptr};
  cudaStreamSynchronize((cudaStream_t)stream);
  h.pStrErr = lastErr("waitForCudaStream");
  return h;
}

CGoCallResHandle DeviceMemoryPoolTrim(int device) {
  CGoCallResHandle h = {nullptr, nullptr};
  int n = 0;
  cudaGetDeviceCount(&n);
  int prev = currentDevice();
  for (int d = 0; d < n && d < kMaxDevices; d++) {
    if (device >= 0 && d != device) continue;
    cudaSetDevice(d);
    std::lock_guard<std::mutex> lock(g_pools[d].mu);
    trimLocked(g_pools[d]);
  }
  cudaSetDevice(prev);
  h.pStrErr = lastErr("DeviceMemoryPoolTrim");
  return h;
}

}  // extern "C"
