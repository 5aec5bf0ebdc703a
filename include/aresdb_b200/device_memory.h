/*
 * device_memory.h — C ABI of libmem.so: device/host memory, streams and copies that the Go
 * caller (cgoutils/memory.go:17-19) and libalgorithm use.  Same symbols, argument meaning
 * and error behaviour as the reference's cgoutils/memory.h:51-99 (CUDA backend:
 * cgoutils/memory/cuda_malloc.cu), re-implemented over per-device size-class pools so
 * that the ~10 alloc/free pairs the Go driver issues per batch never reach cudaMalloc.
 *
 * Unlike the reference header this one only DECLARES (the reference defines fmtError
 * non-inline in the header, cgoutils/memory.h:36-41, which is why it needs two .so files).
 */
#ifndef ARESDB_B200_DEVICE_MEMORY_H_
#define ARESDB_B200_DEVICE_MEMORY_H_

#include <stddef.h>
#include <stdint.h>

#include "cgo_result.h"

#ifdef __cplusplus
extern "C" {
#endif

/* GetFlags() bits (ref: cgoutils/memory.h:28-32).  POOLED_MEMORY_FLAG is deliberately NOT
 * reported: it would switch the Go side to pooledDeviceAllocatorImpl
 * (query/device_allocator.go:158-167), which expects RMM semantics; our pool is internal. */
enum {
  DEVICE_MEMORY_IMPLEMENTATION_FLAG = 1,
  POOLED_MEMORY_FLAG = 1 << 1,
  HASH_REDUCTION_SUPPORT = 1 << 2
};
typedef uint32_t DeviceMemoryFlags;

DeviceMemoryFlags GetFlags();                                        /* ref: memory.h:51 */

CGoCallResHandle HostAlloc(size_t bytes);   /* pinned, portable, zero-filled  ref: :53, cuda_malloc.cu:44-52 */
CGoCallResHandle HostFree(void *p);                                  /* ref: :55 */
CGoCallResHandle HostMemCpy(void *dst, const void *src, size_t bytes); /* ref: :57 */

CGoCallResHandle CreateCudaStream(int device);                       /* ref: :59 */
CGoCallResHandle WaitForCudaStream(void *s, int device);             /* ref: :61 */
CGoCallResHandle DestroyCudaStream(void *s, int device);             /* ref: :63 */

/* Zero-filled device allocation; the zero fill is complete when the call returns
 * (ref: :65, cuda_malloc.cu:97-104 cudaMalloc + cudaMemset). */
CGoCallResHandle DeviceAllocate(size_t bytes, int device);
CGoCallResHandle DeviceFree(void *p, int device);                    /* ref: :67 */

CGoCallResHandle AsyncCopyHostToDevice(void *dst, void *src, size_t bytes, void *stream, int device);   /* ref: :69 */
CGoCallResHandle AsyncCopyDeviceToDevice(void *dst, void *src, size_t bytes, void *stream, int device); /* ref: :72 */
CGoCallResHandle AsyncCopyDeviceToHost(void *dst, void *src, size_t bytes, void *stream, int device);   /* ref: :75 */

CGoCallResHandle GetDeviceCount();                                   /* ref: :78 */
CGoCallResHandle GetDeviceGlobalMemoryInMB(int device);              /* ref: :80 */
CGoCallResHandle CudaProfilerStart();                                /* ref: :82 */
CGoCallResHandle CudaProfilerStop();                                 /* ref: :84 */
/* The reference's CUDA backend answers "Not supported" (cuda_malloc.cu:175-180); we answer. */
CGoCallResHandle GetDeviceMemoryInfo(size_t *freeSize, size_t *totalSize, int device); /* ref: :86 */

/* Internal set used by libalgorithm only; caller has already selected the device (ref: :89-99). */
CGoCallResHandle deviceMalloc(void **devPtr, size_t size);
CGoCallResHandle deviceFree(void *devPtr);
CGoCallResHandle deviceMemset(void *devPtr, int value, size_t count);
CGoCallResHandle asyncCopyHostToDevice(void *dst, const void *src, size_t count, void *stream);
CGoCallResHandle asyncCopyDeviceToHost(void *dst, const void *src, size_t count, void *stream);
CGoCallResHandle waitForCudaStream(void *stream);

/* Additive: return every cached block of `device` (-1: all devices) to the driver. */
CGoCallResHandle DeviceMemoryPoolTrim(int device);

#ifdef __cplusplus
}
#endif

#endif /* ARESDB_B200_DEVICE_MEMORY_H_ */
