// radix_sort.cuh — stable LSD radix sort of (u64 key, payload) pairs, hand-written for sm_100a
// (replaces thrust::stable_sort_by_key of the reference's Sort / HyperLogLog, no CUB/Thrust).
//
// 8-bit digits.  Every pass is three launches over a fixed grid of G persistent blocks, each
// owning one contiguous chunk of the input:
//   1. digitHistogram   per-block 256-bin histogram (shared-memory atomics)        reads  8 B/elem
//   2. scanHistograms   one block: exclusive scan of the bin-major G x 256 table
//   3. scatterByDigit   per tile: warp-synchronous match-based stable ranking, tile staged in
//                       shared memory in digit order, coalesced run-wise write-out    r+w 24 B/elem
// Passes whose digit is constant over the whole input (detected from the histogram) are skipped
// by the caller-visible wrapper only when that does not change the ping-pong parity.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace aresb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                       // keys per thread per tile
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per tile
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

template <typename V>
void radixSortPairs(uint64_t *keys, V *vals, uint64_t *keysTmp, V *valsTmp, int n, int beginBit, int endBit,
                    cudaStream_t s);

// Sort of (key, index) pairs whose indices are pairwise distinct (finalize sorts group hashes with an iota
// payload): ties are broken by the index, so the result equals the stable sort of the iota-ordered input.
// n <= kSmallSortMax runs as ONE launch of one CTA (top-digit bucketing in shared memory, then an in-bucket
// rank sort) instead of 3 launches per 8-bit digit; larger inputs take radixSortPairs.
constexpr int kSmallSortMax = 32768;
void sortKeyIndexPairs(uint64_t *keys, uint32_t *index, uint64_t *keysTmp, uint32_t *indexTmp, int n, int keyBits,
                       cudaStream_t s);

// Bytes of scratch radixSortPairs needs besides the ping-pong buffers.
size_t radixSortScratchBytes(int n);

}  // namespace aresb
