// murmur.cuh — MurmurHash3 x86_32 and x64_128 (Austin Appleby, public domain algorithm).
// The reference hashes the packed dimension row with these (query/utils.cu:113-155 for the
// 32-bit variant used by HashReduce, :157-241 for the 128-bit variant whose LOW word keys
// Sort/Reduce/HyperLogLog); results must be bit-identical.  Keys here live in registers as
// up to four little-endian 64-bit words (a dimension row is <= 32 bytes), zero padded.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ARES_HD __host__ __device__ __forceinline__
#else
#define ARES_HD inline
#endif

namespace aresb {

ARES_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
ARES_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

ARES_HD uint64_t fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

// Low 64 bits of murmur3_x64_128(key[0..len), seed).  w[] holds the key bytes little-endian,
// bytes beyond len MUST be zero.  len <= 32.
ARES_HD uint64_t murmur3_128_lo(const uint64_t w[4], int len, uint32_t seed) {
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = seed, h2 = seed;
  const int nblocks = len >> 4;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    if (i < nblocks) {
      uint64_t k1 = w[2 * i], k2 = w[2 * i + 1];
      k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
      h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
      k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
      h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
  }
  const int tail = len & 15;
  if (tail) {
    // zero padding makes the byte-wise tail switch of the canonical code a word operation
    uint64_t k1 = nblocks == 0 ? w[0] : (nblocks == 1 ? w[2] : 0);
    uint64_t k2 = nblocks == 0 ? w[1] : (nblocks == 1 ? w[3] : 0);
    if (tail > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
  }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2;
  return h1;
}

// murmur3_x86_32(key[0..len), seed); same key convention.
ARES_HD uint32_t murmur3_32(const uint64_t w[4], int len, uint32_t seed) {
  uint32_t h1 = seed;
  const int nblocks = len >> 2;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i < nblocks) {
      uint32_t k1 = (uint32_t)(w[i >> 1] >> ((i & 1) * 32));
      k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u;
      h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
    }
  }
  if (len & 3) {
    uint32_t k1 = nblocks < 8 ? (uint32_t)(w[nblocks >> 1] >> ((nblocks & 1) * 32)) : 0u;
    k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u;
    h1 ^= k1;
  }
  h1 ^= (uint32_t)len;
  h1 ^= h1 >> 16; h1 *= 0x85ebca6bu;
  h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u;
  h1 ^= h1 >> 16;
  return h1;
}

}  // namespace aresb
