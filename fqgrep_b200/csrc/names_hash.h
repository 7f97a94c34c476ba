// names_hash.h -- the one hash shared by the host table builder and the device probe kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FQG_HD __host__ __device__ __forceinline__
#else
#define FQG_HD inline
#endif

// The id is hashed four bytes at a time: little-endian 32-bit words, the last one zero-padded; the length goes in at
// the end (ids are found by scanning for the first space, so the length is not known up front).  Multiply-xorshift
// rounds: cheap in device integer ops, avalanche finish.
FQG_HD uint64_t fqg_name_hash_init() { return 0x9E3779B97F4A7C15ull; }
FQG_HD uint64_t fqg_name_hash_word(uint64_t h, uint32_t w) {
    h = (h ^ w) * 0x9FB21C651E98DF25ull;
    return h ^ (h >> 32);
}
FQG_HD uint64_t fqg_name_hash_finish(uint64_t h, uint32_t n) {
    h = (h ^ n) * 0xD6E8FEB86659FD93ull;
    return h ^ (h >> 32);
}
FQG_HD uint64_t fqg_name_hash(const uint8_t* p, uint32_t n) {
    uint64_t h = fqg_name_hash_init();
    for (uint32_t i = 0; i < n; i += 4) {
        uint32_t w = 0;
        for (uint32_t k = 0; k < 4 && i + k < n; k++) w |= (uint32_t)p[i + k] << (8 * k);
        h = fqg_name_hash_word(h, w);
    }
    return fqg_name_hash_finish(h, n);
}
// The table: 32-byte buckets of four 8-byte slots, slot = tag << 32 | offset of the name's pool entry, 0 = empty.
// bucket = high half of the hash, tag = low half (never 0).  A pool entry sits on a 16-byte boundary:
// [u32 length][bytes, zero-padded to a multiple of four and to at least 12], so ids of up to 12 bytes are confirmed
// with ONE 16-byte read.  names longer than 2^24-1 bytes are rejected at construction.
FQG_HD uint32_t fqg_name_bucket(uint64_t h, uint32_t mask) { return (uint32_t)(h >> 32) & mask; }
FQG_HD uint32_t fqg_name_tag(uint64_t h) { return (uint32_t)h | 1u; }
