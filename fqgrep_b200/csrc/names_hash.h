// names_hash.h -- the one hash shared by the host table builder and the device probe kernel.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FQG_HD __host__ __device__ __forceinline__
#else
#define FQG_HD inline
#endif

// 64-bit multiply-xorshift over bytes; cheap in device integer ops, avalanche finish.
FQG_HD uint64_t fqg_name_hash(const uint8_t* p, uint32_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ n;
    for (uint32_t i = 0; i < n; i++) h = (h ^ p[i]) * 0x100000001B3ull;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    return h;
}
// names longer than 2^24-1 bytes are rejected at construction
FQG_HD uint64_t fqg_slot_word(uint64_t h, uint32_t len) { return ((h | (1ull << 63)) & 0xFFFFFFFFFF000000ull) | (len & 0xFFFFFFu); }
