// names_device.cuh -- device probe of the read-name table (names.cc), shared by the fused FASTQ kernel (ids in shared
// memory) and the SoA name kernel (ids in global memory).
// Replaces AHashSet::contains(read.id_bytes()) of /root/reference/src/lib/matcher.rs:631-638.
#pragma once
#include "kernels.hpp"
#include "names_hash.h"

namespace fqg {

// Walks the id [a, a+len) as zero-padded little-endian 32-bit words; `load` reads one ALIGNED word at an address of
// type A (32-bit shared-window address or 64-bit global address).
template <class A, class Load>
struct IdWords {
    A wp;
    uint32_t sh, len, lo, pos;
    Load load;
    __device__ __forceinline__ IdWords(A a, uint32_t n, Load l) : wp(a & ~(A)3), sh((uint32_t)(a & 3u) * 8u), len(n), lo(0), pos(0), load(l) {
        if (len) lo = load(wp);
    }
    __device__ __forceinline__ bool more() const { return pos < len; }
    __device__ __forceinline__ uint32_t next() {
        const uint32_t rem = len - pos;
        uint32_t hi = 0;
        if (rem + (sh >> 3) > 4u) hi = load(wp + 4u);        // the id continues into the next aligned word
        wp += 4u;
        uint32_t w = __funnelshift_r(lo, hi, sh);
        lo = hi;
        if (rem < 4u) w &= (1u << (8u * rem)) - 1u;
        pos += 4u;
        return w;
    }
};

template <class A, class Load>
__device__ __forceinline__ uint64_t names_hash_id(A a, uint32_t len, Load load) {
    uint64_t h = fqg_name_hash_init();
    for (IdWords<A, Load> it(a, len, load); it.more();) h = fqg_name_hash_word(h, it.next());
    return fqg_name_hash_finish(h, len);
}
// The table is an array of 32-byte buckets of four slot words (one L2 sector per probe); a bucket is filled front to
// back and overflows into the next one, so an empty slot ends the search.
template <class A, class Load>
__device__ __forceinline__ bool names_probe(const NamesView& nv, A a, uint32_t len, Load load, uint64_t h) {
    const uint64_t want = fqg_slot_word(h, len);
    uint32_t bkt = (uint32_t)(h >> 32) & nv.mask;
    for (;;) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(nv.slots) + 2u * (size_t)bkt;
        const ulonglong2 s01 = __ldg(p), s23 = __ldg(p + 1);
        const uint64_t sl[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            if (sl[k] == 0) return false;
            if (sl[k] == want) {      // 40 hash bits + the length agree: confirm on the bytes (pool entries are word-aligned, zero-padded)
                const uint32_t* q = reinterpret_cast<const uint32_t*>(nv.pool + __ldg(nv.slot_off + 4u * (size_t)bkt + k));
                uint32_t diff = 0;
                for (IdWords<A, Load> it(a, len, load); it.more() && !diff;) diff = it.next() ^ __ldg(q++);
                if (!diff) return true;
            }
        }
        bkt = (bkt + 1u) & nv.mask;
    }
}
template <class A, class Load>
__device__ __forceinline__ bool names_contains(const NamesView& nv, A a, uint32_t len, Load load) {
    return names_probe(nv, a, len, load, names_hash_id(a, len, load));
}

}  // namespace fqg
