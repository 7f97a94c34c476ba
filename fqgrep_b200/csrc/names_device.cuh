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
__device__ __forceinline__ ulonglong4 names_load_bucket(const NamesView& nv, uint64_t h) {      // (as two 16-byte reads of one sector)
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(nv.slots) + 2u * (size_t)fqg_name_bucket(h, nv.mask);
    const ulonglong2 s01 = __ldg(p), s23 = __ldg(p + 1);
    return ulonglong4{s01.x, s01.y, s23.x, s23.y};
}
// The table is an array of 32-byte buckets of four slot words (one L2 sector per probe); a bucket is filled front to
// back and overflows into the next one, so an empty slot ends the search.  `first` is the already loaded home bucket.
template <class A, class Load>
__device__ __forceinline__ bool names_probe_from(const NamesView& nv, A a, uint32_t len, Load load, uint64_t h, ulonglong4 first) {
    const uint32_t tag = fqg_name_tag(h);
    uint32_t bkt = fqg_name_bucket(h, nv.mask);
    ulonglong4 b = first;
    for (;;) {
        const uint64_t sl[4] = {b.x, b.y, b.z, b.w};
        // which slots carry the tag (almost always none or one), and does the bucket end the search (an empty slot)?
        uint32_t cand = 0;
        bool open_end = false;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            if (sl[k] == 0) open_end = true;
            if (!open_end && (uint32_t)(sl[k] >> 32) == tag) cand |= 1u << k;
        }
        while (cand) {      // confirm on length and bytes: one 16-byte read covers ids of up to 12 bytes
            const uint32_t k = (uint32_t)__ffs((int)cand) - 1u;
            cand &= cand - 1u;
            const uint64_t slot = k == 0 ? sl[0] : k == 1 ? sl[1] : k == 2 ? sl[2] : sl[3];
            const uint4* q = reinterpret_cast<const uint4*>(nv.pool + (uint32_t)slot);
            const uint4 e = __ldg(q);
            if (e.x != len) continue;
            IdWords<A, Load> it(a, len, load);
            uint32_t diff = 0;
            if (it.more()) diff |= it.next() ^ e.y;
            if (it.more()) diff |= it.next() ^ e.z;
            if (it.more()) diff |= it.next() ^ e.w;
            const uint32_t* rest = reinterpret_cast<const uint32_t*>(q + 1);
            while (it.more() && !diff) diff = it.next() ^ __ldg(rest++);
            if (!diff) return true;
        }
        if (open_end) return false;
        bkt = (bkt + 1u) & nv.mask;
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(nv.slots) + 2u * (size_t)bkt;
        const ulonglong2 s01 = __ldg(p), s23 = __ldg(p + 1);
        b = ulonglong4{s01.x, s01.y, s23.x, s23.y};
    }
}
// The same search for an id of at most 16 bytes that is already at hand as four zero-padded little-endian words: nothing
// is read again, a stored id is compared with registers (one 16-byte read for ids of up to 12 bytes, one more word beyond).
__device__ __forceinline__ bool names_probe_words(const NamesView& nv, const uint32_t (&w)[4], uint32_t len, uint64_t h, ulonglong4 first) {
    const uint32_t tag = fqg_name_tag(h);
    uint32_t bkt = fqg_name_bucket(h, nv.mask);
    ulonglong4 b = first;
    for (;;) {
        const uint64_t sl[4] = {b.x, b.y, b.z, b.w};
        bool open_end = false;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            if (sl[k] == 0) open_end = true;
            if (!open_end && (uint32_t)(sl[k] >> 32) == tag) {
                const uint4* q = reinterpret_cast<const uint4*>(nv.pool + (uint32_t)sl[k]);
                const uint4 e = __ldg(q);
                if (e.x == len && e.y == w[0] && e.z == w[1] && e.w == w[2] && (len <= 12u || __ldg(reinterpret_cast<const uint32_t*>(q + 1)) == w[3])) return true;
            }
        }
        if (open_end) return false;
        bkt = (bkt + 1u) & nv.mask;
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(nv.slots) + 2u * (size_t)bkt;
        const ulonglong2 s01 = __ldg(p), s23 = __ldg(p + 1);
        b = ulonglong4{s01.x, s01.y, s23.x, s23.y};
    }
}
template <class A, class Load>
__device__ __forceinline__ bool names_probe(const NamesView& nv, A a, uint32_t len, Load load, uint64_t h) {
    return names_probe_from(nv, a, len, load, h, names_load_bucket(nv, h));
}
template <class A, class Load>
__device__ __forceinline__ bool names_contains(const NamesView& nv, A a, uint32_t len, Load load) {
    return names_probe(nv, a, len, load, names_hash_id(a, len, load));
}

}  // namespace fqg
