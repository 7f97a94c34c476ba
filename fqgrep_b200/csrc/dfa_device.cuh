// dfa_device.cuh -- device-side DFA stepping shared by the SoA and the fused FASTQ kernels.
#pragma once
#include "device_common.cuh"

namespace fqg {

// Shared memory is addressed through explicit 32-bit shared-window addresses + ld.shared so the hot loop is
// LDS (not generic LD) whatever the optimiser can or cannot prove about pointer provenance.
__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// State encodings (see kernels.hpp): tiers 0/1 keep the state as the BYTE offset of its table row, so one
// 3-input add forms the LDS address; tier 2 keeps the row's entry index into the L2-resident table.
template <int TIER>
struct DfaView {
    uint32_t tab;          // smem address of the table (tiers 0,1)
    const uint32_t* t32;   // global (tier 2)
    uint32_t cls;          // smem address of the byte->class LUT (tiers 1,2); tier 1 stores class*2
    // `b2` is the input byte times two (tier 0) or the plain byte (tiers 1,2)
    __device__ __forceinline__ uint32_t step(uint32_t st, uint32_t b) const {
        if (TIER == 0) return lds_u16(tab + st + b);
        if (TIER == 1) return lds_u16(tab + st + lds_u8(cls + b));
        return __ldg(t32 + st + lds_u8(cls + b));
    }
    // byte k (0..3) of word w in the form step() wants
    __device__ __forceinline__ uint32_t pick(uint32_t w, int k) const {
        if (TIER == 0) return k == 0 ? (w << 1) & 0x1FEu : (w >> (8 * k - 1)) & 0x1FEu;
        return k == 3 ? w >> 24 : (w >> (8 * k)) & 0xFFu;
    }
};

// walk `len` bytes starting at shared address p (any alignment)
template <int TIER>
__device__ __forceinline__ uint32_t walk_shared(const DfaView<TIER>& d, uint32_t p, uint32_t len, uint32_t st) {
    const uint32_t sh = (p & 3u) * 8u;
    uint32_t wp = p & ~3u;
    uint32_t w0 = lds_u32_volatile(wp);
    const uint32_t nw = len >> 2;
#pragma unroll 2
    for (uint32_t k = 0; k < nw; k++) {
        wp += 4;
        const uint32_t w1 = lds_u32_volatile(wp);
        const uint32_t w = __funnelshift_r(w0, w1, sh);
        w0 = w1;
        st = d.step(st, d.pick(w, 0));
        st = d.step(st, d.pick(w, 1));
        st = d.step(st, d.pick(w, 2));
        st = d.step(st, d.pick(w, 3));
    }
    const uint32_t rem = len & 3u;
    if (rem) {
        const uint32_t w = __funnelshift_r(w0, lds_u32_volatile(wp + 4), sh);
        st = d.step(st, d.pick(w, 0));
        if (rem > 1) st = d.step(st, d.pick(w, 1));
        if (rem > 2) st = d.step(st, d.pick(w, 2));
    }
    return st;
}

// fallback for a batch whose span does not fit the smem slot (very long reads): straight from global memory
template <int TIER>
__device__ __noinline__ uint32_t walk_global(const DfaView<TIER>& d, const uint8_t* p, uint32_t len, uint32_t st) {
    for (uint32_t i = 0; i < len; i++) st = d.step(st, d.pick(__ldg(p + i), 0));
    return st;
}


}  // namespace fqg
