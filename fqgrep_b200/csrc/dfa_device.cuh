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

// State encodings (see kernels.hpp).
//   tier 0: a state is the BYTE offset of its row in the shared-memory table.  A row holds 256 four-base
//           transitions indexed by the packed 2-bit codes of four consecutive bases (A=0 C=1 T=2 G=3, from
//           (byte>>1)&3), then the one-byte transitions indexed by byte class.  Four bases cost ONE table
//           lookup on the state-dependency chain; a word holding anything but ACGT (N, lowercase, protein,
//           IUPAC ...) is detected exactly by rebuilding the four letters from the codes, and takes the
//           one-byte path, so results never depend on the input being DNA.
//   tier 1: byte offset of the row in a class-indexed shared-memory table (one lookup per byte + class LUT).
//   tier 2: entry index of the row in a class-indexed table read through L2.
template <int TIER>
struct DfaView {
    uint32_t tab;          // smem address of the table (tiers 0,1)
    const uint32_t* t32;   // global (tier 2)
    uint32_t cls;          // smem address of the byte -> class LUT (tier 0,1: class*2; tier 2: class)

    __device__ __forceinline__ uint32_t step1(uint32_t st, uint32_t byte) const {
        if (TIER == 0) return lds_u16(tab + st + 512u + lds_u8(cls + byte));
        if (TIER == 1) return lds_u16(tab + st + lds_u8(cls + byte));
        return __ldg(t32 + st + lds_u8(cls + byte));
    }
    // four stream-ordered bytes in one little-endian word
    __device__ __forceinline__ uint32_t step4(uint32_t st, uint32_t w) const {
        if (TIER == 0) {
            const uint32_t x = (w >> 1) & 0x03030303u;                       // 2-bit code per byte
            const uint32_t t = (x >> 1) & ~x & 0x01010101u;                  // bytes whose code is 2 ('T')
            const uint32_t letters = x * 2u + 0x41414141u + t * 0x0Fu;       // the ACGT word those codes stand for
            if (letters == w) return lds_u16(tab + st + (((x * 0x01041040u) >> 24) << 1));
        }
        st = step1(st, w & 0xFFu);
        st = step1(st, (w >> 8) & 0xFFu);
        st = step1(st, (w >> 16) & 0xFFu);
        return step1(st, w >> 24);
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
        st = d.step4(st, w);
    }
    const uint32_t rem = len & 3u;
    if (rem) {
        const uint32_t w = __funnelshift_r(w0, lds_u32_volatile(wp + 4), sh);
        st = d.step1(st, w & 0xFFu);
        if (rem > 1) st = d.step1(st, (w >> 8) & 0xFFu);
        if (rem > 2) st = d.step1(st, (w >> 16) & 0xFFu);
    }
    return st;
}

// fallback for reads that do not fit the shared-memory staging (very long reads): straight from global memory
template <int TIER>
__device__ __noinline__ uint32_t walk_global(const DfaView<TIER>& d, const uint8_t* p, uint32_t len, uint32_t st) {
    for (uint32_t i = 0; i < len; i++) st = d.step1(st, __ldg(p + i));
    return st;
}

}  // namespace fqg
