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
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {      // (tables: constant for the lifetime of the kernel)
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;   // zero-extended straight into a 32-bit register: no 16-bit register shuffling in the hot loop
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// State encodings (see kernels.hpp).
//   tiers 0 and 1 keep a state as the 16-bit SHARED-MEMORY ADDRESS of its table row (entries are relocated by the
//   table's base address while the table is staged, see stage_table), so the dependent chain per lookup is a
//   single add + LDS.
//   tier 0: a row holds 256 four-base  A row holds 256 four-base
//           transitions indexed by the packed 2-bit codes of four consecutive bases (A=0 C=1 T=2 G=3, from
//           (byte>>1)&3), then the one-byte transitions indexed by byte class.  Four bases cost ONE table
//           lookup on the state-dependency chain; a word holding anything but ACGT (N, lowercase, protein,
//           IUPAC ...) is detected exactly by rebuilding the four letters from the codes, and takes the
//           one-byte path, so results never depend on the input being DNA.
//   tier 4: same idea for automata too large for 256-entry rows (up to ~1300 states): rows of 16 two-base entries
//           (+ the one-byte entries), two lookups per four bases.
//   tier 1: byte offset of the row in a class-indexed shared-memory table (one lookup per byte + class LUT).
//   tier 2: entry index of the row in a class-indexed table read through L2.
// tiers whose table is staged whole into shared memory with states = shared addresses
template <int TIER>
constexpr bool kSmemTable = (TIER == 0 || TIER == 1 || TIER == 4);
// tiers that consume packed 2-bit base codes (4 bases per lookup in tier 0, 2 per lookup in tier 4)
template <int TIER>
constexpr bool kPackedCodes = (TIER == 0 || TIER == 4);

template <int TIER>
struct DfaView {
    uint32_t tab;          // smem address of the table (tiers 0,1)
    const uint32_t* t32;   // global (tier 2)
    uint32_t cls;          // smem address of the byte -> class LUT (tier 0,1: class*2; tier 2: class)
    uint32_t hot_limit;    // tier 2: states (entry indices) below this have their rows cached in shared memory at `tab`
    // tiers 1, 2, 4: when the automaton has a dead state (^-anchored sets), a read is decided once its state is the dead
    // or the all-accepting one, usually after a few bases; settle != 0 makes the walk stop there (kernels.hpp)
    uint32_t settled_a = 0xFFFFFFFFu, settled_b = 0xFFFFFFFFu, settle = 0;
    // packed tiers, fused FASTQ kernel only: shared address of a 256-entry table "the four letters these four codes stand
    // for" (see step4_acgt<true>)
    uint32_t etab = 0;

    __device__ __forceinline__ uint32_t step1(uint32_t st, uint32_t byte) const {
        if (TIER == 0) return lds_u16(st + 512u + lds_u8(cls + byte));
        if (TIER == 4) return lds_u16(st + 32u + lds_u8(cls + byte));
        if (TIER == 1) return lds_u16(st + lds_u8(cls + byte));
        // tier 2: breadth-first state numbering puts the shallow (hot) states first; their rows live in shared memory
        const uint32_t idx = st + lds_u8(cls + byte);
        if (st < hot_limit) {
            uint32_t v;
            asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tab + (idx << 2)));
            return v;
        }
        return __ldg(t32 + idx);
    }
    // Four stream-ordered bytes in one little-endian word, tier 0 only: ONE lookup, taken unconditionally.
    // `bad` accumulates, exactly, whether any byte was not one of ACGT; the caller then re-walks the read with
    // step1 (rare: reads with N / lowercase / protein), so no per-word branch sits in the hot loop.
    //   code  x = (byte >> 1) & 3           A=0 C=1 T=2 G=3
    //   byte & ~6 is 0x41 for A,C,G and 0x50 for T  =>  ((byte & ~6) ^ 0x41) must equal 0x11 * [code == 2]
    //   index = c3 | c2<<2 | c1<<4 | c0<<6 (times two, byte offset) = (x * 0x40100401) >> 23; bit 0 is provably clear
    // ETAB: the exact "is this word pure ACGT" test as ONE more table lookup instead of arithmetic: the 8-bit code index
    // that selects the transition also selects the word the four letters must spell (etab), so bad |= w ^ etab[index] --
    // 2 ALU + 1 LDS instead of 4 ALU + 1 IMAD per word.  Right for the fused FASTQ kernel, which is bound by integer issue
    // with shared memory at 47 % of its wavefront peak; wrong for the SoA kernel, which sits at 86 % of it.
    template <bool ETAB = false>
    __device__ __forceinline__ uint32_t step4_acgt(uint32_t st, uint32_t w, uint32_t& bad) const {
        const uint32_t x = (w >> 1) & 0x03030303u;
        const uint32_t prod = x * 0x40100401u;                               // packed codes of the word in bits 24..31 (22, 23 clear)
        if (ETAB) {
            bad |= w ^ lds_u32(etab + (prod >> 22));
        } else {
            const uint32_t t = (x >> 1) & ~x & 0x01010101u;
            bad |= ((w & 0xF9F9F9F9u) ^ 0x41414141u) ^ (t * 0x11u);
        }
        const uint32_t idx2 = prod >> 23;                                    // packed codes, times two
        if (TIER == 4) {                                                     // tier 4: rows of 16 two-base entries
            st = lds_u16(st + ((idx2 >> 4) & 0x1Eu));                        // first two bases
            return lds_u16(st + (idx2 & 0x1Eu));                             // last two bases
        }
        return lds_u16(st + idx2);
    }
};

// etab[e] for e = c0 << 6 | c1 << 4 | c2 << 2 | c3 (c0 = the first of the four bases = the low byte of the word):
// the little-endian word of the letters the codes stand for (A=0 C=1 T=2 G=3)
__device__ __forceinline__ void stage_etab(uint32_t* s_etab, uint32_t tid, uint32_t nthreads) {
    for (uint32_t e = tid; e < 256u; e += nthreads) {
        uint32_t word = 0;
#pragma unroll
        for (uint32_t k = 0; k < 4u; k++) {
            const uint32_t c = (e >> (6u - 2u * k)) & 3u;
            word |= (c == 0u ? 0x41u : c == 1u ? 0x43u : c == 2u ? 0x54u : 0x47u) << (8u * k);
        }
        s_etab[e] = word;
    }
}

// Copies a tier-0/1 table into shared memory, turning row offsets into absolute shared addresses.
// The table sits at the start of dynamic shared memory and is < 62 KB, so addresses fit 16 bits.
__device__ __forceinline__ void stage_table(const void* gtable, uint32_t table_bytes, void* s_tab, uint32_t tid, uint32_t nthreads) {
    const uint4* src = reinterpret_cast<const uint4*>(gtable);
    uint4* dst = reinterpret_cast<uint4*>(s_tab);
    const uint32_t base = smem_addr(s_tab);
    if (tid == 0 && base + table_bytes > 0xFFFFu) __trap();   // cannot happen with the tier limits in api.cu
    const uint32_t add = base | (base << 16);
    for (uint32_t i = tid; i < table_bytes / 16; i += nthreads) {
        uint4 v = __ldg(src + i);
        v.x += add; v.y += add; v.z += add; v.w += add;    // two u16 lanes per word; sums stay below 2^16
        dst[i] = v;
    }
}

// walk `len` bytes starting at shared address p (any alignment), one byte per lookup
template <int TIER>
__device__ __noinline__ uint32_t walk_shared_bytes(const DfaView<TIER>& d, uint32_t p, uint32_t len, uint32_t st) {
    for (uint32_t i = 0; i < len; i++) {
        uint32_t b;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(p + i));
        st = d.step1(st, b);
    }
    return st;
}

#ifndef FQG_WALK_UNROLL
#define FQG_WALK_UNROLL 4
#endif
constexpr int kWalkUnroll = FQG_WALK_UNROLL;      // words of the read per trip of the walk loop

// walk `len` bytes starting at shared address p (any alignment)
// SETTLE: leave the walk once the state is the dead or the all-accepting one (only instantiated where the automaton
// has a dead state: the check costs a crossbar-bound loop like tier 4's 20 % when it sits there unused)
template <int TIER, bool SETTLE, bool ETAB = false>
__device__ __forceinline__ uint32_t walk_shared_impl(const DfaView<TIER>& d, uint32_t p, uint32_t len, uint32_t st0) {
    const uint32_t sh = (p & 3u) * 8u;
    uint32_t wp = p & ~3u;
    uint32_t w0 = lds_u32_volatile(wp);
    const uint32_t nw = len >> 2;
    uint32_t st = st0, bad = 0;
#pragma unroll kWalkUnroll
    for (uint32_t k = 0; k < nw; k++) {
        wp += 4;
        const uint32_t w1 = lds_u32_volatile(wp);
        const uint32_t w = __funnelshift_r(w0, w1, sh);
        w0 = w1;
        if (kPackedCodes<TIER>) st = d.template step4_acgt<ETAB>(st, w, bad);
        else {
            st = d.step1(st, w & 0xFFu);
            st = d.step1(st, (w >> 8) & 0xFFu);
            st = d.step1(st, (w >> 16) & 0xFFu);
            st = d.step1(st, w >> 24);
        }
        // decided for good; a word with a non-ACGT byte (`bad`) makes the packed tiers re-walk the read below instead
        if (SETTLE && !bad && (st == d.settled_a || st == d.settled_b)) return st;
    }
    if (kPackedCodes<TIER> && bad) return walk_shared_bytes(d, p, len, st0);
    const uint32_t rem = len & 3u;
    if (rem) {
        const uint32_t w = __funnelshift_r(w0, lds_u32_volatile(wp + 4), sh);
        st = d.step1(st, w & 0xFFu);
        if (rem > 1) st = d.step1(st, (w >> 8) & 0xFFu);
        if (rem > 2) st = d.step1(st, (w >> 16) & 0xFFu);
    }
    return st;
}

template <int TIER, bool ETAB = false>
__device__ __forceinline__ uint32_t walk_shared(const DfaView<TIER>& d, uint32_t p, uint32_t len, uint32_t st0) {
    if (TIER != 0 && d.settle) return walk_shared_impl<TIER, true, ETAB>(d, p, len, st0);      // warp-uniform choice
    return walk_shared_impl<TIER, false, ETAB>(d, p, len, st0);
}

// fallback for reads that do not fit the shared-memory staging (very long reads): straight from global memory
template <int TIER>
__device__ __noinline__ uint32_t walk_global(const DfaView<TIER>& d, const uint8_t* p, uint32_t len, uint32_t st) {
    for (uint32_t i = 0; i < len; i++) st = d.step1(st, __ldg(p + i));
    return st;
}

}  // namespace fqg
