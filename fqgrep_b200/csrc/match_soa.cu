// match_soa.cu -- the read-matching kernel over SoA reads resident in HBM (sm_100a).
//
// Computes, for every read i, what the reference computes per record on a CPU pool thread:
//     Matcher::read_match(&record)            /root/reference/src/lib/matcher.rs:164-175
//  =  (f(seq) || (reverse_complement && f(rc(seq)))) != invert_match
// where f is bases_match of FixedString / FixedStringSet / Regex / RegexSet (matcher.rs:203-205, 249-251,
// 506-508, 569-571).  f and the reverse-complement branch are both folded into one host-compiled DFA
// (regex_compile.cc / ac_build.cc), so a read is walked exactly once, left to right.
//
// Shape of the kernel (B200: 148 SMs, 227 KB smem/CTA, HBM-bound byte work, no tensor cores):
//   * persistent grid, one CTA per SM, up to 32 warps; the DFA table lives in shared memory (tiers 0/1);
//   * a warp owns a batch of 32 consecutive reads.  Their bytes are contiguous in HBM, so ONE bulk async copy
//     (cp.async.bulk global->shared, the TMA engine, completion on an mbarrier) brings the 16 B-aligned span
//     into the warp's private smem slot -- fully coalesced DRAM traffic, no register staging;
//   * ONE slot per warp (32 resident warps for 150-base reads): while the 32 lanes walk batch k, one read per lane,
//     out of shared memory (aligned 32-bit LDS + funnel shift, 4 bases per load), batch k+1 is pulled into L2 with
//     cp.async.bulk.prefetch.L2, so the copy issued when the slot frees only pays L2 latency; the offsets of a batch
//     are requested two batches ahead.  (A second slot per warp, or a CTA-wide ring of slots, leaves room for
//     fewer warps and measured slower on every tier.)
//   * epilogue: ballot -> one mask word per batch (optionally OR-ed with the mate's mask: the pair rule of
//     src/main.rs:749-755), popcount accumulated per warp and added to the global count once.
#include <type_traits>

#include "dfa_device.cuh"
#include "names_device.cuh"
#include "kernels.hpp"

namespace fqg {
namespace {

// ---- tier 3: q-gram sampling filter (gram_filter.cc) --------------------------------------------------------------
struct GramView {
    uint32_t bitmap;          // shared address of the bitmap
    DeviceGram g;
    const uint32_t* t32;      // exact fallback automaton (tier-2 encoding)
    uint32_t cls;             // shared address of the byte -> class LUT
    uint32_t dfa_start, dfa_accept;
};
// Exact check of one filter hit: is the gram that ends at base e of the read at shared address p (len bases) a gram of
// some pattern, and does that pattern occur around it?  A false positive costs one 16-byte L2 read (an empty slot).
__device__ __forceinline__ bool gram_verify(const GramView& v, uint32_t p, uint32_t len, uint32_t gram, uint32_t e) {
    uint32_t i = (gram * 0x85EBCA6Bu) >> v.g.slot_shift;
    for (;;) {
        const uint4 slot = __ldg(v.g.slots + i);      // {gram key, pool offset, pattern length << 5 | offset, -}
        if (slot.z == 0xFFFFFFFFu) return false;
        if (slot.x == gram) {
            const uint32_t plen = slot.z >> 5, j = slot.z & 31u;
            const int32_t start = (int32_t)e - (int32_t)(v.g.q - 1u) - (int32_t)j;     // where the pattern would begin in the read
            if (start >= 0 && (uint32_t)start + plen <= len) {
                // 16 pattern bytes per L2 read against the (unaligned) text words in shared memory
                const uint4* pat = reinterpret_cast<const uint4*>(v.g.pool + slot.y);
                const uint32_t a = p + (uint32_t)start, sh = (a & 3u) * 8u;
                uint32_t wp = a & ~3u, w0 = lds_u32_volatile(wp), diff = 0;
                for (uint32_t k = 0; k < plen && !diff; k += 16u) {
                    const uint4 c = __ldg(pat + (k >> 4));
                    const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t rem = plen - k > 4u * t ? plen - k - 4u * t : 0u;      // pattern bytes left in this word
                        const uint32_t w1 = rem ? lds_u32_volatile(wp + 4u) : 0u;
                        wp += 4u;
                        const uint32_t text = __funnelshift_r(w0, w1, sh);
                        w0 = w1;
                        const uint32_t keep = rem >= 4u ? 0xFFFFFFFFu : rem ? (1u << (8u * rem)) - 1u : 0u;
                        diff |= (text ^ cw[t]) & keep;
                    }
                }
                if (!diff) return true;
            }
        }
        i = (i + 1u) & v.g.slot_mask;
    }
}

// aligned sampling: 2-bit codes of one aligned text word -> one byte (byte k of the word in bits 2k..2k+1), shifted
// into the rolling register `lo` (newest word in the low byte).  gram_filter.cc builds the keys the same way.
__device__ __forceinline__ uint32_t pack_word(uint32_t w, uint32_t lo) {
    return __byte_perm((w & 0x06060606u) * 0x00820820u, lo, 0x6543);
}
__device__ __forceinline__ uint2 lds_u64_volatile(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t gram_at(uint32_t p, uint32_t e, uint32_t q, uint32_t qmask, bool aligned) {
    if (aligned) {          // aligned sampling: re-pack the four staged words that end at base e (see pack_word)
        uint32_t wp = p + e - 15u, g = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) g = pack_word(lds_u32_volatile(wp + 4u * k), g);
        return g & qmask;
    }
    uint32_t g = 0;
    for (uint32_t i = e + 1u - q; i <= e; i++) g = (g << 2) | ((lds_u8(p + i) >> 1) & 3u);
    return g;
}

// true iff some pattern of the set occurs in the read at shared address p.  Must be called by all 32 lanes of the warp
// together (lanes without a read pass len = 0): the filter hits of the whole batch are verified cooperatively.
__device__ __forceinline__ bool gram_walk(const GramView& v, uint32_t p, uint32_t len) {
    // pinned into registers (asm volatile cannot be rematerialised) so the probe does not re-derive them per gram
    uint32_t bitmap, bshift, qmask;
    asm volatile("mov.u32 %0, %1;" : "=r"(bitmap) : "r"(v.bitmap));
    asm volatile("mov.u32 %0, %1;" : "=r"(bshift) : "r"(v.g.bitmap_shift));
    asm volatile("mov.u32 %0, %1;" : "=r"(qmask) : "r"(v.g.qmask));
    const uint32_t q = v.g.q, stride = v.g.s;
    bool found = false, exact_walk = false;
    const bool aligned = (stride & 3u) == 0;
    uint32_t hits0 = 0, hits1 = 0, probes_done = 0, e0 = 0;
    if (len >= v.g.min_len) {
        // Filter hits are recorded branch-free as one bit per probe (probe i ends at base e0 + i*stride); the gram of a
        // hit is re-read from shared memory later.  64 probes cover reads up to ~64*stride bases; longer reads take
        // the exact automaton walk.
        // hits0 takes probes 0..31 (first probe ends up in the highest used bit), hits1 probes 32..63.
        auto probe_into = [&](uint32_t gram, uint32_t& acc) {
            const uint32_t h = gram * 0x9E3779B1u;
            const uint32_t wi = h >> bshift;                                 // word of the blocked Bloom filter (top bits)
            // three bit positions inside it from bits 2..16 of the same product (funnel shifts wrap the amount mod 32)
            const uint32_t need = __funnelshift_l(0u, 1u, h >> 2) | __funnelshift_l(0u, 1u, h >> 7) | __funnelshift_l(0u, 1u, h >> 12);
            uint32_t word;
            asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(bitmap + (wi << 2)));
            acc = (acc << 1) | (uint32_t)((word & need) == need);
        };
        auto probe = [&](uint32_t gram, uint32_t) {
            if (probes_done < 32u) probe_into(gram, hits0); else probe_into(gram, hits1);
            probes_done++;
        };
        uint32_t lo = 0, hi = 0;              // 2-bit codes of the last 16 (+16) bases, newest in the low bits
        uint32_t n_probes = 0;
        if (aligned) {
            // Aligned sampling: the sampled grams end on the last byte of an aligned 8-byte unit (4-byte when
            // stride == 4) of the STAGING BUFFER, whatever the read's own alignment: aligned LDS.64, no funnel shifts,
            // one multiply packs a word.  Sampled ends are e0, e0 + stride, ... <= len - 1 with e0 in [q-1, q-1+7]:
            // every run of `stride` consecutive gram ends inside the read still contains one of them.
            const uint32_t unit = (stride & 7u) ? 3u : 7u;
            const uint32_t end0 = (p + q - 1u) | unit;          // shared address of the first sampled end
            e0 = end0 - p;
            n_probes = len > e0 ? (len - 1u - e0) / stride + 1u : 0u;
            if (n_probes > 64u) exact_walk = true;
            else if (n_probes) {
                uint32_t wp = end0 - 15u;
#pragma unroll
                for (int k = 0; k < 4; k++) lo = pack_word(lds_u32_volatile(wp + 4u * k), lo);
                wp += 16u;
                probe_into(lo & qmask, hits0);
                const uint32_t n_first = n_probes < 32u ? n_probes : 32u;
                auto groups = [&](auto units_c) {       // units_c 8-byte units per probe
                    constexpr uint32_t E = decltype(units_c)::value;
                    for (uint32_t g = 1; g < n_first; g++) {
#pragma unroll
                        for (uint32_t j = 0; j < E; j++) {
                            const uint2 w = lds_u64_volatile(wp + 8u * j);
                            lo = pack_word(w.y, pack_word(w.x, lo));
                        }
                        wp += 8u * E;
                        probe_into(lo & qmask, hits0);
                    }
                    for (uint32_t g = 32; g < n_probes; g++) {
#pragma unroll
                        for (uint32_t j = 0; j < E; j++) {
                            const uint2 w = lds_u64_volatile(wp + 8u * j);
                            lo = pack_word(w.y, pack_word(w.x, lo));
                        }
                        wp += 8u * E;
                        probe_into(lo & qmask, hits1);
                    }
                };
                switch (stride) {
                    case 8: groups(std::integral_constant<uint32_t, 1>{}); break;
                    case 16: groups(std::integral_constant<uint32_t, 2>{}); break;
                    case 24: groups(std::integral_constant<uint32_t, 3>{}); break;
                    case 32: groups(std::integral_constant<uint32_t, 4>{}); break;
                    default:          // stride 4: one aligned word per probe
                        for (uint32_t g = 1; g < n_probes; g++) {
                            lo = pack_word(lds_u32_volatile(wp), lo);
                            wp += 4u;
                            if (g < 32u) probe_into(lo & qmask, hits0); else probe_into(lo & qmask, hits1);
                        }
                }
                probes_done = n_probes;
            }
        } else {
            const uint32_t sh0 = (p & 3u) * 8u, nw = len >> 2;
            uint32_t wp = p & ~3u;
            uint32_t w0 = lds_u32_volatile(wp);
            e0 = q - 1u;
            if ((len - q) / stride + 1u > 64u) exact_walk = true;
            else {
                uint32_t next_e = e0;         // index of the base that ends the next sampled gram
                auto probe_upto = [&](uint32_t latest) {
                    while (next_e <= latest) {
                        probe(__funnelshift_r(lo, hi, 2u * (latest - next_e)) & qmask, 0u);
                        next_e += stride;
                    }
                };
                for (uint32_t k = 0; k < nw; k++) {
                    wp += 4;
                    const uint32_t w1 = lds_u32_volatile(wp);
                    const uint32_t w = __funnelshift_r(w0, w1, sh0);
                    w0 = w1;
                    const uint32_t x = (w >> 1) & 0x03030303u;
                    hi = (hi << 8) | (lo >> 24);
                    lo = (lo << 8) | ((x * 0x40100401u) >> 24);
                    probe_upto(4u * k + 3u);
                }
                const uint32_t rem = len & 3u;
                if (rem) {
                    const uint32_t w = __funnelshift_r(w0, lds_u32_volatile(wp + 4), sh0);
                    for (uint32_t r = 0; r < rem; r++) {
                        hi = (hi << 2) | (lo >> 30);
                        lo = (lo << 2) | ((w >> (8u * r + 1u)) & 3u);
                        probe_upto(4u * nw + r);
                    }
                }
            }
        }
    }
    // Exact verification, shared by the warp.  Hits are rare but unevenly spread (a lane-private loop makes the warp
    // wait for its unluckiest lane, one L2 round trip per hit), so they are numbered across the batch and dealt out one
    // per lane: task t belongs to the lane `owner` whose running hit count first exceeds t; the verifying lane fetches
    // the owner's read (shared address, length, hit bits) by shuffle.  Warp-uniform control flow around every collective.
    {
        const uint32_t mine = (uint32_t)__popc(hits0) + (uint32_t)__popc(hits1);
        uint32_t inc = mine;                       // inclusive prefix sum of the per-lane hit counts
#pragma unroll
        for (uint32_t d = 1; d < 32u; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, inc, d);
            if ((threadIdx.x & 31u) >= d) inc += t;
        }
        const uint32_t total = __shfl_sync(kFullMask, inc, 31);
        uint32_t found_lanes = 0;
        for (uint32_t first = 0; first < total; first += 32u) {
            const uint32_t task = first + (threadIdx.x & 31u);
            uint32_t owner = 0;                    // number of lanes whose inclusive count is <= task
#pragma unroll
            for (uint32_t step = 16; step; step >>= 1) {
                const uint32_t t = __shfl_sync(kFullMask, inc, owner + step - 1u);
                if (t <= task) owner += step;
            }
            owner &= 31u;                          // idle lanes (task >= total) read lane 0..31 harmlessly
            const uint32_t o_inc = __shfl_sync(kFullMask, inc, owner), o_mine = __shfl_sync(kFullMask, mine, owner);
            const uint32_t o_h0 = __shfl_sync(kFullMask, hits0, owner), o_h1 = __shfl_sync(kFullMask, hits1, owner);
            const uint32_t o_p = __shfl_sync(kFullMask, p, owner), o_len = __shfl_sync(kFullMask, len, owner);
            const uint32_t o_e0 = __shfl_sync(kFullMask, e0, owner), o_n = __shfl_sync(kFullMask, probes_done, owner);
            bool ok = false;
            if (task < total) {
                uint32_t k = task - (o_inc - o_mine);          // which of the owner's hits, lowest bit of hits0 first
                const uint32_t c0 = (uint32_t)__popc(o_h0);
                const bool in0 = k < c0;
                uint32_t h = in0 ? o_h0 : o_h1;
                if (!in0) k -= c0;
                for (; k; k--) h &= h - 1u;
                const uint32_t bit = (uint32_t)__ffs((int)h) - 1u;
                // bit b of hits0 is probe (n0 - 1 - b), bit b of hits1 is probe (32 + n1 - 1 - b)
                const uint32_t n0 = o_n < 32u ? o_n : 32u, n1 = o_n - n0;
                const uint32_t probe_idx = in0 ? n0 - 1u - bit : 32u + n1 - 1u - bit;
                const uint32_t e = o_e0 + probe_idx * stride;
                ok = gram_verify(v, o_p, o_len, gram_at(o_p, e, q, qmask, aligned), e);
            }
            found_lanes |= __reduce_or_sync(kFullMask, ok ? 1u << owner : 0u);
        }
        found = (found_lanes >> (threadIdx.x & 31u)) & 1u;
    }
    if (exact_walk) {          // very long read: exact automaton walk through L2
        uint32_t st = v.dfa_start;
        for (uint32_t i = 0; i < len; i++) st = __ldg(v.t32 + st + lds_u8(v.cls + lds_u8(p + i)));
        found = st >= v.dfa_accept;
    }
    return found;
}

struct BatchDesc {
    uint32_t s, e;        // this lane's read [s,e) as offsets into bases
    uintptr_t src;        // 16 B-aligned global address the warp's copy starts at (warp-uniform)
    uint32_t off;         // this lane's first byte relative to src
    uint32_t bytes;       // bytes to copy (multiple of 16); 0 = nothing to copy (warp-uniform)
    bool fits;            // warp-uniform: span fits the slot
};

template <int TIER>
__global__ void __launch_bounds__(1024, 1)
match_soa_kernel(SoaBatch b, DeviceDfa dfa, uint32_t slot_bytes, uint32_t table_smem_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (b.lines_dev) {      // read count known on the device only (FASTQ indexer): b.n_reads was the capacity
        const unsigned long long n = *b.lines_dev >> 2;
        if (n < b.n_reads) b.n_reads = (uint32_t)n;
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    // ---- smem carve-up: [table][class LUT 256][mbarriers][slots]
    uint16_t* s_tab = reinterpret_cast<uint16_t*>(smem);
    uint8_t* s_cls = smem + table_smem_bytes;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_cls + 256);
    uint8_t* s_slots = reinterpret_cast<uint8_t*>(s_bar + 32);
    s_slots = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_slots) + 127) & ~uintptr_t(127));

    if (kSmemTable<TIER>) stage_table(dfa.table, dfa.table_bytes, s_tab, threadIdx.x, blockDim.x);
    if (TIER == 2)
        for (uint32_t i = threadIdx.x; i < dfa.hot_entries; i += blockDim.x) reinterpret_cast<uint32_t*>(s_tab)[i] = __ldg(reinterpret_cast<const uint32_t*>(dfa.table) + i);
    if (TIER == 3) {
        const uint4* src = reinterpret_cast<const uint4*>(dfa.gram.bitmap);
        uint4* dst = reinterpret_cast<uint4*>(s_tab);
        for (uint32_t i = threadIdx.x; i < dfa.gram.bitmap_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_cls[i] = __ldg(dfa.byte_class + i);
    if (lane == 0)
        mbar_init(smem_addr(&s_bar[warp]), 1);
    mbar_fence_init();
    __syncthreads();

    DfaView<(TIER == 3 ? 2 : TIER)> view{smem_addr(s_tab), reinterpret_cast<const uint32_t*>(dfa.table), smem_addr(s_cls), TIER == 2 ? dfa.hot_entries : 0u};
    {
        const uint32_t rel = kSmemTable<TIER> ? smem_addr(s_tab) : 0u;      // same relocation as the start state below
        view.settled_a = dfa.settled[0] == 0xFFFFFFFFu ? 0xFFFFFFFFu : dfa.settled[0] + rel;
        view.settled_b = dfa.settled[1] == 0xFFFFFFFFu ? 0xFFFFFFFFu : dfa.settled[1] + rel;
        view.settle = dfa.settle;
    }
    GramView gview{smem_addr(s_tab), dfa.gram, reinterpret_cast<const uint32_t*>(dfa.table), smem_addr(s_cls), dfa.start, dfa.accept_from};
    const uint32_t reloc = kSmemTable<TIER> ? smem_addr(s_tab) : 0u;      // states are shared addresses in tiers 0/1
    const uint32_t st_start = dfa.start + reloc, st_accept = dfa.accept_from + reloc;
    uint8_t* my_slot = s_slots + (size_t)warp * slot_bytes;      // this warp's one staging slot
    const uint32_t n_batches = (b.n_reads + 31u) >> 5;
    const uint32_t stride = gridDim.x * nwarps;
    const uintptr_t base = reinterpret_cast<uintptr_t>(b.bases);

    // A batch descriptor is fetched in two steps so the offset loads (streamed from HBM) have a whole batch of work to
    // land: load_raw only issues them, finalize (one iteration later) consumes them.
    struct RawDesc { uint32_t s, e; bool valid; };
    auto load_raw = [&](uint32_t batch) -> RawDesc {
        RawDesc d;
        const uint32_t r = batch * 32u + lane;
        d.valid = batch < n_batches && r < b.n_reads;
        d.s = d.valid ? __ldg(b.start + r) : 0u;
        d.e = d.valid ? __ldg(b.end + r) : 0u;
        return d;
    };
    auto finalize = [&](const RawDesc& raw) -> BatchDesc {
        BatchDesc d;
        const bool valid = raw.valid;
        d.s = raw.s;
        d.e = raw.e;
        if (d.e < d.s) d.e = d.s;
        const uint32_t lo = __reduce_min_sync(kFullMask, valid ? d.s : 0xFFFFFFFFu);
        const uint32_t hi = __reduce_max_sync(kFullMask, valid ? d.e : 0u);
        if (hi > lo) {
            const uintptr_t a_lo = (base + lo) & ~uintptr_t(15), a_hi = (base + hi + 15) & ~uintptr_t(15);
            d.src = a_lo;
            d.off = (uint32_t)(base + d.s - a_lo);
            d.fits = (a_hi - a_lo) <= (uintptr_t)(slot_bytes - 16u);
            d.bytes = d.fits ? (uint32_t)(a_hi - a_lo) : 0u;
        } else {
            d.src = base; d.off = 0; d.bytes = 0; d.fits = true;
        }
        return d;
    };
    auto issue = [&](const BatchDesc& d) {
        if (d.fits && d.bytes && lane == 0) {
            const uint32_t bar = smem_addr(&s_bar[warp]);
            mbar_arrive_expect_tx(bar, d.bytes);
            bulk_copy_g2s(smem_addr(my_slot), reinterpret_cast<const void*>(d.src), d.bytes, bar);
        }
    };

    uint32_t batch = blockIdx.x * nwarps + warp;
    uint32_t phase = 0;
    unsigned long long local_count = 0;

    BatchDesc cur = finalize(load_raw(batch));
    issue(cur);
    // nxt: the batch after cur.  The warp has no second slot for it; its bytes are pulled into L2 (126 MB) instead, so
    // the copy issued when the slot frees only pays L2 latency and the HBM requests in flight are not capped by the
    // staging memory.  ahead: offsets in flight for the batch after nxt.
    auto prefetch = [&](const BatchDesc& d) {
        if (d.fits && d.bytes && lane == 0) bulk_prefetch_l2(reinterpret_cast<const void*>(d.src), d.bytes);
    };
    BatchDesc nxt = finalize(load_raw(batch + stride));
    prefetch(nxt);
    RawDesc ahead = load_raw(batch + 2u * stride);

    for (; batch < n_batches; batch += stride) {
        // mate / split-set masks of this batch: requested now, consumed after the walk (an L2/HBM round trip otherwise
        // sits between the walk and the result store of every batch)
        const uint32_t or_word = b.or_mask ? __ldg(b.or_mask + batch) : 0u;
        const uint32_t pre_word = b.pre_or_mask ? __ldg(b.pre_or_mask + batch) : 0u;
        uint32_t st = st_start;
        const uint32_t len = cur.e - cur.s;
        if (cur.fits) {
            if (cur.bytes) {
                mbar_wait(smem_addr(&s_bar[warp]), phase);
                phase ^= 1u;
            }
            if (TIER == 3) st = gram_walk(gview, smem_addr(my_slot) + cur.off, len) ? st_accept : 0u;
            else if (len) st = walk_shared(view, smem_addr(my_slot) + cur.off, len, st);
        } else if (len) {
            st = walk_global(view, b.bases + cur.s, len, st);
        }
        const uint32_t r = batch * 32u + lane;
        bool found = st >= st_accept;
        found |= ((pre_word >> lane) & 1u) != 0;      // split pattern sets, see api.cu
        const bool sel = (r < b.n_reads) && (found != (b.invert != 0));
        uint32_t m = __ballot_sync(kFullMask, sel);
        if (lane == 0) {
            m |= or_word;
            b.mask_out[batch] = m;
            local_count += __popc(m);
        }
        __syncwarp();   // every lane is done reading the slot before it is refilled
        cur = nxt;
        issue(cur);
        nxt = finalize(ahead);
        prefetch(nxt);
        ahead = load_raw(batch + 3u * stride);
    }
    if (lane == 0 && b.count && local_count) atomicAdd(b.count, local_count);
}

template <int TIER>
cudaError_t launch_t(const DeviceDfa& dfa, const SoaBatch& b, int sm_count, cudaStream_t stream) {
    const uint32_t avg = b.avg_read_len ? b.avg_read_len : 150u;
    uint32_t slot = ((32u * avg + 64u + 15u) & ~15u) + 16u;
    slot = (slot + 127u) & ~127u;
    const uint32_t table_smem = kSmemTable<TIER> ? ((dfa.table_bytes + 127u) & ~127u) : TIER == 3 ? ((dfa.gram.bitmap_bytes + 127u) & ~127u) : ((dfa.hot_entries * 4u + 127u) & ~127u);
    const uint32_t fixed = table_smem + 256u + 32u * 8u + 128u;
    const uint32_t budget = 227u * 1024u;
    if (fixed + slot > budget) {
        // slot too large for even one warp: shrink it; oversized batches take the global-memory path
        slot = (budget - fixed) & ~127u;
    }
    uint32_t warps = (budget - fixed) / slot;
    if (warps > 32u) warps = 32u;
    if (warps < 1u) warps = 1u;
    const uint32_t smem = fixed + warps * slot;
    auto kern = match_soa_kernel<TIER>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t n_batches = (b.n_reads + 31u) / 32u;
    uint32_t grid = (n_batches + warps - 1u) / warps;
    if (grid > (uint32_t)sm_count) grid = (uint32_t)sm_count;
    if (grid == 0) return cudaSuccess;
    kern<<<grid, warps * 32u, smem, stream>>>(b, dfa, slot, table_smem);
    return cudaGetLastError();
}

}  // namespace

// ---- read-name lookup over SoA headers (-N) ------------------------------------------------------------------------
// QueryNameMatcher::read_match (/root/reference/src/lib/matcher.rs:631-638): id = header up to the first space,
// selected iff (id in set) != invert.  One read per thread, ids read straight from global memory (a warp's 32 headers
// are contiguous, so its loads share cache lines); the table probe is one random 8-byte L2 read per read.
// Algorithmic bytes per read: header + 4 (offset) + 8 (one slot) + 1/8 (result).
namespace {
// (A software pipeline over the chain offsets -> header -> bucket -> stored id -- every iteration consumes what the previous
// one asked for and asks for the next link of four different batches, header bytes re-read from L1 -- was built and measured
// in round 2 as well: 61 Greads/s against 82 for this kernel.  The lookup is bound by instruction issue as much as by
// latency (8.4 warp-instructions per read, issue slots 62 % busy), and the pipeline's bookkeeping costs more than it hides.)
// (Four reads per thread run stage by stage -- all offset loads, all header loads, all bucket loads in flight together -- was
// built and measured in round 2: 46 Greads/s against 80 for this kernel, because its 73 registers left room for three
// blocks per SM instead of eight; the loads in flight per SM went DOWN.  Removed.)
// eight resident CTAs of 256 threads (32 registers each): measured 77 / 85 / 95 Greads/s with 4 / 6 / 8 of them
#ifndef FQG_NAMES_MIN_BLOCKS
#define FQG_NAMES_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(256, FQG_NAMES_MIN_BLOCKS) match_names_soa_kernel(SoaBatch b, NamesView nv) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_batches = (b.n_reads + 31u) >> 5;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uintptr_t base = reinterpret_cast<uintptr_t>(b.bases);
    auto load = [](uintptr_t p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); };
    unsigned long long local_count = 0;
    for (uint32_t batch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; batch < n_batches; batch += warps) {
        const uint32_t r = batch * 32u + lane;
        bool found = false;
        if (r < b.n_reads) {
            const uint32_t s = __ldg(b.start + r), e = __ldg(b.end + r);
            const uintptr_t a = base + s;
            const uint32_t n = e > s ? e - s : 0u;
            // Fast path, ids that end within the first 16 header bytes: the five aligned words that hold those bytes are
            // requested together, the id is cut at the first space, hashed and later compared out of registers -- no loop,
            // no second pass over the header.  (The generic word-by-word path below cost 8.4 warp-instructions per read,
            // a third of them iterator bookkeeping, profiles/r2h_names_soa_ncu.txt.)
            const uint32_t n16 = n < 16u ? n : 16u, sh = (uint32_t)(a & 3u) * 8u, span = (uint32_t)(a & 3u) + n16;
            const uintptr_t wp = a & ~uintptr_t(3);
            uint32_t raw[5];
#pragma unroll
            for (uint32_t k = 0; k < 5u; k++) raw[k] = 4u * k < span ? load(wp + 4u * k) : 0u;
            uint32_t w[4], id_len = n16;
            bool ended = n <= 16u;      // the header itself ends within the 16 bytes
#pragma unroll
            for (uint32_t k = 0; k < 4u; k++) {
                w[k] = __funnelshift_r(raw[k], raw[k + 1], sh);
                const uint32_t have = n16 > 4u * k ? n16 - 4u * k : 0u;      // header bytes in this word
                const uint32_t keep = have >= 4u ? 0xFFFFFFFFu : (1u << (8u * have)) - 1u;
                w[k] &= keep;
                const uint32_t x = w[k] ^ 0x20202020u;
                const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu) & keep & 0x80808080u;      // 0x80 where a byte is ' '
                if (z && id_len == n16 && 4u * k < id_len) {
                    const uint32_t kk = 4u * k + (((uint32_t)__ffs((int)z) - 1u) >> 3);
                    if (kk < id_len) { id_len = kk; ended = true; }
                }
            }
            if (ended) {
                uint64_t h = fqg_name_hash_init();
#pragma unroll
                for (uint32_t k = 0; k < 4u; k++) {
                    const uint32_t have = id_len > 4u * k ? id_len - 4u * k : 0u;
                    w[k] &= have >= 4u ? 0xFFFFFFFFu : (1u << (8u * have)) - 1u;
                    if (have) h = fqg_name_hash_word(h, w[k]);
                }
                h = fqg_name_hash_finish(h, id_len);
                found = names_probe_words(nv, w, id_len, h, names_load_bucket(nv, h));
            } else {
                // longer ids: hash the header word by word until a word holds a space, which ends the id
                uint64_t h = fqg_name_hash_init();
                id_len = n;
                for (IdWords<uintptr_t, decltype(load)> it(a, n, load); it.more();) {
                    const uint32_t pos = it.pos, ww = it.next(), x = ww ^ 0x20202020u;
                    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);      // 0x80 where a byte is ' '
                    if (z) {                                  // (the zero padding past the end is never a space)
                        const uint32_t k = ((uint32_t)__ffs((int)z) - 1u) >> 3;
                        id_len = pos + k;
                        if (k) h = fqg_name_hash_word(h, ww & ((1u << (8u * k)) - 1u));
                        break;
                    }
                    h = fqg_name_hash_word(h, ww);
                }
                h = fqg_name_hash_finish(h, id_len);
                found = names_probe(nv, a, id_len, load, h);
            }
        }
        const bool sel = (r < b.n_reads) && (found != (b.invert != 0));
        uint32_t m = __ballot_sync(kFullMask, sel);
        if (lane == 0) {
            if (b.or_mask) m |= __ldg(b.or_mask + batch);
            b.mask_out[batch] = m;
            local_count += __popc(m);
        }
    }
    if (lane == 0 && b.count && local_count) atomicAdd(b.count, local_count);
}
}  // namespace

cudaError_t launch_match_names_soa(const NamesView& nv, const SoaBatch& b, int sm_count, cudaStream_t stream) {
    const uint32_t n_batches = (b.n_reads + 31u) / 32u;
    if (!n_batches) return cudaSuccess;
    uint32_t grid = (n_batches + 7u) / 8u;
    const uint32_t cap = (uint32_t)sm_count * 8u * 4u;        // 8 resident CTAs of 256 threads per SM, 4 waves' worth of slack
    if (grid > cap) grid = cap;
    match_names_soa_kernel<<<grid, 256, 0, stream>>>(b, nv);
    return cudaGetLastError();
}

cudaError_t launch_match_soa(const DeviceDfa& dfa, const SoaBatch& b, int sm_count, cudaStream_t stream) {
    // One staging slot per warp + the next batch pulled into L2 (cp.async.bulk.prefetch.L2) instead of a second slot:
    // measured on B200 with 150-base reads, 32 resident warps with one slot each beat 22 warps with two
    // (tier 0: 0.83 -> 0.90 of the HBM roofline), because the walk is a dependent chain per lane and more warps hide it.
    switch (dfa.tier) {
        case 0: return launch_t<0>(dfa, b, sm_count, stream);
        case 1: return launch_t<1>(dfa, b, sm_count, stream);
        case 4: return launch_t<4>(dfa, b, sm_count, stream);
        case 3: return launch_t<3>(dfa, b, sm_count, stream);     // the filter bitmap takes 64-128 KB of the shared memory
        default: return launch_t<2>(dfa, b, sm_count, stream);
    }
}

}  // namespace fqg
