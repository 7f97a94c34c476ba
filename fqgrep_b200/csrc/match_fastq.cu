// match_fastq.cu -- fused FASTQ parse + match kernel (sm_100a): raw FASTQ bytes in HBM are read ONCE.
//
// Replaces, in one pass, three stages of the reference:
//   * the reader thread's record splitting (seq_io 0.3.4 fastq::Reader / RecordSet [third party], driven from
//     /root/reference/src/lib/seq_io.rs:60-80, 149-198, 271-281): newline scan, 4-line record structure,
//     '@' / '+' / equal-length validation, CR stripping;
//   * the worker-pool loop calling Matcher::read_match per record (src/lib/seq_io.rs:314-326 ->
//     src/main.rs:635-640 -> src/lib/matcher.rs:164-175), incl. QueryNameMatcher::read_match (:631-638);
//   * the per-record `found` flag (src/main.rs:558-575, 635-657), emitted as one bit per record.
//
// Shape: the byte stream is cut into 16 KB tiles.  A CTA takes the next tile id from an atomic counter
// (in-order acquisition), pulls tile + 2 KB of overlap into shared memory with one bulk async copy (TMA),
// builds a newline bitmap with SIMD-within-register byte compares, and obtains the number of newlines
// before its tile by a decoupled look-back over per-tile aggregates (single pass, no second read of HBM).
// With the global line number known, every record that STARTS in the tile is handled by one thread:
// the four line ends come from the bitmap, the sequence line is walked through the DFA straight out of
// shared memory, and the selection bit is OR-ed into the global record bitmask.  Records that do not end
// inside tile+overlap (reads longer than ~1 kb) fall back to a global-memory walk.
#include "dfa_device.cuh"
#include "kernels.hpp"
#include "names_device.cuh"

namespace fqg {
namespace {

constexpr uint32_t kTile = 19456;            // 19 KB: ~61 records of 317 B, i.e. two almost full warps in the record phase
constexpr uint32_t kOverlap = 1024;
constexpr uint32_t kLoad = kTile + kOverlap;
constexpr uint32_t kGroups = kLoad / 32;          // one bitmap word per 32 bytes
constexpr uint32_t kNone = 0xFFFFFFFFu;

enum : uint32_t { kErrInvalidStart = 1, kErrInvalidSep = 2, kErrUnequal = 3, kErrUnexpectedEnd = 4, kErrMaskOverflow = 6 };

__device__ __forceinline__ void report(const FastqArgs& a, uint64_t pos, uint32_t code) {
    atomicMin(a.err, (unsigned long long)((pos << 4) | code));
}

__device__ __forceinline__ uint4 lds_v4_volatile(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8_volatile(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// 4 bytes -> 4 bits: bit i set iff byte i == '\n'.  Exact per-byte zero test (no borrow false positives).
__device__ __forceinline__ uint32_t newline_nibble(uint32_t w) {
    const uint32_t x = w ^ 0x0A0A0A0Au;
    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 where the byte is zero
    return (z * 0x00204081u) >> 28;
}

struct NamesProbe {
    NamesView nv;
    // exact membership of id bytes [p, p+len) in the name set; p is a shared-window address / a global pointer
    __device__ bool contains_shared(uint32_t p, uint32_t len) const {
        return names_contains(nv, p, len, [](uint32_t a) { return lds_u32_volatile(a); });
    }
    __device__ bool contains_global(const uint8_t* p, uint32_t len) const {
        return names_contains(nv, reinterpret_cast<uintptr_t>(p), len, [](uintptr_t a) { return __ldg(reinterpret_cast<const uint32_t*>(a)); });
    }
};

// TIER 0,1,2,4: DFA tiers (see kernels.hpp); TIER 3: read-name lookup; TIER 5: index only (no matching).
template <int TIER>
struct RecordEval {
    DfaView<((TIER != 3 && TIER != 5) ? TIER : 0)> view;
    NamesProbe names;
    uint32_t start, accept_from;
    bool invert;

    // record fully in shared memory: s = '@' position, p0..p3 = the four line ends (shared-window addresses)
    __device__ __forceinline__ bool eval_shared(uint32_t tile, uint32_t s, uint32_t p0, uint32_t p1, uint32_t seq_len, uint32_t* idh, bool want_idh) const {
        uint32_t head_end = p0;
        if (head_end > s + 1 && lds_u8_volatile(tile + head_end - 1) == '\r') head_end--;
        uint32_t id_len = 0;
        if (TIER == 3 || want_idh) {      // id = header up to the first space; scanned and hashed four bytes at a time
            // (a malformed record may "start" on an empty line: p0 == s, no header bytes at all)
            const uint32_t a0 = tile + s + 1u, n = head_end > s + 1u ? head_end - (s + 1u) : 0u, sh = (a0 & 3u) * 8u;
            uint32_t wp = a0 & ~3u, w0 = lds_u32_volatile(wp), h = 0x811C9DC5u;
            id_len = n;
            for (uint32_t i = 0; i < n; i += 4u) {
                wp += 4u;
                const uint32_t w1 = lds_u32_volatile(wp);
                uint32_t w = __funnelshift_r(w0, w1, sh);
                w0 = w1;
                const uint32_t v = n - i < 4u ? n - i : 4u;                          // id bytes in this word
                const uint32_t x = w ^ 0x20202020u;
                uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);   // 0x80 where the byte is ' '
                if (v < 4u) z &= (1u << (8u * v)) - 1u;
                if (z) {                                                            // the id ends at the first space
                    const uint32_t bidx = ((uint32_t)__ffs(z) - 1u) >> 3;
                    if (bidx) h = (h ^ (w & ((1u << (8u * bidx)) - 1u))) * 0x01000193u;
                    id_len = i + bidx;
                    break;
                }
                if (v < 4u) w &= (1u << (8u * v)) - 1u;
                h = (h ^ w) * 0x01000193u;
            }
            h ^= id_len;
            if (want_idh) *idh = h ^ (h >> 15);
        }
        bool found;
        if (TIER == 5) found = false;
        else if (TIER == 3) found = names.contains_shared(tile + s + 1, id_len);
        else found = walk_shared(view, tile + p0 + 1, seq_len, start) >= accept_from;
        return found != invert;
    }
    __device__ __noinline__ bool eval_global(const uint8_t* head, uint32_t head_len, const uint8_t* seq, uint32_t seq_len, uint32_t* idh, bool want_idh) const {
        uint32_t id_len = 0;
        while (id_len < head_len && head[id_len] != ' ') id_len++;
        if (want_idh) {                 // same word-wise hash as eval_shared, assembled from bytes
            uint32_t h = 0x811C9DC5u;
            for (uint32_t i = 0; i < id_len; i += 4u) {
                uint32_t w = 0;
                for (uint32_t b = 0; b < 4u && i + b < id_len; b++) w |= (uint32_t)head[i + b] << (8u * b);
                h = (h ^ w) * 0x01000193u;
            }
            h ^= id_len;
            *idh = h ^ (h >> 15);
        }
        bool found;
        if (TIER == 5) found = false;
        else if (TIER == 3) found = names.contains_global(head, id_len);
        else found = walk_global(view, seq, seq_len, start) >= accept_from;
        return found != invert;
    }
};

// ---- the phases of one tile, shared by the two kernel shapes below ---------------------------------------------------

struct TileGeom {
    uint32_t t;            // global tile id
    uint64_t T0;           // stream offset of the tile
    uint32_t avail;        // stream positions visible in shared memory (tile + overlap)
    uint32_t region;       // stream positions this tile owns
    uint32_t real_avail;   // real bytes among `avail` (the rest is the virtual final newline)
};
__device__ __forceinline__ TileGeom tile_geom(const FastqArgs& a, uint32_t t, uint64_t stream_len) {
    TileGeom g;
    g.t = t;
    g.T0 = (uint64_t)t * kTile;
    const uint64_t remaining = stream_len - g.T0;
    g.avail = remaining < kLoad ? (uint32_t)remaining : kLoad;
    g.region = remaining < kTile ? (uint32_t)remaining : kTile;
    const uint64_t real_left = a.nbytes > g.T0 ? a.nbytes - g.T0 : 0;
    g.real_avail = real_left < kLoad ? (uint32_t)real_left : kLoad;
    return g;
}

// newline bitmap: one word per 32 bytes, built by `nthreads` cooperating threads
__device__ __forceinline__ void build_bitmap(uint32_t tile_sa, uint32_t* s_bitmap, const TileGeom& g, uint32_t rtid, uint32_t nthreads) {
    for (uint32_t grp32 = rtid; grp32 < kGroups + 2; grp32 += nthreads) {
        uint32_t bits = 0;
        if (grp32 * 32u < g.real_avail) {
            const uint4 lo = lds_v4_volatile(tile_sa + grp32 * 32u), hi = lds_v4_volatile(tile_sa + grp32 * 32u + 16u);
            bits = newline_nibble(lo.x) | (newline_nibble(lo.y) << 4) | (newline_nibble(lo.z) << 8) | (newline_nibble(lo.w) << 12) |
                   (newline_nibble(hi.x) << 16) | (newline_nibble(hi.y) << 20) | (newline_nibble(hi.z) << 24) | (newline_nibble(hi.w) << 28);
            if (grp32 * 32u + 32u > g.real_avail) bits &= (1u << (g.real_avail - grp32 * 32u)) - 1u;
        }
        if (g.real_avail < g.avail && (g.real_avail >> 5) == grp32) bits |= 1u << (g.real_avail & 31u);   // the virtual final newline
        s_bitmap[grp32] = bits;
    }
}

// exclusive prefix of newline counts per 32-byte group (thread i owns groups [i*GPT, (i+1)*GPT)); `sync` is the team barrier
template <uint32_t NTHREADS, typename Sync>
__device__ __forceinline__ void prefix_counts(const uint32_t* s_bitmap, uint16_t* s_pre, uint32_t* s_wsum, uint32_t rtid, Sync sync) {
    constexpr uint32_t GPT = (kGroups + NTHREADS - 1) / NTHREADS, NW = NTHREADS / 32;
    const uint32_t lane = rtid & 31u, warp = rtid >> 5;
    uint32_t cnt[GPT], mine = 0;
#pragma unroll
    for (uint32_t k = 0; k < GPT; k++) {
        const uint32_t grp32 = rtid * GPT + k;
        cnt[k] = grp32 < kGroups ? __popc(s_bitmap[grp32]) : 0u;
        mine += cnt[k];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= (uint32_t)d) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    sync();
    uint32_t warp_base = 0;
#pragma unroll
    for (uint32_t w = 0; w < NW; w++) if (w < warp) warp_base += s_wsum[w];
    uint32_t run = warp_base + incl - mine;
#pragma unroll
    for (uint32_t k = 0; k < GPT; k++) {
        const uint32_t grp32 = rtid * GPT + k;
        if (grp32 <= kGroups) s_pre[grp32] = (uint16_t)run;
        run += cnt[k];
    }
    sync();
}
__device__ __forceinline__ uint32_t region_newlines(const uint32_t* s_bitmap, const uint16_t* s_pre, uint32_t region) {
    return (region & 31u) == 0 ? s_pre[region >> 5] : s_pre[region >> 5] + __popc(s_bitmap[region >> 5]);
}

// Decoupled look-back by one warp: number of newlines before tile t; publishes this tile's aggregate / inclusive prefix.
// Two levels.  Every resident tile group works on its own tile at about the same time (1184 tiles in flight), so a flat
// look-back finds nothing but aggregates for a whole wave and needs wave/32 dependent L2 round trips per tile -- measured,
// that chain WAS the kernel's run time.  Tiles are therefore grouped in segments of 32: the tile that closes a segment
// publishes the segment's aggregate as soon as it has seen its 31 predecessors' aggregates, and a look-back sums its own
// segment first (one round), then hops 32 segments (1024 tiles) per round.  State words: flag << 62 | value, flag 1 =
// aggregate, 2 = inclusive prefix.  Tiles are handed out in order, so every state a tile waits for belongs to a group that
// is already running and publishes it without waiting for anything later: no circular wait.
__device__ __forceinline__ unsigned long long lookback(const FastqArgs& a, uint32_t t, uint32_t agg, uint32_t lane) {
    const unsigned long long kValue = (1ull << 62) - 1;
    auto poll = [](const unsigned long long* p) {
        unsigned long long st;
        do { st = *reinterpret_cast<const volatile unsigned long long*>(p); } while ((st >> 62) == 0);
        return st;
    };
    // sums the 32 states held by the lanes (lane 0 = nearest) up to and including the first inclusive prefix
    auto fold = [&](unsigned long long st, bool use, bool& hit_prefix) {
        const uint32_t has_prefix = __ballot_sync(kFullMask, use && (st >> 62) == 2);
        const uint32_t first = has_prefix ? (uint32_t)__ffs(has_prefix) - 1u : 32u;
        unsigned long long v = (use && lane <= first) ? (st & kValue) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
        hit_prefix = has_prefix != 0;
        return v;
    };
    unsigned long long excl = 0;
    const uint32_t r = t & 31u;                    // predecessors inside this tile's own segment
    if (t > 0) {
        if (lane == 0) atomicExch(a.tile_state + t, (1ull << 62) | agg);
        bool done = false;
        const bool mine = lane < r;
        excl = fold(mine ? poll(a.tile_state + (t - 1u - lane)) : 0ull, mine, done);
        if (r == 31u && !done && lane == 0) atomicExch(a.seg_state + (t >> 5), (1ull << 62) | (excl + agg));
        int32_t seg = (int32_t)(t >> 5) - 1;       // nearest earlier segment
        while (!done) {
            const int32_t idx = seg - (int32_t)lane;
            excl += fold(idx >= 0 ? poll(a.seg_state + idx) : (2ull << 62), true, done);      // before the stream: prefix 0
            seg -= 32;
        }
    }
    if (lane == 0) {
        atomicExch(a.tile_state + t, (2ull << 62) | (excl + agg));
        if (r == 31u) atomicExch(a.seg_state + (t >> 5), (2ull << 62) | (excl + agg));
    }
    return excl;
}

// one thread per record that starts in the tile; `rwarp`/`nthreads` describe the team of warps that shares the records
template <int TIER>
__device__ __forceinline__ unsigned long long match_records(const FastqArgs& a, const RecordEval<TIER>& ev, uint32_t tile_sa, const uint32_t* s_bitmap,
                                                            const uint16_t* s_pre, const TileGeom& g, unsigned long long Lt, uint32_t agg,
                                                            uint64_t stream_len, uint32_t rwarp, uint32_t lane, uint32_t nthreads) {
    const bool want_idh = a.id_hash != nullptr;
    unsigned long long local_count = 0;
    // record k begins right after the newline of local rank j0 + 4 (k - extra)
    const bool first_global = g.t == 0;
    const uint32_t extra = first_global ? 1u : 0u;
    const uint32_t j0 = 3u - (uint32_t)(Lt & 3ull);          // first local newline rank that ends a record
    const unsigned long long R_base = first_global ? 0ull : (Lt + j0 + 1ull) >> 2;
    uint32_t n_list = extra + (agg > j0 ? (agg - j0 + 3u) / 4u : 0u);
    if (g.T0 + g.region == stream_len && agg > j0 && ((agg - j0 - 1u) & 3u) == 0) n_list--;   // the stream's final newline starts nothing
    // position of the newline with local rank r (< agg): binary search in the per-group prefix counts, then bit select
    const uint32_t n_groups_region = (g.region + 31u) >> 5;
    auto newline_of_rank = [&](uint32_t r) -> uint32_t {
        uint32_t lo = 0, hi = n_groups_region;             // s_pre[lo] <= r < s_pre[hi]
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_pre[mid] <= r) lo = mid; else hi = mid;
        }
        uint32_t bits = s_bitmap[lo];
        for (uint32_t n = r - s_pre[lo]; n; n--) bits &= bits - 1u;
        return lo * 32u + (uint32_t)__ffs(bits) - 1u;
    };
    const uint32_t n_groups_avail = (g.avail + 31u) >> 5;
    auto next_nl = [&](uint32_t p) -> uint32_t {
        uint32_t grp32 = p >> 5;
        if (grp32 >= n_groups_avail) return kNone;
        uint32_t w = s_bitmap[grp32] & (0xFFFFFFFFu << (p & 31u));
        while (!w) {
            if (++grp32 >= n_groups_avail) return kNone;
            w = s_bitmap[grp32];
        }
        return grp32 * 32u + (uint32_t)__ffs(w) - 1u;
    };
    const uint64_t T0 = g.T0;
    for (uint32_t k0 = rwarp * 32u; k0 < n_list; k0 += nthreads) {
        const uint32_t k = k0 + lane;
        const bool valid = k < n_list;
        bool sel = false;
        uint32_t idh = 0, seq_n = 0;
        unsigned long long seq_pos = 0;
        if (valid) {
            const uint32_t s = k < extra ? 0u : newline_of_rank(j0 + 4u * (k - extra)) + 1u;
            const uint32_t p0 = next_nl(s);
            const uint32_t p1 = p0 == kNone ? kNone : next_nl(p0 + 1u);
            const uint32_t p2 = p1 == kNone ? kNone : next_nl(p1 + 1u);
            const uint32_t p3 = p2 == kNone ? kNone : next_nl(p2 + 1u);
            if (p3 != kNone) {
                const uint32_t cr1 = (p1 > p0 + 1u && lds_u8_volatile(tile_sa + p1 - 1u) == '\r') ? 1u : 0u;
                const uint32_t cr3 = (p3 > p2 + 1u && lds_u8_volatile(tile_sa + p3 - 1u) == '\r') ? 1u : 0u;
                const uint32_t seq_len = p1 - p0 - 1u - cr1, qual_len = p3 - p2 - 1u - cr3;
                if (lds_u8_volatile(tile_sa + s) != '@') report(a, T0 + s, kErrInvalidStart);
                else if (lds_u8_volatile(tile_sa + p1 + 1u) != '+') report(a, T0 + p1 + 1u, kErrInvalidSep);
                else if (seq_len != qual_len) report(a, T0 + s, kErrUnequal);
                seq_pos = T0 + p0 + 1u;
                seq_n = seq_len;
                sel = ev.eval_shared(tile_sa, s, p0, p1, seq_len, &idh, want_idh);
            } else {
                // record runs past tile+overlap (or the stream is truncated): walk it in global memory
                const uint8_t* base = a.buf;
                const uint64_t gs = T0 + s;
                uint64_t q[4];
                uint64_t p = gs;
                int found = 0;
                for (; found < 4 && p < stream_len; p++)
                    if (p == a.nbytes || base[p] == '\n') q[found++] = p;
                if (found < 4) report(a, gs, kErrUnexpectedEnd);
                else {
                    const uint32_t cr0 = (q[0] > gs + 1 && base[q[0] - 1] == '\r') ? 1u : 0u;
                    const uint32_t cr1 = (q[1] > q[0] + 1 && base[q[1] - 1] == '\r') ? 1u : 0u;
                    const uint32_t cr3 = (q[3] > q[2] + 1 && base[q[3] - 1] == '\r') ? 1u : 0u;
                    const uint64_t seq_len = q[1] - q[0] - 1 - cr1, qual_len = q[3] - q[2] - 1 - cr3;
                    if (base[gs] != '@') report(a, gs, kErrInvalidStart);
                    else if (base[q[1] + 1] != '+') report(a, q[1] + 1, kErrInvalidSep);
                    else if (seq_len != qual_len) report(a, gs, kErrUnequal);
                    seq_pos = q[0] + 1;
                    seq_n = (uint32_t)seq_len;
                    sel = ev.eval_global(base + gs + 1, q[0] > gs ? (uint32_t)(q[0] - gs - 1 - cr0) : 0u, base + q[0] + 1, (uint32_t)seq_len, &idh, want_idh);
                }
            }
        }
        const unsigned long long R_w = R_base + k0;      // record number of lane 0
        if (valid) {
            if (R_base + k >= a.mask_bits || (want_idh && R_base + k >= a.id_hash_cap)) { report(a, T0, kErrMaskOverflow); sel = false; }
            else if (want_idh) a.id_hash[R_base + k] = idh;
            if (a.seq_start && R_base + k < a.span_cap) { a.seq_start[R_base + k] = seq_pos; a.seq_len[R_base + k] = seq_n; }
        }
        const uint32_t m = __ballot_sync(kFullMask, sel);
        if (lane == 0 && m) {
            const uint32_t sh = (uint32_t)(R_w & 31ull);
            atomicOr(a.mask + (R_w >> 5), m << sh);
            if (sh && (m >> (32u - sh))) atomicOr(a.mask + (R_w >> 5) + 1, m >> (32u - sh));
            local_count += __popc(m);
        }
    }
    return local_count;
}

template <int TIER>
__device__ __forceinline__ RecordEval<TIER> make_eval(const FastqArgs& a, const DeviceDfa& dfa, uint8_t* s_tab, uint8_t* s_cls) {
    RecordEval<TIER> ev;
    ev.view.tab = smem_addr(s_tab);
    ev.view.t32 = reinterpret_cast<const uint32_t*>(dfa.table);
    ev.view.cls = smem_addr(s_cls);
    ev.view.hot_limit = TIER == 2 ? dfa.hot_entries : 0u;
    {
        const uint32_t rel = kSmemTable<TIER> ? smem_addr(s_tab) : 0u;
        ev.view.settled_a = dfa.settled[0] == 0xFFFFFFFFu ? 0xFFFFFFFFu : dfa.settled[0] + rel;
        ev.view.settled_b = dfa.settled[1] == 0xFFFFFFFFu ? 0xFFFFFFFFu : dfa.settled[1] + rel;
        ev.view.settle = dfa.settle;
    }
    ev.names.nv = a.names;
    const uint32_t reloc = kSmemTable<TIER> ? smem_addr(s_tab) : 0u;      // states are shared addresses in the smem-table tiers
    ev.start = dfa.start + reloc;
    ev.accept_from = dfa.accept_from + reloc;
    ev.invert = a.invert != 0;
    return ev;
}
template <int TIER>
__device__ __forceinline__ void stage_automaton(const DeviceDfa& dfa, uint8_t* s_tab, uint8_t* s_cls) {
    if (kSmemTable<TIER>) stage_table(dfa.table, dfa.table_bytes, s_tab, threadIdx.x, blockDim.x);
    if (TIER == 2)
        for (uint32_t i = threadIdx.x; i < dfa.hot_entries; i += blockDim.x) reinterpret_cast<uint32_t*>(s_tab)[i] = __ldg(reinterpret_cast<const uint32_t*>(dfa.table) + i);
    if (TIER != 3 && TIER != 5)
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_cls[i] = __ldg(dfa.byte_class + i);
}

// ---- the kernel: symmetric tile groups ----------------------------------------------------------------------------------
// (A warp-specialised shape -- 2 scanner + 2 matcher warps per group over two tile buffers -- was built and measured this
// round: same results, 0.95x on the DFA path and 0.83x on -N, because two buffers per group halve the groups an SM can hold
// and the 16 remaining warps hide less latency than the overlap wins.  It was removed; see DESIGN.md.)
// per-group shared memory: [mbarrier 8][prefix 8][misc 240][tile kLoad+32][bitmap][pre], 128-byte aligned
constexpr uint32_t kGroupBytes = ((256u + (kLoad + 32u) + (kGroups + 2u) * 4u + (kGroups + 2u) * 2u) + 127u) & ~127u;

// A CTA is up to 8 independent "tile groups" of THREADS threads (named barriers 1..8) that share one copy of the
// automaton in shared memory; each group runs the tile loop on its own buffers, so one SM keeps up to 8 tiles in
// different phases (copy / scan / look-back / match) in flight.
template <int TIER, int THREADS>
__global__ void __launch_bounds__(1024, 1) match_fastq_kernel(FastqArgs a, DeviceDfa dfa, uint32_t table_smem_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t grp = threadIdx.x / THREADS, tid = threadIdx.x % THREADS, lane = tid & 31u, warp = tid >> 5;
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1u), "n"(THREADS) : "memory"); };

    // ---- smem carve-up
    uint8_t* s_tab = smem;
    uint8_t* s_cls = smem + table_smem_bytes;                                // 256 B, shared by all groups
    uint8_t* s_grp = s_cls + 256 + (size_t)grp * kGroupBytes;                // this group's private region (128-byte aligned)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_grp);
    unsigned long long* s_prefix = reinterpret_cast<unsigned long long*>(s_bar + 1);
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_bar + 2);               // warp sums, [40] = tile id
    uint8_t* s_tile = s_grp + 256;                                           // kLoad + 32, 128-byte aligned for the bulk copy
    uint32_t* s_bitmap = reinterpret_cast<uint32_t*>(s_tile + kLoad + 32);   // kGroups + 2
    uint16_t* s_pre = reinterpret_cast<uint16_t*>(s_bitmap + kGroups + 2);   // kGroups + 2

    stage_automaton<TIER>(dfa, s_tab, s_cls);
    if (tid == 0) mbar_init(smem_addr(s_bar), 1);
    mbar_fence_init();
    __syncthreads();       // the only CTA-wide barrier: table staged, mbarriers initialised

    const RecordEval<TIER> ev = make_eval<TIER>(a, dfa, s_tab, s_cls);
    const uint32_t tile_sa = smem_addr(s_tile), bar = smem_addr(s_bar);
    // the stream always ends with a newline: a real one, or a virtual one at offset nbytes (missing final '\n')
    const uint64_t stream_len = a.nbytes + (a.virtual_final_nl ? 1u : 0u);
    uint32_t phase = 0;
    unsigned long long local_count = 0;

    for (;;) {
        if (tid == 0) s_misc[40] = atomicAdd(a.tile_counter, 1u);
        group_sync();
        if (s_misc[40] >= a.tile_count) break;
        const TileGeom g = tile_geom(a, a.tile_first + s_misc[40], stream_len);
        if (g.real_avail) {
            if (tid == 0) {
                const uint32_t bytes = (g.real_avail + 15u) & ~15u;
                mbar_arrive_expect_tx(bar, bytes);
                bulk_copy_g2s(tile_sa, a.buf + g.T0, bytes, bar);
                // Tiles are handed out in order, so the tile one full wave ahead (one per resident group) is the one some
                // group will ask for about a tile-time from now: pull it into L2 (126 MB; a wave is ~24 MB) so that copy
                // only pays L2 latency.  Each group has a single tile buffer, hence no other way to keep HBM busy.
                const uint32_t ahead = s_misc[40] + gridDim.x * (blockDim.x / THREADS);
                if (ahead < a.tile_count) {
                    const TileGeom ga = tile_geom(a, a.tile_first + ahead, stream_len);
                    if (ga.real_avail) bulk_prefetch_l2(a.buf + ga.T0, (ga.real_avail + 15u) & ~15u);
                }
                // one thread polls the transaction barrier; the rest of the group sleeps on the (hardware) named barrier
                // instead of spinning on try_wait and stealing issue slots from the other groups of the SM
                mbar_wait(bar, phase);
            }
            phase ^= 1u;
            group_sync();
        }
        build_bitmap(tile_sa, s_bitmap, g, tid, THREADS);
        group_sync();
        prefix_counts<THREADS>(s_bitmap, s_pre, s_misc, tid, group_sync);
        const uint32_t agg = region_newlines(s_bitmap, s_pre, g.region);
        if (warp == 0) {
            const unsigned long long excl = lookback(a, g.t, agg, lane);
            if (lane == 0) {
                *s_prefix = excl;
                if (g.T0 + g.region == stream_len) *a.lines_out = excl + agg;
            }
        }
        group_sync();
        local_count += match_records<TIER>(a, ev, tile_sa, s_bitmap, s_pre, g, *s_prefix, agg, stream_len, warp, lane, THREADS);
        group_sync();   // tile and bitmap are free for the next iteration
    }
    if (lane == 0 && local_count && a.count) atomicAdd(a.count, local_count);
}

// pair rule (src/main.rs:749-755) for mates in two streams: unit mask = m1 | m2; ids must agree (:741-747).
// Record counts are still on the device (written by the streams' last tiles), so they are read here.
__global__ void combine_pairs_kernel(const uint32_t* m1, const uint32_t* m2, uint32_t* out, uint64_t cap_words, const uint32_t* h1, const uint32_t* h2,
                                     uint64_t idh_cap, const unsigned long long* lines1, const unsigned long long* lines2, unsigned long long* count,
                                     unsigned long long* id_mismatch) {
    const uint64_t r1 = *lines1 >> 2, r2 = *lines2 >> 2;
    uint64_t n_records = r1 < r2 ? r1 : r2;
    uint64_t n_words = (n_records + 31) >> 5;
    if (n_words > cap_words) n_words = cap_words;
    unsigned long long c = 0, bad = ~0ull;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t m = m1[i] | m2[i];
        out[i] = m;
        c += __popc(m);
    }
    if (h1 && h2) {
        if (n_records > idh_cap) n_records = idh_cap;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_records; i += (uint64_t)gridDim.x * blockDim.x)
            if (h1[i] != h2[i] && i < bad) bad = i;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(kFullMask, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
    if (bad != ~0ull) atomicMin(id_mismatch, bad);
}

// interleaved mates (records 2k, 2k+1 of one stream): unit k = bit 2k | bit 2k+1
__global__ void combine_interleaved_kernel(const uint32_t* m, uint32_t* out, uint64_t cap_out_words, const uint32_t* h, uint64_t idh_cap,
                                           const unsigned long long* lines, unsigned long long* count, unsigned long long* id_mismatch) {
    uint64_t n_pairs = *lines >> 3;
    uint64_t n_out_words = (n_pairs + 31) >> 5;
    if (n_out_words > cap_out_words) n_out_words = cap_out_words;
    unsigned long long c = 0, bad = ~0ull;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_out_words; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (uint64_t)m[2 * i] | ((uint64_t)m[2 * i + 1] << 32);
        x = (x | (x >> 1)) & 0x5555555555555555ull;
        x = (x | (x >> 1)) & 0x3333333333333333ull;
        x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
        x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
        x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
        x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
        out[i] = (uint32_t)x;
        c += __popcll(x);
    }
    if (h) {
        if (2 * n_pairs > idh_cap) n_pairs = idh_cap / 2;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x)
            if (h[2 * i] != h[2 * i + 1] && i < bad) bad = i;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(kFullMask, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
    if (bad != ~0ull) atomicMin(id_mismatch, bad);
}

template <int TIER>
cudaError_t launch_fastq_t(const DeviceDfa& dfa, const FastqArgs& a, int sm_count, cudaStream_t stream) {
    const uint32_t table_smem = kSmemTable<TIER> ? ((dfa.table_bytes + 127u) & ~127u) : TIER == 2 ? ((dfa.hot_entries * 4u + 127u) & ~127u) : 0u;
    const uint32_t budget = 227u * 1024u;
    constexpr int THREADS = 128;
    uint32_t groups = (budget - table_smem - 256u - 128u) / kGroupBytes;
    if (groups > 8u) groups = 8u;
    if (groups < 1u) groups = 1u;
    if (a.tile_count < (uint32_t)sm_count * groups) {          // small inputs: spread the tiles over the SMs first
        groups = (a.tile_count + (uint32_t)sm_count - 1u) / (uint32_t)sm_count;
        if (groups < 1u) groups = 1u;
    }
    const uint32_t smem = table_smem + 256u + groups * kGroupBytes + 128u;
    auto kern = match_fastq_kernel<TIER, THREADS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    uint32_t grid = (a.tile_count + groups - 1u) / groups;
    if (grid > (uint32_t)sm_count) grid = (uint32_t)sm_count;
    if (grid == 0) return cudaSuccess;
    kern<<<grid, groups * THREADS, smem, stream>>>(a, dfa, table_smem);
    return cudaGetLastError();
}

}  // namespace

uint32_t fastq_tile_bytes() { return kTile; }
uint32_t fastq_overlap_bytes() { return kOverlap; }

cudaError_t launch_match_fastq(const DeviceDfa& dfa, bool names, bool index_only, const FastqArgs& a, int sm_count, cudaStream_t stream) {
    if (index_only) return launch_fastq_t<5>(dfa, a, sm_count, stream);
    if (names) return launch_fastq_t<3>(dfa, a, sm_count, stream);
    switch (dfa.tier) {
        case 0: return launch_fastq_t<0>(dfa, a, sm_count, stream);
        case 1: return launch_fastq_t<1>(dfa, a, sm_count, stream);
        case 4: return launch_fastq_t<4>(dfa, a, sm_count, stream);
        // tier 3 (q-gram filter) is an SoA-kernel path; the fused FASTQ kernel walks its exact tier-2 table
        default: return launch_fastq_t<2>(dfa, a, sm_count, stream);
    }
}

// ---- K0 part 2: compaction of the indexed sequence lines into SoA (bases back to back + u32 offsets) -------------------
namespace {
constexpr uint32_t kScanBlock = 256, kScanItems = 8, kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_w, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= (uint32_t)d) incl += u;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (uint32_t w = 0; w < kScanBlock / 32; w++) { if (w < warp) base += s_w[w]; tot += s_w[w]; }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}
__global__ void scan_block_sums_kernel(const uint32_t* len, uint64_t n, uint32_t* block_sums) {
    __shared__ uint32_t s_w[kScanBlock / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v = 0;
    for (uint32_t i = 0; i < kScanItems; i++) if (base + i < n) v += len[base + i];
    uint32_t tot;
    block_exclusive_scan(v, s_w, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void scan_sums_kernel(uint32_t* block_sums, uint32_t n_blocks) {     // one block; block_sums -> exclusive prefix, total appended
    __shared__ uint32_t s_w[kScanBlock / 32];
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += kScanBlock) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan(v, s_w, &tot);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) block_sums[n_blocks] = carry;
}
__global__ void scan_offsets_kernel(const uint32_t* len, uint64_t n, const uint32_t* block_sums, uint32_t* offsets) {
    __shared__ uint32_t s_w[kScanBlock / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t item[kScanItems], v = 0;
    for (uint32_t i = 0; i < kScanItems; i++) { item[i] = base + i < n ? len[base + i] : 0u; v += item[i]; }
    uint32_t tot;
    uint32_t run = block_sums[blockIdx.x] + block_exclusive_scan(v, s_w, &tot);
    for (uint32_t i = 0; i < kScanItems; i++) {
        if (base + i < n) offsets[base + i] = run;
        run += item[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanBlock - 1) offsets[n] = block_sums[gridDim.x];
}
// one warp per record: contiguous byte copy of its sequence line
__global__ void gather_bases_kernel(const uint8_t* fastq, const unsigned long long* seq_start, const uint32_t* seq_len, uint64_t n,
                                    const uint32_t* offsets, uint8_t* out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n; r += warps) {
        const uint8_t* src = fastq + seq_start[r];
        uint8_t* dst = out + offsets[r];
        const uint32_t L = seq_len[r];
        for (uint32_t i = lane; i < L; i += 32u) dst[i] = src[i];
    }
}
}  // namespace

uint64_t compact_scratch_words(uint64_t n_records) { return (n_records + kScanTile - 1) / kScanTile + 2; }

cudaError_t launch_compact_bases(const uint8_t* fastq, const unsigned long long* seq_start, const uint32_t* seq_len, uint64_t n_records,
                                 uint8_t* bases_out, uint32_t* offsets_out, uint32_t* block_sums, cudaStream_t stream) {
    if (n_records == 0) return cudaMemsetAsync(offsets_out, 0, 4, stream);
    const uint32_t n_blocks = (uint32_t)((n_records + kScanTile - 1) / kScanTile);
    scan_block_sums_kernel<<<n_blocks, kScanBlock, 0, stream>>>(seq_len, n_records, block_sums);
    scan_sums_kernel<<<1, kScanBlock, 0, stream>>>(block_sums, n_blocks);
    scan_offsets_kernel<<<n_blocks, kScanBlock, 0, stream>>>(seq_len, n_records, block_sums, offsets_out);
    gather_bases_kernel<<<148 * 8, 256, 0, stream>>>(fastq, seq_start, seq_len, n_records, offsets_out, bases_out);
    return cudaGetLastError();
}

cudaError_t launch_combine_pairs(const uint32_t* m1, const uint32_t* m2, uint32_t* out, uint64_t cap_words, const uint32_t* h1, const uint32_t* h2,
                                 uint64_t idh_cap, const unsigned long long* lines1, const unsigned long long* lines2, unsigned long long* count,
                                 unsigned long long* id_mismatch, cudaStream_t stream) {
    combine_pairs_kernel<<<148 * 4, 256, 0, stream>>>(m1, m2, out, cap_words, h1, h2, idh_cap, lines1, lines2, count, id_mismatch);
    return cudaGetLastError();
}
cudaError_t launch_combine_interleaved(const uint32_t* m, uint32_t* out, uint64_t cap_out_words, const uint32_t* h, uint64_t idh_cap,
                                       const unsigned long long* lines, unsigned long long* count, unsigned long long* id_mismatch, cudaStream_t stream) {
    combine_interleaved_kernel<<<148 * 4, 256, 0, stream>>>(m, out, cap_out_words, h, idh_cap, lines, count, id_mismatch);
    return cudaGetLastError();
}

}  // namespace fqg
