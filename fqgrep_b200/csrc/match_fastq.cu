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
// Shape: WARP-AUTONOMOUS TILES.  The byte stream is cut into tiles of 29 x 336 = 9744 bytes.  Every warp of the
// persistent grid runs the whole tile loop on its own, with no CTA-wide barrier after start-up:
//   1. the tile (claimed in order from an atomic counter one step earlier, see 6) + 1008 bytes of overlap arrive in the
//      warp's own shared-memory buffer by ONE bulk async copy (TMA, completion on the warp's mbarrier); the tile one wave
//      ahead is pulled into L2 at the same time (a warp has a single buffer);
//   2. newline masks: lane l owns the 336 bytes [336 l, 336 l + 336) of the buffer -- 21 LDS.128 at a lane stride of 21
//      16-byte units, which is bank-conflict free because 21 is odd -- and turns them into 11 bitmap words held in
//      REGISTERS (exact per-byte test in 3 integer ops per word, bits gathered by IDP.4A);
//   3. popcounts -> warp scan -> the tile's newline count is published (per-tile word + an atomic add into its segment of 32
//      tiles; ONE warp of the grid, the scan warp, turns complete segments into running prefixes, see Lookback);
//   4. every newline's position is scattered into a small shared list at its rank (so line k of the tile is list[k]);
//   5. the line phase (line number mod 4 of the tile's first byte) is guessed from the tile's own first records;
//   6. the next tile is claimed (the atomic's round trip hides behind the walk); lane k takes record k of the tile: five list
//      entries give the record start and its four line ends, the record is validated by seq_io's rules and its sequence line
//      is walked through the DFA straight out of shared memory, results held in registers; the moment the last record is
//      walked the buffer is handed to the next tile's copy;
//   7. ONE TILE LATER -- after steps 1-5 of the next tile, i.e. when every earlier tile has long reported -- the number of
//      newlines before the tile is read (one 64-bit word + the counts of at most 31 tiles of its own segment, both requested
//      at the top of that iteration), the guess is confirmed and the results are committed: ballot -> atomicOr into the
//      global record bitmask, id hashes, spans.  A wrong guess (damaged input only) is re-done from global memory.
// Records that do not end inside tile + overlap (reads longer than ~400 bases) are walked in global memory.  Read-name
// lookups (-N) settle their tile at once instead (their id bytes live in the tile), the index-only kernel defers by a whole
// tile.  Round 1's kernel shared one 19 KB tile between the four warps of a "tile group" and needed four group barriers, a
// shared bitmap, a binary search per record and a bit-scan per line end: 80 warp-instructions per record, 8.9 barrier stalls
// per issue, 0.35 of the raw-FASTQ roofline (profiles/r1zz_match_fastq_ncu_full.txt); this one runs at 0.59-0.62 (match) and
// 0.75 (index only) with 62 warp-instructions per record (profiles/r3*).
#include <cstdlib>

#include "dfa_device.cuh"
#include "kernels.hpp"
#include "names_device.cuh"

namespace fqg {
namespace {

#ifndef FQG_UNITS
#define FQG_UNITS 21
#endif
// FQG_ETAB=1: the walk's exact ACGT check as a table lookup (dfa_device.cuh, step4_acgt<true>) instead of arithmetic.  Measured
// on B200 (A/B of two builds in one call): 0.553 with it against 0.557 without -- two integer instructions saved per word, one
// more conflict-prone LDS and 58 bytes of spills paid.  Off.
#ifndef FQG_ETAB
#define FQG_ETAB 0
#endif
// FQG_FASTQ_TIMING=1 (a variant build, build.build_variant): per-warp cycle counters around the two places a warp can wait --
// the tile's copy and the prefix -- reported through FQG_FASTQ_STATS.  Off in the product build: the four 64-bit accumulators
// would stay live across the whole tile loop of a kernel that has no register to spare.
#ifndef FQG_FASTQ_TIMING
#define FQG_FASTQ_TIMING 0
#endif
// FQG_FASTQ_COUNTERS=0 compiles the tile statistics of FQG_FASTQ_STATS (tiles, guessed / wrongly guessed phases, dense tiles) out
#ifndef FQG_FASTQ_COUNTERS
#define FQG_FASTQ_COUNTERS 1
#endif
#ifndef FQG_EARLY_ISSUE
#define FQG_EARLY_ISSUE 1
#endif
#ifndef FQG_LISTCAP
#define FQG_LISTCAP 1024
#endif
constexpr uint32_t kUnits = FQG_UNITS;                        // 16-byte units per lane; ODD => lane-strided LDS.128 hits 8 distinct bank groups
constexpr uint32_t kLaneBytes = kUnits * 16u;          // 336
constexpr uint32_t kTileLanes = 29;                    // lanes 0..28 hold the tile proper, lanes 29..31 the overlap
constexpr uint32_t kTile = kTileLanes * kLaneBytes;    // 9744: ~30.7 records of 317 bytes, one per lane in the record phase
constexpr uint32_t kLoad = 32u * kLaneBytes;           // 10752 bytes staged per tile
constexpr uint32_t kOverlap = kLoad - kTile;           // 1008
constexpr uint32_t kWords = (kUnits + 1u) / 2u;        // 11 bitmap words per lane (the last one holds 16 bits)
constexpr uint32_t kListCap = FQG_LISTCAP;                    // newline positions per window; denser tiles take several windows
constexpr uint32_t kWindowStep = kListCap - 8u;        // ranks a window is responsible for (a record needs 5 consecutive ranks)
// per-warp shared memory: [mbarrier, 128][newline list, 2 * kListCap + 128][tile, kLoad + 128]
constexpr uint32_t kWarpListOff = 128u, kWarpTileOff = kWarpListOff + 2u * kListCap + 128u;
constexpr uint32_t kWarpBytes = kWarpTileOff + kLoad + 128u;
static_assert(kWarpBytes % 128u == 0 && kWarpTileOff % 128u == 0, "TMA destination alignment");
// tile warps that fit 227 KB of shared memory next to the smallest automaton; the register budget follows from it (launch bounds)
template <int TIER>
constexpr uint32_t kEtabBytes = (kPackedCodes<TIER> && FQG_ETAB) ? 1024u : 0u;
constexpr uint32_t kSmemWarps = (227u * 1024u - 256u - 1024u) / kWarpBytes;
// (17 would fit next to a tiny table, but a CTA of more than 512 threads leaves every thread 96 registers instead of 128)
constexpr uint32_t kMaxTileWarps = kSmemWarps > 16u && kSmemWarps < 20u ? 16u : kSmemWarps;
static_assert(kTile % 16u == 0, "tiles start on 16-byte boundaries (bulk copy source alignment)");

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
__device__ __forceinline__ uint32_t lds_u16_volatile(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}
// Tile claim whose result is NOT consumed on the spot.  A one-lane atomicAdd() on a warp-uniform address -- also as inline
// `atom.global.add` or `atom.global.inc` -- is turned by ptxas into its warp-aggregated form (leader election, ATOMG, then
// a SHFL of the result to all lanes right behind it), i.e. the warp waits for the round trip at the claim: 4.6 % of all
// stall samples sat on that SHFL (profiles/r3o_*).  The address below is the counter plus threadIdx.y (always 0: the blocks
// are one-dimensional), which ptxas has to treat as per-thread: no aggregation, and the value is only waited for where
// take_next() reads it, a walk later (+0.5 %).
__device__ __forceinline__ uint32_t atom_add_u32(uint32_t* p, uint32_t v) {
    uint32_t r;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p + threadIdx.y), "r"(v) : "memory");
    return r;
}
// store only where `on` is non-zero -- as a PREDICATED store, not a branch around one
__device__ __forceinline__ void sts_u16_if(uint32_t a, uint32_t v, uint32_t on) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.shared.u16 [%0], %1;\n}" ::"r"(a), "h"((unsigned short)v), "r"(on) : "memory");
}

// 0x80 in every byte of w that is '\n', 0 elsewhere.  Exact (no borrow between bytes): the low 7 bits of a byte of
// w ^ '\n' overflow into bit 7 of the sum unless they are all zero, and bit 7 of w itself must be clear.  c0a = 0x0A0A0A0A
// and c7f = 0x7F7F7F7F are passed in REGISTERS so that each of the three steps is one instruction (LOP3 takes a single
// immediate, and two are needed otherwise; they arrive as kernel parameters, see FastqArgs).
__device__ __forceinline__ uint32_t newline_flags(uint32_t w, uint32_t c0a, uint32_t c7f) {
    const uint32_t t = ((w ^ c0a) & c7f) + c7f;
    uint32_t z;
    asm("lop3.b32 %0, %1, %2, %3, 0x01;" : "=r"(z) : "r"(t), "r"(w), "r"(c7f));      // ~(t | w | c7f) as ONE instruction
    return z;
}
// 16 bytes -> 128 * (16-bit mask, bit i = byte i is '\n'): the flag bytes (0x80) are gathered by two integer dot products
// per 8 bytes, 0x80 * (1,2,4,8) + 0x80 * (16,32,64,128), so the mask comes out in stream order without any bit shuffling
__device__ __forceinline__ uint32_t newline_mask8x128(uint32_t w0, uint32_t w1, uint32_t c0a, uint32_t c7f) {
    uint32_t a = __dp4a(newline_flags(w0, c0a, c7f), 0x08040201u, 0u);
    return __dp4a(newline_flags(w1, c0a, c7f), 0x80402010u, a);
}

// A read-name lookup in flight: the hash is known and the home bucket has been asked for.  Keeping the two halves apart
// lets the fused kernel resolve its look-back and claim its next tile while the buckets of the 32 records travel.
struct NamePending {
    unsigned long long h = 0;
    ulonglong4 bucket{0, 0, 0, 0};
    uint32_t id_sa = 0, id_len = 0;      // the id in the staged tile (shared address)
    bool active = false;
};
struct NamesProbe {
    NamesView nv;
    __device__ __forceinline__ NamePending start_shared(uint32_t p, uint32_t len) const {
        NamePending np;
        np.h = names_hash_id(p, len, [](uint32_t a) { return lds_u32_volatile(a); });
        np.bucket = names_load_bucket(nv, np.h);
        np.id_sa = p;
        np.id_len = len;
        np.active = true;
        return np;
    }
    __device__ __forceinline__ bool finish_shared(const NamePending& np) const {
        return names_probe_from(nv, np.id_sa, np.id_len, [](uint32_t a) { return lds_u32_volatile(a); }, np.h, np.bucket);
    }
    // exact membership of id bytes [p, p+len) in the name set (records walked in global memory)
    __device__ bool contains_global(const uint8_t* p, uint32_t len) const {
        return names_contains(nv, reinterpret_cast<uintptr_t>(p), len, [](uintptr_t a) { return __ldg(reinterpret_cast<const uint32_t*>(a)); });
    }
};

// Read-id hash for the pair check (src/main.rs:741-747 asserts r1.id_bytes() == r2.id_bytes()).  Two independent 32-bit
// multiply-xor lanes over the id's little-endian words, the length folded in at the end: ids that differ in a single
// aligned word can never collide (each step is a bijection of the state), any other pair of different ids collides
// with probability 2^-64.
struct IdHash {
    uint32_t a = 0x811C9DC5u, b = 0x9E3779B9u;
    __device__ __forceinline__ void word(uint32_t w) {
        a = (a ^ w) * 0x01000193u;
        b = (b ^ w) * 0x85EBCA6Bu;
    }
    __device__ __forceinline__ unsigned long long finish(uint32_t len) const {
        uint32_t x = a ^ len, y = b + len * 0x27D4EB2Fu;
        x ^= x >> 15;
        y ^= y >> 13;
        return ((unsigned long long)y << 32) | x;
    }
};

// TIER 0,1,2,4: DFA tiers (see kernels.hpp); TIER 3: read-name lookup; TIER 5: index only (no matching).
template <int TIER>
struct RecordEval {
    DfaView<((TIER != 3 && TIER != 5) ? TIER : 0)> view;
    NamesProbe names;
    uint32_t start, accept_from;
    bool invert;

    // record fully in shared memory: s = '@' position, p0 = end of the header line (positions in the tile at shared address `tile`)
    __device__ __forceinline__ bool eval_shared(uint32_t tile, uint32_t s, uint32_t p0, uint32_t seq_len, unsigned long long* idh, bool want_idh,
                                                NamePending* pending) const {
        uint32_t head_end = p0;
        if (head_end > s + 1 && lds_u8_volatile(tile + head_end - 1) == '\r') head_end--;
        uint32_t id_len = 0;
        if (TIER == 3 || want_idh) {      // id = header up to the first space; scanned and hashed four bytes at a time
            // (a loop-free variant for ids that end within 16 bytes -- five words read together, cut, masked, hashed -- costs ten
            // more live registers than the kernel has: measured 0.592 -> 0.564 paired, 0.385 -> 0.305 for read names)
            // (a malformed record may "start" on an empty line: p0 == s, no header bytes at all)
            const uint32_t a0 = tile + s + 1u, n = head_end > s + 1u ? head_end - (s + 1u) : 0u, sh = (a0 & 3u) * 8u;
            uint32_t wp = a0 & ~3u, w0 = lds_u32_volatile(wp);
            IdHash h;
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
                    if (bidx) h.word(w & ((1u << (8u * bidx)) - 1u));
                    id_len = i + bidx;
                    break;
                }
                if (v < 4u) w &= (1u << (8u * v)) - 1u;
                h.word(w);
            }
            if (want_idh) *idh = h.finish(id_len);
        }
        bool found;
        if (TIER == 5) found = false;
        else if (TIER == 3) { *pending = names.start_shared(tile + s + 1, id_len); return false; }      // decided by finish_names
        else found = walk_shared<((TIER != 3 && TIER != 5) ? TIER : 0), kPackedCodes<TIER> && FQG_ETAB>(view, tile + p0 + 1, seq_len, start) >= accept_from;
        return found != invert;
    }
    __device__ __noinline__ bool eval_global(const uint8_t* head, uint32_t head_len, const uint8_t* seq, uint32_t seq_len, unsigned long long* idh, bool want_idh) const {
        uint32_t id_len = 0;
        while (id_len < head_len && head[id_len] != ' ') id_len++;
        if (want_idh) {                 // same word-wise hash as eval_shared, assembled from bytes
            IdHash h;
            for (uint32_t i = 0; i < id_len; i += 4u) {
                uint32_t w = 0;
                for (uint32_t b = 0; b < 4u && i + b < id_len; b++) w |= (uint32_t)head[i + b] << (8u * b);
                h.word(w);
            }
            *idh = h.finish(id_len);
        }
        bool found;
        if (TIER == 5) found = false;
        else if (TIER == 3) found = names.contains_global(head, id_len);
        else found = walk_global(view, seq, seq_len, start) >= accept_from;
        return found != invert;
    }
};

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
    if (t < a.full_tiles) {      // all but the last one or two tiles of a stream: no 64-bit arithmetic
        g.avail = kLoad; g.region = kTile; g.real_avail = kLoad;
        return g;
    }
    const uint64_t remaining = stream_len - g.T0;
    g.avail = remaining < kLoad ? (uint32_t)remaining : kLoad;
    g.region = remaining < kTile ? (uint32_t)remaining : kTile;
    const uint64_t real_left = a.nbytes > g.T0 ? a.nbytes - g.T0 : 0;
    g.real_avail = real_left < kLoad ? (uint32_t)real_left : kLoad;
    return g;
}
// real bytes of tile t that a copy has to bring in (0: the tile holds nothing but the virtual final newline)
__device__ __forceinline__ uint32_t tile_real_avail(const FastqArgs& a, uint32_t t) {
    if (t < a.full_tiles) return kLoad;
    const uint64_t T0 = (uint64_t)t * kTile;
    const uint64_t left = a.nbytes > T0 ? a.nbytes - T0 : 0;
    return left < kLoad ? (uint32_t)left : kLoad;
}

// ---- global line numbers: per-tile counts + a scan warp ---------------------------------------------------------------------
// Number of newlines before tile t.  Every resident warp works on its own tile at about the same time (~2400 tiles in
// flight), so a classic decoupled look-back would walk back through ~75 segments of per-tile counts before it met an
// inclusive prefix (round 1 and the first round-2 kernels did: 170 instructions per tile, profiles/r2h_*).  Instead:
//   tile_cnt[t]   0x80000000 | the tile's own newline count, stored the moment the tile has scanned its bytes
//   seg_acc[s]    tiles of segment s (32 consecutive tiles) that have reported << 48 | sum of their counts: every tile ADDS
//                 its count here at the same moment
//   seg_pre[s]    2 << 62 | number of newlines up to the end of segment s -- written by ONE warp of the grid, the scan warp
//                 (an extra warp of CTA 0 that owns no tiles): it follows the segments in order, waits until all 32 tiles of
//                 a segment have reported, and publishes the running sum, up to 32 segments per step.
// A tile's prefix is then seg_pre[its segment - 1] + the counts of the at most 31 tiles before it in its own segment: one
// 64-bit word and one warp reduction.  Both are requested BEFORE the record phase and consumed after it, i.e. most of a
// tile-time after the counts were stored, so the wait loops below hardly ever turn.  Nobody waits in a cycle: counts are
// stored without waiting for anything, the scan warp only waits for counts, tiles only wait for counts and the scan warp.
struct Lookback {
    uint32_t* tile_cnt;
    unsigned long long *seg_acc, *seg_pre;
};
constexpr unsigned long long kLbValue = (1ull << 62) - 1, kLbSum = (1ull << 48) - 1;
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }

__device__ __forceinline__ void lookback_publish(const Lookback& lb, uint32_t t, uint32_t agg, uint32_t lane) {
    if (lane == 0) atomicExch(lb.tile_cnt + t, 0x80000000u | agg);
    if (lane == 1) atomicAdd(lb.seg_acc + (t >> 5), (1ull << 48) | agg);
}
// (Letting a tile fall back to the nearest published prefix among the last 32 segments plus the complete seg_acc sums in
// front of it -- tolerant of a lagging scan warp -- was built and measured: 0.54 -> 0.40 of the roofline.  Every tile then
// reads the 32 seg_acc words that the tiles of the current wave are busy adding to; those few L2 sectors cannot serve 4 M
// reads next to the atomics.  seg_acc is read by the scan warp only.)
struct LookbackPre {
    uint32_t cnt;                 // lane < (t & 31): state of tile t - 1 - lane
    unsigned long long seg;       // every lane: seg_pre of the segment before t's
};
__device__ __forceinline__ LookbackPre lookback_prefetch(const Lookback& lb, uint32_t t, uint32_t lane) {
    LookbackPre p{0u, 0ull};
    if (lane < (t & 31u)) p.cnt = ld_volatile_u32(lb.tile_cnt + (t - 1u - lane));
    if (t >> 5) p.seg = ld_volatile_u64(lb.seg_pre + ((t >> 5) - 1u));
    return p;
}
__device__ __forceinline__ unsigned long long lookback_resolve(const Lookback& lb, uint32_t t, uint32_t lane, LookbackPre pre, long long* t_seg = nullptr) {
    const bool mine = lane < (t & 31u);
    uint32_t st = pre.cnt;
    while (__any_sync(kFullMask, mine && !(st >> 31))) {
        if (mine && !(st >> 31)) {
            __nanosleep(40);      // spinning would steal issue slots
            st = ld_volatile_u32(lb.tile_cnt + (t - 1u - lane));
        }
    }
    const uint32_t own = __reduce_add_sync(kFullMask, mine ? (st & 0x7FFFFFFFu) : 0u);
    unsigned long long p = 0;
    if (t >> 5) {
        const long long t0 = (FQG_FASTQ_TIMING && t_seg) ? clock64() : 0;
        p = pre.seg;
        // (Falling back, on a miss, to the nearest published prefix among the last four segments plus the complete seg_acc
        // sums in front of it was measured twice: the code costs the match kernel 8 registers it does not have -- spills in
        // the walk, 0.56 -> 0.46 -- and buys the index-only kernel nothing.)
        while ((p >> 62) != 2) {      // warp-uniform: every lane reads the same word
            __nanosleep(40);
            p = ld_volatile_u64(lb.seg_pre + ((t >> 5) - 1u));
        }
        p &= kLbValue;
        if (FQG_FASTQ_TIMING && t_seg) *t_seg += clock64() - t0;
    }
    return p + own;
}
// The scan warp of a launch over tiles [first, end): segments whose last tile lies in that range get their prefix here.
// Four segments per lane and step (one 32-byte sector of seg_acc each, two vector loads): a step costs an L2 round trip, and
// at 32 segments per step the warp could only just follow the tile warps (12 segments per microsecond) -- they then waited
// for IT, 15 % of their time in the match kernel and 34 % in the index-only kernel (FQG_FASTQ_STATS).
__device__ __noinline__ void scan_segments(Lookback lb, uint32_t first, uint32_t end, uint32_t lane) {
    const uint32_t s_hi = end >> 5;
    uint32_t base = first >> 5;
    unsigned long long carry = base ? (ld_volatile_u64(lb.seg_pre + (base - 1u)) & kLbValue) : 0ull;      // left by an earlier launch
    while (base < s_hi) {
        const uint32_t s0 = (base & ~3u) + 4u * lane;
        unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
        if (s0 < s_hi) {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(acc[0]), "=l"(acc[1]) : "l"(lb.seg_acc + s0));
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(acc[2]), "=l"(acc[3]) : "l"(lb.seg_acc + s0 + 2u));
        }
        // segments in front of `base` are done (they count as complete and add nothing), segments from s_hi on are not ours
        uint32_t lead = 0;
        bool all = true;
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            const uint32_t sgm = s0 + j;
            const bool ok = sgm < base || (sgm < s_hi && (acc[j] >> 48) == 32);
            all = all && ok;
            lead += all ? 1u : 0u;
        }
        const uint32_t full = __ballot_sync(kFullMask, lead == 4u);
        const uint32_t fp = full == kFullMask ? 32u : (uint32_t)__ffs((int)~full) - 1u;      // first lane that is not complete
        const uint32_t new_end = (base & ~3u) + 4u * fp + (fp < 32u ? __shfl_sync(kFullMask, lead, fp & 31u) : 0u);
        if (new_end <= base) { __nanosleep(60); continue; }
        uint32_t v[4], lane_tot = 0;      // a segment holds at most 32 x 10752 newlines
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            const uint32_t sgm = s0 + j;
            v[j] = (sgm >= base && sgm < new_end) ? (uint32_t)(acc[j] & kLbSum) : 0u;
            lane_tot += v[j];
        }
        uint32_t incl = lane_tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= (uint32_t)d) incl += u;
        }
        unsigned long long running = carry + (incl - lane_tot);
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            const uint32_t sgm = s0 + j;
            running += v[j];
            if (sgm >= base && sgm < new_end) atomicExch(lb.seg_pre + sgm, (2ull << 62) | running);
        }
        carry += __shfl_sync(kFullMask, incl, 31);
        base = new_end;
    }
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

// ---- one record, one lane ---------------------------------------------------------------------------------------------------
// Everything a lane finds out about its record.  Nothing is written to global memory while a record is evaluated: the
// record phase may run under a GUESSED line phase (see the kernel), so outputs -- mask bit, id hash, spans, errors -- are
// committed only once the look-back has confirmed the guess.
struct RecOut {
    NamePending name;                    // TIER 3 only
    bool sel = false;
    uint32_t err = 0;                    // 0 or one of kErr*
    unsigned long long err_pos = 0;
    unsigned long long idh = 0, seq_pos = 0, rec_off = 0;
    uint32_t seq_n = 0, rec_n = 0;
};

struct TileCtx {
    uint32_t tile_sa, list_sa;           // shared addresses of the staged bytes and of the newline list
    uint32_t total;                      // newlines in the staged bytes (list entries when total <= kListCap)
    uint64_t T0, stream_len;
    bool want_idh;
};

// Record that starts at stream position gs, read straight from global memory: records that do not end inside tile + overlap
// (reads longer than ~400 bases) and the slow second look at a tile whose guessed line phase turned out wrong.
template <int TIER>
__device__ __noinline__ RecOut eval_record_global(const FastqArgs& a, const RecordEval<TIER>& ev, const TileCtx& c, uint64_t gs) {
    RecOut o;
    const uint8_t* base = a.buf;
    uint64_t q[4];
    uint64_t p = gs;
    int found = 0;
    for (; found < 4 && p < c.stream_len; p++)
        if (p == a.nbytes || base[p] == '\n') q[found++] = p;
    if (found < 4) { o.err = kErrUnexpectedEnd; o.err_pos = gs; }
    else {
        const uint32_t cr0 = (q[0] > gs + 1 && base[q[0] - 1] == '\r') ? 1u : 0u;
        const uint32_t cr1 = (q[1] > q[0] + 1 && base[q[1] - 1] == '\r') ? 1u : 0u;
        const uint32_t cr3 = (q[3] > q[2] + 1 && base[q[3] - 1] == '\r') ? 1u : 0u;
        const uint64_t seq_len = q[1] - q[0] - 1 - cr1, qual_len = q[3] - q[2] - 1 - cr3;
        if (base[gs] != '@') { o.err = kErrInvalidStart; o.err_pos = gs; }
        else if (base[q[1] + 1] != '+') { o.err = kErrInvalidSep; o.err_pos = q[1] + 1; }
        else if (seq_len != qual_len) { o.err = kErrUnequal; o.err_pos = gs; }
        o.seq_pos = q[0] + 1;
        o.seq_n = (uint32_t)seq_len;
        if (a.rec_off) {
            const bool virt = q[3] == a.nbytes;
            const bool norm = (cr0 | cr1 | cr3) != 0 || q[2] - q[1] != 2u || virt;
            o.rec_off = gs | (norm ? 1ull << 63 : 0ull);
            o.rec_n = (uint32_t)(q[3] - gs + (virt ? 0u : 1u));
        }
        unsigned long long idh = 0;      // (its address goes to an out-of-line function: keep that away from `o`, which lives in registers)
        o.sel = ev.eval_global(base + gs + 1, q[0] > gs ? (uint32_t)(q[0] - gs - 1 - cr0) : 0u, base + q[0] + 1, (uint32_t)seq_len, &idh, c.want_idh);
        o.idh = idh;
    }
    return o;
}

// Record whose line ends have ranks r0 .. r0+3 in the newline list (list entry of rank r = list[r - win_lo]) and that starts
// at tile position s.
template <int TIER>
__device__ __forceinline__ RecOut eval_record(const FastqArgs& a, const RecordEval<TIER>& ev, const TileCtx& c, uint32_t s, uint32_t r0, uint32_t win_lo) {
    RecOut o;
    const uint64_t T0 = c.T0;
    if (r0 + 3u < c.total) {
        const uint32_t li = c.list_sa + 2u * (r0 - win_lo);
        const uint32_t p0 = lds_u16_volatile(li), p1 = lds_u16_volatile(li + 2u), p2 = lds_u16_volatile(li + 4u), p3 = lds_u16_volatile(li + 6u);
        const uint32_t tile_sa = c.tile_sa;
        const uint32_t cr1 = (p1 > p0 + 1u && lds_u8_volatile(tile_sa + p1 - 1u) == '\r') ? 1u : 0u;
        const uint32_t cr3 = (p3 > p2 + 1u && lds_u8_volatile(tile_sa + p3 - 1u) == '\r') ? 1u : 0u;
        const uint32_t seq_len = p1 - p0 - 1u - cr1, qual_len = p3 - p2 - 1u - cr3;
        if (lds_u8_volatile(tile_sa + s) != '@') { o.err = kErrInvalidStart; o.err_pos = T0 + s; }
        else if (lds_u8_volatile(tile_sa + p1 + 1u) != '+') { o.err = kErrInvalidSep; o.err_pos = T0 + p1 + 1u; }
        else if (seq_len != qual_len) { o.err = kErrUnequal; o.err_pos = T0 + s; }
        o.seq_pos = T0 + p0 + 1u;
        o.seq_n = seq_len;
        if (a.rec_off) {
            const uint32_t cr0 = (p0 > s + 1u && lds_u8_volatile(tile_sa + p0 - 1u) == '\r') ? 1u : 0u;
            const bool virt = T0 + p3 == a.nbytes;
            const bool norm = (cr0 | cr1 | cr3) != 0 || p2 - p1 != 2u || virt;
            o.rec_off = (T0 + s) | (norm ? 1ull << 63 : 0ull);
            o.rec_n = p3 - s + (virt ? 0u : 1u);
        }
        unsigned long long idh = 0;
        o.sel = ev.eval_shared(tile_sa, s, p0, seq_len, &idh, c.want_idh, &o.name);
        o.idh = idh;
    } else {
        o = eval_record_global<TIER>(a, ev, c, T0 + s);      // record runs past tile+overlap (or the stream is truncated)
    }
    return o;
}

// second half of a read-name lookup (see NamePending)
template <int TIER>
__device__ __forceinline__ void finish_names(const RecordEval<TIER>& ev, RecOut& o) {
    if (TIER == 3 && o.name.active) o.sel = ev.names.finish_shared(o.name) != ev.invert;
}

// Writes what 32 lanes found out about records R_w .. R_w + 31 (lane i: record R_w + i, `valid` lanes only); returns the number selected.
__device__ __forceinline__ uint32_t commit_records(const FastqArgs& a, uint64_t T0, bool want_idh, const RecOut& o_in, bool valid, unsigned long long R_w, uint32_t lane) {
    RecOut o = o_in;
    if (valid) {
        const unsigned long long R = R_w + lane;
        if (o.err) report(a, o.err_pos, o.err);
        if (R >= a.mask_bits || (want_idh && R >= a.id_hash_cap)) { report(a, T0, kErrMaskOverflow); o.sel = false; }
        else if (want_idh) a.id_hash[R] = o.idh;
        if (a.seq_start && R < a.span_cap) { a.seq_start[R] = o.seq_pos; a.seq_len[R] = o.seq_n; }
        if (a.soa_start && R < a.span_cap) { a.soa_start[R] = (uint32_t)o.seq_pos; a.soa_end[R] = (uint32_t)o.seq_pos + o.seq_n; }
        if (a.rec_off && R < a.rec_cap) { a.rec_off[R] = o.rec_off; a.rec_len[R] = o.rec_n; }
    }
    const uint32_t mk = __ballot_sync(kFullMask, valid && o.sel);
    if (lane == 0 && mk) {
        const uint32_t sh = (uint32_t)(R_w & 31ull);
        atomicOr(a.mask + (R_w >> 5), mk << sh);
        if (sh && (mk >> (32u - sh))) atomicOr(a.mask + (R_w >> 5) + 1, mk >> (32u - sh));
    }
    return (uint32_t)__popc(mk);
}

// records that start in the tile, given the number of newlines before it
struct TilePhase {
    uint32_t extra, j0, n_list;
    unsigned long long R_base;
};
__device__ __forceinline__ TilePhase tile_phase(const TileGeom& g, uint32_t j0, unsigned long long Lt, uint32_t agg, uint64_t stream_len) {
    // record k begins right after the newline of local rank j0 + 4 (k - extra); j0 = first local rank that ends a record
    TilePhase p;
    const bool first_global = g.t == 0;
    p.extra = first_global ? 1u : 0u;
    p.j0 = j0;
    p.R_base = first_global ? 0ull : (Lt + j0 + 1ull) >> 2;
    p.n_list = p.extra + (agg > j0 ? (agg - j0 + 3u) / 4u : 0u);
    if (g.T0 + g.region == stream_len && agg > j0 && ((agg - j0 - 1u) & 3u) == 0) p.n_list--;   // the stream's final newline starts nothing
    return p;
}

// A tile whose records have to be found again under the exact line phase because the guessed one was wrong -- which means
// that the first records of the tile are malformed (a well-formed tile always confirms its guess), so this only ever runs
// on damaged input and may be slow: ONE lane walks the tile's bytes in global memory, record by record.  It does not touch
// the warp's shared-memory tile (which may already hold the next tile).
template <int TIER>
__device__ __noinline__ uint32_t redo_tile_global(const FastqArgs& a, const RecordEval<TIER>& ev, TileCtx c, uint32_t t, unsigned long long Lt, uint32_t lane) {
    c.T0 = (uint64_t)t * kTile;
    const uint64_t end = c.T0 + kTile < c.stream_len ? c.T0 + kTile : c.stream_len;
    uint64_t pos = c.T0;
    unsigned long long lines = Lt;      // newlines before pos
    bool stream_start = t == 0;
    uint32_t selected = 0;
    for (;;) {
        unsigned long long gs = ~0ull, R = 0;      // next record that starts in the tile: right after the 4th newline of its predecessor
        if (lane == 0) {
            if (stream_start) gs = 0;
            else
                while (pos < end) {
                    const bool nl = pos == a.nbytes || a.buf[pos] == '\n';
                    pos++;
                    if (nl && ((++lines) & 3ull) == 0 && pos < c.stream_len) { gs = pos; R = lines >> 2; break; }
                }
        }
        stream_start = false;
        gs = __shfl_sync(kFullMask, gs, 0);
        R = __shfl_sync(kFullMask, R, 0);
        if (gs == ~0ull) break;
        RecOut o;
        if (lane == 0) o = eval_record_global<TIER>(a, ev, c, gs);
        selected += commit_records(a, c.T0, c.want_idh, o, lane == 0, R, lane);
    }
    return selected;
}

// Tiles with more newlines than the list holds (records of a few bytes: the reference's own test fixtures): several windows
// over the ranks, each scattered and consumed in turn.  Rare, so kept out of line; no guessing here.
struct LaneWords { uint32_t w[kWords]; };      // a lane's newline bitmap, by value: the hot path keeps it in registers
template <int TIER>
__device__ __noinline__ uint32_t dense_tile(const FastqArgs& a, const RecordEval<TIER>& ev, const TileCtx& c, const TileGeom& g, LaneWords m, uint32_t excl,
                                            uint32_t agg, unsigned long long Lt, uint32_t lane) {
    const TilePhase ph = tile_phase(g, 3u - (uint32_t)(Lt & 3ull), Lt, agg, c.stream_len);
    const uint32_t lane_pos = lane * kLaneBytes;
    uint32_t selected = 0;
    for (uint32_t win_lo = 0;; win_lo += kWindowStep) {
        uint32_t idx = excl - win_lo;       // may wrap below zero: compared unsigned against the capacity
#pragma unroll
        for (uint32_t j = 0; j < kWords; j++) {
            uint32_t w = m.w[j];
            const uint32_t pos0 = lane_pos + 32u * j;
            while (w) {
                const uint32_t b = (uint32_t)__ffs((int)w) - 1u;
                if (idx < kListCap) sts_u16(c.list_sa + 2u * idx, pos0 + b);
                idx++;
                w &= w - 1u;
            }
        }
        __syncwarp();
        for (uint32_t k0 = 0; k0 < ph.n_list; k0 += 32u) {
            const uint32_t k = k0 + lane;
            // ranks: qa = the newline before the record (none for the stream's first record), r0 .. r0+3 = its four line ends
            const uint32_t qa = k < ph.extra ? 0u : ph.j0 + 4u * (k - ph.extra);
            const uint32_t r0 = k < ph.extra ? 0u : qa + 1u;
            const bool valid = k < ph.n_list && qa >= win_lo && qa < win_lo + kWindowStep;
            RecOut o;
            if (valid) {
                const uint32_t s = k < ph.extra ? 0u : lds_u16_volatile(c.list_sa + 2u * (qa - win_lo)) + 1u;
                o = eval_record<TIER>(a, ev, c, s, r0, win_lo);
                finish_names<TIER>(ev, o);
            }
            selected += commit_records(a, c.T0, c.want_idh, o, valid, ph.R_base + k0, lane);
        }
        __syncwarp();       // every lane is done with the list before the next window overwrites it
        if (win_lo + kWindowStep >= agg) break;
    }
    return selected;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------------
// Every warp of the grid is a tile warp, except the last warp of CTA 0: that one is the scan warp (see Lookback) and owns
// no tiles.  (An extra warp per CTA for it would push the CTA past 512 threads and cost every thread 32 registers.)
// FEAT: which per-record outputs beside the mask the launch asks for -- bit 0 the 64-bit id hashes (paired / interleaved
// input), bit 1 any of the record index / spans (seq_start, soa_start, rec_off).  They are compiled out where they are not
// wanted: the kernel runs at 118-124 of its 128 registers, and results that wait a tile for their commit (step 7) keep every
// output field alive across the next tile's scan.
template <int TIER, int FEAT>
__global__ void __launch_bounds__(kMaxTileWarps * 32u, 1) match_fastq_kernel(FastqArgs a_in, DeviceDfa dfa, uint32_t table_smem_bytes) {
    FastqArgs a = a_in;
    if (!(FEAT & 1)) { a.id_hash = nullptr; a.id_hash_cap = 0; }
    if (!(FEAT & 2)) { a.seq_start = nullptr; a.seq_len = nullptr; a.soa_start = nullptr; a.soa_end = nullptr; a.rec_off = nullptr; a.rec_len = nullptr; }
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, cta_warps = blockDim.x >> 5;

    // ---- smem carve-up: [automaton][class LUT 256][letters-of-codes table 1024, packed tiers][per-warp regions]
    uint8_t* s_tab = smem;
    uint8_t* s_cls = smem + table_smem_bytes;
    uint32_t* s_etab = reinterpret_cast<uint32_t*>(s_cls + 256);
    uint8_t* s_warp = s_cls + 256 + kEtabBytes<TIER> + (size_t)warp * kWarpBytes;
    const uint32_t bar = smem_addr(s_warp), list_sa = bar + kWarpListOff, tile_sa = bar + kWarpTileOff;

    stage_automaton<TIER>(dfa, s_tab, s_cls);
    if (kPackedCodes<TIER> && FQG_ETAB) stage_etab(s_etab, threadIdx.x, blockDim.x);
    if (lane == 0) mbar_init(bar, 1);
    mbar_fence_init();
    __syncthreads();       // the only CTA-wide barrier: table staged, mbarriers initialised

    const Lookback lb{reinterpret_cast<uint32_t*>(a.tile_state), a.seg_acc, a.seg_pre};
    const uint32_t scan_warp = cta_warps - 1u;      // of CTA 0
    if (blockIdx.x == 0 && warp == scan_warp && a.scan_warp) {
        scan_segments(lb, a.tile_first, a.tile_first + a.tile_count, lane);
        return;
    }

    RecordEval<TIER> ev_init = make_eval<TIER>(a, dfa, s_tab, s_cls);
    if (kPackedCodes<TIER> && FQG_ETAB) ev_init.view.etab = smem_addr(s_etab);
    const RecordEval<TIER> ev = ev_init;
    TileCtx c;
    c.tile_sa = tile_sa;
    c.list_sa = list_sa;
    // the stream always ends with a newline: a real one, or a virtual one at offset nbytes (missing final '\n')
    c.stream_len = a.nbytes + (a.virtual_final_nl ? 1u : 0u);
    c.want_idh = a.id_hash != nullptr;
    const uint64_t stream_len = c.stream_len;
    const uint32_t c0a = a.nl_xor, c7f = a.nl_low7;     // see FastqArgs: not compile-time constants on purpose
    const uint32_t lane_pos = lane * kLaneBytes;       // this lane's first byte in the staged tile
    uint32_t phase = 0;
    uint32_t local_count = 0;      // (a stream holds fewer than 2^32 records: <= 4 GiB of at least 8 bytes each)
    const bool timing = FQG_FASTQ_TIMING && a.stats;      // measurement hook: cycles spent waiting (see FQG_FASTQ_TIMING)
    long long t_copy = 0, t_prefix = 0, t_seg = 0, t_all = timing ? clock64() : 0;

    // Tiles are claimed in order from an atomic counter.  (Dealing them out statically -- warp w takes tiles w, w + W, ... --
    // saves the atomic but was measured slower: warps do not run at the same speed (the issue arbiter prefers high warp ids,
    // one SM also hosts the scan warp), a tile needs the counts of ALL earlier tiles, and with a fixed deal everybody ends up
    // waiting for the slowest warp's tiles: 15 % of all warp cycles in the match kernel, 34 % in the index-only kernel,
    // FQG_FASTQ_STATS.  With claims, slow warps simply take fewer tiles.)  The claim for the next tile is issued just before
    // this tile's records are walked -- after the only place where the warp may wait for other tiles (step 7) -- and picked up
    // after the walk, so its round trip is hidden and the warp never waits while it holds a tile it has not started.
    const uint32_t W = gridDim.x * cta_warps - a.scan_warp;      // tile warps of the grid
    // Starts the copy of tile `cl` of this launch into the warp's buffer (the buffer must be free) and pulls the tile after it
    // into L2 (126 MB; a wave of tiles is ~25 MB), so that one's copy will only pay L2 latency: a warp has a single tile
    // buffer, hence no other way to keep HBM busy.
    auto issue_tile = [&](uint32_t cl) {
        if (lane == 0 && cl < a.tile_count) {
            const uint32_t tn = a.tile_first + cl, avail_n = tile_real_avail(a, tn);
            if (avail_n) {
                const uint32_t bytes = (avail_n + 15u) & ~15u;
                mbar_arrive_expect_tx(bar, bytes);
                bulk_copy_g2s(tile_sa, a.buf + (uint64_t)tn * kTile, bytes, bar);
            }
            if (cl + W < a.tile_count) {
                const uint32_t ta = tn + W, avail_a = tile_real_avail(a, ta);
                if (avail_a) bulk_prefetch_l2(a.buf + (uint64_t)ta * kTile, (avail_a + 15u) & ~15u);
            }
        }
    };
    // Results of the previous tile that still wait for their global record numbers (see step 7 in the loop).
    // (Short reads put more than 32 records into a tile.  Letting every lane take a second record and deferring tiles of up to
    // 64 records was built and measured: the second set of pending results spills -- 150-base reads 0.63 -> 0.40, 100-base
    // reads 0.39 -> 0.32; a compact form that keeps only the second record's selected bit and id hash, evaluated in a
    // two-trip loop, loses as well: 0.61 -> 0.50 and 0.44 -> 0.41.  Such tiles are settled at once, after their prefix is
    // resolved and before the next claim.)
    constexpr uint32_t kDeferMax = 32u;
    bool pend = false, p_valid = false, p_last = false;
    RecOut po;
    uint32_t p_t = 0, p_agg = 0, p_j0 = 0;
    LookbackPre p_pre{0u, 0ull};
    auto complete_pending = [&]() {
        if (!pend) return;
        pend = false;
        const long long t0 = timing ? clock64() : 0;
        const unsigned long long Lp = lookback_resolve(lb, p_t, lane, p_pre, timing ? &t_seg : nullptr);
        if (timing) t_prefix += clock64() - t0;
        if (3u - (uint32_t)(Lp & 3ull) != p_j0) {      // damaged input only
            if (FQG_FASTQ_COUNTERS && a.stats && lane == 0) atomicAdd(a.stats + 2, 1ull);
            local_count += redo_tile_global<TIER>(a, ev, c, p_t, Lp, lane);
        } else {
            local_count += commit_records(a, (uint64_t)p_t * kTile, c.want_idh, po, p_valid, (Lp + p_j0 + 1ull) >> 2, lane);
        }
        if (lane == 0 && p_last) *a.lines_out = Lp + p_agg;
    };
    uint32_t claim = 0, next_claim = 0;
    if (lane == 0) claim = atomicAdd(a.tile_counter, 1u);
    claim = __shfl_sync(kFullMask, claim, 0);
    issue_tile(claim);
    bool claimed = false;
    auto claim_next = [&]() {      // warp-uniform; issued once per tile, the value is picked up by take_next
        if (!claimed && lane == 0) next_claim = atom_add_u32(a.tile_counter, 1u);
        claimed = true;
    };
    auto take_next = [&]() {
        claim_next();
        claimed = false;
        return __shfl_sync(kFullMask, next_claim, 0);
    };
    while (claim < a.tile_count) {
        const TileGeom g = tile_geom(a, a.tile_first + claim, stream_len);
        c.T0 = g.T0;
        // The copy of the NEXT tile is started as soon as the last record of this one has been walked (the prefix is resolved
        // and the results are committed from registers while it travels): 7.6 % of all stall samples used to sit on this wait
        // (profiles/r2j_*).
        bool next_issued = false;
        // The prefix words of the pending tile are requested here -- its count went out most of a tile-time ago, so they have
        // been published by now (asked for right after its own count they never were: every resolve then paid a second L2
        // round trip) -- and consumed in step 7, after this tile's bytes have been scanned.
        if (pend) p_pre = lookback_prefetch(lb, p_t, lane);
        if (g.real_avail) {
            const long long t0 = timing ? clock64() : 0;
            mbar_wait(bar, phase);       // try_wait suspends the warp in hardware; no issue slots are burnt
            phase ^= 1u;
            if (timing) t_copy += clock64() - t0;
        }

        // ---- 2. newline masks of this lane's 336 bytes, in registers
        uint32_t m[kWords];
        {
            const uint32_t base = tile_sa + lane_pos;
#pragma unroll
            for (uint32_t j = 0; j < kWords; j++) {
                const uint4 lo = lds_v4_volatile(base + 32u * j);
                const uint32_t a0 = newline_mask8x128(lo.x, lo.y, c0a, c7f), b0 = newline_mask8x128(lo.z, lo.w, c0a, c7f);
                if (2u * j + 1u < kUnits) {
                    const uint4 hi = lds_v4_volatile(base + 32u * j + 16u);
                    const uint32_t a1 = newline_mask8x128(hi.x, hi.y, c0a, c7f), b1 = newline_mask8x128(hi.z, hi.w, c0a, c7f);
                    m[j] = (a0 >> 7) + (b0 << 1) + (a1 << 9) + (b1 << 17);
                } else {
                    m[j] = (a0 >> 7) + (b0 << 1);
                }
            }
            if (g.real_avail < kLoad) {      // warp-uniform and rare: the stream ends inside this buffer
#pragma unroll
                for (uint32_t j = 0; j < kWords; j++) {
                    const uint32_t lo = lane_pos + 32u * j, width = (2u * j + 1u < kUnits) ? 32u : 16u;
                    if (lo + width > g.real_avail) m[j] &= g.real_avail > lo ? (1u << (g.real_avail - lo)) - 1u : 0u;     // stale bytes
                    if (g.real_avail < g.avail && g.real_avail >= lo && g.real_avail < lo + width) m[j] |= 1u << (g.real_avail - lo);   // the virtual final newline
                }
            }
        }

        // ---- 3. ranks: newlines before this lane's bytes; the tile's own count (lanes 0..28, or everything in the stream's last tile)
        uint32_t pc[kWords], cnt = 0;
#pragma unroll
        for (uint32_t j = 0; j < kWords; j++) { pc[j] = (uint32_t)__popc(m[j]); cnt += pc[j]; }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= (uint32_t)d) incl += v;
        }
        const uint32_t total = __shfl_sync(kFullMask, incl, 31);
        const uint32_t agg = g.region == kTile ? __shfl_sync(kFullMask, incl, kTileLanes - 1u) : total;
        c.total = total;
        lookback_publish(lb, g.t, agg, lane);
        const LookbackPre no_pre{0u, 0ull};

        if (FQG_FASTQ_COUNTERS && a.stats && lane == 0) atomicAdd(a.stats + (total > kListCap ? 3 : 0), 1ull);
        if (total > kListCap) {
            complete_pending();
            const unsigned long long Lt = lookback_resolve(lb, g.t, lane, no_pre);
            claim_next();
            if (lane == 0 && g.T0 + g.region == stream_len) *a.lines_out = Lt + agg;
            LaneWords mw;
#pragma unroll
            for (uint32_t j = 0; j < kWords; j++) mw.w[j] = m[j];
            local_count += dense_tile<TIER>(a, ev, c, g, mw, incl - cnt, agg, Lt, lane);
        } else {
            // ---- 4. scatter: list[rank] = position in the tile.  The first two newlines of every 32-byte word are stored
            //         straight away, word by word and independently of each other (a lane rarely has more than two in 32
            //         bytes: the "\n+\n" of a record); only if some lane has a third one somewhere the generic loop runs.
            {
                uint32_t addr = list_sa + 2u * (incl - cnt), leftover = 0;
#pragma unroll
                for (uint32_t j = 0; j < kWords; j++) {
                    const uint32_t rw = __brev(m[j]), pos0 = lane_pos + 32u * j;
                    const uint32_t f1 = (uint32_t)__clz((int)rw);                         // 32 when the word is empty
                    sts_u16_if(addr, pos0 + f1, rw);
                    const uint32_t rw2 = rw & __funnelshift_rc(0x7FFFFFFFu, 0u, f1);      // first one cleared
                    const uint32_t f2 = (uint32_t)__clz((int)rw2);
                    sts_u16_if(addr + 2u, pos0 + f2, rw2);
                    leftover |= rw2 & __funnelshift_rc(0x7FFFFFFFu, 0u, f2);
                    addr += 2u * pc[j];
                }
                if (__any_sync(kFullMask, leftover != 0)) {
                    addr = list_sa + 2u * (incl - cnt);
#pragma unroll
                    for (uint32_t j = 0; j < kWords; j++) {
                        uint32_t w = m[j];
                        const uint32_t pos0 = lane_pos + 32u * j;
                        while (w) {
                            sts_u16(addr, pos0 + (uint32_t)__ffs((int)w) - 1u);
                            addr += 2u;
                            w &= w - 1u;
                        }
                    }
                }
            }
            __syncwarp();

            // ---- 5. the line phase.  Which local newline ends a record depends on the number of newlines before the tile
            //         (mod 4), which is only known once every earlier tile has been scanned.  Instead of waiting, the phase
            //         is GUESSED from the tile itself: lanes 0..7 test, for the four candidates and the first two records
            //         each, "starts with '@', '+' two lines on, equal lengths".  One surviving candidate = the guess; the
            //         records are evaluated under it into registers and only a confirmed guess is committed (step 7).  A
            //         wrong or impossible guess costs a second look, never a wrong answer.
            unsigned long long Lt = 0;
            bool have_Lt = g.t == 0, guessed = false;          // nothing precedes the first tile
            uint32_t j0 = 3u;
            LookbackPre pre = no_pre;
            if (!have_Lt) {
                bool ok = false;
                const uint32_t cand = lane & 3u, r = cand + 4u * (lane >> 2);      // boundary newline rank under this candidate
                if (lane < 8u && r + 4u < total) {
                    const uint32_t li = list_sa + 2u * r;
                    const uint32_t q = lds_u16_volatile(li), p0 = lds_u16_volatile(li + 2u), p1 = lds_u16_volatile(li + 4u), p2 = lds_u16_volatile(li + 6u),
                                   p3 = lds_u16_volatile(li + 8u);
                    ok = lds_u8_volatile(tile_sa + q + 1u) == '@' && lds_u8_volatile(tile_sa + p1 + 1u) == '+' && p1 - p0 == p3 - p2;
                }
                const uint32_t votes = __ballot_sync(kFullMask, ok);
                const uint32_t both = votes & (votes >> 4) & 0xFu;              // candidates whose first AND second record look right
                if (both && (both & (both - 1u)) == 0) {
                    guessed = true;
                    j0 = (uint32_t)__ffs((int)both) - 1u;
                    if (FQG_FASTQ_COUNTERS && a.stats && lane == 0) atomicAdd(a.stats + 1, 1ull);
                    if (TIER == 3) pre = lookback_prefetch(lb, g.t, lane);      // read names settle their tile at once, see step 6
                }
            }
            // ---- 7 (of the PREVIOUS tile).  Its records were evaluated under a guess and wait in registers.  By now this
            //         tile's count has been published and its bytes scanned -- a good third of a tile-time after the previous
            //         tile was walked -- so the prefix of the previous tile is there without waiting: confirm its guess, commit.
            //         (Resolving right after the walk made the warps wait 8 % of their time in the match kernel and 25 % in
            //         the index-only kernel, FQG_FASTQ_STATS.)
            if (TIER != 5 || have_Lt || !guessed) complete_pending();      // (index-only: later still, see below)
            if (!have_Lt && !guessed) {      // no usable guess (no two whole records in the tile, or ambiguous): wait for the prefix
                Lt = lookback_resolve(lb, g.t, lane, no_pre);
                have_Lt = true;
                j0 = 3u - (uint32_t)(Lt & 3ull);
            }
            // ---- 6. one lane per record that starts in the tile
            const bool last_tile = g.T0 + g.region == stream_len;
            TilePhase ph = tile_phase(g, j0, Lt, agg, stream_len);
            // The claim must not be outstanding while the warp waits for other tiles: read names, and guesses that cover more
            // than 32 records (short reads), resolve their prefix below and claim right after that.
            if (have_Lt || (TIER != 3 && ph.n_list <= kDeferMax)) claim_next();
            if (!have_Lt && ph.n_list <= kDeferMax) {
                // the usual case: a guessed phase and at most one record per lane.  Evaluate, hand the buffer to the next
                // tile's copy, keep the results for step 7.
                const bool valid = lane < ph.n_list;
                const uint32_t qa = lane < ph.extra ? 0u : ph.j0 + 4u * (lane - ph.extra);
                RecOut o;
                if (valid) {
                    const uint32_t s = lane < ph.extra ? 0u : lds_u16_volatile(list_sa + 2u * qa) + 1u;
                    o = eval_record<TIER>(a, ev, c, s, lane < ph.extra ? 0u : qa + 1u, 0u);
                }
                if (TIER == 3) {
                    // Read names: a lookup is two dependent L2 / HBM round trips (bucket, then the stored id), and the id bytes it
                    // is compared with sit in the tile, so nothing is deferred here: the prefix is resolved while the buckets
                    // travel, then the lookups are finished, the buffer goes to the next copy and the results are committed.
                    // (Carrying the lookup across the deferral -- bucket checked when the next tile's copy has landed, stored id
                    // compared against the id's bytes re-read from global memory in step 7 -- was built and measured: the pending
                    // lookup state does not fit next to the newline masks, 400-550 bytes of spills, 0.40 -> 0.23.)
                    Lt = lookback_resolve(lb, g.t, lane, pre);
                    have_Lt = true;
                    claim_next();
                    if (3u - (uint32_t)(Lt & 3ull) != j0) {
                        if (FQG_FASTQ_COUNTERS && a.stats && lane == 0) atomicAdd(a.stats + 2, 1ull);
                        local_count += redo_tile_global<TIER>(a, ev, c, g.t, Lt, lane);
                    } else {
                        finish_names<TIER>(ev, o);
                        if (FQG_EARLY_ISSUE) {      // the tile's bytes are not needed any more
                            __syncwarp();
                            claim = take_next();
                            issue_tile(claim);
                            next_issued = true;
                        }
                        local_count += commit_records(a, c.T0, c.want_idh, o, valid, (Lt + j0 + 1ull) >> 2, lane);
                    }
                    if (lane == 0 && last_tile) *a.lines_out = Lt + agg;
                } else {
                if (FQG_EARLY_ISSUE) {
                    __syncwarp();
                    claim = take_next();
                    issue_tile(claim);
                    next_issued = true;
                }
                // The index-only kernel has no walk to spend between a tile's count and its prefix, so there the previous tile is
                // settled only now, a whole tile-time after its count went out.
                if (TIER == 5) complete_pending();
                pend = true;
                po = o; po.name = NamePending();      // (settled by finish_names; nothing of it is needed any more)
                p_valid = valid; p_t = g.t; p_agg = agg; p_j0 = j0; p_last = last_tile;
                p_pre = LookbackPre{0u, 0ull};      // requested at the top of the next iteration (or polled, after the last tile)
                }
            } else {
                // Known prefix (first tile, no usable guess), or a guess that covers more than 32 records (short reads).  In the
                // second case the prefix is only asked for here and resolved inside the loop, after the first 32 records have
                // been evaluated -- their walk is what the request hides behind; nothing is committed before the guess holds.
                bool redone = false;
                if (!have_Lt) {
                    if (TIER == 3) {      // read names: resolved up front (their lookups are finished inside the loop)
                        Lt = lookback_resolve(lb, g.t, lane, pre);
                        have_Lt = true;
                        claim_next();
                        if (3u - (uint32_t)(Lt & 3ull) != j0) {
                            if (FQG_FASTQ_COUNTERS && a.stats && lane == 0) atomicAdd(a.stats + 2, 1ull);
                            local_count += redo_tile_global<TIER>(a, ev, c, g.t, Lt, lane);
                            redone = true;
                        } else ph = tile_phase(g, j0, Lt, agg, stream_len);
                    } else pre = lookback_prefetch(lb, g.t, lane);
                }
                for (uint32_t k0 = 0; !redone && k0 < ph.n_list; k0 += 32u) {
                    const uint32_t k = k0 + lane;
                    const uint32_t qa = k < ph.extra ? 0u : ph.j0 + 4u * (k - ph.extra);
                    const bool valid = k < ph.n_list;
                    RecOut o;
                    if (valid) {
                        const uint32_t s = k < ph.extra ? 0u : lds_u16_volatile(list_sa + 2u * qa) + 1u;
                        o = eval_record<TIER>(a, ev, c, s, k < ph.extra ? 0u : qa + 1u, 0u);
                    }
                    if (TIER != 3 && !have_Lt) {
                        Lt = lookback_resolve(lb, g.t, lane, pre);
                        have_Lt = true;
                        claim_next();      // (only now: the warp must not wait while it holds a tile it has not started)
                        if (3u - (uint32_t)(Lt & 3ull) != j0) {
                            if (FQG_FASTQ_COUNTERS && a.stats && lane == 0) atomicAdd(a.stats + 2, 1ull);
                            local_count += redo_tile_global<TIER>(a, ev, c, g.t, Lt, lane);
                            redone = true;
                            break;
                        }
                        ph = tile_phase(g, j0, Lt, agg, stream_len);      // same records, now with their global numbers
                    }
                    finish_names<TIER>(ev, o);
                    local_count += commit_records(a, c.T0, c.want_idh, o, valid, ph.R_base + k0, lane);
                }
                if (!have_Lt) {      // (a guess always covers two records, so the loop ran; kept for symmetry)
                    Lt = lookback_resolve(lb, g.t, lane, pre);
                    have_Lt = true;
                }
                if (lane == 0 && last_tile) *a.lines_out = Lt + agg;
            }
        }
        if (!next_issued) {
            __syncwarp();       // every lane is done with the list and the tile before the next bulk copy overwrites them
            claim = take_next();
            issue_tile(claim);
        }
    }
    complete_pending();
    if (lane == 0 && local_count && a.count) atomicAdd(a.count, (unsigned long long)local_count);
    if (timing && lane == 0) {
        atomicAdd(a.stats + 4, (unsigned long long)t_copy);
        atomicAdd(a.stats + 5, (unsigned long long)t_prefix);
        atomicAdd(a.stats + 7, (unsigned long long)t_seg);
        atomicAdd(a.stats + 6, (unsigned long long)(clock64() - t_all));
    }
}

// pair rule (src/main.rs:749-755) for mates in two streams: unit mask = m1 | m2; ids must agree (:741-747).
// Record counts are still on the device (written by the streams' last tiles), so they are read here.
__global__ void combine_pairs_kernel(const uint32_t* m1, const uint32_t* m2, uint32_t* out, uint64_t cap_words, const unsigned long long* h1, const unsigned long long* h2,
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
__global__ void combine_interleaved_kernel(const uint32_t* m, uint32_t* out, uint64_t cap_out_words, const unsigned long long* h, uint64_t idh_cap,
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
cudaError_t launch_fastq_t(const DeviceDfa& dfa, const FastqArgs& args, int sm_count, cudaStream_t stream) {
    FastqArgs a = args;
    if (a.tile_count == 0) return cudaSuccess;
    const uint32_t table_smem = kSmemTable<TIER> ? ((dfa.table_bytes + 127u) & ~127u) : TIER == 2 ? ((dfa.hot_entries * 4u + 127u) & ~127u) : 0u;
    const uint32_t budget = 227u * 1024u;
    uint32_t warps = (budget - table_smem - 256u - kEtabBytes<TIER> - 128u) / kWarpBytes;      // one tile buffer per warp
    if (warps > kMaxTileWarps) warps = kMaxTileWarps;
    if (warps < 1u) warps = 1u;
    // the scan warp is only there when this launch completes a segment of 32 tiles (see Lookback)
    a.scan_warp = ((a.tile_first + a.tile_count) >> 5) > (a.tile_first >> 5) ? 1u : 0u;
    a.full_tiles = a.nbytes >= kLoad ? (uint32_t)((a.nbytes - kLoad) / kTile + 1u) : 0u;      // tiles whose tile + overlap lie inside the real bytes
    const uint32_t need = a.tile_count + a.scan_warp;         // warps that find work
    if (need < (uint32_t)sm_count * warps) {                  // small inputs: spread the tiles over the SMs first
        warps = (need + (uint32_t)sm_count - 1u) / (uint32_t)sm_count;
        if (warps < 1u) warps = 1u;
    }
    if (a.scan_warp && warps < 2u && kMaxTileWarps >= 2u && need < 2u * (uint32_t)sm_count) warps = need < 2u ? 2u : warps;
    const uint32_t smem = table_smem + 256u + kEtabBytes<TIER> + warps * kWarpBytes;
    const int feat = (a.id_hash ? 1 : 0) | ((a.seq_start || a.soa_start || a.rec_off) ? 2 : 0);
    auto kern = feat == 0 ? match_fastq_kernel<TIER, 0> : feat == 1 ? match_fastq_kernel<TIER, 1> : feat == 2 ? match_fastq_kernel<TIER, 2> : match_fastq_kernel<TIER, 3>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    uint32_t grid = (need + warps - 1u) / warps;
    if (grid > (uint32_t)sm_count) grid = (uint32_t)sm_count;
    kern<<<grid, warps * 32u, smem, stream>>>(a, dfa, table_smem);
    return cudaGetLastError();
}

}  // namespace

uint32_t fastq_tile_bytes() { return kTile; }
uint32_t fastq_overlap_bytes() { return kOverlap; }

cudaError_t launch_match_fastq(const DeviceDfa& dfa, bool names, bool index_only, const FastqArgs& args, int sm_count, cudaStream_t stream) {
    FastqArgs a = args;
    a.nl_xor = 0x0A0A0A0Au;
    a.nl_low7 = 0x7F7F7F7Fu;
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
__device__ __forceinline__ uint64_t records_on_device(uint64_t cap, const unsigned long long* lines_dev) {
    if (!lines_dev) return cap;
    const uint64_t n = *lines_dev >> 2;
    return n < cap ? n : cap;
}
__global__ void scan_block_sums_kernel(const uint32_t* len, uint64_t n, uint32_t* block_sums, const unsigned long long* lines_dev) {
    __shared__ uint32_t s_w[kScanBlock / 32];
    n = records_on_device(n, lines_dev);
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
__global__ void scan_offsets_kernel(const uint32_t* len, uint64_t n, const uint32_t* block_sums, uint32_t* offsets, const unsigned long long* lines_dev) {
    __shared__ uint32_t s_w[kScanBlock / 32];
    n = records_on_device(n, lines_dev);
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t item[kScanItems], v = 0;
    for (uint32_t i = 0; i < kScanItems; i++) { item[i] = base + i < n ? len[base + i] : 0u; v += item[i]; }
    uint32_t tot;
    uint32_t run = block_sums[blockIdx.x] + block_exclusive_scan(v, s_w, &tot);
    for (uint32_t i = 0; i < kScanItems; i++) {
        if (base + i < n) offsets[base + i] = run;
        run += item[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanBlock - 1) offsets[n] = block_sums[gridDim.x];      // blocks past a device-side n add nothing
}
// one warp per record: its sequence line moves as ALIGNED 32-bit words of the destination -- every lane assembles its word from
// two aligned words of the source (the two are not aligned alike in general) -- with single bytes only in front of the first
// aligned word and behind the last one.  (A byte per lane and trip moved 32 bytes per warp instruction: 1.3 TB/s.)
__global__ void gather_bases_kernel(const uint8_t* fastq, const unsigned long long* seq_start, const uint32_t* seq_len, uint64_t n,
                                    const uint32_t* offsets, uint8_t* out, const unsigned long long* lines_dev) {
    n = records_on_device(n, lines_dev);
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n; r += warps) {
        const uint8_t* src = fastq + seq_start[r];
        uint8_t* dst = out + offsets[r];
        const uint32_t L = seq_len[r];
        uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;      // bytes in front of the first aligned word
        if (head > L) head = L;
        if (lane < head) dst[lane] = src[lane];
        const uint32_t n_words = (L - head) >> 2;
        const uintptr_t s2 = reinterpret_cast<uintptr_t>(src + head);
        const uint32_t sh = (uint32_t)(s2 & 3u) * 8u;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s2 & ~uintptr_t(3));
        uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
        for (uint32_t i = lane; i < n_words; i += 32u) {
            const uint32_t lo = sw[i], hi = sh ? sw[i + 1u] : 0u;      // (the stream is readable 16 bytes past its end)
            dw[i] = __funnelshift_r(lo, hi, sh);
        }
        const uint32_t done = head + 4u * n_words;
        if (lane < L - done) dst[done + lane] = src[done + lane];
    }
}

// ---- K5: spans of the selected records, in output order ----------------------------------------------------------------
// The write-back of the reference (src/main.rs:641-657, write_paired_record :682-727) walks the records again on the
// main thread; here the fused kernel has already left every record's (offset, length, "needs normalising") behind, so
// selecting is an order-preserving compaction of those by the unit mask: popcount per block, scan of the block sums,
// then every thread deals out the spans of its own mask words.  Pairs emit R1 then R2.
constexpr uint32_t kSelWordsPerThread = 4, kSelBlockWords = kScanBlock * kSelWordsPerThread;

__device__ __forceinline__ uint64_t sel_units(int mode, const unsigned long long* lines, uint64_t cap_units) {
    const uint64_t n = mode == FQG_INTERLEAVED ? (*lines >> 3) : (*lines >> 2);
    return n < cap_units ? n : cap_units;
}
__global__ void select_count_kernel(const uint32_t* mask, int mode, const unsigned long long* lines, uint64_t cap_units, uint32_t* block_sums) {
    __shared__ uint32_t s_w[kScanBlock / 32];
    const uint64_t n_words = (sel_units(mode, lines, cap_units) + 31) >> 5;
    const uint64_t w0 = (uint64_t)blockIdx.x * kSelBlockWords + (uint64_t)threadIdx.x * kSelWordsPerThread;
    uint32_t v = 0;
    for (uint32_t i = 0; i < kSelWordsPerThread; i++) if (w0 + i < n_words) v += (uint32_t)__popc(mask[w0 + i]);
    uint32_t tot;
    block_exclusive_scan(v, s_w, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void select_emit_kernel(const uint32_t* mask, int mode, const unsigned long long* lines, uint64_t cap_units, const uint32_t* block_sums, uint32_t n_blocks,
                                   const unsigned long long* off0, const uint32_t* len0, const unsigned long long* off1, const uint32_t* len1,
                                   fqg_span* out, unsigned long long* n_spans_out) {
    __shared__ uint32_t s_w[kScanBlock / 32];
    const uint64_t n_words = (sel_units(mode, lines, cap_units) + 31) >> 5;
    const uint64_t w0 = (uint64_t)blockIdx.x * kSelBlockWords + (uint64_t)threadIdx.x * kSelWordsPerThread;
    uint32_t word[kSelWordsPerThread], v = 0;
    for (uint32_t i = 0; i < kSelWordsPerThread; i++) { word[i] = w0 + i < n_words ? mask[w0 + i] : 0u; v += (uint32_t)__popc(word[i]); }
    uint32_t tot;
    uint64_t rank = (uint64_t)block_sums[blockIdx.x] + block_exclusive_scan(v, s_w, &tot);
    const unsigned long long kOff = (1ull << 63) - 1;
    auto span_of = [&](const unsigned long long* off, const uint32_t* len, uint64_t r, uint32_t flags) {
        const unsigned long long o = off[r];
        fqg_span sp;
        sp.offset = o & kOff;
        sp.len = len[r];
        sp.flags = flags | ((o >> 63) ? FQG_SPAN_NORMALISE : 0u);
        return sp;
    };
    for (uint32_t i = 0; i < kSelWordsPerThread; i++) {
        uint32_t w = word[i];
        while (w) {
            const uint64_t u = (w0 + i) * 32u + (uint32_t)__ffs((int)w) - 1u;
            w &= w - 1u;
            if (mode == FQG_SINGLE) out[rank] = span_of(off0, len0, u, 0u);
            else if (mode == FQG_PAIRED) { out[2 * rank] = span_of(off0, len0, u, 0u); out[2 * rank + 1] = span_of(off1, len1, u, FQG_SPAN_R2); }
            else { out[2 * rank] = span_of(off0, len0, 2 * u, 0u); out[2 * rank + 1] = span_of(off0, len0, 2 * u + 1, 0u); }
            rank++;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_spans_out = (unsigned long long)block_sums[n_blocks] * (mode == FQG_SINGLE ? 1u : 2u);
}
}  // namespace

uint64_t compact_scratch_words(uint64_t n_records) { return (n_records + kScanTile - 1) / kScanTile + 2; }

uint64_t span_scan_words(uint64_t mask_words) { return (mask_words + kSelBlockWords - 1) / kSelBlockWords + 2; }

cudaError_t launch_select_spans(const uint32_t* unit_mask, uint64_t mask_words, int mode, const unsigned long long* lines, const unsigned long long* off0,
                                const uint32_t* len0, const unsigned long long* off1, const uint32_t* len1, fqg_span* out, unsigned long long* n_spans_out,
                                uint32_t* block_sums, cudaStream_t stream) {
    const uint32_t n_blocks = (uint32_t)((mask_words + kSelBlockWords - 1) / kSelBlockWords);
    if (n_blocks == 0) return cudaMemsetAsync(n_spans_out, 0, 8, stream);
    const uint64_t cap_units = mask_words * 32u;
    select_count_kernel<<<n_blocks, kScanBlock, 0, stream>>>(unit_mask, mode, lines, cap_units, block_sums);
    scan_sums_kernel<<<1, kScanBlock, 0, stream>>>(block_sums, n_blocks);
    select_emit_kernel<<<n_blocks, kScanBlock, 0, stream>>>(unit_mask, mode, lines, cap_units, block_sums, n_blocks, off0, len0, off1, len1, out, n_spans_out);
    return cudaGetLastError();
}

cudaError_t launch_compact_bases(const uint8_t* fastq, const unsigned long long* seq_start, const uint32_t* seq_len, uint64_t n_records,
                                 uint8_t* bases_out, uint32_t* offsets_out, uint32_t* block_sums, cudaStream_t stream, const unsigned long long* lines_dev) {
    if (n_records == 0) return cudaMemsetAsync(offsets_out, 0, 4, stream);
    const uint32_t n_blocks = (uint32_t)((n_records + kScanTile - 1) / kScanTile);
    scan_block_sums_kernel<<<n_blocks, kScanBlock, 0, stream>>>(seq_len, n_records, block_sums, lines_dev);
    scan_sums_kernel<<<1, kScanBlock, 0, stream>>>(block_sums, n_blocks);
    scan_offsets_kernel<<<n_blocks, kScanBlock, 0, stream>>>(seq_len, n_records, block_sums, offsets_out, lines_dev);
    gather_bases_kernel<<<148 * 8, 256, 0, stream>>>(fastq, seq_start, seq_len, n_records, offsets_out, bases_out, lines_dev);
    return cudaGetLastError();
}

cudaError_t launch_combine_pairs(const uint32_t* m1, const uint32_t* m2, uint32_t* out, uint64_t cap_words, const unsigned long long* h1, const unsigned long long* h2,
                                 uint64_t idh_cap, const unsigned long long* lines1, const unsigned long long* lines2, unsigned long long* count,
                                 unsigned long long* id_mismatch, cudaStream_t stream) {
    combine_pairs_kernel<<<148 * 4, 256, 0, stream>>>(m1, m2, out, cap_words, h1, h2, idh_cap, lines1, lines2, count, id_mismatch);
    return cudaGetLastError();
}
cudaError_t launch_combine_interleaved(const uint32_t* m, uint32_t* out, uint64_t cap_out_words, const unsigned long long* h, uint64_t idh_cap,
                                       const unsigned long long* lines, unsigned long long* count, unsigned long long* id_mismatch, cudaStream_t stream) {
    combine_interleaved_kernel<<<148 * 4, 256, 0, stream>>>(m, out, cap_out_words, h, idh_cap, lines, count, id_mismatch);
    return cudaGetLastError();
}

}  // namespace fqg
