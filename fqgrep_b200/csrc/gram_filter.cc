// gram_filter.cc -- large literal sets (-F -f with thousands of patterns): q-gram sampling filter + exact verification.
//
// Replaces, for pattern sets whose Aho-Corasick automaton does not fit shared memory, the per-byte automaton walk of
// aho_corasick::AhoCorasick::is_match (/root/reference/src/lib/matcher.rs:250, built at :297-298).  The answer is still
// exact: the filter only decides WHERE to verify.
//
// Every pattern is at least `min_len` long.  Pick a gram length q and a stride s = min_len - q + 1.  Any occurrence of a
// pattern covers >= s consecutive q-gram windows, so probing the text only at window ends q-1, q-1+s, q-1+2s, ... always
// lands on a window that starts at pattern offset 0..s-1.  The device therefore keeps
//   * a blocked Bloom filter (shared memory, 3 bits per gram inside one 32-bit word) of the grams at offsets 0..s-1
//     of every pattern: exact negative filter;
//   * an open-addressing table gram -> (offset in the pattern, pattern bytes) in L2 for the rare positives: one 16-byte
//     slot read decides a false positive; a real gram is verified against the pattern 16 bytes at a time.
// Grams are built from 2-bit codes (byte >> 1) & 3 of ANY byte (not only ACGT): text and patterns use the same map, so
// aliasing (N ~ G ...) can only add candidates, never lose a match.
#include "automata.hpp"

#include <algorithm>
#include <cstdlib>

namespace fqg {

uint32_t gram_code(const uint8_t* p, uint32_t q) {
    uint32_t g = 0;
    for (uint32_t i = 0; i < q; i++) g = (g << 2) | ((p[i] >> 1) & 3u);
    return g;
}

// Key of the q bases p[0..q) when the device samples grams that end on a 4-byte boundary of its staging buffer
// (stride a multiple of 4).  The device packs every aligned 32-bit word of text into one byte with a single multiply
// -- (w & 0x06060606) * 0x00820820 >> 24: byte k of the word lands in bits 2k..2k+1 -- and shifts those bytes
// through a register, newest word in the low byte; the key is that register masked to the last q bases.
// Base p[q-1-i] therefore sits in byte i/4 at bits 2*(3 - i%4).
static uint32_t gram_bit_offset(uint32_t i) { return 8u * (i / 4u) + 2u * (3u - (i % 4u)); }
uint32_t gram_key_aligned(const uint8_t* p, uint32_t q) {
    uint32_t g = 0;
    for (uint32_t i = 0; i < q; i++) g |= (uint32_t)((p[q - 1 - i] >> 1) & 3u) << gram_bit_offset(i);
    return g;
}

bool build_gram_filter(const std::vector<std::string>& patterns_in, bool with_revcomp, GramFilter* out) {
    const uint8_t* comp = complement_table();
    std::vector<std::string> pats = patterns_in;
    if (with_revcomp)
        for (auto& p : patterns_in) {
            std::string r(p.rbegin(), p.rend());
            for (auto& ch : r) ch = (char)comp[(uint8_t)ch];
            pats.push_back(std::move(r));
        }
    out->usable = false;
    if (pats.empty()) return false;
    size_t min_len = SIZE_MAX, total = 0;
    for (auto& p : pats) { min_len = std::min(min_len, p.size()); total += p.size(); }
    if (min_len < 8 || pats.size() >= (1u << 26) || total + 16 * pats.size() >= (1ull << 31)) return false;
    uint32_t q = (uint32_t)std::min<size_t>(16, std::max<size_t>(8, min_len / 2 + 2));
    uint32_t s = (uint32_t)min_len - q + 1;
    if (s > 32) { s = 32; }            // payload keeps the offset in 5 bits; a smaller stride is always valid
    const bool aligned = s >= 4;
    if (aligned) {                     // aligned sampling: one probe per s bytes of staged text, read with aligned
        s = s >= 8 ? (s & ~7u) : 4u;   // 8-byte (or 4-byte) shared loads and no sub-word shifts; the slack goes back
        q = (uint32_t)std::min<size_t>(16, min_len - s + 1);      // into a longer, more selective gram
    }
    out->q = q;
    out->s = s;
    out->min_len = (uint32_t)min_len;
    out->qmask = q == 16 ? 0xFFFFFFFFu : ((1u << (2 * q)) - 1u);
    if (aligned) {
        out->qmask = 0;
        for (uint32_t i = 0; i < q; i++) out->qmask |= 3u << gram_bit_offset(i);
    }
    const uint64_t n_grams = (uint64_t)pats.size() * s;
    if (n_grams > (1ull << 22)) return false;
    // The filter only pays when a random text gram is rarely one of ours: with n_grams of the 4^q possible grams in the
    // table, a 150-base read has ~(150 / s) * n_grams / 4^q real gram hits to verify.  Short patterns in large numbers
    // (thousands of 10-mers: q = 8) would send every read to verification several times; those sets keep the automaton
    // (measured, 1000 x 10-mers: filter 0.17 of roofline, class-indexed table 0.29).  api.cu is stricter (0.1) when the
    // automaton is small enough for the 2-base-stride table, which beats a busy filter (60 x 15-mer regexes: 0.62 vs 0.54).
    out->expected_hits = q < 16 ? (double)n_grams / (double)(1ull << (2 * q)) * (150.0 / s) : 0.0;
    if (out->expected_hits > 0.5 && !std::getenv("FQG_GRAM_NO_DENSITY_GUARD")) return false;      // (knob for measurements)
    // blocked 3-bit Bloom filter in shared memory: the top bits of one multiplicative hash pick a 32-bit word, bits 2..16
    // three bits inside it.  ~40 bits per gram keeps the false-positive rate near 0.2 %; floor 8 KB, cap 128 KB (2^20
    // bits).  The filter shares the SM's 227 KB with the staging slots, i.e. with the number of resident warps, so the
    // last doubling (64 -> 128 KB) is only taken when 64 KB would leave less than ~6 bits per gram (measured on B200,
    // 10 k 20-mers: 80 k grams run 5 % faster with 64 KB and more warps, 160 k grams 28 % faster with 128 KB).
    uint32_t bits_log2 = 16, cap_log2 = n_grams * 6 <= (1ull << 19) ? 19 : 20;
    if (const char* e = std::getenv("FQG_GRAM_BITS_LOG2_MAX")) {      // tuning knob for experiments
        int v = std::atoi(e);
        if (v >= 16 && v <= 20) cap_log2 = (uint32_t)v;
    }
    while (bits_log2 < cap_log2 && (1ull << bits_log2) < n_grams * 40) bits_log2++;
    out->bitmap_bits_log2 = bits_log2;
    out->bitmap.assign((1u << bits_log2) / 32, 0);
    uint32_t slots_log2 = 4;
    while ((1ull << slots_log2) < n_grams * 4 + 2) slots_log2++;        // load <= 0.25: a miss ends after ~1.2 slots
    out->slots_log2 = slots_log2;
    out->slots.assign((size_t)4 << slots_log2, 0xFFFFFFFFu);      // {gram, pool offset, length << 5 | j, -}; all ones = empty
    out->pool.clear();
    for (size_t id = 0; id < pats.size(); id++) {
        const std::string& p = pats[id];
        const uint32_t pool_off = (uint32_t)out->pool.size();      // 16-byte aligned: the device compares uint4 chunks
        for (uint32_t j = 0; j < s; j++) {
            const uint32_t g = aligned ? gram_key_aligned((const uint8_t*)p.data() + j, q) : gram_code((const uint8_t*)p.data() + j, q);
            const uint32_t h = g * 0x9E3779B1u;
            const uint32_t word = h >> (32 - (bits_log2 - 5));
            out->bitmap[word] |= (1u << ((h >> 2) & 31u)) | (1u << ((h >> 7) & 31u)) | (1u << ((h >> 12) & 31u));
            uint32_t i = (g * 0x85EBCA6Bu) >> (32 - slots_log2);
            while (out->slots[4 * (size_t)i + 2] != 0xFFFFFFFFu) i = (i + 1) & ((1u << slots_log2) - 1);
            out->slots[4 * (size_t)i] = g;
            out->slots[4 * (size_t)i + 1] = pool_off;
            out->slots[4 * (size_t)i + 2] = (uint32_t)(p.size() << 5) | j;
        }
        out->pool.insert(out->pool.end(), p.begin(), p.end());
        out->pool.resize((out->pool.size() + 15) & ~(size_t)15, 0);
    }
    out->pool.resize(out->pool.size() + 16, 0);
    out->usable = true;
    return true;
}

}  // namespace fqg
