// fastq_cut.cc -- host batch cutter for the streaming path: prefixes of whole records, mates in lock step.
//
// The reference's reader thread parses every record to build its batches (seq_io RecordSet via
// InterleavedFastqReader::fill_data, /root/reference/src/lib/seq_io.rs:60-80, and PairedFastqReader::fill_data, which
// reads "the same number of records" from R2 as from R1, :149-198).  Here the GPU parses, so the host must not touch
// every byte: a cut is found by looking at a few records next to the wanted position, and R1/R2 are kept in lock step by
// matching the READ ID of the first record after the cut (mates carry the same id, src/main.rs:741-747 asserts it).
// That is a guess -- the device checks every batch (equal record counts, equal ids, whole records), and equal counts in
// every batch imply equal counts at every cut -- so a wrong guess surfaces as an error on exactly that batch and the
// caller repeats the cut with exact=1: record counting at memory speed on all host threads.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/fqgrep_b200.h"
#include "fastq_host.hpp"

namespace fqg {

namespace {

__attribute__((target("avx2"))) uint64_t count_newlines_avx2(const uint8_t* p, size_t n) {
    uint64_t total = 0;
    const __m256i nl = _mm256_set1_epi8('\n');
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const uint32_t a = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i)), nl));
        const uint32_t b = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i + 32)), nl));
        const uint32_t c = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i + 64)), nl));
        const uint32_t d = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i + 96)), nl));
        total += (uint64_t)__builtin_popcountll(((uint64_t)a << 32) | b) + (uint64_t)__builtin_popcountll(((uint64_t)c << 32) | d);
    }
    for (; i < n; i++) total += p[i] == '\n';
    return total;
}
uint64_t count_newlines_plain(const uint8_t* p, size_t n) {
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += p[i] == '\n';
    return total;
}
uint64_t count_newlines_1(const uint8_t* p, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    return avx2 ? count_newlines_avx2(p, n) : count_newlines_plain(p, n);
}

constexpr size_t kBlock = 1u << 20;

// newline counts of consecutive 1 MiB blocks, on all host threads
std::vector<uint64_t> block_newlines(const uint8_t* p, size_t n) {
    const size_t nb = (n + kBlock - 1) / kBlock;
    std::vector<uint64_t> cnt(nb, 0);
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 32) nt = 32;
    if ((size_t)nt > nb) nt = (unsigned)(nb ? nb : 1);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= nb) break;
            const size_t lo = b * kBlock, hi = std::min(n, lo + kBlock);
            cnt[b] = count_newlines_1(p + lo, hi - lo);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return cnt;
}

inline const uint8_t* eol(const uint8_t* p, const uint8_t* e) {
    const void* q = memchr(p, '\n', (size_t)(e - p));
    return q ? (const uint8_t*)q : e;
}
// id of the record that starts at buf[pos] ('@' excluded): up to the first space or the end of the line, CR dropped
void id_at(const uint8_t* buf, size_t n, size_t pos, const uint8_t** id, size_t* len) {
    const uint8_t *e = buf + n, *p = buf + pos + 1;
    if (p > e) p = e;
    const uint8_t* l = eol(p, e);
    size_t k = (size_t)(l - p);
    if (k && p[k - 1] == '\r') k--;
    const void* sp = memchr(p, ' ', k);
    *id = p;
    *len = sp ? (size_t)((const uint8_t*)sp - p) : k;
}
bool same_id(const uint8_t* a, size_t an, size_t ap, const uint8_t* b, size_t bn, size_t bp) {
    const uint8_t *ia, *ib;
    size_t la, lb;
    id_at(a, an, ap, &ia, &la);
    id_at(b, bn, bp, &ib, &lb);
    return la == lb && std::memcmp(ia, ib, la) == 0;
}

// Record starts in [from, to]: walks records forward from the first start at or after `from`.  Stops after `to`.
// The end of the buffer counts as a start only when the last record is certainly whole: it ends with its newline, or
// the input ends here (`final`).  Returns false when nothing there parses (the device reports the input's real error).
bool starts_between(const uint8_t* buf, size_t n, size_t from, size_t to, bool final, std::vector<size_t>* out) {
    out->clear();
    size_t pos = find_record_start(buf, n, from);
    HostRecord r;
    std::string err;
    while (pos <= to && pos < n) {
        out->push_back(pos);
        const int rc = next_host_record(buf, n, &pos, &r, &err);
        if (rc < 0) return true;                     // a partial record at the end: its start is the last cut
        if (rc == 0) { pos = n; break; }
    }
    if (pos >= n && n <= to && n > 0 && (final || buf[n - 1] == '\n')) out->push_back(n);
    return !out->empty();
}

// last record start <= limit (n itself counts, see above), looking back over growing windows for long records
size_t last_start_at_or_before(const uint8_t* buf, size_t n, size_t limit, bool final) {
    if (limit >= n) limit = n;
    std::vector<size_t> st;
    for (size_t back = 1u << 16;; back <<= 2) {
        const size_t from = limit > back ? limit - back : 0;
        if (starts_between(buf, n, from, limit, final, &st)) {
            size_t best = 0;
            bool any = false;
            for (size_t s : st) if (s <= limit) { best = s; any = true; }
            if (any) return best;
        }
        if (from == 0) return 0;
    }
}

// end of the record that starts at pos, or 0 when it is not whole inside buf[0..n) (same "whole" rule as above)
size_t record_end(const uint8_t* buf, size_t n, size_t pos, bool final) {
    HostRecord r;
    std::string err;
    size_t p = pos;
    if (next_host_record(buf, n, &p, &r, &err) != 1) return 0;
    if (p >= n && !final && buf[n - 1] != '\n') return 0;
    return p;
}

}  // namespace

uint64_t count_newlines(const uint8_t* p, size_t n) {
    uint64_t total = 0;
    for (uint64_t c : block_newlines(p, n)) total += c;
    return total;
}

// offset just past the k-th newline (k >= 1) of p[0..n), or n + 1 when there are fewer
size_t offset_after_newline(const uint8_t* p, size_t n, uint64_t k) {
    if (k == 0) return 0;
    const std::vector<uint64_t> cnt = block_newlines(p, n);
    uint64_t seen = 0;
    for (size_t b = 0; b < cnt.size(); b++) {
        if (seen + cnt[b] >= k) {
            const size_t lo = b * kBlock, hi = std::min(n, lo + kBlock);
            for (size_t i = lo; i < hi; i++)
                if (p[i] == '\n' && ++seen == k) return i + 1;
        }
        seen += cnt[b];
    }
    return n + 1;
}

// records in the prefix buf[0..c) that ends on a record end (4 lines each; the last line may lack its newline at the end of input)
static uint64_t records_in_prefix(const uint8_t* buf, size_t c) {
    if (c == 0) return 0;
    return (count_newlines(buf, c) + (buf[c - 1] != '\n' ? 1 : 0)) / 4;
}
// offset just past record k (1-based count) of buf[0..n), n + 1 when there are fewer whole records
static size_t offset_after_records(const uint8_t* buf, size_t n, uint64_t k, bool final) {
    if (k == 0) return 0;
    const size_t o = offset_after_newline(buf, n, 4 * k);
    if (o <= n) return o;
    if (final && n && buf[n - 1] != '\n' && offset_after_newline(buf, n, 4 * k - 1) <= n && records_in_prefix(buf, n) == k) return n;
    return n + 1;
}

int fastq_cut(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, size_t limit, bool final, bool exact, size_t* cut1, size_t* cut2) {
    *cut1 = 0;
    *cut2 = 0;
    if (limit == 0) limit = ~(size_t)0;
    if (mode == FQG_SINGLE) {
        *cut1 = (final && n1 <= limit) ? n1 : last_start_at_or_before(r1, n1, std::min(limit, n1), final);
        return 0;
    }
    if (mode == FQG_INTERLEAVED) {
        if (final && n1 <= limit) { *cut1 = n1; return 0; }
        size_t c = last_start_at_or_before(r1, n1, std::min(limit, n1), final);
        if (exact) {
            // even record count in front of the cut, counted: a cut is a record start, so lines / 4 records precede it
            while (c > 0 && (records_in_prefix(r1, c) & 1)) c = last_start_at_or_before(r1, n1, c - 1, final);
            *cut1 = c;
            return 0;
        }
        // by ids: the mates of a pair carry the same id, the records on either side of a cut between two pairs do not
        for (int guard = 0; guard < 4 && c > 0 && c < n1; guard++) {
            const size_t prev = last_start_at_or_before(r1, n1, c - 1, final);
            if (prev < c && same_id(r1, n1, prev, r1, n1, c)) { c = prev; continue; }      // c splits a pair: step back one record
            break;
        }
        if (c >= n1 && c > 0) {
            // the buffer ends on a record end: whole pairs only if its last two records are mates
            const size_t last = last_start_at_or_before(r1, n1, c - 1, final);
            const size_t before = last > 0 ? last_start_at_or_before(r1, n1, last - 1, final) : last;
            if (!(before < last && same_id(r1, n1, before, r1, n1, last))) c = last;
        }
        *cut1 = c;
        return 0;
    }
    // FQG_PAIRED
    if (final && n1 <= limit && n2 <= limit) { *cut1 = n1; *cut2 = n2; return 0; }
    size_t lim1 = std::min(limit, n1);
    const size_t lim2 = std::min(limit, n2);
    for (int attempt = 0; attempt < 32; attempt++) {
        const size_t c1 = last_start_at_or_before(r1, n1, lim1, final);
        if (c1 == 0) return 0;
        size_t c2 = 0;
        if (exact) {
            c2 = offset_after_records(r2, n2, records_in_prefix(r1, c1), final);
            if (c2 == n2 + 1) {
                // R2 holds fewer whole records than this prefix of R1: hand out the records R2 has (if R1 then has records
                // left at the very end the device reports "out of sync" on the last batch)
                const size_t e2 = last_start_at_or_before(r2, n2, lim2, final);
                const size_t e1 = offset_after_records(r1, n1, records_in_prefix(r2, e2), final);
                *cut1 = e1 > n1 ? 0 : e1;
                *cut2 = e1 > n1 ? 0 : e2;
                return 0;
            }
        } else {
            // anchor on the LAST record of the R1 prefix: its mate sits near the same offset in R2 (same ids, same read
            // lengths in practice); the R2 prefix ends where that mate ends
            const size_t p1 = last_start_at_or_before(r1, n1, c1 - 1, final);
            size_t s2 = n2 + 1;
            std::vector<size_t> st;
            const size_t guess = std::min(p1, n2);
            for (size_t win = 1u << 14; win <= (1u << 24) && s2 == n2 + 1; win <<= 3) {
                const size_t from = guess > win ? guess - win : 0;
                if (!starts_between(r2, n2, from, std::min(n2, guess + win), final, &st)) continue;
                for (size_t s : st)
                    if (s < n2 && same_id(r1, n1, p1, r2, n2, s)) { s2 = s; break; }
                if (from == 0 && guess + win >= n2) break;
            }
            if (s2 == n2 + 1) return fastq_cut(mode, r1, n1, r2, n2, limit, final, true, cut1, cut2);      // ids did not help: count
            c2 = record_end(r2, n2, s2, final);
            if (c2 == 0) {              // the mate is not whole in R2's buffer yet: leave this pair for the next batch
                if (p1 == 0) return 0;
                lim1 = p1;
                continue;
            }
        }
        if (c2 <= lim2) { *cut1 = c1; *cut2 = c2; return 0; }
        // R2's records are longer: its prefix does not fit.  Scale the R1 prefix down by the overshoot and try again
        // (a few rounds land just under the limit; the last ones step back about a record at a time).
        size_t scaled = (size_t)((double)c1 * (double)lim2 / (double)c2);
        if (scaled >= c1) scaled = c1 - 1;
        if (scaled == 0) return 0;
        lim1 = scaled;
    }
    return fastq_cut(mode, r1, n1, r2, n2, limit, final, true, cut1, cut2);
}

}  // namespace fqg
