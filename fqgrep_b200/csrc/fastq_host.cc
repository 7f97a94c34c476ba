// fastq_host.cc -- ordered write-back of the records the device selected.
//
// Replaces the `func` closures of the reference (src/main.rs:561-575, 600-614, 641-657) and write_paired_record
// (:682-727): for every selected unit, in input order, emit seq_io's write_to form "@head\nseq\n+\nqual\n"
// (so CRLF input and "+name" separator lines are normalised exactly like the reference), R1 then R2 for pairs.
// This is host I/O glue driven by the device bitmask; no matching happens here.
#include "fastq_host.hpp"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/fqgrep_b200.h"

namespace fqg {

namespace {
inline const uint8_t* eol(const uint8_t* p, const uint8_t* e) {
    const void* q = memchr(p, '\n', (size_t)(e - p));
    return q ? (const uint8_t*)q : e;
}
inline size_t strip_cr(const uint8_t* p, size_t n) { return n && p[n - 1] == '\r' ? n - 1 : n; }
}  // namespace

int next_host_record(const uint8_t* buf, size_t n, size_t* pos, HostRecord* r, std::string* err) {
    const uint8_t *e = buf + n, *p = buf + *pos;
    if (p >= e) return 0;
    {
        const uint8_t* q = p;
        while (q < e && (*q == '\n' || *q == '\r')) q++;
        if (q == e) return 0;
    }
    if (*p != '@') { *err = "FASTQ parse error: expected '@' at record start (InvalidStart)"; return FQG_E_FASTQ; }
    const uint8_t* l1 = eol(p, e);
    const uint8_t* s = l1 + 1;
    const uint8_t* l2 = l1 < e ? eol(s, e) : e;
    const uint8_t* t = l2 + 1;
    if (l1 >= e || l2 >= e || t >= e) { *err = "FASTQ parse error: truncated record (UnexpectedEnd)"; return FQG_E_FASTQ; }
    if (*t != '+') { *err = "FASTQ parse error: expected '+' separator (InvalidSep)"; return FQG_E_FASTQ; }
    const uint8_t* l3 = eol(t, e);
    if (l3 >= e) { *err = "FASTQ parse error: truncated record (UnexpectedEnd)"; return FQG_E_FASTQ; }
    const uint8_t* q = l3 + 1;
    const uint8_t* l4 = eol(q, e);
    r->head = p + 1; r->head_len = strip_cr(p + 1, (size_t)(l1 - (p + 1)));
    r->seq = s; r->seq_len = strip_cr(s, (size_t)(l2 - s));
    r->qual = q; r->qual_len = strip_cr(q, (size_t)(l4 - q));
    if (r->seq_len != r->qual_len) { *err = "FASTQ parse error: sequence and quality lengths differ (UnequalLengths)"; return FQG_E_FASTQ; }
    *pos = l4 == e ? n : (size_t)(l4 + 1 - buf);
    return 1;
}

// Shard cutting for multi-GPU runs: the first record start at or after `pos`.  A FASTQ quality line may begin with
// '@', so a candidate line must start with '@' AND be followed two lines later by a '+' line (a sequence line never
// starts with '+'), with equal sequence/quality lengths.  Returns n when no record starts after pos.
size_t find_record_start(const uint8_t* buf, size_t n, size_t pos) {
    if (pos == 0) return 0;
    if (pos >= n) return n;
    const uint8_t* e = buf + n;
    const uint8_t* p = buf + pos;
    if (p[-1] != '\n') { p = eol(p, e); if (p >= e) return n; p++; }
    for (int guard = 0; guard < 64 && p < e; guard++) {
        if (*p == '@') {
            const uint8_t* l1 = eol(p, e);
            const uint8_t* l2 = l1 < e ? eol(l1 + 1, e) : e;
            if (l2 < e && l2 + 1 < e && l2[1] == '+') {
                const uint8_t* l3 = eol(l2 + 1, e);
                const uint8_t* l4 = l3 < e ? eol(l3 + 1, e) : e;
                if (l3 < e && strip_cr(l1 + 1, (size_t)(l2 - l1 - 1)) == strip_cr(l3 + 1, (size_t)(l4 - l3 - 1))) return (size_t)(p - buf);
            }
        }
        p = eol(p, e);
        if (p >= e) return n;
        p++;
    }
    return n;
}

namespace {
inline uint8_t* emit(uint8_t* o, const HostRecord& r) {
    *o++ = '@';
    memcpy(o, r.head, r.head_len); o += r.head_len;
    *o++ = '\n';
    memcpy(o, r.seq, r.seq_len); o += r.seq_len;
    *o++ = '\n'; *o++ = '+'; *o++ = '\n';
    memcpy(o, r.qual, r.qual_len); o += r.qual_len;
    *o++ = '\n';
    return o;
}

// ---- parallel write-back ---------------------------------------------------------------------------------------------
// The serial walk above parses every record to find the selected ones: ~memchr speed on one core, i.e. seconds for
// the tens of gigabytes the device matches in one.  The parallel form cuts each input into shards at record starts
// (find_record_start), walks every shard once noting where its records start (pass 1, 4 bytes per record), and with the
// record number of each shard's first record known re-reads and emits only the selected records (pass 2).  It is only trusted when it provably equals the
// serial parse: every shard's walk must end EXACTLY on the next shard's first byte and no walk may hit a malformed
// record; otherwise -- and for small inputs -- the serial walk runs and reports whatever it finds.
struct Span { uint64_t unit; size_t off, len; };
struct ShardOut {
    std::vector<uint8_t> bytes;     // selected records of the shard, write_to form
    std::vector<Span> spans;        // one per selected record: its unit and where it sits in `bytes`
};

size_t write_shard_bytes() {
    if (const char* e = std::getenv("FQG_WRITE_SHARD_BYTES")) {      // test hook: exercise the parallel form on small inputs
        const long v = std::atol(e);
        if (v > 0) return (size_t)v;
    }
    return (size_t)8 << 20;
}

// Parses one stream in parallel.  unit_of(record) maps a record number to its unit; on success fills `shards` (in input
// order) and *n_records.  Returns false when the parallel parse cannot be trusted.
bool parallel_select(const uint8_t* buf, size_t n, const uint32_t* mask, bool interleaved, std::vector<ShardOut>* shards, uint64_t* n_records) {
    const size_t shard_bytes = write_shard_bytes();
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 32) nt = 32;
    if (n < 2 * shard_bytes || nt < 2) return false;
    size_t n_shards = n / shard_bytes;
    if (n_shards > (size_t)nt * 8) n_shards = (size_t)nt * 8;
    if (n_shards < (n >> 31) + 1) n_shards = (n >> 31) + 1;      // record starts are kept as 32-bit offsets into their shard
    std::vector<size_t> cut(n_shards + 1);
    cut[0] = 0;
    for (size_t i = 1; i < n_shards; i++) {
        cut[i] = find_record_start(buf, n, (size_t)((unsigned __int128)n * i / n_shards));
        if (cut[i] < cut[i - 1]) cut[i] = cut[i - 1];
    }
    cut[n_shards] = n;
    const bool trace = std::getenv("FQG_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (trace) std::fprintf(stderr, "[fqg] write_selected %s: %.1f ms since start (%zu shards, %u threads)\n", what,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(), n_shards, nt);
    };
    std::vector<std::vector<uint32_t>> starts(n_shards);
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    auto run = [&](auto&& body) {
        next = 0;
        std::vector<std::thread> pool;
        auto work = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= n_shards || !ok) break;
                body(i);
            }
        };
        for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    };
    // pass 1: record starts of every shard; every walk must land exactly on the next cut (the last one on a clean end of input)
    run([&](size_t i) {
        if (cut[i + 1] - cut[i] > 0xFFFFFFFFull) { ok = false; return; }
        size_t pos = cut[i];
        HostRecord r;
        std::string err;
        std::vector<uint32_t> st;      // local: the headers of starts[] share cache lines between threads
        st.reserve((cut[i + 1] - cut[i]) / 256 + 16);
        while (pos < cut[i + 1]) {
            const size_t at = pos;
            const int rc = next_host_record(buf, n, &pos, &r, &err);
            if (rc < 0) { ok = false; return; }
            if (rc == 0) { pos = n; break; }          // only blank lines left
            st.push_back((uint32_t)(at - cut[i]));
        }
        if (pos != cut[i + 1] && !(i + 1 == n_shards && pos >= n)) { ok = false; return; }
        starts[i] = std::move(st);
    });
    if (!ok) return false;
    lap("pass 1");
    std::vector<uint64_t> first(n_shards + 1, 0);
    for (size_t i = 0; i < n_shards; i++) first[i + 1] = first[i] + starts[i].size();
    if (interleaved && (first[n_shards] & 1)) return false;      // odd record count: the serial walk reports it
    shards->assign(n_shards, ShardOut());
    // pass 2: re-read and emit only the selected records of every shard
    run([&](size_t i) {
        ShardOut so;                   // local for the same reason
        HostRecord r;
        std::string err;
        for (uint64_t rec = first[i]; rec < first[i + 1]; rec++) {
            const uint64_t unit = interleaved ? rec >> 1 : rec;
            if (!((mask[unit >> 5] >> (unit & 31)) & 1u)) continue;
            size_t pos = cut[i] + starts[i][rec - first[i]];
            if (next_host_record(buf, n, &pos, &r, &err) != 1) { ok = false; return; }
            const size_t len = r.head_len + r.seq_len + r.qual_len + 6;
            const size_t off = so.bytes.size();
            so.bytes.resize(off + len);
            emit(so.bytes.data() + off, r);
            so.spans.push_back(Span{unit, off, len});
        }
        (*shards)[i] = std::move(so);
    });
    if (!ok) return false;
    lap("pass 2");
    *n_records = first[n_shards];
    return true;
}

bool write_selected_parallel(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* mask, uint8_t* out, size_t* out_len,
                             uint8_t* out2, size_t* out2_len) {
    std::vector<ShardOut> s1, s2;
    uint64_t c1 = 0, c2 = 0;
    if (!parallel_select(r1, n1, mask, mode == FQG_INTERLEAVED, &s1, &c1)) return false;
    if (mode != FQG_PAIRED) {
        uint8_t* o = out;
        for (auto& so : s1) { if (!so.bytes.empty()) memcpy(o, so.bytes.data(), so.bytes.size()); o += so.bytes.size(); }
        *out_len = (size_t)(o - out);
        if (out2_len) *out2_len = 0;
        return true;
    }
    if (!parallel_select(r2, n2, mask, false, &s2, &c2) || c2 < c1) return false;      // fewer R2 records: serial walk reports it
    // R2 may hold MORE records than R1 (the serial walk stops with R1): units >= c1 never count
    uint8_t *o = out, *o2 = out2;
    size_t j = 0, k = 0;      // cursor into s2
    for (auto& so : s1)
        for (auto& sp : so.spans) {
            memcpy(o, so.bytes.data() + sp.off, sp.len); o += sp.len;
            while (j < s2.size() && k >= s2[j].spans.size()) { j++; k = 0; }
            if (j >= s2.size() || s2[j].spans[k].unit != sp.unit) return false;      // cannot happen: same mask, same units
            const Span& b = s2[j].spans[k++];
            if (o2) { memcpy(o2, s2[j].bytes.data() + b.off, b.len); o2 += b.len; }
            else { memcpy(o, s2[j].bytes.data() + b.off, b.len); o += b.len; }
        }
    *out_len = (size_t)(o - out);
    if (out2_len) *out2_len = out2 ? (size_t)(o2 - out2) : 0;
    return true;
}
}  // namespace

int write_selected_host(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* mask, uint8_t* out, size_t* out_len,
                        uint8_t* out2, size_t* out2_len, std::string* err) {
    if (!out || !out_len || (!mask)) { *err = "NULL argument"; return FQG_E_INVALID_ARG; }
    const bool trace = std::getenv("FQG_TRACE") != nullptr;
    if (write_selected_parallel(mode, r1, n1, r2, n2, mask, out, out_len, out2, out2_len)) {
        if (trace) std::fprintf(stderr, "[fqg] write_selected: parallel walk\n");
        return 0;
    }
    if (trace) std::fprintf(stderr, "[fqg] write_selected: serial walk\n");
    uint8_t *o = out, *o2 = out2;
    size_t p1 = 0, p2 = 0;
    uint64_t unit = 0;
    HostRecord a, b;
    for (;;) {
        int rc = next_host_record(r1, n1, &p1, &a, err);
        if (rc < 0) return rc;
        if (rc == 0) break;
        bool have_b = false;
        if (mode == FQG_INTERLEAVED) {
            rc = next_host_record(r1, n1, &p1, &b, err);
            if (rc < 0) return rc;
            if (rc == 0) { *err = "FASTQ file does not have an even number of records."; return FQG_E_FASTQ; }
            have_b = true;
        } else if (mode == FQG_PAIRED) {
            rc = next_host_record(r2, n2, &p2, &b, err);
            if (rc < 0) return rc;
            if (rc == 0) { *err = "FASTQ files out of sync: R2 has fewer records than R1"; return FQG_E_PAIR_MISMATCH; }
            have_b = true;
        }
        if ((mask[unit >> 5] >> (unit & 31)) & 1u) {
            o = emit(o, a);
            if (have_b) { if (o2) o2 = emit(o2, b); else o = emit(o, b); }
        }
        unit++;
    }
    *out_len = (size_t)(o - out);
    if (out2_len) *out2_len = out2 ? (size_t)(o2 - out2) : 0;
    return 0;
}

// ---- write-back from device spans (K5) -----------------------------------------------------------------------------------
// The fused kernel already knows where every record starts and ends and whether its bytes are in write_to form, so the
// selected records arrive as (offset, length, flags) in output order and the host's share of src/main.rs:641-657 /
// write_paired_record (:682-727) is a gather: byte copies, and re-formatting only for records with CR line ends, a
// "+name" separator line or no final newline.  Output positions are a prefix sum of the span lengths, so large
// batches are copied by all host threads at once.
int write_spans_host(const fqg_span* spans, uint64_t n_spans, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, uint8_t* out, size_t out_cap,
                     size_t* out_len, uint8_t* out2, size_t out2_cap, size_t* out2_len, std::string* err) {
    if (!out_len || (n_spans && (!spans || !out))) { *err = "NULL argument"; return FQG_E_INVALID_ARG; }
    *out_len = 0;
    if (out2_len) *out2_len = 0;
    if (n_spans == 0) return 0;
    auto source = [&](const fqg_span& sp, const uint8_t** buf, size_t* n) { const bool second = (sp.flags & FQG_SPAN_R2) != 0; *buf = second ? r2 : r1; *n = second ? n2 : n1; };
    // pass 1: where every span lands
    std::vector<size_t> at(n_spans);
    size_t o1 = 0, o2 = 0;
    for (uint64_t i = 0; i < n_spans; i++) {
        const fqg_span& sp = spans[i];
        const uint8_t* buf;
        size_t n;
        source(sp, &buf, &n);
        if (!buf || sp.offset > n || sp.len > n - sp.offset) { *err = "span outside its input"; return FQG_E_INVALID_ARG; }
        size_t len = sp.len;
        if (sp.flags & FQG_SPAN_NORMALISE) {
            HostRecord r;
            size_t pos = 0;
            if (next_host_record(buf + sp.offset, sp.len, &pos, &r, err) != 1) return FQG_E_FASTQ;
            len = r.head_len + r.seq_len + r.qual_len + 6;
        }
        size_t& o = (out2 && (sp.flags & FQG_SPAN_R2)) ? o2 : o1;
        at[i] = o;
        o += len;
    }
    if (o1 > out_cap || o2 > out2_cap) { *err = "output buffer too small for the selected records"; return FQG_E_TOO_LARGE; }
    // pass 2: copy / re-format
    auto emit_range = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const fqg_span& sp = spans[i];
            const uint8_t* buf;
            size_t n;
            source(sp, &buf, &n);
            uint8_t* dst = ((out2 && (sp.flags & FQG_SPAN_R2)) ? out2 : out) + at[i];
            if (sp.flags & FQG_SPAN_NORMALISE) {
                HostRecord r;
                size_t pos = 0;
                std::string e;
                if (next_host_record(buf + sp.offset, sp.len, &pos, &r, &e) == 1) emit(dst, r);
            } else {
                memcpy(dst, buf + sp.offset, sp.len);
            }
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 16) nt = 16;
    if (n_spans < 65536 || nt < 2) emit_range(0, n_spans);
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++) pool.emplace_back(emit_range, n_spans * t / nt, n_spans * (t + 1) / nt);
        for (auto& th : pool) th.join();
    }
    *out_len = o1;
    if (out2_len) *out2_len = o2;
    return 0;
}

}  // namespace fqg
