// fastq_host.cc -- ordered write-back of the records the device selected.
//
// Replaces the `func` closures of the reference (src/main.rs:561-575, 600-614, 641-657) and write_paired_record
// (:682-727): for every selected unit, in input order, emit seq_io's write_to form "@head\nseq\n+\nqual\n"
// (so CRLF input and "+name" separator lines are normalised exactly like the reference), R1 then R2 for pairs.
// This is host I/O glue driven by the device bitmask; no matching happens here.
#include "fastq_host.hpp"

#include <cstring>
#include <cuda_runtime.h>

#include "../../include/fqgrep_b200.h"

namespace fqg {

void fastq_state_free(FastqDeviceState* s) {
    for (auto& p : s->scratch) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    for (auto& c : s->cap) c = 0;
}

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
}  // namespace

int write_selected_host(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* mask, uint8_t* out, size_t* out_len,
                        uint8_t* out2, size_t* out2_len, std::string* err) {
    if (!out || !out_len || (!mask)) { *err = "NULL argument"; return FQG_E_INVALID_ARG; }
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

}  // namespace fqg
