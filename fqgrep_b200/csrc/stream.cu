// stream.cu -- the streaming batcher (fqg_stream_*) and the whole-input entry points built on it.
//
// Replaces the reference's three-stage pipeline -- reader thread (spawn_reader, /root/reference/src/main.rs:61-87;
// PairedFastqReader::fill_data lock-step batches, src/lib/seq_io.rs:149-198), worker pool (:314-326) and the ordered replay on
// the main thread (OrderedReader + BTreeMap reorder, :253-282, 327-364) -- with a ring of batches in flight on a GPU:
// while batch k runs its kernels, batch k+1 is being copied in and batch k-1 is being written back by the host.
// Batches complete in submission order, so input order needs no reorder buffer.  Device and pinned memory are bounded
// by the ring (slots x slot_bytes per input stream) whatever the input size.
//
// fqg_grep_fastq_host / _write / _multi cut ONE host-resident input into such batches (fqg_fastq_cut: a few records
// inspected per cut, mates aligned by read id; every batch is verified on the device, and a batch that fails
// verification sends the rest of the input through the exact, counting cutter) and deal them to one ring per GPU.
#include <cstring>
#include <deque>

#include "internal.hpp"

using namespace fqg;

namespace {
struct Slot {
    FastqLane* lane = nullptr;
    uint8_t* pin[2] = {nullptr, nullptr};      // pinned slot buffers, allocated on first fqg_stream_acquire
    bool in_flight = false;
    const uint8_t* r[2] = {nullptr, nullptr};  // what was submitted
    size_t n[2] = {0, 0};
    uint64_t seq = 0;
};
}  // namespace

struct fqg_stream {
    fqg_ctx* ctx = nullptr;
    const fqg_matcher* m = nullptr;
    int mode = FQG_SINGLE;
    bool want_spans = false;
    size_t slot_bytes = 0;
    std::vector<Slot> slots;
    uint64_t submitted = 0, retired = 0;      // batch numbers; batch b lives in slot b % n_slots
    bool acquired = false;
};

namespace {

int stream_submit_slot(fqg_stream* s, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2) {
    if (s->submitted - s->retired >= s->slots.size()) return fail(FQG_E_INVALID_ARG, "all slots are in flight: call fqg_stream_next first");
    Slot& sl = s->slots[s->submitted % s->slots.size()];
    BatchIn in;
    in.mode = s->mode;
    in.r[0] = r1; in.n[0] = n1;
    in.r[1] = r2; in.n[1] = n2;
    in.want_spans = s->want_spans;
    const int rc = lane_submit(s->ctx, sl.lane, s->m, in);
    if (rc) return rc;
    sl.in_flight = true;
    sl.r[0] = r1; sl.n[0] = n1; sl.r[1] = r2; sl.n[1] = n2;
    sl.seq = s->submitted++;
    s->acquired = false;
    return FQG_OK;
}

// OR `n_bits` bits of src (starting at bit 0) into dst starting at bit `at`
void merge_bits(uint32_t* dst, uint64_t at, const uint32_t* src, uint64_t n_bits) {
    if (n_bits == 0) return;
    const uint64_t words = (n_bits + 31) / 32;
    const uint32_t sh = (uint32_t)(at & 31);
    uint32_t* d = dst + (at >> 5);
    const uint32_t last_keep = (n_bits & 31) ? ((1u << (n_bits & 31)) - 1u) : 0xFFFFFFFFu;
    if (sh == 0) {
        for (uint64_t i = 0; i + 1 < words; i++) d[i] |= src[i];
        d[words - 1] |= src[words - 1] & last_keep;
        return;
    }
    for (uint64_t i = 0; i < words; i++) {
        const uint32_t w = i + 1 == words ? (src[i] & last_keep) : src[i];
        d[i] |= w << sh;
        const uint32_t hi = w >> (32 - sh);
        if (hi) d[i + 1] |= hi;
    }
}

struct WholeOut {              // where the whole-input entry points put their results
    uint32_t* unit_mask = nullptr;
    size_t unit_mask_words = 0;
    uint8_t *out = nullptr, *out2 = nullptr;
    size_t out_cap = 0, out2_cap = 0, out_len = 0, out2_len = 0;
    bool write = false;
};

// One host-resident input through rings on n_ctx GPUs, batch b on GPU b % n_ctx, results taken back in batch order.
int grep_whole(fqg_ctx* const* ctxs, uint32_t n_ctx, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, WholeOut* wo,
               fqg_result* res) {
    if (mode != FQG_SINGLE && mode != FQG_INTERLEAVED && mode != FQG_PAIRED) return fail(FQG_E_INVALID_ARG, "unknown mode");
    if ((n1 && !r1) || (mode == FQG_PAIRED && n2 && !r2)) return fail(FQG_E_INVALID_ARG, "NULL input");
    if (mode != FQG_PAIRED) { r2 = nullptr; n2 = 0; }
    std::memset(res, 0, sizeof *res);
    fqg_stream_opts so{};
    so.n_slots = 3;
    so.want_spans = wo->write ? 1u : 0u;
    if (const char* e = std::getenv("FQG_SLOT_BYTES")) so.slot_bytes = (size_t)std::atoll(e);      // test hook: small rings on small inputs
    const size_t slot_bytes = so.slot_bytes ? so.slot_bytes : ((size_t)64 << 20);
    std::vector<fqg_stream*> streams(n_ctx, nullptr);
    struct Closer {
        std::vector<fqg_stream*>& v;
        ~Closer() { for (auto* s : v) fqg_stream_close(s); }
    } closer{streams};
    for (uint32_t g = 0; g < n_ctx; g++) {
        const int rc = fqg_stream_open(ctxs[g], m, mode, &so, &streams[g]);
        if (rc) return rc;
    }
    if (wo->unit_mask && wo->unit_mask_words) std::memset(wo->unit_mask, 0, wo->unit_mask_words * 4);

    struct Pending { size_t p1, p2, c1, c2; };      // start offsets and lengths of a batch in flight
    std::deque<Pending> pending;
    size_t p1 = 0, p2 = 0;
    uint64_t fed = 0, taken = 0, units_done = 0;
    bool exact = false;
    const uint64_t max_in_flight = (uint64_t)n_ctx * so.n_slots;
    auto input_left = [&]() { return p1 < n1 || (mode == FQG_PAIRED && p2 < n2); };
    while (input_left() || !pending.empty()) {
        while (input_left() && pending.size() < max_in_flight) {
            size_t c1 = 0, c2 = 0;
            fastq_cut(mode, r1 + p1, n1 - p1, r2 ? r2 + p2 : nullptr, n2 - p2, slot_bytes, true, exact, &c1, &c2);
            if (c1 == 0 && c2 == 0) {
                // not one whole unit within a slot (a huge record, or bytes that do not parse): hand over everything that
                // is left as one batch; the lane's buffers grow to it and the device reports what is wrong with it
                c1 = n1 - p1;
                c2 = n2 - p2;
            }
            const int rc = fqg_stream_feed(streams[fed % n_ctx], r1 + p1, c1, r2 ? r2 + p2 : nullptr, c2);
            if (rc) return rc;
            pending.push_back(Pending{p1, p2, c1, c2});
            p1 += c1;
            p2 += c2;
            fed++;
        }
        if (pending.empty()) break;
        fqg_batch b;
        const int rc = fqg_stream_next(streams[taken % n_ctx], &b);
        taken++;
        const Pending pd = pending.front();
        pending.pop_front();
        if (rc) {
            const bool cut_suspect = (rc == FQG_E_FASTQ || rc == FQG_E_PAIR_MISMATCH) && !exact;
            const std::string first_error = fqg_last_error();
            // every later batch was cut after this one: drop them
            while (!pending.empty()) {
                fqg_batch drop;
                fqg_stream_next(streams[taken % n_ctx], &drop);
                taken++;
                pending.pop_front();
            }
            if (!cut_suspect) return fail(rc, first_error);
            // the id-guided cut may have been wrong: redo this batch and the rest of the input with counted cuts
            exact = true;
            p1 = pd.p1;
            p2 = pd.p2;
            continue;
        }
        res->n_records += b.n_records;
        res->n_units += b.n_units;
        res->n_selected += b.n_selected;
        res->h2d_ms += b.h2d_ms;
        res->kernel_ms += b.kernel_ms;
        res->h2d_bytes += pd.c1 + pd.c2;
        res->d2h_bytes += 64 + ((b.n_units + 31) / 32) * 4 + b.n_spans * sizeof(fqg_span);
        if (wo->unit_mask) {
            if ((units_done + b.n_units + 31) / 32 > wo->unit_mask_words)
                return fail(FQG_E_INVALID_ARG, "unit_mask_out too small: need " + std::to_string((units_done + b.n_units + 31) / 32) + " words so far");
            merge_bits(wo->unit_mask, units_done, b.unit_mask, b.n_units);
        }
        units_done += b.n_units;
        if (wo->write && b.n_spans) {
            size_t l1 = 0, l2 = 0;
            std::string err;
            const int wrc = write_spans_host(b.spans, b.n_spans, b.r1, b.n1, b.r2, b.n2, wo->out + wo->out_len, wo->out_cap - wo->out_len, &l1,
                                             wo->out2 ? wo->out2 + wo->out2_len : nullptr, wo->out2 ? wo->out2_cap - wo->out2_len : 0, &l2, &err);
            if (wrc) return fail(wrc, err);
            wo->out_len += l1;
            wo->out2_len += l2;
        }
    }
    return FQG_OK;
}

}  // namespace

extern "C" {

int fqg_stream_open(fqg_ctx* ctx, const fqg_matcher* m, int mode, const fqg_stream_opts* opts, fqg_stream** out) {
    if (!ctx || !m || !out) return fail(FQG_E_INVALID_ARG, "NULL argument");
    *out = nullptr;
    if (mode != FQG_SINGLE && mode != FQG_INTERLEAVED && mode != FQG_PAIRED) return fail(FQG_E_INVALID_ARG, "unknown mode");
    auto s = std::make_unique<fqg_stream>();
    s->ctx = ctx;
    s->m = m;
    s->mode = mode;
    s->want_spans = opts && opts->want_spans;
    s->slot_bytes = (opts && opts->slot_bytes) ? opts->slot_bytes : ((size_t)64 << 20);
    const uint32_t n_slots = (opts && opts->n_slots) ? opts->n_slots : 3u;
    if (n_slots > 64) return fail(FQG_E_INVALID_ARG, "at most 64 slots");
    s->slots.resize(n_slots);
    for (auto& sl : s->slots) {
        sl.lane = lane_lease(ctx);
        if (!sl.lane) {
            for (auto& t : s->slots) lane_release(ctx, t.lane);
            return fail(FQG_E_CUDA, "could not create the streams of a FASTQ lane");
        }
    }
    *out = s.release();
    return FQG_OK;
}

int fqg_stream_acquire(fqg_stream* s, uint8_t** r1_buf, uint8_t** r2_buf, size_t* capacity) {
    if (!s || !r1_buf) return fail(FQG_E_INVALID_ARG, "NULL argument");
    if (s->submitted - s->retired >= s->slots.size()) return fail(FQG_E_INVALID_ARG, "all slots are in flight: call fqg_stream_next first");
    Slot& sl = s->slots[s->submitted % s->slots.size()];
    CU_TRY(cudaSetDevice(s->ctx->device));
    const int n_bufs = s->mode == FQG_PAIRED ? 2 : 1;
    for (int i = 0; i < n_bufs; i++)
        if (!sl.pin[i]) CU_TRY(cudaHostAlloc((void**)&sl.pin[i], s->slot_bytes + 64, cudaHostAllocDefault));
    *r1_buf = sl.pin[0];
    if (r2_buf) *r2_buf = sl.pin[1];
    if (capacity) *capacity = s->slot_bytes;
    s->acquired = true;
    return FQG_OK;
}

int fqg_stream_submit(fqg_stream* s, size_t n1, size_t n2) {
    if (!s) return fail(FQG_E_INVALID_ARG, "NULL argument");
    if (!s->acquired) return fail(FQG_E_INVALID_ARG, "fqg_stream_submit without fqg_stream_acquire");
    if (n1 > s->slot_bytes || n2 > s->slot_bytes) return fail(FQG_E_INVALID_ARG, "batch larger than the slot");
    Slot& sl = s->slots[s->submitted % s->slots.size()];
    return stream_submit_slot(s, sl.pin[0], n1, s->mode == FQG_PAIRED ? sl.pin[1] : nullptr, s->mode == FQG_PAIRED ? n2 : 0);
}

int fqg_stream_feed(fqg_stream* s, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2) {
    if (!s) return fail(FQG_E_INVALID_ARG, "NULL argument");
    return stream_submit_slot(s, r1, n1, s->mode == FQG_PAIRED ? r2 : nullptr, s->mode == FQG_PAIRED ? n2 : 0);
}

int fqg_stream_in_flight(const fqg_stream* s) { return s ? (int)(s->submitted - s->retired) : 0; }

int fqg_stream_next(fqg_stream* s, fqg_batch* out) {
    if (!s || !out) return fail(FQG_E_INVALID_ARG, "NULL argument");
    std::memset(out, 0, sizeof *out);
    if (s->retired == s->submitted) return fail(FQG_E_INVALID_ARG, "no batch in flight");
    Slot& sl = s->slots[s->retired % s->slots.size()];
    s->retired++;
    sl.in_flight = false;
    BatchOut bo;
    const int rc = lane_finish(s->ctx, sl.lane, s->m, &bo);
    out->seq = sl.seq;
    out->r1 = sl.r[0]; out->n1 = sl.n[0];
    out->r2 = sl.r[1]; out->n2 = sl.n[1];
    if (rc) return rc;
    out->n_records = bo.res.n_records;
    out->n_units = bo.res.n_units;
    out->n_selected = bo.res.n_selected;
    out->unit_mask = bo.unit_mask;
    out->spans = bo.spans;
    out->n_spans = bo.n_spans;
    out->h2d_ms = bo.res.h2d_ms;
    out->kernel_ms = bo.res.kernel_ms;
    return FQG_OK;
}

void fqg_stream_close(fqg_stream* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    for (auto& sl : s->slots) {
        if (sl.in_flight) { BatchOut bo; lane_wait(s->ctx, sl.lane, &bo); }
        lane_release(s->ctx, sl.lane);
        for (auto& p : sl.pin) if (p) cudaFreeHost(p);
    }
    delete s;
}

int fqg_fastq_cut(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, size_t limit, int final, int exact, size_t* cut1, size_t* cut2) {
    if (!cut1 || (n1 && !r1) || (mode == FQG_PAIRED && (!cut2 || (n2 && !r2)))) return fail(FQG_E_INVALID_ARG, "NULL argument");
    if (mode != FQG_SINGLE && mode != FQG_INTERLEAVED && mode != FQG_PAIRED) return fail(FQG_E_INVALID_ARG, "unknown mode");
    size_t c1 = 0, c2 = 0;
    fastq_cut(mode, r1, n1, mode == FQG_PAIRED ? r2 : nullptr, mode == FQG_PAIRED ? n2 : 0, limit, final != 0, exact != 0, &c1, &c2);
    *cut1 = c1;
    if (cut2) *cut2 = c2;
    return FQG_OK;
}

int fqg_write_spans(const fqg_span* spans, uint64_t n_spans, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, uint8_t* out, size_t out_cap,
                    size_t* out_len, uint8_t* out2, size_t out2_cap, size_t* out2_len) {
    std::string err;
    const int rc = write_spans_host(spans, n_spans, r1, n1, r2, n2, out, out_cap, out_len, out2, out2_cap, out2_len, &err);
    if (rc) return fail(rc, err);
    return FQG_OK;
}

int fqg_grep_fastq_host(fqg_ctx* c, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, uint32_t* unit_mask_out,
                        size_t unit_mask_words, fqg_result* res) {
    if (!c || !m || !res) return fail(FQG_E_INVALID_ARG, "NULL argument");
    WholeOut wo;
    wo.unit_mask = unit_mask_out;
    wo.unit_mask_words = unit_mask_words;
    fqg_ctx* ctxs[1] = {c};
    return grep_whole(ctxs, 1, m, mode, r1, n1, r2, n2, &wo, res);
}

int fqg_grep_fastq_host_write(fqg_ctx* c, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, uint8_t* out,
                              size_t out_cap, size_t* out_len, uint8_t* out2, size_t out2_cap, size_t* out2_len, fqg_result* res) {
    if (!c || !m || !res || !out_len || (out_cap && !out)) return fail(FQG_E_INVALID_ARG, "NULL argument");
    WholeOut wo;
    wo.write = true;
    wo.out = out; wo.out_cap = out_cap;
    wo.out2 = out2; wo.out2_cap = out2 ? out2_cap : 0;
    fqg_ctx* ctxs[1] = {c};
    const int rc = grep_whole(ctxs, 1, m, mode, r1, n1, r2, n2, &wo, res);
    *out_len = wo.out_len;
    if (out2_len) *out2_len = wo.out2_len;
    return rc;
}

int fqg_grep_fastq_host_multi(fqg_ctx* const* ctxs, uint32_t n_ctx, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2,
                              uint32_t* unit_mask_out, size_t unit_mask_words, uint8_t* out, size_t out_cap, size_t* out_len, uint8_t* out2, size_t out2_cap,
                              size_t* out2_len, fqg_result* res) {
    if (!ctxs || n_ctx == 0 || !m || !res) return fail(FQG_E_INVALID_ARG, "NULL argument");
    for (uint32_t g = 0; g < n_ctx; g++) if (!ctxs[g]) return fail(FQG_E_INVALID_ARG, "NULL context");
    WholeOut wo;
    wo.unit_mask = unit_mask_out;
    wo.unit_mask_words = unit_mask_words;
    wo.write = out != nullptr;
    wo.out = out; wo.out_cap = out_cap;
    wo.out2 = out2; wo.out2_cap = out2 ? out2_cap : 0;
    const int rc = grep_whole(ctxs, n_ctx, m, mode, r1, n1, r2, n2, &wo, res);
    if (out_len) *out_len = wo.out_len;
    if (out2_len) *out2_len = wo.out2_len;
    return rc;
}

}  // extern "C"
