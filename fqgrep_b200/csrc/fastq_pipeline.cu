// fastq_pipeline.cu -- one FASTQ batch through the fused parse + match kernel, asynchronously.
//
// A LANE is everything one in-flight batch needs: a compute stream, an H2D copy stream, events, grow-only device scratch
// (the batch's bytes, look-back state, record bitmasks, id hashes, record spans) and pinned host staging for what comes
// back (control block, unit mask, selected spans).  lane_submit enqueues the whole batch and returns; lane_wait blocks on
// it and decodes the outcome.  The one-call entry points use lane 0; fqg_stream_* (stream.cu) keeps several lanes busy
// so that the copy of batch k+1 overlaps the kernels of batch k and the write-back of batch k-1 -- the reference overlaps
// its reader thread, worker pool and writer the same way (src/lib/seq_io.rs:287-368).
//
//   host source  : chunked cudaMemcpyAsync on the copy stream into the lane's device buffer; after each chunk the fused
//                  kernel is launched (compute stream, event-ordered) over the tiles that are now complete.
//   device source: one launch per input stream on the caller's CUDA stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "internal.hpp"

namespace fqg {

namespace {
enum { D_BUF0 = 0, D_BUF1, D_STATE0, D_STATE1, D_COUNTERS0, D_COUNTERS1, D_MASK0, D_MASK1, D_IDH0, D_IDH1, D_RECOFF0, D_RECOFF1, D_RECLEN0, D_RECLEN1,
       D_CTL, D_UNITS, D_SPANS, D_SCAN, D_STATS, D_SOASTART0, D_SOASTART1, D_SOAEND0, D_SOAEND1, D_N };
enum { H_CTL = 0, H_MASK, H_SPANS, H_N };
constexpr size_t kCopyChunk = 32ull << 20;
constexpr uint32_t kMaxLaunches = 1u << 12;

struct StreamPlan {
    const uint8_t* src = nullptr;   // as given (host or device)
    size_t n = 0;                   // bytes with trailing newlines stripped
    const uint8_t* d = nullptr;     // device bytes
    uint32_t tiles = 0;
    unsigned long long* state = nullptr;
    uint32_t* counters = nullptr;
    uint32_t* mask = nullptr;
    unsigned long long* idh = nullptr;
    unsigned long long* rec_off = nullptr;
    uint32_t* rec_len = nullptr;
    unsigned long long* seq_start = nullptr;     // record index (K0): the caller's arrays, or lane scratch when the batch is routed through the SoA kernels
    uint32_t* seq_len = nullptr;
    uint64_t seq_cap = 0;
    uint32_t *soa_start = nullptr, *soa_end = nullptr;      // routed batches: every record's sequence line as offsets into the stream itself
    uint32_t record_bytes_est = 0;               // routed batches: typical record size (sizes the SoA kernel's staging slots)
    uint64_t mask_bits = 0, idh_cap = 0, mask_words = 0;
    size_t copied = 0;
    uint32_t tiles_done = 0, launches = 0;
};

// look-back state of one stream: [per-tile counts (u32, in u64 slots)][seg_acc][seg_pre], the segment arrays on 32-byte
// boundaries with slack (the scan warp reads four segments per lane with vector loads)
size_t seg_offset(uint32_t tiles) { return ((size_t)tiles + 1 + 3) & ~(size_t)3; }
size_t seg_words(uint32_t tiles) { return (((size_t)tiles / 32 + 2) + 7) & ~(size_t)3; }
size_t state_words(uint32_t tiles) { return seg_offset(tiles) + 2 * seg_words(tiles); }

const char* fastq_err_name(unsigned code) {
    switch (code) {
        case 1: return "expected '@' at record start (InvalidStart)";
        case 2: return "expected '+' separator line (InvalidSep)";
        case 3: return "sequence and quality lengths differ (UnequalLengths)";
        case 4: return "truncated record (UnexpectedEnd)";
        case 6: return "record bitmask capacity exceeded (records shorter than 8 bytes, or 16 bytes with pair id check)";
    }
    return "unknown";
}
}  // namespace

struct FastqLane {
    int device = 0;
    cudaStream_t cs = nullptr, xs = nullptr;      // compute, H2D
    cudaEvent_t ev[6] = {nullptr};                // 0 begin, 1 h2d begin, 2 h2d end, 3 kernels end, 4 results on the host
    cudaEvent_t chunk_ev[8] = {nullptr};
    void* d[D_N] = {nullptr};
    size_t d_cap[D_N] = {0};
    void* h[H_N] = {nullptr};
    size_t h_cap[H_N] = {0};
    bool leased = false;
    // the batch in flight
    bool busy = false;
    BatchIn in;
    StreamPlan sp[2];
    cudaStream_t compute = nullptr;
    const uint32_t* unit_mask_dev = nullptr;
    uint64_t units_words_cap = 0, eager_mask_words = 0;
    uint64_t h2d_bytes = 0;
    unsigned stripped = 0;
    bool names = false, have_matcher = false;
    unsigned long long* stats = nullptr;

    int dev_ensure(int slot, size_t bytes, void** out) {
        if (d_cap[slot] < bytes) {
            if (d[slot]) cudaFree(d[slot]);
            d[slot] = nullptr;
            d_cap[slot] = 0;
            const size_t want = bytes + bytes / 16 + 4096;
            CU_TRY(cudaMalloc(&d[slot], want));
            d_cap[slot] = want;
        }
        *out = d[slot];
        return 0;
    }
    int host_ensure(int slot, size_t bytes, void** out) {
        if (h_cap[slot] < bytes) {
            if (h[slot]) cudaFreeHost(h[slot]);
            h[slot] = nullptr;
            h_cap[slot] = 0;
            const size_t want = bytes + bytes / 8 + 4096;
            CU_TRY(cudaHostAlloc(&h[slot], want, cudaHostAllocDefault));
            h_cap[slot] = want;
        }
        *out = h[slot];
        return 0;
    }
    ~FastqLane() {
        cudaSetDevice(device);
        if (cs) cudaStreamSynchronize(cs);
        if (xs) cudaStreamSynchronize(xs);
        for (auto& p : d) if (p) cudaFree(p);
        for (auto& p : h) if (p) cudaFreeHost(p);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& e : chunk_ev) if (e) cudaEventDestroy(e);
        if (cs) cudaStreamDestroy(cs);
        if (xs) cudaStreamDestroy(xs);
    }
};

FastqLane* lane_lease(fqg_ctx* c) {
    std::lock_guard<std::mutex> lock(c->lanes_mu);
    for (auto* l : c->lanes)
        if (!l->leased) { l->leased = true; return l; }
    auto l = std::make_unique<FastqLane>();
    l->device = c->device;
    if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&l->cs, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&l->xs, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto& e : l->ev) if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    for (auto& e : l->chunk_ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    l->leased = true;
    c->lanes.push_back(l.release());
    return c->lanes.back();
}
void lane_release(fqg_ctx* c, FastqLane* lane) {
    if (!lane) return;
    std::lock_guard<std::mutex> lock(c->lanes_mu);
    lane->busy = false;
    lane->leased = false;
}
void lanes_free(fqg_ctx* c) {
    std::lock_guard<std::mutex> lock(c->lanes_mu);
    for (auto* l : c->lanes) delete l;
    c->lanes.clear();
}

int lane_submit(fqg_ctx* c, FastqLane* L, const fqg_matcher* m, const BatchIn& in) {
    if (!L) return fail(FQG_E_CUDA, "could not create the streams of a FASTQ lane");
    if (L->busy) return fail(FQG_E_INVALID_ARG, "lane already holds a batch");
    const int mode = in.mode;
    if (mode != FQG_SINGLE && mode != FQG_INTERLEAVED && mode != FQG_PAIRED) return fail(FQG_E_INVALID_ARG, "unknown mode");
    if ((in.n[0] && !in.r[0]) || (mode == FQG_PAIRED && in.n[1] && !in.r[1])) return fail(FQG_E_INVALID_ARG, "NULL input");
    CU_TRY(cudaSetDevice(c->device));
    DeviceDfa dd;
    int rc = (m && m->kind != FQG_NAMES) ? get_device_dfa(c, m, &dd) : 0;
    if (rc) return rc;
    const bool names = m && m->kind == FQG_NAMES;
    NamesView nv;
    if (names && (rc = get_names_view(c, m, &nv))) return rc;
    const int n_streams = mode == FQG_PAIRED ? 2 : 1;
    const bool check_ids = mode != FQG_SINGLE;
    const bool on_device = in.on_device;
    // Large literal sets (tier 3: q-gram filter, tier 5: split sets) have no table the fused kernel could walk out of shared
    // memory -- their exact automaton lives in L2 and walking it per base runs at 0.04 of the roofline (profiles/r2c_*).  Such
    // batches are ROUTED: the fused kernel only parses, validates and indexes (K0: u32 start / end of every sequence line), and
    // the SoA kernel with the filter then reads those lines where they lie in the stream and decides
    // (src/lib/matcher.rs:249-251 for every record, as before).
    bool route = m && !names && (m->tier == 3 || m->split_lit);
    for (int i = 0; i < n_streams; i++) if (in.n[i] >= 0xFFF00000ull) route = false;      // the index holds u32 offsets into the stream
    if (on_device)
        for (int i = 0; i < n_streams; i++)
            if (reinterpret_cast<uintptr_t>(in.r[i]) & 15u) return fail(FQG_E_INVALID_ARG, "device FASTQ buffers must be 16-byte aligned");
    cudaStream_t cs = on_device ? in.user_stream : L->cs;       // compute stream
    cudaStream_t xs = L->xs;                                    // H2D stream (host sources)
    const uint32_t kTile = fastq_tile_bytes(), kLoad = kTile + fastq_overlap_bytes();

    L->in = in;
    L->compute = cs;
    L->stripped = 0;
    L->names = names;
    L->have_matcher = m != nullptr;
    StreamPlan* sp = L->sp;
    sp[0] = StreamPlan();
    sp[1] = StreamPlan();
    sp[0].src = in.r[0]; sp[0].n = in.n[0];
    sp[1].src = in.r[1]; sp[1].n = in.n[1];
    for (int i = 0; i < n_streams; i++) {
        StreamPlan& s = sp[i];
        // strip trailing newline bytes (seq_io tolerates a missing final newline and trailing blank lines)
        uint8_t tail[4096];
        size_t n = s.n;
        while (n) {
            size_t k = n < sizeof tail ? n : sizeof tail;
            if (on_device) CU_TRY(cudaMemcpy(tail, s.src + n - k, k, cudaMemcpyDeviceToHost));
            else std::memcpy(tail, s.src + n - k, k);
            size_t j = k;
            while (j && (tail[j - 1] == '\n' || tail[j - 1] == '\r')) j--;
            n -= k - j;
            if (j) break;
        }
        if (n < s.n) {
            // Trailing line terminators were dropped.  That is right unless the last record's quality line is EMPTY: then
            // one of them ended its '+' line (seq_io accepts "...+\n<EOF>" as a record with empty quality).  Which case
            // this is depends on the line count of the whole stream, so the caller finds out by trying: a parse that
            // fails with the terminators dropped is repeated with one of them kept (grep_fastq_batch).
            uint8_t first2[2] = {0, 0};
            const size_t k = s.n - n < 2 ? s.n - n : 2;
            if (on_device) CU_TRY(cudaMemcpy(first2, s.src + n, k, cudaMemcpyDeviceToHost));
            else std::memcpy(first2, s.src + n, k);
            const size_t term = (first2[0] == '\r' && k == 2 && first2[1] == '\n') ? 2 : 1;
            if (first2[0] == '\n' || term == 2) {
                L->stripped |= 1u << i;
                if ((in.keep_newline_mask >> i) & 1u) n += term;
            }
        }
        s.n = n;
        const uint64_t stream_len = n ? n + 1 : 0;
        if (stream_len / kTile >= 0xFFFFFFF0ull) return fail(FQG_E_TOO_LARGE, "input stream too large");
        s.tiles = (uint32_t)((stream_len + kTile - 1) / kTile);
        s.mask_bits = n / in.min_record_bytes + 64;      // capacity for records of >= min_record_bytes on average
        s.mask_words = (s.mask_bits + 31) / 32 + 2;
        s.idh_cap = check_ids ? s.mask_bits : 0;
        void* p = nullptr;
        if (!on_device) {
            if ((rc = L->dev_ensure(D_BUF0 + i, n + 64, &p))) return rc;
            s.d = (const uint8_t*)p;
        } else s.d = s.src;
        // look-back state: one word per tile, then two per segment of 32 tiles
        if ((rc = L->dev_ensure(D_STATE0 + i, state_words(s.tiles) * 8, &p))) return rc;
        s.state = (unsigned long long*)p;
        if ((rc = L->dev_ensure(D_COUNTERS0 + i, kMaxLaunches * 4, &p))) return rc;
        s.counters = (uint32_t*)p;
        if ((rc = L->dev_ensure(D_MASK0 + i, s.mask_words * 4, &p))) return rc;
        s.mask = (uint32_t*)p;
        if (check_ids) {
            if ((rc = L->dev_ensure(D_IDH0 + i, s.idh_cap * 8 + 16, &p))) return rc;
            s.idh = (unsigned long long*)p;
        }
        if (in.idx && i == 0) { s.seq_start = in.idx->seq_start; s.seq_len = in.idx->seq_len; s.seq_cap = in.idx->cap; }
        if (route) {
            if ((rc = L->dev_ensure(D_SOASTART0 + i, s.mask_bits * 4 + 16, &p))) return rc;
            s.soa_start = (uint32_t*)p;
            if ((rc = L->dev_ensure(D_SOAEND0 + i, s.mask_bits * 4 + 16, &p))) return rc;
            s.soa_end = (uint32_t*)p;
            // typical record size from the head of the stream (4 lines per record): the SoA kernel stages the span of 32
            // consecutive records per warp; batches that turn out larger than the slot are walked in global memory
            uint8_t head[16384];
            const size_t k = n < sizeof head ? n : sizeof head;
            if (on_device) CU_TRY(cudaMemcpy(head, s.src, k, cudaMemcpyDeviceToHost));
            else std::memcpy(head, s.src, k);
            size_t nl = 0;
            for (size_t j = 0; j < k; j++) nl += head[j] == '\n';
            const size_t est = nl >= 4 ? (4 * k + nl - 1) / nl : k;
            s.record_bytes_est = (uint32_t)(est + est / 16 + 8);
        }
        if (in.want_spans) {
            if ((rc = L->dev_ensure(D_RECOFF0 + i, s.mask_bits * 8 + 16, &p))) return rc;
            s.rec_off = (unsigned long long*)p;
            if ((rc = L->dev_ensure(D_RECLEN0 + i, s.mask_bits * 4 + 16, &p))) return rc;
            s.rec_len = (uint32_t*)p;
        }
    }
    // control block: [0] err, [1] count, [2] id mismatch, [3] lines r1, [4] lines r2, [5] selected spans
    void* p = nullptr;
    if ((rc = L->dev_ensure(D_CTL, 64, &p))) return rc;
    unsigned long long* ctl = (unsigned long long*)p;
    void* d_units = nullptr;
    L->units_words_cap = sp[0].mask_words;
    if ((rc = L->dev_ensure(D_UNITS, L->units_words_cap * 4 + 16, &d_units))) return rc;
    void* h_ctl = nullptr;
    if ((rc = L->host_ensure(H_CTL, 128, &h_ctl))) return rc;
    static const bool want_stats = std::getenv("FQG_FASTQ_STATS") != nullptr;      // measurement hook: tile statistics of the fused kernel on stderr
    unsigned long long* d_stats = nullptr;
    if (want_stats) {
        void* ps = nullptr;
        if ((rc = L->dev_ensure(D_STATS, 64, &ps))) return rc;
        d_stats = (unsigned long long*)ps;
    }
    L->stats = d_stats;

    unsigned long long* ctl_init = (unsigned long long*)h_ctl;      // pinned: the async copy below reads it when it runs
    ctl_init[0] = ~0ull; ctl_init[1] = 0; ctl_init[2] = ~0ull; ctl_init[3] = 0; ctl_init[4] = 0; ctl_init[5] = 0; ctl_init[6] = 0; ctl_init[7] = 0;
    CU_TRY(cudaEventRecord(L->ev[0], cs));
    CU_TRY(cudaMemcpyAsync(ctl, ctl_init, 64, cudaMemcpyHostToDevice, cs));
    if (d_stats) CU_TRY(cudaMemsetAsync(d_stats, 0, 64, cs));
    for (int i = 0; i < n_streams; i++) {
        CU_TRY(cudaMemsetAsync(sp[i].state, 0, state_words(sp[i].tiles) * 8, cs));
        CU_TRY(cudaMemsetAsync(sp[i].counters, 0, kMaxLaunches * 4, cs));
        CU_TRY(cudaMemsetAsync(sp[i].mask, 0, sp[i].mask_words * 4, cs));
    }

    auto launch = [&](int i, uint32_t tile_end) -> int {
        StreamPlan& s = sp[i];
        if (tile_end <= s.tiles_done) return 0;
        if (s.launches >= kMaxLaunches) return fail(FQG_E_TOO_LARGE, "too many launches for one stream");
        FastqArgs a{};
        a.buf = s.d; a.nbytes = s.n; a.virtual_final_nl = 1;
        a.tile_first = s.tiles_done; a.tile_count = tile_end - s.tiles_done;
        a.tile_counter = s.counters + s.launches; a.tile_state = s.state; a.seg_acc = s.state + seg_offset(s.tiles); a.seg_pre = a.seg_acc + seg_words(s.tiles); a.lines_out = ctl + 3 + i;
        a.mask = s.mask; a.mask_bits = s.mask_bits; a.id_hash = s.idh; a.id_hash_cap = s.idh_cap;
        a.count = (mode == FQG_SINGLE && !route) ? ctl + 1 : nullptr;
        a.err = ctl + 0;
        a.invert = (m && (m->opts & FQG_INVERT_MATCH)) ? 1u : 0u;
        a.names = nv;
        a.seq_start = s.seq_start; a.seq_len = s.seq_len; a.span_cap = s.seq_cap;
        a.soa_start = s.soa_start; a.soa_end = s.soa_end;
        if (s.soa_start) a.span_cap = s.mask_bits;
        a.rec_off = s.rec_off; a.rec_len = s.rec_len; a.rec_cap = s.rec_off ? s.mask_bits : 0;
        a.stats = d_stats;
        CU_TRY(launch_match_fastq(dd, names, m == nullptr || route, a, c->sm_count, cs));
        count_launches(1);
        s.launches++;
        s.tiles_done = tile_end;
        return 0;
    };

    L->h2d_bytes = 0;
    if (on_device) {
        for (int i = 0; i < n_streams; i++) if ((rc = launch(i, sp[i].tiles))) return rc;
    } else {
        // interleave the streams chunk by chunk so mates progress together
        CU_TRY(cudaEventRecord(L->ev[1], xs));
        bool more = true;
        int ev_slot = 0;
        while (more) {
            more = false;
            for (int i = 0; i < n_streams; i++) {
                StreamPlan& s = sp[i];
                if (s.copied >= s.n) continue;
                size_t k = s.n - s.copied < kCopyChunk ? s.n - s.copied : kCopyChunk;
                CU_TRY(cudaMemcpyAsync(const_cast<uint8_t*>(s.d) + s.copied, s.src + s.copied, k, cudaMemcpyHostToDevice, xs));
                s.copied += k;
                L->h2d_bytes += k;
                cudaEvent_t e = L->chunk_ev[ev_slot++ % 8];
                CU_TRY(cudaEventRecord(e, xs));
                CU_TRY(cudaStreamWaitEvent(cs, e, 0));
                uint32_t ready = s.copied >= s.n ? s.tiles : (s.copied >= kLoad ? (uint32_t)((s.copied - kLoad) / kTile + 1) : 0u);
                if ((rc = launch(i, ready))) return rc;
                if (s.copied < s.n) more = true;
            }
        }
        CU_TRY(cudaEventRecord(L->ev[2], xs));
        for (int i = 0; i < n_streams; i++) if ((rc = launch(i, sp[i].tiles))) return rc;   // empty streams / leftovers
    }

    if (route) {
        // every record of stream i is indexed by now; the SoA kernel(s) read the sequence lines where they lie in the stream
        // (bases = the stream, start / end = the index) and leave the stream's record mask
        for (int i = 0; i < n_streams; i++) {
            StreamPlan& s = sp[i];
            if (s.tiles == 0) continue;
            SoaBatch b{s.d, s.soa_start, s.soa_end, (uint32_t)(s.mask_bits < 0xFFFFFFFFull ? s.mask_bits : 0xFFFFFFFFull), s.mask, nullptr,
                       mode == FQG_SINGLE ? ctl + 1 : nullptr, (m->opts & FQG_INVERT_MATCH) ? 1u : 0u, s.record_bytes_est, nullptr, ctl + 3 + i};
            if ((rc = match_soa_any(c, m, b, cs))) return rc;
        }
    }
    // ---- mates -> units (src/main.rs:749-755), id agreement (:741-747)
    L->unit_mask_dev = sp[0].mask;
    uint64_t unit_words_upper = sp[0].mask_words;
    if (mode == FQG_PAIRED) {
        const uint64_t w = sp[0].mask_words < sp[1].mask_words ? sp[0].mask_words : sp[1].mask_words;
        const uint64_t nrec = sp[0].idh_cap < sp[1].idh_cap ? sp[0].idh_cap : sp[1].idh_cap;
        CU_TRY(launch_combine_pairs(sp[0].mask, sp[1].mask, (uint32_t*)d_units, w, sp[0].idh, sp[1].idh, nrec, ctl + 3, ctl + 4, ctl + 1, ctl + 2, cs));
        count_launches(1);
        L->unit_mask_dev = (const uint32_t*)d_units;
        unit_words_upper = w;
    } else if (mode == FQG_INTERLEAVED) {
        CU_TRY(launch_combine_interleaved(sp[0].mask, (uint32_t*)d_units, sp[0].mask_words / 2, sp[0].idh, sp[0].idh_cap, ctl + 3, ctl + 1, ctl + 2, cs));
        count_launches(1);
        L->unit_mask_dev = (const uint32_t*)d_units;
        unit_words_upper = sp[0].mask_words / 2;
    }
    if (in.want_spans) {
        // K5: selected records, in output order, as (offset, length, flags) -- the write-back copies spans instead of parsing
        void *d_spans, *d_scan;
        if ((rc = L->dev_ensure(D_SPANS, (sp[0].mask_bits + (mode == FQG_PAIRED ? sp[1].mask_bits : 0)) * sizeof(fqg_span) + 64, &d_spans))) return rc;
        if ((rc = L->dev_ensure(D_SCAN, span_scan_words(unit_words_upper) * 4, &d_scan))) return rc;
        CU_TRY(launch_select_spans(L->unit_mask_dev, unit_words_upper, mode, ctl + 3, sp[0].rec_off, sp[0].rec_len, sp[1].rec_off, sp[1].rec_len,
                                   (fqg_span*)d_spans, ctl + 5, (uint32_t*)d_scan, cs));
        count_launches(3);
    }
    CU_TRY(cudaEventRecord(L->ev[3], cs));
    CU_TRY(cudaMemcpyAsync(h_ctl, ctl, 64, cudaMemcpyDeviceToHost, cs));
    if (d_stats) CU_TRY(cudaMemcpyAsync((uint8_t*)h_ctl + 64, d_stats, 64, cudaMemcpyDeviceToHost, cs));
    // The unit mask goes back without waiting for the record count: every word the count could ask for is copied now
    // (1 bit per min_record_bytes of input) into pinned staging, so a host batch needs ONE synchronisation.
    L->eager_mask_words = 0;
    if (!on_device && m) {
        void* h_mask;
        if ((rc = L->host_ensure(H_MASK, unit_words_upper * 4 + 16, &h_mask))) return rc;
        CU_TRY(cudaMemcpyAsync(h_mask, L->unit_mask_dev, unit_words_upper * 4, cudaMemcpyDeviceToHost, cs));
        L->eager_mask_words = unit_words_upper;
    }
    CU_TRY(cudaEventRecord(L->ev[4], cs));
    L->busy = true;
    return FQG_OK;
}

int lane_wait(fqg_ctx* c, FastqLane* L, BatchOut* out) {
    *out = BatchOut();
    if (!L || !L->busy) return fail(FQG_E_INVALID_ARG, "lane holds no batch");
    L->busy = false;
    out->stripped_mask = L->stripped;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaEventSynchronize(L->ev[4]));
    const BatchIn& in = L->in;
    const StreamPlan* sp = L->sp;
    const int mode = in.mode;
    const int n_streams = mode == FQG_PAIRED ? 2 : 1;
    const unsigned long long* h_ctl = (const unsigned long long*)L->h[H_CTL];
    if (L->stats) {
        std::fprintf(stderr, "[fqg] fused kernel: %llu tiles, %llu with a guessed line phase (%llu guessed wrong), %llu dense", h_ctl[8] + h_ctl[11], h_ctl[9], h_ctl[10], h_ctl[11]);
        if (h_ctl[14])      // only a FQG_FASTQ_TIMING build counts cycles
            std::fprintf(stderr, "; of %llu warp-cycles %.1f %% waited for the copy, %.1f %% for the prefix (%.1f %% on the scan warp's word)", h_ctl[14],
                         100.0 * h_ctl[12] / h_ctl[14], 100.0 * h_ctl[13] / h_ctl[14], 100.0 * h_ctl[15] / h_ctl[14]);
        std::fprintf(stderr, "\n");
    }
    fqg_result* res = &out->res;
    uint64_t d2h_bytes = 64 + L->eager_mask_words * 4;

    if (h_ctl[0] != ~0ull) {
        unsigned code = (unsigned)(h_ctl[0] & 15);
        if (code == 6) out->capacity_hit = true;
        return fail(code >= 6 ? FQG_E_TOO_LARGE : FQG_E_FASTQ, std::string("FASTQ parse error at byte ") + std::to_string(h_ctl[0] >> 4) + ": " + fastq_err_name(code));
    }
    uint64_t lines[2] = {sp[0].n ? h_ctl[3] : 0, sp[1].n ? h_ctl[4] : 0};
    for (int i = 0; i < n_streams; i++)
        if (lines[i] % 4) return fail(FQG_E_FASTQ, "FASTQ parse error: truncated final record (UnexpectedEnd)");
    const uint64_t rec1 = lines[0] / 4, rec2 = lines[1] / 4;
    uint64_t units = rec1;
    if (mode == FQG_PAIRED) {
        if (rec1 != rec2) return fail(FQG_E_PAIR_MISMATCH, "FASTQ files out of sync: R1 has " + std::to_string(rec1) + " records, R2 has " + std::to_string(rec2));
    } else if (mode == FQG_INTERLEAVED) {
        if (rec1 % 2) return fail(FQG_E_FASTQ, "FASTQ file does not have an even number of records.");
        units = rec1 / 2;
    }
    if (mode != FQG_SINGLE && h_ctl[2] != ~0ull) return fail(FQG_E_PAIR_MISMATCH, "Mismatching read pair! (pair " + std::to_string(h_ctl[2]) + ")");
    res->n_records = rec1 + (mode == FQG_PAIRED ? rec2 : 0);
    res->n_units = units;
    res->n_selected = h_ctl[1];
    out->mask_words = (units + 31) / 32;
    out->unit_mask_dev = L->unit_mask_dev;
    if (L->eager_mask_words) {
        if (out->mask_words > L->eager_mask_words) return fail(FQG_E_TOO_LARGE, "internal: unit mask staging too small");
        out->unit_mask = (const uint32_t*)L->h[H_MASK];
    }
    if (in.want_spans) {
        const uint64_t n_spans = h_ctl[5];
        out->n_spans = n_spans;
        if (n_spans) {
            void* h_spans = nullptr;
            int rc = L->host_ensure(H_SPANS, n_spans * sizeof(fqg_span), &h_spans);
            if (rc) return rc;
            CU_TRY(cudaMemcpyAsync(h_spans, L->d[D_SPANS], n_spans * sizeof(fqg_span), cudaMemcpyDeviceToHost, L->compute));
            CU_TRY(cudaStreamSynchronize(L->compute));
            out->spans = (const fqg_span*)h_spans;
            d2h_bytes += n_spans * sizeof(fqg_span);
        }
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, L->ev[0], L->ev[3]);
    res->kernel_ms = ms;
    if (!in.on_device && L->h2d_bytes) {
        cudaEventElapsedTime(&ms, L->ev[1], L->ev[2]);
        res->h2d_ms = ms;
    }
    cudaEventElapsedTime(&ms, L->ev[3], L->ev[4]);
    res->d2h_ms = ms;
    res->h2d_bytes = L->h2d_bytes;
    res->d2h_bytes = d2h_bytes;
    return FQG_OK;
}

// lane_wait plus the retries a batch can need.  The per-record arrays are sized for records of >= 32 bytes first; denser
// inputs (the reference's 4-base test fixtures) are run again at 8.  A batch that fails to parse after its trailing line
// terminators were dropped is run again with one of them kept (a last record with an empty quality line).
int lane_finish(fqg_ctx* c, FastqLane* lane, const fqg_matcher* m, BatchOut* out) {
    BatchIn in = lane ? lane->in : BatchIn();
    int rc = lane_wait(c, lane, out);
    if (rc && out->capacity_hit && in.min_record_bytes > 8) {
        in.min_record_bytes = 8;
        rc = lane_submit(c, lane, m, in);
        if (rc == FQG_OK) rc = lane_wait(c, lane, out);
    }
    const unsigned stripped = out->stripped_mask;
    if ((rc == FQG_E_FASTQ || rc == FQG_E_PAIR_MISMATCH) && stripped && in.keep_newline_mask == 0) {
        const std::string first_error = fqg_last_error();
        for (unsigned keep = 1; keep <= 3; keep++) {
            if ((keep & stripped) != keep) continue;
            in.keep_newline_mask = keep;
            BatchOut again;
            int rc2 = lane_submit(c, lane, m, in);
            if (rc2 == FQG_OK) rc2 = lane_wait(c, lane, &again);
            if (rc2 == FQG_OK) {
                *out = again;
                return FQG_OK;
            }
        }
        return fail(rc, first_error);
    }
    return rc;
}

int grep_fastq_batch(fqg_ctx* c, FastqLane* lane, const fqg_matcher* m, BatchIn in, BatchOut* out) {
    in.min_record_bytes = 32;
    in.keep_newline_mask = 0;
    *out = BatchOut();
    const int rc = lane_submit(c, lane, m, in);
    if (rc) return rc;
    return lane_finish(c, lane, m, out);
}

}  // namespace fqg
