// api.cu -- the C ABI (include/fqgrep_b200.h): matcher construction, device contexts, match entry points.
// No CPU matching path exists in this library: every match call launches the CUDA kernels or fails.
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "internal.hpp"

using namespace fqg;

namespace {
thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
}  // namespace

namespace fqg {

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
void count_launches(uint64_t n) { g_launches += n; }

int ensure_dev(fqg_ctx* c, int slot, size_t bytes, void** out) {
    if (c->d_cap[slot] < bytes) {
        if (c->d_buf[slot]) cudaFree(c->d_buf[slot]);
        c->d_buf[slot] = nullptr;
        c->d_cap[slot] = 0;
        size_t want = bytes + bytes / 8 + 4096;
        CU_TRY(cudaMalloc(&c->d_buf[slot], want));
        c->d_cap[slot] = want;
    }
    *out = c->d_buf[slot];
    return 0;
}

}  // namespace fqg

namespace {

void build_tables(fqg_matcher* m) {
    const Dfa& d = m->dfa;
    const uint64_t n = d.n_states, nc = d.n_classes;
    {   // states that only lead to themselves: at most a dead one and an all-accepting one in a minimal automaton
        int k = 0;
        for (uint32_t s = 0; s < d.n_states && k < 2; s++) {
            bool self = true;
            for (uint32_t c = 0; c < d.n_classes && self; c++) self = d.next[(size_t)s * d.n_classes + c] == s;
            if (!self) continue;
            m->settled[k++] = s;
            if (s < d.accept_from) m->settle = 1;      // a dead state: every read is decided after a few bases
        }
    }
    std::memcpy(m->lut, d.byte_class, 256);
    const uint64_t row0 = 256 + ((nc + 7) / 8) * 8;      // tier-0 row: 256 four-base entries + the one-byte entries
    if (nc <= 127 && n * row0 * 2 <= 40 * 1024) {
        m->tier = 0;
        m->row = (uint32_t)row0 * 2;   // states are byte offsets of their rows
        m->t16.assign(((n * row0 + 7) / 8) * 8, 0);
        static const char kLetter[4] = {'A', 'C', 'T', 'G'};   // code = (byte >> 1) & 3
        for (uint64_t s = 0; s < n; s++) {
            for (uint32_t p = 0; p < 256; p++) {
                uint64_t t = s;
                // index = c3 | c2<<2 | c1<<4 | c0<<6, c0 = first base in stream order (see step4_acgt)
                for (int k = 0; k < 4; k++) t = d.next[t * nc + d.byte_class[(uint8_t)kLetter[(p >> (6 - 2 * k)) & 3]]];
                m->t16[s * row0 + p] = (uint16_t)(t * row0 * 2);
            }
            for (uint64_t c = 0; c < nc; c++) m->t16[s * row0 + 256 + c] = (uint16_t)(d.next[s * nc + c] * row0 * 2);
        }
        for (int b = 0; b < 256; b++) m->lut[b] = (uint8_t)(d.byte_class[b] * 2);
    } else if ((!m->gram.usable || m->gram.s < 8 || m->gram.expected_hits > 0.1) && nc <= 127 && n * (16 + ((nc + 7) / 8) * 8) * 2 <= 60 * 1024) {
        // tier 4: 16 two-base entries + the one-byte entries per row.  Literal sets prefer the q-gram filter (tier 3) when it
        // probes once per 8 bases or less often and rarely hits (1000 x 20-mers: 0.82 vs 0.62 here); a filter that has to
        // probe every 4 bases, or verifies a lot, loses to this table (60 x 15-mer regexes: 0.54 vs 0.62 here)
        const uint64_t row4 = 16 + ((nc + 7) / 8) * 8;
        m->tier = 4;
        m->row = (uint32_t)row4 * 2;
        m->t16.assign(((n * row4 + 7) / 8) * 8, 0);
        static const char kLetter[4] = {'A', 'C', 'T', 'G'};
        for (uint64_t s = 0; s < n; s++) {
            for (uint32_t p = 0; p < 16; p++) {      // index = c0 << 2 | c1, c0 = first base
                uint64_t t = d.next[s * nc + d.byte_class[(uint8_t)kLetter[(p >> 2) & 3]]];
                t = d.next[t * nc + d.byte_class[(uint8_t)kLetter[p & 3]]];
                m->t16[s * row4 + p] = (uint16_t)(t * row4 * 2);
            }
            for (uint64_t c = 0; c < nc; c++) m->t16[s * row4 + 16 + c] = (uint16_t)(d.next[s * nc + c] * row4 * 2);
        }
        for (int b = 0; b < 256; b++) m->lut[b] = (uint8_t)(d.byte_class[b] * 2);
    } else if (!m->gram.usable && n * nc * 2 <= 60 * 1024 && nc <= 127) {
        m->tier = 1;
        m->row = (uint32_t)nc * 2;
        m->t16.assign(((n * nc + 7) / 8) * 8, 0);
        for (uint64_t i = 0; i < n * nc; i++) m->t16[i] = (uint16_t)(d.next[i] * nc * 2);
        for (int b = 0; b < 256; b++) m->lut[b] = (uint8_t)(d.byte_class[b] * 2);
    } else {
        m->tier = m->gram.usable ? 3 : 2;    // tier 3 keeps the tier-2 table as its exact fallback
        m->row = (uint32_t)nc;
        m->t32.resize(((n * nc + 3) / 4) * 4);
        for (uint64_t i = 0; i < n * nc; i++) m->t32[i] = (uint32_t)(d.next[i] * nc);
    }
}

}  // namespace

namespace fqg {

int get_device_dfa(fqg_ctx* c, const fqg_matcher* cm, DeviceDfa* out) {
    fqg_matcher* m = const_cast<fqg_matcher*>(cm);
    std::lock_guard<std::mutex> lock(m->mu);
    auto it = m->dev.find(c->device);
    if (it == m->dev.end()) {
        DeviceTables t;
        if (m->kind == FQG_NAMES) {
            CU_TRY(cudaMalloc(&t.slots, m->names.slots.size() * 8));
            CU_TRY(cudaMalloc(&t.pool, m->names.pool.size()));
            CU_TRY(cudaMemcpy(t.slots, m->names.slots.data(), m->names.slots.size() * 8, cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(t.pool, m->names.pool.data(), m->names.pool.size(), cudaMemcpyHostToDevice));
        } else {
            const void* src = (m->tier == 2 || m->tier == 3) ? (const void*)m->t32.data() : (const void*)m->t16.data();
            size_t bytes = (m->tier == 2 || m->tier == 3) ? m->t32.size() * 4 : m->t16.size() * 2;
            CU_TRY(cudaMalloc(&t.table, bytes + 16));
            CU_TRY(cudaMemcpy(t.table, src, bytes, cudaMemcpyHostToDevice));
            CU_TRY(cudaMalloc(&t.byte_class, 256));
            CU_TRY(cudaMemcpy(t.byte_class, m->lut, 256, cudaMemcpyHostToDevice));
            if (m->tier == 3) {
                const GramFilter& g = m->gram;
                auto up = [&](auto** dst, const auto& v) -> cudaError_t {
                    cudaError_t e = cudaMalloc((void**)dst, v.size() * sizeof(v[0]) + 16);
                    if (e != cudaSuccess) return e;
                    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice);
                };
                CU_TRY(up(&t.g_bitmap, g.bitmap));
                CU_TRY(up(&t.g_kv, g.slots));
                CU_TRY(up(&t.g_pool, g.pool));
            }
        }
        it = m->dev.emplace(c->device, t).first;
    }
    if (out) {
        out->tier = m->tier;
        out->table = it->second.table;
        out->table_bytes = (uint32_t)((m->tier == 2 || m->tier == 3) ? m->t32.size() * 4 : m->t16.size() * 2);
        out->byte_class = it->second.byte_class;
        if (m->tier == 3) {
            const GramFilter& g = m->gram;
            DeviceGram& d = out->gram;
            d.bitmap = it->second.g_bitmap; d.bitmap_bytes = (uint32_t)g.bitmap.size() * 4; d.bitmap_shift = 32 - (g.bitmap_bits_log2 - 5);     // hash -> WORD index
            d.slots = reinterpret_cast<const uint4*>(it->second.g_kv); d.slot_shift = 32 - g.slots_log2; d.slot_mask = (1u << g.slots_log2) - 1;
            d.pool = it->second.g_pool;
            d.q = g.q; d.s = g.s; d.min_len = g.min_len; d.qmask = g.qmask;
        }
        out->row = m->row;
        out->start = m->dfa.start * m->row;
        out->accept_from = m->dfa.accept_from * m->row;
        out->n_states = m->dfa.n_states;
        out->settled[0] = m->settled[0] == 0xFFFFFFFFu ? 0xFFFFFFFFu : m->settled[0] * out->row;
        out->settled[1] = m->settled[1] == 0xFFFFFFFFu ? 0xFFFFFFFFu : m->settled[1] * out->row;
        out->settle = m->settle;
        if (m->tier == 2 || m->tier == 3) {     // rows of the first (shallowest, BFS order) states that fit 60 KB of shared memory
            uint64_t rows = (60ull * 1024) / (m->dfa.n_classes * 4ull);
            if (rows > m->dfa.n_states) rows = m->dfa.n_states;
            out->hot_entries = (uint32_t)(rows * m->dfa.n_classes);
        }
    }
    return 0;
}


int get_names_view(fqg_ctx* c, const fqg_matcher* cm, NamesView* out) {
    int rc = get_device_dfa(c, cm, nullptr);      // uploads the table on first use
    if (rc) return rc;
    fqg_matcher* m = const_cast<fqg_matcher*>(cm);
    std::lock_guard<std::mutex> lock(m->mu);
    const DeviceTables& t = m->dev[c->device];
    out->slots = t.slots; out->pool = t.pool; out->mask = m->names.n_slots / 4 - 1;
    return 0;
}

// SoA reads through whichever kernel(s) the matcher's tier asks for; b.invert / b.or_mask / b.count describe the final result
int match_soa_any(fqg_ctx* c, const fqg_matcher* m, SoaBatch b, cudaStream_t stream) {
    DeviceDfa dd;
    int rc = get_device_dfa(c, m, &dd);
    if (rc) return rc;
    if (m->kind == FQG_NAMES) {      // the spans are header lines (without '@'), not sequences
        NamesView nv;
        if ((rc = get_names_view(c, m, &nv))) return rc;
        CU_TRY(launch_match_names_soa(nv, b, c->sm_count, stream));
        count_launches(1);
        return FQG_OK;
    }
    if (m->split_lit) {
        // pass 1: literal members -> "found" bits in mask_out; pass 2: regex members, OR those bits in before invert
        DeviceDfa d1, d2;
        if ((rc = get_device_dfa(c, m->split_lit.get(), &d1)) || (rc = get_device_dfa(c, m->split_rest.get(), &d2))) return rc;
        SoaBatch b1 = b;
        b1.or_mask = nullptr; b1.count = nullptr; b1.invert = 0u; b1.pre_or_mask = nullptr;
        CU_TRY(launch_match_soa(d1, b1, c->sm_count, stream));
        SoaBatch b2 = b;
        b2.pre_or_mask = b.mask_out;
        CU_TRY(launch_match_soa(d2, b2, c->sm_count, stream));
        count_launches(2);
        return FQG_OK;
    }
    CU_TRY(launch_match_soa(dd, b, c->sm_count, stream));
    count_launches(1);
    return FQG_OK;
}

}  // namespace fqg

extern "C" {

const char* fqg_last_error(void) { return g_err.c_str(); }
const char* fqg_version(void) { return "fqgrep_b200 0.1.0 (sm_100a)"; }
uint64_t fqg_kernel_launches(void) { return g_launches.load(); }

int fqg_validate_fixed_pattern(const char* pattern, int protein) {
    if (!pattern) return fail(FQG_E_INVALID_ARG, "pattern is NULL");
    std::string msg = validate_fixed_pattern(pattern, protein != 0);
    if (msg.empty()) return FQG_OK;
    return fail(FQG_E_PATTERN, msg);
}

int fqg_matcher_new(int kind, const uint8_t* blob, const uint64_t* offsets, uint32_t n, uint32_t opts, fqg_matcher** out) {
    if (!out || (n && (!blob || !offsets))) return fail(FQG_E_INVALID_ARG, "NULL argument");
    *out = nullptr;
    std::vector<std::string> pats(n);
    for (uint32_t i = 0; i < n; i++) {
        if (offsets[i + 1] < offsets[i]) return fail(FQG_E_INVALID_ARG, "offsets must be non-decreasing");
        pats[i].assign((const char*)blob + offsets[i], offsets[i + 1] - offsets[i]);
    }
    auto m = std::make_unique<fqg_matcher>();
    m->kind = kind;
    m->opts = opts;
    const bool rc = (opts & FQG_REVERSE_COMPLEMENT) != 0;
    std::string err;
    switch (kind) {
        case FQG_FIXED:
            if (n != 1) return fail(FQG_E_INVALID_ARG, "FQG_FIXED takes exactly one pattern");
            // fallthrough
        case FQG_FIXED_SET:
            if (!compile_literal_set(pats, rc, &m->dfa, &err)) return fail(FQG_E_PATTERN, err);
            build_gram_filter(pats, rc, &m->gram);     // only used when the automaton does not fit shared memory
            break;
        case FQG_REGEX:
            if (n != 1) return fail(FQG_E_INVALID_ARG, "FQG_REGEX takes exactly one pattern");
            // fallthrough
        case FQG_REGEX_SET: {
            // a set of plain literals (e.g. -f patterns.txt without -F) is the same language as the fixed-string set
            std::vector<std::string> lits(pats.size());
            bool all_literal = !pats.empty();
            for (size_t i = 0; i < pats.size() && all_literal; i++) all_literal = regex_as_literal(pats[i], &lits[i]);
            if (all_literal) {
                if (!compile_literal_set(lits, rc, &m->dfa, &err)) return fail(FQG_E_PATTERN, err);
                build_gram_filter(lits, rc, &m->gram);
            } else {
                if (!compile_regex_set(pats, rc, &m->dfa, &err)) return fail(FQG_E_PATTERN, err);
                // Members with a small finite language ("AC[GT]{3}TT(A|G)CA": 16 strings) are as good as literals: when
                // every member is one and the strings are long enough, the q-gram filter serves the whole set (the
                // automaton above stays the exact fallback and the FASTQ kernel's table); otherwise the long-literal
                // members can still be split off from the real regexes (tier 5).
                std::vector<std::string> lit_members, rest_members, all_strings;
                bool all_finite = true;
                for (size_t i = 0; i < pats.size(); i++) {
                    std::vector<std::string> ls;
                    bool fin = regex_as_literals(pats[i], 64, &ls);
                    bool long_enough = fin;
                    for (auto& l : ls) if (l.size() < 8) long_enough = false;
                    if (fin) all_strings.insert(all_strings.end(), ls.begin(), ls.end()); else all_finite = false;
                    if (long_enough) lit_members.insert(lit_members.end(), ls.begin(), ls.end()); else rest_members.push_back(pats[i]);
                }
                if (all_finite && rest_members.empty()) {
                    build_gram_filter(all_strings, rc, &m->gram);      // usable or not is decided there (lengths, density)
                    break;
                }
                if (lit_members.size() >= 16 && !rest_members.empty()) {
                    auto a = std::make_unique<fqg_matcher>();
                    auto r = std::make_unique<fqg_matcher>();
                    a->kind = FQG_FIXED_SET; a->opts = opts & ~FQG_INVERT_MATCH;
                    r->kind = FQG_REGEX_SET; r->opts = opts;
                    std::string e2;
                    if (compile_literal_set(lit_members, rc, &a->dfa, &e2) && build_gram_filter(lit_members, rc, &a->gram) &&
                        compile_regex_set(rest_members, rc, &r->dfa, &e2)) {
                        build_tables(a.get());
                        build_tables(r.get());
                        m->split_lit = std::move(a);
                        m->split_rest = std::move(r);
                    }
                }
            }
            break;
        }
        case FQG_BITMASK:
            if (n != 1) return fail(FQG_E_INVALID_ARG, "FQG_BITMASK takes exactly one pattern");
            // fallthrough
        case FQG_BITMASK_SET:
            if (!compile_bitmask_set(pats, rc, &m->dfa, &err)) return fail(FQG_E_PATTERN, err);
            break;
        case FQG_NAMES:
            for (auto& p : pats) if (p.size() >= (1u << 24)) return fail(FQG_E_TOO_LARGE, "read name longer than 2^24 bytes");
            build_name_table(pats, &m->names);
            break;
        default:
            return fail(FQG_E_INVALID_ARG, "unknown matcher kind");
    }
    if (kind != FQG_NAMES) build_tables(m.get());
    // the split only pays when the whole-set automaton would walk L2 and the parts do not
    if (m->split_lit && !(m->tier == 2 && m->split_lit->tier != 2 && m->split_rest->tier != 2)) { m->split_lit.reset(); m->split_rest.reset(); }
    *out = m.release();
    return FQG_OK;
}

void fqg_matcher_free(fqg_matcher* m) {
    if (!m) return;
    if (m->split_lit) fqg_matcher_free(m->split_lit.release());
    if (m->split_rest) fqg_matcher_free(m->split_rest.release());
    for (auto& kv : m->dev) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(kv.first);
        cudaFree(kv.second.table);
        cudaFree(kv.second.byte_class);
        cudaFree(kv.second.slots);
        cudaFree(kv.second.pool);
        cudaFree(kv.second.g_bitmap); cudaFree(kv.second.g_kv); cudaFree(kv.second.g_pool);
        cudaSetDevice(prev);
    }
    delete m;
}

int fqg_matcher_dfa_info(const fqg_matcher* m, fqg_dfa_info* out) {
    if (!m || !out) return fail(FQG_E_INVALID_ARG, "NULL argument");
    if (m->kind == FQG_NAMES) return fail(FQG_E_INVALID_ARG, "name matchers have no DFA");
    out->n_states = m->dfa.n_states;
    out->n_classes = m->dfa.n_classes;
    out->start = m->dfa.start;
    out->accept_from = m->dfa.accept_from;
    out->tier = m->split_lit ? 5u : (uint32_t)m->tier;
    out->table_bytes = (uint32_t)((m->tier == 2 || m->tier == 3) ? m->t32.size() * 4 : m->t16.size() * 2);
    return FQG_OK;
}

int fqg_matcher_dfa_export(const fqg_matcher* m, uint8_t* byte_class, uint32_t* next) {
    if (!m) return fail(FQG_E_INVALID_ARG, "NULL argument");
    if (m->kind == FQG_NAMES) return fail(FQG_E_INVALID_ARG, "name matchers have no DFA");
    if (byte_class) std::memcpy(byte_class, m->dfa.byte_class, 256);
    if (next) std::memcpy(next, m->dfa.next.data(), m->dfa.next.size() * 4);
    return FQG_OK;
}

int fqg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int fqg_ctx_create(int device, fqg_ctx** out) {
    if (!out) return fail(FQG_E_INVALID_ARG, "NULL argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(FQG_E_CUDA, std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(FQG_E_INVALID_ARG, "device index out of range");
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(FQG_E_CUDA, "fqgrep_b200 is built for sm_100a (Blackwell B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
    auto c = std::make_unique<fqg_ctx>();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c.release();
    return FQG_OK;
}

void fqg_ctx_destroy(fqg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    lanes_free(c);
    for (auto& b : c->d_buf) cudaFree(b);
    cudaStreamDestroy(c->stream);
    delete c;
}

int fqg_alloc_pinned(size_t bytes, void** out) {
    if (!out) return fail(FQG_E_INVALID_ARG, "NULL argument");
    CU_TRY(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
    return FQG_OK;
}
void fqg_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

int fqg_match_bases_device(fqg_ctx* c, const fqg_matcher* m, const uint8_t* d_bases, const uint32_t* d_start, const uint32_t* d_end,
                           uint64_t n_reads, uint32_t* d_mask_out, const uint32_t* d_or_mask, uint64_t* d_count, uint32_t avg_read_len,
                           void* stream) {
    if (!c || !m || !d_mask_out || (n_reads && (!d_bases || !d_start || !d_end))) return fail(FQG_E_INVALID_ARG, "NULL argument");
    if (n_reads >= (1ull << 32)) return fail(FQG_E_TOO_LARGE, "more than 2^32-1 reads in one call; split the batch");
    if (n_reads == 0) return FQG_OK;
    CU_TRY(cudaSetDevice(c->device));
    SoaBatch b{d_bases, d_start, d_end, (uint32_t)n_reads, d_mask_out, d_or_mask, reinterpret_cast<unsigned long long*>(d_count),
               (m->opts & FQG_INVERT_MATCH) ? 1u : 0u, avg_read_len, nullptr, nullptr};
    return match_soa_any(c, m, b, (cudaStream_t)stream);
}

int fqg_match_bases_host(fqg_ctx* c, const fqg_matcher* m, const uint8_t* bases, const uint64_t* offsets, uint64_t n_reads, uint32_t* mask_out,
                         uint64_t* n_selected) {
    if (!c || !m || (n_reads && (!bases || !offsets))) return fail(FQG_E_INVALID_ARG, "NULL argument");
    if (n_selected) *n_selected = 0;
    CU_TRY(cudaSetDevice(c->device));
    uint64_t total = 0;
    // chunks of <= 1 GiB of bases and a multiple of 32 reads, so 32-bit offsets and mask words line up
    const uint64_t kChunkBytes = 1ull << 30;
    uint64_t r0 = 0;
    std::vector<uint32_t> rel;
    while (r0 < n_reads) {
        uint64_t r1 = r0;
        while (r1 < n_reads && offsets[r1 + 1] - offsets[r0] <= kChunkBytes) r1++;
        if (r1 == r0) return fail(FQG_E_TOO_LARGE, "a single read exceeds 1 GiB");
        if (r1 < n_reads) {
            // chunks must hold a multiple of 32 reads so that their mask words line up with the caller's
            r1 = r0 + ((r1 - r0) & ~31ull);
            if (r1 == r0) return fail(FQG_E_TOO_LARGE, "32 consecutive reads exceed 1 GiB; use fqg_match_bases_device with your own batching");
        }
        const uint64_t n = r1 - r0, nbytes = offsets[r1] - offsets[r0], words = (n + 31) / 32;
        rel.resize(n + 1);
        for (uint64_t i = 0; i <= n; i++) {
            if (offsets[r0 + i] < offsets[r0] || (i && offsets[r0 + i] < offsets[r0 + i - 1])) return fail(FQG_E_INVALID_ARG, "offsets must be non-decreasing");
            rel[i] = (uint32_t)(offsets[r0 + i] - offsets[r0]);
        }
        void *d_bases, *d_off, *d_mask, *d_cnt;
        int rc;
        if ((rc = ensure_dev(c, 0, nbytes + 64, &d_bases)) || (rc = ensure_dev(c, 1, (n + 1) * 4, &d_off)) || (rc = ensure_dev(c, 2, words * 4 + 8, &d_mask)) ||
            (rc = ensure_dev(c, 3, 8, &d_cnt)))
            return rc;
        CU_TRY(cudaMemcpyAsync(d_bases, bases + offsets[r0], nbytes, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemcpyAsync(d_off, rel.data(), (n + 1) * 4, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaMemsetAsync(d_cnt, 0, 8, c->stream));
        rc = fqg_match_bases_device(c, m, (const uint8_t*)d_bases, (const uint32_t*)d_off, (const uint32_t*)d_off + 1, n, (uint32_t*)d_mask, nullptr,
                                    (uint64_t*)d_cnt, (uint32_t)(n ? nbytes / n : 0), c->stream);
        if (rc) return rc;
        uint64_t cnt = 0;
        if (mask_out) CU_TRY(cudaMemcpyAsync(mask_out + r0 / 32, d_mask, words * 4, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        total += cnt;
        r0 = r1;
    }
    if (n_selected) *n_selected = total;
    return FQG_OK;
}

int fqg_synth_reads_device(fqg_ctx* c, uint8_t* d_out, uint64_t n_reads, uint64_t first_index, uint32_t read_len, int fastq, uint64_t seed,
                           const uint8_t* plant_blob, const uint32_t* plant_offsets, uint32_t n_plant, double plant_rate, uint64_t plant_seed,
                           void* stream) {
    if (!c || !d_out || read_len == 0) return fail(FQG_E_INVALID_ARG, "bad argument");
    CU_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    SynthParams p{};
    p.out = d_out; p.n_reads = n_reads; p.first = first_index; p.read_len = read_len; p.fastq = fastq ? 1u : 0u;
    p.seed = seed; p.plant_seed = plant_seed; p.n_plant = n_plant;
    double thr = plant_rate * 18446744073709551616.0;
    p.plant_threshold = plant_rate >= 1.0 ? ~0ull : (plant_rate <= 0 ? 0ull : (uint64_t)thr);
    void *d_blob = nullptr, *d_offs = nullptr;
    if (n_plant) {
        int rc;
        size_t blob_bytes = plant_offsets[n_plant];
        if ((rc = ensure_dev(c, 6, blob_bytes + 16, &d_blob)) || (rc = ensure_dev(c, 7, (n_plant + 1) * 4, &d_offs))) return rc;
        CU_TRY(cudaMemcpyAsync(d_blob, plant_blob, blob_bytes, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(d_offs, plant_offsets, (n_plant + 1) * 4, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));   // host arrays may be temporaries
    }
    p.plant_blob = (const uint8_t*)d_blob;
    p.plant_offs = (const uint32_t*)d_offs;
    CU_TRY(launch_synth(p, st));
    count_launches(1);
    return FQG_OK;
}

int fqg_grep_fastq_device(fqg_ctx* c, const fqg_matcher* m, int mode, const uint8_t* d_r1, size_t n1, const uint8_t* d_r2, size_t n2,
                          uint32_t* d_unit_mask_out, size_t unit_mask_words, fqg_result* res, void* stream) {
    if (!c || !m || !res) return fail(FQG_E_INVALID_ARG, "NULL argument");
    std::memset(res, 0, sizeof *res);
    BatchIn in;
    in.mode = mode;
    in.r[0] = d_r1; in.n[0] = n1; in.r[1] = d_r2; in.n[1] = n2;
    in.on_device = true;
    in.user_stream = (cudaStream_t)stream;
    BatchOut out;
    FastqLane* lane = lane_lease(c);
    struct Release { fqg_ctx* c; FastqLane* l; ~Release() { lane_release(c, l); } } release{c, lane};
    int rc = grep_fastq_batch(c, lane, m, in, &out);
    if (rc) return rc;
    *res = out.res;
    if (d_unit_mask_out) {
        if (out.mask_words > unit_mask_words) return fail(FQG_E_INVALID_ARG, "unit_mask_out too small: need " + std::to_string(out.mask_words) + " words");
        if (out.mask_words) {
            CU_TRY(cudaMemcpyAsync(d_unit_mask_out, out.unit_mask_dev, out.mask_words * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
            CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        }
    }
    return FQG_OK;
}

int fqg_iupac_expand(const char* pattern, int to_regex, char* out, size_t cap, size_t* out_len) {
    if (!pattern || !out_len) return fail(FQG_E_INVALID_ARG, "NULL argument");
    std::string joined;
    if (to_regex) joined = iupac_to_regex(pattern);
    else {
        std::vector<std::string> v;
        std::string err;
        if (!expand_iupac_fixed(pattern, &v, &err)) return fail(FQG_E_PATTERN, err);
        for (size_t i = 0; i < v.size(); i++) { if (i) joined.push_back('\n'); joined += v[i]; }
    }
    *out_len = joined.size();
    if (out) {
        if (cap < joined.size()) return fail(FQG_E_INVALID_ARG, "output buffer too small");
        std::memcpy(out, joined.data(), joined.size());
    }
    return FQG_OK;
}

int fqg_read_input(const char* path, int decompress, int pinned, uint8_t** out, size_t* n_out) {
    if (!path || !out || !n_out) return fail(FQG_E_INVALID_ARG, "NULL argument");
    std::string err;
    int rc = read_input(path, decompress, pinned, out, n_out, &err);
    if (rc) return fail(rc, err);
    return FQG_OK;
}
void fqg_free_input(uint8_t* p, int pinned) { free_input(p, pinned); }

int fqg_reader_open(const char* path, int decompress, fqg_reader** out) {
    if (!path || !out) return fail(FQG_E_INVALID_ARG, "NULL argument");
    std::string err;
    InputReader* r = nullptr;
    const int rc = reader_open(path, decompress, &r, &err);
    if (rc) return fail(rc, err);
    *out = reinterpret_cast<fqg_reader*>(r);
    return FQG_OK;
}
int fqg_reader_read(fqg_reader* r, uint8_t* dst, size_t cap, size_t* n_read, int* eof) {
    if (!r || !n_read || !eof || (cap && !dst)) return fail(FQG_E_INVALID_ARG, "NULL argument");
    std::string err;
    bool e = false;
    const int rc = reader_read(reinterpret_cast<InputReader*>(r), dst, cap, n_read, &e, &err);
    *eof = e ? 1 : 0;
    if (rc) return fail(rc, err);
    return FQG_OK;
}
void fqg_reader_close(fqg_reader* r) { reader_close(reinterpret_cast<InputReader*>(r)); }
int fqg_is_gzip_path(const char* path) { return path && is_gzip_path(path) ? 1 : 0; }
int fqg_is_fastq_path(const char* path) { return path && is_fastq_path(path) ? 1 : 0; }

int fqg_index_fastq_device(fqg_ctx* c, const uint8_t* d_fastq, size_t n, uint64_t* d_seq_start, uint32_t* d_seq_len, uint64_t capacity,
                           uint64_t* n_records, void* stream) {
    if (!c || !d_seq_start || !d_seq_len || !n_records || (n && !d_fastq)) return fail(FQG_E_INVALID_ARG, "NULL argument");
    IndexOut idx;
    idx.seq_start = reinterpret_cast<unsigned long long*>(d_seq_start);
    idx.seq_len = d_seq_len;
    idx.cap = capacity;
    BatchIn in;
    in.mode = FQG_SINGLE;
    in.r[0] = d_fastq; in.n[0] = n;
    in.on_device = true;
    in.user_stream = (cudaStream_t)stream;
    in.idx = &idx;
    BatchOut out;
    FastqLane* lane = lane_lease(c);
    struct Release { fqg_ctx* c; FastqLane* l; ~Release() { lane_release(c, l); } } release{c, lane};
    int rc = grep_fastq_batch(c, lane, nullptr, in, &out);
    if (rc) return rc;
    const fqg_result& res = out.res;
    *n_records = res.n_records;
    if (res.n_records > capacity) return fail(FQG_E_INVALID_ARG, "index arrays too small: " + std::to_string(res.n_records) + " records");
    return FQG_OK;
}

int fqg_compact_bases_device(fqg_ctx* c, const uint8_t* d_fastq, const uint64_t* d_seq_start, const uint32_t* d_seq_len, uint64_t n_records,
                             uint8_t* d_bases_out, uint32_t* d_offsets_out, void* stream) {
    if (!c || !d_bases_out || !d_offsets_out || (n_records && (!d_fastq || !d_seq_start || !d_seq_len))) return fail(FQG_E_INVALID_ARG, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    void* scratch;
    int rc = ensure_dev(c, 6, compact_scratch_words(n_records) * 4, &scratch);
    if (rc) return rc;
    CU_TRY(launch_compact_bases(d_fastq, reinterpret_cast<const unsigned long long*>(d_seq_start), d_seq_len, n_records, d_bases_out, d_offsets_out,
                                (uint32_t*)scratch, (cudaStream_t)stream));
    count_launches(n_records ? 4 : 0);
    return FQG_OK;
}

int fqg_fastq_split(const uint8_t* buf, size_t n, uint32_t parts, uint64_t* cuts) {
    if (!cuts || parts == 0 || (n && !buf)) return fail(FQG_E_INVALID_ARG, "bad argument");
    cuts[0] = 0;
    for (uint32_t k = 1; k < parts; k++) {
        size_t c = find_record_start(buf, n, (size_t)((unsigned __int128)n * k / parts));
        cuts[k] = c < cuts[k - 1] ? cuts[k - 1] : c;
    }
    cuts[parts] = n;
    return FQG_OK;
}

int fqg_write_selected(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* unit_mask, uint8_t* out,
                       size_t* out_len, uint8_t* out2, size_t* out2_len) {
    std::string err;
    int rc = write_selected_host(mode, r1, n1, r2, n2, unit_mask, out, out_len, out2, out2_len, &err);
    if (rc) return fail(rc, err);
    return FQG_OK;
}

}  // extern "C"
