// api.cu -- the C ABI (include/fqgrep_b200.h): matcher construction, device contexts, match entry points.
// No CPU matching path exists in this library: every match call launches the CUDA kernels or fails.
#include "../../include/fqgrep_b200.h"

#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "automata.hpp"
#include "fastq_host.hpp"
#include "kernels.hpp"

using namespace fqg;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU_TRY(expr)                                                                                         \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) return fail(FQG_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct DeviceTables {
    void* table = nullptr;
    uint8_t* byte_class = nullptr;
    // names
    uint64_t* slots = nullptr;
    uint32_t* slot_off = nullptr;
    uint8_t* pool = nullptr;
};

}  // namespace

struct fqg_matcher {
    int kind = 0;
    uint32_t opts = 0;
    Dfa dfa;
    NameTable names;
    int tier = 0;
    uint32_t row = 0;
    std::vector<uint16_t> t16;
    std::vector<uint32_t> t32;
    uint8_t lut[256] = {0};   // byte -> class in the form the tier's kernel wants
    std::mutex mu;
    std::map<int, DeviceTables> dev;
};

struct fqg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    // grow-only device scratch
    void* d_buf[8] = {nullptr};
    size_t d_cap[8] = {0};
    void* h_pinned = nullptr;
    size_t h_cap = 0;
    cudaEvent_t ev[6] = {nullptr};
    FastqDeviceState fq;
};

namespace {

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

void build_tables(fqg_matcher* m) {
    const Dfa& d = m->dfa;
    const uint64_t n = d.n_states, nc = d.n_classes;
    std::memcpy(m->lut, d.byte_class, 256);
    if (n * kTier0Row * 2 <= 40 * 1024) {
        m->tier = 0;
        m->row = kTier0Row * 2;   // states are byte offsets of their rows
        m->t16.assign(((n * kTier0Row + 7) / 8) * 8, 0);
        for (uint64_t s = 0; s < n; s++)
            for (int b = 0; b < 256; b++) m->t16[s * kTier0Row + b] = (uint16_t)(d.next[s * nc + d.byte_class[b]] * kTier0Row * 2);
    } else if (n * nc * 2 <= 65535 && nc <= 127) {
        m->tier = 1;
        m->row = (uint32_t)nc * 2;
        m->t16.assign(((n * nc + 7) / 8) * 8, 0);
        for (uint64_t i = 0; i < n * nc; i++) m->t16[i] = (uint16_t)(d.next[i] * nc * 2);
        for (int b = 0; b < 256; b++) m->lut[b] = (uint8_t)(d.byte_class[b] * 2);
    } else {
        m->tier = 2;
        m->row = (uint32_t)nc;
        m->t32.resize(((n * nc + 3) / 4) * 4);
        for (uint64_t i = 0; i < n * nc; i++) m->t32[i] = (uint32_t)(d.next[i] * nc);
    }
}

int get_device_dfa(fqg_ctx* c, const fqg_matcher* cm, DeviceDfa* out) {
    fqg_matcher* m = const_cast<fqg_matcher*>(cm);
    std::lock_guard<std::mutex> lock(m->mu);
    auto it = m->dev.find(c->device);
    if (it == m->dev.end()) {
        DeviceTables t;
        if (m->kind == FQG_NAMES) {
            CU_TRY(cudaMalloc(&t.slots, m->names.slots.size() * 8));
            CU_TRY(cudaMalloc(&t.slot_off, m->names.slot_off.size() * 4));
            CU_TRY(cudaMalloc(&t.pool, m->names.pool.size()));
            CU_TRY(cudaMemcpy(t.slots, m->names.slots.data(), m->names.slots.size() * 8, cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(t.slot_off, m->names.slot_off.data(), m->names.slot_off.size() * 4, cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(t.pool, m->names.pool.data(), m->names.pool.size(), cudaMemcpyHostToDevice));
        } else {
            const void* src = m->tier == 2 ? (const void*)m->t32.data() : (const void*)m->t16.data();
            size_t bytes = m->tier == 2 ? m->t32.size() * 4 : m->t16.size() * 2;
            CU_TRY(cudaMalloc(&t.table, bytes + 16));
            CU_TRY(cudaMemcpy(t.table, src, bytes, cudaMemcpyHostToDevice));
            CU_TRY(cudaMalloc(&t.byte_class, 256));
            CU_TRY(cudaMemcpy(t.byte_class, m->lut, 256, cudaMemcpyHostToDevice));
        }
        it = m->dev.emplace(c->device, t).first;
    }
    if (out) {
        out->tier = m->tier;
        out->table = it->second.table;
        out->table_bytes = (uint32_t)(m->tier == 2 ? m->t32.size() * 4 : m->t16.size() * 2);
        out->byte_class = it->second.byte_class;
        out->row = m->row;
        out->start = m->dfa.start * m->row;
        out->accept_from = m->dfa.accept_from * m->row;
        out->n_states = m->dfa.n_states;
    }
    return 0;
}

}  // namespace

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
            break;
        case FQG_REGEX:
            if (n != 1) return fail(FQG_E_INVALID_ARG, "FQG_REGEX takes exactly one pattern");
            // fallthrough
        case FQG_REGEX_SET:
            if (!compile_regex_set(pats, rc, &m->dfa, &err)) return fail(FQG_E_PATTERN, err);
            break;
        case FQG_NAMES:
            for (auto& p : pats) if (p.size() >= (1u << 24)) return fail(FQG_E_TOO_LARGE, "read name longer than 2^24 bytes");
            build_name_table(pats, &m->names);
            break;
        default:
            return fail(FQG_E_INVALID_ARG, "unknown matcher kind");
    }
    if (kind != FQG_NAMES) build_tables(m.get());
    *out = m.release();
    return FQG_OK;
}

void fqg_matcher_free(fqg_matcher* m) {
    if (!m) return;
    for (auto& kv : m->dev) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(kv.first);
        cudaFree(kv.second.table);
        cudaFree(kv.second.byte_class);
        cudaFree(kv.second.slots);
        cudaFree(kv.second.slot_off);
        cudaFree(kv.second.pool);
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
    out->tier = (uint32_t)m->tier;
    out->table_bytes = (uint32_t)(m->tier == 2 ? m->t32.size() * 4 : m->t16.size() * 2);
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
    CU_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : c->ev) CU_TRY(cudaEventCreate(&ev));
    *out = c.release();
    return FQG_OK;
}

void fqg_ctx_destroy(fqg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    fastq_state_free(&c->fq);
    for (auto& b : c->d_buf) cudaFree(b);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
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
    if (m->kind == FQG_NAMES) return fail(FQG_E_INVALID_ARG, "read-name matchers need headers: use the FASTQ entry points");
    if (n_reads >= (1ull << 32)) return fail(FQG_E_TOO_LARGE, "more than 2^32-1 reads in one call; split the batch");
    if (n_reads == 0) return FQG_OK;
    CU_TRY(cudaSetDevice(c->device));
    DeviceDfa dd;
    int rc = get_device_dfa(c, m, &dd);
    if (rc) return rc;
    SoaBatch b{d_bases, d_start, d_end, (uint32_t)n_reads, d_mask_out, d_or_mask, reinterpret_cast<unsigned long long*>(d_count),
               (m->opts & FQG_INVERT_MATCH) ? 1u : 0u, avg_read_len};
    CU_TRY(launch_match_soa(dd, b, c->sm_count, (cudaStream_t)stream));
    g_launches++;
    return FQG_OK;
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
        if (r1 < n_reads) { r1 = r0 + ((r1 - r0) & ~31ull); if (r1 == r0) r1 = r0 + 1; }
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
    g_launches++;
    return FQG_OK;
}

int fqg_grep_fastq_device(fqg_ctx* c, const fqg_matcher* m, int mode, const uint8_t* d_r1, size_t n1, const uint8_t* d_r2, size_t n2,
                          uint32_t* d_unit_mask_out, size_t unit_mask_words, fqg_result* res, void* stream) {
    if (!c || !m || !res) return fail(FQG_E_INVALID_ARG, "NULL argument");
    return fail(FQG_E_INVALID_ARG, "fqg_grep_fastq_device: not implemented in this build");
}

int fqg_grep_fastq_host(fqg_ctx* c, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2,
                        uint32_t* unit_mask_out, size_t unit_mask_words, fqg_result* res) {
    if (!c || !m || !res) return fail(FQG_E_INVALID_ARG, "NULL argument");
    return fail(FQG_E_INVALID_ARG, "fqg_grep_fastq_host: not implemented in this build");
}

int fqg_write_selected(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* unit_mask, uint8_t* out,
                       size_t* out_len, uint8_t* out2, size_t* out2_len) {
    std::string err;
    int rc = write_selected_host(mode, r1, n1, r2, n2, unit_mask, out, out_len, out2, out2_len, &err);
    if (rc) return fail(rc, err);
    return FQG_OK;
}

}  // extern "C"
