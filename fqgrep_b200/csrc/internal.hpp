// internal.hpp -- what api.cu, fastq_pipeline.cu and stream.cu share: the opaque handles of the C ABI and the helpers
// that turn a matcher into its device form.  Nothing here is exported.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fqgrep_b200.h"
#include "automata.hpp"
#include "fastq_host.hpp"
#include "kernels.hpp"

namespace fqg {

int fail(int code, const std::string& msg);      // sets the thread-local error text, returns `code`
void count_launches(uint64_t n);                 // fqg_kernel_launches() evidence counter

#define CU_TRY(expr)                                                                                              \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess) return fqg::fail(FQG_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct DeviceTables {
    void* table = nullptr;
    uint8_t* byte_class = nullptr;
    // q-gram filter
    uint32_t *g_bitmap = nullptr, *g_kv = nullptr;
    uint8_t* g_pool = nullptr;
    // names
    uint64_t* slots = nullptr;
    uint8_t* pool = nullptr;
};

struct FastqLane;      // fastq_pipeline.cu: everything one in-flight FASTQ batch needs (streams, events, scratch)

}  // namespace fqg

struct fqg_matcher {
    int kind = 0;
    uint32_t opts = 0;
    fqg::Dfa dfa;
    fqg::NameTable names;
    fqg::GramFilter gram;
    int tier = 0;
    uint32_t row = 0;
    std::vector<uint16_t> t16;
    std::vector<uint32_t> t32;
    uint8_t lut[256] = {0};   // byte -> class in the form the tier's kernel wants
    uint32_t settled[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};   // states that only lead to themselves (see DeviceDfa), found in build_tables
    uint32_t settle = 0;
    std::mutex mu;
    std::map<int, fqg::DeviceTables> dev;
    // Mixed -e/-f sets (many plain literals + a few real regexes) determinise into automata far too large for shared
    // memory.  For the SoA kernels they are split: `split_lit` (the literal members: q-gram filter) runs first and leaves
    // its "found" bits in the mask, `split_rest` (the regex members: small automaton) ORs them in before invert_match.
    // `dfa` above is still the automaton of the WHOLE set (fused FASTQ kernel, introspection).
    std::unique_ptr<fqg_matcher> split_lit, split_rest;
};

struct fqg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // grow-only device scratch of the SoA host entry point and the synthetic generator
    void* d_buf[8] = {nullptr};
    size_t d_cap[8] = {0};
    // FASTQ batches in flight: lane 0 serves the one-call entry points, fqg_stream_* uses as many as its ring has slots
    std::vector<fqg::FastqLane*> lanes;      // owned; freed by lanes_free (fastq_pipeline.cu)
    std::mutex lanes_mu;
};

namespace fqg {

int ensure_dev(fqg_ctx* c, int slot, size_t bytes, void** out);
int get_device_dfa(fqg_ctx* c, const fqg_matcher* m, DeviceDfa* out);
int get_names_view(fqg_ctx* c, const fqg_matcher* m, NamesView* out);
int match_soa_any(fqg_ctx* c, const fqg_matcher* m, SoaBatch b, cudaStream_t stream);

// ---- one FASTQ batch through the fused kernel (fastq_pipeline.cu) --------------------------------------------------------
struct IndexOut {                    // K0: optional per-record index of stream 1 (device arrays owned by the caller)
    unsigned long long* seq_start = nullptr;
    uint32_t* seq_len = nullptr;
    uint64_t cap = 0;
};
struct BatchIn {
    int mode = FQG_SINGLE;
    const uint8_t* r[2] = {nullptr, nullptr};
    size_t n[2] = {0, 0};
    bool on_device = false;
    bool want_spans = false;             // also compute the selected records' spans (fqg_span) for the write-back
    unsigned min_record_bytes = 32;      // capacity of the per-record arrays: n / min_record_bytes records
    unsigned keep_newline_mask = 0;      // see lane_submit: keep one trailing line terminator of stream i (empty last quality line)
    const IndexOut* idx = nullptr;
    cudaStream_t user_stream = nullptr;  // compute stream of device-resident inputs
};
struct BatchOut {
    fqg_result res{};
    const uint32_t* unit_mask = nullptr;     // host batches: pinned staging of the lane, valid until its next submit
    const uint32_t* unit_mask_dev = nullptr; // device copy, valid until the lane's next submit
    uint64_t mask_words = 0;
    const fqg_span* spans = nullptr;         // pinned, want_spans only: selected records in output order
    uint64_t n_spans = 0;
    bool capacity_hit = false;               // the per-record arrays were too small: retry with a smaller min_record_bytes
    unsigned stripped_mask = 0;              // streams whose trailing line terminators were dropped
};
FastqLane* lane_lease(fqg_ctx* c);                           // an idle lane of the context (created when all are taken)
void lane_release(fqg_ctx* c, FastqLane* lane);
void lanes_free(fqg_ctx* c);
// m == nullptr: index only (parse, validate, index; no matching).  Asynchronous: returns once everything is enqueued.
int lane_submit(fqg_ctx* c, FastqLane* lane, const fqg_matcher* m, const BatchIn& in);
// Waits for the lane's batch, decodes errors like the reference's reader would raise them, fills `out`.
int lane_wait(fqg_ctx* c, FastqLane* lane, BatchOut* out);
// lane_wait plus the retries a batch can need (dense records, empty last quality line); result left in the lane
int lane_finish(fqg_ctx* c, FastqLane* lane, const fqg_matcher* m, BatchOut* out);
// submit + finish
int grep_fastq_batch(fqg_ctx* c, FastqLane* lane, const fqg_matcher* m, BatchIn in, BatchOut* out);

}  // namespace fqg
