// fqgrep_b200.hpp -- C++ host side above the C ABI (include/fqgrep_b200.h), mirroring the reference's own interface
// for the read-matching path name by name, so a maintainer of the (Rust) reference can read it side by side:
//
//   MatcherOpts                         src/lib/matcher.rs:20-28
//   Matcher::{opts, read_match}         src/lib/matcher.rs:150-193 (trait), :631-638 (QueryNameMatcher)
//   MatcherFactory::new_matcher         src/lib/matcher.rs:645-676
//   MatcherFactory::new_query_name_matcher   :679-685
//   validate_fixed_pattern              src/lib/matcher.rs:128-147
//   expand_iupac                        src/main.rs:315-372 (over src/lib/mod.rs:70-118)
//   read_patterns / read_query_names    src/main.rs:246-271
//   Opts, fqgrep_from_opts, fqgrep      src/main.rs:142-244, 375-675, 295-311
//
//   spawn_reader + the batch readers    src/main.rs:61-87, src/lib/seq_io.rs:34-81, 121-199 -> InputStream / BatchFeeder
//   ordered_parallel_* + func closures  src/lib/seq_io.rs:287-368, src/main.rs:558-575, 600-614, 641-657 -> run_input
//
// Inputs are STREAMED: reader threads fill the pinned slot buffers of a ring of batches in flight on the GPU(s)
// (fqg_stream_*), batches are cut at record ends with mates in lock step (fqg_fastq_cut), results come back in
// input order and the selected records are copied out from the device's record spans.  Host and device memory are
// bounded by the ring whatever the input size.
//
// Header-only; every matching call goes to the GPU through the C ABI (no host matching path).  Errors are exceptions
// carrying the C status and the reference's user-visible message; FqgrepError::exit_code() is what the reference
// binary would exit with.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <thread>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fqgrep_b200.h"

namespace fqgrep_b200 {

struct FqgrepError : std::runtime_error {
    int code;
    FqgrepError(int c, const std::string& msg) : std::runtime_error(msg), code(c) {}
    // src/main.rs:295-311: anyhow errors exit 2, panics (malformed FASTQ, mismatched pairs) exit 101
    int exit_code() const { return (code == FQG_E_FASTQ || code == FQG_E_PAIR_MISMATCH) ? 101 : 2; }
};
inline void check(int rc) {
    if (rc != FQG_OK) throw FqgrepError(rc, fqg_last_error());
}

struct MatcherOpts {
    bool invert_match = false;
    bool reverse_complement = false;
    uint32_t bits() const { return (invert_match ? FQG_INVERT_MATCH : 0u) | (reverse_complement ? FQG_REVERSE_COMPLEMENT : 0u); }
};

inline void validate_fixed_pattern(const std::string& pattern, bool protein) { check(fqg_validate_fixed_pattern(pattern.c_str(), protein ? 1 : 0)); }

// One device context per (process, GPU), created on first use.
inline fqg_ctx* context(int device = 0) {
    static fqg_ctx* ctxs[64] = {};
    if (device < 0 || device >= 64) throw FqgrepError(FQG_E_INVALID_ARG, "bad device index");
    if (!ctxs[device]) check(fqg_ctx_create(device, &ctxs[device]));
    return ctxs[device];
}

// Host bytes of one input (file, stdin, optionally gunzipped) in pinned memory: spawn_reader, src/main.rs:61-87.
class Input {
  public:
    Input(const std::string& path, bool decompress) { check(fqg_read_input(path.c_str(), decompress ? 1 : 0, 1, &p_, &n_)); }
    ~Input() { if (p_) fqg_free_input(p_, 1); }
    Input(const Input&) = delete;
    Input& operator=(const Input&) = delete;
    const uint8_t* data() const { return p_; }
    size_t size() const { return n_; }

  private:
    uint8_t* p_ = nullptr;
    size_t n_ = 0;
};

struct GrepResult {
    fqg_result res{};
    std::vector<uint32_t> unit_mask;      // bit k = record (single) / pair (paired, interleaved) k selected
};

// `Box<dyn Matcher + Sync + Send>` of the reference: immutable once built, shareable between host threads.
class Matcher {
  public:
    Matcher(int kind, const std::vector<std::string>& patterns, MatcherOpts opts) : kind_(kind), opts_(opts) {
        std::string blob;
        std::vector<uint64_t> offs(1, 0);
        for (auto& p : patterns) { blob += p; offs.push_back(blob.size()); }
        check(fqg_matcher_new(kind, reinterpret_cast<const uint8_t*>(blob.data()), offs.data(), (uint32_t)patterns.size(), opts.bits(), &h_));
    }
    ~Matcher() { if (h_) fqg_matcher_free(h_); }
    Matcher(const Matcher&) = delete;
    Matcher& operator=(const Matcher&) = delete;

    MatcherOpts opts() const { return opts_; }
    int kind() const { return kind_; }
    const fqg_matcher* handle() const { return h_; }

    // Batch seam: parse + match one input (or one R1/R2 pair of inputs) of raw FASTQ bytes in host memory.
    GrepResult grep_fastq(const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, int mode, int device = 0) const {
        GrepResult g;
        g.unit_mask.assign(n1 / 8 / 32 + 2, 0);
        check(fqg_grep_fastq_host(context(device), h_, mode, r1, n1, r2, n2, g.unit_mask.data(), g.unit_mask.size(), &g.res));
        return g;
    }
    // SoA batch in host memory: read i = bases[offsets[i] .. offsets[i+1]) (header lines for a read-name matcher)
    uint64_t match_bases(const uint8_t* bases, const uint64_t* offsets, uint64_t n_reads, std::vector<uint32_t>* mask, int device = 0) const {
        uint64_t n_selected = 0;
        if (mask) mask->assign((n_reads + 31) / 32 + 1, 0);
        check(fqg_match_bases_host(context(device), h_, bases, offsets, n_reads, mask ? mask->data() : nullptr, &n_selected));
        return n_selected;
    }
    // Matcher::read_match for ONE record (tests and known-answer vectors; a 1-record batch on the device).
    bool read_match(const std::string& seq, const std::string& head = "", int device = 0) const {
        if (kind_ == FQG_NAMES) {
            const std::string rec = "@" + head + "\n" + seq + "\n+\n" + std::string(seq.size(), 'I') + "\n";
            GrepResult g = grep_fastq(reinterpret_cast<const uint8_t*>(rec.data()), rec.size(), nullptr, 0, FQG_SINGLE, device);
            return g.res.n_selected == 1;
        }
        const uint64_t offs[2] = {0, seq.size()};
        return match_bases(reinterpret_cast<const uint8_t*>(seq.data()), offs, 1, nullptr, device) == 1;
    }

  private:
    fqg_matcher* h_ = nullptr;
    int kind_;
    MatcherOpts opts_;
};

// `bitencs` (Vec<BitEnc> in the reference) is the list of IUPAC pattern strings; the 4-bit encoding happens in the library.
struct MatcherFactory {
    static std::unique_ptr<Matcher> new_matcher(const std::string* pattern, bool fixed_strings, const std::vector<std::string>& regexp,
                                                const std::vector<std::string>* bitencs, MatcherOpts match_opts) {
        if (bitencs) return std::make_unique<Matcher>(pattern ? FQG_BITMASK : FQG_BITMASK_SET, *bitencs, match_opts);
        if (pattern) return std::make_unique<Matcher>(fixed_strings ? FQG_FIXED : FQG_REGEX, std::vector<std::string>{*pattern}, match_opts);
        return std::make_unique<Matcher>(fixed_strings ? FQG_FIXED_SET : FQG_REGEX_SET, regexp, match_opts);
    }
    static std::unique_ptr<Matcher> new_query_name_matcher(const std::set<std::string>& query_names, MatcherOpts match_opts) {
        return std::make_unique<Matcher>(FQG_NAMES, std::vector<std::string>(query_names.begin(), query_names.end()), match_opts);
    }
};

// ---- IUPAC handling (src/lib/mod.rs:70-118, src/main.rs:315-372) ------------------------------------------------
inline std::string iupac_expand_raw(const std::string& pattern, bool to_regex) {
    size_t n = 0;
    check(fqg_iupac_expand(pattern.c_str(), to_regex ? 1 : 0, nullptr, 0, &n));
    std::string out(n, '\0');
    check(fqg_iupac_expand(pattern.c_str(), to_regex ? 1 : 0, out.empty() ? nullptr : &out[0], n, &n));
    out.resize(n);
    return out;
}
inline std::vector<std::string> expand_iupac_fixed_pattern(const std::string& pattern) {
    std::vector<std::string> v;
    std::stringstream ss(iupac_expand_raw(pattern, false));
    for (std::string line; std::getline(ss, line, '\n');) v.push_back(line);
    return v;
}
inline std::string expand_iupac_regex(const std::string& pattern) { return iupac_expand_raw(pattern, true); }

enum class IupacOption { Never, Expand, Regex, BitMask };

// ---- the driver (src/main.rs) ---------------------------------------------------------------------------------------
struct Opts {
    std::vector<std::string> args;          // [pattern] files...
    std::vector<std::string> regexp;        // -e
    bool has_file = false;                  // -f
    std::string file;
    bool has_read_names_file = false;       // -N
    std::string read_names_file;
    bool fixed_strings = false;             // -F
    bool invert_match = false;              // -v
    bool reverse_complement = false;
    bool paired = false;
    bool count = false;                     // -c
    bool decompress = false;                // -Z
    bool protein = false;
    bool progress = false;
    IupacOption iupac = IupacOption::Never;
    std::vector<std::string> output;        // hidden --output [a [b]]
    int device = 0;                         // first CUDA device
    int gpus = 1;                           // devices device .. device+gpus-1 share the batches of every input
    size_t slot_mb = 64;                    // ring slot size per input stream (memory bound = 3 slots x gpus x streams)
};

inline std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw FqgrepError(FQG_E_PATTERN, "Could not open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
// src/main.rs:246-255: one pattern per line (BufRead::lines: "\n" or "\r\n" terminated), lines not trimmed
inline std::vector<std::string> read_patterns(const std::string& path) {
    std::vector<std::string> v;
    const std::string data = slurp(path);
    size_t pos = 0;
    while (pos < data.size()) {
        size_t nl = data.find('\n', pos);
        std::string line = data.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        v.push_back(line);
        if (nl == std::string::npos) break;
        pos = nl + 1;
    }
    return v;
}
// src/main.rs:258-271: one name per line, trimmed, empty lines skipped
inline std::set<std::string> read_query_names(const std::string& path) {
    std::set<std::string> names;
    std::stringstream ss(slurp(path));
    for (std::string line; std::getline(ss, line, '\n');) {
        const char* ws = " \t\r\n\v\f";
        const size_t a = line.find_first_not_of(ws);
        if (a == std::string::npos) continue;
        names.insert(line.substr(a, line.find_last_not_of(ws) - a + 1));
    }
    return names;
}

namespace detail {
struct Sink {      // a selected-records destination: file or the caller's stream
    FILE* f = nullptr;
    bool owned = false;
    ~Sink() { if (owned && f) std::fclose(f); }
    void open(const std::string& path) {
        f = std::fopen(path.c_str(), "wb");
        if (!f) throw FqgrepError(FQG_E_PATTERN, "Could not create " + path);
        owned = true;
    }
    void write(const std::vector<uint8_t>& v, size_t n) {
        if (n && std::fwrite(v.data(), 1, n, f) != n) throw FqgrepError(FQG_E_PATTERN, "write failed");
    }
};

// proglog's "Searched N reads" line every 50 M reads (src/main.rs:468-480, 636-638, 737-740); counts arrive per batch here
struct Progress {
    bool on;
    uint64_t seen, reported;
    void record(uint64_t n) {
        if (!on) return;
        seen += n;
        while (seen - reported >= 50000000ull) {
            reported += 50000000ull;
            std::string digits = std::to_string(reported), out;
            for (size_t i = 0; i < digits.size(); i++) { if (i && (digits.size() - i) % 3 == 0) out.push_back(','); out.push_back(digits[i]); }
            std::fprintf(stderr, "[fqgrep-progress] Searched %s reads\n", out.c_str());
        }
    }
    void finish() {
        if (!on) return;
        std::fprintf(stderr, "[fqgrep-progress] Searched %llu reads in total\n", (unsigned long long)seen);
    }
};

// One input file as a stream of bytes (spawn_reader, src/main.rs:61-87), delivered into caller-provided buffers.
class InputStream {
  public:
    InputStream(const std::string& path, bool decompress) { check(fqg_reader_open(path.c_str(), decompress ? 1 : 0, &r_)); }
    ~InputStream() { if (r_) fqg_reader_close(r_); }
    InputStream(const InputStream&) = delete;
    InputStream& operator=(const InputStream&) = delete;
    bool eof() const { return eof_; }
    // appends to buf[have .. cap); returns the new fill level.  Errors are kept for the caller's thread (rc / message).
    size_t fill(uint8_t* buf, size_t have, size_t cap) {
        while (have < cap && !eof_) {
            size_t n = 0;
            int e = 0;
            rc = fqg_reader_read(r_, buf + have, cap - have, &n, &e);
            if (rc != FQG_OK) { message = fqg_last_error(); eof_ = true; break; }
            have += n;
            if (e) eof_ = true;
        }
        return have;
    }
    int rc = FQG_OK;
    std::string message;

  private:
    fqg_reader* r_ = nullptr;
    bool eof_ = false;
};

// The per-input body of fqgrep_from_opts' loops (src/main.rs:550-665) on the streaming path: reader threads fill the next
// free slot of the ring (R1 and R2 side by side), the batch is cut at record ends with the mates in lock step (counted,
// like PairedFastqReader::fill_data reads "the same number of records", src/lib/seq_io.rs:155-163), submitted, and the
// oldest batch in flight is retired: count, progress, ordered write-back from the device's record spans.
inline void run_input(const Matcher& matcher, int mode, const std::string& f1, const std::string* f2, const Opts& opts, FILE* w1, FILE* w2,
                      uint64_t* num_matches, uint64_t* num_records, Progress* progress) {
    InputStream in1(f1, opts.decompress);
    std::unique_ptr<InputStream> in2;
    if (f2) in2.reset(new InputStream(*f2, opts.decompress));
    fqg_stream_opts so{};
    so.n_slots = 3;
    so.want_spans = opts.count ? 0u : 1u;
    so.slot_bytes = opts.slot_mb << 20;
    const size_t cap = so.slot_bytes;
    std::vector<fqg_stream*> streams((size_t)opts.gpus, nullptr);
    struct Closer { std::vector<fqg_stream*>& v; ~Closer() { for (auto* s : v) fqg_stream_close(s); } } closer{streams};
    for (int g = 0; g < opts.gpus; g++) check(fqg_stream_open(context(opts.device + g), matcher.handle(), mode, &so, &streams[(size_t)g]));
    std::vector<uint8_t> carry1, carry2, outbuf, outbuf2;
    uint64_t fed = 0, taken = 0;
    auto retire = [&]() {
        fqg_batch b;
        check(fqg_stream_next(streams[taken % streams.size()], &b));
        taken++;
        *num_matches += b.n_selected;
        *num_records += b.n_records;
        progress->record(b.n_records);
        if (opts.count || b.n_spans == 0) return;
        outbuf.resize(b.n1 + b.n2 + 64);
        size_t l1 = 0, l2 = 0;
        if (w2) {
            outbuf2.resize(b.n2 + 64);
            check(fqg_write_spans(b.spans, b.n_spans, b.r1, b.n1, b.r2, b.n2, outbuf.data(), outbuf.size(), &l1, outbuf2.data(), outbuf2.size(), &l2));
        } else {
            check(fqg_write_spans(b.spans, b.n_spans, b.r1, b.n1, b.r2, b.n2, outbuf.data(), outbuf.size(), &l1, nullptr, 0, &l2));
        }
        if (l1 && std::fwrite(outbuf.data(), 1, l1, w1) != l1) throw FqgrepError(FQG_E_PATTERN, "write failed");
        if (l2 && std::fwrite(outbuf2.data(), 1, l2, w2) != l2) throw FqgrepError(FQG_E_PATTERN, "write failed");
    };
    const uint64_t max_in_flight = (uint64_t)so.n_slots * streams.size();
    for (;;) {
        if (fed - taken >= max_in_flight) retire();
        fqg_stream* s = streams[fed % streams.size()];
        uint8_t *b1 = nullptr, *b2 = nullptr;
        size_t slot_cap = 0;
        check(fqg_stream_acquire(s, &b1, &b2, &slot_cap));
        if (carry1.size() > cap || carry2.size() > cap) throw FqgrepError(FQG_E_TOO_LARGE, "a record (pair) is larger than the ring slot; raise --slot-mb");
        if (!carry1.empty()) std::memcpy(b1, carry1.data(), carry1.size());
        if (in2 && !carry2.empty()) std::memcpy(b2, carry2.data(), carry2.size());
        size_t n1 = carry1.size(), n2 = carry2.size();
        // the reader threads of the reference (one per input): file read / gunzip of R1 and R2 side by side
        if (in2) {
            std::thread t2([&]() { n2 = in2->fill(b2, n2, cap); });
            n1 = in1.fill(b1, n1, cap);
            t2.join();
            if (in2->rc != FQG_OK) throw FqgrepError(in2->rc, in2->message);
        } else {
            n1 = in1.fill(b1, n1, cap);
        }
        if (in1.rc != FQG_OK) throw FqgrepError(in1.rc, in1.message);
        const bool final = in1.eof() && (!in2 || in2->eof());
        size_t c1 = 0, c2 = 0;
        check(fqg_fastq_cut(mode, b1, n1, b2, n2, 0, final ? 1 : 0, mode == FQG_SINGLE ? 0 : 1, &c1, &c2));
        if (c1 == 0 && c2 == 0) {
            if (!final && (n1 >= cap || (in2 && n2 >= cap))) throw FqgrepError(FQG_E_TOO_LARGE, "a record (pair) is larger than the ring slot; raise --slot-mb");
            // nothing cut although the input has ended: whatever is left goes to the device, which names the error (or finds nothing)
            c1 = n1;
            c2 = n2;
        }
        if (c1 || c2) {
            check(fqg_stream_submit(s, c1, c2));
            fed++;
        }
        carry1.assign(b1 + c1, b1 + n1);
        if (in2) carry2.assign(b2 + c2, b2 + n2);
        if (final && carry1.empty() && carry2.empty()) break;
        if (final && c1 == 0 && c2 == 0) break;
    }
    while (taken < fed) retire();
}
}  // namespace detail

// Returns the number of selected records (single-end) or pairs; throws FqgrepError.  Selected records go to `out`
// (or to --output files) in input order, in seq_io's write_to form (src/main.rs:641-657, 682-727).
inline uint64_t fqgrep_from_opts(const Opts& opts, FILE* out = stdout) {
    std::vector<std::string> regexp = opts.regexp;
    std::set<std::string> query_names;
    const bool by_name = opts.has_read_names_file;
    if (by_name) {
        query_names = read_query_names(opts.read_names_file);
        // src/main.rs:381-385
        if (opts.reverse_complement) std::fprintf(stderr, "Warning: --reverse-complement is ignored when using query name matching (-N)\n");
    }
    else if (opts.has_file) for (auto& p : read_patterns(opts.file)) regexp.push_back(p);

    bool has_pattern = false;
    std::string pattern;
    std::vector<std::string> files;
    if (by_name || !regexp.empty()) {
        files = opts.args;
    } else {
        if (opts.args.empty()) throw FqgrepError(FQG_E_PATTERN, "Pattern must be given with -e or as the first positional argument ");
        has_pattern = true;
        pattern = opts.args[0];
        files.assign(opts.args.begin() + 1, opts.args.end());
    }

    // expand_iupac, src/main.rs:315-372
    bool fixed_strings = opts.fixed_strings, has_bitencs = false;
    std::vector<std::string> bitencs;
    if (opts.iupac != IupacOption::Never && fixed_strings) {
        std::vector<std::string> patterns = has_pattern ? std::vector<std::string>{pattern} : regexp;
        if (opts.iupac == IupacOption::Expand) {
            regexp.clear();
            for (auto& p : patterns) for (auto& e : expand_iupac_fixed_pattern(p)) regexp.push_back(e);
            has_pattern = false;
        } else if (opts.iupac == IupacOption::Regex) {
            regexp.clear();
            for (auto& p : patterns) regexp.push_back(expand_iupac_regex(p));
            has_pattern = false;
            fixed_strings = false;
        } else {
            has_bitencs = true;
            bitencs = patterns;
            if (patterns.size() == 1) { has_pattern = true; pattern = patterns[0]; regexp.clear(); }
            else { has_pattern = false; regexp = patterns; }
        }
    }

    if (opts.paired && !(files.size() <= 1 || files.size() % 2 == 0))
        throw FqgrepError(FQG_E_PATTERN, "Input files must be a multiple of two, or either a single file or standard input (assume interleaved) with --paired");
    if (opts.output.size() > 2) throw FqgrepError(FQG_E_PATTERN, "Expected at most 2 output files, got " + std::to_string(opts.output.size()));
    if (opts.output.size() == 2 && !opts.paired) throw FqgrepError(FQG_E_PATTERN, "Two output files may only be given with --paired");

    if (!by_name && fixed_strings && !has_bitencs) {
        if (has_pattern) validate_fixed_pattern(pattern, opts.protein);
        else if (!regexp.empty()) for (auto& p : regexp) validate_fixed_pattern(p, opts.protein);
        else throw FqgrepError(FQG_E_PATTERN, "A pattern must be given as a positional argument or with -e/--regexp");
    }

    MatcherOpts match_opts;
    match_opts.invert_match = opts.invert_match;
    match_opts.reverse_complement = opts.reverse_complement;
    std::unique_ptr<Matcher> matcher = by_name ? MatcherFactory::new_query_name_matcher(query_names, match_opts)
                                               : MatcherFactory::new_matcher(has_pattern ? &pattern : nullptr, fixed_strings, regexp,
                                                                             has_bitencs ? &bitencs : nullptr, match_opts);
    detail::Sink w1, w2;
    if (!opts.count) {
        if (opts.output.size() >= 1) w1.open(opts.output[0]); else w1.f = out;
        if (opts.output.size() == 2) w2.open(opts.output[1]);
    }

    uint64_t num_matches = 0, num_records = 0;
    if (files.empty()) files.push_back("-");
    if (opts.gpus < 1 || opts.gpus > 64) throw FqgrepError(FQG_E_INVALID_ARG, "--gpus must be between 1 and 64");
    detail::Progress progress{opts.progress, 0, 0};
    auto run = [&](const std::string& f1, const std::string* f2, int mode) {
        detail::run_input(*matcher, mode, f1, f2, opts, w1.f, w2.f, &num_matches, &num_records, &progress);
    };
    if (opts.paired) {
        if (files.size() == 1) run(files[0], nullptr, FQG_INTERLEAVED);
        else for (size_t i = 0; i + 1 < files.size(); i += 2) run(files[i], &files[i + 1], FQG_PAIRED);
    } else {
        for (auto& f : files) run(f, nullptr, FQG_SINGLE);
    }
    progress.finish();
    if (opts.count) std::fprintf(out, "%llu\n", (unsigned long long)num_matches);
    std::fflush(out);
    return num_matches;
}

// src/main.rs:295-311: 0 if something was selected, 1 if nothing was, 2 on an error, 101 where the reference panics
inline int fqgrep(const Opts& opts, FILE* out = stdout, FILE* err = stderr) {
    try {
        return fqgrep_from_opts(opts, out) == 0 ? 1 : 0;
    } catch (const FqgrepError& e) {
        if (e.exit_code() == 101) {
            std::fprintf(err, "Error: fqgrep panicked.  Please report this as a bug!\n%s\n", e.what());
            return 101;
        }
        std::fprintf(err, "Error: %s\n", e.what());
        return 2;
    }
}

}  // namespace fqgrep_b200
