/*
 * fqgrep_b200.h -- C ABI of the B200-native read-matching path of fqgrep.
 *
 * The reference (fulcrumgenomics/fqgrep v1.1.2, Rust) has no FFI on this path: the seam is the trait
 * object `Box<dyn Matcher + Sync + Send>` built once by `MatcherFactory` and called per record from the
 * seq_io worker pool.  Each entry point below names the reference interface it replaces
 * (file:line under the reference tree).  INTEGRATION.md shows the Rust `extern "C"` block + `impl Matcher`
 * a maintainer adds on the reference side.
 *
 * Plain C: pointers, sizes, ints.  No torch / CUDA types in signatures (a CUDA stream crosses as void*).
 * All functions return 0 on success or a negative fqg_status; fqg_last_error() gives the thread-local text.
 * There is NO CPU fallback: every matching entry point needs a CUDA device and fails with FQG_E_CUDA
 * when none is usable.  Pattern compilation (fqg_matcher_new) is host-only and needs no device.
 */
#ifndef FQGREP_B200_H
#define FQGREP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fqg_ctx fqg_ctx;         /* one per (process, GPU): streams, staging buffers, device tables */
typedef struct fqg_matcher fqg_matcher; /* immutable compiled pattern set; shareable across threads and ctxs */

typedef enum fqg_status {
    FQG_OK = 0,
    FQG_E_INVALID_ARG = -1,
    FQG_E_PATTERN = -2,        /* construction failed; reference: anyhow error -> exit 2 (src/main.rs:305-307) */
    FQG_E_CUDA = -3,           /* no device / CUDA failure */
    FQG_E_FASTQ = -4,          /* malformed FASTQ; reference: .unwrap() panic -> exit 101 (src/main.rs:296-301,659) */
    FQG_E_PAIR_MISMATCH = -5,  /* R1/R2 record counts differ or read ids differ (src/lib/seq_io.rs:149-198, src/main.rs:741-747) */
    FQG_E_TOO_LARGE = -6
} fqg_status;

/* Which reference matcher the set stands for: the MatcherFactory dispatch, src/lib/matcher.rs:652-676. */
typedef enum fqg_kind {
    FQG_FIXED = 0,      /* FixedStringMatcher      matcher.rs:195-238  (pattern.is_some(), fixed_strings) */
    FQG_FIXED_SET = 1,  /* FixedStringSetMatcher   matcher.rs:240-305  (-F with -e/-f)                    */
    FQG_REGEX = 2,      /* RegexMatcher            matcher.rs:489-532  (positional pattern)               */
    FQG_REGEX_SET = 3,  /* RegexSetMatcher         matcher.rs:534-599  (-e/-f without -F)                 */
    FQG_NAMES = 4,      /* QueryNameMatcher        matcher.rs:601-639  (-N)                               */
    FQG_BITMASK = 5,    /* BitMaskMatcher          matcher.rs:384-432  (--iupac bit-mask, one pattern)    */
    FQG_BITMASK_SET = 6 /* BitMaskSetMatcher       matcher.rs:434-487  (--iupac bit-mask, -e/-f)          */
} fqg_kind;

/* MatcherOpts, src/lib/matcher.rs:20-28 */
#define FQG_INVERT_MATCH 1u        /* MatcherOpts::invert_match        (-v) */
#define FQG_REVERSE_COMPLEMENT 2u  /* MatcherOpts::reverse_complement  (--reverse-complement) */

/* Input pairing, src/main.rs:550-665 */
typedef enum fqg_mode {
    FQG_SINGLE = 0,       /* one record = one unit                                   main.rs:630-665 */
    FQG_INTERLEAVED = 1,  /* records 2k,2k+1 of r1 form a pair (InterleavedFastqReader) main.rs:552-589 */
    FQG_PAIRED = 2        /* record k of r1 pairs with record k of r2 (PairedFastqReader) main.rs:590-629 */
} fqg_mode;

const char* fqg_last_error(void);
const char* fqg_version(void);

/* ---- construction (host only) --------------------------------------------------------------------------- */

/* Replaces MatcherFactory::new_matcher / new_query_name_matcher (matcher.rs:645-685).
 * patterns: `n` byte strings, pattern i = blob[offsets[i] .. offsets[i+1]).  For FQG_NAMES they are read ids.
 * Error text follows the reference: "Invalid regular expression: '<p>'" (matcher.rs:499),
 * "Failed to build regex set from patterns" (:554), "Failed to build Aho-Corasick automaton" (:298). */
int fqg_matcher_new(int kind, const uint8_t* blob, const uint64_t* offsets, uint32_t n, uint32_t opts, fqg_matcher** out);
void fqg_matcher_free(fqg_matcher* m);

/* Replaces validate_fixed_pattern (matcher.rs:128-147).  0 if valid; FQG_E_PATTERN and the reference's
 * message ("Fixed pattern must contain only DNA bases: A .. [X] .. GTG") in fqg_last_error() otherwise. */
int fqg_validate_fixed_pattern(const char* pattern, int protein);

/* --iupac expand / regex (src/lib/mod.rs:70-118, applied at src/main.rs:339-360): rewrites ONE -F pattern.
 * to_regex=0: the expanded fixed strings joined by '\n' (error beyond 10 000, same text as the reference);
 * to_regex=1: the character-class regex ("GATK" -> "GAT[GT]").  Call with out=NULL to size the buffer. */
int fqg_iupac_expand(const char* pattern, int to_regex, char* out, size_t cap, size_t* out_len);

/* Compiled-automaton introspection (used by the CPU tests to check the compilers without a GPU). */
typedef struct fqg_dfa_info {
    uint32_t n_states, n_classes, start, accept_from;
    uint32_t tier;          /* 0: 4-base-stride table in shared memory, 1: class-indexed table in shared memory,
                               2: class-indexed table in L2 (hot rows cached in shared memory), 3: q-gram filter in shared memory +
                               exact verify (tier-2 table as fallback), 4: 2-base-stride table in shared memory,
                               5: mixed set split into a literal part (tier 3) and a regex part (small table), two SoA passes */
    uint32_t table_bytes;   /* device table footprint */
} fqg_dfa_info;
int fqg_matcher_dfa_info(const fqg_matcher* m, fqg_dfa_info* out);
/* byte_class: 256 entries; next: n_states*n_classes entries.  Either may be NULL. */
int fqg_matcher_dfa_export(const fqg_matcher* m, uint8_t* byte_class, uint32_t* next);

/* ---- device context --------------------------------------------------------------------------------------- */

int fqg_ctx_create(int device, fqg_ctx** out);
void fqg_ctx_destroy(fqg_ctx* ctx);
int fqg_device_count(void);

/* Pinned host memory for the streaming path (reader threads fill these; src/main.rs:61-87 is the reference's reader). */
int fqg_alloc_pinned(size_t bytes, void** out);
void fqg_free_pinned(void* p);

/* ---- matching: HBM-resident SoA reads ----------------------------------------------------------------------- */

/* Replaces the worker-pool loop `for record in batch { found = matcher.read_match(&record) }`
 * (src/lib/seq_io.rs:314-326 calling src/main.rs:635-640 -> matcher.rs:164-175) for a batch of reads whose bases
 * already sit in device memory.
 *   d_bases            device pointer, bases of all reads; must stay readable 16 bytes past the last read
 *   d_start, d_end     device uint32[n_reads]: read i = d_bases[d_start[i] .. d_end[i]), ascending.
 *                      For back-to-back reads pass one offsets array o[n+1] as (o, o+1).
 *   d_mask_out         device uint32[ceil(n_reads/32)]: bit i%32 of word i/32 = read_match(read i)
 *   d_or_mask          optional device mask OR-ed into the result before it is written and counted
 *                      (pair rule `read_match(r1) || read_match(r2)`, src/main.rs:749-755)
 *   d_count            optional device uint64, incremented by the number of set bits written
 *   avg_read_len       hint used to size shared-memory staging (0 = 150)
 *   stream             cudaStream_t as void* (NULL = the legacy default stream, like any CUDA API).  Asynchronous.
 * With a FQG_NAMES matcher (QueryNameMatcher::read_match, matcher.rs:631-638) the spans are the records' HEADER lines
 * without the leading '@' instead of their bases; the id ends at the first space, as seq_io's id_bytes() has it. */
int fqg_match_bases_device(fqg_ctx* ctx, const fqg_matcher* m, const uint8_t* d_bases, const uint32_t* d_start, const uint32_t* d_end,
                           uint64_t n_reads, uint32_t* d_mask_out, const uint32_t* d_or_mask, uint64_t* d_count, uint32_t avg_read_len,
                           void* stream);

/* Same from host memory: copies bases + offsets in, runs the kernel, copies the mask back.  Synchronous.
 * offsets: uint64[n_reads+1] into bases.  mask_out: uint32[ceil(n/32)] (may be NULL).  n_selected may be NULL. */
int fqg_match_bases_host(fqg_ctx* ctx, const fqg_matcher* m, const uint8_t* bases, const uint64_t* offsets, uint64_t n_reads,
                         uint32_t* mask_out, uint64_t* n_selected);

/* ---- matching: raw FASTQ ---------------------------------------------------------------------------------- */

typedef struct fqg_result {
    uint64_t n_records;   /* records parsed from r1 (and from r2 in FQG_PAIRED) */
    uint64_t n_units;     /* records (single) or pairs */
    uint64_t n_selected;  /* what --count prints, src/main.rs:667-671 */
    double h2d_ms, kernel_ms, d2h_ms;   /* device-event timings of the last call */
    uint64_t h2d_bytes, d2h_bytes;
} fqg_result;

/* Replaces the whole per-input body of fqgrep_from_opts' main loop (src/main.rs:550-665) up to the selected flags:
 * parse records (seq_io rules), decide read_match per record, combine mates.  FASTQ bytes are in HOST memory
 * (pinned memory from fqg_alloc_pinned streams fastest).  unit_mask_out: optional uint32[ceil(n_units/32)]
 * sized by the caller from an upper bound (bytes/8 records is always enough); bit k = unit k selected. */
int fqg_grep_fastq_host(fqg_ctx* ctx, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2,
                        uint32_t* unit_mask_out, size_t unit_mask_words, fqg_result* res);

/* Same with the FASTQ bytes already resident in device memory (d_r1/d_r2 readable 16 bytes past the end). */
int fqg_grep_fastq_device(fqg_ctx* ctx, const fqg_matcher* m, int mode, const uint8_t* d_r1, size_t n1, const uint8_t* d_r2, size_t n2,
                          uint32_t* d_unit_mask_out, size_t unit_mask_words, fqg_result* res, void* stream);

/* K0, for callers that parse once and match many times (the batcher of north_star: "finds record and line boundaries ...
 * and compacts the sequence lines into an SoA bases buffer with per-read offsets").  Replaces the reader thread's record
 * splitting (seq_io 0.3.4 RecordSet, driven from src/lib/seq_io.rs:60-80, 149-198) without matching anything.
 *   fqg_index_fastq_device   d_seq_start[r] = stream offset of record r's sequence line, d_seq_len[r] = its length (CR
 *                            stripped); same validation and errors as fqg_grep_fastq_*.  Arrays hold `capacity` records.
 *   fqg_compact_bases_device gathers those lines back to back into d_bases_out (sum of lengths + 16 bytes) and writes
 *                            d_offsets_out[n_records+1] (u32: at most 4 GiB of bases per call), ready for
 *                            fqg_match_bases_device(d_bases_out, d_offsets_out, d_offsets_out + 1, ...).                */
int fqg_index_fastq_device(fqg_ctx* ctx, const uint8_t* d_fastq, size_t n, uint64_t* d_seq_start, uint32_t* d_seq_len, uint64_t capacity,
                           uint64_t* n_records, void* stream);
int fqg_compact_bases_device(fqg_ctx* ctx, const uint8_t* d_fastq, const uint64_t* d_seq_start, const uint32_t* d_seq_len, uint64_t n_records,
                             uint8_t* d_bases_out, uint32_t* d_offsets_out, void* stream);

/* Selected-record write-back (src/main.rs:641-657 and write_paired_record :682-727 via seq_io write_to):
 * emits "@head\nseq\n+\nqual\n" for every selected unit in input order (pairs: R1 then R2), normalising CRLF
 * and "+name" lines like the reference.  out/out2 must be at least n1(+n2) bytes; out2 non-NULL splits R2
 * into its own stream (the hidden --output a b).  Host-side, driven by the device mask. */
int fqg_write_selected(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* unit_mask,
                       uint8_t* out, size_t* out_len, uint8_t* out2, size_t* out2_len);

/* Input plumbing, host side: replaces spawn_reader (src/main.rs:61-87).  Reads `path` ("-" = stdin) completely into one
 * buffer (pinned=1: cudaHostAlloc, the fast path for fqg_grep_fastq_host; pinned=0: malloc, no CUDA needed).  The input is
 * gunzipped (multi-member gzip / BGZF, like flate2's MultiGzDecoder; pure BGZF files are inflated member-parallel on all
 * host threads) when the extension is .gz/.bgz, or when `decompress`
 * (-Z) is set and the extension is not .fq/.fastq (src/main.rs:77, src/lib/mod.rs:157-183).  Free with fqg_free_input. */
int fqg_read_input(const char* path, int decompress, int pinned, uint8_t** out, size_t* n_out);
void fqg_free_input(uint8_t* p, int pinned);
int fqg_is_gzip_path(const char* path);    /* src/lib/mod.rs:170-175 */
int fqg_is_fastq_path(const char* path);   /* src/lib/mod.rs:177-183 */

/* Multi-GPU sharding helper (host): cut a FASTQ buffer into `parts` contiguous byte ranges that begin at record
 * starts (the reference has no sharding; this replaces nothing, it feeds one fqg_ctx per GPU).  cuts: uint64[parts+1],
 * cuts[0]=0, cuts[parts]=n.  Ranges hold roughly equal BYTES; record counts per range come back in fqg_result. */
int fqg_fastq_split(const uint8_t* buf, size_t n, uint32_t parts, uint64_t* cuts);

/* Benchmark utility: the synthetic-read generator of the benchmark spec as a device kernel (inputs are produced
 * in HBM, never crossing PCIe).  fastq=0: n_reads*read_len bases back to back; fastq=1: 317-byte style records
 * "@r%010d\n<bases>\n+\n<I...>\n".  plant_*: patterns planted (50% reverse-complemented) into a `plant_rate`
 * fraction of the reads; pattern i = plant_blob[plant_offsets[i]..plant_offsets[i+1]) (host pointers).
 * Bit-identical to tests/synth.py. */
int fqg_synth_reads_device(fqg_ctx* ctx, uint8_t* d_out, uint64_t n_reads, uint64_t first_index, uint32_t read_len, int fastq,
                           uint64_t seed, const uint8_t* plant_blob, const uint32_t* plant_offsets, uint32_t n_plant, double plant_rate,
                           uint64_t plant_seed, void* stream);

/* Counters for evidence: number of kernels this library launched since load. */
uint64_t fqg_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* FQGREP_B200_H */
