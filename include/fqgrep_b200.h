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

/* One selected record in write-back order (K5): the device leaves these behind so that the ordered write-back
 * (src/main.rs:641-657, write_paired_record :682-727) copies byte ranges instead of parsing the input a second time. */
typedef struct fqg_span {
    uint64_t offset;   /* byte offset of the record's '@' in its input (r1, or r2 when FQG_SPAN_R2 is set) */
    uint32_t len;      /* bytes the record occupies in the input, final newline included when there is one */
    uint32_t flags;
} fqg_span;
#define FQG_SPAN_R2 1u         /* the record belongs to r2 (FQG_PAIRED) */
#define FQG_SPAN_NORMALISE 2u  /* not already in seq_io's write_to form (CR line ends, "+name" separator line, missing final
                                  newline): emit "@head\nseq\n+\nqual\n" from its fields instead of copying the bytes */

/* Replaces the whole per-input body of fqgrep_from_opts' main loop (src/main.rs:550-665) up to the selected flags:
 * parse records (seq_io rules), decide read_match per record, combine mates.  FASTQ bytes are in HOST memory
 * (pinned memory from fqg_alloc_pinned streams fastest) and may be far larger than device memory: the input is cut into
 * batches of whole records (R1/R2 in lock step) that flow through a ring of device slots, the copy of one batch
 * overlapping the kernels of the one before (fqg_stream_* below is the same machinery, driven by the caller).
 * unit_mask_out: optional uint32[ceil(n_units/32)] sized by the caller from an upper bound (bytes/8 records is always
 * enough); bit k = unit k selected. */
int fqg_grep_fastq_host(fqg_ctx* ctx, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2,
                        uint32_t* unit_mask_out, size_t unit_mask_words, fqg_result* res);

/* Same, and the selected records themselves (src/main.rs:641-657, 682-727): written in input order, in seq_io's write_to
 * form, R1 then R2 for pairs, into out[0..out_cap) -- or R1 into out and R2 into out2 (the hidden `--output a b`) when
 * out2 is not NULL.  *out_len / *out2_len receive the byte counts; FQG_E_TOO_LARGE if a buffer is too small
 * (n1 + n2 bytes always suffice).  The device returns the selected records' spans, the host only copies them. */
int fqg_grep_fastq_host_write(fqg_ctx* ctx, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2,
                              uint8_t* out, size_t out_cap, size_t* out_len, uint8_t* out2, size_t out2_cap, size_t* out2_len, fqg_result* res);

/* Several GPUs, one process: the batches of ONE input are dealt round-robin to the contexts (one per device, any count)
 * and their results taken back in batch order, so masks, counts and output are identical to the single-GPU call.
 * No device-to-device traffic: the only "collective" is this host-side gather. */
int fqg_grep_fastq_host_multi(fqg_ctx* const* ctxs, uint32_t n_ctx, const fqg_matcher* m, int mode, const uint8_t* r1, size_t n1, const uint8_t* r2,
                              size_t n2, uint32_t* unit_mask_out, size_t unit_mask_words, uint8_t* out, size_t out_cap, size_t* out_len, uint8_t* out2,
                              size_t out2_cap, size_t* out2_len, fqg_result* res);

/* Same with the FASTQ bytes already resident in device memory (d_r1/d_r2 16-byte aligned, readable 16 bytes past the end). */
int fqg_grep_fastq_device(fqg_ctx* ctx, const fqg_matcher* m, int mode, const uint8_t* d_r1, size_t n1, const uint8_t* d_r2, size_t n2,
                          uint32_t* d_unit_mask_out, size_t unit_mask_words, fqg_result* res, void* stream);

/* ---- matching: streaming batcher ---------------------------------------------------------------------------------
 * Replaces the reader thread -> worker pool -> ordered main thread pipeline of the reference (spawn_reader
 * src/main.rs:61-87; PairedFastqReader::fill_data lock-step batches src/lib/seq_io.rs:149-198; OrderedReader and the
 * reorder buffer :253-282, 327-364) with a ring of `n_slots` batches in flight on one GPU.  Memory is bounded by the
 * ring: per slot and input stream one pinned host buffer and one device buffer of `slot_bytes`, whatever the input size.
 *
 *   fqg_stream_acquire  the pinned buffers of the next free slot, for reader threads to fill (file read, gunzip)
 *   fqg_stream_submit   the slot now holds n1 (and n2) bytes of WHOLE records -- R1/R2 with equal record counts,
 *                       interleaved input with an even count (fqg_fastq_cut finds such prefixes); enqueues H2D copy,
 *                       kernels and the copies back, returns at once
 *   fqg_stream_feed     same for a batch that lives in the caller's own (ideally pinned) memory: no slot buffer is used
 *   fqg_stream_next     blocks for the OLDEST batch in flight and hands out its result; batches complete in submission
 *                       order, which is input order.  Pointers in fqg_batch stay valid until the slot is reused, i.e.
 *                       until n_slots further batches have been submitted.
 * A batch whose R1/R2 record counts differ, or that ends inside a record, fails in fqg_stream_next with the reference's
 * error (FQG_E_PAIR_MISMATCH / FQG_E_FASTQ); the stream stays usable, so a caller that cut optimistically can re-cut
 * exactly (fqg_fastq_cut with exact=1) and resubmit. */
typedef struct fqg_stream fqg_stream;
typedef struct fqg_stream_opts {
    uint32_t n_slots;     /* batches in flight, 0 = 3 */
    uint32_t want_spans;  /* also return the selected records' spans (for the write-back) */
    size_t slot_bytes;    /* capacity of a slot per input stream, 0 = 64 MiB */
} fqg_stream_opts;
typedef struct fqg_batch {
    uint64_t seq;                 /* batch number, 0-based, in submission order */
    uint64_t n_records, n_units, n_selected;
    const uint32_t* unit_mask;    /* bit k = unit k of this batch selected; ceil(n_units/32) words */
    const fqg_span* spans;        /* want_spans: selected records in output order, offsets relative to r1 / r2 below */
    uint64_t n_spans;
    const uint8_t *r1, *r2;       /* the batch's bytes as submitted (slot buffers, or the caller's memory) */
    size_t n1, n2;
    double h2d_ms, kernel_ms;
} fqg_batch;
int fqg_stream_open(fqg_ctx* ctx, const fqg_matcher* m, int mode, const fqg_stream_opts* opts, fqg_stream** out);
int fqg_stream_acquire(fqg_stream* s, uint8_t** r1_buf, uint8_t** r2_buf, size_t* capacity);
int fqg_stream_submit(fqg_stream* s, size_t n1, size_t n2);
int fqg_stream_feed(fqg_stream* s, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2);
int fqg_stream_in_flight(const fqg_stream* s);
int fqg_stream_next(fqg_stream* s, fqg_batch* out);
void fqg_stream_close(fqg_stream* s);

/* The batch cutter (PairedFastqReader::fill_data's "same number of records from R2", src/lib/seq_io.rs:155-163, and
 * InterleavedFastqReader's even record count, :60-80, without a parsing reader thread): the largest prefixes
 * r1[0..*cut1) / r2[0..*cut2) of at most `limit` bytes each that consist of whole records -- in lock step for
 * FQG_PAIRED, an even number of records for FQG_INTERLEAVED.  final=1 says the inputs end here (the prefixes are then
 * the whole buffers).  exact=0 aligns the mates by their read ids next to the cut (constant work; the device verifies
 * every batch, so a wrong guess cannot go unnoticed); exact=1 counts records on all host threads.  *cut1 == 0 means
 * `limit` does not hold one whole unit. */
int fqg_fastq_cut(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, size_t limit, int final, int exact, size_t* cut1, size_t* cut2);

/* Copies the selected records of one batch to `out` (and R2 records to `out2` when it is not NULL): plain byte copies for
 * spans already in write_to form, re-formatting for FQG_SPAN_NORMALISE ones.  Uses all host threads for large batches. */
int fqg_write_spans(const fqg_span* spans, uint64_t n_spans, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, uint8_t* out, size_t out_cap,
                    size_t* out_len, uint8_t* out2, size_t out2_cap, size_t* out2_len);

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
/* The same plumbing for inputs that must not be held in memory whole (the reader threads of the streaming path): each
 * call delivers up to `cap` further bytes of the -- gunzipped when the rules above say so -- input into dst, e.g. straight
 * into a slot buffer from fqg_stream_acquire.  *eof is set once the input is exhausted.  BGZF members are inflated on all
 * host threads; an input that ends inside a gzip member is an error (flate2's UnexpectedEof -> the reference's exit 101). */
typedef struct fqg_reader fqg_reader;
int fqg_reader_open(const char* path, int decompress, fqg_reader** out);
int fqg_reader_read(fqg_reader* r, uint8_t* dst, size_t cap, size_t* n_read, int* eof);
void fqg_reader_close(fqg_reader* r);
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
