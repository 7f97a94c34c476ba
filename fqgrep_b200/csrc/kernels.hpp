// kernels.hpp -- host-callable launchers of the sm_100a kernels (definitions in *.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fqgrep_b200.h"

namespace fqg {

// Device-resident automaton in the layout the kernels index directly.
//   tier 0: uint16 table[n_states][256 + pad] in shared memory.  Entries 0..255 of a row are FOUR-base transitions
//           indexed by the packed 2-bit codes of four consecutive bases; entries 256.. are one-byte transitions by
//           byte class (used for words that are not pure ACGT and for the <4-byte tail).  An entry is the BYTE
//           offset of the next state's row; the 528 B row stride skews consecutive states by 4 banks.
//   tier 1: uint16 table[n_states][n_classes] in smem, entries = byte offset of the next row (state * n_classes * 2);
//           the 256 B byte->class LUT holds class*2 (tiers 0 and 1).
//   tier 2: uint32 table[n_states][n_classes] read through L2, entries = entry index of the next row
//           (state * n_classes); the LUT holds the plain class.
// Device view of the q-gram filter (tier 3); see gram_filter.cc / match_soa.cu.
struct DeviceGram {
    const uint32_t* bitmap = nullptr;    // device copy, staged into shared memory by the kernel
    uint32_t bitmap_bytes = 0, bitmap_shift = 0;   // h = gram * 0x9E3779B1: word index = h >> bitmap_shift, bit positions (h>>2, h>>7, h>>12) & 31
    const uint4* slots = nullptr;          // {gram key, pool offset, pattern length << 5 | offset in the pattern, -}; .z all ones = empty
    uint32_t slot_shift = 0, slot_mask = 0;
    const uint8_t* pool = nullptr;         // patterns, each on a 16-byte boundary
    uint32_t q = 0, s = 0, min_len = 0, qmask = 0;
};

struct DeviceDfa {
    int tier = 0;                    // 3 = q-gram filter for the SoA kernel, with the tier-2 table below as its exact fallback
    DeviceGram gram;
    const void* table = nullptr;     // device
    uint32_t table_bytes = 0;
    const uint8_t* byte_class = nullptr;   // device, 256 B
    uint32_t row = 0;                // what a state id is multiplied by (see tiers above)
    uint32_t start = 0;              // premultiplied
    uint32_t accept_from = 0;        // premultiplied
    uint32_t n_states = 0;
    uint32_t hot_entries = 0;        // tiers 2/3: leading table entries (whole rows of the shallowest states) cached in smem
    // States every transition of which leads back to themselves (premultiplied like `start`, 0xFFFFFFFF = none): the dead
    // state of a set of ^-anchored patterns and the all-accepting state of patterns without `$`.  `settle` is set when a
    // dead state exists, i.e. when every read is decided after a few bases; the slower tiers then stop walking there.
    uint32_t settled[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    uint32_t settle = 0;
};

struct SoaBatch {
    const uint8_t* bases;
    const uint32_t* start;
    const uint32_t* end;
    uint32_t n_reads;
    uint32_t* mask_out;
    const uint32_t* or_mask;
    unsigned long long* count;
    uint32_t invert;
    uint32_t avg_read_len;
    const uint32_t* pre_or_mask;   // optional: "pattern found" bits of another matcher, OR-ed in BEFORE invert_match is applied
    // optional: the read count is only known on the device (the FASTQ indexer's line count, 4 lines per read); n_reads is
    // then the capacity the launch is sized for and the kernel stops at min(n_reads, *lines_dev / 4)
    const unsigned long long* lines_dev;
};

// K3/K2/K1: per-read DFA walk over SoA reads (bases + start/end), bitmask + count epilogue (K5 fused).
cudaError_t launch_match_soa(const DeviceDfa& dfa, const SoaBatch& b, int sm_count, cudaStream_t stream);

// ---- fused raw-FASTQ parse + match (match_fastq.cu)
struct NamesView {
    const uint64_t* slots = nullptr;
    const uint8_t* pool = nullptr;
    uint32_t mask = 0;     // n_buckets - 1 (a bucket = four consecutive slots = 32 bytes)
};
struct FastqArgs {
    const uint8_t* buf;            // device, readable 16 bytes past nbytes
    uint64_t nbytes;               // stream bytes, trailing newlines stripped
    uint32_t virtual_final_nl;     // 1: treat offset nbytes as a newline
    uint32_t tile_first, tile_count;   // this launch handles global tiles [tile_first, tile_first + tile_count)
    uint32_t* tile_counter;        // zeroed, one per launch
    unsigned long long* tile_state;    // zeroed, one slot per global tile of the stream (used as u32: 0x80000000 | the tile's newline count)
    unsigned long long* seg_acc;       // zeroed, one per segment of 32 consecutive tiles: tiles reported << 48 | their newlines
    unsigned long long* seg_pre;       // zeroed, one per segment: 2 << 62 | newlines up to its end, written by the scan warp (match_fastq.cu, Lookback)
    unsigned long long* lines_out;     // total newline count, written by the stream's last tile
    uint32_t* mask;                // zeroed record bitmask, bit = global record number
    uint64_t mask_bits;
    unsigned long long* id_hash;   // optional, one 64-bit id hash per record (pair id check, src/main.rs:741-747)
    uint64_t id_hash_cap;
    unsigned long long* count;     // optional
    unsigned long long* err;       // min over (position << 4 | code), initialised to ~0
    uint32_t invert;
    NamesView names;
    unsigned long long* seq_start;  // optional record index (K0): stream offset of every record's sequence line ...
    uint32_t* seq_len;              // ... and its length without line terminator
    uint64_t span_cap;
    // optional, for batches routed through the SoA kernels: the sequence line of record R is buf[soa_start[R], soa_end[R]),
    // so the SoA kernel reads the lines where they lie (no compaction pass); capacity span_cap like seq_start
    uint32_t* soa_start;
    uint32_t* soa_end;
    // optional record spans (K5, write-back without re-parsing): stream offset of every record's '@' with bit 63 set when
    // the record is not already in seq_io's write_to form (CR line ends, "+name" separator, missing final newline), and
    // its length in bytes including the final newline when there is one
    unsigned long long* rec_off;
    uint32_t* rec_len;
    uint64_t rec_cap;
    // 0x0A0A0A0A and 0x7F7F7F7F, filled in by launch_match_fastq.  They travel as kernel PARAMETERS because the newline test
    // needs two constants per LOP3 and the instruction takes one immediate: as literals (even behind an asm mov, which
    // ptxas folds) every byte-test costs five instructions per word; as constant-bank operands it costs three.
    uint32_t nl_xor, nl_low7;
    uint32_t full_tiles;            // filled in by the launcher: tiles [0, full_tiles) are whole (tile + overlap inside the real bytes)
    uint32_t scan_warp;             // filled in by the launcher: 1 when the last warp of CTA 0 is the scan warp of this launch (match_fastq.cu, Lookback)
    unsigned long long* stats;      // optional (FQG_FASTQ_STATS): [0] tiles, [1] line phase guessed, [2] guessed wrong, [3] dense tiles
};
// -N over SoA headers: spans are header lines without '@'; the id ends at the first space (matcher.rs:631-638)
cudaError_t launch_match_names_soa(const NamesView& nv, const SoaBatch& b, int sm_count, cudaStream_t stream);
uint32_t fastq_tile_bytes();
uint32_t fastq_overlap_bytes();
// names: read-name lookup instead of an automaton; index_only: no matching at all, only parse / validate / index (K0)
cudaError_t launch_match_fastq(const DeviceDfa& dfa, bool names, bool index_only, const FastqArgs& a, int sm_count, cudaStream_t stream);
// K0 part 2: gather the indexed sequence lines into one contiguous SoA buffer with u32 offsets[n+1]
// lines_dev (optional): device line count of the stream; the record count is then min(n_records, *lines_dev / 4)
cudaError_t launch_compact_bases(const uint8_t* fastq, const unsigned long long* seq_start, const uint32_t* seq_len, uint64_t n_records,
                                 uint8_t* bases_out, uint32_t* offsets_out, uint32_t* block_sums, cudaStream_t stream,
                                 const unsigned long long* lines_dev = nullptr);
uint64_t compact_scratch_words(uint64_t n_records);
cudaError_t launch_combine_pairs(const uint32_t* m1, const uint32_t* m2, uint32_t* out, uint64_t cap_words, const unsigned long long* h1, const unsigned long long* h2,
                                 uint64_t idh_cap, const unsigned long long* lines1, const unsigned long long* lines2, unsigned long long* count,
                                 unsigned long long* id_mismatch, cudaStream_t stream);
cudaError_t launch_combine_interleaved(const uint32_t* m, uint32_t* out, uint64_t cap_out_words, const unsigned long long* h, uint64_t idh_cap,
                                       const unsigned long long* lines, unsigned long long* count, unsigned long long* id_mismatch, cudaStream_t stream);

// K5: order-preserving compaction of the selected records' spans (fqg_span, include/fqgrep_b200.h) by the unit mask.
// `lines` is the device line count of stream 1 (units = lines / 4, / 8 when interleaved); n_spans_out receives the count.
uint64_t span_scan_words(uint64_t mask_words);
cudaError_t launch_select_spans(const uint32_t* unit_mask, uint64_t mask_words, int mode, const unsigned long long* lines, const unsigned long long* off0,
                                const uint32_t* len0, const unsigned long long* off1, const uint32_t* len1, fqg_span* out, unsigned long long* n_spans_out,
                                uint32_t* block_sums, cudaStream_t stream);

// Synthetic benchmark input generator; see synth.cu.
struct SynthParams {
    uint8_t* out;
    uint64_t n_reads, first;
    uint32_t read_len, fastq;
    uint64_t seed, plant_seed, plant_threshold;
    const uint8_t* plant_blob;
    const uint32_t* plant_offs;
    uint32_t n_plant;
};
cudaError_t launch_synth(const SynthParams& p, cudaStream_t stream);

}  // namespace fqg
