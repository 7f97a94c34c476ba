// fastq_host.hpp -- host-side FASTQ glue: record walking for the ordered write-back of selected records.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/fqgrep_b200.h"

namespace fqg {

// One FASTQ record as seq_io 0.3.4 exposes it (head without '@', CR stripped).
struct HostRecord {
    const uint8_t *head, *seq, *qual;
    size_t head_len, seq_len, qual_len;
};
// Returns 1 (record), 0 (clean end of input), <0 = FQG_E_FASTQ with *err set.
int next_host_record(const uint8_t* buf, size_t n, size_t* pos, HostRecord* rec, std::string* err);

// input plumbing (input_host.cc)
bool is_gzip_path(const char* path);
bool is_fastq_path(const char* path);
int read_input(const char* path, int decompress, int pinned, uint8_t** out, size_t* n_out, std::string* err);
void free_input(uint8_t* p, int pinned);

// incremental input (reader_host.cc); see fqg_reader_* in include/fqgrep_b200.h
struct InputReader;
int reader_open(const char* path, int decompress, InputReader** out, std::string* err);
int reader_read(InputReader* r, uint8_t* dst, size_t cap, size_t* n, bool* eof, std::string* err);
void reader_close(InputReader* r);

size_t find_record_start(const uint8_t* buf, size_t n, size_t pos);

// batch cutter of the streaming path (fastq_cut.cc); see fqg_fastq_cut in include/fqgrep_b200.h
int fastq_cut(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, size_t limit, bool final, bool exact, size_t* cut1, size_t* cut2);
uint64_t count_newlines(const uint8_t* p, size_t n);
size_t offset_after_newline(const uint8_t* p, size_t n, uint64_t k);

// write-back from device spans (fastq_host.cc); see fqg_write_spans
int write_spans_host(const fqg_span* spans, uint64_t n_spans, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, uint8_t* out, size_t out_cap,
                     size_t* out_len, uint8_t* out2, size_t out2_cap, size_t* out2_len, std::string* err);

// Ordered write-back of selected units (src/main.rs:641-657, 682-727).
int write_selected_host(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* unit_mask, uint8_t* out,
                        size_t* out_len, uint8_t* out2, size_t* out2_len, std::string* err);

}  // namespace fqg
