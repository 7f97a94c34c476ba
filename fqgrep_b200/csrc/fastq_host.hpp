// fastq_host.hpp -- host-side FASTQ glue: record walking for the ordered write-back of selected records.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace fqg {

// Device-side scratch owned by a context for the raw-FASTQ path (filled in by fastq path code).
struct FastqDeviceState {
    void* scratch[10] = {nullptr};
    size_t cap[10] = {0};
};
void fastq_state_free(FastqDeviceState* s);

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

size_t find_record_start(const uint8_t* buf, size_t n, size_t pos);

// Ordered write-back of selected units (src/main.rs:641-657, 682-727).
int write_selected_host(int mode, const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const uint32_t* unit_mask, uint8_t* out,
                        size_t* out_len, uint8_t* out2, size_t* out2_len, std::string* err);

}  // namespace fqg
