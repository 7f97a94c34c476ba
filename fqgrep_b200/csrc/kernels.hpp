// kernels.hpp -- host-callable launchers of the sm_100a kernels (definitions in *.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fqg {

// Device-resident automaton in the layout the kernels index directly.
//   tier 0: uint16 table[n_states][264], indexed by raw byte, staged in shared memory; an entry is the BYTE offset
//           of the next state's row (state * 528).  Row stride 528 B skews consecutive states by 4 banks, so the
//           hot (state, base) pairs of a warp land in different banks.
//   tier 1: uint16 table[n_states][n_classes] in smem, entries = byte offset of the next row (state * n_classes * 2);
//           the 256 B byte->class LUT holds class*2.
//   tier 2: uint32 table[n_states][n_classes] read through L2, entries = entry index of the next row
//           (state * n_classes); the LUT holds the plain class.
struct DeviceDfa {
    int tier = 0;
    const void* table = nullptr;     // device
    uint32_t table_bytes = 0;
    const uint8_t* byte_class = nullptr;   // device, 256 B
    uint32_t row = 0;                // what a state id is multiplied by (see tiers above)
    uint32_t start = 0;              // premultiplied
    uint32_t accept_from = 0;        // premultiplied
    uint32_t n_states = 0;
};
constexpr uint32_t kTier0Row = 264;

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
};

// K3/K2/K1: per-read DFA walk over SoA reads (bases + start/end), bitmask + count epilogue (K5 fused).
cudaError_t launch_match_soa(const DeviceDfa& dfa, const SoaBatch& b, int sm_count, cudaStream_t stream);

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
