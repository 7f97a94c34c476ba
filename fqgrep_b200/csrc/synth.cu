// synth.cu -- the synthetic-read generator of SURVEY.md 8(d) as a device kernel, so benchmark inputs are produced
// in HBM without crossing PCIe.  Bit-identical to tests/synth.py (numpy), which the parity tests use.
//   base j of read i  = "ACGT"[ splitmix64(seed ^ GOLDEN*(i*L+j)) >> 62 ]
//   planted reads     : h = splitmix64(plant_seed ^ GOLDEN*i) < rate*2^64 ; pattern (h>>8)%P, reverse-complemented
//                       when (h>>7)&1, written at offset (h>>20) % (L-len+1)
//   FASTQ record      : "@r%010d\n" bases "\n+\n" 'I'*L "\n"   (13 + L + 3 + L + 1 bytes)
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.hpp"

namespace fqg {
namespace {

constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ull;

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += kGolden;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ uint8_t complement(uint8_t c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'Y': return 'R'; case 'R': return 'Y'; case 'K': return 'M'; case 'M': return 'K';
        case 'D': return 'H'; case 'H': return 'D'; case 'V': return 'B'; case 'B': return 'V';
        default: return c;
    }
}

__global__ void synth_kernel(SynthParams p) {
    const uint64_t rec_bytes = p.fastq ? (uint64_t)(2 * p.read_len + 17) : (uint64_t)p.read_len;
    const uint64_t total = p.n_reads * rec_bytes;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = t / rec_bytes;
        const uint32_t k = (uint32_t)(t - r * rec_bytes);
        const uint64_t id = p.first + r;
        uint8_t out;
        int j = -1;   // base index when this byte is a base
        if (!p.fastq) j = (int)k;
        else if (k < 13) {
            if (k == 0) out = '@';
            else if (k == 1) out = 'r';
            else if (k == 12) out = '\n';
            else {
                uint64_t v = id;
                for (uint32_t q = 11; q > k; q--) v /= 10;
                out = (uint8_t)('0' + v % 10);
            }
        } else if (k < 13 + p.read_len) j = (int)(k - 13);
        else if (k == 13 + p.read_len) out = '\n';
        else if (k == 14 + p.read_len) out = '+';
        else if (k == 15 + p.read_len) out = '\n';
        else if (k < 16 + 2 * p.read_len) out = 'I';
        else out = '\n';
        if (j >= 0) {
            out = "ACGT"[splitmix64(p.seed ^ (kGolden * (id * p.read_len + (uint64_t)j))) >> 62];
            if (p.n_plant) {
                const uint64_t h = splitmix64(p.plant_seed ^ (kGolden * id));
                if (h < p.plant_threshold) {
                    const uint32_t pi = (uint32_t)((h >> 8) % p.n_plant);
                    const uint32_t o0 = p.plant_offs[pi], len = p.plant_offs[pi + 1] - o0;
                    if (len <= p.read_len) {
                        const uint32_t off = (uint32_t)((h >> 20) % (p.read_len - len + 1));
                        if ((uint32_t)j >= off && (uint32_t)j < off + len) {
                            const uint32_t q = (uint32_t)j - off;
                            out = ((h >> 7) & 1) ? complement(p.plant_blob[o0 + len - 1 - q]) : p.plant_blob[o0 + q];
                        }
                    }
                }
            }
        }
        p.out[t] = out;
    }
}

}  // namespace

cudaError_t launch_synth(const SynthParams& p, cudaStream_t stream) {
    if (p.n_reads == 0) return cudaSuccess;
    synth_kernel<<<148 * 16, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace fqg
