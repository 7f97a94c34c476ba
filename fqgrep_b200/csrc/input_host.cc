// input_host.cc -- host input plumbing: file / stdin -> (optional multi-member gzip inflate) -> one host buffer that
// the FASTQ entry points stream to the GPU.  Replaces spawn_reader (/root/reference/src/main.rs:61-87: File or stdin ->
// BufReader -> optional flate2 MultiGzDecoder -> fastq::Reader) and the extension rules of src/lib/mod.rs:157-183.
// gzip decoding stays on the host (north_star: excluded from the headline).  Plain (multi-member) gzip is one serial
// zlib stream.  BGZF -- what bgzip / htslib write, the usual form of .fastq.gz from sequencing pipelines -- is a chain of
// independent <= 64 KiB members whose compressed size sits in a 'BC' extra field, so the members are located without
// inflating anything and then inflated by all host threads at once into their final positions.  No matching happens here.
#include <zlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <cuda_runtime.h>

#include "../../include/fqgrep_b200.h"
#include "fastq_host.hpp"

namespace fqg {

static bool has_extension(const char* path, const char* a, const char* b) {
    // Path::extension(): the part after the last '.' of the file name, unless the name starts with that dot
    const char* base = std::strrchr(path, '/');
    base = base ? base + 1 : path;
    const char* dot = std::strrchr(base, '.');
    if (!dot || dot == base) return false;
    return std::strcmp(dot + 1, a) == 0 || std::strcmp(dot + 1, b) == 0;
}
bool is_gzip_path(const char* path) { return has_extension(path, "gz", "bgz"); }       // mod.rs:170-175
bool is_fastq_path(const char* path) { return has_extension(path, "fastq", "fq"); }    // mod.rs:177-183

namespace {
struct Buffer {
    uint8_t* p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    bool reserve(size_t want, std::string* err) {
        if (want + 64 <= cap) return true;
        size_t ncap = cap ? cap : (1u << 20);
        while (ncap < want + 64) ncap += ncap / 2 + (1u << 20);
        uint8_t* q = nullptr;
        if (pinned) {
            if (cudaHostAlloc((void**)&q, ncap, cudaHostAllocDefault) != cudaSuccess) { *err = "cudaHostAlloc failed for the input buffer"; return false; }
        } else if (!(q = (uint8_t*)std::malloc(ncap))) { *err = "out of memory for the input buffer"; return false; }
        if (n) std::memcpy(q, p, n);
        release();
        p = q;
        cap = ncap;
        return true;
    }
    void release() {
        if (p) { if (pinned) cudaFreeHost(p); else std::free(p); }
        p = nullptr;
    }
};
}  // namespace

namespace {
struct BgzfBlock {
    size_t in_off, in_len;      // the raw deflate payload
    size_t out_off;
    uint32_t out_len, crc;
};
// Splits `z` into BGZF members.  False unless EVERY byte belongs to a well-formed member with a BC field.
bool bgzf_index(const std::vector<uint8_t>& z, std::vector<BgzfBlock>* blocks, size_t* total_out) {
    size_t o = 0, out = 0;
    while (o < z.size()) {
        if (z.size() - o < 28 || z[o] != 0x1f || z[o + 1] != 0x8b || z[o + 2] != 8 || !(z[o + 3] & 4)) return false;
        if (z[o + 3] & ~4) return false;                     // FNAME / FCOMMENT / FHCRC: not what bgzip writes
        const size_t xlen = z[o + 10] | (z[o + 11] << 8);
        if (o + 12 + xlen > z.size()) return false;
        size_t bsize = 0;
        for (size_t x = o + 12; x + 4 <= o + 12 + xlen;) {
            const size_t slen = z[x + 2] | (z[x + 3] << 8);
            if (z[x] == 'B' && z[x + 1] == 'C' && slen == 2 && x + 6 <= o + 12 + xlen) bsize = (size_t)(z[x + 4] | (z[x + 5] << 8)) + 1;
            x += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || o + bsize > z.size()) return false;
        const uint8_t* t = &z[o + bsize - 8];
        BgzfBlock b;
        b.in_off = o + 12 + xlen;
        b.in_len = bsize - 12 - xlen - 8;
        b.crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        b.out_len = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        b.out_off = out;
        if (b.out_len > (1u << 16)) return false;            // BGZF blocks hold at most 64 KiB
        out += b.out_len;
        blocks->push_back(b);
        o += bsize;
    }
    *total_out = out;
    return !blocks->empty();
}
// Inflates the members on all host threads; false on any size / CRC / stream error.
bool bgzf_inflate(const std::vector<uint8_t>& z, const std::vector<BgzfBlock>& blocks, uint8_t* out) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 64) nt = 64;
    if (nt > blocks.size() / 8 + 1) nt = (unsigned)(blocks.size() / 8 + 1);
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    auto work = [&]() {
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -MAX_WBITS) != Z_OK) { ok = false; return; }
        for (;;) {
            const size_t first = next.fetch_add(16);
            if (first >= blocks.size() || !ok) break;
            const size_t last = first + 16 < blocks.size() ? first + 16 : blocks.size();
            for (size_t i = first; i < last; i++) {
                const BgzfBlock& b = blocks[i];
                inflateReset(&zs);
                zs.next_in = (Bytef*)&z[b.in_off];
                zs.avail_in = (uInt)b.in_len;
                zs.next_out = out + b.out_off;
                zs.avail_out = b.out_len;
                const int zr = inflate(&zs, Z_FINISH);
                if (zr != Z_STREAM_END || zs.avail_out != 0 || zs.avail_in != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), out + b.out_off, b.out_len) != b.crc) { ok = false; break; }
            }
        }
        inflateEnd(&zs);
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return ok;
}
}  // namespace

int read_input(const char* path, int decompress, int pinned, uint8_t** out, size_t* n_out, std::string* err) {
    FILE* f = std::strcmp(path, "-") == 0 ? stdin : std::fopen(path, "rb");
    if (!f) { *err = std::string("Error opening input: ") + path; return FQG_E_INVALID_ARG; }
    const bool gz = is_gzip_path(path) || (!is_fastq_path(path) && decompress);   // main.rs:77
    Buffer buf;
    buf.pinned = pinned != 0;
    std::string raw_err;
    static const size_t kChunk = 1u << 20;
    std::string in(kChunk, '\0');
    int rc = FQG_OK;
    bool done = false;
    if (gz && f != stdin) {
        // regular file: look at the whole compressed image first; BGZF goes the parallel way
        std::vector<uint8_t> z;
        if (std::fseek(f, 0, SEEK_END) == 0) {
            const long sz = std::ftell(f);
            std::rewind(f);
            if (sz > 0) {
                z.resize((size_t)sz);
                if (std::fread(z.data(), 1, z.size(), f) != z.size()) z.clear();
                std::rewind(f);
            }
        }
        std::vector<BgzfBlock> blocks;
        size_t total = 0;
        if (!z.empty() && bgzf_index(z, &blocks, &total)) {
            if (!buf.reserve(total + 1, err)) rc = FQG_E_CUDA;
            else if (!bgzf_inflate(z, blocks, buf.p)) { *err = std::string("gzip decode error in ") + path; rc = FQG_E_FASTQ; }
            else buf.n = total;
            done = true;
        }
    }
    if (done) {
    } else if (!gz) {
        if (f != stdin && std::fseek(f, 0, SEEK_END) == 0) {      // regular file: one allocation of the right size
            const long sz = std::ftell(f);
            std::rewind(f);
            if (sz > 0 && !buf.reserve((size_t)sz + kChunk + 1, err)) rc = FQG_E_CUDA;      // (+ one chunk: the loop asks for a full chunk of room)
        }
        while (rc == FQG_OK) {
            if (!buf.reserve(buf.n + kChunk, err)) { rc = FQG_E_CUDA; break; }
            size_t k = std::fread(buf.p + buf.n, 1, kChunk, f);
            buf.n += k;
            if (k < kChunk) break;
        }
    } else {
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) { *err = "zlib init failed"; rc = FQG_E_INVALID_ARG; }
        bool stream_open = true;
        size_t member_in = 0;      // compressed bytes consumed by the member being inflated (inflateReset zeroes total_in)
        while (rc == FQG_OK) {
            size_t k = std::fread(&in[0], 1, kChunk, f);
            if (k == 0) break;
            zs.next_in = (Bytef*)&in[0];
            zs.avail_in = (uInt)k;
            while (zs.avail_in && rc == FQG_OK) {
                if (!stream_open) {               // MultiGzDecoder: another member follows
                    inflateReset(&zs);
                    stream_open = true;
                }
                if (!buf.reserve(buf.n + kChunk, err)) { rc = FQG_E_CUDA; break; }
                zs.next_out = buf.p + buf.n;
                const size_t room = buf.cap - 64 - buf.n;
                zs.avail_out = (uInt)(room > (1u << 30) ? (1u << 30) : room);
                const uInt before = zs.avail_out, before_in = zs.avail_in;
                int zr = inflate(&zs, Z_NO_FLUSH);
                buf.n += before - zs.avail_out;
                member_in += before_in - zs.avail_in;
                if (zr == Z_STREAM_END) { stream_open = false; member_in = 0; }
                else if (zr != Z_OK && zr != Z_BUF_ERROR) { *err = std::string("gzip decode error in ") + path; rc = FQG_E_FASTQ; }
            }
        }
        // The input ended inside a member -- the first or any later one (flate2's MultiGzDecoder raises UnexpectedEof there,
        // which the reference turns into a panic, exit 101): partial data must never be returned as if it were the file.
        if (rc == FQG_OK && stream_open && member_in) { *err = std::string("truncated gzip stream in ") + path; rc = FQG_E_FASTQ; }
        inflateEnd(&zs);
    }
    if (f != stdin) std::fclose(f);
    if (rc != FQG_OK) { buf.release(); return rc; }
    if (!buf.p && !buf.reserve(1, err)) return FQG_E_CUDA;
    *out = buf.p;
    *n_out = buf.n;
    return FQG_OK;
}

void free_input(uint8_t* p, int pinned) {
    if (!p) return;
    if (pinned) cudaFreeHost(p); else std::free(p);
}

}  // namespace fqg
