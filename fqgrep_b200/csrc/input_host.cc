// input_host.cc -- host input plumbing: file / stdin -> (optional multi-member gzip inflate) -> one host buffer that
// the FASTQ entry points stream to the GPU.  Replaces spawn_reader (/root/reference/src/main.rs:61-87: File or stdin ->
// BufReader -> optional flate2 MultiGzDecoder -> fastq::Reader) and the extension rules of src/lib/mod.rs:157-183.
// gzip decoding stays on the host (north_star: excluded from the headline); BGZF is a multi-member gzip, so the same
// loop handles .bgz.  No matching happens here.
#include <zlib.h>

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
    if (!gz) {
        for (;;) {
            if (!buf.reserve(buf.n + kChunk, err)) { rc = FQG_E_CUDA; break; }
            size_t k = std::fread(buf.p + buf.n, 1, kChunk, f);
            buf.n += k;
            if (k < kChunk) break;
        }
    } else {
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) { *err = "zlib init failed"; rc = FQG_E_INVALID_ARG; }
        bool stream_open = true, any_member = false;
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
                const uInt before = zs.avail_out;
                int zr = inflate(&zs, Z_NO_FLUSH);
                buf.n += before - zs.avail_out;
                if (zr == Z_STREAM_END) { stream_open = false; any_member = true; }
                else if (zr != Z_OK && zr != Z_BUF_ERROR) { *err = std::string("gzip decode error in ") + path; rc = FQG_E_FASTQ; }
            }
        }
        if (rc == FQG_OK && stream_open && (any_member || zs.total_in)) {
            if (zs.total_in && !any_member) { *err = std::string("truncated gzip stream in ") + path; rc = FQG_E_FASTQ; }
        }
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
