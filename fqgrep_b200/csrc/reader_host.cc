// reader_host.cc -- incremental host input for the streaming path: file / stdin -> (optional multi-member gzip inflate) ->
// the caller's buffer, a slot at a time.  Replaces spawn_reader (/root/reference/src/main.rs:61-87: File or stdin ->
// BufReader -> optional flate2 MultiGzDecoder -> fastq::Reader) for inputs that must not be held in memory whole; the
// extension rules are those of src/lib/mod.rs:157-183 (is_gzip_path / is_fastq_path in input_host.cc).
// gzip decoding stays on the host (north_star: excluded from the headline).  BGZF members (independent, <= 64 KiB, size in
// a 'BC' extra field) are located without inflating and then inflated on all host threads straight into the caller's
// buffer; anything else is one serial zlib stream.  Truncated members are errors wherever the cut falls.
#include <zlib.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fqgrep_b200.h"
#include "fastq_host.hpp"

namespace fqg {

struct InputReader {
    FILE* f = nullptr;
    bool own = false, gz = false, file_eof = false;
    std::string path;
    // compressed bytes read from the file and not yet consumed
    std::vector<uint8_t> in;
    size_t in_pos = 0, in_len = 0;
    // serial inflate state
    z_stream zs;
    bool zs_init = false, stream_open = false;
    size_t member_in = 0;
    bool bgzf = true;      // until a member that is not a BGZF block shows up
    ~InputReader() {
        if (zs_init) inflateEnd(&zs);
        if (f && own) std::fclose(f);
    }
    // makes at least `want` unconsumed bytes available if the file has them
    void refill(size_t want) {
        if (in_len - in_pos >= want || file_eof) return;
        if (in_pos) {
            std::memmove(in.data(), in.data() + in_pos, in_len - in_pos);
            in_len -= in_pos;
            in_pos = 0;
        }
        const size_t target = want > ((size_t)4 << 20) ? want : ((size_t)4 << 20);
        if (in.size() < target) in.resize(target);
        while (in_len < target && !file_eof) {
            const size_t k = std::fread(in.data() + in_len, 1, in.size() - in_len, f);
            in_len += k;
            if (k == 0) file_eof = true;
        }
    }
};

int reader_open(const char* path, int decompress, InputReader** out, std::string* err) {
    auto r = new InputReader();
    r->path = path;
    if (std::strcmp(path, "-") == 0) r->f = stdin;
    else {
        r->f = std::fopen(path, "rb");
        r->own = true;
    }
    if (!r->f) {
        *err = std::string("Error opening input: ") + path;
        delete r;
        return FQG_E_INVALID_ARG;
    }
    r->gz = is_gzip_path(path) || (!is_fastq_path(path) && decompress);   // main.rs:77
    *out = r;
    return FQG_OK;
}
void reader_close(InputReader* r) { delete r; }

namespace {
struct Block { size_t in_off, in_len, out_off; uint32_t out_len, crc; };

// one BGZF block at r->in[pos..): 1 = parsed, 0 = not (yet) whole in the buffer, -1 = not a BGZF block
int parse_bgzf(const InputReader* r, size_t pos, size_t* bsize, Block* b) {
    const uint8_t* z = r->in.data();
    const size_t have = r->in_len - pos;
    if (have < 18) return r->file_eof && have == 0 ? 0 : (r->file_eof ? -1 : 0);
    if (z[pos] != 0x1f || z[pos + 1] != 0x8b || z[pos + 2] != 8 || z[pos + 3] != 4) return -1;
    const size_t xlen = z[pos + 10] | (z[pos + 11] << 8);
    if (have < 12 + xlen) return r->file_eof ? -1 : 0;
    size_t bs = 0;
    for (size_t x = pos + 12; x + 4 <= pos + 12 + xlen;) {
        const size_t slen = z[x + 2] | (z[x + 3] << 8);
        if (z[x] == 'B' && z[x + 1] == 'C' && slen == 2 && x + 6 <= pos + 12 + xlen) bs = (size_t)(z[x + 4] | (z[x + 5] << 8)) + 1;
        x += 4 + slen;
    }
    if (bs < 12 + xlen + 8) return -1;
    if (have < bs) return r->file_eof ? -1 : 0;
    const uint8_t* t = z + pos + bs - 8;
    b->in_off = pos + 12 + xlen;
    b->in_len = bs - 12 - xlen - 8;
    b->crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    b->out_len = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
    if (b->out_len > (1u << 16)) return -1;
    *bsize = bs;
    return 1;
}

bool inflate_blocks(const uint8_t* z, const std::vector<Block>& blocks, uint8_t* dst) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 32) nt = 32;
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
                const Block& b = blocks[i];
                inflateReset(&zs);
                zs.next_in = (Bytef*)(z + b.in_off);
                zs.avail_in = (uInt)b.in_len;
                zs.next_out = dst + b.out_off;
                zs.avail_out = b.out_len;
                const int zr = inflate(&zs, Z_FINISH);
                if (zr != Z_STREAM_END || zs.avail_out != 0 || zs.avail_in != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst + b.out_off, b.out_len) != b.crc) { ok = false; break; }
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

// Up to `cap` bytes of (decompressed) input into dst.  *n == 0 with *eof set means the input is exhausted.
int reader_read(InputReader* r, uint8_t* dst, size_t cap, size_t* n, bool* eof, std::string* err) {
    *n = 0;
    *eof = false;
    if (!r->gz) {
        while (*n < cap) {
            const size_t k = std::fread(dst + *n, 1, cap - *n, r->f);
            *n += k;
            if (k == 0) { *eof = true; break; }
        }
        return FQG_OK;
    }
    // ---- BGZF: whole members, inflated in parallel, as long as the input keeps being BGZF
    if (cap < ((size_t)1 << 16)) r->bgzf = false;      // a block may not fit: tiny buffers take the serial decoder
    if (r->bgzf && !r->stream_open) {
        std::vector<Block> blocks;
        size_t out = 0;
        for (;;) {
            r->refill((size_t)1 << 17);      // a block is at most 64 KiB: two of them are always enough to decide
            size_t pos = r->in_pos;
            bool stop = false;
            blocks.clear();
            const size_t base_out = out;
            while (!stop) {
                Block b;
                size_t bs = 0;
                const int pr = parse_bgzf(r, pos, &bs, &b);
                if (pr == 1) {
                    if (out + b.out_len > cap) { stop = true; break; }
                    b.out_off = out;
                    out += b.out_len;
                    blocks.push_back(b);
                    pos += bs;
                } else {
                    if (pr < 0) r->bgzf = false;      // from here on: the serial decoder
                    stop = true;
                }
            }
            if (!blocks.empty()) {
                if (!inflate_blocks(r->in.data(), blocks, dst)) { *err = std::string("gzip decode error in ") + r->path; return FQG_E_FASTQ; }
                r->in_pos = pos;
            } else out = base_out;
            const bool more_input = !(r->file_eof && r->in_pos >= r->in_len);
            // stop when the buffer is (nearly) full, the input is not BGZF any more, or the file has ended
            if (!r->bgzf || !more_input || cap - out < (1u << 16)) break;
            if (blocks.empty() && r->in_len - r->in_pos >= ((size_t)1 << 17)) break;      // a block that does not fit what is left of cap
        }
        *n = out;
        if (r->bgzf || out == cap) {
            *eof = r->file_eof && r->in_pos >= r->in_len;
            if (*n > 0 || *eof) return FQG_OK;
        }
        // fall through: the next member is plain gzip
    }
    // ---- serial multi-member gzip (MultiGzDecoder)
    if (!r->zs_init) {
        std::memset(&r->zs, 0, sizeof r->zs);
        if (inflateInit2(&r->zs, 16 + MAX_WBITS) != Z_OK) { *err = "zlib init failed"; return FQG_E_INVALID_ARG; }
        r->zs_init = true;
        r->stream_open = false;
        r->member_in = 0;
    }
    while (*n < cap) {
        r->refill(1);
        if (r->in_pos >= r->in_len) {      // end of the compressed input
            if (r->stream_open && r->member_in) { *err = std::string("truncated gzip stream in ") + r->path; return FQG_E_FASTQ; }
            *eof = true;
            break;
        }
        if (!r->stream_open) {               // MultiGzDecoder: another member follows
            inflateReset(&r->zs);
            r->stream_open = true;
            r->member_in = 0;
        }
        r->zs.next_in = r->in.data() + r->in_pos;
        const size_t avail = r->in_len - r->in_pos;
        r->zs.avail_in = (uInt)(avail > (1u << 30) ? (1u << 30) : avail);
        r->zs.next_out = dst + *n;
        const size_t room = cap - *n;
        r->zs.avail_out = (uInt)(room > (1u << 30) ? (1u << 30) : room);
        const uInt before_out = r->zs.avail_out, before_in = r->zs.avail_in;
        const int zr = inflate(&r->zs, Z_NO_FLUSH);
        *n += before_out - r->zs.avail_out;
        r->in_pos += before_in - r->zs.avail_in;
        r->member_in += before_in - r->zs.avail_in;
        if (zr == Z_STREAM_END) { r->stream_open = false; r->member_in = 0; }
        else if (zr != Z_OK && zr != Z_BUF_ERROR) { *err = std::string("gzip decode error in ") + r->path; return FQG_E_FASTQ; }
        else if (zr == Z_BUF_ERROR && before_in == r->zs.avail_in && before_out == r->zs.avail_out && r->file_eof) {
            *err = std::string("truncated gzip stream in ") + r->path;
            return FQG_E_FASTQ;
        }
    }
    if (*n == cap && !*eof) {
        r->refill(1);
        if (r->in_pos >= r->in_len && !(r->stream_open && r->member_in)) *eof = true;
    }
    return FQG_OK;
}

}  // namespace fqg
