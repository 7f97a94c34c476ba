// names.cc -- read-name set -> open-addressing table probed on the device.
// Replaces AHashSet<Vec<u8>> (/root/reference/src/lib/matcher.rs:604, :632; filled at src/main.rs:258-271).
// Membership is exact: the 64-bit hash only picks the bucket and a 32-bit tag, the kernel then compares length and bytes.
#include "automata.hpp"
#include "names_hash.h"

#include <unordered_set>

namespace fqg {

uint64_t name_hash(const uint8_t* p, size_t n) { return fqg_name_hash(p, (uint32_t)n); }

void build_name_table(const std::vector<std::string>& names, NameTable* out) {
    std::unordered_set<std::string> uniq(names.begin(), names.end());
    // 32-byte buckets of four slots, at least one bucket per name (load <= 0.25): a lookup that misses reads one
    // bucket (one L2 sector) 98 % of the time.  A bucket fills front to back and overflows into the next bucket.
    // A slot carries the offset of the name's pool entry next to its tag, so a hit costs ONE further dependent read (the
    // round-1 table kept the offsets in a separate array: bucket -> offset -> bytes, three round trips per hit).
    // (FQG_NAMES_PER_BUCKET = 2, i.e. half the table at load 0.5, was measured: more of it stays in L2 but 14 % of the buckets
    // overflow -- 79 instead of 94 Greads/s over SoA headers, 7.1 instead of 7.9 from raw FASTQ.)
#ifndef FQG_NAMES_PER_BUCKET
#define FQG_NAMES_PER_BUCKET 1
#endif
    uint32_t buckets = 4;
    while ((size_t)buckets * FQG_NAMES_PER_BUCKET < uniq.size() + 1) buckets <<= 1;
    const uint32_t slots = buckets * 4;
    out->n_slots = slots;
    out->slots.assign(slots, 0);
    out->pool.clear();
    out->max_len = 0;
    for (auto& s : uniq) {
        const uint64_t h = fqg_name_hash((const uint8_t*)s.data(), (uint32_t)s.size());
        uint32_t i = fqg_name_bucket(h, buckets - 1) * 4;
        while (out->slots[i]) i = (i + 1) & (slots - 1);
        const size_t off = out->pool.size();      // a multiple of 16
        out->slots[i] = ((uint64_t)fqg_name_tag(h) << 32) | (uint32_t)off;
        const uint32_t len = (uint32_t)s.size();
        for (int k = 0; k < 4; k++) out->pool.push_back((uint8_t)(len >> (8 * k)));
        out->pool.insert(out->pool.end(), s.begin(), s.end());
        out->pool.resize((out->pool.size() + 15) & ~(size_t)15, 0);      // zero-padded: compared four bytes at a time
        if (s.size() > out->max_len) out->max_len = (uint32_t)s.size();
    }
    out->pool.resize(out->pool.size() + 16, 0);
}

}  // namespace fqg
