// ac_build.cc -- literal set -> Aho-Corasick automaton, emitted as the same dense byte-class DFA the
// regex path produces, so one CUDA kernel family serves -F, -F -f and -e.
//
// Replaces aho_corasick::AhoCorasick::new(&patterns) + is_match (/root/reference/src/lib/matcher.rs:297-298, :250)
// and bstr's find for the single-pattern case (:204).  is_match only needs "some pattern occurs", so every
// node that completes a pattern (directly or through a suffix link) collapses into one absorbing state and
// the trie below it is never built.  --reverse-complement adds rc(pattern) for each pattern (mod.rs:145-155:
// P occurs in rc(seq) <=> rc(P) occurs in seq, because complement is a byte involution).
#include "automata.hpp"

#include <algorithm>
#include <deque>

namespace fqg {

bool compile_literal_set(const std::vector<std::string>& patterns_in, bool with_revcomp, Dfa* out, std::string* err) {
    const uint8_t* comp = complement_table();
    std::vector<std::string> pats = patterns_in;
    if (with_revcomp)
        for (auto& p : patterns_in) {
            std::string r(p.rbegin(), p.rend());
            for (auto& ch : r) ch = (char)comp[(uint8_t)ch];
            pats.push_back(std::move(r));
        }
    Dfa d;
    // byte classes: every byte that occurs in a pattern is its own class, everything else shares class 0
    bool used[256] = {false};
    for (auto& p : pats) for (unsigned char ch : p) used[ch] = true;
    uint32_t ncls = 1;
    for (int b = 0; b < 256; b++) d.byte_class[b] = used[b] ? (uint8_t)(ncls++) : 0;
    if (ncls > 255) { *err = "Failed to build Aho-Corasick automaton (alphabet too large)"; return false; }
    d.n_classes = ncls;

    bool any_empty = false;
    for (auto& p : pats) if (p.empty()) any_empty = true;
    if (any_empty) {   // the empty needle occurs in every haystack
        std::vector<uint32_t> next(ncls, 0);
        std::vector<uint8_t> acc{1};
        finalize_dfa(1, ncls, next, acc, 0, false, &d);
        *out = std::move(d);
        return true;
    }

    // trie; node 0 = root
    std::vector<int32_t> child(ncls, -1);
    std::vector<uint8_t> term(1, 0);
    for (auto& p : pats) {
        uint32_t u = 0;
        bool dead = false;
        for (unsigned char ch : p) {
            if (term[u]) { dead = true; break; }   // a shorter pattern is a prefix: this one can never be the first hit
            uint32_t c = d.byte_class[ch];
            int32_t v = child[(size_t)u * ncls + c];
            if (v < 0) {
                v = (int32_t)term.size();
                child[(size_t)u * ncls + c] = v;
                term.push_back(0);
                child.resize(child.size() + ncls, -1);
            }
            u = (uint32_t)v;
        }
        if (!dead) term[u] = 1;
    }
    uint32_t n = (uint32_t)term.size();
    if (n > (1u << 24)) { *err = "Failed to build Aho-Corasick automaton (more than 2^24 trie nodes)"; return false; }
    // BFS: failure links folded straight into a total transition function
    std::vector<uint32_t> delta((size_t)n * ncls, 0), fail(n, 0);
    std::deque<uint32_t> q;
    for (uint32_t c = 0; c < ncls; c++) {
        int32_t v = child[c];
        if (v >= 0) { delta[c] = (uint32_t)v; fail[v] = 0; q.push_back((uint32_t)v); }
        else delta[c] = 0;
    }
    while (!q.empty()) {
        uint32_t u = q.front();
        q.pop_front();
        term[u] |= term[fail[u]];
        for (uint32_t c = 0; c < ncls; c++) {
            int32_t v = child[(size_t)u * ncls + c];
            uint32_t via = delta[(size_t)fail[u] * ncls + c];
            if (v >= 0) { fail[v] = via; delta[(size_t)u * ncls + c] = (uint32_t)v; q.push_back((uint32_t)v); }
            else delta[(size_t)u * ncls + c] = via;
        }
    }
    // collapse terminal nodes into one absorbing state; keep only nodes reachable without passing a terminal
    const uint32_t kNone = UINT32_MAX;
    std::vector<uint32_t> id(n, kNone);
    std::vector<uint32_t> order;
    q.push_back(0);
    id[0] = 0;
    order.push_back(0);
    bool hit = false;
    while (!q.empty()) {
        uint32_t u = q.front();
        q.pop_front();
        for (uint32_t c = 0; c < ncls; c++) {
            uint32_t v = delta[(size_t)u * ncls + c];
            if (term[v]) { hit = true; continue; }
            if (id[v] == kNone) { id[v] = (uint32_t)order.size(); order.push_back(v); q.push_back(v); }
        }
    }
    uint32_t m = (uint32_t)order.size();
    uint32_t match_id = m;
    uint32_t total = m + (hit ? 1 : 0);
    std::vector<uint32_t> next((size_t)total * ncls, match_id);
    std::vector<uint8_t> acc(total, 0);
    for (uint32_t k = 0; k < m; k++)
        for (uint32_t c = 0; c < ncls; c++) {
            uint32_t v = delta[(size_t)order[k] * ncls + c];
            next[(size_t)k * ncls + c] = term[v] ? match_id : id[v];
        }
    if (hit) acc[match_id] = 1;
    finalize_dfa(total, ncls, next, acc, 0, total <= 20000, &d);
    *out = std::move(d);
    return true;
}

}  // namespace fqg
