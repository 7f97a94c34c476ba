// automata.hpp -- host-side pattern compilers for the B200 read matcher.
//
// Every matcher kind of the reference's MatcherFactory (/root/reference/src/lib/matcher.rs:645-685)
// that looks at bases is lowered to ONE dense byte-class DFA answering
//     f(seq) || (reverse_complement && f(rc(seq)))
// in a single left-to-right pass over the read, so the device never materialises rc(seq)
// (the reference allocates it per read: src/lib/mod.rs:145-155).  The final `!= invert`
// (matcher.rs:204,250,507,570 folded by :164-175) is one XOR in the kernel epilogue.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fqg {

// COMPLEMENT table of src/lib/mod.rs:129-143 (IUPAC letters + lowercase, identity elsewhere).
const uint8_t* complement_table();

struct Dfa {
    uint32_t n_states = 0;
    uint32_t n_classes = 0;
    uint8_t byte_class[256] = {0};
    std::vector<uint32_t> next;   // [n_states][n_classes], state ids
    uint32_t start = 0;
    // States are numbered so that `state >= accept_from` <=> "the read is selected if input ends here".
    // The absorbing match state (if any) is the last one.
    uint32_t accept_from = 0;
    bool always = false;          // matches every read (empty pattern), table still valid
};

struct CompileError {
    std::string message;
};

// Regex set -> DFA (regex-syntax 0.8 surface, is_match semantics).
// patterns are UTF-8 byte strings.  with_revcomp adds the reversed+complemented language of every member.
// Returns false and fills err (message starts with "Invalid regular expression: '<p>'" like matcher.rs:499).
bool compile_regex_set(const std::vector<std::string>& patterns, bool with_revcomp, Dfa* out, std::string* err);

// --iupac bit-mask needles (BitMaskMatcher / BitMaskSetMatcher, matcher.rs:307-487) -> DFA.
bool compile_bitmask_set(const std::vector<std::string>& patterns, bool with_revcomp, Dfa* out, std::string* err);
const uint8_t* iupac_masks();   // IUPAC_MASKS, mod.rs:14-53
// --iupac expand / regex rewrites (mod.rs:70-118, driven from main.rs:339-360)
bool expand_iupac_fixed(const std::string& pattern, std::vector<std::string>* out, std::string* err);
std::string iupac_to_regex(const std::string& pattern);

bool regex_as_literal(const std::string& pattern, std::string* literal);
// the regex's whole (finite, small, non-empty-string) language, see regex_compile.cc
bool regex_as_literals(const std::string& pattern, size_t max_strings, std::vector<std::string>* literals);

// Literal set -> Aho-Corasick DFA (replaces aho_corasick::AhoCorasick::new, matcher.rs:297-298, and
// bstr find, matcher.rs:204).  with_revcomp adds rc(pattern) for every pattern.
bool compile_literal_set(const std::vector<std::string>& patterns, bool with_revcomp, Dfa* out, std::string* err);

// Shared back end: optional Moore minimisation, then renumbering (accepting states last).
void finalize_dfa(uint32_t n_states, uint32_t n_classes, std::vector<uint32_t>& next, std::vector<uint8_t>& accept, uint32_t start,
                  bool minimize, Dfa* out);

// Large literal sets: q-gram sampling filter + exact verification table (gram_filter.cc).
struct GramFilter {
    bool usable = false;
    uint32_t q = 0, s = 0, min_len = 0, qmask = 0;
    uint32_t bitmap_bits_log2 = 0, slots_log2 = 0;
    double expected_hits = 0;          // real gram hits a 150-base read of random bases sends to verification
    std::vector<uint32_t> bitmap;      // shared-memory resident on the device
    std::vector<uint32_t> slots;       // 4 words per slot {gram key, pool offset, pattern length << 5 | offset in the pattern, -}; all ones = empty
    std::vector<uint8_t> pool;         // patterns (+ reverse complements), each starting on a 16-byte boundary
};
uint32_t gram_code(const uint8_t* p, uint32_t q);
bool build_gram_filter(const std::vector<std::string>& patterns, bool with_revcomp, GramFilter* out);

// validate_fixed_pattern (matcher.rs:128-147); returns empty string when ok, else the reference's message.
std::string validate_fixed_pattern(const std::string& pattern, bool protein);

// Read-name set -> open-addressing table (replaces AHashSet<Vec<u8>>, matcher.rs:604,632).
struct NameTable {
    uint32_t n_slots = 0;                 // power of two
    std::vector<uint64_t> slots;          // 0 = empty; else tag << 32 | offset of the name's pool entry (names_hash.h)
    std::vector<uint8_t> pool;            // entries on 16-byte boundaries: u32 length, then the bytes, zero-padded
    uint32_t max_len = 0;
};
uint64_t name_hash(const uint8_t* p, size_t n);
void build_name_table(const std::vector<std::string>& names, NameTable* out);

}  // namespace fqg
