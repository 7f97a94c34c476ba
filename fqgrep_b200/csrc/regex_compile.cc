// regex_compile.cc -- regex set -> byte-class DFA, ahead of time, on the host.
//
// Replaces, for the device path, what the reference gets from the `regex` crate at
// /root/reference/src/lib/matcher.rs:497-499 (RegexBuilder::new(p).build(), all defaults) and
// :551-554 (RegexSetBuilder), used only through is_match (:507, :570).  Only the LANGUAGE matters
// for is_match ("some substring of the read is in L(p)"), so greediness, captures and match priority
// are dropped and the whole set (plus, for --reverse-complement, the reversed+complemented language
// of every member) is determinised into a single DFA that the CUDA kernel walks once per read.
//
// Surface: regex-syntax 0.8 as restated in SURVEY.md Appendix A.  Deviations (all confined to
// non-ASCII haystack bytes, i.e. never DNA/protein reads) are listed in DESIGN.md: \p{..} is
// rejected, \w \d are ASCII-only, case folding is ASCII plus the two ASCII-reaching Unicode folds.
#include "automata.hpp"

#include <algorithm>
#include <array>
#include <cstring>
#include <deque>
#include <memory>
#include <unordered_map>

namespace fqg {

const uint8_t* complement_table() {
    static uint8_t tab[256];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; i++) tab[i] = (uint8_t)i;
        const char* a = "AGCTYRWSKMDVHBN";
        const char* b = "TCGARYWSMKHBDVN";
        for (int i = 0; a[i]; i++) {
            tab[(uint8_t)a[i]] = (uint8_t)b[i];
            tab[(uint8_t)a[i] + 32] = (uint8_t)(b[i] + 32);
        }
        init = true;
    }
    return tab;
}

namespace {

constexpr uint32_t kMaxScalar = 0x10FFFF;
constexpr size_t kMaxNfaStates = 400000;
constexpr size_t kMaxDfaStates = 1u << 20;

struct ParseError {
    std::string msg;
};

// ---------------------------------------------------------------- scalar-range sets
using Range = std::pair<uint32_t, uint32_t>;
struct ScalarSet {
    std::vector<Range> r;
    void add(uint32_t lo, uint32_t hi) { r.emplace_back(lo, hi); }
    void canon() {
        std::sort(r.begin(), r.end());
        std::vector<Range> o;
        for (auto& x : r) {
            if (!o.empty() && x.first <= o.back().second + 1) o.back().second = std::max(o.back().second, x.second);
            else o.push_back(x);
        }
        r.swap(o);
    }
    void unite(const ScalarSet& b) { r.insert(r.end(), b.r.begin(), b.r.end()); canon(); }
    void negate(uint32_t maxv) {
        canon();
        std::vector<Range> o;
        uint32_t nxt = 0;
        bool open = true;
        for (auto& x : r) {
            if (x.first > nxt) o.emplace_back(nxt, x.first - 1);
            if (x.second >= maxv) { open = false; break; }
            nxt = x.second + 1;
        }
        if (open && nxt <= maxv) o.emplace_back(nxt, maxv);
        r.swap(o);
    }
    void intersect(const ScalarSet& b) {
        std::vector<Range> o;
        for (auto& x : r)
            for (auto& y : b.r) {
                uint32_t lo = std::max(x.first, y.first), hi = std::min(x.second, y.second);
                if (lo <= hi) o.emplace_back(lo, hi);
            }
        r.swap(o);
        canon();
    }
    void subtract(ScalarSet b) { b.negate(kMaxScalar); intersect(b); }
    void symdiff(const ScalarSet& b) {
        ScalarSet u = *this, i = *this;
        u.unite(b);
        i.intersect(b);
        u.subtract(i);
        r.swap(u.r);
    }
    bool has(uint32_t c) const {
        for (auto& x : r) if (x.first <= c && c <= x.second) return true;
        return false;
    }
};

// ---------------------------------------------------------------- AST over bytes
using ByteSet = std::array<uint64_t, 4>;
inline void bs_set(ByteSet& s, int b) { s[b >> 6] |= 1ull << (b & 63); }
inline bool bs_has(const ByteSet& s, int b) { return (s[b >> 6] >> (b & 63)) & 1; }

enum class Look : uint8_t { StartText, EndText, StartLine, EndLine, WordB, NotWordB, WordStart, WordEnd, WordStartHalf, WordEndHalf };

struct Ast;
using AstP = std::unique_ptr<Ast>;
struct Ast {
    enum Kind { Empty, Bytes, Cat, Alt, Rep, Assert, Never } kind = Empty;
    ByteSet set{};
    std::vector<AstP> kids;
    int min = 0, max = -1;
    Look look = Look::StartText;
};
AstP mk(Ast::Kind k) { auto a = std::make_unique<Ast>(); a->kind = k; return a; }
AstP mk_bytes(int lo, int hi) { auto a = mk(Ast::Bytes); for (int b = lo; b <= hi; b++) bs_set(a->set, b); return a; }
AstP mk_alt2(AstP a, AstP b) {
    if (!a) return b;
    if (!b) return a;
    auto x = mk(Ast::Alt);
    if (a->kind == Ast::Alt) x->kids = std::move(a->kids); else x->kids.push_back(std::move(a));
    x->kids.push_back(std::move(b));
    return x;
}
AstP mk_cat2(AstP a, AstP b) {
    if (!a) return b;
    if (!b) return a;
    auto x = mk(Ast::Cat);
    if (a->kind == Ast::Cat) x->kids = std::move(a->kids); else x->kids.push_back(std::move(a));
    x->kids.push_back(std::move(b));
    return x;
}

int utf8_encode(uint32_t cp, uint8_t* o) {
    if (cp < 0x80) { o[0] = (uint8_t)cp; return 1; }
    if (cp < 0x800) { o[0] = 0xC0 | (cp >> 6); o[1] = 0x80 | (cp & 63); return 2; }
    if (cp < 0x10000) { o[0] = 0xE0 | (cp >> 12); o[1] = 0x80 | ((cp >> 6) & 63); o[2] = 0x80 | (cp & 63); return 3; }
    o[0] = 0xF0 | (cp >> 18); o[1] = 0x80 | ((cp >> 12) & 63); o[2] = 0x80 | ((cp >> 6) & 63); o[3] = 0x80 | (cp & 63);
    return 4;
}

// Scalar range -> alternation of byte-range sequences.  Ranges are first cut at the surrogate gap and
// at the 1/2/3/4-byte encoding boundaries, then at continuation-byte boundaries so each piece is a
// product of per-byte ranges.
void utf8_pieces(uint32_t lo, uint32_t hi, std::vector<std::pair<uint32_t, uint32_t>>& out) {
    if (lo > hi) return;
    if (lo <= 0xDFFF && hi >= 0xD800) {
        if (lo < 0xD800) utf8_pieces(lo, 0xD7FF, out);
        if (hi > 0xDFFF) utf8_pieces(0xE000, hi, out);
        return;
    }
    for (uint32_t lim : {0x7Fu, 0x7FFu, 0xFFFFu})
        if (lo <= lim && hi > lim) { utf8_pieces(lo, lim, out); utf8_pieces(lim + 1, hi, out); return; }
    if (hi >= 0x80)
        for (int i = 1; i < 4; i++) {
            uint32_t m = (1u << (6 * i)) - 1;
            if ((lo & ~m) != (hi & ~m)) {
                if ((lo & m) != 0) { utf8_pieces(lo, lo | m, out); utf8_pieces((lo | m) + 1, hi, out); return; }
                if ((hi & m) != m) { utf8_pieces(lo, (hi & ~m) - 1, out); utf8_pieces(hi & ~m, hi, out); return; }
            }
        }
    out.emplace_back(lo, hi);
}
AstP scalar_set_ast(const ScalarSet& s, bool unicode) {
    if (!unicode) {
        auto b = mk(Ast::Bytes);
        bool any = false;
        for (auto& x : s.r)
            for (uint32_t c = x.first; c <= x.second && c < 256; c++) { bs_set(b->set, (int)c); any = true; }
        return any ? std::move(b) : mk(Ast::Never);
    }
    AstP alt;
    auto ascii = mk(Ast::Bytes);
    bool any_ascii = false;
    std::vector<std::pair<uint32_t, uint32_t>> pieces;
    for (auto& x : s.r) utf8_pieces(x.first, x.second, pieces);
    for (auto& p : pieces) {
        if (p.second < 0x80) { for (uint32_t c = p.first; c <= p.second; c++) bs_set(ascii->set, (int)c); any_ascii = true; continue; }
        uint8_t a[4], b[4];
        int n = utf8_encode(p.first, a);
        utf8_encode(p.second, b);
        AstP seq;
        for (int k = 0; k < n; k++) seq = mk_cat2(std::move(seq), mk_bytes(a[k], b[k]));
        alt = mk_alt2(std::move(alt), std::move(seq));
    }
    if (any_ascii) alt = mk_alt2(std::move(ascii), std::move(alt));
    return alt ? std::move(alt) : mk(Ast::Never);
}

inline bool is_word_byte(int c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_'; }

// ---------------------------------------------------------------- parser
struct Flags {
    bool i = false, m = false, s = false, u = true, x = false, R = false;
};

class Parser {
public:
    explicit Parser(const std::string& p) : p_(p) {}
    AstP parse() {
        AstP a = alternation();
        if (pos_ < p_.size()) {
            if (p_[pos_] == ')') throw ParseError{"unopened group"};
            throw ParseError{"unexpected character"};
        }
        return a;
    }

private:
    const std::string& p_;
    size_t pos_ = 0;
    Flags f_;
    int depth_ = 0;

    int peek(size_t k = 0) const { return pos_ + k < p_.size() ? (uint8_t)p_[pos_ + k] : -1; }
    bool eat(int c) { if (peek() == c) { pos_++; return true; } return false; }
    uint32_t next_scalar() {
        if (pos_ >= p_.size()) throw ParseError{"unexpected end of pattern"};
        uint32_t c = (uint8_t)p_[pos_++];
        if (c < 0x80) return c;
        int n = c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : 1;
        uint32_t cp = c & (0x3F >> n);
        while (n-- > 0 && pos_ < p_.size()) cp = (cp << 6) | ((uint8_t)p_[pos_++] & 63);
        return cp;
    }
    void skip_ws() {
        if (!f_.x) return;
        for (;;) {
            int c = peek();
            if (c == ' ' || (c >= 9 && c <= 13)) pos_++;
            else if (c == '#') { while (pos_ < p_.size() && p_[pos_] != '\n') pos_++; }
            else break;
        }
    }
    uint32_t top() const { return f_.u ? kMaxScalar : 255; }

    static void perl_class(ScalarSet& s, int which, bool unicode) {
        if (which == 'd') s.add('0', '9');
        else if (which == 'w') { s.add('0', '9'); s.add('A', 'Z'); s.add('a', 'z'); s.add('_', '_'); }
        else {
            s.add(9, 13); s.add(' ', ' ');
            if (unicode)
                for (auto r : std::initializer_list<Range>{{0x85, 0x85}, {0xA0, 0xA0}, {0x1680, 0x1680}, {0x2000, 0x200A}, {0x2028, 0x2029}, {0x202F, 0x202F}, {0x205F, 0x205F}, {0x3000, 0x3000}})
                    s.add(r.first, r.second);
        }
        s.canon();
    }
    static void case_fold(ScalarSet& s, bool unicode) {
        ScalarSet extra;
        for (uint32_t c = 'a'; c <= 'z'; c++) {
            bool lo = s.has(c), up = s.has(c - 32);
            if (lo) extra.add(c - 32, c - 32);
            if (up) extra.add(c, c);
        }
        if (unicode) {   // the only non-ASCII scalars whose simple fold reaches ASCII
            if (s.has('k') || s.has('K')) extra.add(0x212A, 0x212A);
            if (s.has('s') || s.has('S')) extra.add(0x17F, 0x17F);
            if (s.has(0x212A)) { extra.add('k', 'k'); extra.add('K', 'K'); }
            if (s.has(0x17F)) { extra.add('s', 's'); extra.add('S', 'S'); }
        }
        s.unite(extra);
    }
    static int hexval(int c) {
        if (c >= '0' && c <= '9') return c - '0';
        if (c >= 'a' && c <= 'f') return c - 'a' + 10;
        if (c >= 'A' && c <= 'F') return c - 'A' + 10;
        return -1;
    }
    uint32_t hex_escape(int kind) {
        uint32_t cp = 0;
        int nd = 0;
        if (eat('{')) {
            while (peek() != '}') {
                int h = hexval(peek());
                if (h < 0 || ++nd > 8) throw ParseError{"invalid hexadecimal escape"};
                cp = cp * 16 + h;
                pos_++;
            }
            pos_++;
            if (!nd) throw ParseError{"empty hexadecimal escape"};
        } else {
            int want = kind == 'x' ? 2 : kind == 'u' ? 4 : 8;
            for (; nd < want; nd++) {
                int h = hexval(peek());
                if (h < 0) throw ParseError{"invalid hexadecimal escape"};
                cp = cp * 16 + h;
                pos_++;
            }
        }
        if (cp > kMaxScalar || (cp >= 0xD800 && cp <= 0xDFFF)) throw ParseError{"hexadecimal escape is not a Unicode scalar value"};
        return cp;
    }
    static bool is_meta(int c) { return c > 0 && c < 128 && std::strchr("\\.+*?()|[]{}^$#&-~", c) != nullptr; }
    static bool escapable(int c) {
        if (is_meta(c)) return true;
        if (c < 0 || c >= 128) return false;
        if (is_word_byte(c)) return false;
        if (c == '<' || c == '>') return false;
        return true;
    }
    // escapes that stand for one scalar; -1 if `e` is not such an escape
    int64_t scalar_escape(uint32_t e) {
        switch (e) {
            case 'a': return 7;
            case 'f': return 12;
            case 't': return 9;
            case 'n': return 10;
            case 'r': return 13;
            case 'v': return 11;
            case 'x': case 'u': case 'U': return hex_escape((int)e);
        }
        return escapable((int)e) ? (int64_t)e : -1;
    }

    // ---- character classes
    struct ClassItem { bool is_set; uint32_t scalar; ScalarSet set; };
    ClassItem class_item() {
        ClassItem it{false, 0, {}};
        uint32_t c = next_scalar();
        if (c != '\\') { it.scalar = c; return it; }
        uint32_t e = next_scalar();
        if (e == 'd' || e == 's' || e == 'w' || e == 'D' || e == 'S' || e == 'W') {
            perl_class(it.set, (int)(e | 32), f_.u);
            if (e < 'a') it.set.negate(top());
            it.is_set = true;
            return it;
        }
        if (e == 'p' || e == 'P') throw ParseError{"Unicode property classes (\\p) are not supported"};
        if (e == 'b') { it.scalar = 8; return it; }
        int64_t v = scalar_escape(e);
        if (v < 0) throw ParseError{"unrecognized escape sequence"};
        it.scalar = (uint32_t)v;
        return it;
    }
    bool posix_class(ScalarSet& into) {
        static const struct { const char* name; std::vector<Range> r; } kPosix[] = {
            {"alnum", {{'0', '9'}, {'A', 'Z'}, {'a', 'z'}}}, {"alpha", {{'A', 'Z'}, {'a', 'z'}}}, {"ascii", {{0, 0x7F}}},
            {"blank", {{'\t', '\t'}, {' ', ' '}}}, {"cntrl", {{0, 0x1F}, {0x7F, 0x7F}}}, {"digit", {{'0', '9'}}}, {"graph", {{'!', '~'}}},
            {"lower", {{'a', 'z'}}}, {"print", {{' ', '~'}}}, {"punct", {{'!', '/'}, {':', '@'}, {'[', '`'}, {'{', '~'}}},
            {"space", {{'\t', '\r'}, {' ', ' '}}}, {"upper", {{'A', 'Z'}}}, {"word", {{'0', '9'}, {'A', 'Z'}, {'a', 'z'}, {'_', '_'}}},
            {"xdigit", {{'0', '9'}, {'A', 'F'}, {'a', 'f'}}}};
        // at "[:"
        size_t q = pos_ + 2;
        bool neg = false;
        if (q < p_.size() && p_[q] == '^') { neg = true; q++; }
        size_t e = q;
        while (e < p_.size() && p_[e] != ':') e++;
        if (e + 1 >= p_.size() || p_[e + 1] != ']') return false;
        std::string name = p_.substr(q, e - q);
        for (auto& pc : kPosix)
            if (name == pc.name) {
                ScalarSet t;
                t.r = pc.r;
                t.canon();
                if (neg) t.negate(top());
                into.unite(t);
                pos_ = e + 2;
                return true;
            }
        return false;
    }
    static void apply_op(ScalarSet& acc, int op, const ScalarSet& rhs) {
        if (op == '&') acc.intersect(rhs);
        else if (op == '-') acc.subtract(rhs);
        else acc.symdiff(rhs);
    }
    // called just after '['
    ScalarSet char_class() {
        if (++depth_ > 200) throw ParseError{"pattern nesting is too deep"};
        bool neg = eat('^');
        ScalarSet cur, acc;
        bool have_acc = false, first = true;
        int pending = 0;
        for (;;) {
            skip_ws();
            int c = peek();
            if (c < 0) throw ParseError{"unclosed character class"};
            if (c == ']' && !first) { pos_++; break; }
            first = false;
            if (c == '[') {
                if (peek(1) == ':' && posix_class(cur)) continue;
                pos_++;
                cur.unite(char_class());
                continue;
            }
            if ((c == '&' || c == '-' || c == '~') && peek(1) == c) {
                if (f_.i) case_fold(cur, f_.u);
                if (!have_acc) { acc = cur; have_acc = true; } else apply_op(acc, pending, cur);
                cur = ScalarSet{};
                pending = c;
                pos_ += 2;
                continue;
            }
            ClassItem lo = class_item();
            if (lo.is_set) { cur.unite(lo.set); continue; }
            if (peek() == '-' && peek(1) >= 0 && peek(1) != ']' && peek(1) != '-') {
                pos_++;
                ClassItem hi = class_item();
                if (hi.is_set) throw ParseError{"invalid character class range"};
                if (hi.scalar < lo.scalar) throw ParseError{"invalid character class range, start is greater than end"};
                cur.add(lo.scalar, hi.scalar);
            } else cur.add(lo.scalar, lo.scalar);
            cur.canon();
        }
        if (f_.i) case_fold(cur, f_.u);
        if (have_acc) { apply_op(acc, pending, cur); cur = acc; }
        if (neg) cur.negate(top());
        depth_--;
        return cur;
    }

    AstP literal(uint32_t cp) {
        ScalarSet s;
        s.add(cp, cp);
        if (f_.i) case_fold(s, f_.u);
        return scalar_set_ast(s, true);   // ASCII -> one byte, otherwise its UTF-8 encoding
    }
    static AstP assertion(Look l) { auto a = mk(Ast::Assert); a->look = l; return a; }

    // after "(?", positioned on the first flag char.  true: "(?flags:" (a group body follows)
    bool flag_run() {
        bool neg = false, any = false;
        for (;;) {
            int c = peek();
            if (c == ')') { pos_++; if (neg && !any) throw ParseError{"dangling flag negation operator"}; return false; }
            if (c == ':') { pos_++; return true; }
            if (c == '-') { if (neg) throw ParseError{"flag negation operator repeated"}; neg = true; any = false; pos_++; continue; }
            bool v = !neg;
            switch (c) {
                case 'i': f_.i = v; break;
                case 'm': f_.m = v; break;
                case 's': f_.s = v; break;
                case 'U': break;
                case 'u': f_.u = v; break;
                case 'x': f_.x = v; break;
                case 'R':
                    // CRLF mode changes `.`, and `^` / `$` under (?m) (\r\n as one line terminator); only the former was ever
                    // modelled here, so switching it ON is refused instead of being half honoured
                    if (v) throw ParseError{"the CRLF flag (?R) is not supported"};
                    f_.R = false;
                    break;
                default: throw ParseError{"unrecognized flag"};
            }
            any = true;
            pos_++;
        }
    }

    AstP atom() {
        int c = peek();
        if (c == '(') {
            pos_++;
            if (++depth_ > 200) throw ParseError{"pattern nesting is too deep"};
            Flags saved = f_;
            if (peek() == '?') {
                if (peek(1) == '=' || peek(1) == '!' || (peek(1) == '<' && (peek(2) == '=' || peek(2) == '!')))
                    throw ParseError{"look-around, including look-ahead and look-behind, is not supported"};
                pos_++;
                if (peek() == 'P' && peek(1) == '<') pos_++;
                if (peek() == '<') {
                    pos_++;
                    size_t s = pos_;
                    while (pos_ < p_.size() && p_[pos_] != '>') {
                        int ch = (uint8_t)p_[pos_];
                        if (!(is_word_byte(ch) || ch == '.' || ch == '[' || ch == ']' || ch >= 0x80)) throw ParseError{"invalid capture group character"};
                        pos_++;
                    }
                    if (pos_ >= p_.size() || pos_ == s || (p_[s] >= '0' && p_[s] <= '9')) throw ParseError{"invalid capture group name"};
                    pos_++;
                } else {
                    if (peek() == ')') throw ParseError{"missing flags"};
                    if (!flag_run()) { depth_--; return mk(Ast::Empty); }   // "(?flags)" applies to the rest of the enclosing group
                }
            }
            AstP body = alternation();
            if (!eat(')')) throw ParseError{"unclosed group"};
            f_ = saved;
            depth_--;
            return body;
        }
        if (c == '[') { pos_++; return scalar_set_ast(char_class(), f_.u); }
        if (c == '.') {
            pos_++;
            ScalarSet s;
            if (!f_.s) { s.add('\n', '\n'); if (f_.R) s.add('\r', '\r'); }
            s.negate(top());
            return scalar_set_ast(s, f_.u);
        }
        if (c == '^') { pos_++; return assertion(f_.m ? Look::StartLine : Look::StartText); }
        if (c == '$') { pos_++; return assertion(f_.m ? Look::EndLine : Look::EndText); }
        if (c == '\\') {
            pos_++;
            uint32_t e = next_scalar();
            if (e >= '0' && e <= '9') throw ParseError{"backreferences are not supported"};
            switch (e) {
                case 'A': return assertion(Look::StartText);
                case 'z': return assertion(Look::EndText);
                case 'B': return assertion(Look::NotWordB);
                case '<': return assertion(Look::WordStart);
                case '>': return assertion(Look::WordEnd);
                case 'b':
                    if (peek() == '{') {
                        static const std::pair<const char*, Look> kSpecial[] = {{"{start}", Look::WordStart}, {"{end}", Look::WordEnd}, {"{start-half}", Look::WordStartHalf}, {"{end-half}", Look::WordEndHalf}};
                        for (auto& sp : kSpecial) {
                            size_t l = std::strlen(sp.first);
                            if (p_.compare(pos_, l, sp.first) == 0) { pos_ += l; return assertion(sp.second); }
                        }
                        if (peek(1) >= '0' && peek(1) <= '9') return assertion(Look::WordB);   // \b{2}: counted repetition of \b
                        throw ParseError{"invalid word boundary assertion"};
                    }
                    return assertion(Look::WordB);
                case 'd': case 's': case 'w': case 'D': case 'S': case 'W': {
                    ScalarSet s;
                    perl_class(s, (int)(e | 32), f_.u);
                    if (e < 'a') s.negate(top());
                    return scalar_set_ast(s, f_.u);
                }
                case 'p': case 'P': throw ParseError{"Unicode property classes (\\p) are not supported"};
            }
            int64_t v = scalar_escape(e);
            if (v < 0) throw ParseError{"unrecognized escape sequence"};
            if (!f_.u && e == 'x' && v < 256) {   // (?-u)\xFF is the raw byte
                ScalarSet s;
                s.add((uint32_t)v, (uint32_t)v);
                if (f_.i) case_fold(s, false);
                return scalar_set_ast(s, false);
            }
            return literal((uint32_t)v);
        }
        return literal(next_scalar());
    }

    bool number(long& v) {
        int d = 0;
        v = 0;
        while (peek() >= '0' && peek() <= '9') {
            v = v * 10 + (peek() - '0');
            if (v > 100000) throw ParseError{"repetition count is too large"};
            pos_++;
            d++;
        }
        return d > 0;
    }
    AstP repeat() {
        skip_ws();
        int c = peek();
        if (c == '*' || c == '+' || c == '?' || c == '{') throw ParseError{"repetition operator missing expression"};
        AstP x = atom();
        for (;;) {
            skip_ws();
            c = peek();
            int mn, mx;
            if (c == '*') { mn = 0; mx = -1; pos_++; }
            else if (c == '+') { mn = 1; mx = -1; pos_++; }
            else if (c == '?') { mn = 0; mx = 1; pos_++; }
            else if (c == '{') {
                pos_++;
                skip_ws();
                long v;
                if (!number(v)) throw ParseError{"repetition quantifier expects a valid decimal"};
                mn = mx = (int)v;
                skip_ws();
                if (eat(',')) { skip_ws(); mx = number(v) ? (int)v : -1; skip_ws(); }
                if (!eat('}')) throw ParseError{"unclosed counted repetition"};
                if (mx >= 0 && mx < mn) throw ParseError{"invalid repetition count range"};
            } else break;
            eat('?');   // lazy marker: same language
            auto r = mk(Ast::Rep);
            r->min = mn;
            r->max = mx;
            r->kids.push_back(std::move(x));
            x = std::move(r);
        }
        return x;
    }
    AstP concatenation() {
        AstP x;
        for (;;) {
            skip_ws();
            int c = peek();
            if (c < 0 || c == '|' || c == ')') break;
            x = mk_cat2(std::move(x), repeat());
        }
        return x ? std::move(x) : mk(Ast::Empty);
    }
    AstP alternation() {
        AstP x = concatenation();
        while (peek() == '|') { pos_++; x = mk_alt2(std::move(x), concatenation()); }
        return x;
    }
};

// L(revcomp_ast(a)) = { rc(w) : w in L(a) }, assertions mirrored, so that  rc(seq) ~ a  <=>  seq ~ revcomp_ast(a).
AstP revcomp_ast(const Ast& a) {
    const uint8_t* comp = complement_table();
    auto o = mk(a.kind);
    o->min = a.min;
    o->max = a.max;
    switch (a.kind) {
        case Ast::Bytes:
            for (int b = 0; b < 256; b++) if (bs_has(a.set, b)) bs_set(o->set, comp[b]);
            break;
        case Ast::Assert:
            switch (a.look) {
                case Look::StartText: o->look = Look::EndText; break;
                case Look::EndText: o->look = Look::StartText; break;
                case Look::StartLine: o->look = Look::EndLine; break;
                case Look::EndLine: o->look = Look::StartLine; break;
                case Look::WordStart: o->look = Look::WordEnd; break;
                case Look::WordEnd: o->look = Look::WordStart; break;
                case Look::WordStartHalf: o->look = Look::WordEndHalf; break;
                case Look::WordEndHalf: o->look = Look::WordStartHalf; break;
                default: o->look = a.look;
            }
            break;
        case Ast::Cat:
            for (auto it = a.kids.rbegin(); it != a.kids.rend(); ++it) o->kids.push_back(revcomp_ast(**it));
            break;
        default:
            for (auto& k : a.kids) o->kids.push_back(revcomp_ast(*k));
    }
    return o;
}

// ---------------------------------------------------------------- Thompson NFA
struct NState {
    enum T : uint8_t { Bytes, Split, Assert, Match } t;
    Look look;
    uint32_t set_id;      // Bytes: index into Nfa::sets
    int32_t out = -1, out2 = -1;
};
struct Nfa {
    std::vector<NState> st;
    std::vector<ByteSet> sets;
    std::unordered_map<std::string, uint32_t> set_index;
    bool has_look = false;
    uint32_t intern(const ByteSet& s) {
        std::string key((const char*)s.data(), sizeof(ByteSet));
        auto it = set_index.find(key);
        if (it != set_index.end()) return it->second;
        sets.push_back(s);
        return set_index[key] = (uint32_t)sets.size() - 1;
    }
    int32_t add(NState s) {
        if (st.size() >= kMaxNfaStates) throw ParseError{"compiled regex exceeds size limit"};
        st.push_back(s);
        return (int32_t)st.size() - 1;
    }
    // returns the entry state of `a` continuing into `next`
    int32_t build(const Ast& a, int32_t next) {
        switch (a.kind) {
            case Ast::Empty: return next;
            case Ast::Never: { NState s{NState::Split, Look::StartText, 0, -1, -1}; return add(s); }   // dead end
            case Ast::Bytes: { NState s{NState::Bytes, Look::StartText, intern(a.set), next, -1}; return add(s); }
            case Ast::Assert: { has_look = true; NState s{NState::Assert, a.look, 0, next, -1}; return add(s); }
            case Ast::Cat: { int32_t n = next; for (auto it = a.kids.rbegin(); it != a.kids.rend(); ++it) n = build(**it, n); return n; }
            case Ast::Alt: {
                int32_t entry = -1;
                for (auto it = a.kids.rbegin(); it != a.kids.rend(); ++it) {
                    int32_t k = build(**it, next);
                    if (entry < 0) entry = k;
                    else { NState s{NState::Split, Look::StartText, 0, k, entry}; entry = add(s); }
                }
                return entry;
            }
            case Ast::Rep: {
                const Ast& body = *a.kids[0];
                int32_t n = next;
                if (a.max < 0) {
                    NState loop{NState::Split, Look::StartText, 0, -1, next};
                    int32_t l = add(loop);
                    st[l].out = build(body, l);
                    n = l;
                } else {
                    for (int k = 0; k < a.max - a.min; k++) {
                        int32_t b = build(body, n);
                        NState opt{NState::Split, Look::StartText, 0, b, next};
                        n = add(opt);
                    }
                }
                for (int k = 0; k < a.min; k++) n = build(body, n);
                return n;
            }
        }
        return next;
    }
};

// ---------------------------------------------------------------- subset construction with 1-byte look-around
enum Ctx : uint8_t { kStart = 0, kLf = 1, kWord = 2, kOther = 3 };
constexpr int kEoi = -1;

struct Determinizer {
    const Nfa& nfa;
    int32_t nfa_start;
    std::vector<uint32_t> mark;
    uint32_t gen = 0;
    std::vector<int32_t> stack;

    Determinizer(const Nfa& n, int32_t s) : nfa(n), nfa_start(s), mark(n.st.size(), 0) {}

    static bool look_holds(Look l, Ctx prev, int next) {
        bool pw = prev == kWord, nw = next >= 0 && is_word_byte(next);
        switch (l) {
            case Look::StartText: return prev == kStart;
            case Look::EndText: return next == kEoi;
            case Look::StartLine: return prev == kStart || prev == kLf;
            case Look::EndLine: return next == kEoi || next == '\n';
            case Look::WordB: return pw != nw;
            case Look::NotWordB: return pw == nw;
            case Look::WordStart: return !pw && nw;
            case Look::WordEnd: return pw && !nw;
            case Look::WordStartHalf: return !pw;
            case Look::WordEndHalf: return !nw;
        }
        return false;
    }
    // epsilon closure of `seeds` between a byte of category `prev` and byte `next` (or EOI);
    // appends byte-consuming states to `out`; returns true if Match is reachable.
    bool closure(const std::vector<uint32_t>& seeds, Ctx prev, int next, std::vector<uint32_t>& out) {
        gen++;
        out.clear();
        bool matched = false;
        stack.clear();
        for (auto it = seeds.rbegin(); it != seeds.rend(); ++it) stack.push_back((int32_t)*it);
        while (!stack.empty()) {
            int32_t s = stack.back();
            stack.pop_back();
            if (s < 0 || mark[s] == gen) continue;
            mark[s] = gen;
            const NState& n = nfa.st[s];
            switch (n.t) {
                case NState::Match: matched = true; break;
                case NState::Bytes: out.push_back((uint32_t)s); break;
                case NState::Split: stack.push_back(n.out2); stack.push_back(n.out); break;
                case NState::Assert: if (look_holds(n.look, prev, next)) stack.push_back(n.out); break;
            }
        }
        return matched;
    }
};

struct VecHash {
    size_t operator()(const std::vector<uint32_t>& v) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (uint32_t x : v) { h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};

void compute_byte_classes(const Nfa& nfa, uint8_t cls[256], uint32_t* n_classes) {
    std::vector<uint32_t> c(256, 0);
    uint32_t n = 1;
    auto split = [&](const ByteSet& s) {
        std::unordered_map<uint64_t, uint32_t> remap;
        uint32_t nn = 0;
        std::vector<uint32_t> nc(256);
        for (int b = 0; b < 256; b++) {
            uint64_t key = ((uint64_t)c[b] << 1) | (bs_has(s, b) ? 1 : 0);
            auto it = remap.find(key);
            if (it == remap.end()) it = remap.emplace(key, nn++).first;
            nc[b] = it->second;
        }
        c.swap(nc);
        n = nn;
    };
    for (auto& s : nfa.sets) split(s);
    if (nfa.has_look) {
        ByteSet w{}, lf{};
        for (int b = 0; b < 256; b++) if (is_word_byte(b)) bs_set(w, b);
        bs_set(lf, '\n');
        split(w);
        split(lf);
    }
    for (int b = 0; b < 256; b++) cls[b] = (uint8_t)c[b];
    *n_classes = n;
}

}  // namespace

// Moore partition refinement + renumbering (accepting states last, absorbing match state very last).
// `accept` = accept-at-end-of-input per state.  Shared with the Aho-Corasick builder.
void finalize_dfa(uint32_t n_states, uint32_t n_classes, std::vector<uint32_t>& next, std::vector<uint8_t>& accept, uint32_t start,
                  bool minimize, Dfa* out) {
    std::vector<uint32_t> part(n_states);
    uint32_t n_parts = 0;
    if (minimize) {
        // absorbing accepting states form their own initial block so "matched" stays distinguishable
        for (uint32_t s = 0; s < n_states; s++) part[s] = accept[s] ? 1 : 0;
        n_parts = 2;
        std::vector<uint32_t> sig(n_classes + 1), np(n_states);
        for (;;) {
            std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> ids;
            ids.reserve(n_parts * 2);
            for (uint32_t s = 0; s < n_states; s++) {
                sig[0] = part[s];
                for (uint32_t c = 0; c < n_classes; c++) sig[c + 1] = part[next[(size_t)s * n_classes + c]];
                auto it = ids.find(sig);
                if (it == ids.end()) it = ids.emplace(sig, (uint32_t)ids.size()).first;
                np[s] = it->second;
            }
            uint32_t cnt = (uint32_t)ids.size();
            part.swap(np);
            if (cnt == n_parts) break;
            n_parts = cnt;
        }
    } else {
        for (uint32_t s = 0; s < n_states; s++) part[s] = s;
        n_parts = n_states;
    }
    // representative per block, block-level table
    std::vector<uint32_t> rep(n_parts, UINT32_MAX);
    for (uint32_t s = 0; s < n_states; s++) if (rep[part[s]] == UINT32_MAX) rep[part[s]] = s;
    auto bnext = [&](uint32_t b, uint32_t c) { return part[next[(size_t)rep[b] * n_classes + c]]; };
    // BFS order from the start block
    std::vector<uint32_t> order;
    std::vector<uint8_t> seen(n_parts, 0);
    std::deque<uint32_t> q;
    q.push_back(part[start]);
    seen[part[start]] = 1;
    while (!q.empty()) {
        uint32_t b = q.front();
        q.pop_front();
        order.push_back(b);
        for (uint32_t c = 0; c < n_classes; c++) {
            uint32_t t = bnext(b, c);
            if (!seen[t]) { seen[t] = 1; q.push_back(t); }
        }
    }
    auto absorbing = [&](uint32_t b) {
        if (!accept[rep[b]]) return false;
        for (uint32_t c = 0; c < n_classes; c++) if (bnext(b, c) != b) return false;
        return true;
    };
    std::vector<uint32_t> newid(n_parts, UINT32_MAX);
    uint32_t k = 0;
    for (uint32_t b : order) if (!accept[rep[b]]) newid[b] = k++;
    out->accept_from = k;
    for (uint32_t b : order) if (accept[rep[b]] && !absorbing(b)) newid[b] = k++;
    for (uint32_t b : order) if (accept[rep[b]] && absorbing(b)) newid[b] = k++;
    out->n_states = k;
    out->n_classes = n_classes;
    out->start = newid[part[start]];
    out->next.assign((size_t)k * n_classes, 0);
    for (uint32_t b : order)
        for (uint32_t c = 0; c < n_classes; c++) out->next[(size_t)newid[b] * n_classes + c] = newid[bnext(b, c)];
    out->always = out->start >= out->accept_from && absorbing(part[start]);
}

namespace {
bool determinize(Nfa& nfa, int32_t nfa_start, Dfa* out, std::string* err);

// NFA for the union of the ASTs (plus their reverse-complement languages), then the shared determiniser.
bool compile_asts(const std::vector<AstP>& asts, bool with_revcomp, Dfa* out, std::string* err) {
    Nfa nfa;
    int32_t nfa_start = -1;
    try {
        NState m{NState::Match, Look::StartText, 0, -1, -1};
        int32_t match = nfa.add(m);
        for (auto& ast : asts) {
            int32_t e1 = nfa.build(*ast, match);
            if (nfa_start < 0) nfa_start = e1;
            else { NState s{NState::Split, Look::StartText, 0, nfa_start, e1}; nfa_start = nfa.add(s); }
            if (with_revcomp) {
                AstP rc = revcomp_ast(*ast);
                int32_t e2 = nfa.build(*rc, match);
                NState s{NState::Split, Look::StartText, 0, nfa_start, e2};
                nfa_start = nfa.add(s);
            }
        }
        if (nfa_start < 0) {   // empty set never matches (RegexSet::empty().is_match == false)
            NState s{NState::Split, Look::StartText, 0, -1, -1};
            nfa_start = nfa.add(s);
        }
    } catch (ParseError& e) {
        *err = "Failed to build regex set from patterns (" + e.msg + ")";
        return false;
    }
    return determinize(nfa, nfa_start, out, err);
}
}  // namespace

bool compile_regex_set(const std::vector<std::string>& patterns, bool with_revcomp, Dfa* out, std::string* err) {
    std::vector<AstP> asts;
    for (auto& p : patterns) {
        try {
            asts.push_back(Parser(p).parse());
        } catch (ParseError& e) {
            *err = "Invalid regular expression: '" + p + "' (" + e.msg + ")";
            return false;
        }
    }
    return compile_asts(asts, with_revcomp, out, err);
}

// True when the regex is nothing but a literal byte string (what the regex crate's literal optimisations also detect):
// such sets are routed to the Aho-Corasick / q-gram machinery, which is the same language, compiled faster and --
// for large pattern files given without -F -- eligible for the shared-memory filter tier.
bool regex_as_literal(const std::string& pattern, std::string* literal) {
    AstP ast;
    try {
        ast = Parser(pattern).parse();
    } catch (ParseError&) {
        return false;
    }
    literal->clear();
    auto one_byte = [&](const Ast& a) -> bool {
        if (a.kind != Ast::Bytes) return false;
        int found = -1;
        for (int b = 0; b < 256; b++)
            if (bs_has(a.set, b)) { if (found >= 0) return false; found = b; }
        if (found < 0) return false;
        literal->push_back((char)found);
        return true;
    };
    if (ast->kind == Ast::Empty) return true;
    if (ast->kind == Ast::Bytes) return one_byte(*ast);
    if (ast->kind != Ast::Cat) return false;
    for (auto& k : ast->kids) if (!one_byte(*k)) return false;
    return true;
}

// The finite language of a regex, when it has one and it is small: classes of a few bytes, alternation, groups and
// bounded repetition are multiplied out ("AC[GT]{2}(A|G)" -> 8 strings).  An unanchored search for such a regex is a
// search for any of its strings, so large sets of them can take the q-gram filter like literal sets do.  False for
// anything else (assertions, unbounded repetition, big classes, more than `max_strings` strings, the empty string).
static bool finite_language(const Ast& a, size_t max_strings, std::vector<std::string>* out) {
    out->clear();
    switch (a.kind) {
        case Ast::Empty: out->push_back(std::string()); return true;
        case Ast::Bytes: {
            for (int b = 0; b < 256; b++)
                if (bs_has(a.set, b)) { if (out->size() >= 8 || out->size() >= max_strings) return false; out->push_back(std::string(1, (char)b)); }
            return !out->empty();
        }
        case Ast::Cat: {
            out->push_back(std::string());
            std::vector<std::string> kid, next;
            for (auto& k : a.kids) {
                if (!finite_language(*k, max_strings, &kid)) return false;
                if (out->size() * kid.size() > max_strings) return false;
                next.clear();
                for (auto& x : *out) for (auto& y : kid) next.push_back(x + y);
                out->swap(next);
            }
            return true;
        }
        case Ast::Alt: {
            std::vector<std::string> kid;
            for (auto& k : a.kids) {
                if (!finite_language(*k, max_strings, &kid)) return false;
                if (out->size() + kid.size() > max_strings) return false;
                out->insert(out->end(), kid.begin(), kid.end());
            }
            return !out->empty();
        }
        case Ast::Rep: {
            if (a.max < 0 || a.kids.size() != 1) return false;
            std::vector<std::string> kid, power(1, std::string()), next;
            if (!finite_language(*a.kids[0], max_strings, &kid)) return false;
            for (int k = 0; k <= a.max; k++) {
                if (k >= a.min) {
                    if (out->size() + power.size() > max_strings) return false;
                    out->insert(out->end(), power.begin(), power.end());
                }
                if (k == a.max) break;
                if (power.size() * kid.size() > max_strings) return false;
                next.clear();
                for (auto& x : power) for (auto& y : kid) next.push_back(x + y);
                power.swap(next);
            }
            return !out->empty();
        }
        default: return false;
    }
}
bool regex_as_literals(const std::string& pattern, size_t max_strings, std::vector<std::string>* literals) {
    AstP ast;
    try {
        ast = Parser(pattern).parse();
    } catch (ParseError&) {
        return false;
    }
    if (!finite_language(*ast, max_strings, literals)) return false;
    for (auto& l : *literals) if (l.empty()) return false;
    return true;
}

// --iupac bit-mask (BitMaskMatcher / BitMaskSetMatcher, matcher.rs:307-487): a needle is a fixed-length string of
// byte classes.  With IUPAC_MASKS m[] (mod.rs:14-53), needle symbol n_i = m[pattern byte]:
//   needle has an ambiguity code (or a non-IUPAC byte): position i accepts every byte b with (n_i & m[b]) == m[b]
//       (matcher.rs:343-348) -- which includes every byte whose mask is 0;
//   needle is pure ACGT(U): the additive rolling hash (matcher.rs:320-342) only lets windows through whose masks sum to
//       the needle's, so position i accepts exactly the bytes with m[b] == n_i.
// The empty needle matches every read, like the loop at matcher.rs:330 does for offset 0.
bool compile_bitmask_set(const std::vector<std::string>& patterns, bool with_revcomp, Dfa* out, std::string* err) {
    const uint8_t* m = iupac_masks();
    std::vector<AstP> asts;
    for (auto& p : patterns) {
        bool has_iupac = false;
        for (unsigned char ch : p) {
            const uint8_t v = m[ch];
            if (!(v == 1 || v == 2 || v == 4 || v == 8)) has_iupac = true;
        }
        AstP cat;
        for (unsigned char ch : p) {
            auto cls = mk(Ast::Bytes);
            bool any = false;
            for (int b = 0; b < 256; b++) {
                const bool ok = has_iupac ? ((m[ch] & m[b]) == m[b]) : (m[b] == m[ch]);
                if (ok) { bs_set(cls->set, b); any = true; }
            }
            cat = mk_cat2(std::move(cat), any ? std::move(cls) : mk(Ast::Never));
        }
        asts.push_back(cat ? std::move(cat) : mk(Ast::Empty));
    }
    return compile_asts(asts, with_revcomp, out, err);
}

namespace {
bool determinize(Nfa& nfa, int32_t nfa_start, Dfa* out, std::string* err) {
    Dfa d;
    compute_byte_classes(nfa, d.byte_class, &d.n_classes);
    std::vector<int> rep(d.n_classes, -1);
    for (int b = 255; b >= 0; b--) rep[d.byte_class[b]] = b;

    Determinizer det(nfa, nfa_start);
    auto ctx_after = [&](int b) -> Ctx {
        if (!nfa.has_look) return kOther;
        return is_word_byte(b) ? kWord : b == '\n' ? kLf : kOther;
    };
    // DFA state = (kernel set of NFA states *before* closure, look-behind context)
    std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> ids;
    std::vector<std::vector<uint32_t>> sets;
    std::vector<uint8_t> ctxs;
    auto intern = [&](std::vector<uint32_t>& set, Ctx ctx) -> uint32_t {
        std::sort(set.begin(), set.end());
        set.erase(std::unique(set.begin(), set.end()), set.end());
        set.push_back(0x80000000u | ctx);   // ctx folded into the key
        auto it = ids.find(set);
        if (it != ids.end()) return it->second;
        uint32_t id = (uint32_t)sets.size();
        ids.emplace(set, id);
        set.pop_back();
        sets.push_back(set);
        ctxs.push_back(ctx);
        return id;
    };
    constexpr uint32_t kMatch = UINT32_MAX;
    std::vector<uint32_t> next;
    std::vector<uint8_t> accept;
    {
        std::vector<uint32_t> s0{(uint32_t)nfa_start};
        intern(s0, nfa.has_look ? kStart : kOther);
    }
    std::vector<uint32_t> cl, tgt;
    for (uint32_t cur = 0; cur < sets.size(); cur++) {
        if (sets.size() > kMaxDfaStates) { *err = "Failed to build regex set from patterns (DFA exceeds " + std::to_string(kMaxDfaStates) + " states)"; return false; }
        std::vector<uint32_t> seeds = sets[cur];
        Ctx ctx = (Ctx)ctxs[cur];
        next.resize((size_t)(cur + 1) * d.n_classes);
        bool shared = !nfa.has_look, shared_match = false;
        if (shared) shared_match = det.closure(seeds, ctx, 'A', cl);
        for (uint32_t c = 0; c < d.n_classes; c++) {
            int b = rep[c];
            bool matched = shared ? shared_match : det.closure(seeds, ctx, b, cl);
            if (matched) { next[(size_t)cur * d.n_classes + c] = kMatch; continue; }
            tgt.clear();
            for (uint32_t s : cl) if (bs_has(nfa.sets[nfa.st[s].set_id], b)) tgt.push_back((uint32_t)nfa.st[s].out);
            tgt.push_back((uint32_t)nfa_start);   // unanchored search: a match may begin at the next offset
            next[(size_t)cur * d.n_classes + c] = intern(tgt, ctx_after(b));
        }
        accept.push_back(det.closure(seeds, ctx, kEoi, cl) ? 1 : 0);
    }
    uint32_t n = (uint32_t)sets.size();
    // materialise the absorbing match state
    uint32_t match_id = n;
    bool used = false;
    for (auto& t : next) if (t == kMatch) { t = match_id; used = true; }
    if (used) {
        n++;
        next.resize((size_t)n * d.n_classes, match_id);
        accept.push_back(1);
    }
    finalize_dfa(n, d.n_classes, next, accept, 0, true, &d);
    *out = std::move(d);
    return true;
}
}  // namespace

const uint8_t* iupac_masks() {
    static uint8_t t[256];
    static bool init = false;
    if (!init) {
        const struct { char c; uint8_t m; } kv[] = {{'A', 1}, {'C', 2}, {'G', 4}, {'T', 8}, {'U', 8}, {'M', 3}, {'R', 5}, {'W', 9}, {'S', 6}, {'Y', 10},
                                                    {'K', 12}, {'V', 7}, {'H', 11}, {'D', 13}, {'B', 14}, {'N', 15}};
        for (auto& e : kv) { t[(uint8_t)e.c] = e.m; t[(uint8_t)e.c + 32] = e.m; }
        init = true;
    }
    return t;
}

// expand_iupac_fixed_pattern (mod.rs:70-97): depth-first over A,C,G,T, at most 10 000 results.
bool expand_iupac_fixed(const std::string& pattern, std::vector<std::string>* out, std::string* err) {
    const uint8_t* m = iupac_masks();
    out->clear();
    std::string prefix;
    bool failed = false;
    auto rec = [&](auto&& self, size_t i) -> void {
        if (failed) return;
        if (out->size() >= 10000) {
            *err = "IUPAC expansion exceeded 10000 patterns for input '" + pattern + "'";
            failed = true;
            return;
        }
        if (i == pattern.size()) { out->push_back(prefix); return; }
        for (char base : {'A', 'C', 'G', 'T'})
            if (m[(uint8_t)base] & m[(uint8_t)pattern[i]]) {
                prefix.push_back(base);
                self(self, i + 1);
                prefix.pop_back();
            }
    };
    rec(rec, 0);
    return !failed;
}

// expand_iupac_regex (mod.rs:101-118)
std::string iupac_to_regex(const std::string& pattern) {
    const uint8_t* m = iupac_masks();
    std::string o;
    for (unsigned char ch : pattern) {
        std::string bases;
        for (char base : {'A', 'C', 'G', 'T'}) if (m[(uint8_t)base] & m[ch]) bases.push_back(base);
        if (bases.size() == 1) o += bases;
        else { o.push_back('['); o += bases; o.push_back(']'); }
    }
    return o;
}

std::string validate_fixed_pattern(const std::string& pattern, bool protein) {
    const char* valid = protein ? "ARNDCEQGHILKMFPSTWYVBJOUXZ" : "ACGTN";
    const char* kind = protein ? "amino acids" : "DNA bases";
    for (size_t i = 0; i < pattern.size();) {
        uint8_t c = (uint8_t)pattern[i];
        size_t len = c < 0x80 ? 1 : c >= 0xF0 ? 4 : c >= 0xE0 ? 3 : 2;
        len = std::min(len, pattern.size() - i);
        if (c >= 0x80 || !std::strchr(valid, c) || c == 0)
            return std::string("Fixed pattern must contain only ") + kind + ": " + pattern.substr(0, i) + " .. [" + pattern.substr(i, len) + "] .. " + pattern.substr(i + len);
        i += len;
    }
    return std::string();
}

}  // namespace fqg
