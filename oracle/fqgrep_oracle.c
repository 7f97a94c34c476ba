/*
 * fqgrep_oracle.c -- CPU restatement of fqgrep's read-matching hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product
 * (fqgrep_b200/) never links, imports or calls anything in this directory.
 *
 * Parity status: the reference is Rust and cannot be compiled in this image (no cargo /
 * rustc / vendored crates), so there is no oracle/_ref.  This restatement is pinned
 * against every golden vector the reference's own tests hold for the path
 * (tests/golden/kat_reference.json, transcribed from /root/reference/src/lib/matcher.rs
 * :786-990,1285-1373 and src/main.rs:926-1402) and cross-checked against Python `re`
 * and GNU grep (tests/test_oracle_*.py).  Beyond those vectors parity is "unpinned":
 * the arithmetic lives in third-party crates that are not in /root/reference
 * (regex 1.12.3 / regex-automata 0.4.14 / regex-syntax 0.8.10, aho-corasick 1.1.4,
 * bstr 0.2.17 + memchr 2.8.0, ahash 0.8.12, seq_io 0.3.4 -- versions from Cargo.lock);
 * their published semantics are restated here.
 *
 * What each piece follows (file:line under /root/reference):
 *   fqo_revcomp            src/lib/mod.rs:129-155   (COMPLEMENT LUT + reverse)
 *   bases_match()          src/lib/matcher.rs:203-205 (fixed), :249-251 (set),
 *                          :506-508 (regex), :569-571 (regex set)  -- all XOR invert
 *   fqo_read_match         src/lib/matcher.rs:164-175 (truth table), :631-638 (names)
 *   fqo_validate_fixed     src/lib/matcher.rs:128-147
 *   pair decision          src/main.rs:731-756 (id assert, r1 || r2)
 *   count / selection      src/main.rs:558-575, 597-614, 635-657, 667-671
 *   FASTQ record rules     seq_io 0.3.4 [3P] as used at src/main.rs:61-87,
 *                          src/lib/seq_io.rs:60-80,149-198
 *   regex surface          regex-syntax 0.8 [3P] as called at matcher.rs:497-499,551-554
 *
 * The regex engine is a Thompson-NFA (Pike) simulation over bytes: deliberately a
 * different algorithm from the product's ahead-of-time DFA compiler.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ utils */
typedef struct { int lo, hi; } Rng;              /* inclusive codepoint range */
typedef struct { Rng *r; int n, cap; } RSet;      /* sorted, non-adjacent ranges */

static void rs_init(RSet *s) { s->r = NULL; s->n = s->cap = 0; }
static void rs_free(RSet *s) { free(s->r); rs_init(s); }
static void rs_push(RSet *s, int lo, int hi) {
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 8; s->r = realloc(s->r, sizeof(Rng) * s->cap); }
    s->r[s->n].lo = lo; s->r[s->n].hi = hi; s->n++;
}
static int rng_cmp(const void *a, const void *b) { return ((const Rng *)a)->lo - ((const Rng *)b)->lo; }
static void rs_canon(RSet *s) {
    if (!s->n) return;
    qsort(s->r, s->n, sizeof(Rng), rng_cmp);
    int w = 0;
    for (int i = 1; i < s->n; i++) {
        if (s->r[i].lo <= s->r[w].hi + 1) { if (s->r[i].hi > s->r[w].hi) s->r[w].hi = s->r[i].hi; }
        else s->r[++w] = s->r[i];
    }
    s->n = w + 1;
}
static void rs_union(RSet *d, const RSet *a) { for (int i = 0; i < a->n; i++) rs_push(d, a->r[i].lo, a->r[i].hi); rs_canon(d); }
static void rs_negate(RSet *s, int maxcp) {
    RSet o; rs_init(&o); int next = 0;
    for (int i = 0; i < s->n; i++) { if (s->r[i].lo > next) rs_push(&o, next, s->r[i].lo - 1); next = s->r[i].hi + 1; }
    if (next <= maxcp) rs_push(&o, next, maxcp);
    rs_free(s); *s = o;
}
static void rs_intersect(RSet *a, const RSet *b) {
    RSet o; rs_init(&o);
    for (int i = 0; i < a->n; i++) for (int j = 0; j < b->n; j++) {
        int lo = a->r[i].lo > b->r[j].lo ? a->r[i].lo : b->r[j].lo;
        int hi = a->r[i].hi < b->r[j].hi ? a->r[i].hi : b->r[j].hi;
        if (lo <= hi) rs_push(&o, lo, hi);
    }
    rs_canon(&o); rs_free(a); *a = o;
}
static void rs_copy(RSet *d, const RSet *a) { rs_init(d); for (int i = 0; i < a->n; i++) rs_push(d, a->r[i].lo, a->r[i].hi); }

/* --------------------------------------------------------------- regex AST */
enum { N_EMPTY, N_BYTES, N_CAT, N_ALT, N_REP, N_LOOK, N_FAIL };
enum { LK_START_TEXT, LK_END_TEXT, LK_START_LINE, LK_END_LINE, LK_WORD, LK_NOT_WORD,
       LK_WORD_START, LK_WORD_END, LK_WORD_START_HALF, LK_WORD_END_HALF };
typedef struct Node {
    int kind;
    uint8_t set[32];          /* N_BYTES: 256-bit byte set */
    struct Node *a, *b;
    int min, max;             /* N_REP; max = -1 for unbounded */
    int look;
} Node;

typedef struct {
    const uint8_t *p; int n, pos;
    int fi, fm, fs, fu, fx, fR;   /* flags: i m s u x R */
    int depth; int err; char msg[160];
    Node **all; int nall, capall;
} Parser;

static Node *mk(Parser *P, int kind) {
    Node *x = calloc(1, sizeof(Node)); x->kind = kind;
    if (P->nall == P->capall) { P->capall = P->capall ? P->capall * 2 : 64; P->all = realloc(P->all, sizeof(Node *) * P->capall); }
    P->all[P->nall++] = x; return x;
}
static Node *fail_(Parser *P, const char *m) { if (!P->err) { P->err = 1; snprintf(P->msg, sizeof P->msg, "%s", m); } return NULL; }
static Node *mkcat(Parser *P, Node *a, Node *b) { if (!a) return b; if (!b) return a; Node *x = mk(P, N_CAT); x->a = a; x->b = b; return x; }
static Node *mkalt(Parser *P, Node *a, Node *b) { Node *x = mk(P, N_ALT); x->a = a; x->b = b; return x; }
static Node *mkbyte_rng(Parser *P, int lo, int hi) { Node *x = mk(P, N_BYTES); for (int c = lo; c <= hi; c++) x->set[c >> 3] |= 1 << (c & 7); return x; }

/* UTF-8 encode */
static int u8enc(int cp, uint8_t *o) {
    if (cp < 0x80) { o[0] = cp; return 1; }
    if (cp < 0x800) { o[0] = 0xC0 | (cp >> 6); o[1] = 0x80 | (cp & 63); return 2; }
    if (cp < 0x10000) { o[0] = 0xE0 | (cp >> 12); o[1] = 0x80 | ((cp >> 6) & 63); o[2] = 0x80 | (cp & 63); return 3; }
    o[0] = 0xF0 | (cp >> 18); o[1] = 0x80 | ((cp >> 12) & 63); o[2] = 0x80 | ((cp >> 6) & 63); o[3] = 0x80 | (cp & 63); return 4;
}
/* Split a scalar range into byte-range sequences (the utf8-ranges algorithm, restated). */
static Node *u8range(Parser *P, int lo, int hi) {
    if (lo > hi) return NULL;
    if (lo <= 0xDFFF && hi >= 0xD800) {          /* carve out surrogates */
        Node *a = lo < 0xD800 ? u8range(P, lo, 0xD7FF) : NULL, *b = hi > 0xDFFF ? u8range(P, 0xE000, hi) : NULL;
        if (a && b) return mkalt(P, a, b); return a ? a : b;
    }
    static const int lim[3] = { 0x7F, 0x7FF, 0xFFFF };
    for (int i = 0; i < 3; i++) if (lo <= lim[i] && hi > lim[i]) return mkalt(P, u8range(P, lo, lim[i]), u8range(P, lim[i] + 1, hi));
    if (hi < 0x80) return mkbyte_rng(P, lo, hi);
    for (int i = 1; i < 4; i++) {
        int m = (1 << (6 * i)) - 1;
        if ((lo & ~m) != (hi & ~m)) {
            if ((lo & m) != 0) return mkalt(P, u8range(P, lo, lo | m), u8range(P, (lo | m) + 1, hi));
            if ((hi & m) != m) return mkalt(P, u8range(P, lo, (hi & ~m) - 1), u8range(P, hi & ~m, hi));
        }
    }
    uint8_t a[4], b[4]; int n = u8enc(lo, a); u8enc(hi, b);
    Node *x = NULL;
    for (int k = 0; k < n; k++) x = mkcat(P, x, mkbyte_rng(P, a[k], b[k]));
    return x;
}
/* A class (set of scalars, or of bytes when !unicode) -> AST over bytes. */
static Node *class_node(Parser *P, const RSet *s, int unicode) {
    Node *x = NULL;
    if (!unicode) {
        Node *b = mk(P, N_BYTES); int any = 0;
        for (int i = 0; i < s->n; i++) for (int c = s->r[i].lo; c <= s->r[i].hi && c < 256; c++) { b->set[c >> 3] |= 1 << (c & 7); any = 1; }
        return any ? b : mk(P, N_FAIL);
    }
    for (int i = 0; i < s->n; i++) { Node *r = u8range(P, s->r[i].lo, s->r[i].hi); if (r) x = x ? mkalt(P, x, r) : r; }
    return x ? x : mk(P, N_FAIL);
}

static int peek(Parser *P) { return P->pos < P->n ? P->p[P->pos] : -1; }
static int eat(Parser *P, int c) { if (peek(P) == c) { P->pos++; return 1; } return 0; }
/* decode one UTF-8 scalar from the pattern */
static int next_cp(Parser *P) {
    if (P->pos >= P->n) return -1;
    int c = P->p[P->pos++];
    if (c < 0x80) return c;
    int n = c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : 1; int cp = c & (0x3F >> n);
    while (n-- && P->pos < P->n) cp = (cp << 6) | (P->p[P->pos++] & 63);
    return cp;
}
static void skip_ws(Parser *P) {
    if (!P->fx) return;
    for (;;) {
        int c = peek(P);
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') P->pos++;
        else if (c == '#') { while (P->pos < P->n && P->p[P->pos] != '\n') P->pos++; }
        else break;
    }
}
static int isword(int c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_'; }

static void add_perl(RSet *s, int which, int unicode) {   /* d s w */
    if (which == 'd') rs_push(s, '0', '9');
    else if (which == 'w') { rs_push(s, '0', '9'); rs_push(s, 'A', 'Z'); rs_push(s, 'a', 'z'); rs_push(s, '_', '_'); }
    else {
        rs_push(s, '\t', '\r'); rs_push(s, ' ', ' ');
        if (unicode) { rs_push(s, 0x85, 0x85); rs_push(s, 0xA0, 0xA0); rs_push(s, 0x1680, 0x1680); rs_push(s, 0x2000, 0x200A);
            rs_push(s, 0x2028, 0x2029); rs_push(s, 0x202F, 0x202F); rs_push(s, 0x205F, 0x205F); rs_push(s, 0x3000, 0x3000); }
    }
    rs_canon(s);
}
static void fold_case(RSet *s, int unicode) {
    RSet add; rs_init(&add);
    for (int i = 0; i < s->n; i++) {
        for (int c = 'a'; c <= 'z'; c++) if (c >= s->r[i].lo && c <= s->r[i].hi) rs_push(&add, c - 32, c - 32);
        for (int c = 'A'; c <= 'Z'; c++) if (c >= s->r[i].lo && c <= s->r[i].hi) rs_push(&add, c + 32, c + 32);
        if (unicode) {
            if (s->r[i].lo <= 0x212A && 0x212A <= s->r[i].hi) { rs_push(&add, 'k', 'k'); rs_push(&add, 'K', 'K'); }
            if (s->r[i].lo <= 0x17F && 0x17F <= s->r[i].hi) { rs_push(&add, 's', 's'); rs_push(&add, 'S', 'S'); }
        }
    }
    if (unicode) for (int i = 0; i < s->n; i++) {
        if ((s->r[i].lo <= 'k' && 'k' <= s->r[i].hi) || (s->r[i].lo <= 'K' && 'K' <= s->r[i].hi)) rs_push(&add, 0x212A, 0x212A);
        if ((s->r[i].lo <= 's' && 's' <= s->r[i].hi) || (s->r[i].lo <= 'S' && 'S' <= s->r[i].hi)) rs_push(&add, 0x17F, 0x17F);
    }
    rs_union(s, &add); rs_free(&add);
}
static int hexv(int c) { if (c >= '0' && c <= '9') return c - '0'; if (c >= 'a' && c <= 'f') return c - 'a' + 10; if (c >= 'A' && c <= 'F') return c - 'A' + 10; return -1; }
static int parse_hex(Parser *P, int kind) {   /* after \x \u \U */
    int cp = 0, nd = 0;
    if (eat(P, '{')) {
        while (peek(P) != '}' ) { int h = hexv(peek(P)); if (h < 0) { fail_(P, "invalid hex escape"); return -1; } cp = cp * 16 + h; P->pos++; if (++nd > 8) { fail_(P, "invalid hex escape"); return -1; } }
        P->pos++; if (!nd) { fail_(P, "empty hex escape"); return -1; }
    } else {
        int want = kind == 'x' ? 2 : kind == 'u' ? 4 : 8;
        for (; nd < want; nd++) { int h = hexv(peek(P)); if (h < 0) { fail_(P, "invalid hex escape"); return -1; } cp = cp * 16 + h; P->pos++; }
    }
    if (cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) { fail_(P, "invalid scalar value in hex escape"); return -1; }
    return cp;
}
static int is_meta(int c) { return c > 0 && strchr("\\.+*?()|[]{}^$#&-~", c) != NULL; }
static int is_escapeable(int c) {
    if (is_meta(c)) return 1;
    if (c >= 0x80 || c < 0) return 0;
    if (isword(c)) return 0;
    if (c == '<' || c == '>') return 0;
    return 1;
}
/* Escape that denotes a single scalar (shared by class and top level). Returns cp or -1 if not one. */
static int escape_scalar(Parser *P, int c) {
    switch (c) {
        case 'a': return 7; case 'f': return 12; case 't': return 9; case 'n': return 10; case 'r': return 13; case 'v': return 11;
        case 'x': case 'u': case 'U': return parse_hex(P, c);
    }
    if (is_escapeable(c)) return c;
    return -1;
}

static Node *parse_alt(Parser *P);

static const struct { const char *name; const char *spec; } POSIX[] = {
    {"alnum", "09AZaz"}, {"alpha", "AZaz"}, {"ascii", "\x01\x7f"}, {"blank", "\t\t  "}, {"cntrl", "\x01\x1f\x7f\x7f"},
    {"digit", "09"}, {"graph", "!~"}, {"lower", "az"}, {"print", " ~"}, {"punct", "!/:@[`{~"}, {"space", "\t\r  "},
    {"upper", "AZ"}, {"word", "09AZaz__"}, {"xdigit", "09AFaf"}, {NULL, NULL}
};

/* parse one class item set or nested class into *out; handles [..] recursion and set ops */
static void parse_class(Parser *P, RSet *out);   /* called with pos just after '[' */

static int class_atom(Parser *P, RSet *tmp, int *is_set) {
    /* returns scalar, or -2 with *is_set=1 and tmp filled */
    *is_set = 0;
    int c = next_cp(P);
    if (c < 0) { fail_(P, "unclosed character class"); return -1; }
    if (c != '\\') return c;
    int e = next_cp(P);
    if (e < 0) { fail_(P, "incomplete escape"); return -1; }
    if (e == 'd' || e == 's' || e == 'w' || e == 'D' || e == 'S' || e == 'W') {
        add_perl(tmp, e | 32, P->fu);
        if (e < 'a') rs_negate(tmp, P->fu ? 0x10FFFF : 255);
        *is_set = 1; return -2;
    }
    if (e == 'p' || e == 'P') { fail_(P, "Unicode property classes are not supported"); return -1; }
    if (e == 'b') return 8;   /* \b inside a class is backspace */
    int cp = escape_scalar(P, e);
    if (cp < 0) { fail_(P, "unrecognized escape sequence"); return -1; }
    return cp;
}

static void parse_class(Parser *P, RSet *out) {
    if (++P->depth > 200) { fail_(P, "nesting too deep"); return; }
    int negate = eat(P, '^');
    RSet cur; rs_init(&cur);         /* union being accumulated */
    RSet acc; rs_init(&acc); int have_acc = 0, pend_op = 0;   /* for && -- ~~ */
    int first = 1;
    for (;;) {
        skip_ws(P);
        if (P->err) break;
        int c = peek(P);
        if (c < 0) { fail_(P, "unclosed character class"); break; }
        if (c == ']' && !first) { P->pos++; break; }
        first = 0;
        if (c == '[') {
            /* POSIX class or nested class */
            if (P->pos + 1 < P->n && P->p[P->pos + 1] == ':') {
                int q = P->pos + 2, neg = 0; if (q < P->n && P->p[q] == '^') { neg = 1; q++; }
                int e = q; while (e < P->n && P->p[e] != ':') e++;
                if (e + 1 < P->n && P->p[e + 1] == ']') {
                    int found = 0;
                    for (int k = 0; POSIX[k].name; k++) if ((int)strlen(POSIX[k].name) == e - q && !memcmp(POSIX[k].name, P->p + q, e - q)) {
                        RSet t; rs_init(&t); const char *s = POSIX[k].spec;
                        if (k == 2) { rs_push(&t, 0, 0x7f); } else for (; s[0]; s += 2) rs_push(&t, (uint8_t)s[0], (uint8_t)s[1]);
                        rs_canon(&t); if (neg) rs_negate(&t, P->fu ? 0x10FFFF : 255);
                        rs_union(&cur, &t); rs_free(&t); found = 1; break;
                    }
                    if (found) { P->pos = e + 2; continue; }
                }
            }
            P->pos++;
            RSet t; rs_init(&t); parse_class(P, &t); rs_union(&cur, &t); rs_free(&t);
            continue;
        }
        if ((c == '&' || c == '-' || c == '~') && P->pos + 1 < P->n && P->p[P->pos + 1] == c) {
            /* set operation: fold cur into acc */
            if (P->fi) fold_case(&cur, P->fu);
            if (!have_acc) { acc = cur; have_acc = 1; }
            else {
                if (pend_op == '&') rs_intersect(&acc, &cur);
                else if (pend_op == '-') { RSet n; rs_copy(&n, &cur); rs_negate(&n, 0x10FFFF); rs_intersect(&acc, &n); rs_free(&n); }
                else { RSet u, i2; rs_copy(&u, &acc); rs_union(&u, &cur); rs_copy(&i2, &acc); rs_intersect(&i2, &cur); rs_negate(&i2, 0x10FFFF); rs_intersect(&u, &i2); rs_free(&i2); rs_free(&acc); acc = u; }
                rs_free(&cur);
            }
            rs_init(&cur); pend_op = c; P->pos += 2; continue;
        }
        RSet t; rs_init(&t); int is_set;
        int lo = class_atom(P, &t, &is_set);
        if (P->err) { rs_free(&t); break; }
        if (is_set) { rs_union(&cur, &t); rs_free(&t); continue; }
        rs_free(&t);
        /* range? */
        if (peek(P) == '-' && P->pos + 1 < P->n && P->p[P->pos + 1] != ']' && !(P->p[P->pos + 1] == '-')) {
            P->pos++;
            RSet t2; rs_init(&t2); int is2;
            int hi = class_atom(P, &t2, &is2); rs_free(&t2);
            if (P->err) break;
            if (is2) { fail_(P, "invalid range in character class"); break; }
            if (hi < lo) { fail_(P, "invalid character class range"); break; }
            rs_push(&cur, lo, hi); rs_canon(&cur);
        } else { rs_push(&cur, lo, lo); rs_canon(&cur); }
    }
    if (P->fi) fold_case(&cur, P->fu);
    if (have_acc) {
        if (pend_op == '&') rs_intersect(&acc, &cur);
        else if (pend_op == '-') { RSet n; rs_copy(&n, &cur); rs_negate(&n, 0x10FFFF); rs_intersect(&acc, &n); rs_free(&n); }
        else { RSet u, i2; rs_copy(&u, &acc); rs_union(&u, &cur); rs_copy(&i2, &acc); rs_intersect(&i2, &cur); rs_negate(&i2, 0x10FFFF); rs_intersect(&u, &i2); rs_free(&i2); rs_free(&acc); acc = u; }
        rs_free(&cur); cur = acc;
    }
    if (negate) rs_negate(&cur, P->fu ? 0x10FFFF : 255);
    rs_union(out, &cur); rs_free(&cur);
    P->depth--;
}

static Node *lit_node(Parser *P, int cp) {
    /* a literal scalar: ASCII -> one byte; anything else -> its UTF-8 encoding */
    RSet s; rs_init(&s); rs_push(&s, cp, cp);
    if (P->fi) fold_case(&s, P->fu);
    Node *x = class_node(P, &s, 1);
    rs_free(&s); return x;
}
static Node *mklook(Parser *P, int k) { Node *x = mk(P, N_LOOK); x->look = k; return x; }

static int parse_flags(Parser *P) {   /* after "(?" when next is a flag char; returns 1 if "(?flags:" (group follows), 0 if "(?flags)" */
    int neg = 0, any = 0;
    for (;;) {
        int c = peek(P);
        if (c == ')') { P->pos++; if (neg && !any) fail_(P, "dangling flag negation"); return 0; }
        if (c == ':') { P->pos++; return 1; }
        if (c == '-') { if (neg) { fail_(P, "repeated flag negation"); return 0; } neg = 1; any = 0; P->pos++; continue; }
        int v = !neg;
        switch (c) {
            case 'i': P->fi = v; break; case 'm': P->fm = v; break; case 's': P->fs = v; break;
            case 'U': break; case 'u': P->fu = v; break; case 'x': P->fx = v; break; case 'R': P->fR = v; break;
            default: fail_(P, "unrecognized flag"); return 0;
        }
        any = 1; P->pos++;
    }
}

static Node *parse_atom(Parser *P) {
    int c = peek(P);
    if (c == '(') {
        P->pos++;
        if (++P->depth > 200) return fail_(P, "nesting too deep");
        int sv[6] = { P->fi, P->fm, P->fs, P->fu, P->fx, P->fR };
        if (peek(P) == '?') {
            if (P->pos + 1 < P->n && (P->p[P->pos + 1] == '=' || P->p[P->pos + 1] == '!')) return fail_(P, "look-around is not supported");
            if (P->pos + 2 < P->n && P->p[P->pos + 1] == '<' && (P->p[P->pos + 2] == '=' || P->p[P->pos + 2] == '!')) return fail_(P, "look-around is not supported");
            P->pos++;
            if (peek(P) == 'P' && P->pos + 1 < P->n && P->p[P->pos + 1] == '<') P->pos++;
            if (peek(P) == '<') {   /* named group */
                P->pos++; int s = P->pos;
                while (P->pos < P->n && P->p[P->pos] != '>') { int ch = P->p[P->pos]; if (!(isword(ch) || ch == '.' || ch == '[' || ch == ']' || ch >= 0x80)) return fail_(P, "invalid capture group name"); P->pos++; }
                if (P->pos >= P->n || P->pos == s || (P->p[s] >= '0' && P->p[s] <= '9')) return fail_(P, "invalid capture group name");
                P->pos++;
            } else {
                if (peek(P) == ')' ) return fail_(P, "empty flags");
                if (!parse_flags(P)) {
                    if (P->err) return NULL;
                    /* "(?flags)": flags apply to the remainder of the enclosing group */
                    P->depth--;
                    return mk(P, N_EMPTY);
                }
            }
        }
        Node *x = parse_alt(P);
        if (P->err) return NULL;
        if (!eat(P, ')')) return fail_(P, "unclosed group");
        P->fi = sv[0]; P->fm = sv[1]; P->fs = sv[2]; P->fu = sv[3]; P->fx = sv[4]; P->fR = sv[5];
        P->depth--;
        return x;
    }
    if (c == '[') { P->pos++; RSet s; rs_init(&s); parse_class(P, &s); if (P->err) { rs_free(&s); return NULL; } Node *x = class_node(P, &s, P->fu); rs_free(&s); return x; }
    if (c == '.') {
        P->pos++; RSet s; rs_init(&s);
        if (!P->fs) { rs_push(&s, '\n', '\n'); if (P->fR) rs_push(&s, '\r', '\r'); rs_canon(&s); }
        rs_negate(&s, P->fu ? 0x10FFFF : 255);
        Node *x = class_node(P, &s, P->fu); rs_free(&s); return x;
    }
    if (c == '^') { P->pos++; return mklook(P, P->fm ? LK_START_LINE : LK_START_TEXT); }
    if (c == '$') { P->pos++; return mklook(P, P->fm ? LK_END_LINE : LK_END_TEXT); }
    if (c == '\\') {
        P->pos++; int e = next_cp(P);
        if (e < 0) return fail_(P, "incomplete escape sequence");
        if (e >= '0' && e <= '9') return fail_(P, "backreferences are not supported");
        switch (e) {
            case 'A': return mklook(P, LK_START_TEXT);
            case 'z': return mklook(P, LK_END_TEXT);
            case 'B': return mklook(P, LK_NOT_WORD);
            case '<': return mklook(P, LK_WORD_START);
            case '>': return mklook(P, LK_WORD_END);
            case 'b':
                if (peek(P) == '{') {
                    static const struct { const char *s; int k; } W[] = { {"{start}", LK_WORD_START}, {"{end}", LK_WORD_END}, {"{start-half}", LK_WORD_START_HALF}, {"{end-half}", LK_WORD_END_HALF} };
                    for (int k = 0; k < 4; k++) { int l = strlen(W[k].s); if (P->pos + l <= P->n && !memcmp(P->p + P->pos, W[k].s, l)) { P->pos += l; return mklook(P, W[k].k); } }
                    /* "\b{" followed by something else: if it looks like a counted repetition it is \b repeated; else error */
                    int q = P->pos + 1; if (q < P->n && P->p[q] >= '0' && P->p[q] <= '9') return mklook(P, LK_WORD);
                    return fail_(P, "invalid word boundary assertion");
                }
                return mklook(P, LK_WORD);
            case 'd': case 's': case 'w': case 'D': case 'S': case 'W': {
                RSet s; rs_init(&s); add_perl(&s, e | 32, P->fu);
                if (e < 'a') rs_negate(&s, P->fu ? 0x10FFFF : 255);
                Node *x = class_node(P, &s, P->fu); rs_free(&s); return x;
            }
            case 'p': case 'P': return fail_(P, "Unicode property classes are not supported");
        }
        int cp = escape_scalar(P, e);
        if (P->err) return NULL;
        if (cp < 0) return fail_(P, "unrecognized escape sequence");
        if (!P->fu && (e == 'x') && cp < 256) { Node *x = mkbyte_rng(P, cp, cp); if (P->fi && isword(cp) && cp != '_' && !(cp >= '0' && cp <= '9')) { int o = cp ^ 32; x->set[o >> 3] |= 1 << (o & 7); } return x; }
        return lit_node(P, cp);
    }
    int cp = next_cp(P);
    return lit_node(P, cp);
}

static Node *parse_repeat(Parser *P) {
    skip_ws(P);
    int c = peek(P);
    if (c == '*' || c == '+' || c == '?' || c == '{') return fail_(P, "repetition operator missing expression");
    Node *x = parse_atom(P);
    if (P->err) return NULL;
    for (;;) {
        skip_ws(P);
        c = peek(P);
        int mn, mx;
        if (c == '*') { mn = 0; mx = -1; P->pos++; }
        else if (c == '+') { mn = 1; mx = -1; P->pos++; }
        else if (c == '?') { mn = 0; mx = 1; P->pos++; }
        else if (c == '{') {
            P->pos++; skip_ws(P);
            int d = 0; long v = 0;
            while (peek(P) >= '0' && peek(P) <= '9') { v = v * 10 + (peek(P) - '0'); if (v > 100000) return fail_(P, "repetition count too large"); P->pos++; d++; }
            if (!d) return fail_(P, "repetition quantifier expects a valid decimal");
            mn = (int)v; mx = mn; skip_ws(P);
            if (eat(P, ',')) {
                skip_ws(P); d = 0; v = 0;
                while (peek(P) >= '0' && peek(P) <= '9') { v = v * 10 + (peek(P) - '0'); if (v > 100000) return fail_(P, "repetition count too large"); P->pos++; d++; }
                mx = d ? (int)v : -1; skip_ws(P);
            }
            if (!eat(P, '}')) return fail_(P, "unclosed counted repetition");
            if (mx >= 0 && mx < mn) return fail_(P, "invalid repetition count range");
        } else break;
        eat(P, '?');   /* lazy suffix: irrelevant for is_match */
        Node *r = mk(P, N_REP); r->a = x; r->min = mn; r->max = mx; x = r;
    }
    return x;
}
static Node *parse_concat(Parser *P) {
    Node *x = NULL;
    for (;;) {
        skip_ws(P);
        int c = peek(P);
        if (c < 0 || c == '|' || c == ')') break;
        Node *y = parse_repeat(P);
        if (P->err) return NULL;
        x = mkcat(P, x, y);
    }
    return x ? x : mk(P, N_EMPTY);
}
static Node *parse_alt(Parser *P) {
    Node *x = parse_concat(P);
    while (!P->err && peek(P) == '|') { P->pos++; Node *y = parse_concat(P); if (P->err) return NULL; x = mkalt(P, x, y); }
    return x;
}

/* ------------------------------------------------------------- Pike program */
enum { I_BYTES, I_SPLIT, I_JMP, I_LOOK, I_MATCH, I_FAIL };
typedef struct { int op, x, y; uint8_t set[32]; } Inst;
typedef struct { Inst *in; int n, cap; int err; } Prog;
#define PROG_LIMIT 400000
static int emit(Prog *g, int op) {
    if (g->n >= PROG_LIMIT) { g->err = 1; return 0; }
    if (g->n == g->cap) { g->cap = g->cap ? g->cap * 2 : 64; g->in = realloc(g->in, sizeof(Inst) * g->cap); }
    memset(&g->in[g->n], 0, sizeof(Inst)); g->in[g->n].op = op; return g->n++;
}
static void comp(Prog *g, Node *x) {
    if (g->err) return;
    switch (x->kind) {
        case N_EMPTY: break;
        case N_FAIL: emit(g, I_FAIL); break;
        case N_BYTES: { int i = emit(g, I_BYTES); memcpy(g->in[i].set, x->set, 32); break; }
        case N_LOOK: { int i = emit(g, I_LOOK); g->in[i].x = x->look; break; }
        case N_CAT: comp(g, x->a); comp(g, x->b); break;
        case N_ALT: { int s = emit(g, I_SPLIT); g->in[s].x = g->n; comp(g, x->a); int j = emit(g, I_JMP); g->in[s].y = g->n; comp(g, x->b); g->in[j].x = g->n; break; }
        case N_REP: {
            for (int k = 0; k < x->min && !g->err; k++) comp(g, x->a);
            if (x->max < 0) { int s = emit(g, I_SPLIT); g->in[s].x = g->n; comp(g, x->a); int j = emit(g, I_JMP); g->in[j].x = s; g->in[s].y = g->n; }
            else {
                /* (a(a(a)?)?)? : nested optionals */
                int nopt = x->max - x->min; int *sp = malloc(sizeof(int) * (nopt + 1));
                for (int k = 0; k < nopt && !g->err; k++) { sp[k] = emit(g, I_SPLIT); g->in[sp[k]].x = g->n; comp(g, x->a); }
                for (int k = 0; k < nopt && !g->err; k++) g->in[sp[k]].y = g->n;
                free(sp);
            }
            break;
        }
    }
}

typedef struct {
    Prog g; int literal_len; uint8_t *literal;   /* pure-literal fast path (memmem), like regex's literal optimisation */
} Regex;

static int regex_compile(const uint8_t *pat, int n, Regex *re, char *msg, int msglen) {
    Parser P; memset(&P, 0, sizeof P); P.p = pat; P.n = n; P.fu = 1;
    Node *root = parse_alt(&P);
    if (!P.err && P.pos < P.n) { if (P.p[P.pos] == ')') fail_(&P, "unopened group"); else fail_(&P, "unexpected character"); }
    memset(re, 0, sizeof *re);
    int rc = 0;
    if (P.err) { snprintf(msg, msglen, "%s", P.msg); rc = -1; }
    else {
        comp(&re->g, root); emit(&re->g, I_MATCH);
        if (re->g.err) { snprintf(msg, msglen, "compiled regex exceeds size limit"); rc = -1; }
        else {
            /* literal detection: program is BYTES{1 member}* MATCH */
            int lit = 1;
            for (int i = 0; i + 1 < re->g.n && lit; i++) {
                if (re->g.in[i].op != I_BYTES) { lit = 0; break; }
                int cnt = 0; for (int b = 0; b < 32; b++) cnt += __builtin_popcount(re->g.in[i].set[b]);
                if (cnt != 1) lit = 0;
            }
            if (lit && re->g.n > 1) {
                re->literal_len = re->g.n - 1; re->literal = malloc(re->literal_len);
                for (int i = 0; i < re->literal_len; i++) for (int c = 0; c < 256; c++) if (re->g.in[i].set[c >> 3] & (1 << (c & 7))) re->literal[i] = c;
            }
        }
    }
    for (int i = 0; i < P.nall; i++) free(P.all[i]);
    free(P.all);
    if (rc) { free(re->g.in); memset(re, 0, sizeof *re); }
    return rc;
}
static void regex_free(Regex *re) { free(re->g.in); free(re->literal); memset(re, 0, sizeof *re); }

static int look_ok(int k, const uint8_t *h, size_t n, size_t i) {
    int pw = i > 0 && isword(h[i - 1]), nw = i < n && isword(h[i]);
    switch (k) {
        case LK_START_TEXT: return i == 0;
        case LK_END_TEXT: return i == n;
        case LK_START_LINE: return i == 0 || h[i - 1] == '\n';
        case LK_END_LINE: return i == n || h[i] == '\n';
        case LK_WORD: return pw != nw;
        case LK_NOT_WORD: return pw == nw;
        case LK_WORD_START: return !pw && nw;
        case LK_WORD_END: return pw && !nw;
        case LK_WORD_START_HALF: return !pw;
        case LK_WORD_END_HALF: return !nw;
    }
    return 0;
}
typedef struct { int *list; int n; unsigned *mark; unsigned gen; } TSet;
typedef struct { TSet a, b; int *stack; int cap; } Scratch;
static void scratch_fit(Scratch *s, int n) {
    if (s->cap >= n) return;
    s->a.list = realloc(s->a.list, sizeof(int) * n); s->b.list = realloc(s->b.list, sizeof(int) * n);
    s->a.mark = realloc(s->a.mark, sizeof(unsigned) * n); s->b.mark = realloc(s->b.mark, sizeof(unsigned) * n);
    memset(s->a.mark, 0, sizeof(unsigned) * n); memset(s->b.mark, 0, sizeof(unsigned) * n);
    s->a.gen = s->b.gen = 0; s->stack = realloc(s->stack, sizeof(int) * n * 2 + 16); s->cap = n;
}
/* add pc with epsilon closure at haystack position i; returns 1 if MATCH reached */
static int addthread(const Prog *g, TSet *t, int *stack, int pc0, const uint8_t *h, size_t n, size_t i) {
    int sp = 0; stack[sp++] = pc0; int matched = 0;
    while (sp) {
        int pc = stack[--sp];
        if (t->mark[pc] == t->gen) continue;
        t->mark[pc] = t->gen;
        const Inst *in = &g->in[pc];
        switch (in->op) {
            case I_JMP: stack[sp++] = in->x; break;
            case I_SPLIT: stack[sp++] = in->y; stack[sp++] = in->x; break;
            case I_LOOK: if (look_ok(in->x, h, n, i)) stack[sp++] = pc + 1; break;
            case I_MATCH: matched = 1; break;
            case I_BYTES: t->list[t->n++] = pc; break;
            case I_FAIL: break;
        }
    }
    return matched;
}
static __thread Scratch tls_scratch;
static int regex_is_match(const Regex *re, const uint8_t *h, size_t n) {
    if (re->literal) return memmem(h, n, re->literal, re->literal_len) != NULL;
    const Prog *g = &re->g; Scratch *s = &tls_scratch; scratch_fit(s, g->n + 1);
    TSet *cur = &s->a, *nxt = &s->b; cur->n = 0; cur->gen++;
    for (size_t i = 0;; i++) {
        if (addthread(g, cur, s->stack, 0, h, n, i)) return 1;     /* unanchored: new thread at every offset */
        if (i == n) break;
        nxt->n = 0; nxt->gen++;
        uint8_t c = h[i];
        for (int k = 0; k < cur->n; k++) {
            const Inst *in = &g->in[cur->list[k]];
            if (in->set[c >> 3] & (1 << (c & 7))) if (addthread(g, nxt, s->stack, cur->list[k] + 1, h, n, i + 1)) return 1;
        }
        TSet *t = cur; cur = nxt; nxt = t;
    }
    return 0;
}

/* ------------------------------------------------------------------ matcher */
enum { FQO_FIXED = 0, FQO_FIXED_SET = 1, FQO_REGEX = 2, FQO_REGEX_SET = 3, FQO_NAMES = 4, FQO_BITMASK = 5, FQO_BITMASK_SET = 6 };

/* IUPAC_MASKS, src/lib/mod.rs:14-53 */
static uint8_t IUPAC[256]; static int iupac_init = 0;
static void init_iupac(void) {
    if (iupac_init) return;
    const char *k = "ACGTUMRWSYKVHDBN"; const uint8_t v[] = {1, 2, 4, 8, 8, 3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15};
    for (int i = 0; k[i]; i++) { IUPAC[(uint8_t)k[i]] = v[i]; IUPAC[(uint8_t)k[i] + 32] = v[i]; }
    iupac_init = 1;
}
/* bitenc_match_offsets / bitenc_find, src/lib/matcher.rs:307-375, restated literally (rolling additive hash included) */
static int bitenc_find(const uint8_t *bases, size_t nb, const uint8_t *needle, size_t nn, int has_iupac) {
    if (nb < nn) return 0;
    size_t bases_hash = 0, needle_hash = 0;
    if (!has_iupac) for (size_t i = 0; i < nn; i++) needle_hash += needle[i];
    for (size_t offset = 0; offset <= nb - nn; offset++) {
        if (!has_iupac) {
            if (offset == 0) { bases_hash = 0; for (size_t i = 0; i < nn; i++) bases_hash += bases[i]; }
            else { bases_hash += bases[offset + nn - 1]; bases_hash -= bases[offset - 1]; }
            if (bases_hash != needle_hash) continue;
        }
        size_t i = 0;
        for (; i < nn; i++) { uint8_t base = bases[offset + i]; if ((needle[i] & base) != base) break; }
        if (i == nn) return 1;
    }
    return 0;
}
typedef struct fqo_matcher {
    int kind, invert, revcomp, npat;
    uint8_t **pat; size_t *plen;
    Regex *re;
    /* names: open addressing */
    uint64_t *slot_hash; int32_t *slot_idx; size_t nslots;
    /* FIXED_SET whose patterns all have the same length: window hashing instead of one memmem per pattern (set_len > 0).
       Same answers as the memmem loop -- tests/test_oracle_kat.py checks the two against each other on random sets; it
       exists so that bench.py can check WHOLE 2 M-read masks for 10 000-pattern sets in seconds. */
    size_t set_len; int use_naive_set;
} fqo_matcher;

static uint64_t fnv(const uint8_t *p, size_t n) { uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; } return h | 1; }

void fqo_matcher_free(fqo_matcher *m) {
    if (!m) return;
    for (int i = 0; i < m->npat; i++) { if (m->pat) free(m->pat[i]); if (m->re) regex_free(&m->re[i]); }
    free(m->pat); free(m->plen); free(m->re); free(m->slot_hash); free(m->slot_idx); free(m);
}

/* patterns: blob + offs[n+1].  Error text follows matcher.rs:499 / :554. */
fqo_matcher *fqo_matcher_new(int kind, const uint8_t *blob, const uint64_t *offs, int n, int invert, int revcomp, char *err, int errlen) {
    fqo_matcher *m = calloc(1, sizeof *m);
    m->kind = kind; m->invert = invert; m->revcomp = revcomp; m->npat = n;
    m->pat = calloc(n ? n : 1, sizeof(uint8_t *)); m->plen = calloc(n ? n : 1, sizeof(size_t));
    for (int i = 0; i < n; i++) { m->plen[i] = offs[i + 1] - offs[i]; m->pat[i] = malloc(m->plen[i] + 1); memcpy(m->pat[i], blob + offs[i], m->plen[i]); m->pat[i][m->plen[i]] = 0; }
    if (kind == FQO_REGEX || kind == FQO_REGEX_SET) {
        m->re = calloc(n ? n : 1, sizeof(Regex));
        for (int i = 0; i < n; i++) {
            char msg[200];
            if (regex_compile(m->pat[i], (int)m->plen[i], &m->re[i], msg, sizeof msg)) {
                if (err) snprintf(err, errlen, "Invalid regular expression: '%s' (%s)", (char *)m->pat[i], msg);
                fqo_matcher_free(m); return NULL;
            }
        }
    }
    if (kind == FQO_FIXED_SET && n >= 32) {
        size_t L = m->plen[0]; int same = L > 0;
        for (int i = 1; i < n && same; i++) same = m->plen[i] == L;
        if (same) {
            m->set_len = L;
            m->nslots = 16; while (m->nslots < (size_t)n * 2 + 2) m->nslots *= 2;
            m->slot_hash = calloc(m->nslots, sizeof(uint64_t)); m->slot_idx = calloc(m->nslots, sizeof(int32_t));
            for (int i = 0; i < n; i++) {
                uint64_t h = fnv(m->pat[i], L); size_t s = h & (m->nslots - 1);
                while (m->slot_hash[s]) s = (s + 1) & (m->nslots - 1);
                m->slot_hash[s] = h; m->slot_idx[s] = i;
            }
        }
    }
    if (kind == FQO_NAMES) {
        m->nslots = 16; while (m->nslots < (size_t)n * 2 + 2) m->nslots *= 2;
        m->slot_hash = calloc(m->nslots, sizeof(uint64_t)); m->slot_idx = calloc(m->nslots, sizeof(int32_t));
        for (int i = 0; i < n; i++) {
            uint64_t h = fnv(m->pat[i], m->plen[i]); size_t s = h & (m->nslots - 1);
            while (m->slot_hash[s]) s = (s + 1) & (m->nslots - 1);
            m->slot_hash[s] = h; m->slot_idx[s] = i;
        }
    }
    return m;
}

void fqo_matcher_set_naive(fqo_matcher *m, int naive) { m->use_naive_set = naive; }

/* mod.rs:129-155 */
static uint8_t COMP[256]; static int comp_init = 0;
static void init_comp(void) {
    if (comp_init) return;
    for (int i = 0; i < 256; i++) COMP[i] = i;
    const char *a = "AGCTYRWSKMDVHBN", *b = "TCGARYWSMKHBDVN";
    for (int i = 0; i < 15; i++) { COMP[(uint8_t)a[i]] = b[i]; COMP[(uint8_t)a[i] + 32] = b[i] + 32; }
    comp_init = 1;
}
void fqo_revcomp(const uint8_t *s, size_t n, uint8_t *out) { init_comp(); for (size_t i = 0; i < n; i++) out[i] = COMP[s[n - 1 - i]]; }

static int found_in(const fqo_matcher *m, const uint8_t *s, size_t n) {
    switch (m->kind) {
        case FQO_FIXED: return m->plen[0] == 0 || memmem(s, n, m->pat[0], m->plen[0]) != NULL;          /* bstr find */
        case FQO_FIXED_SET:
            if (m->set_len && !m->use_naive_set) {      /* every window of the common length: is it one of the patterns? */
                const size_t L = m->set_len;
                for (size_t i = 0; i + L <= n; i++) {
                    uint64_t h = fnv(s + i, L); size_t q = h & (m->nslots - 1);
                    while (m->slot_hash[q]) {
                        if (m->slot_hash[q] == h && !memcmp(m->pat[m->slot_idx[q]], s + i, L)) return 1;
                        q = (q + 1) & (m->nslots - 1);
                    }
                }
                return 0;
            }
            for (int i = 0; i < m->npat; i++) if (m->plen[i] == 0 || memmem(s, n, m->pat[i], m->plen[i])) return 1;      /* aho-corasick is_match */
            return 0;
        case FQO_REGEX: return regex_is_match(&m->re[0], s, n);
        case FQO_REGEX_SET: for (int i = 0; i < m->npat; i++) if (regex_is_match(&m->re[i], s, n)) return 1; return 0;
        case FQO_BITMASK: case FQO_BITMASK_SET: {      /* matcher.rs:392-395, 442-450: encode(bases) then any needle */
            init_iupac();
            uint8_t stackbuf[512], *enc = n <= sizeof stackbuf ? stackbuf : malloc(n);
            for (size_t i = 0; i < n; i++) enc[i] = IUPAC[s[i]];
            int found = 0;
            for (int k = 0; k < m->npat && !found; k++) {
                uint8_t nbuf[512], *ne = m->plen[k] <= sizeof nbuf ? nbuf : malloc(m->plen[k]);
                int has_iupac = 0;
                for (size_t i = 0; i < m->plen[k]; i++) { ne[i] = IUPAC[m->pat[k][i]]; if (!(ne[i] == 1 || ne[i] == 2 || ne[i] == 4 || ne[i] == 8)) has_iupac = 1; }
                found = bitenc_find(enc, n, ne, m->plen[k], has_iupac);
                if (ne != nbuf) free(ne);
            }
            if (enc != stackbuf) free(enc);
            return found;
        }
        default: return 1;   /* names: bases_match is `!invert` (matcher.rs:616-618) */
    }
}
int fqo_bases_match(const fqo_matcher *m, const uint8_t *s, size_t n) { return found_in(m, s, n) != m->invert; }

static int names_contains(const fqo_matcher *m, const uint8_t *id, size_t n) {
    uint64_t h = fnv(id, n); size_t s = h & (m->nslots - 1);
    while (m->slot_hash[s]) {
        if (m->slot_hash[s] == h) { int i = m->slot_idx[s]; if (m->plen[i] == n && !memcmp(m->pat[i], id, n)) return 1; }
        s = (s + 1) & (m->nslots - 1);
    }
    return 0;
}
static size_t id_len(const uint8_t *head, size_t hl) { const uint8_t *sp = memchr(head, ' ', hl); return sp ? (size_t)(sp - head) : hl; }

/* matcher.rs:164-175 written exactly in the reference's two-branch form; :631-638 for names */
int fqo_read_match(const fqo_matcher *m, const uint8_t *head, size_t hl, const uint8_t *seq, size_t sl) {
    if (m->kind == FQO_NAMES) { int idm = names_contains(m, head, id_len(head, hl)); return m->invert ? !idm : idm; }
    int bm = fqo_bases_match(m, seq, sl);
    if (m->invert) {
        if (!bm) return 0;
        if (!m->revcomp) return 1;
        uint8_t stackbuf[512], *rc = sl <= sizeof stackbuf ? stackbuf : malloc(sl);
        fqo_revcomp(seq, sl, rc); int r = fqo_bases_match(m, rc, sl); if (rc != stackbuf) free(rc); return r;
    } else {
        if (bm) return 1;
        if (!m->revcomp) return 0;
        uint8_t stackbuf[512], *rc = sl <= sizeof stackbuf ? stackbuf : malloc(sl);
        fqo_revcomp(seq, sl, rc); int r = fqo_bases_match(m, rc, sl); if (rc != stackbuf) free(rc); return r;
    }
}

/* matcher.rs:128-147.  Returns 0 ok, -1 with the reference's message. */
int fqo_validate_fixed(const char *pattern, int protein, char *err, int errlen) {
    const char *valid = protein ? "ARNDCEQGHILKMFPSTWYVBJOUXZ" : "ACGTN"; const char *kind = protein ? "amino acids" : "DNA bases";
    size_t n = strlen(pattern);
    for (size_t i = 0; i < n;) {
        uint8_t c = pattern[i]; size_t l = c < 0x80 ? 1 : c >= 0xF0 ? 4 : c >= 0xE0 ? 3 : 2; if (i + l > n) l = n - i;
        if (c >= 0x80 || !strchr(valid, c)) {
            if (err) snprintf(err, errlen, "Fixed pattern must contain only %s: %.*s .. [%.*s] .. %s", kind, (int)i, pattern, (int)l, pattern + i, pattern + i + l);
            return -1;
        }
        i += l;
    }
    return 0;
}

/* expand_iupac_fixed_pattern (mod.rs:70-97) and expand_iupac_regex (mod.rs:101-118).  Results joined by '\n'.
 * Returns length, or -1 when the 10 000 limit is hit (err gets the reference's message). */
static int expand_rec(const uint8_t *p, size_t n, size_t i, char *prefix, char *out, size_t cap, size_t *olen, int *count, const char *orig, char *err, int errlen) {
    if (*count >= 10000) { if (err) snprintf(err, errlen, "IUPAC expansion exceeded 10000 patterns for input '%s'", orig); return -1; }
    if (i == n) {
        if (*olen + n + 1 > cap) return -1;
        if (*count) out[(*olen)++] = '\n';
        memcpy(out + *olen, prefix, n); *olen += n; (*count)++;
        return 0;
    }
    for (const char *b = "ACGT"; *b; b++)
        if (IUPAC[(uint8_t)*b] & IUPAC[p[i]]) { prefix[i] = *b; if (expand_rec(p, n, i + 1, prefix, out, cap, olen, count, orig, err, errlen)) return -1; }
    return 0;
}
long fqo_iupac_expand(const char *pattern, int to_regex, char *out, size_t cap, char *err, int errlen) {
    init_iupac();
    size_t n = strlen(pattern), olen = 0;
    if (to_regex) {
        for (size_t i = 0; i < n; i++) {
            char bases[5]; int nb = 0;
            for (const char *b = "ACGT"; *b; b++) if (IUPAC[(uint8_t)*b] & IUPAC[(uint8_t)pattern[i]]) bases[nb++] = *b;
            if (olen + 8 > cap) return -1;
            if (nb == 1) out[olen++] = bases[0];
            else { out[olen++] = '['; memcpy(out + olen, bases, nb); olen += nb; out[olen++] = ']'; }
        }
        return (long)olen;
    }
    char *prefix = malloc(n + 1); int count = 0;
    int rc = expand_rec((const uint8_t *)pattern, n, 0, prefix, out, cap, &olen, &count, pattern, err, errlen);
    free(prefix);
    return rc ? -1 : (long)olen;
}

/* ------------------------------------------------------------- FASTQ records */
typedef struct { const uint8_t *head, *seq, *qual; size_t hl, sl, ql; } Rec;
enum { FQO_OK = 0, FQO_E_INVALID_START = -1, FQO_E_INVALID_SEP = -2, FQO_E_UNEQUAL = -3, FQO_E_UNEXPECTED_END = -4,
       FQO_E_ODD_INTERLEAVED = -5, FQO_E_OUT_OF_SYNC = -6, FQO_E_ID_MISMATCH = -7 };

static const uint8_t *line_end(const uint8_t *p, const uint8_t *e) { const uint8_t *q = memchr(p, '\n', e - p); return q ? q : e; }
static size_t trim_cr(const uint8_t *p, size_t n) { return n && p[n - 1] == '\r' ? n - 1 : n; }
/* returns 1 record parsed, 0 clean end, <0 error.  seq_io 0.3.4 rules. */
static int next_rec(const uint8_t *buf, size_t n, size_t *pos, Rec *r) {
    const uint8_t *e = buf + n, *p = buf + *pos;
    if (p >= e) return 0;
    { const uint8_t *q = p; while (q < e && (*q == '\n' || *q == '\r')) q++; if (q == e) return 0; }   /* only blank lines left */
    if (*p != '@') return FQO_E_INVALID_START;
    const uint8_t *l1 = line_end(p, e); if (l1 == e) return FQO_E_UNEXPECTED_END;
    const uint8_t *s = l1 + 1, *l2 = line_end(s, e); if (l2 == e) return FQO_E_UNEXPECTED_END;
    const uint8_t *t = l2 + 1; if (t >= e) return FQO_E_UNEXPECTED_END;
    if (*t != '+') return FQO_E_INVALID_SEP;
    const uint8_t *l3 = line_end(t, e); if (l3 == e) return FQO_E_UNEXPECTED_END;
    const uint8_t *q = l3 + 1, *l4 = line_end(q, e);    /* final newline optional */
    r->head = p + 1; r->hl = trim_cr(p + 1, l1 - (p + 1));
    r->seq = s; r->sl = trim_cr(s, l2 - s);
    r->qual = q; r->ql = trim_cr(q, l4 - q);
    if (r->sl != r->ql) return l4 == e && r->ql < r->sl ? FQO_E_UNEXPECTED_END : FQO_E_UNEQUAL;
    *pos = (l4 == e) ? n : (size_t)(l4 + 1 - buf);
    return 1;
}
static size_t write_rec(uint8_t *out, const Rec *r) {   /* seq_io write_to: @head\nseq\n+\nqual\n */
    uint8_t *o = out; *o++ = '@'; memcpy(o, r->head, r->hl); o += r->hl; *o++ = '\n'; memcpy(o, r->seq, r->sl); o += r->sl;
    *o++ = '\n'; *o++ = '+'; *o++ = '\n'; memcpy(o, r->qual, r->ql); o += r->ql; *o++ = '\n'; return o - out;
}

/* Parse every record of a FASTQ buffer. Returns count or <0 error; fills recs (malloc'd). */
static int64_t parse_all(const uint8_t *buf, size_t n, Rec **out) {
    size_t cap = n / 64 + 16, cnt = 0, pos = 0; Rec *recs = malloc(sizeof(Rec) * cap);
    for (;;) {
        if (cnt == cap) { cap *= 2; recs = realloc(recs, sizeof(Rec) * cap); }
        int rc = next_rec(buf, n, &pos, &recs[cnt]);
        if (rc == 0) break;
        if (rc < 0) { free(recs); return rc; }
        cnt++;
    }
    *out = recs; return (int64_t)cnt;
}

typedef struct { const fqo_matcher *m; const Rec *r1, *r2; size_t lo, hi; uint8_t *flag; int err; } Job;
static void *job_run(void *arg) {
    Job *j = arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        if (j->r2) {     /* main.rs:731-756 */
            const Rec *a = &j->r1[i], *b = &j->r2[i];
            size_t la = id_len(a->head, a->hl), lb = id_len(b->head, b->hl);
            if (la != lb || memcmp(a->head, b->head, la)) { j->err = FQO_E_ID_MISMATCH; return NULL; }
            j->flag[i] = fqo_read_match(j->m, a->head, a->hl, a->seq, a->sl) ? 1 : fqo_read_match(j->m, b->head, b->hl, b->seq, b->sl) ? 2 : 0;
        } else { const Rec *a = &j->r1[i]; j->flag[i] = (uint8_t)fqo_read_match(j->m, a->head, a->hl, a->seq, a->sl); }
    }
    return NULL;
}

/*
 * Whole-input driver (main.rs:550-665): mode 0 single-end (r2 unused), 1 interleaved pairs in r1,
 * 2 pairs across r1/r2.  flags_out (optional, caller-sized) gets one byte per record/pair:
 * single 0/1, paired 0/1/2 like process_paired_reads.  out/out_len (optional): selected records
 * in input order, paired interleaved R1,R2 (main.rs:682-727).  out2 (optional, mode 1/2): when
 * non-NULL R2 records go there instead (the hidden --output a b split).
 * Returns number selected (records or pairs) or a negative FQO_E_* (the reference panics: exit 101).
 */
int64_t fqo_grep(const fqo_matcher *m, int mode, const uint8_t *r1, size_t n1, const uint8_t *r2, size_t n2,
                 int threads, uint8_t *flags_out, int64_t *n_units, uint8_t *out, size_t *out_len, uint8_t *out2, size_t *out2_len) {
    Rec *a = NULL, *b = NULL, *pa = NULL, *pb = NULL; int64_t na = parse_all(r1, n1, &a), nb = 0, units;
    if (na < 0) return na;
    if (mode == 2) { nb = parse_all(r2, n2, &b); if (nb < 0) { free(a); return nb; } if (na != nb) { free(a); free(b); return na < nb ? FQO_E_OUT_OF_SYNC : FQO_E_UNEXPECTED_END; } pa = a; pb = b; units = na; }
    else if (mode == 1) {
        if (na % 2) { free(a); return FQO_E_ODD_INTERLEAVED; }
        units = na / 2; pa = malloc(sizeof(Rec) * (units + 1)); pb = malloc(sizeof(Rec) * (units + 1));
        for (int64_t i = 0; i < units; i++) { pa[i] = a[2 * i]; pb[i] = a[2 * i + 1]; }
    } else { pa = a; units = na; }
    uint8_t *flag = flags_out ? flags_out : malloc(units + 1);
    if (threads < 1) threads = 1; if (threads > 256) threads = 256;
    pthread_t th[256]; Job jobs[256]; int err = 0;
    for (int t = 0; t < threads; t++) {
        jobs[t] = (Job){ m, pa, mode ? pb : NULL, (size_t)(units * t / threads), (size_t)(units * (t + 1) / threads), flag, 0 };
        if (threads > 1) pthread_create(&th[t], NULL, job_run, &jobs[t]); else job_run(&jobs[t]);
    }
    for (int t = 0; t < threads; t++) { if (threads > 1) pthread_join(th[t], NULL); if (jobs[t].err) err = jobs[t].err; }
    int64_t cnt = 0; size_t o1 = 0, o2 = 0;
    if (!err) for (int64_t i = 0; i < units; i++) if (flag[i]) {
        cnt++;
        if (out) { o1 += write_rec(out + o1, &pa[i]); if (mode) { if (out2) o2 += write_rec(out2 + o2, &pb[i]); else o1 += write_rec(out + o1, &pb[i]); } }
    }
    if (out_len) *out_len = o1; if (out2_len) *out2_len = o2; if (n_units) *n_units = units;
    if (!flags_out) free(flag);
    if (mode == 1) { free(pa); free(pb); }
    free(a); free(b);
    return err ? err : cnt;
}

/* SoA entry: bases[offs[i]..offs[i+1]) for i<n; mask bit i = read_match. Header-less, so not for NAMES. */
typedef struct { const fqo_matcher *m; const uint8_t *bases; const uint64_t *offs; size_t lo, hi; uint8_t *flag; } SJob;
/* SoA spans are sequences, except for a read-name matcher where they are header lines without '@' (matcher.rs:631-638) */
static void *sjob_run(void *arg) {
    SJob *j = arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        const uint8_t *p = j->bases + j->offs[i];
        size_t n = j->offs[i + 1] - j->offs[i];
        j->flag[i] = (uint8_t)(j->m->kind == FQO_NAMES ? fqo_read_match(j->m, p, n, (const uint8_t *)"", 0) : fqo_read_match(j->m, (const uint8_t *)"", 0, p, n));
    }
    return NULL;
}
int64_t fqo_match_soa(const fqo_matcher *m, const uint8_t *bases, const uint64_t *offs, size_t n, int threads, uint8_t *flags_out) {
    if (threads < 1) threads = 1; if (threads > 256) threads = 256;
    pthread_t th[256]; SJob jobs[256];
    for (int t = 0; t < threads; t++) { jobs[t] = (SJob){ m, bases, offs, n * t / threads, n * (t + 1) / threads, flags_out }; if (threads > 1) pthread_create(&th[t], NULL, sjob_run, &jobs[t]); else sjob_run(&jobs[t]); }
    if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    int64_t c = 0; for (size_t i = 0; i < n; i++) c += flags_out[i] != 0; return c;
}
