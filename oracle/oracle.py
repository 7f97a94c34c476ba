"""ctypes loader for the CPU oracle (oracle/fqgrep_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Never imported by the fqgrep_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfqgrep_oracle.so")

FIXED, FIXED_SET, REGEX, REGEX_SET, NAMES, BITMASK, BITMASK_SET = 0, 1, 2, 3, 4, 5, 6
SINGLE, INTERLEAVED, PAIRED = 0, 1, 2
ERRORS = {-1: "InvalidStart", -2: "InvalidSep", -3: "UnequalLengths", -4: "UnexpectedEnd",
          -5: "OddInterleaved", -6: "OutOfSync", -7: "IdMismatch"}


def build(force=False):
    src = os.path.join(_HERE, "fqgrep_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libfqgrep_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.fqo_matcher_new.restype = C.c_void_p
        L.fqo_matcher_new.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.fqo_matcher_free.argtypes = [C.c_void_p]
        L.fqo_matcher_set_naive.argtypes = [C.c_void_p, C.c_int]
        L.fqo_bases_match.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.fqo_read_match.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.fqo_revcomp.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.fqo_validate_fixed.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.fqo_grep.restype = C.c_int64
        L.fqo_grep.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int,
                               C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p,
                               C.POINTER(C.c_size_t)]
        L.fqo_iupac_expand.restype = C.c_long
        L.fqo_iupac_expand.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_int]
        L.fqo_match_soa.restype = C.c_int64
        L.fqo_match_soa.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        _lib = L
    return _lib


class OracleError(Exception):
    pass


def _b(x):
    return x.encode() if isinstance(x, str) else bytes(x)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _as_u8(buf):
    if buf is None:
        return None
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


class Matcher:
    """Oracle matcher: kind in {FIXED, FIXED_SET, REGEX, REGEX_SET, NAMES} (matcher.rs:645-685 dispatch)."""

    def __init__(self, kind, patterns, invert=False, revcomp=False):
        pats = [_b(p) for p in patterns]
        blob = b"".join(pats)
        offs = np.zeros(len(pats) + 1, dtype=np.uint64)
        if pats:
            offs[1:] = np.cumsum([len(p) for p in pats])
        err = C.create_string_buffer(512)
        self.kind = kind
        self._h = lib().fqo_matcher_new(kind, blob, _ptr(offs), len(pats), int(invert), int(revcomp), err, 512)
        if not self._h:
            raise OracleError(err.value.decode("utf-8", "replace"))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fqo_matcher_free(self._h)
            self._h = None

    def set_naive(self, naive=True):
        """FIXED_SET only: use the plain one-memmem-per-pattern loop even when all patterns have the same length."""
        lib().fqo_matcher_set_naive(self._h, int(naive))

    def bases_match(self, seq):
        s = _b(seq)
        return bool(lib().fqo_bases_match(self._h, s, len(s)))

    def read_match(self, seq, head=b""):
        s, h = _b(seq), _b(head)
        return bool(lib().fqo_read_match(self._h, h, len(h), s, len(s)))

    def grep(self, r1, r2=None, mode=SINGLE, threads=1, want_output=False, split=False):
        """Returns dict(count, units, flags[, out, out2]); raises OracleError on malformed input (reference: panic, exit 101)."""
        a, b = _as_u8(r1), _as_u8(r2)
        cap = a.size // 4 + 2
        flags = np.zeros(cap, dtype=np.uint8)
        units = C.c_int64(0)
        out = out2 = None
        ol, ol2 = C.c_size_t(0), C.c_size_t(0)
        if want_output:
            out = np.zeros(a.size + (b.size if b is not None else 0) + 16, dtype=np.uint8)
            if split:
                out2 = np.zeros(out.size, dtype=np.uint8)
        rc = lib().fqo_grep(self._h, mode, _ptr(a), a.size, _ptr(b), b.size if b is not None else 0, threads,
                            _ptr(flags), C.byref(units), _ptr(out), C.byref(ol), _ptr(out2), C.byref(ol2))
        if rc < 0:
            raise OracleError(ERRORS.get(rc, str(rc)))
        res = {"count": int(rc), "units": units.value, "flags": flags[:units.value].copy()}
        if want_output:
            res["out"] = out[:ol.value].tobytes()
            if split:
                res["out2"] = out2[:ol2.value].tobytes()
        return res

    def match_soa(self, bases, offsets, threads=1):
        bases = _as_u8(bases)
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offs.size - 1
        flags = np.zeros(max(n, 1), dtype=np.uint8)
        cnt = lib().fqo_match_soa(self._h, _ptr(bases), _ptr(offs), n, threads, _ptr(flags))
        return int(cnt), flags[:n]


def revcomp(seq):
    s = _b(seq)
    out = C.create_string_buffer(len(s) + 1)
    lib().fqo_revcomp(s, len(s), out)
    return out.raw[:len(s)]


def validate_fixed(pattern, protein=False):
    """None if ok, else the reference's error text (matcher.rs:138-143)."""
    err = C.create_string_buffer(1024)
    rc = lib().fqo_validate_fixed(pattern.encode("utf-8"), int(protein), err, 1024)
    return None if rc == 0 else err.value.decode("utf-8", "replace")


def expand_iupac_fixed_pattern(pattern):
    """mod.rs:70-97 -> list of expanded fixed strings; raises OracleError beyond 10 000."""
    out = C.create_string_buffer(1 << 22)
    err = C.create_string_buffer(512)
    n = lib().fqo_iupac_expand(pattern.encode(), 0, out, len(out), err, 512)
    if n < 0:
        raise OracleError(err.value.decode())
    return out.raw[:n].decode().split("\n")


def expand_iupac_regex(pattern):
    """mod.rs:101-118"""
    out = C.create_string_buffer(16 * len(pattern) + 16)
    n = lib().fqo_iupac_expand(pattern.encode(), 1, out, len(out), None, 0)
    return out.raw[:n].decode()


def new_matcher(pattern, fixed_strings, regexp, invert=False, revcomp=False):
    """MatcherFactory::new_matcher dispatch (matcher.rs:652-676), bitmask variants excluded."""
    if pattern is not None:
        return Matcher(FIXED if fixed_strings else REGEX, [pattern], invert, revcomp)
    return Matcher(FIXED_SET if fixed_strings else REGEX_SET, list(regexp), invert, revcomp)
