"""Host-side mirror of the reference's matcher API over the C ABI.

Names follow /root/reference/src/lib/matcher.rs: MatcherOpts (:20-28), the Matcher trait's
`opts / bases_match / read_match` (:150-193), MatcherFactory::new_matcher / new_query_name_matcher
(:645-685), validate_fixed_pattern (:128-147).  The batch-level calls (`match_bases`, `grep_fastq`)
are the seam the GPU needs: one call per batch instead of one virtual call per record
(src/lib/seq_io.rs:314-326).  Everything executes on the device; there is no host matching path.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi

FIXED, FIXED_SET, REGEX, REGEX_SET, NAMES, BITMASK, BITMASK_SET = 0, 1, 2, 3, 4, 5, 6
SINGLE, INTERLEAVED, PAIRED = 0, 1, 2
INVERT_MATCH, REVERSE_COMPLEMENT = 1, 2


class FqgrepError(Exception):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code
        self.message = message

    @property
    def exit_code(self):
        """What the reference binary would exit with (src/main.rs:295-311)."""
        return {-2: 2, -4: 101, -5: 101}.get(self.code, 2)


def _check(rc):
    if rc != 0:
        raise FqgrepError(rc, _ffi.last_error())


@dataclass(frozen=True)
class MatcherOpts:
    invert_match: bool = False
    reverse_complement: bool = False

    def bits(self):
        return (INVERT_MATCH if self.invert_match else 0) | (REVERSE_COMPLEMENT if self.reverse_complement else 0)


def validate_fixed_pattern(pattern, protein=False):
    """Raises FqgrepError with the reference's message when the pattern is not a valid fixed pattern."""
    _check(_ffi.lib().fqg_validate_fixed_pattern(pattern.encode("utf-8"), int(protein)))


def _b(x):
    return x.encode("utf-8") if isinstance(x, str) else bytes(x)


_contexts = {}


def context(device=0):
    """One device context per (process, GPU).  Raises when no CUDA device is usable."""
    h = _contexts.get(device)
    if h is None:
        p = C.c_void_p()
        _check(_ffi.lib().fqg_ctx_create(device, C.byref(p)))
        h = _contexts[device] = p
    return h


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Matcher:
    """A compiled pattern set (`Box<dyn Matcher + Sync + Send>` in the reference)."""

    def __init__(self, kind, patterns, opts=MatcherOpts()):
        pats = [_b(p) for p in patterns]
        blob = b"".join(pats)
        offs = np.zeros(len(pats) + 1, dtype=np.uint64)
        if pats:
            offs[1:] = np.cumsum([len(p) for p in pats])
        h = C.c_void_p()
        _check(_ffi.lib().fqg_matcher_new(kind, blob, _ptr(offs), len(pats), opts.bits(), C.byref(h)))
        self._h = h
        self.kind = kind
        self._opts = opts
        self._pats = pats

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _ffi is not None and getattr(_ffi, "_lib", None) is not None:      # (module globals are gone at interpreter shutdown)
            _ffi._lib.fqg_matcher_free(h)
            self._h = None

    def opts(self):
        return self._opts

    # -- introspection of the compiled automaton (CPU-side tests of the compilers)
    def dfa_info(self):
        info = _ffi.DfaInfo()
        _check(_ffi.lib().fqg_matcher_dfa_info(self._h, C.byref(info)))
        return info

    def dfa_export(self):
        info = self.dfa_info()
        cls = np.zeros(256, dtype=np.uint8)
        nxt = np.zeros(info.n_states * info.n_classes, dtype=np.uint32)
        _check(_ffi.lib().fqg_matcher_dfa_export(self._h, _ptr(cls), _ptr(nxt)))
        return cls, nxt.reshape(info.n_states, info.n_classes), info.start, info.accept_from

    # -- batch seam
    def match_bases(self, bases, offsets, device=0):
        """read_match for every read of an SoA batch in host memory -> (bool array, n_selected)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size - 1
        mask = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
        cnt = C.c_uint64(0)
        _check(_ffi.lib().fqg_match_bases_host(context(device), self._h, _ptr(bases), _ptr(offsets), n, _ptr(mask), C.byref(cnt)))
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n].astype(bool)
        return bits, cnt.value

    def grep_fastq(self, r1, r2=None, mode=SINGLE, device=0, want_mask=True):
        """Parse + match raw FASTQ bytes (host memory).  Returns dict(count, units, records, mask, result)."""
        a = r1 if isinstance(r1, np.ndarray) else np.frombuffer(bytes(r1), dtype=np.uint8)
        b = None if r2 is None else (r2 if isinstance(r2, np.ndarray) else np.frombuffer(bytes(r2), dtype=np.uint8))
        words = a.size // 8 // 32 + 2
        mask = np.zeros(words, dtype=np.uint32) if want_mask else None
        res = _ffi.Result()
        _check(_ffi.lib().fqg_grep_fastq_host(context(device), self._h, mode, _ptr(a), a.size, _ptr(b), 0 if b is None else b.size,
                                              _ptr(mask), words if want_mask else 0, C.byref(res)))
        out = {"count": res.n_selected, "units": res.n_units, "records": res.n_records, "result": res}
        if want_mask:
            out["mask_words"] = mask
            out["mask"] = np.unpackbits(mask.view(np.uint8), bitorder="little")[:res.n_units].astype(bool)
        return out

    def grep_fastq_write(self, r1, r2=None, mode=SINGLE, device=0, split=False, devices=None):
        """Parse + match + ordered write-back of the selected records (src/main.rs:641-657, 682-727): the device returns
        the selected records' spans, the host copies them.  devices=[...] deals the batches to several GPUs.
        Returns dict(count, units, records, out[, out2], result)."""
        a = r1 if isinstance(r1, np.ndarray) else np.frombuffer(bytes(r1), dtype=np.uint8)
        b = None if r2 is None else (r2 if isinstance(r2, np.ndarray) else np.frombuffer(bytes(r2), dtype=np.uint8))
        cap = a.size + (0 if b is None else b.size) + 64
        out = np.zeros(cap, dtype=np.uint8)
        out2 = np.zeros(cap, dtype=np.uint8) if split else None
        n1, n2 = C.c_size_t(0), C.c_size_t(0)
        res = _ffi.Result()
        L = _ffi.lib()
        if devices is None:
            _check(L.fqg_grep_fastq_host_write(context(device), self._h, mode, _ptr(a), a.size, _ptr(b), 0 if b is None else b.size,
                                               _ptr(out), cap, C.byref(n1), _ptr(out2), cap if split else 0, C.byref(n2), C.byref(res)))
        else:
            ctxs = (C.c_void_p * len(devices))(*[context(d) for d in devices])
            _check(L.fqg_grep_fastq_host_multi(ctxs, len(devices), self._h, mode, _ptr(a), a.size, _ptr(b), 0 if b is None else b.size, None, 0,
                                               _ptr(out), cap, C.byref(n1), _ptr(out2), cap if split else 0, C.byref(n2), C.byref(res)))
        ret = {"count": res.n_selected, "units": res.n_units, "records": res.n_records, "result": res, "out": out[:n1.value].tobytes()}
        if split:
            ret["out2"] = out2[:n2.value].tobytes()
        return ret

    def grep_fastq_multi(self, r1, r2=None, mode=SINGLE, devices=(0,)):
        """One input over several GPUs of the box (batches dealt round-robin, results gathered in batch order on the host):
        same mask and count as grep_fastq."""
        a = r1 if isinstance(r1, np.ndarray) else np.frombuffer(bytes(r1), dtype=np.uint8)
        b = None if r2 is None else (r2 if isinstance(r2, np.ndarray) else np.frombuffer(bytes(r2), dtype=np.uint8))
        words = a.size // 8 // 32 + 2
        mask = np.zeros(words, dtype=np.uint32)
        res = _ffi.Result()
        ctxs = (C.c_void_p * len(devices))(*[context(d) for d in devices])
        _check(_ffi.lib().fqg_grep_fastq_host_multi(ctxs, len(devices), self._h, mode, _ptr(a), a.size, _ptr(b), 0 if b is None else b.size,
                                                    _ptr(mask), words, None, 0, None, None, 0, None, C.byref(res)))
        return {"count": res.n_selected, "units": res.n_units, "records": res.n_records, "result": res, "mask_words": mask,
                "mask": np.unpackbits(mask.view(np.uint8), bitorder="little")[:res.n_units].astype(bool)}

    # -- per-record convenience with the reference's names (1-record batches; tests and KATs only)
    def bases_match(self, bases, device=0):
        """Matcher::bases_match: pattern found != invert_match (matcher.rs:204,250,507,570); no reverse complement."""
        if self._opts.reverse_complement:
            fwd = Matcher(self.kind, self._pats, MatcherOpts(self._opts.invert_match, False))
            return fwd.bases_match(bases, device)
        b = np.frombuffer(_b(bases), dtype=np.uint8)
        bits, _ = self.match_bases(b, np.array([0, b.size], dtype=np.uint64), device)
        return bool(bits[0])

    def read_match(self, seq, head=b"", device=0):
        """Matcher::read_match (matcher.rs:164-175; QueryNameMatcher override :631-638)."""
        if self.kind == NAMES:
            s = _b(seq)
            rec = b"@" + _b(head) + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n"
            return bool(self.grep_fastq(rec)["mask"][0])
        b = np.frombuffer(_b(seq), dtype=np.uint8)
        bits, _ = self.match_bases(b, np.array([0, b.size], dtype=np.uint64), device)
        return bool(bits[0])


class FastqStream:
    """The streaming batcher (fqg_stream_*): a ring of batches in flight on one GPU, results in submission order.
    Replaces reader thread -> worker pool -> ordered main thread of the reference (src/main.rs:61-87,
    src/lib/seq_io.rs:149-198, 253-282, 327-364)."""

    def __init__(self, matcher, mode=SINGLE, device=0, n_slots=3, slot_bytes=64 << 20, want_spans=False):
        self._m = matcher
        self.mode = mode
        opts = _ffi.StreamOpts(n_slots, int(want_spans), slot_bytes)
        h = C.c_void_p()
        _check(_ffi.lib().fqg_stream_open(context(device), matcher._h, mode, C.byref(opts), C.byref(h)))
        self._h = h
        self._keep = []

    def feed(self, r1, r2=None):
        """Submit a batch of whole records living in the caller's memory (kept alive until its result is taken)."""
        a = r1 if isinstance(r1, np.ndarray) else np.frombuffer(bytes(r1), dtype=np.uint8)
        b = None if r2 is None else (r2 if isinstance(r2, np.ndarray) else np.frombuffer(bytes(r2), dtype=np.uint8))
        _check(_ffi.lib().fqg_stream_feed(self._h, _ptr(a), a.size, _ptr(b), 0 if b is None else b.size))
        self._keep.append((a, b))

    def fill(self, r1, r2=None):
        """Copy a batch into the next free slot's pinned buffers and submit it (what a reader thread does)."""
        p1, p2, cap = C.c_void_p(), C.c_void_p(), C.c_size_t(0)
        _check(_ffi.lib().fqg_stream_acquire(self._h, C.byref(p1), C.byref(p2), C.byref(cap)))
        for src, dst in ((r1, p1), (r2, p2)):
            if src is not None:
                b = bytes(src)
                assert len(b) <= cap.value
                C.memmove(dst, b, len(b))
        _check(_ffi.lib().fqg_stream_submit(self._h, len(bytes(r1)), 0 if r2 is None else len(bytes(r2))))
        self._keep.append(None)

    def in_flight(self):
        return _ffi.lib().fqg_stream_in_flight(self._h)

    def next(self):
        """Result of the oldest batch in flight: dict(seq, count, units, records, mask, spans, out)."""
        b = _ffi.Batch()
        rc = _ffi.lib().fqg_stream_next(self._h, C.byref(b))
        if self._keep:
            self._keep.pop(0)
        _check(rc)
        words = (b.n_units + 31) // 32
        mw = np.ctypeslib.as_array(b.unit_mask, shape=(max(words, 1),))[:words].copy() if words else np.zeros(0, dtype=np.uint32)
        out = {"seq": b.seq, "count": b.n_selected, "units": b.n_units, "records": b.n_records, "mask_words": mw,
               "mask": np.unpackbits(mw.view(np.uint8), bitorder="little")[:b.n_units].astype(bool)}
        if b.n_spans:
            cap = b.n1 + b.n2 + 64
            buf = np.zeros(cap, dtype=np.uint8)
            n1, n2 = C.c_size_t(0), C.c_size_t(0)
            _check(_ffi.lib().fqg_write_spans(b.spans, b.n_spans, b.r1, b.n1, b.r2, b.n2, _ptr(buf), cap, C.byref(n1), None, 0, C.byref(n2)))
            out["out"] = buf[:n1.value].tobytes()
            out["spans"] = [(b.spans[i].offset, b.spans[i].len, b.spans[i].flags) for i in range(min(b.n_spans, 1 << 16))]
        else:
            out["out"] = b""
            out["spans"] = []
        return out

    def close(self):
        if self._h:
            _ffi.lib().fqg_stream_close(self._h)
            self._h = None

    def __del__(self):
        self.close()


def fastq_cut(r1, r2=None, mode=SINGLE, limit=0, final=False, exact=False):
    """fqg_fastq_cut: largest prefixes of whole records (mates in lock step) of at most `limit` bytes."""
    a = r1 if isinstance(r1, np.ndarray) else np.frombuffer(bytes(r1), dtype=np.uint8)
    b = None if r2 is None else (r2 if isinstance(r2, np.ndarray) else np.frombuffer(bytes(r2), dtype=np.uint8))
    c1, c2 = C.c_size_t(0), C.c_size_t(0)
    _check(_ffi.lib().fqg_fastq_cut(mode, _ptr(a), a.size, _ptr(b), 0 if b is None else b.size, limit, int(final), int(exact), C.byref(c1), C.byref(c2)))
    return (c1.value, c2.value) if mode == PAIRED else c1.value


def is_gzip_path(path):
    """is_gzip_path (src/lib/mod.rs:170-175)."""
    return bool(_ffi.lib().fqg_is_gzip_path(str(path).encode()))


def is_fastq_path(path):
    """is_fastq_path (src/lib/mod.rs:177-183)."""
    return bool(_ffi.lib().fqg_is_fastq_path(str(path).encode()))


def read_input(path, decompress=False, pinned=False):
    """spawn_reader of the reference (src/main.rs:61-87): file -> optional (multi-member) gunzip -> bytes as a numpy array.
    pinned=True keeps the bytes in CUDA pinned memory (the caller must keep the returned array alive and not resize it)."""
    p = C.c_void_p()
    n = C.c_size_t(0)
    _check(_ffi.lib().fqg_read_input(str(path).encode(), int(decompress), int(pinned), C.byref(p), C.byref(n)))
    view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(n.value, 1),))[: n.value]
    if pinned:
        return _PinnedInput(view, p)
    out = view.copy()
    _ffi.lib().fqg_free_input(p, 0)
    return out


class _PinnedInput(np.ndarray):
    """numpy view of a pinned input buffer; frees it with the library when collected."""

    def __new__(cls, view, handle):
        obj = np.asarray(view).view(cls)
        obj._handle = handle
        return obj

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            _ffi.lib().fqg_free_input(h, 1)
            self._handle = None


def write_selected(mask_words, r1, r2=None, mode=SINGLE, split=False):
    """Selected records in input order, seq_io write_to form (src/main.rs:641-657, 682-727)."""
    a = r1 if isinstance(r1, np.ndarray) else np.frombuffer(bytes(r1), dtype=np.uint8)
    b = None if r2 is None else (r2 if isinstance(r2, np.ndarray) else np.frombuffer(bytes(r2), dtype=np.uint8))
    cap = a.size + (0 if b is None else b.size) + 64
    out = np.zeros(cap, dtype=np.uint8)
    out2 = np.zeros(cap, dtype=np.uint8) if split else None
    n1, n2 = C.c_size_t(0), C.c_size_t(0)
    mw = np.ascontiguousarray(mask_words, dtype=np.uint32)
    _check(_ffi.lib().fqg_write_selected(mode, _ptr(a), a.size, _ptr(b), 0 if b is None else b.size, _ptr(mw), _ptr(out), C.byref(n1),
                                         _ptr(out2), C.byref(n2)))
    if split:
        return out[:n1.value].tobytes(), out2[:n2.value].tobytes()
    return out[:n1.value].tobytes()


def expand_iupac_fixed_pattern(pattern):
    """expand_iupac_fixed_pattern (src/lib/mod.rs:70-97): all ACGT strings an IUPAC pattern stands for (<= 10 000)."""
    n = C.c_size_t(0)
    _check(_ffi.lib().fqg_iupac_expand(pattern.encode(), 0, None, 0, C.byref(n)))
    buf = C.create_string_buffer(n.value + 1)
    _check(_ffi.lib().fqg_iupac_expand(pattern.encode(), 0, buf, n.value, C.byref(n)))
    return buf.raw[:n.value].decode().split("\n")


def expand_iupac_regex(pattern):
    """expand_iupac_regex (src/lib/mod.rs:101-118): "GATK" -> "GAT[GT]"."""
    n = C.c_size_t(0)
    _check(_ffi.lib().fqg_iupac_expand(pattern.encode(), 1, None, 0, C.byref(n)))
    buf = C.create_string_buffer(n.value + 1)
    _check(_ffi.lib().fqg_iupac_expand(pattern.encode(), 1, buf, n.value, C.byref(n)))
    return buf.raw[:n.value].decode()


def expand_iupac(pattern, regexp, fixed_strings, iupac):
    """expand_iupac of the reference (src/main.rs:315-372).  iupac in {"never","expand","regex","bit-mask"}.
    Returns (pattern, regexp, fixed_strings, bitencs) ready for MatcherFactory.new_matcher."""
    if iupac == "never" or not fixed_strings:
        return pattern, list(regexp), fixed_strings, None
    patterns = [pattern] if pattern is not None else list(regexp)
    if iupac == "expand":
        return None, [e for p in patterns for e in expand_iupac_fixed_pattern(p)], True, None
    if iupac == "regex":
        return None, [expand_iupac_regex(p) for p in patterns], False, None
    if iupac == "bit-mask":
        if len(patterns) == 1:
            return patterns[0], [], True, patterns
        return None, patterns, True, patterns
    raise ValueError(iupac)


class MatcherFactory:
    """MatcherFactory of the reference (matcher.rs:641-686).  `bitencs` (the reference passes Vec<BitEnc>) is the list
    of IUPAC pattern strings; encoding to 4-bit masks happens inside the library."""

    @staticmethod
    def new_matcher(pattern, fixed_strings, regexp, bitencs=None, match_opts=MatcherOpts()):
        if bitencs is not None:
            return Matcher(BITMASK if pattern is not None else BITMASK_SET, list(bitencs), match_opts)
        if pattern is not None:
            return Matcher(FIXED if fixed_strings else REGEX, [pattern], match_opts)
        return Matcher(FIXED_SET if fixed_strings else REGEX_SET, list(regexp), match_opts)

    @staticmethod
    def new_query_name_matcher(query_names, match_opts=MatcherOpts()):
        return Matcher(NAMES, sorted(_b(n) for n in query_names), match_opts)
