"""GPU tests of the streaming batcher (fqg_stream_*), the whole-input entry points built on it (bounded device memory,
ordered write-back from device spans) and the several-GPUs-one-process path, all against the oracle.

Reference behaviour being replaced: reader thread -> worker pool -> ordered main thread (src/main.rs:61-87,
src/lib/seq_io.rs:149-198, 253-282, 327-364) and the ordered write-back (src/main.rs:641-657, 682-727)."""
import ctypes as C
import os
import random

import numpy as np
import pytest

import fqgrep_b200 as F
from fqgrep_b200 import _ffi
from oracle import oracle as O
from tests import synth
from tests.fq_util import fastq

pytestmark = pytest.mark.gpu


def opts(inv=False, rc=False):
    return F.MatcherOpts(invert_match=inv, reverse_complement=rc)


def ragged_pair(rng, n, ids=None):
    ids = ids or ["q%d/%d" % (i, rng.randrange(10 ** rng.randrange(1, 8))) for i in range(n)]
    mk = lambda lo, hi: ["".join(rng.choice("ACGTN") for _ in range(rng.randrange(lo, hi))) for _ in range(n)]
    return fastq(mk(30, 200), heads=[i + " 1" for i in ids]), fastq(mk(30, 260), heads=[i + " 2:%s" % ("y" * rng.randrange(5)) for i in ids])


class slot_bytes:
    """FQG_SLOT_BYTES: the ring's slot size for the whole-input entry points (a test hook; fqg_stream_open takes it as an option)."""

    def __init__(self, n):
        self.n = n

    def __enter__(self):
        os.environ["FQG_SLOT_BYTES"] = str(self.n)

    def __exit__(self, *a):
        os.environ.pop("FQG_SLOT_BYTES", None)


@pytest.mark.parametrize("kind,pats,inv,rc", [
    (F.REGEX_SET, ["GACGAGATTA", "AC[GT]{3}TT"], False, True),
    (F.FIXED, ["ACGTAC"], True, False),
    (F.NAMES, ["q7", "q100", "q4999", "nope"], False, False),
])
def test_many_small_batches_equal_one_big_batch(kind, pats, inv, rc):
    """An input 40x the slot size: cut by the id-guided cutter, pushed through the ring, results concatenated."""
    rng = random.Random(21)
    n = 6000
    ids = ["q%d" % i for i in range(n)]
    r1, r2 = ragged_pair(rng, n, ids)
    # interleave record by record
    def recs(buf):
        out, pos, lines = [], 0, 0
        start = 0
        while True:
            nl = buf.find(b"\n", pos)
            if nl < 0:
                break
            lines += 1
            pos = nl + 1
            if lines % 4 == 0:
                out.append(buf[start:pos])
                start = pos
        return out
    inter = b"".join(x + y for x, y in zip(recs(r1), recs(r2)))
    m = F.Matcher(kind, pats, opts(inv, rc))
    o = O.Matcher(kind, pats, inv, rc)
    for mode, x, y in ((F.SINGLE, r1, None), (F.PAIRED, r1, r2), (F.INTERLEAVED, inter, None)):
        e = o.grep(x, y, mode=mode, want_output=True, threads=8)
        with slot_bytes(32 * 1024):
            g = m.grep_fastq(x, y, mode=mode)
            w = m.grep_fastq_write(x, y, mode=mode)
        assert g["units"] == e["units"] and g["count"] == e["count"], mode
        assert (g["mask"] == (e["flags"] != 0)).all(), mode
        assert w["count"] == e["count"] and w["out"] == e["out"], mode
    e = o.grep(r1, r2, mode=O.PAIRED, want_output=True, threads=8)
    with slot_bytes(48 * 1024):
        w = m.grep_fastq_write(r1, r2, mode=F.PAIRED, split=True)
    bits = np.zeros(((e["units"] + 31) // 32 + 1) * 32, dtype=np.uint8)
    bits[: e["units"]] = e["flags"] != 0
    o1, o2 = F.write_selected(np.packbits(bits, bitorder="little").view(np.uint32), r1, r2, mode=F.PAIRED, split=True)
    assert w["out"] == o1 and w["out2"] == o2


def test_input_four_times_the_resident_cap_64mb():
    """an input 4x larger than a 64 MB cap, byte-identical output (default 64 MiB slots)."""
    n = 900_000                                      # 285 MB per mate
    m1, m2 = synth.bases(n, seed=1), synth.bases(n, seed=2)
    pats = ["GACGAGATTA", "GACGTGATTA"]
    synth.plant(m1, pats, rate=0.01, seed=99)
    synth.plant(m2, pats, rate=0.01, seed=98)
    f1, f2 = synth.fastq(m1), synth.fastq(m2)
    m = F.Matcher(F.REGEX_SET, pats, opts(False, True))
    e = O.Matcher(O.REGEX_SET, pats, False, True).grep(f1, f2, mode=O.PAIRED, want_output=True, threads=os.cpu_count())
    g = m.grep_fastq(f1, f2, mode=F.PAIRED)
    assert g["units"] == n and g["count"] == e["count"] and (g["mask"] == (e["flags"] != 0)).all()
    w = m.grep_fastq_write(f1, f2, mode=F.PAIRED)
    assert w["out"] == e["out"]
    # -v: nearly everything is selected, so the write-back moves almost the whole input
    mi = F.Matcher(F.REGEX_SET, pats, opts(True, True))
    ei = O.Matcher(O.REGEX_SET, pats, True, True).grep(f1, None, mode=O.SINGLE, want_output=True, threads=os.cpu_count())
    wi = mi.grep_fastq_write(f1)
    assert wi["count"] == ei["count"] and wi["out"] == ei["out"]


def test_stream_api_orders_batches_and_returns_spans():
    rng = random.Random(2)
    m = F.Matcher(F.REGEX_SET, ["ACGT"], opts())
    o = O.Matcher(O.REGEX_SET, ["ACGT"], False, False)
    st = F.FastqStream(m, mode=F.SINGLE, n_slots=2, slot_bytes=1 << 20, want_spans=True)
    batches = [fastq(["".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 300))) for _ in range(rng.randrange(1, 400))],
                     heads=["b%d_%d" % (k, i) for i in range(400)]) for k in range(7)]
    got = []
    for k, b in enumerate(batches):
        if st.in_flight() == 2:
            got.append(st.next())
        if k % 2:
            st.feed(b)
        else:
            st.fill(b)
    while st.in_flight():
        got.append(st.next())
    with pytest.raises(F.FqgrepError):
        st.next()
    assert [g["seq"] for g in got] == list(range(7))
    for g, b in zip(got, batches):
        e = o.grep(b, want_output=True)
        assert g["count"] == e["count"] and (g["mask"] == (e["flags"] != 0)).all() and g["out"] == e["out"]
    # a batch that ends inside a record fails like the reference's reader, and the stream keeps working
    st.feed(batches[0][:-9])
    with pytest.raises(F.FqgrepError) as ex:
        st.next()
    assert ex.value.exit_code == 101
    st.feed(batches[1])
    assert st.next()["count"] == o.grep(batches[1])["count"]
    # more batches than slots without taking results is refused, not queued
    st.feed(batches[2]); st.feed(batches[3])
    with pytest.raises(F.FqgrepError):
        st.feed(batches[4])
    st.close()


def test_wrong_id_guided_cut_is_caught_and_recut_exactly():
    """Every read carries the same id, mates have different lengths: the id-guided cutter pairs up the wrong records,
    the device sees batches with different record counts, and the rest of the input is cut by counting."""
    rng = random.Random(8)
    n = 3000
    s1 = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 90))) for _ in range(n)]
    s2 = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 140))) for _ in range(n)]
    r1, r2 = fastq(s1, heads=["same"] * n), fastq(s2, heads=["same"] * n)
    m = F.Matcher(F.FIXED_SET, ["ACGTA", "TTTTT"], opts(False, True))
    e = O.Matcher(O.FIXED_SET, ["ACGTA", "TTTTT"], False, True).grep(r1, r2, mode=O.PAIRED, want_output=True)
    with slot_bytes(16 * 1024):
        g = m.grep_fastq(r1, r2, mode=F.PAIRED)
        w = m.grep_fastq_write(r1, r2, mode=F.PAIRED)
    assert g["units"] == n and g["count"] == e["count"] and (g["mask"] == (e["flags"] != 0)).all()
    assert w["out"] == e["out"]
    # genuinely broken pairing is still an error (and still exit 101)
    with slot_bytes(16 * 1024):
        with pytest.raises(F.FqgrepError) as ex:
            m.grep_fastq(r1, fastq(s2[:-1], heads=["same"] * (n - 1)), mode=F.PAIRED)
    assert ex.value.exit_code == 101
    ids = ["i%d" % i for i in range(n)]
    bad = list(ids)
    bad[2500] = "other"
    with slot_bytes(16 * 1024):
        with pytest.raises(F.FqgrepError) as ex:
            m.grep_fastq(fastq(s1, heads=ids), fastq(s2, heads=bad), mode=F.PAIRED)
    assert ex.value.exit_code == 101


def test_two_contexts_one_input_equals_one_context():
    """Multi-GPU product path on a one-GPU box: two contexts on device 0, batches dealt round-robin, host-side gather."""
    L = _ffi.lib()
    extra = C.c_void_p()
    assert L.fqg_ctx_create(0, C.byref(extra)) == 0, _ffi.last_error()
    try:
        rng = random.Random(13)
        n = 8000
        r1, r2 = ragged_pair(rng, n)
        pats = ["GATTACA", "AC[GT]T{2,}"]
        m = F.Matcher(F.REGEX_SET, pats, opts(True, True))
        o = O.Matcher(O.REGEX_SET, pats, True, True)
        ctxs = (C.c_void_p * 3)(F.context(0), extra, F.context(0))      # a context may also take every other batch twice over
        for mode, x, y in ((F.SINGLE, r1, None), (F.PAIRED, r1, r2)):
            e = o.grep(x, y, mode=mode, want_output=True, threads=8)
            a = np.frombuffer(x, dtype=np.uint8)
            b = None if y is None else np.frombuffer(y, dtype=np.uint8)
            words = a.size // 8 // 32 + 2
            mask = np.zeros(words, dtype=np.uint32)
            out = np.zeros(a.size + (0 if b is None else b.size) + 64, dtype=np.uint8)
            n_out = C.c_size_t(0)
            res = _ffi.Result()
            with slot_bytes(64 * 1024):
                rc = L.fqg_grep_fastq_host_multi(ctxs, 3, m._h, mode, a.ctypes.data, a.size, None if b is None else b.ctypes.data, 0 if b is None else b.size,
                                                 mask.ctypes.data, words, out.ctypes.data, out.size, C.byref(n_out), None, 0, None, C.byref(res))
            assert rc == 0, _ffi.last_error()
            assert res.n_units == e["units"] and res.n_selected == e["count"]
            got = np.unpackbits(mask.view(np.uint8), bitorder="little")[: res.n_units].astype(bool)
            assert (got == (e["flags"] != 0)).all()
            assert out[: n_out.value].tobytes() == e["out"]
    finally:
        L.fqg_ctx_destroy(extra)


def test_one_million_read_names():
    """BASELINE configs[4b] at full name count: 1 M ids (900 k present), 2 M records, -v too."""
    n, n_names = 2_000_000, 1_000_000
    f1 = synth.fastq(synth.bases(n, seed=1))
    rng = np.random.default_rng(5)
    chosen = rng.choice(n, 900_000, replace=False)
    names = ["r%010d" % i for i in chosen] + ["r%010d" % (n + i) for i in range(n_names - chosen.size)]
    want = np.zeros(n, dtype=bool)
    want[chosen] = True
    for inv in (False, True):
        g = F.Matcher(F.NAMES, names, opts(inv)).grep_fastq(f1)
        assert g["count"] == (n - chosen.size if inv else chosen.size)
        assert (g["mask"] == (want != inv)).all()
    e = O.Matcher(O.NAMES, names, False, False).grep(f1[: 317 * 200_000], threads=os.cpu_count())
    assert ((e["flags"] != 0) == want[:200_000]).all()
