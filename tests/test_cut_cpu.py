"""Host side of the streaming path, no GPU needed: the batch cutter (fqg_fastq_cut) and the span writer (fqg_write_spans).

The cutter stands in for the lock-step batching of the reference's readers (InterleavedFastqReader::fill_data,
src/lib/seq_io.rs:60-80; PairedFastqReader::fill_data "same number of records", :149-198): every prefix it returns must
consist of whole records, with equal record counts for R1/R2 and an even count for interleaved input."""
import ctypes as C
import random

import numpy as np
import pytest

import fqgrep_b200 as F
from fqgrep_b200 import _ffi
from tests.fq_util import fastq


def records(buf):
    """Offsets of record starts by a plain 4-line walk + the end offset."""
    starts, pos, lines = [0], 0, 0
    while True:
        nl = buf.find(b"\n", pos)
        if nl < 0:
            break
        lines += 1
        pos = nl + 1
        if lines % 4 == 0:
            starts.append(pos)
    return starts


def random_pair(rng, n, crlf=False, same_len=False):
    """R1/R2 with the same ids; read and header lengths vary independently per mate unless same_len."""
    ids = ["read%d:%d" % (i, rng.randrange(10 ** rng.randrange(1, 9))) for i in range(n)]
    s1 = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 120))) for _ in range(n)]
    s2 = s1 if same_len else ["".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 160))) for _ in range(n)]
    h1 = [i + " 1:N:0" for i in ids]
    h2 = [i + (" 2:N:0" if same_len else " 2:N:0:%s" % ("x" * rng.randrange(0, 9))) for i in ids]
    return fastq(s1, heads=h1, crlf=crlf), fastq(s2, heads=h2, crlf=crlf)


@pytest.mark.parametrize("exact", [False, True])
def test_single_cut_is_last_record_start(exact):
    rng = random.Random(1)
    buf = fastq(["".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 200))) for _ in range(400)], qual_char="@")      # '@' quality lines
    st = records(buf)
    for limit in [1, 50, 317, 1000, 4096, len(buf) - 1, len(buf), len(buf) + 10]:
        cut = F.fastq_cut(buf, limit=limit, final=False, exact=exact)
        want = max([s for s in st if s <= limit] or [0])
        assert cut == want, (limit, cut, want)
    # a buffer that ends inside a record: the cut is that record's start; with final=True and room, everything
    part = buf[: st[200] + 7]
    assert F.fastq_cut(part, limit=0, final=False) == st[200]
    assert F.fastq_cut(part, limit=0, final=True) == len(part)
    # a record without its final newline only counts when the input ends there
    assert F.fastq_cut(buf[:-1], limit=0, final=False) == st[-2]
    assert F.fastq_cut(buf[:-1], limit=0, final=True) == len(buf) - 1


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("same_len", [False, True])
def test_paired_cut_keeps_mates_in_lock_step(exact, same_len):
    rng = random.Random(7 + same_len)
    r1, r2 = random_pair(rng, 600, same_len=same_len)
    st1, st2 = records(r1), records(r2)
    for limit in [300, 2000, 9000, 30000, max(len(r1), len(r2)) + 1]:
        for final in (False, True):
            c1, c2 = F.fastq_cut(r1, r2, mode=F.PAIRED, limit=limit, final=final, exact=exact)
            assert c1 in st1 and c2 in st2, (limit, c1, c2)
            assert st1.index(c1) == st2.index(c2), (limit, st1.index(c1), st2.index(c2))
            assert c1 <= limit and c2 <= limit
            # close to maximal: a handful of further pairs would not fit (R2's longer records make the cutter step back)
            k = st1.index(c1)
            if k + 4 < len(st1):
                assert st1[k + 4] > limit or st2[k + 4] > limit
    # buffers that end mid-record (what a reader thread has after a read()): only pairs whole in BOTH count
    a, b = r1[: st1[300] + 11], r2[: st2[280] + 3]
    c1, c2 = F.fastq_cut(a, b, mode=F.PAIRED, limit=0, final=False, exact=exact)
    assert (c1, c2) == (st1[280], st2[280])


def test_paired_cut_walks_through_a_whole_input():
    rng = random.Random(3)
    r1, r2 = random_pair(rng, 2000)
    st1, st2 = records(r1), records(r2)
    p1 = p2 = 0
    n_batches = 0
    while p1 < len(r1):
        c1, c2 = F.fastq_cut(r1[p1:], r2[p2:], mode=F.PAIRED, limit=5000, final=True)
        assert c1 > 0
        p1 += c1
        p2 += c2
        assert st1.index(p1) == st2.index(p2)
        n_batches += 1
    assert p2 == len(r2) and n_batches > 10


@pytest.mark.parametrize("exact", [False, True])
def test_interleaved_cut_never_splits_a_pair(exact):
    rng = random.Random(11)
    r1, r2 = random_pair(rng, 300)
    a, b = records(r1), records(r2)
    inter = b"".join(r1[a[i]:a[i + 1]] + r2[b[i]:b[i + 1]] for i in range(300))
    st = records(inter)
    for limit in range(400, 20000, 977):
        cut = F.fastq_cut(inter, mode=F.INTERLEAVED, limit=limit, final=False, exact=exact)
        assert cut in st and st.index(cut) % 2 == 0 and cut <= limit
        k = st.index(cut)
        assert k + 2 >= len(st) or st[k + 2] > limit
    # ends after an R1: that record stays for the next batch
    assert F.fastq_cut(inter[: st[101]], mode=F.INTERLEAVED, limit=0, final=False, exact=exact) == st[100]
    assert F.fastq_cut(inter[: st[100]], mode=F.INTERLEAVED, limit=0, final=False, exact=exact) == st[100]


def test_exact_cut_when_ids_do_not_help():
    # every read has the same id: the id-guided cut cannot align the mates, the counting cut can (and the id-guided
    # entry point falls back to it by itself when it finds nothing... here ids always "match", so only exact is right)
    rng = random.Random(5)
    n = 300
    s1 = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 90))) for _ in range(n)]
    s2 = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(20, 90))) for _ in range(n)]
    r1, r2 = fastq(s1, heads=["same"] * n), fastq(s2, heads=["same"] * n)
    st1, st2 = records(r1), records(r2)
    c1, c2 = F.fastq_cut(r1, r2, mode=F.PAIRED, limit=9000, final=False, exact=True)
    assert st1.index(c1) == st2.index(c2) and c1 > 0


def _spans(items):
    arr = (_ffi.Span * len(items))()
    for i, (o, l, f) in enumerate(items):
        arr[i].offset, arr[i].len, arr[i].flags = o, l, f
    return arr


def test_write_spans_copies_and_normalises():
    L = _ffi.lib()
    plain = fastq(["ACGT", "GGGG", "TTTTT"])
    crlf = fastq(["ACGT", "GG"], heads=["a b", "c"], crlf=True, plus_name=True)
    st, sc = records(plain), [0, crlf.index(b"@c")]
    items = [(st[2], st[3] - st[2], 0), (st[0], st[1] - st[0], 0), (sc[1], len(crlf) - sc[1], 1 | 2), (sc[0], sc[1], 1 | 2)]
    out = np.zeros(1024, dtype=np.uint8)
    n1, n2 = C.c_size_t(0), C.c_size_t(0)
    a, b = np.frombuffer(plain, dtype=np.uint8), np.frombuffer(crlf, dtype=np.uint8)
    rc = L.fqg_write_spans(_spans(items), 4, a.ctypes.data, a.size, b.ctypes.data, b.size, out.ctypes.data, out.size, C.byref(n1), None, 0, C.byref(n2))
    assert rc == 0, _ffi.last_error()
    assert out[: n1.value].tobytes() == plain[st[2]:st[3]] + plain[st[0]:st[1]] + b"@c\nGG\n+\nXX\n" + b"@a b\nACGT\n+\nXXXX\n"
    # split outputs: R2 spans go to the second buffer
    out2 = np.zeros(1024, dtype=np.uint8)
    rc = L.fqg_write_spans(_spans(items), 4, a.ctypes.data, a.size, b.ctypes.data, b.size, out.ctypes.data, out.size, C.byref(n1), out2.ctypes.data, out2.size,
                           C.byref(n2))
    assert rc == 0
    assert out[: n1.value].tobytes() == plain[st[2]:st[3]] + plain[st[0]:st[1]]
    assert out2[: n2.value].tobytes() == b"@c\nGG\n+\nXX\n" + b"@a b\nACGT\n+\nXXXX\n"
    # a record without its final newline, flagged for normalisation
    tail = plain[st[2]:-1]
    t = np.frombuffer(tail, dtype=np.uint8)
    rc = L.fqg_write_spans(_spans([(0, len(tail), 2)]), 1, t.ctypes.data, t.size, None, 0, out.ctypes.data, out.size, C.byref(n1), None, 0, C.byref(n2))
    assert rc == 0 and out[: n1.value].tobytes() == plain[st[2]:]
    # too small an output buffer and spans outside the input are errors, not overruns
    assert L.fqg_write_spans(_spans([(0, len(tail), 2)]), 1, t.ctypes.data, t.size, None, 0, out.ctypes.data, 3, C.byref(n1), None, 0, C.byref(n2)) == -6
    assert L.fqg_write_spans(_spans([(5, len(tail), 0)]), 1, t.ctypes.data, t.size, None, 0, out.ctypes.data, out.size, C.byref(n1), None, 0, C.byref(n2)) == -1
