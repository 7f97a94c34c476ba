"""Input plumbing: extension rules and host-side gunzip, against the reference's vectors
(src/lib/mod.rs:209-233, src/main.rs:1279-1297) and Python's gzip.  CPU only except the end-to-end count."""
import gzip
import os

import numpy as np
import pytest

import fqgrep_b200 as F
from oracle import oracle as O
from tests.fq_util import fastq


def test_extension_rules_match_reference_vectors():
    for name, exp in (("test_fastq.fq.gz", True), ("test_fastq.fq.bgz", True), ("test_fastq.fq.tar", False)):
        assert F.is_gzip_path("/tmp/x/" + name) == exp
    for name, exp in (("test_fastq.fq", True), ("test_fastq.fastq", True), ("test_fastq.sam", False)):
        assert F.is_fastq_path("/tmp/x/" + name) == exp
    assert not F.is_gzip_path("/tmp/.gz") and not F.is_gzip_path("/tmp/a.b/gz")


def _write(path, data, mode):
    if mode == "plain":
        open(path, "wb").write(data)
    elif mode == "gz":
        with gzip.open(path, "wb") as f:
            f.write(data)
    else:   # multi-member (what bgzip / concatenated gzips look like)
        with open(path, "wb") as f:
            for i in range(0, len(data), 700):
                f.write(gzip.compress(data[i:i + 700]))


def test_read_input_plain_gzip_multimember_and_decompress_flag(tmp_path):
    data = fastq(["ACGT" * 40, "TTTTGGGGCCCC" * 12, "N" * 7] * 200)
    cases = [("a.fq", "plain", False), ("a.fastq", "plain", True),     # -Z does not apply to .fq/.fastq (main.rs:77)
             ("a.fq.gz", "gz", False), ("a.fq.bgz", "multi", False), ("a.dat", "gz", True), ("a.txt", "plain", False)]
    for name, mode, z in cases:
        p = str(tmp_path / name)
        _write(p, data, mode)
        got = F.read_input(p, decompress=z)
        assert got.tobytes() == data, name
    with pytest.raises(F.FqgrepError):
        F.read_input(str(tmp_path / "missing.fq"))
    bad = str(tmp_path / "bad.fq.gz")
    open(bad, "wb").write(b"this is not gzip")
    with pytest.raises(F.FqgrepError):
        F.read_input(bad)
    empty = str(tmp_path / "empty.fq")
    open(empty, "wb").write(b"")
    assert F.read_input(empty).size == 0


@pytest.mark.gpu
def test_count_is_the_same_for_fq_fastq_gz_bgz(tmp_path, kat):
    """main.rs:1279-1297: the same reads behind .fq / .fastq / .fq.gz / .fq.bgz give the same count."""
    seqs = ["AAAA", "TTTT", "GGGG", "CCCC", "ACGT"] * 50
    data = fastq(seqs)
    m = F.MatcherFactory.new_matcher(None, False, ["A"])
    exp = O.Matcher(O.REGEX_SET, ["A"]).grep(data)["count"]
    for name, mode in (("r.fq", "plain"), ("r.fastq", "plain"), ("r.fq.gz", "gz"), ("r.fq.bgz", "multi")):
        p = str(tmp_path / name)
        _write(p, data, mode)
        buf = F.read_input(p, pinned=True)
        assert m.grep_fastq(buf)["count"] == exp == 100


def bgzf(data, block=900, eof=True):
    """What bgzip writes: independent gzip members with a 'BC' extra field holding the member size - 1."""
    import struct
    import zlib
    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof else [])
    for ch in chunks:
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = c.compress(ch) + c.flush()
        bsize = 12 + 6 + len(payload) + 8
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += payload + struct.pack("<II", zlib.crc32(ch) & 0xFFFFFFFF, len(ch))
    return bytes(out)


def test_bgzf_members_are_inflated_in_parallel_and_checked(tmp_path):
    """BGZF input takes the parallel path (members located through their BC fields); the bytes are those Python's gzip
    reads, damage is an error like in the serial decoder, and anything that is not pure BGZF falls back to it."""
    data = fastq(["ACGT" * 40, "TTTTGGGGCCCC" * 12, "N" * 7] * 3000)
    z = bgzf(data)
    assert gzip.decompress(z) == data and z.count(b"BC\x02\x00") > 500
    p = str(tmp_path / "r.fq.gz")
    open(p, "wb").write(z)
    assert F.read_input(p).tobytes() == data
    assert F.read_input(p, pinned=False).tobytes() == data
    # one flipped payload byte: CRC or stream error
    bad = bytearray(z)
    bad[len(z) // 2] ^= 0x55
    open(str(tmp_path / "bad.fq.gz"), "wb").write(bytes(bad))
    with pytest.raises(F.FqgrepError):
        F.read_input(str(tmp_path / "bad.fq.gz"))
    # BGZF followed by a plain gzip member: not pure BGZF, the serial multi-member decoder takes it
    open(str(tmp_path / "mixed.fq.gz"), "wb").write(bgzf(data[:5000], eof=False) + gzip.compress(data[5000:9000]))
    assert F.read_input(str(tmp_path / "mixed.fq.gz")).tobytes() == data[:9000]
    # empty BGZF file (only the EOF marker)
    open(str(tmp_path / "empty.fq.bgz"), "wb").write(bgzf(b""))
    assert F.read_input(str(tmp_path / "empty.fq.bgz")).size == 0


def test_truncated_gzip_is_an_error_wherever_the_cut_falls(tmp_path):
    """flate2's MultiGzDecoder raises UnexpectedEof when the input ends inside a member -- the first or a later one --
    and the reference turns that into a panic (exit 101).  Partial data must never come back as if it were the file."""
    a = fastq(["ACGT" * 30] * 300)
    b = fastq(["TTGA" * 25] * 400)
    za, zb = gzip.compress(a), gzip.compress(b)
    cases = {
        "first_member_cut.fq.gz": za[: len(za) // 2],
        "second_member_cut.fq.gz": za + zb[: len(zb) // 2],
        "second_member_header_only.fq.gz": za + zb[:5],
        "third_member_cut.fq.gz": za + zb + za[: len(za) - 3],
        "bgzf_cut.fq.gz": bgzf(a + b)[: len(bgzf(a + b)) // 2],            # no longer pure BGZF: the serial decoder sees the cut
        "bgzf_last_block_cut.fq.bgz": bgzf(a + b, eof=False)[:-9],
    }
    for name, z in cases.items():
        p = str(tmp_path / name)
        open(p, "wb").write(z)
        with pytest.raises(F.FqgrepError) as e:
            F.read_input(p)
        assert "truncated" in str(e.value) or "gzip" in str(e.value), name
        assert e.value.exit_code == 101, name
    # whole members are fine, with or without the empty BGZF EOF block
    open(str(tmp_path / "ok.fq.gz"), "wb").write(za + zb)
    assert F.read_input(str(tmp_path / "ok.fq.gz")).tobytes() == a + b
    open(str(tmp_path / "ok2.fq.bgz"), "wb").write(bgzf(a + b, eof=False))
    assert F.read_input(str(tmp_path / "ok2.fq.bgz")).tobytes() == a + b


def _read_all(path, cap, decompress=False):
    """fqg_reader_*: the incremental reader of the streaming path, `cap` bytes per call."""
    import ctypes as C
    from fqgrep_b200 import _ffi
    L = _ffi.lib()
    h = C.c_void_p()
    if L.fqg_reader_open(str(path).encode(), int(decompress), C.byref(h)) != 0:
        raise F.FqgrepError(-1, _ffi.last_error())
    out = bytearray()
    buf = np.zeros(cap, dtype=np.uint8)
    try:
        for _ in range(1 << 20):
            n, eof = C.c_size_t(0), C.c_int(0)
            rc = L.fqg_reader_read(h, buf.ctypes.data, cap, C.byref(n), C.byref(eof))
            if rc != 0:
                raise F.FqgrepError(rc, _ffi.last_error())
            assert n.value <= cap
            out += buf[: n.value].tobytes()
            if eof.value:
                break
            assert n.value > 0
        else:
            raise AssertionError("reader never reached eof")
    finally:
        L.fqg_reader_close(h)
    return bytes(out)


def test_incremental_reader_plain_gzip_bgzf_mixed_and_truncated(tmp_path):
    data = fastq(["ACGT" * 40, "TTTTGGGGCCCC" * 12, "N" * 7] * 4000)      # ~1.3 MB
    files = {"p.fq": data, "g.fq.gz": gzip.compress(data), "m.fq.gz": b"".join(gzip.compress(data[i:i + 70000]) for i in range(0, len(data), 70000)),
             "b.fq.bgz": bgzf(data), "b_noeof.fq.gz": bgzf(data, eof=False), "mixed.fq.gz": bgzf(data[:300000], eof=False) + gzip.compress(data[300000:]),
             "mixed2.fq.gz": gzip.compress(data[:300000]) + bgzf(data[300000:]), "e.fq": b"", "e.fq.gz": gzip.compress(b""), "z.dat": gzip.compress(data)}
    for name, z in files.items():
        p = tmp_path / name
        p.write_bytes(z)
        want = data if not name.startswith("e.") else b""
        for cap in (1000, 65536, 100_000, 1 << 20, 8 << 20):
            assert _read_all(p, cap, decompress=name.endswith(".dat")) == want, (name, cap)
    (tmp_path / "raw.dat").write_bytes(data)
    assert _read_all(tmp_path / "raw.dat", 4096) == data                   # no -Z: read as it is
    a, b = data[:400000], data[400000:]
    za, zb = gzip.compress(a), gzip.compress(b)
    for name, z in {"t1.fq.gz": za[: len(za) // 2], "t2.fq.gz": za + zb[: len(zb) // 2], "t3.fq.gz": za + zb[:5], "t4.fq.bgz": bgzf(data)[:-40],
                    "t5.fq.bgz": bgzf(data, eof=False)[: len(bgzf(data)) // 2], "t6.fq.gz": b"not gzip at all"}.items():
        p = tmp_path / name
        p.write_bytes(z)
        for cap in (70000, 4 << 20):
            with pytest.raises(F.FqgrepError):
                _read_all(p, cap)
    with pytest.raises(F.FqgrepError):
        _read_all(tmp_path / "missing.fq", 100)
