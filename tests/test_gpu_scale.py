"""Parity at benchmark scale (BASELINE.json configs[1] shape): millions of device-generated reads, full mask compared with
the oracle, and the three product paths (SoA kernel, fused FASTQ kernel on HBM-resident bytes, host streaming pipeline)
compared with each other bit for bit -- a size-independent cross-check that holds at any N."""
import ctypes as C

import numpy as np
import pytest

import fqgrep_b200 as F
from fqgrep_b200 import _ffi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

PATS = ["GACGAGATTA", "GACGTGATTA"]
RL = 150


def _gen(torch, n, seed, pseed, fastq):
    L = _ffi.lib()
    dev = torch.device("cuda:0")
    blob = "".join(PATS).encode()
    offs = np.array([0, 10, 20], dtype=np.uint32)
    rec = (2 * RL + 17) if fastq else RL
    t = torch.empty(n * rec + 64, dtype=torch.uint8, device=dev)
    rc = L.fqg_synth_reads_device(F.context(0), t.data_ptr(), n, 0, RL, fastq, seed, blob, offs.ctypes.data_as(C.c_void_p), 2, 0.01, pseed,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _ffi.last_error()
    torch.cuda.synchronize()
    return t


def test_device_generator_is_bit_identical_to_the_numpy_one():
    import torch
    from tests import synth
    n = 50_000
    for seed, pseed in ((1, 99), (2, 98)):
        mat = synth.bases(n, RL, seed=seed)
        synth.plant(mat, PATS, rate=0.01, seed=pseed)
        got = _gen(torch, n, seed, pseed, 0)[: n * RL].cpu().numpy()
        assert (got == mat.reshape(-1)).all()
        gotq = _gen(torch, n, seed, pseed, 1)[: n * (2 * RL + 17)].cpu().numpy()
        assert (gotq == synth.fastq(mat)).all()


def test_four_million_pairs_three_paths_and_oracle_agree():
    import torch
    L = _ffi.lib()
    ctx = F.context(0)
    n = 4_000_000
    m = F.MatcherFactory.new_matcher(None, False, PATS, None, F.MatcherOpts(False, True))
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    dev = torch.device("cuda:0")
    words = n // 32 + 2
    # path 1: SoA kernel, mate masks OR-ed (main.rs:749-755)
    masks = []
    offs = (torch.arange(n + 1, dtype=torch.int64, device=dev) * RL).to(torch.int32)
    m1 = torch.zeros(words, dtype=torch.int32, device=dev)
    m2 = torch.zeros(words, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    host_bases = []
    for (seed, pseed), out, orm in (((1, 99), m1, None), ((2, 98), m2, m1)):
        tb = _gen(torch, n, seed, pseed, 0)
        rc = L.fqg_match_bases_device(ctx, m._h, tb.data_ptr(), offs.data_ptr(), offs.data_ptr() + 4, n, out.data_ptr(),
                                      orm.data_ptr() if orm is not None else None, cnt.data_ptr() if orm is not None else None, RL, sp)
        assert rc == 0, _ffi.last_error()
        torch.cuda.synchronize()
        host_bases.append(tb[: n * RL].cpu().numpy())
        del tb
    soa_mask = m2.cpu().numpy().view(np.uint32)[: (n + 31) // 32].copy()
    soa_count = int(cnt.item())
    # oracle on the very same bytes, every bit
    o = O.Matcher(O.REGEX_SET, PATS, False, True)
    ho = np.arange(n + 1, dtype=np.uint64) * RL
    f1 = o.match_soa(host_bases[0], ho, threads=32)[1] != 0
    f2 = o.match_soa(host_bases[1], ho, threads=32)[1] != 0
    exp = f1 | f2
    got = np.unpackbits(soa_mask.view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (got == exp).all() and soa_count == int(exp.sum())
    del host_bases
    # path 2: fused FASTQ kernel on HBM-resident bytes; path 3: host streaming pipeline (pinned FASTQ)
    q1, q2 = _gen(torch, n, 1, 99, 1), _gen(torch, n, 2, 98, 1)
    nbytes = n * (2 * RL + 17)
    dm = torch.zeros(words, dtype=torch.int32, device=dev)
    res = _ffi.Result()
    rc = L.fqg_grep_fastq_device(ctx, m._h, F.PAIRED, q1.data_ptr(), nbytes, q2.data_ptr(), nbytes, dm.data_ptr(), words, C.byref(res), sp)
    assert rc == 0, _ffi.last_error()
    assert res.n_units == n and res.n_selected == soa_count
    assert (dm.cpu().numpy().view(np.uint32)[: (n + 31) // 32] == soa_mask).all()
    h1, h2 = q1[:nbytes].cpu().numpy(), q2[:nbytes].cpu().numpy()
    del q1, q2
    r = m.grep_fastq(h1, h2, mode=F.PAIRED)
    assert r["count"] == soa_count and (r["mask_words"][: (n + 31) // 32] == soa_mask).all()
    # idempotence: a second pass over the selected records selects all of them
    sel = F.write_selected(r["mask_words"], h1[: 317 * 200_000], h2[: 317 * 200_000], mode=F.PAIRED)
    il = np.frombuffer(sel, dtype=np.uint8)
    again = m.grep_fastq(il, mode=F.INTERLEAVED)
    assert again["count"] == again["units"] == int(exp[:200_000].sum())


def test_host_soa_entry_chunks_above_one_gib():
    """fqg_match_bases_host splits batches above 1 GiB of bases into several launches; masks must line up."""
    import torch
    n = 7_400_000                                   # 1.11 GB of bases -> two chunks
    m = F.Matcher(F.FIXED, ["GACGAGATTA"], F.MatcherOpts(False, True))
    tb = _gen(torch, n, 1, 99, 0)
    offs32 = (torch.arange(n + 1, dtype=torch.int64, device=tb.device) * RL).to(torch.int32)
    words = (n + 31) // 32 + 1
    dmask = torch.zeros(words, dtype=torch.int32, device=tb.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=tb.device)
    rc = _ffi.lib().fqg_match_bases_device(F.context(0), m._h, tb.data_ptr(), offs32.data_ptr(), offs32.data_ptr() + 4, n, dmask.data_ptr(), None,
                                           cnt.data_ptr(), RL, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _ffi.last_error()
    torch.cuda.synchronize()
    host = tb[: n * RL].cpu().numpy()
    bits, count = m.match_bases(host, np.arange(n + 1, dtype=np.uint64) * RL)
    exp = np.unpackbits(dmask.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    assert count == int(cnt.item()) and (bits == exp).all() and count > 0


def test_paired_files_with_different_read_lengths_stream_in_step():
    """R1 and R2 with the same record count but different byte sizes: the chunked host pipeline interleaves the two
    copies without the mates drifting apart (PairedFastqReader's job, src/lib/seq_io.rs:149-198)."""
    from tests import synth
    n = 260_000
    m1, m2 = synth.bases(n, 150, seed=1), synth.bases(n, 90, seed=2)
    synth.plant(m1, PATS, rate=0.01, seed=99)
    synth.plant(m2, PATS, rate=0.02, seed=98)
    f1, f2 = synth.fastq(m1), synth.fastq(m2)      # 82 MB and 51 MB: different numbers of 64 MB copy chunks
    m = F.MatcherFactory.new_matcher(None, False, PATS, None, F.MatcherOpts(True, True))
    g = m.grep_fastq(f1, f2, mode=F.PAIRED)
    e = O.Matcher(O.REGEX_SET, PATS, True, True).grep(f1, f2, mode=O.PAIRED, threads=16)
    assert g["count"] == e["count"] and (g["mask"] == (e["flags"] != 0)).all()


def test_one_million_read_names_full_scale():
    """configs[4]b at its stated size: a read-names file of 1,000,000 ids (matcher.rs:601-639) over 2 M device-generated
    records -- fused FASTQ kernel, SoA name kernel and oracle agree on every bit; ids of three lengths (11, 23 and 36
    bytes: the last two go past the 12 bytes a pool entry's first read covers)."""
    import torch
    L = _ffi.lib()
    ctx = F.context(0)
    dev = torch.device("cuda:0")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = 2_000_000
    rng = np.random.default_rng(11)
    chosen = np.sort(rng.choice(n, 600_000, replace=False))
    names = ["r%010d" % i for i in chosen] + ["r%010d" % (n + i) for i in range(400_000)]
    assert len(names) == 1_000_000
    m = F.Matcher(F.NAMES, names)
    fq = _gen(torch, n, 1, 99, 1)
    nbytes = n * (2 * RL + 17)
    words = n // 32 + 2
    dm = torch.zeros(words, dtype=torch.int32, device=dev)
    res = _ffi.Result()
    rc = L.fqg_grep_fastq_device(ctx, m._h, F.SINGLE, fq.data_ptr(), nbytes, None, 0, dm.data_ptr(), words, C.byref(res), sp)
    assert rc == 0, _ffi.last_error()
    want = np.zeros(n, dtype=bool)
    want[chosen] = True
    got = np.unpackbits(dm.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    assert res.n_units == n and res.n_selected == chosen.size and (got == want).all()
    host = fq[:nbytes].cpu().numpy()
    e = O.Matcher(O.NAMES, names, False, False).grep(host, threads=16)
    assert e["count"] == chosen.size and ((e["flags"] != 0) == want).all()
    # SoA name kernel over the header lines (ids back to back + u32 offsets)
    ids = np.frombuffer(("".join("r%010d" % i for i in range(200_000))).encode(), dtype=np.uint8)
    offs = np.arange(200_001, dtype=np.uint64) * 11
    bits, count = m.match_bases(ids, offs)
    assert count == int(want[:200_000].sum()) and (bits == want[:200_000]).all()
    # longer ids, with and without a comment after the id
    long_names = ["A00123:45:HXXXXDSXX:1:1101:%07d" % i for i in range(0, 300_000, 3)] + ["read_with_a_23_byte_%03d" % i for i in range(0, 1000, 2)]
    ml = F.Matcher(F.NAMES, long_names)
    heads = ["A00123:45:HXXXXDSXX:1:1101:%07d 1:N:0:ACGT" % i for i in range(40_000)] + ["read_with_a_23_byte_%03d" % i for i in range(1000)]
    from tests.fq_util import fastq
    data = fastq(["ACGT"] * len(heads), heads=heads)
    g = ml.grep_fastq(data)
    el = O.Matcher(O.NAMES, long_names, False, False).grep(data, threads=8)
    assert g["count"] == el["count"] == (40_000 + 2) // 3 + 500 and (g["mask"] == (el["flags"] != 0)).all()
