"""GPU parity of the fused FASTQ parse+match path (fqg_grep_fastq_host / _device) against the oracle and the
reference's end-to-end golden vectors (src/main.rs tests): counts, selection masks, ordered output bytes."""
import random

import numpy as np
import pytest

import fqgrep_b200 as F
from oracle import oracle as O
from tests import synth
from tests.fq_util import fastq, seqs_of

pytestmark = pytest.mark.gpu


def opts(inv=False, rc=False):
    return F.MatcherOpts(invert_match=inv, reverse_complement=rc)


def both(kind, pats, inv, rc, r1, r2=None, mode=F.SINGLE, check_output=True):
    m = F.Matcher(kind, pats, opts(inv, rc))
    o = O.Matcher(kind, pats, inv, rc)
    g = m.grep_fastq(r1, r2, mode=mode)
    e = o.grep(r1, r2, mode=mode, want_output=check_output, threads=4)
    assert g["units"] == e["units"]
    assert g["count"] == e["count"], (kind, pats, inv, rc, mode)
    assert (g["mask"] == (e["flags"] != 0)).all()
    if check_output:
        assert F.write_selected(g["mask_words"], r1, r2, mode=mode) == e["out"]
        # the same bytes from the device's record spans (K5): no second parse on the host
        w = m.grep_fastq_write(r1, r2, mode=mode)
        assert w["count"] == e["count"] and w["out"] == e["out"]
    return g


def test_kat_e2e_counts(kat):
    e = kat["e2e_counts"]
    for paired, pats, exp in e["cases"]:
        m = F.MatcherFactory.new_matcher(None, False, pats)
        total = 0
        if paired:
            for i in range(0, 4, 2):
                total += m.grep_fastq(fastq(e["files"][i]), fastq(e["files"][i + 1]), mode=F.PAIRED)["count"]
        else:
            for f in e["files"]:
                total += m.grep_fastq(fastq(f))["count"]
        assert total == exp, (paired, pats)


def test_kat_e2e_order_and_split(kat):
    e = kat["e2e_order"]
    for paired, pats, exp in e["cases"]:
        m = F.MatcherFactory.new_matcher(None, False, pats)
        if paired:
            r1, r2 = fastq(e["files"][0]), fastq(e["files"][1])
            out = F.write_selected(m.grep_fastq(r1, r2, mode=F.PAIRED)["mask_words"], r1, r2, mode=F.PAIRED)
        else:
            out = b"".join(F.write_selected(m.grep_fastq(fastq(f))["mask_words"], fastq(f)) for f in e["files"])
        assert seqs_of(out) == exp, (paired, pats)
    s = kat["e2e_split_output"]
    r1, r2 = fastq(s["r1"]), fastq(s["r2"])
    g = F.Matcher(F.REGEX_SET, [s["pattern"]]).grep_fastq(r1, r2, mode=F.PAIRED)
    assert g["count"] == s["count"]
    o1, o2 = F.write_selected(g["mask_words"], r1, r2, mode=F.PAIRED, split=True)
    assert seqs_of(o1) == s["out_r1"] and seqs_of(o2) == s["out_r2"]


def test_kat_rc_invert_paired(kat):
    e = kat["e2e_rc_invert_paired"]
    for paired, inv, rc, exp in e["cases"]:
        m = F.Matcher(F.REGEX_SET, [e["pattern"]], opts(inv, rc))
        if paired:
            got = m.grep_fastq(fastq(e["files"][0]), fastq(e["files"][1]), mode=F.PAIRED)["count"]
        else:
            got = sum(m.grep_fastq(fastq(f))["count"] for f in e["files"])
        assert got == exp, (paired, inv, rc)


def test_kat_names(kat):
    for inv, names, rid, exp in kat["query_name"]["cases"]:
        m = F.MatcherFactory.new_query_name_matcher(names, opts(inv))
        assert m.read_match("ACGT", head=rid + " description") == exp
    hc = kat["query_name"]["header_case"]
    assert F.MatcherFactory.new_query_name_matcher(hc["names"]).read_match("ACGT", head=hc["header"]) == hc["expect"]
    e = kat["e2e_names"]
    for inv, names, exp in e["single"]["cases"]:
        assert F.Matcher(F.NAMES, names, opts(inv)).grep_fastq(fastq(["ACGT"] * e["single"]["n_reads"]))["count"] == exp
    n = e["paired"]["n_reads_per_file"]
    for inv, names, exp in e["paired"]["cases"]:
        assert F.Matcher(F.NAMES, names, opts(inv)).grep_fastq(fastq(["ACGT"] * n), fastq(["TTTT"] * n), mode=F.PAIRED)["count"] == exp
    for inv, names, exp in e["empty"]["cases"]:
        assert F.Matcher(F.NAMES, names, opts(inv)).grep_fastq(fastq(["ACGT"] * e["empty"]["n_reads"]))["count"] == exp


def test_kat_exit_codes(kat):
    e = kat["e2e_exit"]
    for c in e["cases"]:
        try:
            if c["fixed"]:
                F.validate_fixed_pattern(c["pattern"])
            m = F.MatcherFactory.new_matcher(c["pattern"], c["fixed"], [])
            code = 0 if sum(m.grep_fastq(fastq(f))["count"] for f in e["files"]) else 1
        except F.FqgrepError as ex:
            code = ex.exit_code
        assert code == c["exit"], c


def test_kat_protein(kat):
    e = kat["e2e_protein"]
    r = fastq(e["reads"])
    for fixed, pats, exp in e["cases"]:
        m = F.MatcherFactory.new_matcher(None, fixed, pats)
        assert seqs_of(F.write_selected(m.grep_fastq(r)["mask_words"], r)) == exp


@pytest.mark.parametrize("kind,pats,inv,rc", [
    (F.REGEX, ["GACGAGATTA"], False, False),
    (F.REGEX_SET, ["GACGAGATTA", "GACGTGATTA"], False, True),
    (F.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"], True, False),
    (F.FIXED_SET, ["GACGAGATTA", "TTTTTTTT", "ACGTAC"], True, True),
])
def test_synthetic_fastq_single_paired_interleaved(kind, pats, inv, rc):
    n = 120_000     # ~38 MB per mate: thousands of tiles, several look-back windows
    m1, m2 = synth.bases(n, seed=1), synth.bases(n, seed=2)
    synth.plant(m1, ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"], rate=0.01, seed=99)
    synth.plant(m2, ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"], rate=0.01, seed=98)
    f1, f2 = synth.fastq(m1), synth.fastq(m2)
    g = both(kind, pats, inv, rc, f1, check_output=False)
    assert 0 < g["count"] < n
    both(kind, pats, inv, rc, f1, f2, mode=F.PAIRED, check_output=False)
    # interleave the two mates record by record
    il = np.stack([f1.reshape(n, 317), f2.reshape(n, 317)], axis=1).reshape(-1)
    both(kind, pats, inv, rc, il, mode=F.INTERLEAVED, check_output=False)


def test_output_bytes_match_oracle_on_medium_file():
    n = 30_000
    m1 = synth.bases(n, seed=4)
    synth.plant(m1, ["GACGAGATTA"], rate=0.02)
    f1 = synth.fastq(m1)
    both(F.FIXED, ["GACGAGATTA"], False, True, f1)


def test_names_one_million_ids_scaled_down():
    n = 200_000
    f1 = synth.fastq(synth.bases(n, seed=1))
    rng = np.random.default_rng(5)
    chosen = rng.choice(n, 20_000, replace=False)
    names = ["r%010d" % i for i in chosen] + ["r%010d" % (n + i) for i in range(2_000)]
    for inv in (False, True):
        g = both(F.NAMES, names, inv, False, f1, check_output=False)
        assert g["count"] == (n - 20_000 if inv else 20_000)
    f2 = synth.fastq(synth.bases(n, seed=2))
    both(F.NAMES, names, False, False, f1, f2, mode=F.PAIRED, check_output=False)


def test_large_pattern_set_through_fastq_path():
    """A tier-3 matcher (q-gram filter in the SoA kernel) runs its exact tier-2 automaton inside the fused FASTQ kernel."""
    rng = np.random.default_rng(17)
    pats = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(5000)]
    n = 40_000
    mat = synth.bases(n, seed=6)
    synth.plant(mat, pats[:100], rate=0.03)
    f1 = synth.fastq(mat)
    assert F.Matcher(F.FIXED_SET, pats, opts(False, True)).dfa_info().tier == 3
    g = both(F.FIXED_SET, pats, False, True, f1, check_output=False)
    assert 0 < g["count"] < n


def test_format_edge_cases():
    seqs = ["ACGT", "", "TTACGG", "N", "ACGTACGTAC" * 30]
    for crlf in (False, True):
        for plus_name in (False, True):
            for final_nl in (False, True):
                data = fastq(seqs, crlf=crlf, plus_name=plus_name, final_newline=final_nl)
                for pats in (["ACG"], ["T$"], ["^$"], ["^A", "C$"]):
                    both(F.REGEX_SET, pats, False, True, data)
                both(F.NAMES, ["@0", "@3"], False, False, data)
    both(F.REGEX, ["A"], False, False, fastq(seqs) + b"\n\n\r\n")       # trailing blank lines
    # a LAST record with empty sequence and quality: the newline that ends its '+' line must survive the trimming of
    # trailing terminators (seq_io accepts "...+\n<EOF>"), in one stream or in both
    for tail in (b"@e\n\n+\n", b"@e\n\n+\n\n", b"@e\r\n\r\n+\r\n", b"@e\n\n+\n\n\n\n"):
        both(F.REGEX_SET, ["^$", "A"], False, False, fastq(seqs) + tail)
        both(F.REGEX_SET, ["^$", "A"], False, False, fastq(seqs, heads=["p%d" % i for i in range(5)]) + tail,
             fastq(seqs, heads=["p%d" % i for i in range(5)]) + b"@e\nAC\n+\nII\n", mode=F.PAIRED)
        both(F.REGEX_SET, ["^$", "A"], False, False, fastq(seqs) + tail, fastq(seqs) + tail, mode=F.PAIRED)
    with pytest.raises(F.FqgrepError):
        F.Matcher(F.REGEX, ["A"]).grep_fastq(fastq(seqs) + b"@e\n\n+")         # no newline after '+': truncated
    assert F.Matcher(F.REGEX, ["A"]).grep_fastq(b"")["count"] == 0
    # ids end at the first space only (a tab belongs to the id)
    both(F.NAMES, ["a\tb"], False, False, fastq(["ACGT"], heads=["a\tb c"]))
    both(F.NAMES, ["a"], False, False, fastq(["ACGT"], heads=["a b"]))


def test_ragged_and_long_records_cross_tiles():
    rng = random.Random(9)
    seqs = []
    for i in range(3000):
        L = rng.choice([0, 1, 5, 50, 150, 151, 300, 1000, 2500, 6000]) if i % 7 else rng.choice([20000, 40000])
        seqs.append("".join(rng.choice("ACGT") for _ in range(L)))
    seqs[10] = seqs[10] + "GACGAGATTA"
    data = fastq(seqs)
    for kind, pats, inv, rc in ((F.FIXED, ["GACGAGATTA"], False, True), (F.REGEX_SET, ["^ACG", "TT$"], True, False)):
        both(kind, pats, inv, rc, data)
    both(F.NAMES, ["@5", "@2999", "@0"], False, False, data)


def test_tiny_records_like_reference_fixtures():
    seqs = ["AAAA", "TTTT", "GGGG", "CCCC"] * 5000      # 16-byte records, 1000+ records per tile
    data = fastq(seqs, heads=["%d" % (i % 10) for i in range(len(seqs))])
    both(F.REGEX_SET, ["^A", "GG"], False, False, data)
    both(F.REGEX_SET, ["^A", "GG"], False, False, data, data, mode=F.PAIRED)


def test_line_phase_guess_wrong_and_ambiguous_in_multi_tile_files():
    """The fused kernel guesses a tile's line phase from its first two records and confirms the guess a tile later.
    (a) One line dropped in the middle of a file of many tiles: every tile behind it guesses a phase that the line count
        contradicts -- those tiles are walked again from global memory; the file is an error, like in the oracle, and the
        part in front of the damage alone still gives the oracle's answer.
    (b) Records whose lines 2 and 4 could be lines 1 and 3 ("@..", "@..", "+..", "+.."): two phases look right, no guess is
        made; results must not depend on it.  Once with ~24 records per tile, once with ~40 (more than one per lane)."""
    rng = random.Random(5)
    seqs = ["".join(rng.choice("ACGT") for _ in range(150)) for _ in range(6000)]      # ~1.9 MB, ~190 tiles
    seqs[4000] = seqs[4000][:50] + "GACGAGATTA" + seqs[4000][60:]
    good = fastq(seqs)
    lines = good.split(b"\n")
    m = F.Matcher(F.FIXED, ["GACGAGATTA"], opts(False, True))
    for drop in (4 * 3000 + 1, 4 * 3000 + 3, 4 * 17):      # a sequence line, a quality line, a header
        bad = b"\n".join(lines[:drop] + lines[drop + 1:])
        with pytest.raises(O.OracleError):
            O.Matcher(O.FIXED, ["GACGAGATTA"], False, True).grep(bad)
        with pytest.raises(F.FqgrepError) as e:
            m.grep_fastq(bad)
        assert e.value.exit_code == 101
    both(F.FIXED, ["GACGAGATTA"], False, True, good)
    for width in (100, 60):
        recs = []
        for i in range(5000):
            recs.append(b"@" + (b"%d" % i).ljust(width, b"x") + b"\n@" + b"AC"[i % 2:i % 2 + 1] * width + b"\n+" + b"z" * width + b"\n+" + b"I" * width + b"\n")
        amb = b"".join(recs)
        both(F.REGEX_SET, ["^@A", "C{5}$"], False, False, amb)
        both(F.NAMES, ["%d" % i + "x" * (width - len("%d" % i)) for i in (0, 7, 4999)], False, False, amb)


def test_malformed_fastq_is_an_error_like_the_reference_panic():
    m = F.Matcher(F.FIXED, ["AC"])
    good = fastq(["ACGT"] * 100)
    for bad in (b"ACGT\n" + good, good + b"@a\nACGT\n-\nXXXX\n", good + b"@a\nACGT\n+\nXXX\n", good + b"@a\nACGT\n+\n", good + b"@a\nACGT\n"):
        with pytest.raises(F.FqgrepError) as e:
            m.grep_fastq(bad)
        assert e.value.exit_code == 101
        with pytest.raises(O.OracleError):
            O.Matcher(O.FIXED, ["AC"]).grep(bad)
    with pytest.raises(F.FqgrepError):
        m.grep_fastq(fastq(["ACGT"] * 3), mode=F.INTERLEAVED)
    with pytest.raises(F.FqgrepError) as e:
        m.grep_fastq(fastq(["ACGT"] * 3), fastq(["ACGT"] * 2), mode=F.PAIRED)
    assert e.value.exit_code == 101
    with pytest.raises(F.FqgrepError) as e:     # mismatching ids: assert_eq! panic in the reference (main.rs:741-747)
        m.grep_fastq(fastq(["ACGT"], heads=["a 1"]), fastq(["ACGT"], heads=["b 1"]), mode=F.PAIRED)
    assert e.value.exit_code == 101
    assert m.grep_fastq(fastq(["ACGT"], heads=["a 1"]), fastq(["GGGG"], heads=["a 2"]), mode=F.PAIRED)["count"] == 1


def test_device_resident_fastq_entry():
    import ctypes as C
    import torch
    from fqgrep_b200 import _ffi
    n = 150_000
    m1, m2 = synth.bases(n, seed=1), synth.bases(n, seed=2)
    synth.plant(m1, ["GACGAGATTA"], rate=0.01, seed=99)
    f1, f2 = synth.fastq(m1), synth.fastq(m2)
    pats = ["GACGAGATTA", "GACGTGATTA"]
    m = F.Matcher(F.REGEX_SET, pats, opts(False, True))
    dev = torch.device("cuda:0")
    bufs = []
    for f in (f1, f2):
        t = torch.zeros(f.size + 64, dtype=torch.uint8, device=dev)
        t[: f.size] = torch.from_numpy(f).to(dev)
        bufs.append(t)
    words = n // 32 + 2
    mask = torch.zeros(words, dtype=torch.int32, device=dev)
    res = _ffi.Result()
    torch.cuda.synchronize()
    rc = _ffi.lib().fqg_grep_fastq_device(F.context(0), m._h, F.PAIRED, bufs[0].data_ptr(), f1.size, bufs[1].data_ptr(), f2.size, mask.data_ptr(), words,
                                          C.byref(res), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _ffi.last_error()
    e = O.Matcher(O.REGEX_SET, pats, False, True).grep(f1, f2, mode=O.PAIRED, threads=8)
    assert res.n_selected == e["count"] and res.n_units == n
    got = np.unpackbits(mask.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (got == (e["flags"] != 0)).all()


def test_k0_index_and_compaction_then_soa_match():
    """Parse once, match many: fqg_index_fastq_device + fqg_compact_bases_device give the SoA layout of the same reads;
    the SoA kernel over it must select exactly what the fused kernel selects."""
    import ctypes as C
    import torch
    from fqgrep_b200 import _ffi
    L = _ffi.lib()
    ctx = F.context(0)
    dev = torch.device("cuda:0")
    rng = random.Random(3)
    seqs = ["".join(rng.choice("ACGT") for _ in range(rng.choice([0, 1, 50, 150, 151, 300, 5000]))) for _ in range(20_000)]
    seqs[5] = "GACGAGATTA" + seqs[5]
    data = np.frombuffer(fastq(seqs, crlf=True, plus_name=True, final_newline=False), dtype=np.uint8)
    tb = torch.zeros(data.size + 64, dtype=torch.uint8, device=dev)
    tb[: data.size] = torch.from_numpy(data.copy()).to(dev)
    cap = len(seqs) + 10
    d_start = torch.zeros(cap, dtype=torch.int64, device=dev)
    d_len = torch.zeros(cap, dtype=torch.int32, device=dev)
    n_rec = C.c_uint64(0)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.fqg_index_fastq_device(ctx, tb.data_ptr(), data.size, d_start.data_ptr(), d_len.data_ptr(), cap, C.byref(n_rec), sp)
    assert rc == 0, _ffi.last_error()
    assert n_rec.value == len(seqs)
    starts, lens = d_start.cpu().numpy()[: len(seqs)], d_len.cpu().numpy()[: len(seqs)]
    assert [int(x) for x in lens] == [len(s) for s in seqs]
    for i in (0, 5, 777, len(seqs) - 1):
        assert bytes(data[starts[i]: starts[i] + lens[i]]).decode() == seqs[i]
    total = int(lens.sum())
    d_bases = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
    d_offs = torch.zeros(len(seqs) + 1, dtype=torch.int32, device=dev)
    rc = L.fqg_compact_bases_device(ctx, tb.data_ptr(), d_start.data_ptr(), d_len.data_ptr(), len(seqs), d_bases.data_ptr(), d_offs.data_ptr(), sp)
    assert rc == 0, _ffi.last_error()
    torch.cuda.synchronize()
    offs = d_offs.cpu().numpy().astype(np.int64)
    assert offs[0] == 0 and offs[-1] == total and (np.diff(offs) == lens).all()
    assert d_bases[:total].cpu().numpy().tobytes() == "".join(seqs).encode()
    pats = ["GACGAGATTA", "^ACG", "TT$"]
    m = F.Matcher(F.REGEX_SET, pats, opts(False, True))
    words = len(seqs) // 32 + 2
    mask = torch.zeros(words, dtype=torch.int32, device=dev)
    rc = L.fqg_match_bases_device(ctx, m._h, d_bases.data_ptr(), d_offs.data_ptr(), d_offs.data_ptr() + 4, len(seqs), mask.data_ptr(), None, None, 0, sp)
    assert rc == 0, _ffi.last_error()
    torch.cuda.synchronize()
    fused = m.grep_fastq(data)
    got = np.unpackbits(mask.cpu().numpy().view(np.uint8), bitorder="little")[: len(seqs)].astype(bool)
    assert (got == fused["mask"]).all() and fused["count"] > 0
    # malformed input is reported by the indexer too
    bad = np.frombuffer(b"@a\nACGT\n+\nXXX\n", dtype=np.uint8)
    tb2 = torch.zeros(bad.size + 64, dtype=torch.uint8, device=dev)
    tb2[: bad.size] = torch.from_numpy(bad.copy()).to(dev)
    assert L.fqg_index_fastq_device(ctx, tb2.data_ptr(), bad.size, d_start.data_ptr(), d_len.data_ptr(), cap, C.byref(n_rec), sp) == -4


def _random_fastq(rng, n_records, pair_heads=None):
    """A valid FASTQ file with everything the format allows mixed in: empty and long lines, CRLF (whole file), '+name'
    separators, quality lines that start with '@' or '+', headers with spaces and tabs, optional final newline."""
    nl = b"\r\n" if rng.random() < 0.3 else b"\n"
    recs, heads = [], []
    for i in range(n_records):
        L = rng.choice((0, 1, 2, 3, 4, 7, 33, 64, 150, 150, 150, 151, 299, 2000)) if rng.random() < 0.97 else rng.randrange(15000, 25000)
        seq = bytes(rng.choice(b"ACGTNacgt") for _ in range(L))
        qual = bytes(rng.choice(b"@+I#5?") for _ in range(L))
        if pair_heads is not None:
            head = pair_heads[i]
        else:
            head = ("r%d" % i + rng.choice(("", " 1:N:0", "\tx", " a b", "/1"))).encode()
        heads.append(head)
        plus = b"+" + (head if rng.random() < 0.2 else b"")
        recs.append(b"@" + head + nl + seq + nl + plus + nl + qual + nl)
    data = b"".join(recs)
    if rng.random() < 0.4 and data:
        data = data[:-len(nl)]
    return data, heads


def _dump_failing_input(tag, *blobs):
    """Keeps the inputs of a failing fuzz case where a gpurun call brings them back (gpurun_out/)."""
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        for k, b in enumerate(blobs):
            if b is not None:
                open(os.path.join(d, "fuzz_fail_%s_%d.bin" % (tag, k)), "wb").write(bytes(b))


def test_fuzz_valid_and_damaged_fastq_against_the_oracle():
    """Randomised files through the fused kernel: valid ones must give the oracle's units / count / mask / output bytes in
    all three pairing modes; damaged ones (a byte range cut out, a line dropped, a quality line shortened, truncation)
    must be an error exactly when the oracle says so, and agree with it otherwise."""
    import os
    rng = random.Random(int(os.environ.get("FQG_FUZZ_SEED", "20240611")))       # longer campaigns: FQG_FUZZ_ITERS / FQG_FUZZ_SEED
    iters = int(os.environ.get("FQG_FUZZ_ITERS", "40"))
    matchers = [(F.REGEX_SET, ["ACG", "T{3,}", "^N", "G$"], False, True), (F.FIXED, ["acgt"], True, False), (F.NAMES, ["r1", "r7", "r64", "r501"], False, False)]
    n_err = n_ok = 0
    for it in range(iters):
        n = rng.choice((1, 2, 3, 17, 64, 200, 900))
        data, heads = _random_fastq(rng, n)
        kind, pats, inv, rc = matchers[it % len(matchers)]
        try:
            both(kind, pats, inv, rc, data)
        except Exception:
            _dump_failing_input("single_it%d" % it, data)
            raise
        if n % 2 == 0:
            ids = [heads[i // 2 * 2].split(b" ")[0] for i in range(n)]
            inter, _ = _random_fastq(rng, n, pair_heads=[ids[i] + rng.choice((b"", b" x")) for i in range(n)])
            try:
                both(kind, pats, inv, rc, inter, mode=F.INTERLEAVED)
            except Exception:
                _dump_failing_input("inter_it%d" % it, inter)
                raise
        mate, _ = _random_fastq(rng, n, pair_heads=[h.split(b" ")[0] + b" 2" for h in heads])
        try:
            both(kind, pats, inv, rc, data, mate, mode=F.PAIRED)
        except Exception:
            _dump_failing_input("paired_it%d" % it, data, mate)
            raise
        # damage
        for _ in range(4):
            bad = bytearray(data)
            how = rng.randrange(6)
            if how == 0 and len(bad) > 2:
                a = rng.randrange(len(bad) - 1); del bad[a:a + rng.randrange(1, min(40, len(bad) - a))]
            elif how == 1 and bad.count(b"\n") > 1:
                lines = bytes(bad).split(b"\n"); del lines[rng.randrange(len(lines) - 1)]; bad = bytearray(b"\n".join(lines))
            elif how == 2 and len(bad) > 1:
                bad = bad[:rng.randrange(1, len(bad))]
            elif how == 4:      # harmless: another letter inside some line, blank lines at the end
                spots = [k for k in range(len(bad)) if bad[k] in b"ACGTNacgtI#5?"]
                if spots:
                    bad[rng.choice(spots)] = ord("A")
                bad += b"\n" * rng.randrange(0, 3)
            elif how == 5:
                bad += rng.choice((b"\n", b"\r\n", b"\n\n\n"))
            else:
                a = rng.randrange(len(bad) + 1); bad[a:a] = rng.choice((b"\n", b"@", b"+\n", b"\r", b"X"))
            bad = bytes(bad)
            m = F.Matcher(kind, pats, opts(inv, rc))
            o = O.Matcher(kind, pats, inv, rc)
            try:
                e = o.grep(bad)
            except O.OracleError:
                e = None
            try:
                if e is None:
                    with pytest.raises(F.FqgrepError) as ex:
                        m.grep_fastq(bad)
                    assert ex.value.exit_code == 101
                    n_err += 1
                else:
                    g = m.grep_fastq(bad)
                    assert g["units"] == e["units"] and g["count"] == e["count"] and (g["mask"] == (e["flags"] != 0)).all()
                    n_ok += 1
                # the same damaged file as one mate of a pair and as an interleaved file: error or not, like the oracle
                for mode, second in ((F.PAIRED, mate), (F.INTERLEAVED, None)):
                    try:
                        ep = o.grep(bad, second, mode=mode)
                    except O.OracleError:
                        ep = None
                    if ep is None:
                        with pytest.raises(F.FqgrepError) as ex:
                            m.grep_fastq(bad, second, mode=mode)
                        assert ex.value.exit_code == 101
                    else:
                        gp = m.grep_fastq(bad, second, mode=mode)
                        assert gp["units"] == ep["units"] and gp["count"] == ep["count"] and (gp["mask"] == (ep["flags"] != 0)).all()
            except BaseException:
                _dump_failing_input("damaged_it%d_%s" % (it, "oracle_err" if e is None else "oracle_ok"), bad, mate)
                raise
    assert n_err > 20 and n_ok > 20     # the damage produced both kinds


def test_anchored_barcode_set_through_the_fastq_path():
    """The fused kernel shares the table walk with the SoA kernel, including the stop at settled states (dead state of a
    ^-anchored set, tiers 1 / 2 / 4)."""
    rng = random.Random(43)
    codes = ["".join(rng.choice("ACGT") for _ in range(rng.randint(8, 14))) for _ in range(1500)]
    pats = ["^" + c for c in codes]
    assert F.Matcher(F.REGEX_SET, pats).dfa_info().tier in (1, 2, 4)
    seqs = []
    for i in range(20000):
        s = "".join(rng.choice("ACGTN" if i % 5 == 0 else "ACGT") for _ in range(rng.choice([0, 5, 36, 150, 151])))
        if i % 3 == 0:
            s = codes[rng.randrange(len(codes))] + s
        seqs.append(s)
    data = fastq(seqs)
    for inv in (False, True):
        g = both(F.REGEX_SET, pats, inv, False, data, check_output=False)
        assert 0 < g["count"] < len(seqs)
    both(F.REGEX_SET, pats, False, False, data, fastq([s[::-1] for s in seqs]), mode=F.PAIRED, check_output=False)
