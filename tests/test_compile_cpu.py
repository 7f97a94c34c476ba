"""CPU tests of the product's host compilers (regex -> DFA, Aho-Corasick -> DFA) through the C ABI:
the exported tables are walked by a test-side simulator and compared with the oracle, the reference's
golden vectors, and Python's `re` as a third, independent implementation.  No compute call needs a GPU."""
import random
import re

import numpy as np
import pytest

import fqgrep_b200 as F
from oracle import oracle as O
from tests import synth
from tests.dfa_sim import run_dfa, run_dfa_soa


def opts(inv, rc):
    return F.MatcherOpts(invert_match=inv, reverse_complement=rc)


def test_library_exports_every_declared_symbol():
    import os
    from fqgrep_b200 import _ffi
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "fqgrep_b200.h")).read()
    declared = set(re.findall(r"\b(fqg_[a-z_0-9]+)\s*\(", hdr))
    L = _ffi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_ffi.SYMBOLS)
    assert b"sm_100a" in L.fqg_version()
    # the reference-side bindings shown in INTEGRATION.md (Rust shim, source only) and the C++ host header name them all too
    root = os.path.join(os.path.dirname(__file__), "..")
    rust = open(os.path.join(root, "rust", "fqgrep-b200", "src", "lib.rs")).read()
    assert not [n for n in declared if n not in rust]
    hpp = open(os.path.join(root, "include", "fqgrep_b200.hpp")).read()
    for n in ("fqg_matcher_new", "fqg_grep_fastq_host", "fqg_match_bases_host", "fqg_stream_open", "fqg_stream_next", "fqg_fastq_cut", "fqg_write_spans",
              "fqg_reader_open", "fqg_iupac_expand"):
        assert n in hpp


def test_kat_fixed(kat):
    for rc, pat, seq, exp in kat["fixed_string"]["cases"]:
        for inv in (False, True):
            m = F.MatcherFactory.new_matcher(pat, True, [], None, opts(inv, rc))
            assert run_dfa(m, [seq])[0] == (exp != inv)
    for pat, seq, fwd, rcv in kat["rc_special"]["cases"]:
        assert run_dfa(F.Matcher(F.FIXED, [pat], opts(False, False)), [seq])[0] == fwd
        assert run_dfa(F.Matcher(F.FIXED, [pat], opts(False, True)), [seq])[0] == rcv


def test_kat_fixed_set(kat):
    for rc, pats, seq, exp in kat["fixed_string_set"]["cases"]:
        for inv in (False, True):
            m = F.MatcherFactory.new_matcher(None, True, pats, None, opts(inv, rc))
            assert run_dfa(m, [seq])[0] == (exp != inv)


def test_kat_regex(kat):
    for rc, pat, seq, exp in kat["regex"]["cases"]:
        for inv in (False, True):
            m = F.MatcherFactory.new_matcher(pat, False, [], None, opts(inv, rc))
            assert run_dfa(m, [seq])[0] == (exp != inv), (rc, pat, seq, inv)


def test_kat_regex_set(kat):
    for rc, pats, seq, exp in kat["regex_set"]["cases"]:
        for inv in (False, True):
            m = F.MatcherFactory.new_matcher(None, False, pats, None, opts(inv, rc))
            assert run_dfa(m, [seq])[0] == (exp != inv), (rc, pats, seq, inv)


def test_kat_validate_fixed(kat):
    v = kat["validate_fixed"]
    for p, prot in v["ok"]:
        F.validate_fixed_pattern(p, prot)
    for p, prot, msg in v["err"]:
        with pytest.raises(F.FqgrepError) as e:
            F.validate_fixed_pattern(p, prot)
        assert e.value.message == msg and e.value.exit_code == 2
    for p, prot, sub in v["err_contains"]:
        with pytest.raises(F.FqgrepError) as e:
            F.validate_fixed_pattern(p, prot)
        assert sub in e.value.message


def test_invalid_regex_message():
    with pytest.raises(F.FqgrepError) as e:
        F.MatcherFactory.new_matcher("*", False, [])
    assert e.value.message.startswith("Invalid regular expression: '*'") and e.value.exit_code == 2
    for bad in ["(", "a)", "[a", "a{2", "a{3,2}", r"\1", "(?=a)", "(?<!a)b", r"\p{Greek}", r"\q", "a{", "(?P<1>a)", "(?z)"]:
        with pytest.raises(F.FqgrepError):
            F.Matcher(F.REGEX, [bad])
        with pytest.raises(O.OracleError):
            O.Matcher(O.REGEX, [bad])


def test_crlf_flag_is_refused_not_half_honoured():
    """(?R) would make \\r\\n one line terminator for `.` AND for ^ / $ under (?m); only the former was ever modelled, so
    turning it on is a construction error (exit 2) rather than a silently different language.  Turning it off is a no-op."""
    for bad in ["(?R)A$", "(?Rm)A$", "(?m:(?R)^A)", "A(?R:.)C"]:
        with pytest.raises(F.FqgrepError) as e:
            F.Matcher(F.REGEX, [bad])
        assert e.value.exit_code == 2 and e.value.message.startswith("Invalid regular expression")
    assert F.Matcher(F.REGEX, ["(?-R)A.C"]).dfa_info().n_states > 0


BASELINE_REGEXES = ["GACGAGATTA", "GACGTGATTA", "AC[GT]{3}TT(A|G)CA"]

SYNTAX_CASES = [
    r"A.A", r"^A", r"T$", r"^ACGT$", r"^$", r"", r"A|", r"(|A)C", r"A*", r"A+C", r"AC?G", r"A{3}", r"A{2,}", r"A{1,3}C", r"(AC){2}G",
    r"[ACG]T", r"[^A]C", r"[A-C]G", r"[^A-C]", r"[[:alpha:]]N", r"[[:^upper:]]", r"\dA", r"\w+", r"\W", r"\s", r"\S{4}", r"A\b", r"\bA", r"\BA",
    r"(?i)acgt", r"(?i:a)C", r"a(?i)c", r"(?x) A C  # comment", r"(?s).", r".", r"\x41C", r"\x{41}C", r"\u0041", r"A\.C", r"\.", r"\\",
    r"[A&&[^C]]", r"[A-G&&[^C-F]]", r"[A-Z--[C-X]]", r"[AC~~CG]", r"[\]A]", r"[]A]", r"[A\-C]", r"[-A]", r"[A-]", r"(?P<n>AC)G", r"(?<n>AC)G", r"(?:AC)+G",
    r"A*?C", r"A+?C", r"A??C", r"A{2,3}?C", r"(A|C|G)(A|T)", r"(A*)*C", r"(A|AA)+$", r"^(A|C)*$", r"\AAC", r"AC\z", r"(?m)^A", r"(?m)A$",
    r"\b{start}A", r"A\b{end}", r"\<A", r"A\>", r"\b{start-half}A", r"A\b{end-half}", r"[\d\s]", r"[^\d]", r"[\w&&[^_]]", r"(?-u:\xFF)", r"(?-u)[^A]",
    r"é", r"[é]", r"[^é]", r"[α-ω]", r"\u{1F600}", r"(?i)k", r"(?i)s", r"(?i)[k]", r"}", r"]", r"a{,3}" if False else r"A{0}C", r"(){3}", r"^*A",
]


def _haystacks(rng, n=400):
    alph = ["ACGT", "ACGTN", "ACGTacgt", "AC", "ACGT_ 09\n\t", "ACGTkKsS"]
    hs = ["", "A", "C", "AC", "ACGT", "AAAA", "N"]
    for _ in range(n):
        a = rng.choice(alph)
        hs.append("".join(rng.choice(a) for _ in range(rng.randint(0, 24))))
    return hs


def test_regex_dfa_vs_oracle_syntax_surface():
    rng = random.Random(1)
    hs = _haystacks(rng) + ["é", "aé", "\u212a", "\u017f", "αβγ", "😀", "A\u00e9C"]
    hb = [h.encode("utf-8") for h in hs] + [b"\xff", b"A\xffC", b"\xc3", b"\xe2\x84"]
    for pat in SYNTAX_CASES:
        for rc in (False, True):
            m = F.Matcher(F.REGEX, [pat], opts(False, rc))
            o = O.Matcher(O.REGEX, [pat], False, rc)
            got = run_dfa(m, hb)
            exp = [o.read_match(h) for h in hb]
            bad = [h for h, g, e in zip(hb, got, exp) if g != e]
            assert not bad, (pat, rc, bad[:5])


# patterns whose meaning is the same in Python `re` (bytes) and regex-syntax on ASCII haystacks
PY_SAFE = [r"A.A", r"^A", r"^ACGT", r"A*C", r"A+C", r"AC?G", r"A{3}", r"A{2,}", r"A{1,3}C", r"(AC){2}G", r"[ACG]T", r"[^A]C", r"[A-C]G",
           r"\dA", r"\w+", r"\W", r"\s", r"A\b", r"\bA", r"\BA", r"(?i)acgt", r"(A|C|G)(A|T)", r"(A*)*C", r"AC[GT]{3}TT(A|G)CA", r"G.{3,5}C",
           r"(?:AC)+G", r"A|C{2}|GT", r"[^ACGT]", r"\AAC", r"T.....G", r"C+.T", r"A.+G"]


def test_regex_vs_python_re():
    rng = random.Random(2)
    hs = [h.encode() for h in _haystacks(rng, 600)]
    for pat in PY_SAFE:
        pr = re.compile(pat.encode())
        m = F.Matcher(F.REGEX, [pat])
        o = O.Matcher(O.REGEX, [pat])
        got = run_dfa(m, hs)
        for h, g in zip(hs, got):
            e = pr.search(h) is not None
            assert g == e, (pat, h)
            assert o.read_match(h) == e, ("oracle", pat, h)
    # `$` differs from Python before a trailing newline: regex crate's $ is end-of-haystack only
    assert run_dfa(F.Matcher(F.REGEX, ["A$"]), [b"A\n"])[0] is False or not run_dfa(F.Matcher(F.REGEX, ["A$"]), [b"A\n"])[0]
    assert not O.Matcher(O.REGEX, ["A$"]).read_match(b"A\n")


def test_random_regex_differential():
    rng = random.Random(3)
    atoms = ["A", "C", "G", "T", "N", ".", "[AC]", "[^G]", "[ACGT]", "(A|T)", "(CG|GC)", "^", "$", r"\b"]
    quants = ["", "", "", "*", "+", "?", "{2}", "{1,2}", "{0,1}"]
    hs = [h.encode() for h in _haystacks(rng, 300)]
    for _ in range(150):
        k = rng.randint(1, 5)
        pat = "".join(rng.choice(atoms) + rng.choice(quants) for _ in range(k))
        if rng.random() < 0.3:
            pat = pat + "|" + "".join(rng.choice(atoms[:9]) for _ in range(rng.randint(1, 4)))
        for rc, inv in ((False, False), (True, False), (True, True)):
            try:
                o = O.Matcher(O.REGEX_SET, [pat], inv, rc)
            except O.OracleError:
                with pytest.raises(F.FqgrepError):
                    F.Matcher(F.REGEX_SET, [pat], opts(inv, rc))
                continue
            m = F.Matcher(F.REGEX_SET, [pat], opts(inv, rc))
            got = run_dfa(m, hs)
            for h, g in zip(hs, got):
                assert g == o.read_match(h), (pat, rc, inv, h)


def test_literal_sets_vs_oracle():
    rng = random.Random(4)
    hs = [h.encode() for h in _haystacks(rng, 500)]
    for _ in range(60):
        pats = ["".join(rng.choice("ACGTN") for _ in range(rng.randint(0 if rng.random() < 0.05 else 1, 7))) for _ in range(rng.randint(1, 6))]
        for rc, inv in ((False, False), (True, False), (False, True), (True, True)):
            m = F.Matcher(F.FIXED_SET, pats, opts(inv, rc))
            o = O.Matcher(O.FIXED_SET, pats, inv, rc)
            got = run_dfa(m, hs)
            for h, g in zip(hs, got):
                assert g == o.read_match(h), (pats, rc, inv, h)
            # the reference's own property: FixedString == Regex on literal patterns (matcher.rs:1391-1486)
            if all(pats):
                r = F.Matcher(F.REGEX_SET, pats, opts(inv, rc))
                assert run_dfa(r, hs) == got


@pytest.mark.parametrize("name,kind,pats,inv,rc", [
    ("cfg1_literal_regex", F.REGEX, ["GACGAGATTA"], False, False),
    ("cfg1_fixed", F.FIXED, ["GACGAGATTA"], False, False),
    ("cfg2_two_regex_rc", F.REGEX_SET, ["GACGAGATTA", "GACGTGATTA"], False, True),
    ("cfg4_class_regex_invert", F.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"], True, False),
])
def test_baseline_configs_on_synthetic_reads(name, kind, pats, inv, rc):
    n = 20000
    mat = synth.bases(n, seed=1)
    lit = ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"]
    planted = synth.plant(mat, lit, rate=0.01)
    assert 100 < planted.size < 300
    bases, offs = synth.soa(mat)
    m = F.Matcher(kind, pats, opts(inv, rc))
    got = run_dfa_soa(m, bases, offs)
    cnt, flags = O.Matcher(kind, pats, inv, rc).match_soa(bases, offs, threads=4)
    assert (got == (flags != 0)).all()
    assert 0 < cnt < n


def test_large_literal_set_aho_corasick():
    rng = np.random.default_rng(7)
    pats = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(10000)]
    m = F.Matcher(F.FIXED_SET, pats, opts(False, True))
    info = m.dfa_info()
    assert info.tier == 3 and info.n_states > 100000     # q-gram filter on the device, exact automaton kept as fallback
    n = 3000
    mat = synth.bases(n, seed=5)
    synth.plant(mat, pats[:50], rate=0.05)
    bases, offs = synth.soa(mat)
    got = run_dfa_soa(m, bases, offs)
    cnt, flags = O.Matcher(O.FIXED_SET, pats[:50], False, True).match_soa(bases, offs, threads=4)
    # every read the 50-pattern oracle selects must be selected by the 10k set; spot-check the rest exactly
    assert (got[flags != 0]).all()
    full = O.Matcher(O.FIXED_SET, pats, False, True)
    idx = rng.integers(0, n, 200)
    for i in idx:
        assert got[i] == full.read_match(bytes(mat[i]))


def test_write_selected_is_host_glue_and_matches_oracle_output():
    """Ordered write-back (main.rs:641-657, 682-727) driven by a given mask: no device needed."""
    from tests.fq_util import fastq
    seqs = ["ACGT", "TTTT", "ACAC", "GGGG", "AC"]
    r1 = fastq(seqs, crlf=True, plus_name=True, final_newline=False)
    r2 = fastq([s[::-1] for s in seqs])
    o = O.Matcher(O.FIXED, ["AC"])
    e = o.grep(r1, want_output=True)
    def words(flags):
        b = np.packbits(flags.astype(bool), bitorder="little").tobytes()
        return np.frombuffer(b + b"\0" * (8 - len(b) % 4), dtype=np.uint32)
    mask = words(e["flags"])
    assert F.write_selected(mask, r1) == e["out"] == fastq(["ACGT", "ACAC", "AC"], heads=["@0", "@2", "@4"])
    e2 = o.grep(r1, r2, mode=O.PAIRED, want_output=True, split=True)
    mask2 = words(e2["flags"])
    a, b = F.write_selected(mask2, r1, r2, mode=F.PAIRED, split=True)
    assert a == e2["out"] and b == e2["out2"]
    e3 = o.grep(r1, r2, mode=O.PAIRED, want_output=True)
    assert F.write_selected(mask2, r1, r2, mode=F.PAIRED) == e3["out"]
    with pytest.raises(F.FqgrepError):
        F.write_selected(mask, b"@a\nACGT\n+\nXX\n")


def test_parallel_write_back_equals_the_serial_walk(monkeypatch, capfd):
    """fqg_write_selected cuts big inputs into shards and walks them on all host threads; it must equal the serial walk
    (= the oracle's output) byte for byte, and step aside for it whenever the input is not clean FASTQ.  The shard size
    is shrunk through the test hook so that a few hundred kilobytes already run the parallel form."""
    import random
    from tests.fq_util import fastq
    rng = random.Random(3)
    seqs = ["".join(rng.choice("ACGTN") for _ in range(rng.choice((1, 7, 36, 150, 151, 400)))) for _ in range(3000)]
    quals = None
    def words(flags):
        b = np.packbits(flags.astype(bool), bitorder="little").tobytes()
        return np.frombuffer(b + b"\0" * (8 - len(b) % 4), dtype=np.uint32)
    o = O.Matcher(O.REGEX_SET, ["ACG", "T{3}"])
    monkeypatch.setenv("FQG_TRACE", "1")
    for crlf, plus_name, final_nl in ((False, False, True), (True, True, False)):
        r1 = fastq(seqs, crlf=crlf, plus_name=plus_name, final_newline=final_nl)
        r2 = fastq([s[::-1] for s in seqs], crlf=not crlf)
        want1 = o.grep(r1, want_output=True)
        wantp = o.grep(r1, r2, mode=O.PAIRED, want_output=True)
        wants = o.grep(r1, r2, mode=O.PAIRED, want_output=True, split=True)
        ri = fastq(seqs, heads=["p%d" % (i // 2) for i in range(len(seqs))], crlf=crlf, plus_name=plus_name, final_newline=final_nl)
        wanti = o.grep(ri, mode=O.INTERLEAVED, want_output=True)
        for shard in ("1000", "4096", "100000"):
            monkeypatch.setenv("FQG_WRITE_SHARD_BYTES", shard)
            assert F.write_selected(words(want1["flags"]), r1) == want1["out"]
            assert F.write_selected(words(wantp["flags"]), r1, r2, mode=F.PAIRED) == wantp["out"]
            a, b = F.write_selected(words(wants["flags"]), r1, r2, mode=F.PAIRED, split=True)
            assert a == wants["out"] and b == wants["out2"]
            assert F.write_selected(words(wanti["flags"]), ri, mode=F.INTERLEAVED) == wanti["out"]
            err = capfd.readouterr().err
            assert err.count("parallel walk") == 4 and "serial walk" not in err, (shard, err)
    # quality lines that start with '@' (a shard cut must not take them for headers)
    monkeypatch.setenv("FQG_WRITE_SHARD_BYTES", "600")
    tricky = b"".join(b"@r%d\n%s\n+\n@%s\n" % (i, b"ACGT" * 10, b"I" * 39) for i in range(400))
    e = O.Matcher(O.FIXED, ["ACGT"]).grep(tricky, want_output=True)
    assert F.write_selected(words(e["flags"]), tricky) == e["out"] == tricky
    assert "parallel walk" in capfd.readouterr().err
    # malformed input inside a later shard: same error as the serial walk
    bad = tricky + b"@x\nACGT\n+\nII\n" + tricky
    with pytest.raises(F.FqgrepError):
        F.write_selected(words(np.ones(900)), bad)
    assert "serial walk" in capfd.readouterr().err
    # R2 longer than R1 is tolerated by the serial walk (it stops with R1), shorter is an error
    r1 = fastq(seqs[:500])
    r2 = fastq([s[::-1] for s in seqs[:520]])
    sel = np.zeros(500); sel[::7] = 1
    monkeypatch.delenv("FQG_WRITE_SHARD_BYTES")
    serial = F.write_selected(words(sel), r1, r2, mode=F.PAIRED)
    monkeypatch.setenv("FQG_WRITE_SHARD_BYTES", "2000")
    assert F.write_selected(words(sel), r1, r2, mode=F.PAIRED) == serial
    with pytest.raises(F.FqgrepError):
        F.write_selected(words(sel), r2, r1, mode=F.PAIRED)


def test_parallel_write_back_fuzz_against_the_serial_walk(monkeypatch):
    """Random valid and damaged files, random masks: the shard-parallel write-back returns exactly what the serial walk
    returns -- the same bytes, or an error when the serial walk reports one."""
    import random
    rng = random.Random(5)

    def rand_file(n):
        nl = b"\r\n" if rng.random() < 0.3 else b"\n"
        recs = []
        for i in range(n):
            L = rng.choice((0, 1, 4, 33, 150, 151, 400))
            seq = bytes(rng.choice(b"ACGTN") for _ in range(L))
            qual = bytes(rng.choice(b"@+I#") for _ in range(L))
            recs.append(b"@r%d x" % i + nl + seq + nl + b"+" + (b"r%d" % i if rng.random() < 0.2 else b"") + nl + qual + nl)
        data = b"".join(recs)
        return data[:-len(nl)] if rng.random() < 0.4 else data

    def run(mask, r1, r2, mode, split):
        try:
            return F.write_selected(mask, r1, r2, mode=mode, split=split)
        except F.FqgrepError as e:
            return ("error", e.code)

    for it in range(60):
        n = rng.choice((2, 50, 400, 1500))
        r1, r2 = rand_file(n), rand_file(n)
        if rng.random() < 0.5:            # damage one of them
            bad = bytearray(r1)
            a = rng.randrange(len(bad))
            if rng.random() < 0.5:
                del bad[a:a + rng.randrange(1, 30)]
            else:
                bad[a:a] = rng.choice((b"\n", b"@", b"+\n", b"\r"))
            r1 = bytes(bad)
        mask = np.frombuffer(bytes(rng.getrandbits(8) for _ in range((n // 32 + 3) * 4)), dtype=np.uint32)
        for mode, second, split in ((F.SINGLE, None, False), (F.INTERLEAVED, None, False), (F.PAIRED, r2, False), (F.PAIRED, r2, True)):
            monkeypatch.setenv("FQG_WRITE_SHARD_BYTES", str(1 << 40))
            serial = run(mask, r1, second, mode, split)
            monkeypatch.setenv("FQG_WRITE_SHARD_BYTES", rng.choice(("700", "5000", "60000")))
            assert run(mask, r1, second, mode, split) == serial, (it, mode, split)


def test_tier_choice_follows_the_measured_policy():
    """Which device layout a pattern set gets (fqg_dfa_info.tier), as DESIGN.md section 3 states it: the q-gram filter only
    where it pays, regexes with small finite languages counted as literals, short dense sets left to the automaton."""
    rng = np.random.default_rng(7)
    k20 = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(3000)]
    tier = lambda kind, pats, rc=False: F.Matcher(kind, pats, opts(False, rc)).dfa_info().tier
    assert tier(F.REGEX_SET, ["GACGAGATTA", "GACGTGATTA"], True) == 0            # configs[1]
    assert tier(F.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"]) == 0                          # small automaton: stays a table
    assert tier(F.FIXED_SET, k20[:1000]) == 3                                      # long literals: filter, one probe per 8 bases
    assert tier(F.REGEX_SET, k20[:1000]) == 3                                      # the same set given without -F
    assert tier(F.FIXED_SET, [p[:10] for p in k20[:1000]]) in (1, 2)               # 1000 x 10-mers: too dense for the filter
    assert tier(F.FIXED_SET, [p[:15] for p in k20[:60]]) == 4                      # stride-4 filter loses to 2-base rows
    classy = [p[:9] + "[" + p[9] + "N]" + p[10:] for p in k20[:300]]
    assert tier(F.REGEX_SET, classy, True) == 3                                    # finite languages multiplied out
    assert tier(F.REGEX_SET, classy + ["ACGT.*TTTTTTTTTTTT"], True) == 5           # one real regex: split set
    assert tier(F.REGEX_SET, ["^" + p[:12] for p in k20[:2000]]) in (1, 2)         # anchored barcodes: table + early stop
