"""Pins the CPU oracle against every golden vector the reference's own tests hold for the hot path."""
import pytest

from oracle import oracle as O
from tests.fq_util import fastq, seqs_of


def test_fixed_string(kat):
    for rc, pat, seq, exp in kat["fixed_string"]["cases"]:
        for inv in (False, True):
            assert O.Matcher(O.FIXED, [pat], inv, rc).read_match(seq) == (exp != inv), (rc, pat, seq, inv)


def test_fixed_string_set(kat):
    for rc, pats, seq, exp in kat["fixed_string_set"]["cases"]:
        for inv in (False, True):
            assert O.Matcher(O.FIXED_SET, pats, inv, rc).read_match(seq) == (exp != inv), (rc, pats, seq, inv)


def test_regex(kat):
    for rc, pat, seq, exp in kat["regex"]["cases"]:
        for inv in (False, True):
            assert O.Matcher(O.REGEX, [pat], inv, rc).read_match(seq) == (exp != inv), (rc, pat, seq, inv)


def test_regex_set(kat):
    for rc, pats, seq, exp in kat["regex_set"]["cases"]:
        for inv in (False, True):
            assert O.Matcher(O.REGEX_SET, pats, inv, rc).read_match(seq) == (exp != inv), (rc, pats, seq, inv)


def test_query_name(kat):
    for inv, names, rid, exp in kat["query_name"]["cases"]:
        m = O.Matcher(O.NAMES, names, inv, False)
        assert m.read_match("ACGT", head=rid + " description") == exp
    hc = kat["query_name"]["header_case"]
    assert O.Matcher(O.NAMES, hc["names"]).read_match("ACGT", head=hc["header"]) == hc["expect"]


def test_rc_special(kat):
    for pat, seq, fwd, rc in kat["rc_special"]["cases"]:
        assert O.Matcher(O.FIXED, [pat], False, False).read_match(seq) == fwd
        assert O.Matcher(O.FIXED, [pat], False, True).read_match(seq) == rc


def test_revcomp(kat):
    for s, r in kat["revcomp"]["cases"]:
        assert O.revcomp(s) == r.encode()


def test_validate_fixed(kat):
    v = kat["validate_fixed"]
    for p, prot in v["ok"]:
        assert O.validate_fixed(p, prot) is None
    for p, prot, msg in v["err"]:
        assert O.validate_fixed(p, prot) == msg
    for p, prot, sub in v["err_contains"]:
        assert sub in O.validate_fixed(p, prot)


def _count_files(m, files, paired):
    total = 0
    if paired:
        for i in range(0, len(files), 2):
            total += m.grep(fastq(files[i]), fastq(files[i + 1]), mode=O.PAIRED)["count"]
    else:
        for f in files:
            total += m.grep(fastq(f))["count"]
    return total


def test_e2e_counts(kat):
    e = kat["e2e_counts"]
    for paired, pats, exp in e["cases"]:
        assert _count_files(O.Matcher(O.REGEX_SET, pats), e["files"], paired) == exp, (paired, pats)


def test_e2e_order(kat):
    e = kat["e2e_order"]
    for paired, pats, exp in e["cases"]:
        m = O.Matcher(O.REGEX_SET, pats)
        if paired:
            out = m.grep(fastq(e["files"][0]), fastq(e["files"][1]), mode=O.PAIRED, want_output=True, threads=4)["out"]
        else:
            out = b"".join(m.grep(fastq(f), want_output=True, threads=4)["out"] for f in e["files"])
        assert seqs_of(out) == exp, (paired, pats)


def test_e2e_rc_invert_paired(kat):
    e = kat["e2e_rc_invert_paired"]
    for paired, inv, rc, exp in e["cases"]:
        m = O.Matcher(O.REGEX_SET, [e["pattern"]], inv, rc)
        assert _count_files(m, e["files"], paired) == exp, (paired, inv, rc)


def test_e2e_split_output(kat):
    e = kat["e2e_split_output"]
    r = O.Matcher(O.REGEX_SET, [e["pattern"]]).grep(fastq(e["r1"]), fastq(e["r2"]), mode=O.PAIRED, want_output=True, split=True)
    assert r["count"] == e["count"]
    assert seqs_of(r["out"]) == e["out_r1"] and seqs_of(r["out2"]) == e["out_r2"]


def test_e2e_names(kat):
    e = kat["e2e_names"]
    seqs = ["ACGT"] * e["single"]["n_reads"]
    for inv, names, exp in e["single"]["cases"]:
        assert O.Matcher(O.NAMES, names, inv).grep(fastq(seqs))["count"] == exp
    n = e["paired"]["n_reads_per_file"]
    for inv, names, exp in e["paired"]["cases"]:
        assert O.Matcher(O.NAMES, names, inv).grep(fastq(["ACGT"] * n), fastq(["TTTT"] * n), mode=O.PAIRED)["count"] == exp
    for inv, names, exp in e["empty"]["cases"]:
        assert O.Matcher(O.NAMES, names, inv).grep(fastq(["ACGT"] * e["empty"]["n_reads"]))["count"] == exp


def test_e2e_exit(kat):
    e = kat["e2e_exit"]
    for c in e["cases"]:
        code = None
        try:
            if c["fixed"]:
                if O.validate_fixed(c["pattern"]) is not None:
                    raise O.OracleError("validation")
                m = O.Matcher(O.FIXED, [c["pattern"]])
            else:
                m = O.Matcher(O.REGEX, [c["pattern"]])
            code = 0 if _count_files(m, e["files"], False) > 0 else 1
        except O.OracleError:
            code = 2
        assert code == c["exit"], c


def test_e2e_protein(kat):
    e = kat["e2e_protein"]
    for fixed, pats, exp in e["cases"]:
        m = O.Matcher(O.FIXED_SET if fixed else O.REGEX_SET, pats)
        assert seqs_of(m.grep(fastq(e["reads"]), want_output=True)["out"]) == exp


def test_fastq_rules():
    m = O.Matcher(O.FIXED, ["AC"])
    good = fastq(["ACGT", "TTTT"])
    assert m.grep(good)["count"] == 1
    # CRLF and +name are normalised on output (seq_io write_to); missing final newline accepted
    r = m.grep(fastq(["ACGT", "TTAC"], crlf=True, plus_name=True, final_newline=False), want_output=True)
    assert r["count"] == 2 and r["out"] == fastq(["ACGT", "TTAC"])
    for bad in (b"ACGT\n", b"@a\nACGT\n-\nXXXX\n", b"@a\nACGT\n+\nXXX\n", b"@a\nACGT\n+\n", b"@a\nACGT\n"):
        with pytest.raises(O.OracleError):
            m.grep(bad)
    with pytest.raises(O.OracleError):
        m.grep(fastq(["ACGT"] * 3), mode=O.INTERLEAVED)
    with pytest.raises(O.OracleError):
        m.grep(fastq(["ACGT"] * 3), fastq(["ACGT"] * 2), mode=O.PAIRED)
    with pytest.raises(O.OracleError):   # id mismatch panics (main.rs:741-747)
        m.grep(fastq(["ACGT"], heads=["a 1"]), fastq(["ACGT"], heads=["b 1"]), mode=O.PAIRED)
    assert m.grep(fastq(["ACGT"], heads=["a 1"]), fastq(["GGGG"], heads=["a 2"]), mode=O.PAIRED)["count"] == 1
    # id ends at the first space only; a tab is part of the id (seq_io 0.3.4 id_bytes)
    assert O.Matcher(O.NAMES, ["a\tb"]).grep(fastq(["ACGT"], heads=["a\tb c"]))["count"] == 1
    assert m.grep(b"")["count"] == 0


def test_same_length_set_fast_path_equals_the_memmem_loop():
    """The oracle answers large same-length literal sets by window hashing (so that bench.py can check whole 2 M-read masks
    of the 10 000-pattern config); it must say exactly what its plain one-memmem-per-pattern loop says."""
    import numpy as np
    from tests import synth
    rng = np.random.default_rng(3)
    for L, npat in ((20, 500), (8, 64), (31, 33)):
        pats = ["".join("ACGT"[i] for i in rng.integers(0, 4, L)) for _ in range(npat)]
        mat = synth.bases(20_000, seed=9)
        synth.plant(mat, pats[:40], rate=0.05, seed=5)
        mat[7, :3] = np.frombuffer(b"NNN", dtype=np.uint8)
        bases, offs = synth.soa(mat)
        for inv, rc in ((False, False), (True, True)):
            fast = O.Matcher(O.FIXED_SET, pats, inv, rc)
            slow = O.Matcher(O.FIXED_SET, pats, inv, rc)
            slow.set_naive(True)
            c1, f1 = fast.match_soa(bases, offs, threads=4)
            c2, f2 = slow.match_soa(bases, offs, threads=4)
            assert c1 == c2 and (f1 == f2).all() and 0 < c1 < 20_000


def test_oracle_agrees_with_gnu_grep_on_the_baseline_patterns(tmp_path):
    """GNU grep -c -F / -E on the extracted sequence lines is an independent implementation; the oracle (and so
    everything checked against it) must count the same lines for the BASELINE.json patterns, -v included."""
    import shutil
    import subprocess
    import numpy as np
    from tests import synth
    if not shutil.which("grep"):
        import pytest
        pytest.skip("no grep")
    n = 30_000
    mat = synth.bases(n, seed=1)
    rng = np.random.default_rng(7)
    p1k = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(1000)]
    synth.plant(mat, ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"] + p1k[:30], rate=0.05, seed=99, revcomp_half=False)
    seqfile = tmp_path / "seqs.txt"
    seqfile.write_bytes(b"\n".join(bytes(r) for r in mat) + b"\n")
    patfile = tmp_path / "pats.txt"
    patfile.write_text("\n".join(p1k) + "\n")
    bases, offs = synth.soa(mat)

    def grep(args):
        r = subprocess.run(["grep", "-c"] + args + [str(seqfile)], capture_output=True, text=True, env={"LC_ALL": "C"})
        return int(r.stdout.strip() or 0)
    cases = [
        (O.FIXED, ["GACGAGATTA"], False, ["-F", "GACGAGATTA"]),                                   # configs[0] with -F
        (O.REGEX, ["GACGAGATTA"], False, ["-E", "GACGAGATTA"]),                                   # configs[0]
        (O.REGEX_SET, ["GACGAGATTA", "GACGTGATTA"], False, ["-E", "-e", "GACGAGATTA", "-e", "GACGTGATTA"]),   # configs[1] without RC
        (O.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"], True, ["-v", "-E", "AC[GT]{3}TT(A|G)CA"]),          # configs[3]
        (O.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"], False, ["-E", "AC[GT]{3}TT(A|G)CA"]),
        (O.FIXED_SET, p1k, False, ["-F", "-f", str(patfile)]),                                    # configs[2] / [4a] shape
    ]
    for kind, pats, inv, args in cases:
        cnt, _ = O.Matcher(kind, pats, inv, False).match_soa(bases, offs, threads=4)
        assert cnt == grep(args), (kind, pats[:2], inv)
        assert 0 < cnt < n
