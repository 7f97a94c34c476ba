"""The reference's integration tests (src/main.rs `mod tests`, built on write_fastq / write_pattern / build_opts /
fqgrep_from_opts) restated against fqgrep_b200.driver: temp FASTQ files, pattern files, name files, output files, exit
codes.  Everything matches on the GPU, so the module is marked gpu; the option-normalisation errors need no device."""
import gzip
import io
import os

import pytest

import fqgrep_b200 as F
from tests.fq_util import fastq, seqs_of


def write_fastq(d, seqs_per_file, ext=".fq"):
    """main.rs:790-821: sample_{i}{ext}, heads "@{n}" (so ids are "@0", "@1", ...), qualities 'X'."""
    paths = []
    for i, seqs in enumerate(seqs_per_file):
        p = os.path.join(d, "sample_%d%s" % (i, ext))
        data = fastq(seqs)
        if ext.endswith(".gz") or ext.endswith(".bgz"):
            with gzip.open(p, "wb") as f:
                f.write(data)
        else:
            open(p, "wb").write(data)
        paths.append(p)
    return paths


def write_lines(d, name, lines):
    p = os.path.join(d, name)
    open(p, "w").write("".join(l + "\n" for l in lines))
    return p


def build_opts(d, seqs, regexp, pattern_from_file, output=(), ext=".fq"):
    """main.rs:863-900"""
    paths = write_fastq(str(d), seqs, ext)
    o = F.Opts(args=paths, count=not output, output=list(output))
    if pattern_from_file:
        o.file = write_lines(str(d), "pattern.txt", regexp)
    else:
        o.regexp = list(regexp)
    return o


def test_option_errors_need_no_device(tmp_path):
    seqs = [["GGGG", "TTTT"], ["AAAA", "CCCC"]]
    for fixed, protein, pats, msg in ((True, False, ["^G"], "Fixed pattern must contain only DNA bases:  .. [^] .. G"),
                                     (True, False, ["^A", "AA"], "Fixed pattern must contain only DNA bases:  .. [^] .. A"),
                                     (True, True, ["^Q"], "Fixed pattern must contain only amino acids:  .. [^] .. Q"),
                                     (True, True, ["^Q", "QQ"], "Fixed pattern must contain only amino acids:  .. [^] .. Q")):
        o = build_opts(tmp_path, seqs, pats, True)          # main.rs:1176-1210
        o.fixed_strings, o.protein = fixed, protein
        with pytest.raises(F.FqgrepError) as e:
            F.fqgrep_from_opts(o)
        assert e.value.message == msg
    o = build_opts(tmp_path, [["A"], ["C"], ["G"]], ["A"], True)      # main.rs:1212-1237
    o.paired = True
    with pytest.raises(F.FqgrepError) as e:
        F.fqgrep_from_opts(o)
    assert e.value.message.startswith("Input files must be a multiple of two")
    with pytest.raises(F.FqgrepError):
        F.fqgrep_from_opts(F.Opts(args=[]))
    o = build_opts(tmp_path, [["A"]], ["A"], False, output=["a", "b"])
    with pytest.raises(F.FqgrepError) as e:
        F.fqgrep_from_opts(o)
    assert e.value.message == "Two output files may only be given with --paired"
    assert F.read_patterns(write_lines(str(tmp_path), "p.txt", [" A ", "", "C"])) == [" A ", "", "C"]     # not trimmed (main.rs:251-253)
    assert F.read_query_names(write_lines(str(tmp_path), "n.txt", [" r1 ", "", "r2"])) == {b"r1", b"r2"}   # trimmed, empties skipped


@pytest.mark.gpu
def test_counts_order_and_split_outputs(tmp_path, kat):
    e = kat["e2e_counts"]
    for i, (paired, pats, exp) in enumerate(e["cases"]):
        d = tmp_path / ("c%d" % i)
        d.mkdir()
        o = build_opts(d, e["files"], pats, True)
        o.paired = paired
        out = io.BytesIO()
        assert F.fqgrep_from_opts(o, stdout=out) == exp, (paired, pats)
        assert out.getvalue() == b"%d\n" % exp                 # --count prints one line (main.rs:667-671)
    e = kat["e2e_order"]
    for i, (paired, pats, exp) in enumerate(e["cases"]):
        d = tmp_path / ("o%d" % i)
        d.mkdir()
        outp = str(d / "output.fq")
        o = build_opts(d, e["files"], pats, True, output=[outp])
        o.paired = paired
        F.fqgrep_from_opts(o)
        assert seqs_of(open(outp, "rb").read()) == exp, (paired, pats)
    s = kat["e2e_split_output"]
    d = tmp_path / "split"
    d.mkdir()
    o = build_opts(d, [s["r1"], s["r2"]], [s["pattern"]], False, output=[str(d / "r1.fq"), str(d / "r2.fq")])
    o.paired = True
    assert F.fqgrep_from_opts(o) == s["count"]
    assert seqs_of(open(str(d / "r1.fq"), "rb").read()) == s["out_r1"] and seqs_of(open(str(d / "r2.fq"), "rb").read()) == s["out_r2"]
    # default output goes to stdout, interleaved for pairs (main.rs:519-525, 692-701)
    o = build_opts(d, [s["r1"], s["r2"]], [s["pattern"]], False)
    o.paired, o.count = True, False
    out = io.BytesIO()
    F.fqgrep_from_opts(o, stdout=out)
    assert seqs_of(out.getvalue()) == [x for pair in zip(s["out_r1"], s["out_r2"]) for x in pair]


@pytest.mark.gpu
def test_rc_invert_paired_protein_and_pattern_sources(tmp_path, kat):
    e = kat["e2e_rc_invert_paired"]
    for i, (paired, inv, rc, exp) in enumerate(e["cases"]):
        d = tmp_path / ("r%d" % i)
        d.mkdir()
        o = build_opts(d, e["files"], [e["pattern"]], False)
        o.paired, o.invert_match, o.reverse_complement = paired, inv, rc
        assert F.fqgrep_from_opts(o, stdout=io.BytesIO()) == exp, (paired, inv, rc)
    e = kat["e2e_protein"]
    for i, (fixed, pats, exp) in enumerate(e["cases"]):
        d = tmp_path / ("p%d" % i)
        d.mkdir()
        outp = str(d / "output.fq")
        o = build_opts(d, [e["reads"]], pats, True, output=[outp])
        o.protein, o.fixed_strings = True, fixed
        F.fqgrep_from_opts(o)
        assert seqs_of(open(outp, "rb").read()) == exp
    # main.rs:1239-1273: the same patterns from a file and from -e
    d = tmp_path / "src"
    d.mkdir()
    seqs = [["GTCAGCTCGAGCATCAGCTACGCACT"], ["AGTGCGTAGCTGATGCTCGAGCTGAC"], ["GGGTCTGAATCCATGGAAAGCTATTG"]]
    o = build_opts(d, seqs, ["^G"], True)
    assert F.fqgrep_from_opts(o, stdout=io.BytesIO()) == 2
    o.regexp = ["^G"]
    assert F.fqgrep_from_opts(o, stdout=io.BytesIO()) == 2
    # positional pattern (RegexMatcher / FixedStringMatcher paths of the factory)
    paths = write_fastq(str(d), seqs)
    assert F.fqgrep_from_opts(F.Opts(args=["^G"] + paths, count=True), stdout=io.BytesIO()) == 2
    assert F.fqgrep_from_opts(F.Opts(args=["GTCAGC"] + paths, count=True, fixed_strings=True), stdout=io.BytesIO()) == 1


@pytest.mark.gpu
def test_compression_exit_codes_names_and_iupac(tmp_path, kat):
    for i, ext in enumerate((".fq", ".fastq", ".fq.gz", ".fq.bgz")):            # main.rs:1275-1297
        d = tmp_path / ("z%d" % i)
        d.mkdir()
        seqs = [["GTCAG", "ACGT", "GGTG"], ["AGTGCGTAGCTGATGCTCGAGCTGAC"], ["GGGTCTGAATCCATGGAAAGCTATTG"]]
        assert F.fqgrep_from_opts(build_opts(d, seqs, ["^G"], True, ext=ext), stdout=io.BytesIO()) == 3
    e = kat["e2e_exit"]                                                         # main.rs:1299-1318
    for i, c in enumerate(e["cases"]):
        d = tmp_path / ("x%d" % i)
        d.mkdir()
        paths = write_fastq(str(d), e["files"])
        o = F.Opts(args=[c["pattern"]] + paths, count=True, fixed_strings=c["fixed"])
        code = F.fqgrep(o, stdout=io.BytesIO(), stderr=io.StringIO())
        assert (code or 0) == c["exit"], c
    n = kat["e2e_names"]                                                        # main.rs:1320-1402
    for i, (inv, names, exp) in enumerate(n["single"]["cases"]):
        d = tmp_path / ("n%d" % i)
        d.mkdir()
        o = build_opts(d, [["ACGT"] * n["single"]["n_reads"]], [], False)
        o.read_names_file, o.invert_match = write_lines(str(d), "names.txt", names), inv
        assert F.fqgrep_from_opts(o, stdout=io.BytesIO()) == exp
    for i, (inv, names, exp) in enumerate(n["paired"]["cases"]):
        d = tmp_path / ("np%d" % i)
        d.mkdir()
        k = n["paired"]["n_reads_per_file"]
        o = build_opts(d, [["ACGT"] * k, ["TTTT"] * k], [], False)
        o.read_names_file, o.invert_match, o.paired = write_lines(str(d), "names.txt", names), inv, True
        assert F.fqgrep_from_opts(o, stdout=io.BytesIO()) == exp
    for i, (inv, names, exp) in enumerate(n["empty"]["cases"]):
        d = tmp_path / ("ne%d" % i)
        d.mkdir()
        o = build_opts(d, [["ACGT"] * n["empty"]["n_reads"]], [], False)
        o.read_names_file, o.invert_match = write_lines(str(d), "names.txt", names), inv
        assert F.fqgrep_from_opts(o, stdout=io.BytesIO()) == exp
    for i, c in enumerate(kat["e2e_iupac_modes"]["cases"]):                      # main.rs:1408-1522
        for mode in ("expand", "regex", "bit-mask"):
            d = tmp_path / ("i%d%s" % (i, mode))
            d.mkdir()
            o = build_opts(d, [c["seqs"]], c["patterns"], False)
            o.fixed_strings, o.iupac, o.invert_match = True, mode, c["invert"]
            assert F.fqgrep_from_opts(o, stdout=io.BytesIO()) == c["count"], (mode, c)
    d = tmp_path / "lim"
    d.mkdir()
    o = build_opts(d, [["ACGT"]], ["NNNNNNNN"], False)
    o.fixed_strings, o.iupac = True, "expand"
    assert F.fqgrep(o, stdout=io.BytesIO(), stderr=io.StringIO()) == 2
    # malformed input: the reference panics -> 101
    bad = os.path.join(str(d), "bad.fq")
    open(bad, "wb").write(b"@a\nACGT\n+\nXXX\n")
    assert F.fqgrep(F.Opts(args=["A", bad], count=True), stdout=io.BytesIO(), stderr=io.StringIO()) == 101
