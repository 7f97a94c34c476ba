"""The C++ host side (include/fqgrep_b200.hpp + fqgrep_b200/csrc/cli.cc -> fqgrep_b200/fqgrep-b200) run as a process,
against the reference's own end-to-end vectors (src/main.rs `mod tests`, transcribed in tests/golden/kat_reference.json):
the same command lines a user of the reference types, the same stdout / output files / exit codes.  Every vector also runs
through the Python mirror in tests/test_driver.py; here a spread of them is enough (each process start pays a CUDA context)."""
import os
import subprocess

import pytest

import fqgrep_b200.build as B
from tests.fq_util import seqs_of
from tests.test_driver import write_fastq, write_lines


def run(*args, stdin=None):
    B.build()
    p = subprocess.run([B.CLI] + [str(a) for a in args], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    return p.returncode, p.stdout, p.stderr.decode()


def test_cli_errors_need_no_device(tmp_path):
    paths = write_fastq(str(tmp_path), [["GTCAGC"], ["AGTGCG"]])
    code, _, err = run("-F", "^T", *paths)                       # main.rs:1315-1317 -> exit 2
    assert code == 2 and err == "Error: Fixed pattern must contain only DNA bases:  .. [^] .. T\n"
    code, _, err = run("-c", "*", *paths)                        # main.rs:1303-1306
    assert code == 2 and err.startswith("Error: Invalid regular expression: '*'")
    code, _, err = run("--protein", "-F", "^Q", *paths)
    assert code == 2 and err == "Error: Fixed pattern must contain only amino acids:  .. [^] .. Q\n"
    code, _, err = run("--paired", "-e", "A", *paths, paths[0])
    assert code == 2 and err.startswith("Error: Input files must be a multiple of two")
    code, _, err = run("-e", "A", "--output", "a", "b", "--", *paths)
    assert code == 2 and err == "Error: Two output files may only be given with --paired\n"
    code, _, err = run()
    assert code == 2 and err.startswith("Error: Pattern must be given with -e or as the first positional argument")
    names = write_lines(str(tmp_path), "n.txt", ["@0"])
    assert run("-N", names, "-e", "A", *paths)[0] == 2           # clap conflicts_with (main.rs:182)
    assert run("--protein", "--reverse-complement", "A", *paths)[0] == 2
    assert run("--iupac", "sometimes", "A", *paths)[0] == 2
    assert run("--color", "always", "A", *paths)[0] == 2         # colouring is host-side, out of scope: refused, not ignored
    assert run("--bogus")[0] == 2
    code, out, _ = run("--help")
    assert code == 0 and b"--read-names-file" in out and b"--iupac" in out
    assert run("--version")[0] == 0


@pytest.mark.gpu
def test_cli_counts_order_split_and_stdin(tmp_path, kat):
    e = kat["e2e_counts"]
    for i, (paired, pats, exp) in enumerate(e["cases"][::3]):
        d = tmp_path / ("c%d" % i)
        d.mkdir()
        paths = write_fastq(str(d), e["files"])
        pf = write_lines(str(d), "pattern.txt", pats)
        code, out, err = run("-c", "-f", pf, *(["--paired"] if paired else []), "-t", "4", *paths)
        assert out == b"%d\n" % exp and code == (0 if exp else 1), (paired, pats, err)
    e = kat["e2e_order"]
    for i, (paired, pats, exp) in enumerate(e["cases"][::2]):
        d = tmp_path / ("o%d" % i)
        d.mkdir()
        paths = write_fastq(str(d), e["files"])
        args = [x for p in pats for x in ("-e", p)] + (["--paired"] if paired else [])
        code, out, _ = run(*args, *paths)                        # default: stdout
        assert seqs_of(out) == exp, (paired, pats)
        if i == 0:
            outp = str(d / "output.fq")
            run(*args, "--output", outp, "--", *paths)
            assert seqs_of(open(outp, "rb").read()) == exp
    s = kat["e2e_split_output"]
    d = tmp_path / "split"
    d.mkdir()
    paths = write_fastq(str(d), [s["r1"], s["r2"]])
    code, out, _ = run("--paired", "-e", s["pattern"], "--output", d / "r1.fq", d / "r2.fq", "--", *paths)
    assert code == 0 and out == b""
    assert seqs_of(open(str(d / "r1.fq"), "rb").read()) == s["out_r1"] and seqs_of(open(str(d / "r2.fq"), "rb").read()) == s["out_r2"]
    # standard input: no files, or "-"; --paired on stdin means interleaved (main.rs:552-589)
    inter = b"".join(b"@%d\n%s\n+\n%s\n" % (k // 2, q.encode(), b"X" * len(q)) for k, q in enumerate(x for pair in zip(s["r1"], s["r2"]) for x in pair))
    code, out, _ = run("-c", "--paired", "-e", s["pattern"], stdin=inter)
    assert out == b"%d\n" % s["count"]
    code, out, _ = run("--paired", s["pattern"], "-", stdin=inter)
    assert seqs_of(out) == [x for pair in zip(s["out_r1"], s["out_r2"]) for x in pair]


@pytest.mark.gpu
def test_cli_flags_exit_codes_names_iupac_gzip(tmp_path, kat):
    e = kat["e2e_rc_invert_paired"]
    for i, (paired, inv, rc, exp) in enumerate(e["cases"][1::2]):
        d = tmp_path / ("r%d" % i)
        d.mkdir()
        paths = write_fastq(str(d), e["files"])
        flags = (["--paired"] if paired else []) + (["-v"] if inv else []) + (["--reverse-complement"] if rc else [])
        code, out, _ = run("-c", "-e", e["pattern"], *flags, *paths)
        assert out == b"%d\n" % exp, (paired, inv, rc)
    e = kat["e2e_exit"]
    for i, c in enumerate(e["cases"]):
        d = tmp_path / ("x%d" % i)
        d.mkdir()
        paths = write_fastq(str(d), e["files"])
        code, _, _ = run("-c", *(["-F"] if c["fixed"] else []), c["pattern"], *paths)
        assert code == c["exit"], c
    e = kat["e2e_protein"]
    for i, (fixed, pats, exp) in enumerate(e["cases"][::3]):
        d = tmp_path / ("p%d" % i)
        d.mkdir()
        paths = write_fastq(str(d), [e["reads"]])
        code, out, _ = run("--protein", *(["-F"] if fixed else []), "-f", write_lines(str(d), "pattern.txt", pats), *paths)
        assert seqs_of(out) == exp
    n = kat["e2e_names"]
    for i, (inv, names, exp) in enumerate(n["single"]["cases"][1::2]):
        d = tmp_path / ("n%d" % i)
        d.mkdir()
        paths = write_fastq(str(d), [["ACGT"] * n["single"]["n_reads"]])
        code, out, _ = run("-c", "-N", write_lines(str(d), "names.txt", names), *(["-v"] if inv else []), *paths)
        assert out == b"%d\n" % exp
    for i, (inv, names, exp) in enumerate(n["paired"]["cases"][:2]):
        d = tmp_path / ("np%d" % i)
        d.mkdir()
        k = n["paired"]["n_reads_per_file"]
        paths = write_fastq(str(d), [["ACGT"] * k, ["TTTT"] * k])
        code, out, _ = run("-c", "--paired", "--read-names-file=" + write_lines(str(d), "names.txt", names), *(["-v"] if inv else []), *paths)
        assert out == b"%d\n" % exp
    for i, c in enumerate(kat["e2e_iupac_modes"]["cases"][::3]):
        for mode in ("expand", "regex", "bit-mask"):
            d = tmp_path / ("i%d%s" % (i, mode))
            d.mkdir()
            paths = write_fastq(str(d), [c["seqs"]])
            args = [x for p in c["patterns"] for x in ("-e", p)]
            code, out, _ = run("-cF", "--iupac", mode, *(["-v"] if c["invert"] else []), *args, *paths)
            assert out == b"%d\n" % c["count"], (mode, c)
    seqs = [["GTCAG", "ACGT", "GGTG"], ["AGTGCGTAGCTGATGCTCGAGCTGAC"], ["GGGTCTGAATCCATGGAAAGCTATTG"]]
    for i, ext in enumerate((".fq.gz",)):        # main.rs:1275-1297
        d = tmp_path / ("z%d" % i)
        d.mkdir()
        code, out, _ = run("-c", "^G", *write_fastq(str(d), seqs, ext))
        assert out == b"3\n"
    d = tmp_path / "bad"
    d.mkdir()
    bad = os.path.join(str(d), "bad.fq")
    open(bad, "wb").write(b"@a\nACGT\n+\nXXX\n")
    code, _, err = run("-c", "A", bad)                           # the reference panics on malformed input -> 101
    assert code == 101 and "fqgrep panicked" in err
