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


@pytest.mark.gpu
def test_cli_streams_large_inputs_through_a_small_ring(tmp_path):
    """The command line reads its inputs a slot at a time (reader threads -> ring -> ordered write-back).  With 1 MiB slots
    a 30 MB pair of gzip / BGZF / plain inputs takes ~30 batches per stream; counts and output bytes must be the oracle's,
    on one device and with the batches dealt to two streams (--gpus 2 needs two devices: skipped on a one-GPU box)."""
    import gzip
    import random
    from oracle import oracle as O
    from tests.test_gpu_stream import ragged_pair
    from tests.test_input import bgzf
    rng = random.Random(5)
    r1, r2 = ragged_pair(rng, 60_000)
    p1, p2, pz1, pz2, pb1 = (str(tmp_path / n) for n in ("a_1.fq", "a_2.fq", "z_1.fq.gz", "z_2.fq.gz", "b_1.fq.bgz"))
    open(p1, "wb").write(r1); open(p2, "wb").write(r2)
    open(pz1, "wb").write(gzip.compress(r1, 1)); open(pz2, "wb").write(gzip.compress(r2, 1))
    open(pb1, "wb").write(bgzf(r1, block=60000))
    pats = ["GATTACA", "AC[GT]T{2,}"]
    o = O.Matcher(O.REGEX_SET, pats, False, True)
    e1 = o.grep(r1, want_output=True, threads=8)
    ep = o.grep(r1, r2, mode=O.PAIRED, want_output=True, split=True, threads=8)
    args = ["-e", pats[0], "-e", pats[1], "--reverse-complement", "--slot-mb", "1"]
    for path in (p1, pz1, pb1):
        code, out, err = run(*args, path)
        assert code == 0 and out == e1["out"], (path, err[-300:])
        code, out, err = run("-c", *args, path)
        assert out == b"%d\n" % e1["count"]
    code, out, err = run("--paired", *args, p1, p2)
    assert code == 0 and out == O.Matcher(O.REGEX_SET, pats, False, True).grep(r1, r2, mode=O.PAIRED, want_output=True, threads=8)["out"], err[-300:]
    code, out, err = run("--paired", *args, "--output", tmp_path / "o1.fq", tmp_path / "o2.fq", "--", pz1, pz2)
    assert code == 0 and open(str(tmp_path / "o1.fq"), "rb").read() == ep["out"] and open(str(tmp_path / "o2.fq"), "rb").read() == ep["out2"], err[-300:]
    code, out, err = run("-c", "--paired", "--progress", *args, p1, p2, pz1, pz2)         # counts add up over the input pairs
    assert out == b"%d\n" % (2 * ep["count"]) and "Searched" in err
    # mates that run out of step are the reference's panic, wherever in the stream it happens
    open(str(tmp_path / "short_2.fq"), "wb").write(r2[: r2.rfind(b"@q")])
    code, _, err = run("-c", "--paired", *args, p1, tmp_path / "short_2.fq")
    assert code == 101 and "out of sync" in err
    # a record larger than a slot cannot be batched: a clear error, not a hang
    big = b"@big\n" + b"A" * (3 << 20) + b"\n+\n" + b"I" * (3 << 20) + b"\n"
    open(str(tmp_path / "big.fq"), "wb").write(r1[: r1.index(b"@q40/")] + big)
    code, _, err = run("-c", *args, tmp_path / "big.fq")
    assert code == 2 and "--slot-mb" in err
    code, out, _ = run("-c", "-e", "A{100}", "--slot-mb", "16", tmp_path / "big.fq")
    assert out == b"1\n"
    import torch
    if torch.cuda.device_count() >= 2:
        code, out, err = run("--paired", "--gpus", "2", *args, p1, p2)
        assert code == 0 and out == O.Matcher(O.REGEX_SET, pats, False, True).grep(r1, r2, mode=O.PAIRED, want_output=True, threads=8)["out"], err[-300:]
