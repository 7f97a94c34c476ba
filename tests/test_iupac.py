"""--iupac modes: expand / regex are host rewrites, bit-mask is compiled to the same DFA family.
CPU part: product compilers (exported DFA walked by the test-side simulator) and oracle vs the reference's vectors.
GPU part: the three modes agree end to end through the fused FASTQ path."""
import random

import pytest

import fqgrep_b200 as F
from oracle import oracle as O
from tests.dfa_sim import run_dfa
from tests.fq_util import fastq


def opts(inv=False, rc=False):
    return F.MatcherOpts(invert_match=inv, reverse_complement=rc)


def test_expand_functions_match_reference_vectors(kat):
    e = kat["iupac_expand"]
    for pat, exp in e["fixed"]:
        assert F.expand_iupac_fixed_pattern(pat) == exp
        assert O.expand_iupac_fixed_pattern(pat) == exp
    for pat, exp in e["regex"]:
        assert F.expand_iupac_regex(pat) == exp
        assert O.expand_iupac_regex(pat) == exp
    with pytest.raises(F.FqgrepError) as ex:
        F.expand_iupac_fixed_pattern(e["limit"]["pattern"])
    assert e["limit"]["message_contains"] in ex.value.message and ex.value.exit_code == 2
    with pytest.raises(O.OracleError) as ex2:
        O.expand_iupac_fixed_pattern(e["limit"]["pattern"])
    assert e["limit"]["message_contains"] in str(ex2.value)


def test_bitmask_kat_oracle_and_compiled_dfa(kat):
    cases = [(rc, [p], s, e) for rc, p, s, e in kat["bitmask_iupac"]["cases"]]
    cases += [(rc, ps, s, e) for rc, ps, s, e in kat["bitmask_set_iupac"]["cases"]]
    # the DNA-only vectors of FixedString / FixedStringSet are asserted for the bit-mask matchers too (matcher.rs:1114-1147, 1197-1233)
    cases += [(rc, [p], s, e) for rc, p, s, e in kat["fixed_string"]["cases"]]
    cases += [(rc, ps, s, e) for rc, ps, s, e in kat["fixed_string_set"]["cases"]]
    b = kat["bitmask_short_read"]
    cases.append((False, [b["pattern"]], b["seq"], b["expect"]))
    for rc, pats, seq, exp in cases:
        kind_o = O.BITMASK if len(pats) == 1 else O.BITMASK_SET
        kind_f = F.BITMASK if len(pats) == 1 else F.BITMASK_SET
        assert O.Matcher(kind_o, pats, False, rc).read_match(seq) == exp, (rc, pats, seq)
        assert run_dfa(F.Matcher(kind_f, pats, opts(False, rc)), [seq])[0] == exp, (rc, pats, seq)


def test_bitmask_compiled_dfa_vs_oracle_random_incl_quirks():
    """Non-IUPAC read bytes (mask 0) match every symbol of an IUPAC needle but never a pure-ACGT needle
    (the rolling-hash prefilter, matcher.rs:320-342); lowercase and U behave like their uppercase / T."""
    rng = random.Random(8)
    needle_alpha = "ACGTUMRWSYKVHDBNacgtn" + "X"
    read_alpha = "ACGTNacgtURYKMXZ.*-"
    for _ in range(120):
        pats = ["".join(rng.choice(needle_alpha) for _ in range(rng.randint(0 if rng.random() < 0.03 else 1, 6))) for _ in range(rng.randint(1, 4))]
        reads = ["".join(rng.choice(read_alpha) for _ in range(rng.randint(0, 14))) for _ in range(150)]
        for rc, inv in ((False, False), (True, False), (True, True)):
            kind_o = O.BITMASK if len(pats) == 1 else O.BITMASK_SET
            kind_f = F.BITMASK if len(pats) == 1 else F.BITMASK_SET
            o = O.Matcher(kind_o, pats, inv, rc)
            got = run_dfa(F.Matcher(kind_f, pats, opts(inv, rc)), reads)
            for r, g in zip(reads, got):
                assert g == o.read_match(r), (pats, rc, inv, r)


def _modes(patterns, inv):
    for mode in ("expand", "regex", "bit-mask"):
        pattern, regexp, fixed, bitencs = F.expand_iupac(patterns[0] if len(patterns) == 1 else None, patterns if len(patterns) > 1 else [], True, mode)
        yield mode, F.MatcherFactory.new_matcher(pattern, fixed, regexp, bitencs, opts(inv))


def test_expand_iupac_dispatch_mirrors_main_rs():
    assert F.expand_iupac("GATK", [], True, "never") == ("GATK", [], True, None)
    assert F.expand_iupac("GATK", [], False, "expand") == ("GATK", [], False, None)       # ignored without -F (main.rs:327-330)
    assert F.expand_iupac("GATK", [], True, "expand") == (None, ["GATG", "GATT"], True, None)
    assert F.expand_iupac("GATK", [], True, "regex") == (None, ["GAT[GT]"], False, None)
    assert F.expand_iupac("GATK", [], True, "bit-mask") == ("GATK", [], True, ["GATK"])
    assert F.expand_iupac(None, ["GATK", "ACS"], True, "bit-mask") == (None, ["GATK", "ACS"], True, ["GATK", "ACS"])


def test_iupac_modes_agree_cpu(kat):
    for c in kat["e2e_iupac_modes"]["cases"]:
        for mode, m in _modes(c["patterns"], c["invert"]):
            assert sum(run_dfa(m, c["seqs"])) == c["count"], (mode, c)


@pytest.mark.gpu
def test_iupac_modes_agree_gpu(kat):
    for c in kat["e2e_iupac_modes"]["cases"]:
        for mode, m in _modes(c["patterns"], c["invert"]):
            assert m.grep_fastq(fastq(c["seqs"]))["count"] == c["count"], (mode, c)
    with pytest.raises(F.FqgrepError) as ex:
        F.expand_iupac("NNNNNNNN", [], True, "expand")
    assert ex.value.exit_code == 2


@pytest.mark.gpu
def test_bitmask_gpu_vs_oracle_on_reads():
    import numpy as np
    from tests import synth
    mat = synth.bases(60_000, seed=13)
    mat[::97, 40] = ord("N")
    mat[::101, 10:14] = np.frombuffer(b"acgt", dtype=np.uint8)
    bases, offs = synth.soa(mat)
    for pats in (["GATKACR"], ["ACGTAC", "TTNNAA", "GGYR"]):
        kind_o = O.BITMASK if len(pats) == 1 else O.BITMASK_SET
        kind_f = F.BITMASK if len(pats) == 1 else F.BITMASK_SET
        for inv, rc in ((False, False), (True, True)):
            bits, cnt = F.Matcher(kind_f, pats, opts(inv, rc)).match_bases(bases, offs)
            ocnt, flags = O.Matcher(kind_o, pats, inv, rc).match_soa(bases, offs, threads=8)
            assert cnt == ocnt and (bits == (flags != 0)).all()
