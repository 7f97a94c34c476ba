"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle and the reference's
golden vectors.  Bit-exact (integer / byte work): the bar is equality of every selection bit and of --count."""
import random

import numpy as np
import pytest

import fqgrep_b200 as F
from oracle import oracle as O
from tests import synth
from tests.fq_util import soa

pytestmark = pytest.mark.gpu


def opts(inv, rc):
    return F.MatcherOpts(invert_match=inv, reverse_complement=rc)


def check(kind, pats, inv, rc, bases, offs):
    m = F.Matcher(kind, pats, opts(inv, rc))
    bits, cnt = m.match_bases(bases, offs)
    ocnt, flags = O.Matcher(kind, pats, inv, rc).match_soa(bases, offs, threads=8)
    bad = np.nonzero(bits != (flags != 0))[0]
    assert bad.size == 0, (kind, pats[:3], inv, rc, bad[:10])
    assert cnt == ocnt
    return cnt


def test_kat_all_matchers_on_gpu(kat):
    for rc, pat, seq, exp in kat["fixed_string"]["cases"]:
        for inv in (False, True):
            assert F.MatcherFactory.new_matcher(pat, True, [], None, opts(inv, rc)).read_match(seq) == (exp != inv)
    for rc, pats, seq, exp in kat["fixed_string_set"]["cases"]:
        for inv in (False, True):
            assert F.MatcherFactory.new_matcher(None, True, pats, None, opts(inv, rc)).read_match(seq) == (exp != inv)
    for rc, pat, seq, exp in kat["regex"]["cases"]:
        for inv in (False, True):
            assert F.MatcherFactory.new_matcher(pat, False, [], None, opts(inv, rc)).read_match(seq) == (exp != inv)
    for rc, pats, seq, exp in kat["regex_set"]["cases"]:
        for inv in (False, True):
            assert F.MatcherFactory.new_matcher(None, False, pats, None, opts(inv, rc)).read_match(seq) == (exp != inv)
    for pat, seq, fwd, rcv in kat["rc_special"]["cases"]:
        assert F.Matcher(F.FIXED, [pat], opts(False, False)).read_match(seq) == fwd
        assert F.Matcher(F.FIXED, [pat], opts(False, True)).read_match(seq) == rcv


@pytest.mark.parametrize("kind,pats,inv,rc", [
    (F.REGEX, ["GACGAGATTA"], False, False),
    (F.FIXED, ["GACGAGATTA"], False, False),
    (F.FIXED, ["GACGAGATTA"], True, True),
    (F.REGEX_SET, ["GACGAGATTA", "GACGTGATTA"], False, True),
    (F.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"], True, False),
    (F.REGEX_SET, ["^ACG", "TTT$", "G.{3,5}C{2}A"], False, True),
])
def test_baseline_configs_200k_reads(kind, pats, inv, rc):
    n = 200_000
    mat = synth.bases(n, seed=1)
    synth.plant(mat, ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"], rate=0.01)
    bases, offs = synth.soa(mat)
    cnt = check(kind, pats, inv, rc, bases, offs)
    assert 0 < cnt < n


def test_tail_batches_and_tiny_inputs():
    mat = synth.bases(1000, seed=3)
    synth.plant(mat, ["GACGAGATTA"], rate=0.05)
    for n in (1, 2, 31, 32, 33, 63, 64, 65, 999):
        bases, offs = synth.soa(mat[:n])
        check(F.FIXED, ["GACGAGATTA"], False, True, bases, offs)
        check(F.FIXED, ["GACGAGATTA"], True, False, bases, offs)


def test_ragged_empty_and_unaligned_reads():
    rng = random.Random(5)
    seqs = []
    for i in range(5000):
        L = rng.choice([0, 0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 33, 64, 100, 149, 150, 151, 255, 300])
        seqs.append("".join(rng.choice("ACGTN") for _ in range(L)))
    bases, offs = soa(seqs)
    for kind, pats in ((F.FIXED_SET, ["ACG", "TTN", "GG"]), (F.REGEX_SET, ["^$", "^A.*T$", "N{2}"]), (F.REGEX, [""]), (F.FIXED, [""])):
        for inv, rc in ((False, False), (True, True)):
            check(kind, pats, inv, rc, bases, offs)


def test_long_reads_take_the_global_memory_path():
    rng = np.random.default_rng(11)
    lens = [20000, 150, 150, 70000, 1, 0, 9000] + [150] * 100 + [30000]
    seqs = [bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)]) for L in lens]
    seqs[3] = seqs[3][:-10] + b"GACGAGATTA"
    bases, offs = soa(seqs)
    cnt = check(F.FIXED, ["GACGAGATTA"], False, False, bases, offs)
    assert cnt >= 1


def test_all_three_table_tiers():
    mat = synth.bases(50_000, seed=9)
    rng = np.random.default_rng(3)
    small = ["ACGTAC", "TTTTT"]
    medium = ["".join("ACGT"[i] for i in rng.integers(0, 4, 7)) for _ in range(300)]      # too short for the q-gram filter -> tier 4
    medium2 = ["".join("ACGT"[i] for i in rng.integers(0, 4, 7)) for _ in range(1500)]    # too many states for 2-base rows -> tier 1
    large = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(10_000)]
    synth.plant(mat, small + medium[:20] + large[:200], rate=0.05)
    bases, offs = synth.soa(mat)
    tiers = []
    for pats in (small, medium, medium2, large):
        m = F.Matcher(F.FIXED_SET, pats, opts(False, True))
        tiers.append(m.dfa_info().tier)
        cnt = check(F.FIXED_SET, pats, False, True, bases, offs)
        assert cnt > 0
    assert tiers == [0, 4, 1, 3]     # 4-base rows, 2-base rows, class-indexed, q-gram filter


def test_tier4_two_base_rows_with_non_acgt_reads_and_regex_sets():
    rng = np.random.default_rng(31)
    pats = ["".join("ACGT"[i] for i in rng.integers(0, 4, 8)) + "[AC]N?" + "".join("ACGT"[i] for i in rng.integers(0, 4, 6)) for _ in range(25)]
    assert F.Matcher(F.REGEX_SET, pats, opts(False, True)).dfa_info().tier == 4
    mat = synth.bases(60_000, seed=8)
    mat[::7, 30] = ord("N")
    mat[::11, 3:9] = np.frombuffer(b"acgtnn", dtype=np.uint8)
    lit = [p.replace("[AC]N?", "A") for p in pats[:10]] + [p.replace("[AC]N?", "CN") for p in pats[10:20]]
    lit = [x for x in lit if x]
    synth.plant(mat, lit, rate=0.05)
    bases, offs = synth.soa(mat)
    for inv, rc in ((False, False), (False, True), (True, True)):
        cnt = check(F.REGEX_SET, pats, inv, rc, bases, offs)
        assert 0 < cnt < 60_000


def test_mixed_literal_and_regex_set_is_split_into_two_passes():
    """-f file with hundreds of plain literals plus a few real regexes: literal part via the q-gram filter, regex part via
    a small automaton, found-bits OR-ed before invert_match (tier 5)."""
    rng = np.random.default_rng(41)
    lits = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(250)]
    pats = lits + ["AC.G[AT]CA{2}T", "^TTTT", "GG$"]
    assert F.Matcher(F.REGEX_SET, pats, opts(False, True)).dfa_info().tier == 5
    mat = synth.bases(80_000, seed=14)
    synth.plant(mat, lits[:30], rate=0.03)
    bases, offs = synth.soa(mat)
    for inv, rc in ((False, False), (False, True), (True, False), (True, True)):
        cnt = check(F.REGEX_SET, pats, inv, rc, bases, offs)
        assert 0 < cnt < 80_000


def test_tier2_l2_automaton_for_large_sets_of_short_patterns():
    rng = np.random.default_rng(12)
    pats = list({"".join("ACGTN"[i] for i in rng.integers(0, 5, 7)) for _ in range(14000)})     # 7-mers: too short for the q-gram filter
    m = F.Matcher(F.FIXED_SET, pats, opts(False, False))
    assert m.dfa_info().tier == 2
    mat = synth.bases(20_000, L=40, seed=3)
    mat[::3, 5] = ord("N")
    bases, offs = synth.soa(mat)
    cnt = check(F.FIXED_SET, pats, False, False, bases, offs)
    assert 0 < cnt < 20_000


def test_gram_filter_edge_cases(monkeypatch):
    """Tier 3: variable pattern lengths, reverse complement, N in patterns and reads, candidate overflow, ragged reads.
    4000 patterns as short as 9 bases make a filter that verifies several grams per read; the library would rather walk
    the automaton for such a set (density guard in gram_filter.cc), so the guard is lifted here to keep the unaligned
    stride-2 scan and the many-hits verification rounds under test."""
    rng = np.random.default_rng(21)
    pats = ["".join("ACGT"[i] for i in rng.integers(0, 4, int(L))) for L in rng.integers(9, 40, 4000)]
    pats += ["ACGTNNNNACGTAC", "NNNNNNNNNN", "GATTACAGATTACAGATTACA"]
    assert F.Matcher(F.FIXED_SET, pats).dfa_info().tier == 2
    monkeypatch.setenv("FQG_GRAM_NO_DENSITY_GUARD", "1")
    for rc in (False, True):
        m = F.Matcher(F.FIXED_SET, pats, opts(False, rc))
        assert m.dfa_info().tier == 3
    seqs = []
    r = random.Random(4)
    for i in range(6000):
        L = r.choice([0, 5, 8, 9, 10, 17, 33, 64, 150, 151, 250])
        s = "".join(r.choice("ACGTN" if i % 5 == 0 else "ACGT") for _ in range(L))
        if i % 11 == 0 and L >= 40:
            p = pats[r.randrange(len(pats))]
            if len(p) <= L:
                o = r.randrange(L - len(p) + 1)
                s = s[:o] + p + s[o + len(p):]
        if i % 13 == 0 and L >= 40:      # reverse-complement plant
            p = O.revcomp(pats[r.randrange(4000)]).decode()
            if len(p) <= L:
                o = r.randrange(L - len(p) + 1)
                s = s[:o] + p + s[o + len(p):]
        seqs.append(s)
    # reads made of pattern prefixes: many filter hits, few real matches (exercises the candidate overflow fallback)
    for i in range(300):
        seqs.append("".join(pats[r.randrange(4000)][:12] for _ in range(12)))
    seqs.append("N" * 150)
    seqs.append("ACGTNNNNACGTAC")
    for L in (600, 2000):        # more than 64 sampled grams: exact automaton walk
        body = "".join(r.choice("ACGT") for _ in range(L))
        seqs.append(body)
        seqs.append(body[: L - 30] + pats[7] + body[L - 30 + len(pats[7]):] if len(pats[7]) <= 30 else body + pats[7])
    bases, offs = soa(seqs)
    for inv, rc in ((False, False), (False, True), (True, True)):
        cnt = check(F.FIXED_SET, pats, inv, rc, bases, offs)
        assert 0 < cnt < len(seqs)
    # with the guard in place: a set small enough for a stride-2 filter, too big for the 2-base-stride table
    monkeypatch.delenv("FQG_GRAM_NO_DENSITY_GUARD")
    few = pats[:120] + ["NNNNNNNNNN"]
    assert F.Matcher(F.FIXED_SET, few).dfa_info().tier == 3
    assert check(F.FIXED_SET, few, False, False, bases, offs) > 0


def test_protein_alphabet(kat):
    e = kat["e2e_protein"]
    bases, offs = soa(e["reads"])
    for fixed, pats, exp in e["cases"]:
        m = F.Matcher(F.FIXED_SET if fixed else F.REGEX_SET, pats)
        bits, cnt = m.match_bases(bases, offs)
        assert [r for r, b in zip(e["reads"], bits) if b] == exp


def test_device_resident_entry_with_pair_or_mask():
    """fqg_match_bases_device with torch-owned HBM buffers, mate OR-mask and device-side count (main.rs:749-755)."""
    import ctypes as C
    import torch
    from fqgrep_b200 import _ffi
    n = 100_003
    m1, m2 = synth.bases(n, seed=1), synth.bases(n, seed=2)
    synth.plant(m1, ["GACGAGATTA"], rate=0.01, seed=5)
    synth.plant(m2, ["GACGTGATTA"], rate=0.01, seed=6)
    pats = ["GACGAGATTA", "GACGTGATTA"]
    m = F.Matcher(F.REGEX_SET, pats, opts(False, True))
    o = O.Matcher(O.REGEX_SET, pats, False, True)
    dev = torch.device("cuda:0")
    words = (n + 31) // 32
    outs = []
    mask1 = torch.zeros(words, dtype=torch.int32, device=dev)
    mask2 = torch.zeros(words, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx = F.context(0)
    for mat, mask, orm in ((m1, mask1, None), (m2, mask2, mask1)):
        b, offs = synth.soa(mat)
        tb = torch.zeros(b.size + 64, dtype=torch.uint8, device=dev)
        tb[: b.size] = torch.from_numpy(b).to(dev)
        to = torch.from_numpy(offs.astype(np.int64)).to(dev).to(torch.int32)
        count.zero_()
        torch.cuda.synchronize()
        rc = _ffi.lib().fqg_match_bases_device(ctx, m._h, tb.data_ptr(), to.data_ptr(), to.data_ptr() + 4, n, mask.data_ptr(),
                                               orm.data_ptr() if orm is not None else None, count.data_ptr(), 150,
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, _ffi.last_error()
        torch.cuda.synchronize()
        outs.append((tb, to))
    f1 = o.match_soa(*synth.soa(m1), threads=8)[1] != 0
    f2 = o.match_soa(*synth.soa(m2), threads=8)[1] != 0
    got = np.unpackbits(mask2.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (got == (f1 | f2)).all()
    assert int(count.item()) == int((f1 | f2).sum())


def test_fuzz_random_pattern_sets_on_ragged_mixed_alphabet_reads():
    """Random regex / literal / bit-mask sets against reads with N, lowercase, protein letters and ragged lengths:
    exercises the 4-base and 2-base table paths, their exact non-ACGT fallback, the tails and the class-indexed tiers."""
    import os
    rng = random.Random(int(os.environ.get("FQG_FUZZ_SEED", "77")))      # longer campaigns: FQG_FUZZ_SEED / FQG_FUZZ_ITERS
    seqs = []
    for i in range(12000):
        L = rng.choice([0, 1, 3, 4, 5, 8, 16, 31, 32, 33, 75, 150, 151, 152, 153, 301])
        alpha = "ACGT" if i % 3 else rng.choice(["ACGTN", "ACGTacgt", "ACGTNRYKM", "ARNDCEQGHILKMFPSTWYV"])
        seqs.append("".join(rng.choice(alpha) for _ in range(L)))
    bases, offs = soa(seqs)
    atoms = ["A", "C", "G", "T", "N", ".", "[AC]", "[^G]", "[ACGT]", "(A|T)", "(CG|GC)", "a", "[Nn]"]
    quants = ["", "", "", "", "*", "+", "?", "{2}", "{1,3}"]
    for trial in range(int(os.environ.get("FQG_FUZZ_ITERS", "40"))):
        kind = rng.choice([F.REGEX_SET, F.REGEX_SET, F.FIXED_SET, F.BITMASK_SET])
        if kind == F.REGEX_SET:
            pats = []
            for _ in range(rng.randint(1, 12)):
                body = "".join(rng.choice(atoms) + rng.choice(quants) for _ in range(rng.randint(2, 9)))
                pats.append(rng.choice(["", "^", ""]) + body + rng.choice(["", "$", ""]))
        elif kind == F.FIXED_SET:
            lo, hi, cnt = rng.choice(((1, 14, 40), (8, 20, 400), (12, 24, 3000), (5, 9, 600)))      # small / q-gram filter / 2-base rows
            pats = ["".join(rng.choice("ACGTN") for _ in range(rng.randint(lo, hi))) for _ in range(rng.randint(1, cnt))]
        else:
            pats = ["".join(rng.choice("ACGTRYKMSWN") for _ in range(rng.randint(2, 10))) for _ in range(rng.randint(2, 6))]
        inv, rc = rng.random() < 0.3, rng.random() < 0.5
        try:
            O.Matcher(kind, pats, inv, rc)
        except O.OracleError:
            continue
        check(kind, pats, inv, rc, bases, offs)


def test_names_over_soa_headers():
    """QueryNameMatcher::read_match (matcher.rs:631-638) through the SoA entry point: the spans are header lines, the id
    ends at the first space.  Checked against the oracle and against plain Python set membership."""
    rng = random.Random(11)
    alphabet = "abcXYZ0189._:/-"
    ids = ["".join(rng.choice(alphabet) for _ in range(rng.choice((0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 31, 40)))) for _ in range(6000)]
    ids += ["read1", "read1_extra", "SRR001666.1", "r0000000001", "r0000000002"]
    chosen = sorted(set(rng.sample(ids, 2500)) | {"read1", "SRR001666.1"})
    heads = []
    for i in range(40_037):        # ragged: not a multiple of 32
        h = rng.choice(ids)
        k = rng.random()
        if k < 0.3:
            h += " " + "".join(rng.choice(alphabet + " ") for _ in range(rng.randrange(0, 20)))      # description after a space
        elif k < 0.35:
            h = h[: max(0, len(h) - 1)]          # near miss (a prefix)
        elif k < 0.4:
            h += "x"                             # near miss (an extension)
        heads.append(h.encode())
    blob = np.frombuffer(b"".join(heads), dtype=np.uint8)
    offs = np.zeros(len(heads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(h) for h in heads])
    want = np.array([h.split(b" ")[0].decode() in set(chosen) for h in heads])
    for inv in (False, True):
        m = F.Matcher(F.NAMES, chosen, opts(inv, False))
        bits, cnt = m.match_bases(blob, offs)
        ocnt, flags = O.Matcher(F.NAMES, chosen, inv, False).match_soa(blob, offs, threads=4)
        assert (bits == (want != inv)).all()
        assert (bits == (flags != 0)).all() and cnt == ocnt == int((want != inv).sum())


def test_regexes_with_finite_languages_take_the_gram_filter():
    """Members like `ACGT(A|C)G{1,2}[GT]...` have a small finite language: the set is multiplied out for the q-gram filter
    (the automaton stays the exact fallback).  Every construct of the expansion is exercised and checked against the
    oracle's NFA, with reads that carry real occurrences of random members of the languages."""
    rng = random.Random(31)

    def lit(n):
        return "".join(rng.choice("ACGT") for _ in range(n))

    pats, samples = [], []
    for i in range(400):
        a, b, c = lit(rng.randint(6, 10)), lit(rng.randint(5, 9)), lit(rng.randint(4, 8))
        kind = i % 5
        if kind == 0:
            pats.append(a + "[AG]" + b + "(C|T)" + c); samples.append(a + "G" + b + "T" + c)
        elif kind == 1:
            pats.append(a + "(AC|GT|T)" + b + c); samples.append(a + "GT" + b + c)
        elif kind == 2:
            pats.append(a + "N{1,2}" + b + "[CT]" + c); samples.append(a + "NN" + b + "C" + c)
        elif kind == 3:
            pats.append(a + b + "(?:A|C){2}" + c); samples.append(a + b + "CA" + c)
        else:
            pats.append(a + "G?" + b + c + "[ACGT]"); samples.append(a + b + c + "T")
    for rc in (False, True):
        assert F.Matcher(F.REGEX_SET, pats, opts(False, rc)).dfa_info().tier == 3
    seqs = []
    for i in range(20000):
        L = rng.choice([20, 36, 75, 150, 151, 250])
        s = "".join(rng.choice("ACGT") for _ in range(L))
        if i % 9 == 0:
            p = samples[rng.randrange(len(samples))]
            if i % 18 == 0:
                p = O.revcomp(p).decode()
            if len(p) <= L:
                o = rng.randrange(L - len(p) + 1)
                s = s[:o] + p + s[o + len(p):]
        seqs.append(s)
    bases, offs = soa(seqs)
    for inv, rc in ((False, False), (False, True), (True, True)):
        cnt = check(F.REGEX_SET, pats, inv, rc, bases, offs)
        assert 0 < cnt < len(seqs)
    # a member without a finite language (or with the empty string in it) keeps the whole set on the automaton / split path
    assert F.Matcher(F.REGEX_SET, pats + ["ACGTACGTAC.*TTTTGGGGCC"]).dfa_info().tier in (2, 5)
    assert F.Matcher(F.REGEX_SET, pats[:3] + ["A*"]).dfa_info().tier in (0, 1, 2, 4)


def test_anchored_sets_stop_at_settled_states():
    """^-anchored pattern sets (barcodes at the start of the read) have a dead state: the slower tiers stop walking a read
    once it is dead or accepted for good.  Same selection as the oracle, with and without -v / reverse complement, on
    reads that also carry N and lowercase (the exact non-ACGT path) and on reads shorter than the patterns."""
    rng = random.Random(41)
    codes = ["".join(rng.choice("ACGT") for _ in range(rng.randint(8, 14))) for _ in range(1500)]
    pats = ["^" + c for c in codes]
    info = F.Matcher(F.REGEX_SET, pats).dfa_info()
    assert info.tier in (1, 2, 4)
    seqs = []
    for i in range(30000):
        L = rng.choice([0, 3, 9, 20, 75, 150, 151])
        alpha = "ACGT" if i % 4 else rng.choice(["ACGTN", "ACGTacgt"])
        s = "".join(rng.choice(alpha) for _ in range(L))
        if i % 3 == 0:
            c = codes[rng.randrange(len(codes))]
            s = (c + s)[:max(L, len(c))] if i % 6 else s[:5] + c + s[5:]        # at the start (a match) or further in (not one)
        seqs.append(s)
    bases, offs = soa(seqs)
    for inv, rc in ((False, False), (True, False), (False, True)):
        cnt = check(F.REGEX_SET, pats, inv, rc, bases, offs)
        assert 0 < cnt < len(seqs)
    # a set that mixes anchored and unanchored members has no dead state: nothing to stop at, same answers
    check(F.REGEX_SET, pats[:200] + ["GGGGGGGG"], False, False, bases, offs)
