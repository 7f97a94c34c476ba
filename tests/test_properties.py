"""The reference's property tests (src/lib/matcher.rs:1379-1486, proptest), restated with hypothesis against the
product's compiled automata (walked by the test-side simulator) and against the oracle.  CPU only."""
from hypothesis import given, settings, strategies as st

import fqgrep_b200 as F
from oracle import oracle as O
from tests.dfa_sim import run_dfa

dna_sequence = st.text(alphabet="ACGT", min_size=1, max_size=100)
dna_pattern = st.text(alphabet="ACGT", min_size=1, max_size=10)
any_bases = st.text(alphabet="ACGTNacgtn", min_size=0, max_size=60)


def opts(inv=False, rc=False):
    return F.MatcherOpts(invert_match=inv, reverse_complement=rc)


@settings(max_examples=60, deadline=None)
@given(any_bases)
def test_reverse_complement_involutive(seq):
    assert O.revcomp(O.revcomp(seq)) == seq.encode()


@settings(max_examples=40, deadline=None)
@given(dna_sequence)
def test_sequence_matches_itself(seq):
    assert run_dfa(F.Matcher(F.FIXED, [seq], opts()), [seq])[0]
    assert O.Matcher(O.FIXED, [seq]).bases_match(seq)


@settings(max_examples=60, deadline=None)
@given(dna_sequence, dna_pattern)
def test_rc_matching_consistent(seq, pattern):
    """A forward match implies a match with reverse_complement enabled; and rc-enabled == fwd(seq) or fwd(rc(seq))."""
    fwd = F.Matcher(F.FIXED, [pattern], opts(False, False))
    both = F.Matcher(F.FIXED, [pattern], opts(False, True))
    m_fwd, m_rc_only = run_dfa(fwd, [seq, O.revcomp(seq)])
    m_both = run_dfa(both, [seq])[0]
    assert (not m_fwd) or m_both
    assert m_both == (m_fwd or m_rc_only)
    assert m_both == O.Matcher(O.FIXED, [pattern], False, True).read_match(seq)


@settings(max_examples=60, deadline=None)
@given(dna_sequence, dna_pattern, st.booleans())
def test_invert_match_negates(seq, pattern, rc):
    normal = run_dfa(F.Matcher(F.FIXED, [pattern], opts(False, rc)), [seq])[0]
    inverted = run_dfa(F.Matcher(F.FIXED, [pattern], opts(True, rc)), [seq])[0]
    assert normal != inverted


@settings(max_examples=60, deadline=None)
@given(dna_sequence, dna_pattern, st.booleans(), st.booleans())
def test_fixed_and_regex_agree(seq, pattern, inv, rc):
    fixed = run_dfa(F.Matcher(F.FIXED, [pattern], opts(inv, rc)), [seq])[0]
    regex = run_dfa(F.Matcher(F.REGEX, [pattern], opts(inv, rc)), [seq])[0]
    fset = run_dfa(F.Matcher(F.FIXED_SET, [pattern], opts(inv, rc)), [seq])[0]
    rset = run_dfa(F.Matcher(F.REGEX_SET, [pattern], opts(inv, rc)), [seq])[0]
    bitm = run_dfa(F.Matcher(F.BITMASK, [pattern], opts(inv, rc)), [seq])[0]
    assert fixed == regex == fset == rset == bitm == O.Matcher(O.REGEX, [pattern], inv, rc).read_match(seq)


@settings(max_examples=40, deadline=None)
@given(st.lists(dna_pattern, min_size=1, max_size=5), st.lists(any_bases, min_size=1, max_size=8), st.booleans(), st.booleans())
def test_sets_equal_union_of_members(patterns, seqs, inv, rc):
    got = run_dfa(F.Matcher(F.FIXED_SET, patterns, opts(inv, rc)), seqs)
    members = [run_dfa(F.Matcher(F.FIXED, [p], opts(False, rc)), seqs) for p in patterns]
    for i in range(len(seqs)):
        assert got[i] == (any(m[i] for m in members) != inv)
