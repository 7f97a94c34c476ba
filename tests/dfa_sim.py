"""Test-side simulator of the DFA the product compiles (exported through fqg_matcher_dfa_export).
Lets the CPU suite check the host compilers against the oracle without a GPU.  Not product code."""
import numpy as np


def run_dfa(matcher, seqs):
    """-> list of bool: read_match per sequence, computed by walking the exported table."""
    cls, nxt, start, accept_from = matcher.dfa_export()
    inv = matcher.opts().invert_match
    out = []
    for s in seqs:
        b = s.encode() if isinstance(s, str) else bytes(s)
        st = start
        for ch in b:
            st = nxt[st, cls[ch]]
        out.append((st >= accept_from) != inv)
    return out


def run_dfa_soa(matcher, bases, offsets):
    """Vectorised over reads of equal length; falls back to the scalar walk otherwise."""
    cls, nxt, start, accept_from = matcher.dfa_export()
    inv = matcher.opts().invert_match
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = np.diff(offsets)
    n = lens.size
    if n and (lens == lens[0]).all() and offsets[0] == 0:
        L = int(lens[0])
        mat = cls[np.asarray(bases[: n * L]).reshape(n, L)]
        st = np.full(n, start, dtype=np.int64)
        flat = nxt.astype(np.int64)
        for j in range(L):
            st = flat[st, mat[:, j]]
        return (st >= accept_from) != inv
    return np.array(run_dfa(matcher, [bytes(bases[offsets[i]:offsets[i + 1]]) for i in range(n)]), dtype=bool)
