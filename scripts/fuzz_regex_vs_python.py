#!/usr/bin/env python
"""Differential campaign, CPU only: random regexes (classes, ranges, Perl classes, groups, alternation, greedy and lazy counted
repetition, anchors, word boundaries, (?i)) compiled by the product and simulated from the exported DFA (tests/dfa_sim.py),
against Python re on ASCII haystacks without newlines -- a third implementation that shares no code with the product or the
oracle.  Usage: fuzz_regex_vs_python.py SEED ITERATIONS.  Known, excluded difference: Python's \\B never matches in an empty
haystack.  "REJECTED" lines are patterns whose DFA exceeds the documented 2^20-state limit."""
import os
import random
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fqgrep_b200 as F
from tests.dfa_sim import run_dfa
seed = int(sys.argv[1]); iters = int(sys.argv[2])
rng = random.Random(seed)
LIT = list("ACGTNacgt09_ -")
def atom(d):
    r = rng.random()
    if r < 0.45: return re.escape(rng.choice(LIT)) if rng.random() < 0.1 else rng.choice("ACGTNacgt0_")
    if r < 0.55: return "."
    if r < 0.70:
        items = []
        for _ in range(rng.randint(1, 4)):
            if rng.random() < 0.3:
                a, b = sorted(rng.sample("ACGTNZaz09", 2)); items.append(a + "-" + b)
            else: items.append(rng.choice(["A","C","G","T","N","a","_","0",r"\d",r"\w",r"\s"]))
        return "[" + ("^" if rng.random() < 0.3 else "") + "".join(items) + "]"
    if r < 0.80: return rng.choice([r"\d", r"\w", r"\s", r"\D", r"\W", r"\S"])
    if r < 0.86: return rng.choice([r"\b", r"\B", "^", "$", r"\A"])
    if d > 2: return rng.choice("ACGT")
    kind = rng.choice(["(", "(?:"])
    return kind + alt(d + 1) + ")"
def piece(d):
    a = atom(d)
    if a in (r"\b", r"\B", "^", "$", r"\A"): return a
    q = rng.choice(["", "", "", "*", "+", "?", "{2}", "{1,3}", "{0,2}", "{2,}", "*?", "+?", "??"])
    return a + q
def concat(d): return "".join(piece(d) for _ in range(rng.randint(1, 5)))
def alt(d): return "|".join(concat(d) for _ in range(rng.choice([1, 1, 1, 2, 3])))
alph = ["ACGT", "ACGTN", "ACGTacgt", "ACGT_ 09-", "AaCc09_ "]
hs = [b"", b"A", b"a", b" ", b"_", b"0"]
for _ in range(500):
    a = rng.choice(alph); hs.append("".join(rng.choice(a) for _ in range(rng.randint(0, 20))).encode())
bad = 0
for it in range(iters):
    pat = alt(0)
    if rng.random() < 0.15: pat = "(?i)" + pat
    try:
        pr = re.compile(pat.encode())
    except re.error:
        continue
    try:
        m = F.Matcher(F.REGEX, [pat])
    except F.FqgrepError as e:
        print("REJECTED", repr(pat), e.message); bad += 1; continue
    got = run_dfa(m, hs)
    for h, g in zip(hs, got):
        e = pr.search(h) is not None
        if g != e and not (h == b"" and "\\B" in pat):
            print("MISMATCH", repr(pat), h, "dfa", g, "re", e); bad += 1; break
    if bad > 15: break
print("done", seed, "bad", bad)
