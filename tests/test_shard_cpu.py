"""world_size-2 gloo test of the multi-GPU host logic (sharding + gather).  The per-shard matching is done by the
oracle here (no GPU in this container); on GPUs the same shard/gather code wraps the CUDA path (bench.py --gpus N)."""
import os
import subprocess
import sys

import numpy as np

from fqgrep_b200 import shard
from tests import synth
from tests.fq_util import fastq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from fqgrep_b200 import shard
from oracle import oracle as O
from tests import synth
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
n = 20011
mat = synth.bases(n, seed=1)
synth.plant(mat, ["GACGAGATTA"], rate=0.02)
pats = ["GACGAGATTA", "GACGTGATTA"]
m = O.Matcher(O.REGEX_SET, pats, False, True)
lo, hi = shard.shard_bounds(n, world)[rank]
bases, offs = synth.soa(mat[lo:hi])
cnt, flags = m.match_soa(bases, offs)
words = np.frombuffer(np.packbits(flags.astype(bool), bitorder="little").tobytes() + b"\0" * 8, dtype=np.uint8)
words = words[: (words.size // 4) * 4].view(np.uint32)
total = shard.gather_count(cnt, dist)
allbits = shard.gather_masks(words, hi - lo, dist)
full_cnt, full_flags = m.match_soa(*synth.soa(mat))
assert total == full_cnt, (total, full_cnt)
assert allbits.size == n and (allbits == (full_flags != 0)).all()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", total)
'''


def test_shard_bounds_are_contiguous_and_word_aligned():
    for n in (0, 1, 31, 32, 33, 1000, 50_000_000):
        for w in (1, 2, 4, 8):
            b = shard.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and lo <= hi and hi % 32 == 0


def test_fastq_cuts_land_on_record_starts():
    rng = np.random.default_rng(3)
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, int(L))) for L in rng.integers(1, 300, 2000)]
    data = fastq(seqs, qual_char="@")     # quality lines full of '@' must not fool the cutter
    starts = set()
    pos = 0
    for s in seqs:
        starts.add(pos)
        pos += len(fastq([s], heads=["@%d" % len(starts)]))
    # recompute true starts from the real buffer
    true_starts, p = [], 0
    lines = data.split(b"\n")
    off = 0
    for i, ln in enumerate(lines[:-1]):
        if i % 4 == 0:
            true_starts.append(off)
        off += len(ln) + 1
    for parts in (1, 2, 3, 8):
        cuts = shard.fastq_cuts(data, parts)
        assert cuts[0] == 0 and cuts[-1] == len(data) and cuts == sorted(cuts)
        for c in cuts[1:-1]:
            assert c in true_starts


def test_two_rank_gloo_shard_and_gather(tmp_path):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
