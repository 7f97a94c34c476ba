#!/usr/bin/env python
"""Host-side ordered write-back rate (no GPU): fqg_write_selected over synthetic FASTQ with 1 % of the records selected,
serial walk (shard size forced above the input) vs the parallel walk.  Prints one JSON line; GB/s of INPUT bytes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fqgrep_b200 as F          # noqa: E402
from tests import synth          # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    data = synth.fastq(synth.bases(n, seed=1))
    sel = (np.arange(n) % 100 == 0)
    mask = np.frombuffer(np.packbits(sel, bitorder="little").tobytes() + b"\0" * 8, dtype=np.uint8)[: (n // 32 + 2) * 4].view(np.uint32)
    out = {"fastq_gb": data.size / 1e9, "records": n, "selected": int(sel.sum()), "host_threads": os.cpu_count()}
    ref = None
    for name, shard in (("serial", str(1 << 40)), ("parallel", "")):
        if shard:
            os.environ["FQG_WRITE_SHARD_BYTES"] = shard
        else:
            os.environ.pop("FQG_WRITE_SHARD_BYTES", None)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            got = F.write_selected(mask, data)
            best = min(best, time.perf_counter() - t0)
        ref = ref or got
        assert got == ref
        out[name] = {"seconds": best, "gb_per_s": data.size / 1e9 / best}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
