#!/usr/bin/env python
"""Host-side input decode rate (no GPU): fqg_read_input on the same FASTQ as one plain gzip stream (serial zlib) and as
BGZF (members inflated on all host threads).  Prints one JSON line; numbers are MB/s of DECOMPRESSED bytes."""
import gzip
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fqgrep_b200 as F          # noqa: E402
from tests import synth          # noqa: E402
from tests.test_input import bgzf   # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000
    data = synth.fastq(synth.bases(n, seed=1)).tobytes()
    out = {"fastq_mb": len(data) / 1e6, "host_threads": os.cpu_count()}
    with tempfile.TemporaryDirectory() as d:
        for name, blob in (("plain_gzip", gzip.compress(data, 6)), ("bgzf", bgzf(data, block=65280))):
            p = os.path.join(d, name + ".fq.gz")
            open(p, "wb").write(blob)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                got = F.read_input(p)
                best = min(best, time.perf_counter() - t0)
            assert got.tobytes() == data
            out[name] = {"compressed_mb": len(blob) / 1e6, "seconds": best, "mb_per_s": len(data) / 1e6 / best}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
