#!/usr/bin/env python
"""Times fqg_grep_fastq_device (single-end, configs[1] patterns, 8 M records resident in HBM) with CUDA events around the call,
whatever the call returns: used for A/B runs of library variants (FQG_LIB_VARIANT), including measurement-only variants
whose results are wrong on purpose."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np               # noqa: E402
import fqgrep_b200 as F          # noqa: E402
from fqgrep_b200 import _ffi     # noqa: E402

L = _ffi.lib()
dev = torch.device("cuda:0")
ctx = F.context(0)
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
RL = int(sys.argv[2]) if len(sys.argv) > 2 else 150      # read length; a record is 2 * RL + 17 bytes
pats = ["GACGAGATTA", "GACGTGATTA"]
blob = "".join(pats).encode()
offs = np.array([0, 10, 20], dtype=np.uint32)
REC = 2 * RL + 17
t = torch.empty(n * REC + 64, dtype=torch.uint8, device=dev)
assert L.fqg_synth_reads_device(ctx, t.data_ptr(), n, 0, RL, 1, 1, blob, offs.ctypes.data_as(C.c_void_p), 2, 0.01, 99, sp) == 0
m = F.MatcherFactory.new_matcher(None, False, pats, None, F.MatcherOpts(False, True))
words = n // 32 + 2
mask = torch.zeros(words, dtype=torch.int32, device=dev)
res = _ffi.Result()
ms, rcs = [], []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rcs.append(L.fqg_grep_fastq_device(ctx, m._h, 0, t.data_ptr(), n * REC, None, 0, mask.data_ptr(), words, C.byref(res), sp))
    e1.record(stream)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
best = min(ms[1:])
print(os.environ.get("FQG_LIB_VARIANT", "default").split("/")[-1], "rc", rcs[-1], "selected", res.n_selected, "ms %.3f" % best, "frac %.4f" % ((REC + 8.125) * n / (best / 1e3) / 1e9 / 6558.1), "read length", RL)
