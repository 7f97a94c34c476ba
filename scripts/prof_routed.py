#!/usr/bin/env python
"""A tier-3 pattern set (10,000 x 20-mers, configs[2]) over raw FASTQ resident in HBM: the routed path (index -> compact -> SoA
kernel with the q-gram filter), a few calls, for the ncu launch list."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fqgrep_b200 as F          # noqa: E402
from fqgrep_b200 import _ffi     # noqa: E402

L = _ffi.lib()
dev = torch.device("cuda:0")
ctx = F.context(0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rng = np.random.default_rng(7)
pats = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(10_000)]
blob = "".join(pats[:200]).encode()
offs = (np.arange(201, dtype=np.uint32) * 20)
t = torch.empty(n * 317 + 64, dtype=torch.uint8, device=dev)
assert L.fqg_synth_reads_device(ctx, t.data_ptr(), n, 0, 150, 1, 1, blob, offs.ctypes.data_as(C.c_void_p), 200, 0.01, 99, sp) == 0
m = F.Matcher(F.FIXED_SET, pats)
words = n // 32 + 2
mask = torch.zeros(words, dtype=torch.int32, device=dev)
res = _ffi.Result()
ms = []
for _ in range(4):
    assert L.fqg_grep_fastq_device(ctx, m._h, 0, t.data_ptr(), n * 317, None, 0, mask.data_ptr(), words, C.byref(res), sp) == 0, _ffi.last_error()
    ms.append(res.kernel_ms)
print("routed tier", m.dfa_info().tier, "selected", res.n_selected, "ms", ms, "frac", 325.125 * n / (min(ms) / 1e3) / 1e9 / 6558.1)
