#!/usr/bin/env python
"""-N over raw FASTQ resident in HBM (configs[4]b shape): a few calls of the fused kernel in read-name mode, for ncu."""
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
n, n_names = 4_000_000, 1_000_000
t = torch.empty(n * 317 + 64, dtype=torch.uint8, device=dev)
assert L.fqg_synth_reads_device(ctx, t.data_ptr(), n, 0, 150, 1, 1, b"", np.zeros(1, dtype=np.uint32).ctypes.data_as(C.c_void_p), 0, 0.0, 0, sp) == 0
rng = np.random.default_rng(5)
chosen = rng.choice(n, 900_000, replace=False)
names = ["r%010d" % i for i in chosen] + ["r%010d" % (n + i) for i in range(n_names - chosen.size)]
m = F.Matcher(F.NAMES, names)
words = n // 32 + 2
mask = torch.zeros(words, dtype=torch.int32, device=dev)
res = _ffi.Result()
ms = []
for _ in range(5):
    assert L.fqg_grep_fastq_device(ctx, m._h, 0, t.data_ptr(), n * 317, None, 0, mask.data_ptr(), words, C.byref(res), sp) == 0, _ffi.last_error()
    ms.append(res.kernel_ms)
print("names fastq: selected", res.n_selected, "expected", chosen.size, "ms", ms, "Greads/s", n / (min(ms) / 1e3) / 1e9, "frac", 325.125 * n / (min(ms) / 1e3) / 1e9 / 6558.1)
