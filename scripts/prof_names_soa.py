#!/usr/bin/env python
"""-N over SoA header ids resident in HBM (configs[4]b shape, 50 M reads, 1 M names): a few calls for ncu."""
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
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
n, n_names, idl = 50_000_000, 1_000_000, 11
idx = torch.arange(n, device=dev, dtype=torch.int64)
hb = torch.empty((n, idl), dtype=torch.uint8, device=dev)
hb[:, 0] = ord("r")
for k in range(10):
    hb[:, 10 - k] = ((idx // (10 ** k)) % 10 + 48).to(torch.uint8)
del idx
offs = (torch.arange(n + 1, device=dev, dtype=torch.int64) * idl).to(torch.int32)
rng = np.random.default_rng(5)
chosen = rng.choice(n, 900_000, replace=False)
names = ["r%010d" % i for i in chosen] + ["r%010d" % (n + i) for i in range(n_names - chosen.size)]
m = F.Matcher(F.NAMES, names)
mask = torch.zeros(n // 32 + 2, dtype=torch.int32, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
ms = []
for _ in range(5):
    cnt.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    assert L.fqg_match_bases_device(ctx, m._h, hb.data_ptr(), offs.data_ptr(), offs.data_ptr() + 4, n, mask.data_ptr(), None, cnt.data_ptr(), idl, sp) == 0, _ffi.last_error()
    e1.record(stream)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print("names soa: selected", int(cnt.item()), "expected", chosen.size, "ms", ms, "Greads/s", n / (min(ms) / 1e3) / 1e9, "frac", 23.125 * n / (min(ms) / 1e3) / 1e9 / 6558.1)
