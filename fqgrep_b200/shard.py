"""Multi-GPU plumbing: contiguous sharding of reads across ranks and the host-side gather of counts / masks.

The path is embarrassingly parallel: one process per GPU, every rank runs the same kernels on its
own contiguous range, and the only exchange is a gather of `--count` partial sums and selection masks
(torch.distributed; NCCL on GPUs, gloo in the CPU tests).  No collective sits on the data path.
"""
import ctypes as C

import numpy as np

from . import _ffi


def shard_bounds(n_units, world):
    """Contiguous [lo, hi) per rank; interior cuts are multiples of 32 so per-rank mask words concatenate."""
    cuts = [0]
    for r in range(1, world):
        c = (n_units * r // world) & ~31
        cuts.append(max(c, cuts[-1]))
    cuts.append(n_units)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def fastq_cuts(buf, parts):
    """Byte offsets of `parts` contiguous ranges of a FASTQ buffer that begin at record starts."""
    a = buf if isinstance(buf, np.ndarray) else np.frombuffer(bytes(buf), dtype=np.uint8)
    cuts = np.zeros(parts + 1, dtype=np.uint64)
    rc = _ffi.lib().fqg_fastq_split(a.ctypes.data_as(C.c_void_p), a.size, parts, cuts.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(_ffi.last_error())
    return [int(c) for c in cuts]


def gather_count(local_count, dist=None, device=None):
    """Sum of per-rank selected counts (what --count prints, src/main.rs:667-671)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(local_count)
    import torch
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_masks(local_mask_words, n_local_units, dist=None):
    """Rank-ordered concatenation of selection bits -> bool array over all units (rank 0 .. world-1)."""
    bits = np.unpackbits(np.ascontiguousarray(local_mask_words, dtype=np.uint32).view(np.uint8), bitorder="little")[:n_local_units]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bits.astype(bool)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.packbits(bits, bitorder="little").tobytes() + n_local_units.to_bytes(8, "little"))
    out = []
    for p in parts:
        n = int.from_bytes(p[-8:], "little")
        out.append(np.unpackbits(np.frombuffer(p[:-8], dtype=np.uint8), bitorder="little")[:n])
    return np.concatenate(out).astype(bool)
