"""Synthetic reads exactly as SURVEY.md 8(d) specifies: record i = '@r%010d' / 150 uniform bases from
splitmix64(seed ^ GOLDEN*(i*L+j)) >> 62 / '+' / 150 x 'I'  (317 B per record).  numpy, deterministic.
Also used by bench.py; the same generator exists as a device kernel in the product for HBM-resident inputs."""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def splitmix64(x):
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += GOLDEN
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def bases(n, L=150, seed=1, first=0):
    """uint8 [n, L] of ACGT."""
    with np.errstate(over="ignore"):
        idx = (np.arange(first, first + n, dtype=np.uint64)[:, None] * np.uint64(L) + np.arange(L, dtype=np.uint64)[None, :])
        h = splitmix64(np.uint64(seed) ^ (GOLDEN * idx))
    return ACGT[(h >> np.uint64(62)).astype(np.int64)]


def plant(mat, patterns, rate=0.01, seed=99, first=0, revcomp_half=True):
    """Overwrite a pattern (or its reverse complement, 50/50) into a pseudo-random `rate` of the reads."""
    n, L = mat.shape
    comp = np.arange(256, dtype=np.uint8)
    for a, b in zip(b"AGCTYRWSKMDVHBN", b"TCGARYWSMKHBDVN"):
        comp[a] = b
        comp[a + 32] = b + 32
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(seed) ^ (GOLDEN * np.arange(first, first + n, dtype=np.uint64)))
    chosen = np.nonzero(h < np.uint64(int(rate * 2.0 ** 64)))[0]
    pats = [np.frombuffer(p.encode() if isinstance(p, str) else p, dtype=np.uint8) for p in patterns]
    for r in chosen:
        hv = int(h[r])
        p = pats[(hv >> 8) % len(pats)]
        if revcomp_half and (hv >> 7) & 1:
            p = comp[p[::-1]]
        if p.size > L:
            continue
        off = (hv >> 20) % (L - p.size + 1)
        mat[r, off:off + p.size] = p
    return chosen


def soa(mat):
    n, L = mat.shape
    return np.ascontiguousarray(mat).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))


def fastq(mat, first=0):
    """uint8 FASTQ bytes, 13 + L + 1 + 2 + L + 1 per record."""
    n, L = mat.shape
    rec = 13 + L + 1 + 2 + L + 1
    out = np.empty((n, rec), dtype=np.uint8)
    out[:, 0] = ord("@")
    out[:, 1] = ord("r")
    ids = np.arange(first, first + n, dtype=np.int64)
    for k in range(10):
        out[:, 11 - k] = (ids // 10 ** k % 10 + 48).astype(np.uint8)
    out[:, 12] = 10
    out[:, 13:13 + L] = mat
    out[:, 13 + L] = 10
    out[:, 14 + L] = ord("+")
    out[:, 15 + L] = 10
    out[:, 16 + L:16 + 2 * L] = ord("I")
    out[:, 16 + 2 * L] = 10
    return out.reshape(-1)
