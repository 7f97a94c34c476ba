"""Shared helpers for building FASTQ fixtures the way the reference's tests do (main.rs:790-821)."""
import numpy as np


def fastq(seqs, heads=None, qual_char="X", crlf=False, plus_name=False, final_newline=True):
    """FASTQ bytes for a list of sequences.  Default headers follow the reference fixture: head "@{n}" so the
    on-disk header line is "@@{n}" (main.rs:806-813)."""
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i, s in enumerate(seqs):
        s = s.encode() if isinstance(s, str) else s
        h = (heads[i] if heads is not None else "@%d" % i)
        h = h.encode() if isinstance(h, str) else h
        out.append(b"@" + h + nl + s + nl + b"+" + (h if plus_name else b"") + nl + qual_char.encode() * len(s) + nl)
    data = b"".join(out)
    if not final_newline and data:
        data = data[:-len(nl)]
    return data


def seqs_of(fastq_bytes):
    lines = fastq_bytes.split(b"\n")
    return [lines[i].decode() for i in range(1, len(lines), 4)]


def soa(seqs):
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs) or b"", dtype=np.uint8).copy(), offs
