#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples out of an .ncu-rep captured with --import-source on
(ncu --page source --print-source cuda,sass --csv).  Usage: ncu_source_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main(path, top=60):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr_len = None, None
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr_len = len(r)
            continue
        if hdr_len is None or not r or not r[0].strip().isdigit():
            continue
        sh = len(r) - hdr_len
        try:
            samples = int(r[4 + sh]) if r[4 + sh] not in ("-", "") else 0
            inst = int(r[7 + sh]) if r[7 + sh] not in ("-", "") else 0
        except (ValueError, IndexError):
            continue
        src = ",".join(r[1:2 + sh]).strip()
        lines.append((cur_file, int(r[0]), inst, samples, src))
    ti = sum(x[2] for x in lines) or 1
    ts = sum(x[3] for x in lines) or 1
    print("total warp instructions %d, stall samples %d" % (ti, ts))
    print("%-22s %5s %12s %6s %8s %6s  %s" % ("file", "line", "warp-inst", "%", "samples", "%", "source"))
    for f, ln, inst, s, src in sorted(lines, key=lambda x: -x[2])[:top]:
        print("%-22s %5d %12d %6.2f %8d %6.2f  %s" % (f, ln, inst, 100.0 * inst / ti, s, 100.0 * s / ts, src[:110]))
    print("---- by stall samples")
    for f, ln, inst, s, src in sorted(lines, key=lambda x: -x[3])[:top // 2]:
        print("%-22s %5d %12d %6.2f %8d %6.2f  %s" % (f, ln, inst, 100.0 * inst / ti, s, 100.0 * s / ts, src[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
