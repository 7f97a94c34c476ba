#!/usr/bin/env python
"""Measures every BASELINE.json config on one B200 (HBM-resident SoA reads unless noted) and checks a sample of each
against the oracle.  Writes gpurun_out/configs_<tag>.json.  Parity-test cases per section 4 -- not bench.py lines."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fqgrep_b200 as F                      # noqa: E402
from fqgrep_b200 import _ffi                 # noqa: E402
from oracle import oracle as O               # noqa: E402
from tests import synth                      # noqa: E402

L = _ffi.lib()
dev = torch.device("cuda:0")
ctx = F.context(0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
RL = 150


def gen(n, seed, plant, rate, pseed, fastq=0):
    blob = "".join(plant).encode()
    offs = np.zeros(len(plant) + 1, dtype=np.uint32)
    offs[1:] = np.cumsum([len(p) for p in plant])
    rec = (2 * RL + 17) if fastq else RL
    t = torch.empty(n * rec + 64, dtype=torch.uint8, device=dev)
    rc = L.fqg_synth_reads_device(ctx, t.data_ptr(), n, 0, RL, fastq, seed, blob, offs.ctypes.data_as(C.c_void_p), len(plant), rate, pseed, sp)
    assert rc == 0, _ffi.last_error()
    return t


def time_soa(m, tb, n, steps=5):
    offs = (torch.arange(n + 1, dtype=torch.int64, device=dev) * RL).to(torch.int32)
    mask = torch.zeros((n + 31) // 32 + 1, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)

    def once():
        cnt.zero_()
        rc = L.fqg_match_bases_device(ctx, m._h, tb.data_ptr(), offs.data_ptr(), offs.data_ptr() + 4, n, mask.data_ptr(), None, cnt.data_ptr(), RL, sp)
        assert rc == 0, _ffi.last_error()
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, int(cnt.item()), mask


def check_sample(kind, pats, inv, rc, tb, mask, n_check=200_000):
    host = tb[: n_check * RL].cpu().numpy()
    offs = np.arange(n_check + 1, dtype=np.uint64) * RL
    ocnt, flags = O.Matcher(kind, pats, inv, rc).match_soa(host, offs, threads=os.cpu_count())
    got = np.unpackbits(mask[: n_check // 32 + 1].cpu().numpy().view(np.uint8), bitorder="little")[:n_check].astype(bool)
    return bool((got == (flags != 0)).all()), int(ocnt)


def run_soa(name, kind, pats, inv, rc, n, plant, out):
    if ONLY and ONLY not in name:
        return
    tb = gen(n, 1, plant, 0.01, 99)
    t0 = time.perf_counter()
    m = F.Matcher(kind, pats, F.MatcherOpts(inv, rc))
    build_s = time.perf_counter() - t0
    info = m.dfa_info()
    ms, cnt, mask = time_soa(m, tb, n)
    ok, ocnt = check_sample(kind, pats, inv, rc, tb, mask)
    gb = n * RL / (ms / 1e3) / 1e9
    out[name] = {"reads": n, "ms": ms, "gbases_s": gb, "roofline_frac": (RL + 4.125) * n / (ms / 1e3) / 1e9 / PEAK, "selected": cnt,
                 "tier": info.tier, "dfa_states": info.n_states, "dfa_classes": info.n_classes, "table_bytes": info.table_bytes,
                 "compile_s": build_s, "parity_sample_ok": ok}
    print(name, json.dumps(out[name]), flush=True)
    del tb


def run_names(out, n=20_000_000, n_names=1_000_000):
    tb = gen(n, 1, [], 0.0, 0, fastq=1)
    rng = np.random.default_rng(5)
    chosen = rng.choice(n, int(n_names * 0.9), replace=False)
    names = ["r%010d" % i for i in chosen] + ["r%010d" % (n + i) for i in range(n_names - chosen.size)]
    t0 = time.perf_counter()
    m = F.Matcher(F.NAMES, names)
    build_s = time.perf_counter() - t0
    nbytes = n * (2 * RL + 17)
    words = n // 32 + 2
    mask = torch.zeros(words, dtype=torch.int32, device=dev)
    res = _ffi.Result()

    def once():
        rc = L.fqg_grep_fastq_device(ctx, m._h, 0, tb.data_ptr(), nbytes, None, 0, mask.data_ptr(), words, C.byref(res), sp)
        assert rc == 0, _ffi.last_error()
    for _ in range(2):
        once()
    kms = 0.0
    for _ in range(5):
        once()
        kms += res.kernel_ms
    kms /= 5
    out["cfg5b_names_1M_ids_fastq"] = {"reads": n, "ms": kms, "greads_s": n / (kms / 1e3) / 1e9, "gbases_s": n * RL / (kms / 1e3) / 1e9,
                                       "raw_fastq_roofline_frac": (2 * RL + 17 + 8.125) * n / (kms / 1e3) / 1e9 / PEAK,
                                       "selected": int(res.n_selected), "expected_selected": int(chosen.size), "compile_s": build_s,
                                       "parity_ok": int(res.n_selected) == int(chosen.size)}
    print("cfg5b", json.dumps(out["cfg5b_names_1M_ids_fastq"]), flush=True)


ONLY = sys.argv[2] if len(sys.argv) > 2 else ""


def run_names_soa(out, n=50_000_000, n_names=1_000_000):
    """cfg5b on SoA headers: ids r%010d back to back (11 B each) + u32 offsets; SURVEY 8(d): header + 4 + 8 + 1/8 B per read."""
    idl = 11
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    hb = torch.empty((n, idl), dtype=torch.uint8, device=dev)
    hb[:, 0] = ord("r")
    for k in range(10):
        hb[:, 10 - k] = ((idx // (10 ** k)) % 10 + 48).to(torch.uint8)
    del idx
    offs = (torch.arange(n + 1, device=dev, dtype=torch.int64) * idl).to(torch.int32)
    rng = np.random.default_rng(5)
    chosen = rng.choice(n, int(n_names * 0.9), replace=False)
    names = ["r%010d" % i for i in chosen] + ["r%010d" % (n + i) for i in range(n_names - chosen.size)]
    m = F.Matcher(F.NAMES, names)
    words = n // 32 + 2
    mask = torch.zeros(words, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)

    def once():
        rc = L.fqg_match_bases_device(ctx, m._h, hb.data_ptr(), offs.data_ptr(), offs.data_ptr() + 4, n, mask.data_ptr(), None, cnt.data_ptr(), idl, sp)
        assert rc == 0, _ffi.last_error()
    for _ in range(3):
        once()
    cnt.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    sel = int(cnt.item()) // 5
    out["cfg5b_names_1M_ids_soa"] = {"reads": n, "ms": ms, "greads_s": n / (ms / 1e3) / 1e9, "roofline_frac": (idl + 12.125) * n / (ms / 1e3) / 1e9 / PEAK,
                                     "selected": sel, "expected_selected": int(chosen.size), "parity_ok": sel == int(chosen.size)}
    print("cfg5b_soa", json.dumps(out["cfg5b_names_1M_ids_soa"]), flush=True)


def run_index_only(out, n=20_000_000):
    """K0 alone: parse + validate + index, no matching: the parse floor of the fused kernel."""
    tb = gen(n, 1, [], 0.0, 0, fastq=1)
    nbytes = n * (2 * RL + 17)
    d_start = torch.zeros(n + 8, dtype=torch.int64, device=dev)
    d_len = torch.zeros(n + 8, dtype=torch.int32, device=dev)
    nrec = C.c_uint64(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(5):
        if it == 2:
            torch.cuda.synchronize()
            e0.record()
        rc = L.fqg_index_fastq_device(ctx, tb.data_ptr(), nbytes, d_start.data_ptr(), d_len.data_ptr(), n + 8, C.byref(nrec), sp)
        assert rc == 0, _ffi.last_error()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out["k0_index_only_fastq"] = {"reads": n, "ms_incl_host_sync": ms, "greads_s": n / (ms / 1e3) / 1e9,
                                  "raw_fastq_roofline_frac": (2 * RL + 17 + 12) * n / (ms / 1e3) / 1e9 / PEAK, "records": int(nrec.value)}
    print("k0_index_only", json.dumps(out["k0_index_only_fastq"]), flush=True)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    out = {"peak_gbs": PEAK}
    rng = np.random.default_rng(7)
    p10k = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(10_000)]
    p1k = p10k[:1000]
    lit = ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"]
    run_soa("cfg1_regex_literal_10M", F.REGEX, ["GACGAGATTA"], False, False, 10_000_000, lit, out)
    run_soa("cfg1_fixed_10M", F.FIXED, ["GACGAGATTA"], False, False, 10_000_000, lit, out)
    run_soa("cfg2_two_regex_rc_25M", F.REGEX_SET, ["GACGAGATTA", "GACGTGATTA"], False, True, 25_000_000, lit, out)
    run_soa("cfg4_class_regex_invert_25M", F.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"], True, False, 25_000_000, lit, out)
    p200 = ["".join("ACGT"[i] for i in rng.integers(0, 4, 12)) for _ in range(200)]
    run_soa("extra_200x12mers_rc_25M", F.FIXED_SET, p200, False, True, 25_000_000, p200[:20], out)
    mixed = p10k[:300] + ["AC.G[AT]C"]
    run_soa("extra_mixed_300lit_1regex_25M", F.REGEX_SET, mixed, False, True, 25_000_000, p10k[:20], out)
    r60 = ["".join("ACGT"[i] for i in rng.integers(0, 4, 8)) + "[AC]" + "".join("ACGT"[i] for i in rng.integers(0, 4, 6)) for _ in range(60)]
    run_soa("extra_regexset_60_25M", F.REGEX_SET, r60, False, False, 25_000_000, p10k[:20], out)
    r400 = []
    for _ in range(1200):
        t = ["ACGT"[i] for i in rng.integers(0, 4, 10)]
        t[int(rng.integers(2, 8))] = "[AC]"
        r400.append("".join(t))
    r300 = []
    for p in p10k[1000:1300]:      # 300 regexes "20-mer with one two-letter class": finite languages -> q-gram filter
        k = int(rng.integers(3, 17))
        r300.append(p[:k] + "[" + p[k] + "N]" + p[k + 1:])
    run_soa("extra_regexset_300x20mers_one_class_25M", F.REGEX_SET, r300, False, True, 25_000_000, p10k[1000:1020], out)
    anch = ["^" + p[:12] for p in p10k[2000:4000]]      # 2000 barcodes anchored at the read start: dead state, early stop
    run_soa("extra_anchored_2000x12_25M", F.REGEX_SET, anch, False, False, 25_000_000, [], out)
    run_soa("extra_regexset_400_25M", F.REGEX_SET, r400[:400], False, False, 25_000_000, p10k[:20], out)      # class-indexed smem table (tier 1)
    run_soa("extra_regexset_1200_25M", F.REGEX_SET, r400, False, False, 25_000_000, p10k[:20], out)           # table through L2 (tier 2)
    run_soa("extra_1000x10mers_25M", F.FIXED_SET, [p[:10] for p in p1k], False, False, 25_000_000, [p[:10] for p in p1k[:20]], out)
    run_soa("cfg5a_1000_fixed_25M", F.FIXED_SET, p1k, False, False, 25_000_000, p1k[:50], out)
    run_soa("cfg3_10k_20mers_25M", F.FIXED_SET, p10k, False, False, 25_000_000, p10k[:200], out)
    run_soa("extra_cfg3_10k_20mers_revcomp_25M", F.FIXED_SET, p10k, False, True, 25_000_000, p10k[:200], out)
    if not ONLY or ONLY in "cfg5b_names":
        run_names(out)
        run_names_soa(out)
    if not ONLY or ONLY in "k0_index_only":
        run_index_only(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_%s.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    main()
