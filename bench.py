#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 read-matching path (contract in the task statement, section 4).

Workload (BASELINE.json configs[1]): --paired, two -e regexes (GACGAGATTA, GACGTGATTA), --reverse-complement,
50 M pairs of 150 bp synthetic reads (SURVEY.md 8(d) generator, 1 % planted hits), produced directly in HBM.
A "step" is one pass of the hot path over all pairs a rank owns.

  value     Gbases/s, reads resident in HBM as SoA bases + u32 offsets (device-event timed, max over ranks)
  e2e       same metric through the C-ABI host entry point on raw FASTQ in pinned host memory: H2D copy, on-GPU
            parsing, matching, pair combine and the D2H of the selection mask are all inside the timed region
  roofline  match_soa kernel: algorithmic bytes (154.125 B/read) / measured launch time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (no Rust toolchain here, so no reference binary) on a bounded sample, all host threads

`--impl reference` times the CPU implementation (oracle port) on the same workload definition.
Multi-GPU: one process per GPU (torchrun); pairs are sharded contiguously, no collective on the data path; the
only communication is the barrier and the max-reduction of the timing (weak scaling: every rank owns a full slab).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATTERNS = ["GACGAGATTA", "GACGTGATTA"]
READ_LEN = 150
ALGO_BYTES_PER_READ = READ_LEN + 4 + 1.0 / 8.0     # bases + u32 offset + result bit
FASTQ_BYTES_PER_READ = 2 * READ_LEN + 17


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_from=0.0, t_to=1e30):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t_from <= t <= t_to] or [r for _, r in self.rows[-3:]]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for k, nme in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port (kind "port") with all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle as O
    from tests import synth
    cores = os.cpu_count() or 1
    n_pairs = args.ref_pairs
    r1 = synth.bases(n_pairs, READ_LEN, seed=1)
    r2 = synth.bases(n_pairs, READ_LEN, seed=2)
    synth.plant(r1, PATTERNS, rate=0.01, seed=99)
    synth.plant(r2, PATTERNS, rate=0.01, seed=98)
    f1, f2 = synth.fastq(r1), synth.fastq(r2)
    m = O.Matcher(O.REGEX_SET, PATTERNS, invert=False, revcomp=True)
    times = []
    cnt = None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cnt = m.grep(f1, f2, mode=O.PAIRED, threads=cores)["count"]
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    bases = 2.0 * n_pairs * READ_LEN * len(times)
    val = bases / total / 1e9
    line = {
        "impl": "reference", "metric": "Gbases/s matched", "value": val, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "--paired -e GACGAGATTA -e GACGTGATTA --reverse-complement, 150bp synthetic FASTQ", "pairs_per_step": n_pairs,
                   "selected": int(cnt)},
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": cores, "kind": "port",
                         "sample": "%d pairs of the configs[1] workload per step, FASTQ in host memory, parse+match+pair (oracle port, pthreads)" % n_pairs},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=50_000_000, help="pairs per GPU per step (configs[1]: 50M)")
    ap.add_argument("--e2e-pairs", type=int, default=8_000_000, help="pairs per step of the host-memory e2e leg")
    ap.add_argument("--ref-pairs", type=int, default=2_000_000, help="pairs per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-raw", action="store_true")
    ap.add_argument("--raw-pairs", type=int, default=20_000_000, help="pairs per step of the HBM-resident raw-FASTQ leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import fqgrep_b200 as F
    from fqgrep_b200 import _ffi

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _ffi.lib()
    ctx = F.context(local)
    sampler = ClockSampler(local)      # nvidia-smi takes a second to start: launch it before the inputs are generated
    if rank == 0:
        sampler.start()
    matcher = F.MatcherFactory.new_matcher(None, False, PATTERNS, None, F.MatcherOpts(invert_match=False, reverse_complement=True))
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    # ---- inputs, generated in HBM: slabs of <= 2^32 bytes so offsets are u32
    n_pairs = args.pairs
    slab_reads = 25_000_000
    blob = "".join(PATTERNS).encode()
    poffs = np.array([0, 10, 20], dtype=np.uint32)
    first = rank * n_pairs
    slabs = []   # (bases tensor, offsets tensor, n_reads, mate index, mask slice start)
    offs_proto = (torch.arange(slab_reads + 1, dtype=torch.int64, device=dev) * READ_LEN).to(torch.int32)
    for mate, (seed, pseed) in enumerate(((1, 99), (2, 98))):
        done = 0
        while done < n_pairs:
            n = min(slab_reads, n_pairs - done)
            tb = torch.empty(n * READ_LEN + 64, dtype=torch.uint8, device=dev)
            rc = L.fqg_synth_reads_device(ctx, tb.data_ptr(), n, first + done, READ_LEN, 0, seed, blob, poffs.ctypes.data_as(C.c_void_p), 2, 0.01, pseed, sp)
            assert rc == 0, _ffi.last_error()
            to = offs_proto[: n + 1].clone()
            slabs.append((tb, to, n, mate, done))
            done += n
    words = (n_pairs + 31) // 32 + 1
    mask1 = torch.zeros(words, dtype=torch.int32, device=dev)
    mask2 = torch.zeros(words, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    launches_before = L.fqg_kernel_launches()

    def step():
        count.zero_()
        for tb, to, n, mate, done in slabs:
            assert done % 32 == 0
            out = mask1 if mate == 0 else mask2
            orm = None if mate == 0 else mask1.data_ptr() + done // 8
            rc = L.fqg_match_bases_device(ctx, matcher._h, tb.data_ptr(), to.data_ptr(), to.data_ptr() + 4, n, out.data_ptr() + done // 8, orm,
                                          count.data_ptr() if mate == 1 else None, READ_LEN, sp)
            assert rc == 0, _ffi.last_error()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0:
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 5.0:
            time.sleep(0.05)
    t_from = time.perf_counter()
    l0 = L.fqg_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches_timed = L.fqg_kernel_launches() - l0
    t_to = time.perf_counter()
    selected = int(count.item())
    # The timed region lasts tens of milliseconds, i.e. one or two nvidia-smi samples.  Keep the identical load running
    # (untimed) for another ~0.4 s so that the clock / throttle report rests on more than that; both counts are reported.
    clocks = None
    if rank == 0:
        t_probe = time.perf_counter()
        while time.perf_counter() - t_probe < 0.4:
            step()
            torch.cuda.synchronize()
        in_region = len([1 for t, r in sampler.rows if t_from <= t <= t_to and r and r[0].isdigit()])
        clocks = sampler.stop(t_from, time.perf_counter())
        clocks["samples_in_timed_region"] = in_region
        clocks["note"] = "samples = timed region + 0.4 s of the same steps run untimed right after it"
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    bases_per_step_rank = 2.0 * n_pairs * READ_LEN
    value = bases_per_step_rank * world * args.steps / (ms / 1e3) / 1e9

    # roofline of the dominant (only) kernel in the step
    peak, peak_src = peaks()
    reads_per_launch = 2.0 * n_pairs / len(slabs)
    launch_ms = ms / launches_timed
    achieved = ALGO_BYTES_PER_READ * reads_per_launch / (launch_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")     # dram__bytes_read+write per launch from the committed `ncu --set full` capture
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if abs(tj.get("reads_per_launch", 0) - reads_per_launch) < 1:
            traffic = tj["dram_bytes_per_launch"]
    roofline = {"bound": "hbm", "kernel": "match_soa_kernel<tier0,1 slot + L2 prefetch>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)", "traffic": traffic,
                "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ, "reads_per_launch": reads_per_launch, "launch_ms": launch_ms}

    # ---- second resident layout: raw FASTQ bytes in HBM through the fused parse+match kernel (SURVEY 8(d), 2nd roofline column)
    raw = None
    if not args.no_raw:
        try:
            raw = run_raw_fastq(args, L, ctx, matcher, dev, sp, rank, world, peak)
        except Exception as ex:
            raw = {"value": None, "error": str(ex)[:200]}

    # ---- e2e through the host entry point: raw FASTQ in pinned host memory
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, L, ctx, matcher, dev, sp, rank, world)
        except Exception as ex:   # reported, never silently replaced by the device number
            e2e = {"value": None, "unit": "Gbases/s", "error": str(ex)[:200]}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    if rank == 0:
        line = {
            "metric": "Gbases/s matched", "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: --paired -e GACGAGATTA -e GACGTGATTA --reverse-complement, %d pairs x 150bp per GPU" % n_pairs,
                       "layout": "SoA bases + u32 offsets resident in HBM (%.1f GB per GPU, >> 126 MB L2, no L2 flush needed)" % (bases_per_step_rank / 1e9),
                       "pairs_per_gpu": n_pairs, "selected_pairs_rank0": selected, "parallelism": "shard%d" % world},
            "roofline": roofline, "raw_fastq_resident": raw, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_timed), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_raw_fastq(args, L, ctx, matcher, dev, sp, rank, world, peak):
    """Raw FASTQ resident in HBM -> fqg_grep_fastq_device (fused parse+match, pair combine).  Device-event timed."""
    import torch
    import torch.distributed as dist
    from fqgrep_b200 import _ffi
    n = args.raw_pairs
    nbytes = n * FASTQ_BYTES_PER_READ
    blob = "".join(PATTERNS).encode()
    poffs = np.array([0, 10, 20], dtype=np.uint32)
    bufs = []
    for seed, pseed in ((1, 99), (2, 98)):
        tb = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
        rc = L.fqg_synth_reads_device(ctx, tb.data_ptr(), n, rank * n, READ_LEN, 1, seed, blob, poffs.ctypes.data_as(C.c_void_p), 2, 0.01, pseed, sp)
        assert rc == 0, _ffi.last_error()
        bufs.append(tb)
    words = n // 32 + 2
    mask = torch.zeros(words, dtype=torch.int32, device=dev)
    res = _ffi.Result()
    stream = torch.cuda.current_stream()

    def once():
        rc = L.fqg_grep_fastq_device(ctx, matcher._h, 2, bufs[0].data_ptr(), nbytes, bufs[1].data_ptr(), nbytes, mask.data_ptr(), words, C.byref(res), sp)
        if rc != 0:
            raise RuntimeError(_ffi.last_error())

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    steps = max(3, args.steps // 2)
    kms = 0.0
    for _ in range(steps):
        once()
        kms += res.kernel_ms          # device events around memsets + both fused kernels + combine
    if world > 1:
        t = torch.tensor([kms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms = float(t.item())
    val = 2.0 * n * READ_LEN * world * steps / (kms / 1e3) / 1e9
    gbs = (FASTQ_BYTES_PER_READ + 8 + 0.125) * 2.0 * n * steps / (kms / 1e3) / 1e9
    return {"value": val, "unit": "Gbases/s", "pairs_per_step_per_gpu": n, "selected": int(res.n_selected), "ms_per_step": kms / steps,
            "api": "fqg_grep_fastq_device (FASTQ bytes resident in HBM, mode=PAIRED)",
            "roofline": {"bound": "hbm", "kernel": "match_fastq_kernel<tier0,128>", "algorithmic_bytes_per_record": FASTQ_BYTES_PER_READ + 8 + 0.125,
                         "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "note": "whole call (memsets + 2 fused kernels + combine) over device events; kernel share in profiles/"}}


def run_e2e(args, L, ctx, matcher, dev, sp, rank, world):
    import torch
    import torch.distributed as dist
    from fqgrep_b200 import _ffi
    n = args.e2e_pairs
    nbytes = n * FASTQ_BYTES_PER_READ
    blob = "".join(PATTERNS).encode()
    poffs = np.array([0, 10, 20], dtype=np.uint32)
    hosts = []
    for seed, pseed in ((1, 99), (2, 98)):
        tb = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
        rc = L.fqg_synth_reads_device(ctx, tb.data_ptr(), n, rank * n, READ_LEN, 1, seed, blob, poffs.ctypes.data_as(C.c_void_p), 2, 0.01, pseed, sp)
        assert rc == 0, _ffi.last_error()
        p = C.c_void_p()
        assert L.fqg_alloc_pinned(nbytes + 64, C.byref(p)) == 0, _ffi.last_error()
        torch.cuda.synchronize()
        host = (C.c_uint8 * (nbytes + 64)).from_address(p.value)
        ht = torch.frombuffer(host, dtype=torch.uint8)
        ht.copy_(tb.cpu())
        hosts.append((p, ht))
        del tb
    words = n // 32 + 2
    mask = np.zeros(words, dtype=np.uint32)
    res = _ffi.Result()

    def once():
        rc = L.fqg_grep_fastq_host(ctx, matcher._h, 2, hosts[0][0], nbytes, hosts[1][0], nbytes, mask.ctypes.data_as(C.c_void_p), words, C.byref(res))
        if rc != 0:
            raise RuntimeError(_ffi.last_error())

    for _ in range(max(1, args.warmup)):
        once()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    val = 2.0 * n * READ_LEN * world * args.steps / dt / 1e9
    out = {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": int(res.h2d_bytes), "d2h_bytes_per_step": int(res.d2h_bytes),
           "pairs_per_step_per_gpu": n, "selected": int(res.n_selected), "api": "fqg_grep_fastq_host (pinned host FASTQ, mode=PAIRED)",
           "h2d_ms": res.h2d_ms, "kernel_ms": res.kernel_ms,
           "pcie_roofline_gbases_s": None}
    if res.h2d_ms > 0:
        gbs = res.h2d_bytes / (res.h2d_ms / 1e3) / 1e9
        out["h2d_gb_s"] = gbs
        out["pcie_roofline_gbases_s"] = gbs / FASTQ_BYTES_PER_READ * READ_LEN
    for p, _ in hosts:
        L.fqg_free_pinned(p)
    return out


def cpu_baseline(args):
    from oracle import oracle as O
    from tests import synth
    cores = os.cpu_count() or 1
    n = args.ref_pairs
    r1 = synth.bases(n, READ_LEN, seed=1)
    r2 = synth.bases(n, READ_LEN, seed=2)
    synth.plant(r1, PATTERNS, rate=0.01, seed=99)
    synth.plant(r2, PATTERNS, rate=0.01, seed=98)
    f1, f2 = synth.fastq(r1), synth.fastq(r2)
    m = O.Matcher(O.REGEX_SET, PATTERNS, invert=False, revcomp=True)
    m.grep(f1[: 317 * 1000], f2[: 317 * 1000], mode=O.PAIRED, threads=cores)
    t0 = time.perf_counter()
    reps = 0
    while True:
        m.grep(f1, f2, mode=O.PAIRED, threads=cores)
        reps += 1
        if time.perf_counter() - t0 > 8.0 or reps >= 5:
            break
    dt = time.perf_counter() - t0
    return {"value": 2.0 * n * READ_LEN * reps / dt / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "port",
            "sample": "%d pairs x %d passes of the configs[1] workload (FASTQ in host memory, parse+match+pair), oracle port with %d pthreads; "
                      "no Rust toolchain in the image so the reference binary cannot be built" % (n, reps, cores)}


if __name__ == "__main__":
    main()
