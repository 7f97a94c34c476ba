#!/usr/bin/env python
"""bench.py -- benchmark of the B200 read-matching path (contract in the task statement, section 4).

Headline workload (BASELINE.json configs[1]): --paired, two -e regexes (GACGAGATTA, GACGTGATTA), --reverse-complement,
50 M pairs of 150 bp synthetic reads (SURVEY.md 8(d) generator, 1 % planted hits), produced directly in HBM.
A "step" is one pass of the hot path over all pairs a rank owns.

  value        Gbases/s, reads resident in HBM as SoA bases + u32 offsets (device-event timed, max over ranks)
  roofline     match_soa kernel: algorithmic bytes (154.125 B/read) / measured launch time vs MEASURED_PEAKS.json hbm_gbs
  raw_fastq_resident   the same pairs as raw FASTQ bytes in HBM through the fused parse+match kernel (325.125 B/record)
  raw_fastq_read_lengths   the fused kernel on single-end reads of 250 / 100 / 75 / 50 bases (4 M records each), every mask
               compared bit for bit with the SoA kernel's on the same reads
  configs      every other BASELINE.json config on one GPU (cfg1, cfg3, cfg4, cfg5a, cfg5b + the headline's own parity):
               SoA-resident rate and roofline fraction, raw-FASTQ-resident rate, tier, and parity -- the whole device mask
               of the first 2 M reads against the oracle, and the SoA kernel against the fused FASTQ kernel bit for bit
               on every record both saw
  e2e          same metric through the C-ABI host entry point on raw FASTQ in pinned host memory: the input is cut into
               batches and streamed through a ring of device slots; H2D copies, on-GPU parsing, matching, pair combine and
               the D2H of the selection mask are all inside the timed region
  e2e_write    as e2e, plus the ordered write-back of the selected records into a host buffer (device record spans ->
               host byte copies); the output is compared with the oracle's once, outside the timed region
  h2d_probe    a plain 1 GiB pinned cudaMemcpyAsync, best of 5, all ranks at once: the PCIe roofline the e2e legs are held to
  sharded      N ranks process disjoint record ranges of ONE deterministic input; rank 0 gathers masks and counts and
               compares them with the unsharded run of the same input on its own GPU (and with the oracle)
  one_process_multi_gpu   rank 0 alone drives all N GPUs through fqg_grep_fastq_host_multi (the product's multi-GPU path)
  cpu_baseline the oracle port (no Rust toolchain here, so no reference binary) on a bounded sample, all host threads

`--impl reference` times the CPU implementation (oracle port) on the same workload definition.
Multi-GPU: one process per GPU (torchrun); pairs are sharded contiguously, no collective on the data path; the
only communication is the barrier, the max-reduction of the timing and the gather of the sharded leg's masks.
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
FASTQ_ALGO_BYTES = FASTQ_BYTES_PER_READ + 8 + 1.0 / 8.0
SLAB_READS = 25_000_000


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


def host_pairs(n_pairs):
    """configs[1] input on the host: R1/R2 FASTQ of n_pairs pairs, planted like the device generator plants them."""
    from tests import synth
    r1 = synth.bases(n_pairs, READ_LEN, seed=1)
    r2 = synth.bases(n_pairs, READ_LEN, seed=2)
    synth.plant(r1, PATTERNS, rate=0.01, seed=99)
    synth.plant(r2, PATTERNS, rate=0.01, seed=98)
    return synth.fastq(r1), synth.fastq(r2)


def run_reference(args, rank, world):
    """CPU arm: the oracle port (kind "port") with all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    n_pairs = args.ref_pairs
    f1, f2 = host_pairs(n_pairs)
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
                         "sample": "%d pairs of the configs[1] workload per step, FASTQ in host memory, parse+match+pair (oracle port, pthreads); "
                                   "vs oracle port, not the fqgrep binary (no Rust toolchain in the image)" % n_pairs},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


class Env:
    """What every leg needs: the library, one context per rank, torch device and stream."""

    def __init__(self, args, rank, world, local):
        import torch
        import fqgrep_b200 as F
        from fqgrep_b200 import _ffi
        self.torch, self.F, self.ffi = torch, F, _ffi
        self.args, self.rank, self.world, self.local = args, rank, world, local
        torch.cuda.set_device(local)
        self.dev = torch.device("cuda", local)
        self.L = _ffi.lib()
        self.ctx = F.context(local)
        self.stream = torch.cuda.current_stream()
        self.sp = C.c_void_p(self.stream.cuda_stream)
        self.peak, self.peak_src = peaks()

    def synth(self, n, first, seed, plant, rate, pseed, fastq, read_len=READ_LEN):
        """n synthetic reads starting at read index `first`, generated in HBM (SoA bases or 317-byte FASTQ records)."""
        blob = "".join(plant).encode()
        offs = np.zeros(len(plant) + 1, dtype=np.uint32)
        if plant:
            offs[1:] = np.cumsum([len(p) for p in plant])
        rec = (2 * read_len + 17) if fastq else read_len
        t = self.torch.empty(n * rec + 64, dtype=self.torch.uint8, device=self.dev)
        rc = self.L.fqg_synth_reads_device(self.ctx, t.data_ptr(), n, first, read_len, int(fastq), seed, blob, offs.ctypes.data_as(C.c_void_p), len(plant),
                                           rate, pseed, self.sp)
        assert rc == 0, self.ffi.last_error()
        return t

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world > 1:
            import torch.distributed as dist
            t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x


def unpack(mask_words, n):
    return np.unpackbits(np.ascontiguousarray(mask_words).view(np.uint8), bitorder="little")[:n].astype(bool)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=50_000_000, help="pairs per GPU per step (configs[1]: 50M)")
    ap.add_argument("--e2e-pairs", type=int, default=8_000_000, help="pairs per step of the host-memory e2e leg")
    ap.add_argument("--write-pairs", type=int, default=4_000_000, help="pairs per step of the e2e_write leg")
    ap.add_argument("--ref-pairs", type=int, default=2_000_000, help="pairs per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--parity-reads", type=int, default=2_000_000, help="reads of every config whose whole mask is checked against the oracle")
    ap.add_argument("--shard-reads", type=int, default=2_000_000, help="records per rank of the sharded-parity leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-raw", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config legs (they run on one GPU only in any case)")
    ap.add_argument("--configs-scale", type=float, default=1.0, help="scale the read counts of the config legs (quick runs)")
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

    env = Env(args, rank, world, local)
    dev, L, ctx, sp, stream = env.dev, env.L, env.ctx, env.sp, env.stream
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local)      # nvidia-smi takes a second to start: launch it before the inputs are generated
    if rank == 0:
        sampler.start()
    matcher = F.MatcherFactory.new_matcher(None, False, PATTERNS, None, F.MatcherOpts(invert_match=False, reverse_complement=True))

    # ---- inputs, generated in HBM: slabs of <= 2^32 bytes so offsets are u32
    n_pairs = args.pairs
    first = rank * n_pairs
    slabs = []   # (bases tensor, offsets tensor, n_reads, mate index, mask slice start)
    offs_proto = (torch.arange(SLAB_READS + 1, dtype=torch.int64, device=dev) * READ_LEN).to(torch.int32)
    for mate, (seed, pseed) in enumerate(((1, 99), (2, 98))):
        done = 0
        while done < n_pairs:
            n = min(SLAB_READS, n_pairs - done)
            tb = env.synth(n, first + done, seed, PATTERNS, 0.01, pseed, fastq=False)
            slabs.append((tb, offs_proto[: n + 1], n, mate, done))
            done += n
    words = (n_pairs + 31) // 32 + 1
    mask1 = torch.zeros(words, dtype=torch.int32, device=dev)
    mask2 = torch.zeros(words, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step():
        count.zero_()
        for tb, to, n, mate, done in slabs:
            assert done % 32 == 0
            out = mask1 if mate == 0 else mask2
            orm = None if mate == 0 else mask1.data_ptr() + done // 8
            rc = L.fqg_match_bases_device(ctx, matcher._h, tb.data_ptr(), to.data_ptr(), to.data_ptr() + 4, n, out.data_ptr() + done // 8, orm,
                                          count.data_ptr() if mate == 1 else None, READ_LEN, sp)
            assert rc == 0, _ffi.last_error()

    for _ in range(args.warmup):
        step()
    env.barrier()
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
    env.barrier()
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
    ms = env.max_over_ranks(ms)
    bases_per_step_rank = 2.0 * n_pairs * READ_LEN
    value = bases_per_step_rank * world * args.steps / (ms / 1e3) / 1e9

    # roofline of the dominant (only) kernel in the step
    peak, peak_src = env.peak, env.peak_src
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

    # parity of the headline itself: pairs [0, parity) of rank 0's slab against the oracle (whole mask)
    headline_parity = None
    if rank == 0 and not args.no_configs:
        headline_parity = guarded(lambda: headline_parity_check(env, slabs, mask2, n_pairs))
    del slabs, mask1, mask2
    torch.cuda.empty_cache()

    # ---- second resident layout: raw FASTQ bytes in HBM through the fused parse+match kernel (SURVEY 8(d), 2nd roofline column)
    raw = None if args.no_raw else guarded(lambda: run_raw_fastq(env, matcher))
    read_lengths = guarded(lambda: run_read_lengths(env, matcher)) if (world == 1 and not args.no_raw and not args.no_configs) else None
    torch.cuda.empty_cache()

    # ---- every other BASELINE config, one GPU
    configs = None
    if world == 1 and not args.no_configs:
        configs = guarded(lambda: run_configs(env))
        if isinstance(configs, dict) and headline_parity is not None:
            configs["cfg2_headline"] = headline_parity
        torch.cuda.empty_cache()

    # ---- the PCIe roofline, measured by itself; then e2e through the host entry points
    probe = guarded(lambda: h2d_probe(env))
    e2e = e2e_write = None
    if not args.no_e2e:
        e2e = guarded(lambda: run_e2e(env, matcher, probe), {"value": None, "unit": "Gbases/s"})
        e2e_write = guarded(lambda: run_e2e_write(env, matcher), {"value": None, "unit": "Gbases/s"})

    # ---- one deterministic input, sharded over the ranks; and the one-process multi-GPU product path
    sharded = guarded(lambda: run_sharded(env, matcher))
    multi = guarded(lambda: run_one_process_multi(env, matcher)) if world > 1 else None

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
            "roofline": roofline, "raw_fastq_resident": raw, "raw_fastq_read_lengths": read_lengths, "configs": configs, "h2d_probe": probe, "e2e": e2e, "e2e_write": e2e_write,
            "sharded": sharded, "sharded_parity_ok": bool(isinstance(sharded, dict) and sharded.get("parity_ok")),
            "one_process_multi_gpu": multi, "cpu_baseline": cpu, "gpu_launches": int(launches_timed), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def guarded(fn, base=None):
    """A leg that fails is reported as an error in its own key, never silently replaced by another number."""
    try:
        return fn()
    except Exception as ex:     # noqa: BLE001
        out = dict(base or {})
        out["error"] = "%s: %s" % (type(ex).__name__, str(ex)[:300])
        return out


def headline_parity_check(env, slabs, mask2, n_pairs):
    from oracle import oracle as O
    n = min(env.args.parity_reads, n_pairs, SLAB_READS)
    host = []
    for mate in (0, 1):
        tb = [s for s in slabs if s[3] == mate and s[4] == 0][0][0]
        host.append(tb[: n * READ_LEN].cpu().numpy())
    offs = np.arange(n + 1, dtype=np.uint64) * READ_LEN
    o = O.Matcher(O.REGEX_SET, PATTERNS, False, True)
    threads = os.cpu_count() or 1
    _, f1 = o.match_soa(host[0], offs, threads=threads)
    _, f2 = o.match_soa(host[1], offs, threads=threads)
    got = unpack(mask2[: n // 32 + 1].cpu().numpy(), n)
    ok = bool((got == ((f1 != 0) | (f2 != 0))).all())
    return {"workload": "configs[1] (the headline)", "parity": {"pairs_checked_against_oracle": n, "oracle_ok": ok}, "parity_ok": ok}


# ---- the other BASELINE configs ----------------------------------------------------------------------------------------
def run_configs(env):
    F = env.F
    rng = np.random.default_rng(7)
    p10k = ["".join("ACGT"[i] for i in rng.integers(0, 4, 20)) for _ in range(10_000)]
    p1k = p10k[:1000]
    s = env.args.configs_scale
    out = {}
    out["cfg1"] = config_leg(env, "configs[0]: GACGAGATTA --count, single-end (RegexMatcher via the positional pattern)", F.REGEX, ["GACGAGATTA"], False, False,
                             int(10_000_000 * s), ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"])
    out["cfg1_fixed"] = config_leg(env, "configs[0] with -F (FixedStringMatcher)", F.FIXED, ["GACGAGATTA"], False, False, int(10_000_000 * s),
                                   ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"], parity=False)
    out["cfg3"] = config_leg(env, "configs[2]: -F -f 10,000 random 20-mers (seed 7), 1% planted", F.FIXED_SET, p10k, False, False, int(100_000_000 * s), p10k[:200])
    out["cfg4"] = config_leg(env, "configs[3]: -e 'AC[GT]{3}TT(A|G)CA' -v", F.REGEX_SET, ["AC[GT]{3}TT(A|G)CA"], True, False, int(200_000_000 * s),
                             ["GACGAGATTA", "GACGTGATTA", "ACGTGTTACA"])
    out["cfg5a"] = config_leg(env, "configs[4]a: -F -f 1,000 fixed patterns; the 1 B pairs are sharded over 8 GPUs, this is one GPU's resident slab run",
                              F.FIXED_SET, p1k, False, False, int(100_000_000 * s), p1k[:50])
    out["cfg5b"] = names_leg(env, int(50_000_000 * s))
    return out


def time_on_stream(env, fn, min_ms=40.0, max_reps=20):
    """Device-event time of fn() per call: one warm-up call, then repeated until min_ms have been timed."""
    torch = env.torch
    fn()
    torch.cuda.synchronize()
    reps, total = 0, 0.0
    while total < min_ms and reps < max_reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(env.stream)
        fn()
        e1.record(env.stream)
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
        reps += 1
    return total / reps, reps


def config_leg(env, workload, kind, pats, inv, rc, n_reads, plant, parity=True):
    """One single-end config: SoA-resident rate over n_reads (slabs of 25 M), raw-FASTQ-resident rate over the first 8 M,
    whole-mask parity of the first 2 M against the oracle, SoA kernel == fused FASTQ kernel on the 8 M."""
    from oracle import oracle as O
    torch, F, L, ffi = env.torch, env.F, env.L, env.ffi
    t0 = time.perf_counter()
    m = F.Matcher(kind, pats, F.MatcherOpts(inv, rc))
    compile_s = time.perf_counter() - t0
    info = m.dfa_info()
    slabs, done = [], 0
    while done < n_reads:
        n = min(SLAB_READS, n_reads - done)
        slabs.append((env.synth(n, done, 1, plant, 0.01, 99, fastq=False), n, done))
        done += n
    offs = (torch.arange(min(SLAB_READS, n_reads) + 1, dtype=torch.int64, device=env.dev) * READ_LEN).to(torch.int32)
    mask = torch.zeros((n_reads + 31) // 32 + 1, dtype=torch.int32, device=env.dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=env.dev)

    def soa_pass():
        cnt.zero_()
        for tb, n, first in slabs:
            rc_ = L.fqg_match_bases_device(env.ctx, m._h, tb.data_ptr(), offs.data_ptr(), offs.data_ptr() + 4, n, mask.data_ptr() + first // 8, None,
                                           cnt.data_ptr(), READ_LEN, env.sp)
            assert rc_ == 0, ffi.last_error()
    ms, reps = time_on_stream(env, soa_pass)
    selected = int(cnt.item())
    out = {"workload": workload, "reads": n_reads, "tier": int(info.tier), "dfa_states": int(info.n_states), "compile_s": round(compile_s, 3), "selected": selected,
           "soa": {"gbases_s": n_reads * READ_LEN / (ms / 1e3) / 1e9, "ms_per_pass": ms, "passes_timed": reps,
                   "roofline_frac": ALGO_BYTES_PER_READ * n_reads / (ms / 1e3) / 1e9 / env.peak, "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ}}
    # raw FASTQ of the same reads [0, n_fq) through the fused kernel
    n_fq = min(8_000_000, n_reads)
    fq = env.synth(n_fq, 0, 1, plant, 0.01, 99, fastq=True)
    nbytes = n_fq * FASTQ_BYTES_PER_READ
    fwords = n_fq // 32 + 2
    fmask = torch.zeros(fwords, dtype=torch.int32, device=env.dev)
    res = ffi.Result()
    kms = []

    def fq_pass():
        rc_ = L.fqg_grep_fastq_device(env.ctx, m._h, 0, fq.data_ptr(), nbytes, None, 0, fmask.data_ptr(), fwords, C.byref(res), env.sp)
        assert rc_ == 0, ffi.last_error()
        kms.append(res.kernel_ms)
    for _ in range(4):
        fq_pass()
    fms = float(np.mean(kms[1:]))
    out["raw_fastq"] = {"gbases_s": n_fq * READ_LEN / (fms / 1e3) / 1e9, "ms_per_pass": fms, "records": n_fq,
                        "roofline_frac": FASTQ_ALGO_BYTES * n_fq / (fms / 1e3) / 1e9 / env.peak, "algorithmic_bytes_per_record": FASTQ_ALGO_BYTES,
                        "note": "whole call: memsets + fused kernel, device events"}
    same = bool(torch.equal(fmask[: n_fq // 32], mask[: n_fq // 32]) and int(res.n_selected) == int(unpack(mask[: n_fq // 32 + 1].cpu().numpy(), n_fq).sum()))
    par = {"soa_eq_fastq_records": n_fq, "soa_eq_fastq_ok": same}
    ok = same
    if parity:
        n_par = min(env.args.parity_reads, n_reads)
        host = slabs[0][0][: n_par * READ_LEN].cpu().numpy()
        ocnt, flags = O.Matcher(kind, pats, inv, rc).match_soa(host, np.arange(n_par + 1, dtype=np.uint64) * READ_LEN, threads=os.cpu_count() or 1)
        got = unpack(mask[: n_par // 32 + 1].cpu().numpy(), n_par)
        par["reads_checked_against_oracle"] = n_par
        par["oracle_ok"] = bool((got == (flags != 0)).all())
        par["oracle_selected"] = int(ocnt)
        ok = ok and par["oracle_ok"]
    out["parity"] = par
    out["parity_ok"] = ok
    return out


def names_leg(env, n_reads, n_names=1_000_000):
    """configs[4]b: -N with 1 M ids (90 % present).  SoA = the header lines' ids back to back + u32 offsets; raw FASTQ = fused kernel."""
    from oracle import oracle as O
    torch, F, L, ffi = env.torch, env.F, env.L, env.ffi
    dev = env.dev
    idl = 11
    idx = torch.arange(n_reads, device=dev, dtype=torch.int64)
    hb = torch.empty((n_reads, idl), dtype=torch.uint8, device=dev)
    hb[:, 0] = ord("r")
    for k in range(10):
        hb[:, 10 - k] = ((idx // (10 ** k)) % 10 + 48).to(torch.uint8)
    del idx
    offs = (torch.arange(n_reads + 1, device=dev, dtype=torch.int64) * idl).to(torch.int32)
    rng = np.random.default_rng(5)
    n_fq = min(8_000_000, n_reads)
    # the names are drawn from the records both layouts hold, so that both legs select
    chosen = rng.choice(n_fq, min(int(n_names * 0.9), n_fq // 2), replace=False)
    names = ["r%010d" % i for i in chosen] + ["r%010d" % (n_reads + i) for i in range(n_names - chosen.size)]
    t0 = time.perf_counter()
    m = F.Matcher(F.NAMES, names)
    compile_s = time.perf_counter() - t0
    mask = torch.zeros(n_reads // 32 + 2, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)

    def soa_pass():
        cnt.zero_()
        rc_ = L.fqg_match_bases_device(env.ctx, m._h, hb.data_ptr(), offs.data_ptr(), offs.data_ptr() + 4, n_reads, mask.data_ptr(), None, cnt.data_ptr(), idl, env.sp)
        assert rc_ == 0, ffi.last_error()
    ms, reps = time_on_stream(env, soa_pass)
    selected = int(cnt.item())
    algo = idl + 4 + 8 + 0.125
    out = {"workload": "configs[4]b: -N read-names file of 1,000,000 ids (%d present)" % chosen.size, "reads": n_reads, "names": len(names), "compile_s": round(compile_s, 3),
           "selected": selected,
           "soa": {"greads_s": n_reads / (ms / 1e3) / 1e9, "ms_per_pass": ms, "passes_timed": reps, "roofline_frac": algo * n_reads / (ms / 1e3) / 1e9 / env.peak,
                   "algorithmic_bytes_per_read": algo}}
    fq = env.synth(n_fq, 0, 1, [], 0.0, 0, fastq=True)
    nbytes = n_fq * FASTQ_BYTES_PER_READ
    fwords = n_fq // 32 + 2
    fmask = torch.zeros(fwords, dtype=torch.int32, device=dev)
    res = ffi.Result()
    kms = []
    for _ in range(4):
        rc_ = L.fqg_grep_fastq_device(env.ctx, m._h, 0, fq.data_ptr(), nbytes, None, 0, fmask.data_ptr(), fwords, C.byref(res), env.sp)
        assert rc_ == 0, ffi.last_error()
        kms.append(res.kernel_ms)
    fms = float(np.mean(kms[1:]))
    out["raw_fastq"] = {"greads_s": n_fq / (fms / 1e3) / 1e9, "gbases_s": n_fq * READ_LEN / (fms / 1e3) / 1e9, "ms_per_pass": fms, "records": n_fq,
                        "roofline_frac": FASTQ_ALGO_BYTES * n_fq / (fms / 1e3) / 1e9 / env.peak, "algorithmic_bytes_per_record": FASTQ_ALGO_BYTES}
    want = np.zeros(n_fq, dtype=bool)
    want[chosen] = True
    got_fq = unpack(fmask.cpu().numpy(), n_fq)
    got_soa = unpack(mask[: n_fq // 32 + 1].cpu().numpy(), n_fq)
    n_par = min(env.args.parity_reads, n_fq)
    e = O.Matcher(O.NAMES, names, False, False).grep(fq[: n_par * FASTQ_BYTES_PER_READ].cpu().numpy(), threads=os.cpu_count() or 1)
    par = {"soa_eq_fastq_records": n_fq, "soa_eq_fastq_ok": bool((got_fq == got_soa).all()), "closed_form_ok": bool((got_fq == want).all() and selected == chosen.size),
           "reads_checked_against_oracle": n_par, "oracle_ok": bool(((e["flags"] != 0) == got_fq[:n_par]).all())}
    out["parity"] = par
    out["parity_ok"] = bool(par["soa_eq_fastq_ok"] and par["closed_form_ok"] and par["oracle_ok"])
    return out


# ---- raw FASTQ resident in HBM (headline workload) -------------------------------------------------------------------------
def run_raw_fastq(env, matcher):
    """Raw FASTQ resident in HBM -> fqg_grep_fastq_device (fused parse+match, pair combine).  Device-event timed."""
    torch, L, ffi = env.torch, env.L, env.ffi
    n = env.args.raw_pairs
    nbytes = n * FASTQ_BYTES_PER_READ
    bufs = [env.synth(n, env.rank * n, seed, PATTERNS, 0.01, pseed, fastq=True) for seed, pseed in ((1, 99), (2, 98))]
    words = n // 32 + 2
    mask = torch.zeros(words, dtype=torch.int32, device=env.dev)
    res = ffi.Result()

    def once():
        rc = L.fqg_grep_fastq_device(env.ctx, matcher._h, 2, bufs[0].data_ptr(), nbytes, bufs[1].data_ptr(), nbytes, mask.data_ptr(), words, C.byref(res), env.sp)
        if rc != 0:
            raise RuntimeError(ffi.last_error())

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    steps = max(3, env.args.steps // 2)
    kms = 0.0
    for _ in range(steps):
        once()
        kms += res.kernel_ms          # device events around memsets + both fused kernels + combine
    kms = env.max_over_ranks(kms)
    val = 2.0 * n * READ_LEN * env.world * steps / (kms / 1e3) / 1e9
    gbs = FASTQ_ALGO_BYTES * 2.0 * n * steps / (kms / 1e3) / 1e9
    return {"value": val, "unit": "Gbases/s", "pairs_per_step_per_gpu": n, "selected": int(res.n_selected), "ms_per_step": kms / steps,
            "api": "fqg_grep_fastq_device (FASTQ bytes resident in HBM, mode=PAIRED)",
            "roofline": {"bound": "hbm", "kernel": "match_fastq_kernel<tier0> (warp-autonomous tiles)", "algorithmic_bytes_per_record": FASTQ_ALGO_BYTES,
                         "achieved": gbs, "peak": env.peak, "unit": "GB/s", "frac": gbs / env.peak,
                         "note": "whole call (memsets + 2 fused kernels + combine) over device events; kernel share in profiles/"}}


def run_read_lengths(env, matcher, lengths=(250, 100, 75, 50), n=4_000_000):
    """The fused kernel away from the benchmark's 150 bases: single-end, HBM-resident, each length with the SoA kernel's mask
    of the same reads as the check (bit for bit).  Below ~130 bases a tile holds more than one record per lane."""
    torch, L, ffi = env.torch, env.L, env.ffi
    out = {}
    for rl in lengths:
        rec = 2 * rl + 17
        fq = env.synth(n, 0, 1, PATTERNS, 0.01, 99, fastq=True, read_len=rl)
        words = n // 32 + 2
        fmask = torch.zeros(words, dtype=torch.int32, device=env.dev)
        res = ffi.Result()
        kms = []
        for _ in range(4):
            rc = L.fqg_grep_fastq_device(env.ctx, matcher._h, 0, fq.data_ptr(), n * rec, None, 0, fmask.data_ptr(), words, C.byref(res), env.sp)
            if rc != 0:
                raise RuntimeError(ffi.last_error())
            kms.append(res.kernel_ms)
        del fq
        tb = env.synth(n, 0, 1, PATTERNS, 0.01, 99, fastq=False, read_len=rl)
        offs = (torch.arange(n + 1, dtype=torch.int64, device=env.dev) * rl).to(torch.int32)
        smask = torch.zeros(words, dtype=torch.int32, device=env.dev)
        rc = L.fqg_match_bases_device(env.ctx, matcher._h, tb.data_ptr(), offs.data_ptr(), offs.data_ptr() + 4, n, smask.data_ptr(), None, None, rl, env.sp)
        if rc != 0:
            raise RuntimeError(ffi.last_error())
        torch.cuda.synchronize()
        fms = float(np.mean(kms[1:]))
        out["%d" % rl] = {"records": n, "record_bytes": rec, "ms_per_pass": fms, "gbases_s": n * rl / (fms / 1e3) / 1e9,
                          "roofline_frac": (rec + 8.125) * n / (fms / 1e3) / 1e9 / env.peak, "selected": int(res.n_selected),
                          "fastq_eq_soa_ok": bool(torch.equal(fmask[: n // 32], smask[: n // 32]))}
        del tb, offs
    return out


# ---- host-memory legs ---------------------------------------------------------------------------------------------------
def pinned_pair(env, n, first):
    """R1/R2 of n pairs as raw FASTQ in pinned host memory (generated on the device, copied down once)."""
    torch, L, ffi = env.torch, env.L, env.ffi
    nbytes = n * FASTQ_BYTES_PER_READ
    hosts = []
    for seed, pseed in ((1, 99), (2, 98)):
        tb = env.synth(n, first, seed, PATTERNS, 0.01, pseed, fastq=True)
        p = C.c_void_p()
        assert L.fqg_alloc_pinned(nbytes + 64, C.byref(p)) == 0, ffi.last_error()
        torch.cuda.synchronize()
        host = (C.c_uint8 * (nbytes + 64)).from_address(p.value)
        ht = torch.frombuffer(host, dtype=torch.uint8)
        ht.copy_(tb.cpu())
        hosts.append((p, ht))
        del tb
    return hosts, nbytes


def h2d_probe(env):
    """The PCIe roofline, measured on its own: 1 GiB from pinned host memory, best of 5, every rank copying at the same time."""
    torch = env.torch
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=env.dev)
    best = None
    for _ in range(6):
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(env.stream)
        d.copy_(h, non_blocking=True)
        e1.record(env.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    worst_rank_ms = env.max_over_ranks(best)
    gbs = n / (worst_rank_ms / 1e3) / 1e9
    return {"bytes": n, "best_ms_slowest_rank": worst_rank_ms, "h2d_gb_s_per_gpu": gbs, "pcie_roofline_gbases_s_per_gpu": gbs / FASTQ_BYTES_PER_READ * READ_LEN,
            "pcie_roofline_gbases_s_all_gpus": gbs / FASTQ_BYTES_PER_READ * READ_LEN * env.world,
            "how": "torch pinned -> device copy_ (cudaMemcpyAsync), CUDA events, best of 5 after a warm-up, all %d ranks at once" % env.world}


def run_e2e(env, matcher, probe):
    torch, L, ffi = env.torch, env.L, env.ffi
    n = env.args.e2e_pairs
    hosts, nbytes = pinned_pair(env, n, env.rank * n)
    words = n // 32 + 2
    mask = np.zeros(words, dtype=np.uint32)
    res = ffi.Result()

    def once():
        rc = L.fqg_grep_fastq_host(env.ctx, matcher._h, 2, hosts[0][0], nbytes, hosts[1][0], nbytes, mask.ctypes.data_as(C.c_void_p), words, C.byref(res))
        if rc != 0:
            raise RuntimeError(ffi.last_error())

    for _ in range(max(1, env.args.warmup)):
        once()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(env.args.steps):
        once()
    torch.cuda.synchronize()
    dt = env.max_over_ranks(time.perf_counter() - t0)
    val = 2.0 * n * READ_LEN * env.world * env.args.steps / dt / 1e9
    out = {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": int(res.h2d_bytes), "d2h_bytes_per_step": int(res.d2h_bytes),
           "pairs_per_step_per_gpu": n, "selected": int(res.n_selected),
           "api": "fqg_grep_fastq_host (pinned host FASTQ, mode=PAIRED) -> id-guided batch cutter -> fqg_stream ring of 3 x 64 MiB slots per input stream",
           "device_memory_bound": "ring: 3 slots x 2 streams x 64 MiB of FASTQ + per-record scratch, whatever the input size",
           "timing": "host wall clock around %d calls (each call synchronises), max over ranks" % env.args.steps}
    if isinstance(probe, dict) and probe.get("pcie_roofline_gbases_s_all_gpus"):
        out["pcie_roofline_gbases_s"] = probe["pcie_roofline_gbases_s_all_gpus"]
        out["frac_of_pcie_roofline"] = val / probe["pcie_roofline_gbases_s_all_gpus"]
    for p, _ in hosts:
        L.fqg_free_pinned(p)
    return out


def run_e2e_write(env, matcher):
    """e2e with output: pinned FASTQ in -> the selected records, in order, in a host buffer.  Checked against the oracle's bytes once."""
    torch, L, ffi = env.torch, env.L, env.ffi
    n = env.args.write_pairs
    hosts, nbytes = pinned_pair(env, n, env.rank * n)
    cap = 2 * nbytes + 64
    po = C.c_void_p()
    assert L.fqg_alloc_pinned(cap, C.byref(po)) == 0, ffi.last_error()
    out_len = C.c_size_t(0)
    res = ffi.Result()

    def once():
        rc = L.fqg_grep_fastq_host_write(env.ctx, matcher._h, 2, hosts[0][0], nbytes, hosts[1][0], nbytes, po, cap, C.byref(out_len), None, 0, None, C.byref(res))
        if rc != 0:
            raise RuntimeError(ffi.last_error())

    once()
    env.barrier()
    steps = max(3, env.args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    torch.cuda.synchronize()
    dt = env.max_over_ranks(time.perf_counter() - t0)
    val = 2.0 * n * READ_LEN * env.world * steps / dt / 1e9
    out = {"value": val, "unit": "Gbases/s", "pairs_per_step_per_gpu": n, "selected_pairs": int(res.n_selected), "output_bytes": int(out_len.value),
           "h2d_bytes_per_step": int(res.h2d_bytes), "d2h_bytes_per_step": int(res.d2h_bytes),
           "api": "fqg_grep_fastq_host_write (pinned host FASTQ in, selected records out; device record spans, host byte copies)"}
    if env.rank == 0:
        from oracle import oracle as O
        n_chk = min(n, 1_000_000)      # the first n_chk pairs: the oracle writes them in the same order, so its output is a prefix of ours
        f1 = np.frombuffer((C.c_uint8 * (n_chk * FASTQ_BYTES_PER_READ)).from_address(hosts[0][0].value), dtype=np.uint8)
        f2 = np.frombuffer((C.c_uint8 * (n_chk * FASTQ_BYTES_PER_READ)).from_address(hosts[1][0].value), dtype=np.uint8)
        e = O.Matcher(O.REGEX_SET, PATTERNS, False, True).grep(f1, f2, mode=O.PAIRED, want_output=True, threads=os.cpu_count() or 1)
        got = bytes((C.c_uint8 * len(e["out"])).from_address(po.value)) if len(e["out"]) <= out_len.value else b""
        out["output_checked_pairs"] = n_chk
        out["output_parity_ok"] = bool(got == e["out"] and len(e["out"]) > 0)
    for p, _ in hosts:
        L.fqg_free_pinned(p)
    L.fqg_free_pinned(po)
    return out


# ---- one input, N ranks ------------------------------------------------------------------------------------------------
def run_sharded(env, matcher):
    """ONE deterministic single-end input of world x S records (configs[1]'s patterns, --reverse-complement): rank r parses and
    matches records [r S, (r+1) S) as raw FASTQ on its own GPU; rank 0 gathers the masks and counts (the only exchange, after
    the timed region) and compares them with the unsharded run of the whole input on its own GPU, and with the oracle."""
    torch, L, ffi = env.torch, env.L, env.ffi
    S, world, rank = env.args.shard_reads, env.world, env.rank
    S -= S % 32

    def grep(first, n):
        fq = env.synth(n, first, 1, PATTERNS, 0.01, 99, fastq=True)
        words = n // 32 + 2
        mask = torch.zeros(words, dtype=torch.int32, device=env.dev)
        res = ffi.Result()
        rc = L.fqg_grep_fastq_device(env.ctx, matcher._h, 0, fq.data_ptr(), n * FASTQ_BYTES_PER_READ, None, 0, mask.data_ptr(), words, C.byref(res), env.sp)
        if rc != 0:
            raise RuntimeError(ffi.last_error())
        return mask[: n // 32].clone(), int(res.n_selected), fq

    mine, cnt, fq = grep(rank * S, S)
    if world > 1:
        import torch.distributed as dist
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        counts = [torch.zeros(1, dtype=torch.int64, device=env.dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([cnt], dtype=torch.int64, device=env.dev))
        gathered = torch.cat(parts)
        total = int(sum(int(c.item()) for c in counts))
    else:
        gathered, total = mine, cnt
    out = None
    if rank == 0:
        del fq
        whole, whole_cnt, wfq = grep(0, world * S)
        same = bool(torch.equal(gathered, whole) and total == whole_cnt)
        from oracle import oracle as O
        n_chk = min(env.args.parity_reads, S)
        e = O.Matcher(O.REGEX_SET, PATTERNS, False, True).grep(wfq[: n_chk * FASTQ_BYTES_PER_READ].cpu().numpy(), threads=os.cpu_count() or 1)
        oracle_ok = bool(((e["flags"] != 0) == unpack(whole[: n_chk // 32 + 1].cpu().numpy(), n_chk)).all())
        h = int(torch.sum(gathered.to(torch.int64) * (torch.arange(gathered.numel(), device=env.dev, dtype=torch.int64) % 1000003 + 1)).item())
        out = {"records_per_rank": S, "ranks": world, "selected_total": total, "selected_unsharded": whole_cnt, "mask_checksum": h,
               "sharded_equals_unsharded": same, "oracle_records_checked": n_chk, "oracle_ok": oracle_ok, "parity_ok": bool(same and oracle_ok)}
    env.barrier()
    return out


def run_one_process_multi(env, matcher):
    """The product's multi-GPU path: ONE process (rank 0) deals the batches of one pinned input to a context per GPU
    (fqg_grep_fastq_host_multi) while the other ranks wait; checked against the same call on one GPU."""
    torch, L, ffi, F = env.torch, env.L, env.ffi, env.F
    out = None
    env.barrier()
    if env.rank == 0:
        n_dev = min(env.world, torch.cuda.device_count())
        n = 2_000_000 * n_dev
        hosts, nbytes = pinned_pair(env, n, 0)
        words = n // 32 + 2
        res1, resn = ffi.Result(), ffi.Result()
        m1, mn = np.zeros(words, dtype=np.uint32), np.zeros(words, dtype=np.uint32)
        rc = L.fqg_grep_fastq_host(env.ctx, matcher._h, 2, hosts[0][0], nbytes, hosts[1][0], nbytes, m1.ctypes.data_as(C.c_void_p), words, C.byref(res1))
        if rc != 0:
            raise RuntimeError(ffi.last_error())
        ctxs = (C.c_void_p * n_dev)(*[F.context(d) for d in range(n_dev)])
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            rc = L.fqg_grep_fastq_host_multi(ctxs, n_dev, matcher._h, 2, hosts[0][0], nbytes, hosts[1][0], nbytes, mn.ctypes.data_as(C.c_void_p), words,
                                             None, 0, None, None, 0, None, C.byref(resn))
            if rc != 0:
                raise RuntimeError(ffi.last_error())
            times.append(time.perf_counter() - t0)
        torch.cuda.set_device(env.local)
        ok = bool(resn.n_selected == res1.n_selected and resn.n_units == res1.n_units and (m1 == mn).all())
        out = {"gpus": n_dev, "pairs": n, "value": 2.0 * n * READ_LEN / min(times[1:]) / 1e9, "unit": "Gbases/s", "selected": int(resn.n_selected),
               "equals_one_gpu": ok, "api": "fqg_grep_fastq_host_multi, one host thread, batches dealt round-robin, host-side gather in batch order"}
        for p, _ in hosts:
            L.fqg_free_pinned(p)
    env.barrier()
    return out


def cpu_baseline(args):
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    n = args.ref_pairs
    f1, f2 = host_pairs(n)
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
                      "vs oracle port, not the fqgrep binary: no Rust toolchain in the image so the reference cannot be built" % (n, reps, cores)}


if __name__ == "__main__":
    main()
