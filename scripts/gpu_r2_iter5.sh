#!/bin/bash
# smoke + FASTQ / stream / parity tests + bench line without the host-memory legs + the routed and names paths (plain runs)
TAG=${1:-r2}
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_stream.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "raw", d["raw_fastq_resident"].get("value"), d["raw_fastq_resident"].get("roofline",{}).get("frac"))
for k,v in (d.get("configs") or {}).items():
    if not isinstance(v, dict): print(k, v); continue
    so=v.get("soa",{}); rf=v.get("raw_fastq",{})
    print(k, "soa", round(so.get("gbases_s",so.get("greads_s",0)),1), round(so.get("roofline_frac",0),3), "raw", round(rf.get("gbases_s",0),1), round(rf.get("roofline_frac",0),3), "tier", v.get("tier"), "parity", v.get("parity_ok"), v.get("error"))
PY
timeout 300 python scripts/prof_routed.py 2>&1 | tail -1
timeout 300 python scripts/prof_names.py 2>&1 | tail -1
FQG_FASTQ_STATS=1 timeout 300 python scripts/prof_routed.py 2>&1 | grep "fused kernel" | tail -2
