#!/bin/bash
# smoke + names / FASTQ tests + config legs + launch list of the routed path + ncu of the names kernel
TAG=${1:-r2}
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_stream.py tests/test_gpu_parity.py -m gpu -q --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
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
CMDR="python scripts/prof_routed.py"
timeout 300 $CMDR > gpurun_out/profr_plain_$TAG.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_routed_$TAG.csv $CMDR > gpurun_out/ncu_routed_$TAG.log 2>&1
tail -2 gpurun_out/profr_plain_$TAG.log
CMDN="python scripts/prof_names.py"
timeout 300 $CMDN > gpurun_out/profn_plain_$TAG.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fastq_kernel -s 2 -c 1 -o gpurun_out/prof_names_$TAG -f $CMDN > gpurun_out/ncu_names_$TAG.log 2>&1
tail -2 gpurun_out/profn_plain_$TAG.log
