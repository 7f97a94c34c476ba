#!/bin/bash
# smoke + FASTQ-path tests + raw-FASTQ leg + ncu of the fused kernel
TAG=${1:-r2}
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_stream.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
CMD="python bench.py --pairs 1000000 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --raw-pairs 20000000"
FQG_FASTQ_STATS=1 timeout 300 $CMD > gpurun_out/raw_$TAG.json 2> gpurun_out/raw_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/raw_$TAG.json"))
r=d["raw_fastq_resident"]; print("raw paired:", r.get("value"), r.get("ms_per_step"), r.get("roofline",{}).get("frac"), r.get("error"))
PY
grep "fused kernel" gpurun_out/raw_$TAG.err | sort | uniq -c | sort -rn | head -3
bash scripts/gpu_r2_prof_fastq.sh $TAG
