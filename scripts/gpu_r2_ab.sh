#!/bin/bash
# A/B of fused-kernel switches on the raw-FASTQ-resident leg (device-event timed), no ncu.
TAG=${1:-ab}
CMD="python bench.py --pairs 1000000 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --raw-pairs 20000000"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
for V in 0 1; do
  FQG_FASTQ_EARLY_CLAIM=$V FQG_FASTQ_STATS=1 timeout 300 $CMD > gpurun_out/ab_${TAG}_early$V.json 2> gpurun_out/ab_${TAG}_early$V.err
  echo "early_claim=$V"; python - <<PY
import json
d=json.load(open("gpurun_out/ab_${TAG}_early$V.json"))
r=d["raw_fastq_resident"]; print(r.get("value"), r.get("ms_per_step"), r.get("roofline",{}).get("frac"), r.get("error"))
print(d["sharded"])
PY
  grep "fused kernel" gpurun_out/ab_${TAG}_early$V.err | sort | uniq -c | sort -rn | head -4
done
