#!/bin/bash
# Round-2 profiling call: every command runs plain first (exit 0), then under ncu.
#   1. fused FASTQ kernel (tier 0): launch list + full capture with source correlation
#   2. routed tier-3 path (index -> compact -> SoA): launch list
#   3. -N over raw FASTQ and over SoA headers: full captures
TAG=${1:-r2}
CMD="python bench.py --pairs 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --raw-pairs 4000000"
timeout 300 $CMD > gpurun_out/profq_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fastq_$TAG.csv $CMD > gpurun_out/ncu_listq_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_fastq_kernel -s 2 -c 1 -o gpurun_out/prof_fastq_$TAG -f $CMD > gpurun_out/ncu_fullq_$TAG.log 2>&1
tail -2 gpurun_out/ncu_fullq_$TAG.log; tail -c 300 gpurun_out/profq_plain_$TAG.log
CMDR="python scripts/prof_routed.py"
timeout 300 $CMDR > gpurun_out/profr_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_routed_$TAG.csv $CMDR > gpurun_out/ncu_routed_$TAG.log 2>&1
tail -2 gpurun_out/profr_plain_$TAG.log
CMDN="python scripts/prof_names.py"
timeout 300 $CMDN > gpurun_out/profn_plain_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fastq_kernel -s 2 -c 1 -o gpurun_out/prof_names_$TAG -f $CMDN > gpurun_out/ncu_names_$TAG.log 2>&1
tail -2 gpurun_out/profn_plain_$TAG.log
CMDS="python scripts/prof_names_soa.py"
timeout 300 $CMDS > gpurun_out/profs_plain_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_names_soa_kernel -s 2 -c 1 -o gpurun_out/prof_names_soa_$TAG -f $CMDS > gpurun_out/ncu_names_soa_$TAG.log 2>&1
tail -2 gpurun_out/profs_plain_$TAG.log
