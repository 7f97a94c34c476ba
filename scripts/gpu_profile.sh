#!/bin/bash
# Round-1 profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture.
set -x
TAG=${1:-r1}
CMD="python bench.py --pairs 25000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --raw-pairs 4000000"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/prof_plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:match_soa_kernel -s 2 -c 2 -o gpurun_out/prof_soa_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_fastq_kernel -s 2 -c 2 -o gpurun_out/prof_fastq_$TAG -f $CMD > gpurun_out/ncu_fullq_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_full_$TAG.log; tail -n 3 gpurun_out/ncu_fullq_$TAG.log
