#!/bin/bash
# ncu on the fused FASTQ kernel: plain run first, then the launch list, then one full capture with source correlation.
TAG=${1:-r2}
CMD="python bench.py --pairs 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --raw-pairs 4000000"
timeout 300 $CMD > gpurun_out/profq_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fastq_$TAG.csv $CMD > gpurun_out/ncu_listq_$TAG.log 2>&1
timeout 300 $CMD > gpurun_out/profq_plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_fastq_kernel -s 2 -c 1 -o gpurun_out/prof_fastq_$TAG -f $CMD > gpurun_out/ncu_fullq_$TAG.log 2>&1
tail -2 gpurun_out/ncu_fullq_$TAG.log; tail -c 600 gpurun_out/profq_plain_$TAG.log
