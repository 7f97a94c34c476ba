#!/bin/bash
TAG=${1:-r1}
CMD="python bench.py --pairs 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --raw-pairs 4000000"
timeout 300 $CMD > gpurun_out/profq_plain_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_fastq_kernel -s 2 -c 1 -o gpurun_out/prof_fastq_$TAG -f $CMD > gpurun_out/ncu_fullq_$TAG.log 2>&1
tail -2 gpurun_out/ncu_fullq_$TAG.log
