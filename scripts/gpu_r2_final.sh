#!/bin/bash
# Final evidence of the round: smoke, the whole GPU suite, the full bench line, then (each command after its plain run has
# exited 0) the launch list of the same bench command and one full capture of the headline SoA kernel.
TAG=${1:-r3z}
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
tail -1 gpurun_out/smoke_$TAG.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_$TAG.err
CMD="python bench.py --pairs 25000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --raw-pairs 4000000"
timeout 300 $CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 300 $CMD > gpurun_out/prof_plain2_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_soa_kernel -s 2 -c 2 -o gpurun_out/prof_soa_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_full_$TAG.log
