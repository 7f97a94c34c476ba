#!/bin/bash
# Iteration call: smoke, the FASTQ-path GPU tests, one full bench line, then an ncu capture of the fused kernel.
TAG=${1:-r2}
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
tail -1 gpurun_out/smoke_$TAG.log
timeout 1500 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_stream.py -m gpu -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 9000 gpurun_out/bench_$TAG.json; tail -n 8 gpurun_out/bench_$TAG.err
bash scripts/gpu_r2_prof_fastq.sh $TAG
