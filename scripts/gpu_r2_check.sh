#!/bin/bash
# Round-2 GPU check: smoke first (a hang or a wrong answer there stops the call early), then the whole GPU suite, then
# one bench line.  Every step under its own timeout so a hang cannot eat the gpurun limit.
TAG=${1:-r2}
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
tail -2 gpurun_out/smoke_$TAG.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/bench_$TAG.json; tail -n 5 gpurun_out/bench_$TAG.err
