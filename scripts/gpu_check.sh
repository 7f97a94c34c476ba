#!/bin/bash
# GPU parity suite + one bench line; every step under its own timeout so a hang cannot eat the gpurun limit.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; tail -c 3500 gpurun_out/bench_check.json; tail -n 5 gpurun_out/bench_check.err
