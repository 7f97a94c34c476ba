python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; tail -c 3500 gpurun_out/bench_r1_c.json; tail -5 gpurun_out/bench_r1_c.err
