#!/bin/bash
# N-GPU run of bench.py exactly as the driver launches it (one rank per GPU, torchrun, NCCL only for barrier/max).
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 2500 gpurun_out/bench_n$N.json; tail -n 5 gpurun_out/bench_n$N.err
