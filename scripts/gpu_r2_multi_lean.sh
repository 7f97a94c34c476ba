#!/bin/bash
# N-GPU call without the test suite: bench.py exactly as the driver launches it (torchrun, one rank per GPU), then the CPU arm
N=${1:-2}
TAG=${2:-r3n$N}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -n 4 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", d["value"], "raw", d["raw_fastq_resident"].get("value"))
for k in ("h2d_probe","e2e","e2e_write","sharded","one_process_multi_gpu"):
    print(k, json.dumps(d.get(k))[:600])
print("sharded_parity_ok", d.get("sharded_parity_ok"))
PY
timeout 300 python bench.py --impl reference --gpus $N --steps 2 --warmup 1 --ref-pairs 500000 | cut -c1-300
