#!/bin/bash
# N-GPU call: whole GPU suite on one of the GPUs, then bench.py exactly as the driver launches it (torchrun, one rank per GPU)
N=${1:-2}
TAG=${2:-r2n$N}
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -30 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -n 6 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", d["value"], "raw", d["raw_fastq_resident"].get("value"))
for k in ("h2d_probe","e2e","e2e_write","sharded","one_process_multi_gpu"):
    print(k, d.get(k))
print("sharded_parity_ok", d.get("sharded_parity_ok"))
PY
timeout 300 python bench.py --impl reference --gpus $N --steps 2 --warmup 1 --ref-pairs 500000 | cut -c1-400
