python scripts/ab_fastq.py | tail -1
python scripts/ab_fastq.py | tail -1
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('paired raw frac', d['raw_fastq_resident']['roofline']['frac'])"
python scripts/prof_routed.py | tail -1
python scripts/prof_names.py | tail -1
timeout 800 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_stream.py -m gpu -q -x 2>&1 | tail -2
