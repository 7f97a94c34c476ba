FQG_FASTQ_STATS=1 python scripts/ab_fastq.py 2>&1 | tail -2
python scripts/ab_fastq.py | tail -1
FQG_FASTQ_STATS=1 timeout 300 python scripts/prof_routed.py 2>&1 | tail -2
timeout 300 python scripts/prof_routed.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_stream.py -m gpu -q -x --timeout 600 2>&1 | tail -2
