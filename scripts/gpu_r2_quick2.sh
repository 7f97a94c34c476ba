FQG_FASTQ_STATS=1 python scripts/ab_fastq.py 2>&1 | tail -2
FQG_FASTQ_STATS=1 timeout 300 python scripts/prof_routed.py 2>&1 | tail -2
