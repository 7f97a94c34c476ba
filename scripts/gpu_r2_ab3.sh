#!/bin/bash
# A/B of library variants: single-end call (ab_fastq.py) and the paired raw-FASTQ leg of bench.py
for so in "" $(ls fqgrep_b200/libfqgrep_b200_*.so 2>/dev/null); do
  export FQG_LIB_VARIANT=${so:+$PWD/$so}
  echo "${so:-default}: $(timeout 300 python scripts/ab_fastq.py 2>&1 | tail -1 | cut -c1-120) | paired $(python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['raw_fastq_resident']['roofline']['frac'],4))")"
done
