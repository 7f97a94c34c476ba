#!/bin/bash
# A/B of library variants over read lengths (single-end) and the paired 150-base leg
for so in "" $(ls fqgrep_b200/libfqgrep_b200_*.so 2>/dev/null); do
  export FQG_LIB_VARIANT=${so:+$PWD/$so}
  for rl in 150 100 75; do echo "${so:-default}: $(timeout 300 python scripts/ab_fastq.py 8000000 $rl 2>&1 | tail -1 | cut -c1-120)"; done
  echo "${so:-default}: paired $(python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['raw_fastq_resident']['roofline']['frac'],4))")"
done
