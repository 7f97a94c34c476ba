#!/bin/bash
# A/B of kernel-parameter variants of the library inside one GPU call (variants built beforehand with build.build_variant)
TAG=${1:-ab}
CMD="python bench.py --pairs 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --raw-pairs 8000000"
for so in "" $(ls fqgrep_b200/libfqgrep_b200_*.so); do
  for rep in 1 2; do
    FQG_LIB_VARIANT=${so:+$PWD/$so} timeout 300 $CMD 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['raw_fastq_resident']
print('${so:-default}', 'raw', round(r.get('value',0),1), round(r.get('roofline',{}).get('frac',0),4), 'selected', r.get('selected'), r.get('error'))"
  done
done
