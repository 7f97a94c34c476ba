#!/bin/bash
for so in "" $(ls fqgrep_b200/libfqgrep_b200_*.so 2>/dev/null); do
  echo "${so:-default}: $(FQG_LIB_VARIANT=${so:+$PWD/$so} timeout 300 python scripts/prof_routed.py 8000000 2>&1 | tail -1 | cut -c1-200)"
done
