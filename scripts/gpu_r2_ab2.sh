#!/bin/bash
for so in "" $(ls fqgrep_b200/libfqgrep_b200_*.so 2>/dev/null); do
  FQG_LIB_VARIANT=${so:+$PWD/$so} timeout 300 python scripts/ab_fastq.py 2>&1 | tail -1
done
