#!/bin/bash
# A/B of attention library variants on one box: correctness first, then timing per variant, then the clock-stamp timeline.
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or attn" -p no:cacheprovider 2>&1 | tail -5
for v in "$@"; do
  echo "=== variant: ${v}"
  if [ "$v" = "default" ]; then timeout 120 python scripts/attn_ab.py 2>&1 | tail -6
  else DIG_B200_LIB=libdig_b200_${v}.so timeout 120 python scripts/attn_ab.py 2>&1 | tail -6; fi
done
timeout 120 python scripts/attn_timeline.py > gpurun_out/attn_timeline.txt 2>&1; head -45 gpurun_out/attn_timeline.txt
