#!/bin/bash
# attention kernels: correctness, timing (optionally per library variant), clock-stamp timelines
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or attn" -p no:cacheprovider 2>&1 | tail -3
for v in "$@"; do
  echo "=== variant: ${v}"
  if [ "$v" = "default" ]; then timeout 120 python scripts/attn_ab.py 2>&1 | tail -6
  else DIG_B200_LIB=libdig_b200_${v}.so timeout 120 python scripts/attn_ab.py 2>&1 | tail -6; fi
done
timeout 120 python scripts/attn_timeline.py > gpurun_out/attn_timeline.txt 2>&1; grep -A40 "backward kernel" gpurun_out/attn_timeline.txt
