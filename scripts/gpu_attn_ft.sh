#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_finetune.py -m gpu -q -x -p no:cacheprovider -k "attention or attn or finetune" 2>&1 | tail -3
for v in base default; do
  echo "=== variant: ${v}"
  if [ "$v" = "default" ]; then timeout 120 python scripts/attn_ab.py 2>&1 | tail -6
  else DIG_B200_LIB=libdig_b200_${v}.so timeout 120 python scripts/attn_ab.py 2>&1 | tail -6; fi
done
DIG_BENCH_VIT_BASE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_ft.json
python -c "import json; d=json.loads(open('gpurun_out/bench_ft.json').read()); print('ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'attn', d['roofline']['attention']['fwd']['avg_launch_ms'], d['roofline']['attention']['bwd']['avg_launch_ms']); print('parity', d['parity']); print('finetune', d['finetune'])"
