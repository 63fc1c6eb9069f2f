#!/bin/bash
# quick GPU round: kernel tests, row-kernel and attention micro-benchmarks (base vs current library), one bench line
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in base default; do
  echo "=== variant: ${v}"
  if [ "$v" = "default" ]; then timeout 120 python scripts/rowops_ab.py 2>&1 | tail -4
  else DIG_B200_LIB=libdig_b200_${v}.so timeout 120 python scripts/rowops_ab.py 2>&1 | tail -4; fi
done
DIG_B200_LIB=libdig_b200_base.so timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BASE ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['attention']['fwd']['avg_launch_ms'], d['clocks'])"
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_quick.json
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_quick.json').read()); print('NEW  ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['attention']['fwd']['avg_launch_ms'], d['clocks'])"
