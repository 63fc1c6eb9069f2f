#!/bin/bash
# regression + bench round on one GPU
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 200 python scripts/gemm_ab.py "proj fwd" 2>&1 | tail -2
for pf in 0 1; do
DIG_PREFETCH=$pf DIG_BENCH_VIT_BASE=0 DIG_BENCH_FINETUNE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/bench_pf$pf.json
python -c "import json; d=json.loads(open('gpurun_out/bench_pf$pf.json').read()); print('PREFETCH=$pf ms/step', d['ms_per_step'], 'e2e', d['e2e'])"
done
