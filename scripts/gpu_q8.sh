#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "gemm" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -x -p no:cacheprovider -k "gelu" 2>&1 | tail -4
timeout 200 python scripts/gemm_ab.py GELU "fc1 fwd plain" "fc2 dgrad plain" 2>&1 | tail -9
for q in 0 1; do
DIG_GELU_Q8=$q timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('Q8=$q ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'loss', d['loss'], d['clocks'])"
done
timeout 900 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider -k "step or parity or golden" 2>&1 | tail -4
cat gpurun_out/parity_bench_config.json 2>/dev/null | head -40
