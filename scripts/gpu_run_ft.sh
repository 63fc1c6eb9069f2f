set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_finetune.py -m gpu -q -p no:cacheprovider > gpurun_out/r2h_ft_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error|^E " gpurun_out/r2h_ft_pytest.log | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.log 2>gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err; grep '^{' gpurun_out/r2h_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','finetune','vit_base','gpu_eager_baseline','cpu_baseline','parity') if k in d})"
python scripts/gaps.py 128 2>&1 | grep -v branch > gpurun_out/r2h_gaps.txt; head -32 gpurun_out/r2h_gaps.txt
