set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multirank.py -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/r2d_pytest.log | tail -20
python scripts/gemm_ab.py GELU > gpurun_out/r2d_gemm_ab.log 2>&1; cat gpurun_out/r2d_gemm_ab.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.log 2>gpurun_out/r2d_bench.err; grep '^{' gpurun_out/r2d_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','parity','vit_base') if k in d}); print(d['roofline']['frac'], d['roofline']['attention'])"
