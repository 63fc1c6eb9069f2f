set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multirank.py -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/r2e_pytest.log | tail -20
for pdl in 1 0; do
DIG_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2e_bench_pdl$pdl.log 2>gpurun_out/r2e_bench_pdl$pdl.err; grep '^{' gpurun_out/r2e_bench_pdl$pdl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=$pdl', {k:d[k] for k in ('value','ms_per_step','e2e','loss') if k in d})"
done
DIG_PDL=1 DIG_TWO_STREAMS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r2e_bench_1s.log 2>&1; grep '^{' gpurun_out/r2e_bench_1s.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1-stream PDL=1', d['ms_per_step'])"
DIG_PDL=0 DIG_TWO_STREAMS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r2e_bench_1s0.log 2>&1; grep '^{' gpurun_out/r2e_bench_1s0.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1-stream PDL=0', d['ms_per_step'])"
