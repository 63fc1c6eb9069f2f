set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_finetune.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|^E |^FAILED" | head
python scripts/profile_finetune.py 256 2>&1 | grep -v "branch\|Warn\|warn" > gpurun_out/r2k_finetune_profile.txt; head -8 gpurun_out/r2k_finetune_profile.txt
DIG_TWO_STREAMS=0 bash scripts/ncu_step_metrics.sh r2 128 2>&1 | tail -3
bash scripts/ncu_capture.sh r2 128 2>&1 | tail -12
ls -la gpurun_out | grep r2_ | head -40
