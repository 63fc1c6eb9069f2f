set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_finetune.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2h_ft_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error|assert|^E " gpurun_out/r2h_ft_pytest.log | head -40
