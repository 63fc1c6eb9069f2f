set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 420 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2f_multirank.log 2>&1
grep -E "passed|failed|worst|Error|^\[|^E " gpurun_out/r2f_multirank.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 scripts/timeline_multi.py 128 2>/dev/null | grep -v branch > gpurun_out/r2f_timeline_n2_peer.txt; head -20 gpurun_out/r2f_timeline_n2_peer.txt
DIG_PEER=0 timeout 300 $TR --master-port 29522 scripts/timeline_multi.py 128 2>/dev/null | grep -v branch > gpurun_out/r2f_timeline_n2_nccl.txt; head -20 gpurun_out/r2f_timeline_n2_nccl.txt
