set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n8_peer.log 2> gpurun_out/r2g_bench_n8_peer.err
grep '^{' gpurun_out/r2g_bench_n8_peer.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','torch_ddp','vit_base') if k in d})"
DIG_PEER=0 timeout 300 $TR --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2g_bench_n8_nccl.log 2> gpurun_out/r2g_bench_n8_nccl.err
grep '^{' gpurun_out/r2g_bench_n8_nccl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e') if k in d})"
timeout 200 $TR --master-port 29533 scripts/timeline_multi.py 128 2>/dev/null | grep -v branch > gpurun_out/r2g_timeline_n8_peer.txt; head -16 gpurun_out/r2g_timeline_n8_peer.txt
DIG_PEER=0 timeout 200 $TR --master-port 29534 scripts/timeline_multi.py 128 2>/dev/null | grep -v branch > gpurun_out/r2g_timeline_n8_nccl.txt; head -16 gpurun_out/r2g_timeline_n8_nccl.txt
