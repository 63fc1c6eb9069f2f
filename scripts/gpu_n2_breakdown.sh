#!/bin/bash
# where does the N=2 overhead come from?  same box: N=1 on each GPU, N=2 default, N=2 nearly independent replicas
cd /root/repo; mkdir -p gpurun_out
B="bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], 'ms/step', round(d['ms_per_step'],3), 'crops/s', round(d['value']), d.get('clocks'))" $1 "$2"; }
CUDA_VISIBLE_DEVICES=0 timeout 300 python $B > gpurun_out/bd_n1.json 2>/dev/null; show gpurun_out/bd_n1.json "N=1 GPU0"
CUDA_VISIBLE_DEVICES=1 timeout 300 python $B > gpurun_out/bd_n1b.json 2>/dev/null; show gpurun_out/bd_n1b.json "N=1 GPU1"
(CUDA_VISIBLE_DEVICES=0 timeout 300 python $B > gpurun_out/bd_c0.json 2>/dev/null &  CUDA_VISIBLE_DEVICES=1 timeout 300 python $B > gpurun_out/bd_c1.json 2>/dev/null; wait)
show gpurun_out/bd_c0.json "two concurrent independent N=1 runs: GPU0"; show gpurun_out/bd_c1.json "two concurrent independent N=1 runs: GPU1"
DIG_BENCH_RANK_TIMES=1 timeout 300 $R $B --gpus 2 > gpurun_out/bd_n2.json 2> gpurun_out/bd_n2.err; show gpurun_out/bd_n2.json "N=2 default"; grep "per-rank" gpurun_out/bd_n2.err
DIG_BENCH_RANK_TIMES=1 DIG_BENCH_NO_SYNCBN=1 DIG_BENCH_DDP=none timeout 300 $R $B --gpus 2 > gpurun_out/bd_n2_rep.json 2> gpurun_out/bd_n2_rep.err; show gpurun_out/bd_n2_rep.json "N=2 replicas (no SyncBN, no grad averaging)"; grep "per-rank" gpurun_out/bd_n2_rep.err
