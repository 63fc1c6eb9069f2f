#!/bin/bash
# 8-GPU round: gradient-averaging variants (peer-memory exchanges overlapped with the backward / one at the end / NCCL segments)
cd /root/repo; mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
B="bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e"
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], 'ms/step', round(d['ms_per_step'],3), 'crops/s', round(d['value']), 'loss', d.get('loss'), d.get('clocks'))" $1 "$2" || tail -3 $1; }
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/n8_n1.json 2>/dev/null; show gpurun_out/n8_n1.json "N=1 (GPU 0 of the same box)"
DIG_BENCH_RANK_TIMES=1 timeout 400 $R $B > gpurun_out/n8_ov3.json 2> gpurun_out/n8_ov3.err; show gpurun_out/n8_ov3.json "N=8 peer grads, 3 overlapped exchanges + 1 at the end"; grep per-rank gpurun_out/n8_ov3.err
DIG_PEER_GRAD_OVERLAP=2 timeout 400 $R $B > gpurun_out/n8_ov2.json 2>/dev/null; show gpurun_out/n8_ov2.json "N=8 peer grads, 2 overlapped + 1"
DIG_PEER_GRAD_OVERLAP=0 timeout 400 $R $B > gpurun_out/n8_ov0.json 2>/dev/null; show gpurun_out/n8_ov0.json "N=8 peer grads, one exchange at the end"
