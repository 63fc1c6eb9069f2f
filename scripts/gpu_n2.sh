#!/bin/bash
# 2-GPU round: multi-rank parity (peer exchanges incl. the gradient averaging vs NCCL vs one process at 2B), peer all-reduce micro-benchmark, bench A/B
cd /root/repo; mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -p no:cacheprovider -s 2>&1 | tail -4
timeout 300 $R scripts/peer_grad_ab.py 2>&1 | grep -E "peer|NCCL all" | tail -12
B="bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e"
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], 'ms/step', round(d['ms_per_step'],3), 'crops/s', round(d['value']), 'loss', d.get('loss'), d.get('clocks'))" $1 "$2"; }
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/n2_n1.json 2>/dev/null; show gpurun_out/n2_n1.json "N=1"
timeout 400 $R $B > gpurun_out/n2_overlap.json 2>/dev/null; show gpurun_out/n2_overlap.json "N=2 peer grads, 2 exchanges overlapped + 1 at the end"
DIG_PEER_GRAD_OVERLAP=0 timeout 400 $R $B > gpurun_out/n2_end.json 2>/dev/null; show gpurun_out/n2_end.json "N=2 peer grads, one exchange at the end"
DIG_PEER_GRADS=0 timeout 400 $R $B > gpurun_out/n2_nccl.json 2>/dev/null; show gpurun_out/n2_nccl.json "N=2 NCCL segments"
timeout 400 $R $B > gpurun_out/n2_overlap2.json 2>/dev/null; show gpurun_out/n2_overlap2.json "N=2 peer grads overlapped (repeat)"
