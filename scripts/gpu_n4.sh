#!/bin/bash
# 4-GPU round: the bench exactly as the driver launches it (extras on), plus the all-reduce micro-benchmark's correctness line at W=4
cd /root/repo; mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611"
timeout 200 $R scripts/peer_grad_ab.py 2>&1 | grep -E "max \|peer|148 blocks|148 small|NCCL all_reduce \(one" | tail -5
t0=$(date +%s)
timeout 900 $R bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_full.log 2> gpurun_out/n4_full.err
t1=$(date +%s); echo "bench N=4 default: $((t1-t0)) s, rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/n4_full.log').read().strip().splitlines()[-1])
print('N=4 ms/step', round(d['ms_per_step'],3), 'crops/s', round(d['value']), 'e2e', d['e2e']['value'], 'loss', d['loss'], d['clocks'])
print('torch_ddp', d.get('torch_ddp')); print('config', d['config'])" || tail -20 gpurun_out/n4_full.err
timeout 300 $R bench.py --impl reference --gpus 4 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
