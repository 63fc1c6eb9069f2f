#!/bin/bash
# 8-GPU round: bench A/B of the gradient averaging path (peer-memory kernel vs NCCL segments)
cd /root/repo; mkdir -p gpurun_out
for pg in 1 0; do
DIG_PEER_GRADS=$pg timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/n8_pg$pg.json
python -c "import json; d=json.loads(open('gpurun_out/n8_pg$pg.json').read()); print('PEER_GRADS=$pg ms/step', d['ms_per_step'], 'crops/s', d['value'], 'loss', d.get('loss'), d.get('clocks'))" || tail -5 gpurun_out/n8_pg$pg.json
done
