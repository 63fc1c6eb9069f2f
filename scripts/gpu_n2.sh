#!/bin/bash
# 2-GPU round: multi-rank parity (peer exchanges incl. the gradient all-reduce vs NCCL vs one process at 2B), then bench A/B
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -p no:cacheprovider -s 2>&1 | tail -6
for pg in 1 0; do
DIG_PEER_GRADS=$pg timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/n2_pg$pg.json
python -c "import json; d=json.loads(open('gpurun_out/n2_pg$pg.json').read()); print('PEER_GRADS=$pg ms/step', d['ms_per_step'], 'crops/s', d['value'], 'loss', d.get('loss'))" || tail -5 gpurun_out/n2_pg$pg.json
done
