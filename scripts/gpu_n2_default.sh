#!/bin/bash
# the bench exactly as the driver launches it at N=2 (extras, e2e with the prefetching engine)
cd /root/repo; mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_full.log 2> gpurun_out/n2_full.err
echo "rc=$? $(( $(date +%s) - t0 )) s"
python -c "
import json; d=json.loads(open('gpurun_out/n2_full.log').read().strip().splitlines()[-1])
print('N=2 ms/step', round(d['ms_per_step'],3), 'crops/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'loss', d['loss'], d['clocks'])
print('torch_ddp', d.get('torch_ddp'))" || tail -20 gpurun_out/n2_full.err
