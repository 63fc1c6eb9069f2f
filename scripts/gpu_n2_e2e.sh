#!/bin/bash
# N=2: end-to-end (train_one_epoch) vs resident step time for the prefetch / overlapped-exchange switches
cd /root/repo; mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
B="bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
show() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], 'resident', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))" $1 "$2" || tail -3 $1; }
timeout 200 $R $B > gpurun_out/e2e_a.json 2>/dev/null; show gpurun_out/e2e_a.json "default           "
DIG_PREFETCH=0 timeout 200 $R $B > gpurun_out/e2e_b.json 2>/dev/null; show gpurun_out/e2e_b.json "PREFETCH=0        "
DIG_PEER_GRAD_OVERLAP=0 timeout 200 $R $B > gpurun_out/e2e_c.json 2>/dev/null; show gpurun_out/e2e_c.json "GRAD_OVERLAP=0    "
