#!/bin/bash
# regression + bench round on one GPU: the GPU test suite, smoke, the default bench line
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'crops/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), 'parity loss_rel', d['parity']['loss_rel'], 'vit_base', round(d['vit_base']['value']), 'finetune', round(d['finetune']['value']), d['clocks'])" || tail -5 gpurun_out/bench_final.err
