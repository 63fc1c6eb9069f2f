#!/bin/bash
# final profile pass of a round (under gpurun): per-launch metrics of one step, ncu --set full of the top kernels, fine-tune kernel list
set -x
cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r2}
python scripts/profile_finetune.py 256 2>&1 | grep -v "branch\|Warn\|warn" > gpurun_out/${TAG}_finetune_profile.txt; head -5 gpurun_out/${TAG}_finetune_profile.txt
DIG_TWO_STREAMS=0 bash scripts/ncu_step_metrics.sh ${TAG} 128 2>&1 | tail -3
bash scripts/ncu_capture.sh ${TAG} 128 2>&1 | tail -12
python scripts/gaps.py 128 > gpurun_out/${TAG}_gaps.txt 2>&1; head -12 gpurun_out/${TAG}_gaps.txt
ls -la gpurun_out | grep ${TAG}_ | head -40
