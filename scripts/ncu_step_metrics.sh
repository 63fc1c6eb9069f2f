#!/bin/bash
# Per-launch device time, DRAM traffic and tensor-pipe activity of ONE resident-input step (bs=128), every kernel (under gpurun).
TAG=${1:-r1}; B=${2:-128}
mkdir -p gpurun_out
ncu --profile-from-start off --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --csv --log-file gpurun_out/${TAG}_step_metrics.csv python scripts/profile_step.py $B 2 > gpurun_out/${TAG}_step_metrics.log 2>&1
tail -1 gpurun_out/${TAG}_step_metrics.log
python scripts/summarize_step_metrics.py gpurun_out/${TAG}_step_metrics.csv | head -60
