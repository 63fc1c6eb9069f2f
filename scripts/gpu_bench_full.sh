#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
t1=$(date +%s); echo "bench.py default run: $((t1-t0)) s"
tail -1 gpurun_out/bench_full.log | cut -c1-400
t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
t1=$(date +%s); echo "bench.py --impl reference: $((t1-t0)) s"
tail -1 gpurun_out/bench_ref.log | cut -c1-600
