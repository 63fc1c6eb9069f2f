#!/bin/bash
# ncu --set full captures of the step's top kernels (one launch each, skipping the warm-up occurrences), under gpurun.
# usage: scripts/ncu_capture.sh <tag> [batch]   -> gpurun_out/<tag>_<name>.ncu-rep
TAG=${1:-cap}; B=${2:-128}
mkdir -p gpurun_out
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:$2" -s ${3:-3} -c ${4:-1} -f -o gpurun_out/${TAG}_$1 python scripts/profile_step.py $B 2 > gpurun_out/${TAG}_$1.log 2>&1
  tail -1 gpurun_out/${TAG}_$1.log
  # the .ncu-rep files are large (gpurun_out is capped at 64 MiB): keep the raw-metric and per-instruction pages as text
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_$1.source.csv.gz
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap fc1_gelu   'gemm2_bf16_tcgen05<[^,]*256, [^,]*0, [^,]*0, [^,]*6, [^,]*0, [^,]*1>' 8    # fc1 + GELU, 8-bit pre-activation codes (online branch)
cap gelu_bwd   'gemm2_bf16_tcgen05<[^,]*256, [^,]*0, [^,]*1, [^,]*7, [^,]*0, [^,]*1>' 5    # fc2 dgrad x gelu'(8-bit level)
cap wgrad      'gemm2_bf16_tcgen05<[^,]*256, [^,]*1, [^,]*1, [^,]*4, [^,]*1, [^,]*1>' 20 3
cap res_f32    'gemm2_bf16_tcgen05<[^,]*192, [^,]*0, [^,]*0, [^,]*0, [^,]*1, [^,]*1>' 30 2
cap qkv        'gemm2_bf16_tcgen05<[^,]*192, [^,]*0, [^,]*0, [^,]*0, [^,]*0, [^,]*1>' 14
cap dgrad      'gemm2_bf16_tcgen05<[^,]*128, [^,]*0, [^,]*1, [^,]*0, [^,]*0, [^,]*1>' 6
cap attn_fwd   'attn_fwd_persist_kernel' 14
cap attn_bwd   'attn_bwd_persist_kernel' 5
cap ln_bwd     'layernorm_bwd_kernel' 5
cap ln_fwd     'layernorm_fwd_vec_kernel' 30
