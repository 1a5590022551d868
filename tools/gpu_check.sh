#!/bin/bash
# Run the GPU numerics groups in isolated processes (a hung kernel cannot take the other groups down);
# logs land in gpurun_out/.  Usage: tools/gpu_check.sh [group ...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv 2>&1
GROUPS_DEFAULT="gemm_nt gemm_dgrad gemm_wgrad gemm_epilogue linear_autograd layernorm cross_entropy fused_optimizer bert_small smoke_entry"
for g in ${@:-$GROUPS_DEFAULT}; do
  echo "=== $g" | tee -a gpurun_out/check_summary.txt
  timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -k "$g" -p no:cacheprovider > gpurun_out/check_$g.log 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/check_$g.log)" | tee -a gpurun_out/check_summary.txt
done
