#!/bin/bash
# One `ncu --set full` capture per kernel family (single GPU, never under a bench).  Reports land in gpurun_out/ncu_<target>.ncu-rep;
# summarise here with tools/summarize_ncu.py into profiles/.   Usage: tools/gpu_ncu.sh target[:count] ...
mkdir -p gpurun_out
for spec in "$@"; do
  t=${spec%%:*}; c=${spec#*:}; [ "$c" = "$spec" ] && c=10
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:rvn_ -c $c -f -o gpurun_out/ncu_$t \
      python tools/ncu_targets.py $t > gpurun_out/ncu_$t.log 2>&1
  echo "ncu $t exit=$? $(tail -n 1 gpurun_out/ncu_$t.log | cut -c1-200)"
done
