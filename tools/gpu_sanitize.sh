#!/bin/bash
# compute-sanitizer passes over the kernel numerics tests (one GPU; slow: run a group at a time).
# Usage: tools/gpu_sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
TOOL=${1:-memcheck}; EXPR=${2:-"gemm_nt or layernorm or cross_entropy or column_sum or embedding or dropout"}
mkdir -p gpurun_out
timeout -k 10 1500 compute-sanitizer --tool $TOOL --error-exitcode 77 --log-file gpurun_out/sanitize_$TOOL.log \
    python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider -k "$EXPR" > gpurun_out/sanitize_$TOOL.pytest.log 2>&1
echo "compute-sanitizer $TOOL exit=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$TOOL.log) summaries"
grep "ERROR SUMMARY" gpurun_out/sanitize_$TOOL.log | sort | uniq -c | head
tail -n 3 gpurun_out/sanitize_$TOOL.pytest.log
