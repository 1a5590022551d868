#!/bin/bash
# round-2 session F (N GPUs): K1 two-shot vs one-shot — numerics (dist_gpu_worker) and micro-benchmark
N=${1:-4}
mkdir -p gpurun_out
trun() { timeout -k 10 ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
for mode in 1 0; do
  echo "== numerics two_shot=$mode"; RAVNEST_K1_TWOSHOT=$mode trun tests/dist_gpu_worker.py 2>&1 | grep -E "DIST_GPU_OK|Error|error|assert" | head -5
  echo "== k1 bench two_shot=$mode"; RAVNEST_K1_TWOSHOT=$mode trun tools/k1_bench.py 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['mode'], d['optimizer'], d['params'], 'fused', d['fused_ms'], 'nccl+torch', d['nccl_plus_torch_ms'], 'algo floor', d['algorithmic_floor_ms'], 'frac', d['fraction_of_algorithmic_floor'])"
done
