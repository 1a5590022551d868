#!/bin/bash
# bench at N GPUs (native + baseline): tools/gpu_scale.sh N
N=${1:-4}
mkdir -p gpurun_out
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
run() {
  name=$1; shift
  timeout -k 10 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) bench.py --gpus $N "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/scale_summary.txt
  grep '^{"metric"' gpurun_out/$name.log | tail -n 1 > gpurun_out/$name.json
  python - <<PY >> gpurun_out/scale_summary.txt
import json
try:
    d = json.load(open('gpurun_out/$name.json'))
    print('  value', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), d['config']['parallelism'], 'clocks', d['clocks'])
except Exception as e:
    print('  no json', e)
PY
}
run bench_n${N}_native --steps ${STEPS:-32} --warmup 8
run bench_n${N}_baseline --steps ${STEPS:-32} --warmup 8 --impl baseline --no-graphs --no-e2e
cat gpurun_out/scale_summary.txt
grep -E "Error|error|Traceback" gpurun_out/bench_n${N}_native.log | head -5
