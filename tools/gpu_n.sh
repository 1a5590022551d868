#!/bin/bash
# N-GPU validation + measurements: dist kernels tests (N>=2), K1 micro-bench, BERT and ResNet-50 bench at the default topology.
# Usage: tools/gpu_n.sh N [bert] [resnet] [k1] [tests] [baseline]
N=$1; shift
mkdir -p gpurun_out
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
S=gpurun_out/n${N}_summary.txt; : > $S
tr() { timeout -k 10 ${TMO:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
bench() {
  name=$1; shift
  tr bench.py --gpus $N "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" >> $S
  grep '^{"metric"' gpurun_out/$name.log | tail -n 1 > gpurun_out/$name.json
  python - <<PY >> $S
import json
try:
    d = json.load(open('gpurun_out/$name.json'))
    print('  value', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), d['config']['parallelism'],
          'exposed_comm_ms', d.get('exposed_comm_ms_per_step'), 'clocks', d['clocks'])
except Exception as e:
    print('  no json', e)
PY
  grep -E "Error|Traceback" gpurun_out/$name.log | head -3 >> $S
}
for what in "$@"; do
  case $what in
    tests) tr tests/dist_gpu_worker.py > gpurun_out/n${N}_worker.log 2>&1; echo "dist_gpu_worker exit=$? $(grep DIST_GPU_OK gpurun_out/n${N}_worker.log)" >> $S;;
    k1) tr tools/k1_bench.py > gpurun_out/n${N}_k1.log 2>&1; echo "k1 exit=$?" >> $S; grep '^{' gpurun_out/n${N}_k1.log >> $S;;
    bert) bench bench_n${N}_native --steps ${STEPS:-32} --warmup 8;;
    resnet) bench bench_n${N}_resnet50_native --model resnet50 --steps ${STEPS:-24} --warmup 8;;
    baseline) bench bench_n${N}_baseline --steps ${STEPS:-32} --warmup 8 --impl baseline --no-graphs --no-e2e;;
    resnet_baseline) bench bench_n${N}_resnet50_baseline --model resnet50 --steps ${STEPS:-24} --warmup 8 --impl baseline --no-graphs --no-e2e;;
  esac
done
cat $S
