#!/bin/bash
# multi-GPU validation: fused DP kernel numerics, pipeline over NCCL p2p, bench at N GPUs.  Usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
echo "== dist_gpu_worker" > gpurun_out/multi_summary.txt
timeout -k 10 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 tests/dist_gpu_worker.py > gpurun_out/multi_worker.log 2>&1
echo "exit=$? $(grep DIST_GPU_OK gpurun_out/multi_worker.log)" >> gpurun_out/multi_summary.txt
echo "== bcast worker (K2)" >> gpurun_out/multi_summary.txt
timeout -k 10 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29873 tests/dist_gpu_bcast_worker.py > gpurun_out/multi_bcast.log 2>&1
echo "exit=$? $(grep BCAST_GPU_OK gpurun_out/multi_bcast.log)" >> gpurun_out/multi_summary.txt
run() {
  name=$1; shift
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) bench.py --gpus $N "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/multi_summary.txt
  grep '^{"metric"' gpurun_out/$name.log | tail -n 1 > gpurun_out/$name.json
  python - <<PY >> gpurun_out/multi_summary.txt
import json
try:
    d = json.load(open('gpurun_out/$name.json'))
    print('  value', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), d['config']['parallelism'], 'clocks', d['clocks'])
except Exception as e:
    print('  no json', e)
PY
}
run bench_n${N}_pp --steps ${STEPS:-24} --warmup 6
run bench_n${N}_dp --steps ${STEPS:-24} --warmup 6 --topology 1x$N --update-frequency 2
run bench_n${N}_pp_baseline --steps ${STEPS:-24} --warmup 6 --impl baseline --no-e2e
cat gpurun_out/multi_summary.txt
tail -n 25 gpurun_out/multi_worker.log
