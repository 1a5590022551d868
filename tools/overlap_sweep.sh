#!/bin/bash
# BASELINE.json config #4: Inception-V3 3-stage pipeline, activation send/recv overlap sweep.  For every per-replica batch size the
# step time, the exposed (non-overlapped) communication per step and the per-rank busy time are printed; needs >= 3 GPUs.
#   tools/overlap_sweep.sh [model] [SxR] [batches...]        default: inception 3x1 16 32 64 128 256
MODEL=${1:-inception}; TOPO=${2:-3x1}; shift 2 2>/dev/null
BATCHES=${@:-16 32 64 128 256}
N=$(( ${TOPO%x*} * ${TOPO#*x} ))
mkdir -p gpurun_out
export RVN_BENCH_WATCHDOG=${WD:-150}
OUT=gpurun_out/overlap_sweep_${MODEL}_${TOPO}.txt; : > $OUT
for b in $BATCHES; do
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) \
      bench.py --gpus $N --model $MODEL --topology $TOPO --batch $b --steps 32 --warmup 8 --no-e2e --rebalance 0 > gpurun_out/sweep_${MODEL}_b$b.log 2>&1
  grep '^{"metric"' gpurun_out/sweep_${MODEL}_b$b.log | tail -n 1 | python -c "
import json, sys
line = sys.stdin.read().strip()
if not line:
    print('batch $b: no result')
else:
    d = json.loads(line)
    print('batch %4d  %9.1f %s  %7.3f ms/step  exposed comm %6.3f ms (%4.1f %%)  busy %s' % ($b, d['value'], d['unit'], d['ms_per_step'],
          d['exposed_comm_ms_per_step'], 100 * d['exposed_comm_ms_per_step'] / d['ms_per_step'], d['stage_busy_ms_per_rank']))
" >> $OUT
done
cat $OUT
