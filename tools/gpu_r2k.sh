#!/bin/bash
# round-2 session K (8 GPUs): final driver-style native run (BERT + ResNet-50, re-planning on) and BASELINE.json config #5 (GPT-Sorter, 8-stage pure pipeline)
N=${1:-8}
mkdir -p gpurun_out
trun() { timeout -k 10 ${TMO:-500} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
show() { python - <<PY
import json
try:
    d = json.load(open('gpurun_out/$1.json'))
    for name, r in [(d['config']['model'][:14], d)] + list((d.get('also') or {}).items()):
        c = r['config']
        print('$1', name, round(r['value'], 1), r['unit'], 'ms/step', round(r['ms_per_step'], 3), 'e2e', r['e2e'] and round(r['e2e']['value'], 1), c['parallelism'], 'busy', r.get('stage_busy_ms_per_rank'),
              'exposed', round(r.get('exposed_comm_ms_per_step') or 0, 3), 'shares', c.get('stage_shares') and [round(v, 3) for v in c['stage_shares']], 'replicas', r.get('replica_weights'), r.get('invalid'))
except Exception as e:
    print('$1 no json', e)
PY
}
echo "== native n$N"; trun bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/k_native_n$N.log 2>&1; grep '^{"metric"' gpurun_out/k_native_n$N.log | tail -1 > gpurun_out/k_native_n$N.json; show k_native_n$N
grep -E "Error|Traceback" gpurun_out/k_native_n$N.log | head -5
echo "== sorter ${N}-stage"; trun bench.py --model sorter --gpus $N --steps 64 --warmup 8 --rebalance 0 > gpurun_out/k_sorter_n$N.log 2>&1; grep '^{"metric"' gpurun_out/k_sorter_n$N.log | tail -1 > gpurun_out/k_sorter_n$N.json; show k_sorter_n$N
grep -E "Error|Traceback" gpurun_out/k_sorter_n$N.log | head -5
