#!/bin/bash
# round-2 session O: driver-style validation — full GPU test suite, build + smoke, then the default bench (BERT + ResNet-50 under "also") at N GPUs
N=${1:-1}
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d = json.load(open('gpurun_out/$1.json'))
    for name, r in [(d['config']['model'][:14], d)] + list((d.get('also') or {}).items()):
        c = r['config']
        print('$1', name, round(r['value'], 1), r['unit'], 'ms/step', round(r['ms_per_step'], 3), 'e2e', r['e2e'] and round(r['e2e']['value'], 1), c['parallelism'], 'busy', r.get('stage_busy_ms_per_rank'),
              'exposed', round(r.get('exposed_comm_ms_per_step') or 0, 3), 'shares', c.get('stage_shares') and [round(v, 3) for v in c['stage_shares']], 'replicas', r.get('replica_weights'), r.get('invalid'), 'loss', r.get('last_loss_per_rank'))
except Exception as e:
    print('$1 no json', e)
PY
}
if [ "$N" = "1" ]; then
  echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
  echo "== build + smoke"; timeout -k 10 300 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -3
  echo "== bench n1"; timeout -k 10 400 python bench.py --gpus 1 > gpurun_out/o_native_n1.log 2>&1; grep '^{"metric"' gpurun_out/o_native_n1.log | tail -1 > gpurun_out/o_native_n1.json; show o_native_n1
  grep -E "Error|Traceback" gpurun_out/o_native_n1.log | head -5
else
  echo "== bench n$N"; timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) bench.py --gpus $N > gpurun_out/o_native_n$N.log 2>&1
  grep '^{"metric"' gpurun_out/o_native_n$N.log | tail -1 > gpurun_out/o_native_n$N.json; show o_native_n$N
  grep -E "Error|Traceback" gpurun_out/o_native_n$N.log | head -5
fi
