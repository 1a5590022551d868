#!/bin/bash
# round-2 session E (N GPUs): the driver's scaling commands — native arm (BERT + ResNet-50 in one job) and the reference arm
N=${1:-4}
mkdir -p gpurun_out
export RVN_REF_TIMEOUT=600
trun() { timeout -k 10 ${TMO:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
if [ "$2" = "tests" ]; then echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
echo "== native n$N"; trun bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_native_n$N.log 2>&1; grep '^{"metric"' gpurun_out/e_native_n$N.log | tail -1 > gpurun_out/e_native_n$N.json
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/e_native_n$N.json'))
    for name, r in [('bert', d)] + list((d.get('also') or {}).items()):
        print(name, round(r['value'], 1), r['unit'], 'ms/step', round(r['ms_per_step'], 3), 'e2e', r['e2e'] and round(r['e2e']['value'], 1), r['config']['parallelism'], 'busy', r.get('stage_busy_ms_per_rank'),
              'exposed', round(r.get('exposed_comm_ms_per_step') or 0, 3), 'replicas', r.get('replica_weights'), 'upd', r.get('optimizer_updates_in_window'), r.get('invalid'))
except Exception as e:
    print('no json', e)
PY
grep -E "Error|Traceback|error" gpurun_out/e_native_n$N.log | head -5
echo "== reference n$N"; trun bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/e_ref_n$N.log 2>&1; grep '"impl": "reference"' gpurun_out/e_ref_n$N.log | tail -1 > gpurun_out/e_ref_n$N.json; tail -c 700 gpurun_out/e_ref_n$N.json
echo; nvidia-smi --query-gpu=memory.used --format=csv | tr '\n' ' '
