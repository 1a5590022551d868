#!/bin/bash
# round-2 session G (8 GPUs): K1 two-shot numerics + micro-benchmark, then the driver's scaling commands at N=8 (native: BERT + ResNet-50; reference: BERT)
N=${1:-8}
mkdir -p gpurun_out
export RVN_REF_TIMEOUT=600
trun() { timeout -k 10 ${TMO:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
echo "== K1 numerics"; TMO=200 trun tests/dist_gpu_worker.py 2>&1 | grep -E "DIST_GPU_OK|Error|assert" | head -4
echo "== K1 bench"; TMO=200 trun tools/k1_bench.py 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['mode'], d['optimizer'], d['params'], 'fused', d['fused_ms'], 'nccl+torch', d['nccl_plus_torch_ms'], 'algo floor', d['algorithmic_floor_ms'], 'frac', d['fraction_of_algorithmic_floor'])"
echo "== native n$N"; trun bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g_native_n$N.log 2>&1; grep '^{"metric"' gpurun_out/g_native_n$N.log | tail -1 > gpurun_out/g_native_n$N.json
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/g_native_n$N.json'))
    for name, r in [('bert', d)] + list((d.get('also') or {}).items()):
        print(name, round(r['value'], 1), r['unit'], 'ms/step', round(r['ms_per_step'], 3), 'e2e', r['e2e'] and round(r['e2e']['value'], 1), r['config']['parallelism'], 'busy', r.get('stage_busy_ms_per_rank'),
              'exposed', round(r.get('exposed_comm_ms_per_step') or 0, 3), 'replicas', r.get('replica_weights'), 'upd', r.get('optimizer_updates_in_window'), r.get('invalid'))
except Exception as e:
    print('no json', e)
PY
grep -E "Error|Traceback" gpurun_out/g_native_n$N.log | head -5
echo "== reference n$N"; trun bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/g_ref_n$N.log 2>&1; grep '"impl": "reference"' gpurun_out/g_ref_n$N.log | tail -1 > gpurun_out/g_ref_n$N.json; tail -c 600 gpurun_out/g_ref_n$N.json
