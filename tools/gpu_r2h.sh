#!/bin/bash
# round-2 session H (4 GPUs): new kernel tests, ResNet-50 pp4 and BERT pp2xdp2 with profile-guided re-planning
N=${1:-4}
mkdir -p gpurun_out
trun() { timeout -k 10 ${TMO:-500} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
show() { python - <<PY
import json
try:
    d = json.load(open('gpurun_out/$1.json'))
    for name, r in [(d['config']['model'][:12], d)] + list((d.get('also') or {}).items()):
        c = r['config']
        print('$1', name, round(r['value'], 1), r['unit'], 'ms/step', round(r['ms_per_step'], 3), 'e2e', r['e2e'] and round(r['e2e']['value'], 1), c['parallelism'], 'busy', r.get('stage_busy_ms_per_rank'),
              'exposed', round(r.get('exposed_comm_ms_per_step') or 0, 3), 'shares', c.get('stage_shares') and [round(v, 3) for v in c['stage_shares']], 'before', c.get('stage_busy_before_replan_ms'), r.get('invalid'))
except Exception as e:
    print('$1 no json', e)
PY
}
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== resnet n$N"; trun bench.py --model resnet50 --gpus $N --steps 24 --warmup 8 --diagnose 8 > gpurun_out/h_resnet_n$N.log 2>&1; grep '^{"metric"' gpurun_out/h_resnet_n$N.log | tail -1 > gpurun_out/h_resnet_n$N.json; show h_resnet_n$N
grep -E "Error|Traceback" gpurun_out/h_resnet_n$N.log | head -5
echo "== resnet n$N two rounds"; trun bench.py --model resnet50 --gpus $N --steps 24 --warmup 8 --rebalance 2 --no-e2e > gpurun_out/h_resnet2_n$N.log 2>&1; grep '^{"metric"' gpurun_out/h_resnet2_n$N.log | tail -1 > gpurun_out/h_resnet2_n$N.json; show h_resnet2_n$N
echo "== bert n$N"; trun bench.py --gpus $N --steps 32 --warmup 8 --also "" > gpurun_out/h_bert_n$N.log 2>&1; grep '^{"metric"' gpurun_out/h_bert_n$N.log | tail -1 > gpurun_out/h_bert_n$N.json; show h_bert_n$N
grep -E "Error|Traceback" gpurun_out/h_bert_n$N.log | head -5
