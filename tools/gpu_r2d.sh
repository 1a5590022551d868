#!/bin/bash
# round-2 session D (2 GPUs): pipeline with just-in-time NCCL rendezvous; NCCL p2p channel sweep
mkdir -p gpurun_out
tr() { timeout -k 10 ${TMO:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
run() { name=$1; shift; tr bench.py --gpus 2 "$@" > gpurun_out/$name.log 2>&1; grep '^{"metric"' gpurun_out/$name.log | tail -1 > gpurun_out/$name.json
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/$name.json'))
    print('$name', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), 'busy', d.get('stage_busy_ms_per_rank'), 'exposed', round(d.get('exposed_comm_ms_per_step') or 0, 3))
except Exception as e:
    print('$name no json', e)
PY
}
run d_bert_jit --steps 32 --warmup 8 --diagnose 8
RAVNEST_P2P_JIT=0 run d_bert_nojit --steps 32 --warmup 8 --no-e2e
NCCL_MAX_P2P_NCHANNELS=2 run d_bert_jit_ch2 --steps 32 --warmup 8 --no-e2e
NCCL_MAX_P2P_NCHANNELS=8 run d_bert_jit_ch8 --steps 32 --warmup 8 --no-e2e
run d_resnet_jit --model resnet50 --steps 24 --warmup 8 --diagnose 8
NCCL_MAX_P2P_NCHANNELS=4 run d_resnet_jit_ch4 --model resnet50 --steps 24 --warmup 8 --no-e2e
RAVNEST_P2P_JIT=0 run d_resnet_nojit --model resnet50 --steps 24 --warmup 8 --no-e2e
