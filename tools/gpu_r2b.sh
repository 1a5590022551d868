#!/bin/bash
# round-2 session B (2 GPUs): pipeline diagnosis at pp2, reference arm at N=2
mkdir -p gpurun_out
export RVN_REF_TIMEOUT=500
tr() { timeout -k 10 ${TMO:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "$@"; }
echo "== stage calls"; timeout 300 python tools/stage_calls.py 2>&1 | grep -v Warning | tail -30
echo "== native bert n2 diag"; tr bench.py --gpus 2 --steps 32 --warmup 8 --diagnose 8 > gpurun_out/native_bert_n2.log 2>&1; grep '^{"metric"' gpurun_out/native_bert_n2.log | tail -1 > gpurun_out/native_bert_n2.json; tail -c 700 gpurun_out/native_bert_n2.json
echo "== native resnet n2 diag"; tr bench.py --model resnet50 --gpus 2 --steps 24 --warmup 8 --diagnose 8 > gpurun_out/native_resnet_n2.log 2>&1; grep '^{"metric"' gpurun_out/native_resnet_n2.log | tail -1 > gpurun_out/native_resnet_n2.json; tail -c 700 gpurun_out/native_resnet_n2.json
echo "== ref bert n2"; tr bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/ref_bert_n2.log 2>&1; grep '"impl": "reference"' gpurun_out/ref_bert_n2.log | tail -1 > gpurun_out/ref_bert_n2.json; tail -c 900 gpurun_out/ref_bert_n2.json
echo "== ref resnet n2"; tr bench.py --impl reference --model resnet50 --gpus 2 --steps 10 --warmup 3 > gpurun_out/ref_resnet_n2.log 2>&1; grep '"impl": "reference"' gpurun_out/ref_resnet_n2.log | tail -1 > gpurun_out/ref_resnet_n2.json; tail -c 900 gpurun_out/ref_resnet_n2.json
