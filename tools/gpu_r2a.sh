#!/bin/bash
# round-2 session A (1 GPU): reference arm at N=1 (BERT, ResNet-50), GPU tests, native bench
mkdir -p gpurun_out
export RVN_REF_TIMEOUT=500
echo "== ref bert"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/ref_bert_n1.json 2> gpurun_out/ref_bert_n1.err; tail -c 600 gpurun_out/ref_bert_n1.json
cp /tmp/rvn_refbench_*/provider_*.log gpurun_out/ 2>/dev/null; for f in gpurun_out/provider_*.log; do tail -c 3000 $f > $f.tail; rm $f; done
echo "== ref resnet"; timeout 600 python bench.py --impl reference --model resnet50 --gpus 1 --steps 10 --warmup 3 > gpurun_out/ref_resnet_n1.json 2> gpurun_out/ref_resnet_n1.err; tail -c 600 gpurun_out/ref_resnet_n1.json
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== native bert"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/native_bert_n1.json 2> gpurun_out/native_bert_n1.err; tail -c 400 gpurun_out/native_bert_n1.json
echo "== native resnet"; timeout 300 python bench.py --model resnet50 --gpus 1 --steps 20 --warmup 5 > gpurun_out/native_resnet_n1.json 2> gpurun_out/native_resnet_n1.err; tail -c 400 gpurun_out/native_resnet_n1.json
