#!/bin/bash
# round-2 session L (1 GPU): image stem on the implicit-GEMM kernel, conv1x1 fork, staged conv weight gradients — numerics, then ResNet-50 / Inception N=1
mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d = json.load(open('gpurun_out/$1.json'))
    print('$1', d['config']['model'][:14], round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), 'launches', d.get('gpu_launches'), d.get('invalid'), 'loss', d.get('last_loss_per_rank'))
except Exception as e:
    print('$1 no json', e)
PY
}
echo "== new kernel tests"
timeout -k 10 400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "small_cin or stem_pack or fork or staging or conv2d_implicit or rectangular" -s 2>&1 | tail -25
echo "== resnet50 n1 (all on)"
timeout -k 10 300 python bench.py --model resnet50 --gpus 1 --steps 20 --warmup 5 --diagnose 4 > gpurun_out/l_resnet_n1.log 2>&1; grep '^{"metric"' gpurun_out/l_resnet_n1.log | tail -1 > gpurun_out/l_resnet_n1.json; show l_resnet_n1
grep -E "Error|Traceback|error" gpurun_out/l_resnet_n1.log | head -8
head -45 gpurun_out/diag_resnet50_n1_rank0.txt 2>/dev/null | cut -c1-150
echo "== inception n1"
timeout -k 10 200 python bench.py --model inception --gpus 1 --steps 20 --warmup 5 > gpurun_out/l_inception_n1.log 2>&1; grep '^{"metric"' gpurun_out/l_inception_n1.log | tail -1 > gpurun_out/l_inception_n1.json; show l_inception_n1
grep -E "Error|Traceback" gpurun_out/l_inception_n1.log | head -5
