#!/bin/bash
# round-2 session M (1 GPU): BatchNorm statistics from the GEMM / convolution epilogues — numerics, then ResNet-50 / Inception / BERT N=1
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
timeout -k 10 400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "statistics or batchnorm or gemm_nt or gemm_2cta or conv2d_implicit or small_cin or fork or staging" 2>&1 | tail -15
echo "== resnet50 n1"
timeout -k 10 300 python bench.py --model resnet50 --gpus 1 --steps 20 --warmup 5 --diagnose 4 > gpurun_out/m_resnet_n1.log 2>&1; grep '^{"metric"' gpurun_out/m_resnet_n1.log | tail -1 > gpurun_out/m_resnet_n1.json; show m_resnet_n1
grep -E "Error|Traceback|error" gpurun_out/m_resnet_n1.log | head -8
head -40 gpurun_out/diag_resnet50_n1_rank0.txt 2>/dev/null | cut -c1-150
tail -2 gpurun_out/diag_resnet50_n1_rank0.txt | cut -c1-600
echo "== resnet50 n1, statistics kernel (RAVNEST_BN_STATS_FROM_CONV=0)"
RAVNEST_BN_STATS_FROM_CONV=0 timeout -k 10 300 python bench.py --model resnet50 --gpus 1 --steps 20 --warmup 5 > gpurun_out/m_resnet_n1_off.log 2>&1; grep '^{"metric"' gpurun_out/m_resnet_n1_off.log | tail -1 > gpurun_out/m_resnet_n1_off.json; show m_resnet_n1_off
echo "== inception n1"
timeout -k 10 200 python bench.py --model inception --gpus 1 --steps 20 --warmup 5 > gpurun_out/m_inception_n1.log 2>&1; grep '^{"metric"' gpurun_out/m_inception_n1.log | tail -1 > gpurun_out/m_inception_n1.json; show m_inception_n1
grep -E "Error|Traceback" gpurun_out/m_inception_n1.log | head -5
echo "== bert n1 (GEMM kernels gained a statistics branch: regression check)"
timeout -k 10 300 python bench.py --model bert --also "" --gpus 1 --steps 32 --warmup 5 > gpurun_out/m_bert_n1.log 2>&1; grep '^{"metric"' gpurun_out/m_bert_n1.log | tail -1 > gpurun_out/m_bert_n1.json; show m_bert_n1
grep -E "Error|Traceback" gpurun_out/m_bert_n1.log | head -5
