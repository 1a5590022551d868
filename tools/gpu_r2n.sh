#!/bin/bash
# round-2 session N (1 GPU): stride-2 data gradients on the implicit-GEMM kernel (parity phases); where to collect BatchNorm statistics for pointwise convs
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
echo "== kernel tests"
timeout -k 10 400 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv2d or statistics or staging or fork" -s 2>&1 | grep -v "^stem mat" | tail -25
run() { name=$1; shift; env "$@" timeout -k 10 300 python bench.py --model resnet50 --gpus 1 --steps 20 --warmup 5 ${DIAG} > gpurun_out/$name.log 2>&1; grep '^{"metric"' gpurun_out/$name.log | tail -1 > gpurun_out/$name.json; show $name; grep -E "Error|Traceback" gpurun_out/$name.log | head -4; }
echo "== resnet50 n1: statistics in every epilogue"; DIAG="--diagnose 4" run n_resnet_all RAVNEST_BN_STATS_GEMM=all
grep -E "cutlass|cudnn|dgrad|fallbacks" gpurun_out/diag_resnet50_n1_rank0.txt | cut -c1-150
cp gpurun_out/diag_resnet50_n1_rank0.txt gpurun_out/diag_resnet50_n1_all.txt
echo "== resnet50 n1: pair-kernel GEMMs leave the statistics to BatchNorm"; DIAG="" run n_resnet_skip_pair RAVNEST_BN_STATS_GEMM=skip_pair
echo "== resnet50 n1: statistics GEMMs stay on the single-CTA kernel"; DIAG="" run n_resnet_nopair RVN_STATS_PAIR=0
echo "== resnet50 n1: library stride-2 dgrad (RAVNEST_STRIDE2_DGRAD=0)"; DIAG="" run n_resnet_s2lib RAVNEST_STRIDE2_DGRAD=0
