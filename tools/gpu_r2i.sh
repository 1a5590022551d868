#!/bin/bash
# round-2 session I (1 GPU): kernel tests (2-CTA GEMM under its own timeout), sweeps, N=1 bench, sanitizer, ncu captures
mkdir -p gpurun_out
echo "== 2cta tests"; timeout 240 python -m pytest tests/test_kernels_gpu.py -x -q -k "2cta" 2>&1 | tail -3
echo "== pytest gpu (rest)"; timeout 600 python -m pytest tests -m gpu -x -q -k "not 2cta" 2>&1 | tail -3
echo "== gemm sweep"; timeout 400 python tools/gemm_bench.py --sweep 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['name'], 'best', d['best'], d[d['best']], 'auto', d['auto_us'], '2cta', d.get('2cta_bn128'), d.get('2cta_bn256'), 'cublas', d['cublas_us'])"
echo "== gemm bench (rvn first)"; timeout 300 python tools/gemm_bench.py 2>&1 | grep "^{'name'" | cut -c1-220
echo "== fmha bench"; timeout 200 python tools/fmha_bench.py 2>&1 | tail -1
echo "== bench n1"; timeout 400 python bench.py --gpus 1 --steps 32 --warmup 5 > gpurun_out/i_native_n1.json 2> gpurun_out/i_native_n1.err; python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/i_native_n1.json') if l.startswith('{')][-1])
    for name, r in [('bert', d)] + list((d.get('also') or {}).items()):
        print(name, round(r['value'], 1), r['unit'], 'ms/step', round(r['ms_per_step'], 3), 'e2e', round(r['e2e']['value'], 1), r.get('invalid'))
except Exception as e:
    print('no json', e)
PY
echo "== sanitizer"; for tool in memcheck racecheck; do timeout 500 compute-sanitizer --tool $tool --error-exitcode 77 --log-file gpurun_out/sanitize_tcgen05_$tool.log python tools/sanitize_targets.py > gpurun_out/sanitize_tcgen05_$tool.out 2>&1; echo "$tool exit=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitize_tcgen05_$tool.log | tail -1) $(tail -1 gpurun_out/sanitize_tcgen05_$tool.out)"; done
echo "== ncu"; bash tools/gpu_ncu.sh gemm:5 gemm2:3 fmha:6 conv:9
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
