#!/bin/bash
# round-2 session J (1 GPU): rectangular convs, pair-kernel routing, racecheck of the 2-CTA kernel, the three other configs at N=1
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== racecheck"; timeout 400 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file gpurun_out/sanitize_tcgen05_racecheck.log python tools/sanitize_targets.py > gpurun_out/sanitize_tcgen05_racecheck.out 2>&1; echo "racecheck exit=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitize_tcgen05_racecheck.log | tail -1) $(tail -1 gpurun_out/sanitize_tcgen05_racecheck.out)"
one() { name=$1; shift; timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/j_$name.out 2> gpurun_out/j_$name.err; grep '^{"metric"' gpurun_out/j_$name.out | tail -1 > gpurun_out/j_$name.json; python - <<PY
import json
try:
    r = json.load(open('gpurun_out/j_$name.json'))
    print('$name', round(r['value'], 1), r['unit'], 'ms/step', round(r['ms_per_step'], 3), 'e2e', r['e2e'] and round(r['e2e']['value'], 1), 'launches ours/all', r['gpu_launches_detail'].get('ours_per_step'), r['gpu_launches_detail'].get('all_per_step'))
except Exception as e:
    print('$name no json', e)
PY
}
one inception --model inception --steps 32 --warmup 5 --diagnose 4
grep "fallbacks" gpurun_out/diag_inception_n1_rank0.txt
one sorter --model sorter --steps 64 --warmup 8 --diagnose 4
grep "fallbacks" gpurun_out/diag_sorter_n1_rank0.txt
one resnet --model resnet50 --steps 24 --warmup 5 --diagnose 4
grep "fallbacks" gpurun_out/diag_resnet50_n1_rank0.txt
one bert --steps 32 --warmup 5 --also ""
