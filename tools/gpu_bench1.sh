#!/bin/bash
# single-GPU bench sweep (native / baseline, graphs on/off); JSON lines land in gpurun_out/bench_*.json
mkdir -p gpurun_out
for cfg in "native:" "baseline:" "native:--no-graphs"; do
  impl=${cfg%%:*}; extra=${cfg#*:}
  name=bench_${impl}${extra// /}
  timeout -k 10 300 python bench.py --steps ${STEPS:-20} --warmup 5 --impl $impl $extra > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/bench_summary.txt
  grep '^{"metric"' gpurun_out/$name.log | tail -n 1 > gpurun_out/$name.json
  python - <<PY >> gpurun_out/bench_summary.txt
import json
try:
    d = json.load(open('gpurun_out/$name.json'))
    print('  value', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
except Exception as e:
    print('  no json', e)
PY
done
cat gpurun_out/bench_summary.txt
