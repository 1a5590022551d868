#!/bin/bash
# Validation + measurement matrix for an N-GPU box, every run bounded by its own watchdog (a hang costs <= WD seconds, with a
# stack dump of every thread in the log).  Results: gpurun_out/matrix_n<N>.txt + one JSON per run.
#   tools/gpu_matrix.sh 1            # tests, smoke, BERT / ResNet / Inception / Sorter on one GPU, native + baseline
#   tools/gpu_matrix.sh 2|4|8        # dist kernel tests (K1, K2), K1 micro-bench, BERT + ResNet at the BASELINE.json topologies,
#                                    # plus (N >= 3) Inception 3-stage and (N == 8) Sorter 8-stage pipelines
N=${1:-1}
WD=${WD:-150}
mkdir -p gpurun_out
OUT=gpurun_out/matrix_n${N}.txt; : > $OUT
export RVN_BENCH_WATCHDOG=$WD NCCL_DEBUG=${NCCL_DEBUG:-WARN}
tr() { timeout -k 10 $((WD * 3)) python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) "${@:2}"; }
bench() {                       # bench <gpus> <name> <bench.py args...>
  local g=$1 name=$2; shift 2
  if [ "$g" = 1 ]; then timeout -k 10 $((WD * 3)) python bench.py "$@" > gpurun_out/$name.log 2>&1; else tr $g bench.py --gpus $g "$@" > gpurun_out/$name.log 2>&1; fi
  local rc=$?
  grep '^{"metric"' gpurun_out/$name.log | tail -n 1 > gpurun_out/$name.json
  python - "$name" "$rc" >> $OUT <<'PY'
import json, sys
name, rc = sys.argv[1], sys.argv[2]
try:
    d = json.load(open('gpurun_out/%s.json' % name))
    print('%-34s rc=%s %10.1f %-10s %8.3f ms/step  e2e %s  %s  busy %s  loss %s' % (name, rc, d['value'], d['unit'], d['ms_per_step'],
          d['e2e'] and round(d['e2e']['value'], 1), d['config']['parallelism'], d.get('stage_busy_ms_per_rank'), d.get('last_loss_per_rank')))
except Exception as err:
    print('%-34s rc=%s NO RESULT (%s)' % (name, rc, err))
PY
  grep -E "Timeout \(|Error" gpurun_out/$name.log | head -3 >> $OUT
}
if [ "$N" = 1 ]; then
  timeout -k 10 600 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/matrix_pytest.log 2>&1; echo "pytest -m gpu rc=$? $(tail -n 1 gpurun_out/matrix_pytest.log)" >> $OUT
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/matrix_smoke.log 2>&1; echo "smoke rc=$? $(tail -n 1 gpurun_out/matrix_smoke.log | cut -c1-120)" >> $OUT
  for m in bert resnet50 inception sorter; do
    bench 1 m1_${m}_native --model $m --steps 30 --warmup 5
    bench 1 m1_${m}_baseline --model $m --steps 30 --warmup 5 --impl baseline --no-graphs --no-e2e
  done
else
  tr 2 tests/dist_gpu_worker.py > gpurun_out/matrix_k1_test.log 2>&1; echo "K1 test rc=$? $(grep DIST_GPU_OK gpurun_out/matrix_k1_test.log)" >> $OUT
  tr 2 tests/dist_gpu_bcast_worker.py > gpurun_out/matrix_k2_test.log 2>&1; echo "K2 test rc=$? $(grep BCAST_GPU_OK gpurun_out/matrix_k2_test.log)" >> $OUT
  tr $N tools/k1_bench.py > gpurun_out/matrix_k1_bench.log 2>&1; echo "K1 bench rc=$?" >> $OUT; grep '^{' gpurun_out/matrix_k1_bench.log >> $OUT
  bench $N m${N}_bert_native --steps 32 --warmup 8
  bench $N m${N}_resnet50_native --model resnet50 --steps 24 --warmup 8
  bench $N m${N}_bert_baseline --steps 32 --warmup 8 --impl baseline --no-graphs --no-e2e
  [ "$N" -ge 3 ] && bench 3 m3_inception_pp3 --model inception --topology 3x1 --steps 24 --warmup 8
  [ "$N" -ge 4 ] && bench 4 m4_resnet50_pp4 --model resnet50 --topology 4x1 --steps 24 --warmup 8
  [ "$N" = 8 ] && bench 8 m8_sorter_pp8 --model sorter --topology 8x1 --steps 48 --warmup 8
fi
cat $OUT
