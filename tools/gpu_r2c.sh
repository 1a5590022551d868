#!/bin/bash
# round-2 session C (1 GPU): kernel tests incl. split-K / BN=192 GEMM and the dQ bulk-reduce FMHA backward; GEMM sweep; FMHA bench; bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== fmha bench"; timeout 300 python tools/fmha_bench.py 2>&1 | tail -2
echo "== gemm sweep"; timeout 600 python tools/gemm_bench.py --sweep 2>&1 | grep -v Warn | tail -20
echo "== native bert"; timeout 300 python bench.py --gpus 1 --steps 32 --warmup 5 --diagnose 8 > gpurun_out/native_bert_n1.json 2> gpurun_out/native_bert_n1.err; tail -c 300 gpurun_out/native_bert_n1.json
