#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2; do
  timeout 300 python bench.py --mode retrieval --steps 3 --warmup 3 --no-cpu-baseline --retrieval-diag $d > gpurun_out/bench_retrieval_diag$d.json 2> gpurun_out/bench_retrieval_diag$d.err
  python -c "import json; d=json.load(open('gpurun_out/bench_retrieval_diag$d.json')); print('diag $d', d['roofline']['kernel_ms'], d['roofline']['achieved'])"
done
C="python bench.py --mode retrieval --steps 1 --warmup 3 --no-cpu-baseline --users 16384 --catalog-per-gpu 262144"
$C > gpurun_out/plain_rt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:retrieve_topk -s 4 -c 1 -o gpurun_out/prof_retrieval $C > gpurun_out/ncu_rt.log 2>&1
tail -3 gpurun_out/ncu_rt.log
