#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_cache.py -m gpu -x -q > gpurun_out/pytest_modules.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_modules.log
tail -15 gpurun_out/pytest_modules.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_zipf_v3.json 2> gpurun_out/bench_zipf_v3.err
python -c "import json; d=json.load(open('gpurun_out/bench_zipf_v3.json')); print('eval', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e'], d['clocks'])"
C="python bench.py --mode retrieval --steps 1 --warmup 3 --no-cpu-baseline --users 37888 --catalog-per-gpu 262144"
$C > gpurun_out/plain_rt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:retrieve_topk -s 4 -c 1 -o gpurun_out/prof_retrieval_v2 $C > gpurun_out/ncu_rt.log 2>&1
tail -2 gpurun_out/ncu_rt.log
