#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pair.log
tail -4 gpurun_out/pytest_pair.log
for pr in 1 0; do
  timeout 300 python bench.py --mode retrieval --steps 5 --warmup 3 --no-cpu-baseline --retrieval-pair $pr > gpurun_out/bench_rt_pair$pr.json 2> gpurun_out/bench_x.err
  python -c "import json; d=json.load(open('gpurun_out/bench_rt_pair$pr.json')); print('pair $pr', round(d['roofline']['kernel_ms'],2), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), d['clocks'])" || tail -3 gpurun_out/bench_x.err
done
python tools/retrieval_stats.py
