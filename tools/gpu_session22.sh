#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pair.log
tail -8 gpurun_out/pytest_pair.log
for pr in 0 1; do
  for d in 0 2; do
    timeout 300 python bench.py --mode retrieval --steps 3 --warmup 3 --no-cpu-baseline --retrieval-pair $pr --retrieval-diag $d > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
    python -c "import json; d=json.load(open('gpurun_out/bench_x.json')); print('pair $pr diag $d', round(d['roofline']['kernel_ms'],2), round(d['roofline']['achieved'],1), d['clocks'])" || tail -3 gpurun_out/bench_x.err
  done
done
