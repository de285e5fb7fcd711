#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -x -q > gpurun_out/pytest_retrieval.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_retrieval.log
tail -25 gpurun_out/pytest_retrieval.log
for d in 0 1; do
  timeout 300 python bench.py --mode retrieval --steps 3 --warmup 3 --no-cpu-baseline --retrieval-diag $d > gpurun_out/bench_retrieval_diag$d.json 2> gpurun_out/bench_retrieval_diag$d.err
  python -c "import json; d=json.load(open('gpurun_out/bench_retrieval_diag$d.json')); print('diag $d', d['roofline']['kernel_ms'], d['roofline']['achieved'])"
done
timeout 300 python bench.py --mode retrieval --steps 3 --warmup 3 --no-cpu-baseline --users 37888 > gpurun_out/bench_retrieval_2waves.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/bench_retrieval_2waves.json')); print('2 waves', d['roofline']['kernel_ms'], d['roofline']['achieved'], d['clocks'])"
