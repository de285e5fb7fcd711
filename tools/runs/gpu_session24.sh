#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_par.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_par.log
tail -3 gpurun_out/pytest_par.log
for extra in "--table-dtype bf16 --variant 3" "--table-dtype bf16 --variant 5" "--table-dtype bf16 --variant 6" "--table-dtype bf16 --variant 5 --uniform-ids" "--table-dtype bf16 --variant 6 --uniform-ids" "--table-dtype bf16" ""; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $extra > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python -c "import json; d=json.load(open('gpurun_out/bench_x.json')); print('$extra |', round(d['value']/1e6,2), 'M/s step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'GB/s', round(d['roofline']['achieved']))" || tail -3 gpurun_out/bench_x.err
done
