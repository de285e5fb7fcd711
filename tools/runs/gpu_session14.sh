#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_cache.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_mod.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mod.log
tail -6 gpurun_out/pytest_mod.log
for extra in "--modules 1" "--modules 1 --early-fusion" "--modules 1 --early-fusion --loss supcon" "--modules 1 --loss ce" ""; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $extra > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python -c "import json; d=json.load(open('gpurun_out/bench_x.json')); print('$extra |', round(d['value']/1e6,2), 'M/s step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value']/1e6,2), d['check'])" || tail -3 gpurun_out/bench_x.err
done
