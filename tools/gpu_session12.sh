#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streaming" > gpurun_out/pytest_stream.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_stream.log
tail -30 gpurun_out/pytest_stream.log
for v in 2 4; do
  for extra in "" "--uniform-ids" "--modules 1" "--modules 3"; do
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --variant $v $extra > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
    python -c "import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('variant $v $extra', round(d['value']/1e6,2), 'M/s step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), d['check'])" || tail -3 gpurun_out/bench_v$v.err
  done
done
