#!/bin/bash
mkdir -p gpurun_out
for v in 4; do
  for extra in "" "--uniform-ids" "--modules 1"; do
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --variant $v $extra > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
    python -c "import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('variant $v $extra', round(d['value']/1e6,2), 'M/s step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3))" || tail -3 gpurun_out/bench_v$v.err
  done
done
