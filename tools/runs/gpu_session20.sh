#!/bin/bash
mkdir -p gpurun_out
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --modules 3 --sweep 121"
$C > gpurun_out/plain_sweep.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_eval_kernel -s 3 -c 1 -o gpurun_out/prof_sweep $C > gpurun_out/ncu_sweep.log 2>&1
tail -2 gpurun_out/ncu_sweep.log; python -c "
import json; d=json.loads(open('gpurun_out/plain_sweep.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
