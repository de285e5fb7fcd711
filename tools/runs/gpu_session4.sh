#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_zipf.json 2> gpurun_out/bench_zipf.err
python bench.py --steps 20 --warmup 3 --uniform-ids --no-cpu-baseline > gpurun_out/bench_uniform.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --modules 1 > gpurun_out/bench_m1.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --modules 3 > gpurun_out/bench_m3.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $C > gpurun_out/ncu_launches.log 2>&1
$C > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_eval_kernel -s 3 -c 2 -o gpurun_out/prof_zipf $C > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for f in bench_zipf bench_uniform bench_m1 bench_m3 bench_reference; do python -c "import sys,json; d=json.loads(open('gpurun_out/$f.json').read()); print('$f', d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('kernel_ms'), (d.get('roofline') or {}).get('achieved'), d['e2e'])"; done
