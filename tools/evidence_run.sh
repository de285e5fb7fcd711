#!/bin/bash
# One-GPU evidence run (under gpurun): GPU tests, smoke, the benches of every mode, ncu launch lists and full captures.
# Outputs land in gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into the JSON summaries kept under profiles/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_zipf.json 2> gpurun_out/bench_zipf.err
python bench.py --steps 20 --warmup 3 --uniform-ids --no-cpu-baseline > gpurun_out/bench_uniform.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --modules 3 --sweep 121 > gpurun_out/bench_sweep121.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --table-dtype bf16 > gpurun_out/bench_bf16.json 2>/dev/null
python bench.py --mode retrieval --steps 5 --warmup 3 > gpurun_out/bench_retrieval.json 2> gpurun_out/bench_retrieval.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null
for f in bench_zipf bench_uniform bench_sweep121 bench_bf16 bench_retrieval bench_reference; do python -c "import sys,json; d=json.loads(open('gpurun_out/$f.json').read()); print('$f', d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('kernel_ms'), (d.get('roofline') or {}).get('achieved'), (d.get('roofline') or {}).get('frac'), d['e2e']['value'])"; done
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $C > gpurun_out/ncu_launches.log 2>&1
$C > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_eval_kernel -s 3 -c 2 -o gpurun_out/prof_zipf $C > gpurun_out/ncu_full.log 2>&1
R="python bench.py --mode retrieval --steps 1 --warmup 3 --no-cpu-baseline --users 37888 --catalog-per-gpu 524288"
$R > gpurun_out/plain_rt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:retrieve_topk -s 4 -c 1 -o gpurun_out/prof_retrieval $R > gpurun_out/ncu_rt.log 2>&1
R2="python bench.py --mode retrieval --steps 2 --warmup 3 --no-cpu-baseline"
$R2 > gpurun_out/plain_rt2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_retrieval.csv $R2 > gpurun_out/ncu_launches_rt.log 2>&1
