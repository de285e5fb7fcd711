#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_zipf.json 2> gpurun_out/bench_zipf.err
python bench.py --mode retrieval --steps 5 --warmup 3 > gpurun_out/bench_retrieval.json 2> gpurun_out/bench_retrieval.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --table-dtype bf16 --uniform-ids > gpurun_out/bench_bf16_uniform.json 2>/dev/null
for f in bench_zipf bench_retrieval bench_bf16_uniform; do python -c "import sys,json; d=json.loads(open('gpurun_out/$f.json').read()); print('$f', d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('kernel_ms'), (d.get('roofline') or {}).get('achieved'), (d.get('roofline') or {}).get('frac'), d['e2e']['value'], (d.get('cpu_baseline') or {}).get('value'))"; done
