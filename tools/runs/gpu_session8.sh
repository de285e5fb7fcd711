#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_early_fusion_loss.py -m gpu -x -q > gpurun_out/pytest_ef.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ef.log
tail -40 gpurun_out/pytest_ef.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zipf_v3.json 2> gpurun_out/bench_zipf_v3.err
python -c "import json; d=json.load(open('gpurun_out/bench_zipf_v3.json')); print('eval', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
