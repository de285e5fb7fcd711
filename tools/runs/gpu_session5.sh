#!/bin/bash
# retrieval mode bring-up: parity tests first (bounded), then the whole GPU suite, then the retrieval bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -x -q > gpurun_out/pytest_retrieval.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_retrieval.log
tail -30 gpurun_out/pytest_retrieval.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --mode retrieval --steps 5 --warmup 3 > gpurun_out/bench_retrieval.json 2> gpurun_out/bench_retrieval.err; echo "bench exit $?"
cat gpurun_out/bench_retrieval.json; tail -5 gpurun_out/bench_retrieval.err
