#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_metrics.py -m gpu -x -q > gpurun_out/pytest_metrics.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_metrics.log
tail -30 gpurun_out/pytest_metrics.log
