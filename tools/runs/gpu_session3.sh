#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep3.log
for cfg in "0 3" "1 3" "2 4" "3 5" "3 6"; do set -- $cfg
  for ids in "" "--uniform-ids"; do
  echo -n "variant=$1 ctas=$2 $ids : " >> gpurun_out/sweep3.log
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant $1 --ctas-per-sm $2 --chunks-per-warp 1 $ids 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['e2e']['value'])" >> gpurun_out/sweep3.log
  done
done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep3.log
