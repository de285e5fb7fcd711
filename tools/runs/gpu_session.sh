#!/bin/bash
# One gpurun session: GPU tests, bench variants, ncu launch list + full capture (profiling recipe order:
# the plain command must exit 0 right before the same command runs under ncu).
mkdir -p gpurun_out
set -o pipefail
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
B="python bench.py --steps 20 --warmup 3"
$B > gpurun_out/bench_zipf.json 2> gpurun_out/bench_zipf.err
$B --uniform-ids --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err
for v in 0 1; do for c in 1 2 3; do for k in 1 8; do
  echo "variant=$v ctas=$c chunks=$k" >> gpurun_out/sweep.log
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant $v --ctas-per-sm $c --chunks-per-warp $k 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['achieved'])" >> gpurun_out/sweep.log
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --uniform-ids --variant $v --ctas-per-sm $c --chunks-per-warp $k 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('uniform', d['value'], d['roofline']['kernel_ms'], d['roofline']['achieved'])" >> gpurun_out/sweep.log
done; done; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --modules 1 2>/dev/null > gpurun_out/bench_m1.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --modules 3 2>/dev/null > gpurun_out/bench_m3.json
# ncu: launch list (every launch with its device time), then one full capture of the fused kernel
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $C > gpurun_out/ncu_launches.log 2>&1
$C > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_eval_kernel -s 3 -c 2 -o gpurun_out/prof_zipf $C > gpurun_out/ncu_full.log 2>&1
C2="$C --uniform-ids"
$C2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_eval_kernel -s 3 -c 2 -o gpurun_out/prof_uniform $C2 > gpurun_out/ncu_full_uniform.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep.log; ls -la gpurun_out
