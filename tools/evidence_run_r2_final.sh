#!/bin/bash
# Round-2 final one-GPU evidence run (under gpurun): all GPU tests, smoke, the default bench line + the reference arm, ncu launch list of
# the bench, fresh ncu --set full captures of the shipped fused kernel on the three id / dtype cases.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_u_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_u_pytest_gpu.log; tail -3 gpurun_out/r2_u_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_u_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2_u_smoke.log | cut -c1-300
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_u_bench_n1.json 2> gpurun_out/r2_u_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_u_bench_reference.json 2> gpurun_out/r2_u_bench_reference.err
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-checks"
$C > gpurun_out/r2_u_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_u_launches.csv $C > gpurun_out/r2_u_ncu_launches.log 2>&1
for c in zipf uniform bf16; do
  python tools/ncu_target.py --case $c > gpurun_out/r2_u_target_$c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:score_eval_kernel -s 2 -c 2 -o gpurun_out/r2_u_prof_$c python tools/ncu_target.py --case $c > gpurun_out/r2_u_ncu_$c.log 2>&1
done
python - <<EOF
import json
for f in ("gpurun_out/r2_u_bench_n1.json", "gpurun_out/r2_u_bench_reference.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],4), d["e2e"]["value"], (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("kernel_ms"), d.get("api"))
    except Exception as e: print(f, "ERR", e)
EOF
