#!/bin/bash
# Round-2 one-GPU evidence run (under gpurun): sanitizers, ncu launch list + full captures, the default bench line, MIND-large on one GPU.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_l_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_l_pytest_gpu.log; tail -3 gpurun_out/r2_l_pytest_gpu.log
python tools/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1; tail -4 gpurun_out/sanitize_plain.log
bash tools/sanitize_run.sh
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --no-checks"
$C > gpurun_out/r2_l_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_l_launches.csv $C > gpurun_out/r2_l_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_eval_kernel -s 3 -c 1 -o gpurun_out/r2_l_prof_zipf $C > gpurun_out/r2_l_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"auc_rank_sum|auc_build_keys|reduce_partials|partition_kernel" -s 12 -c 4 -o gpurun_out/r2_l_prof_small $C > gpurun_out/r2_l_ncu_small.log 2>&1
python bench.py --steps 20 > gpurun_out/r2_l_bench_n1.json 2> gpurun_out/r2_l_bench_n1.err
python bench.py --steps 10 --workload large --shard --no-cpu-baseline --no-extra > gpurun_out/r2_l_bench_large_n1.json 2> gpurun_out/r2_l_bench_large_n1.err
python - <<EOF
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_l_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]/1e6,2), round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4), round(d["e2e"]["value"]/1e6,2), round(d["e2e"]["ms_per_step"],4), d["check"])
    except Exception as e: print(f, "ERR", e)
EOF
