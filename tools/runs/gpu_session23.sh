#!/bin/bash
mkdir -p gpurun_out
R="python bench.py --mode retrieval --steps 1 --warmup 3 --no-cpu-baseline"
$R > gpurun_out/plain_rt3.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:retrieve_topk -s 4 -c 1 --csv --log-file gpurun_out/retrieval_bench_shape.csv $R > gpurun_out/ncu_rt3.log 2>&1
tail -3 gpurun_out/retrieval_bench_shape.csv
