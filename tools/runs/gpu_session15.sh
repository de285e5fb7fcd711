#!/bin/bash
# 8 GPUs: BASELINE.json configs[4] (1 M users x 10 M news, top-100) and the eval scaling point
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 300 $T tools/dist_retrieval_check.py > gpurun_out/dist_retrieval_check_n8.json 2> gpurun_out/dist_retrieval_check_n8.err; echo "check exit $?"
tail -1 gpurun_out/dist_retrieval_check_n8.json
for ex in all_gather all_to_all; do
  timeout 600 $T bench.py --gpus 8 --mode retrieval --users 1000000 --steps 2 --warmup 3 --exchange $ex > gpurun_out/bench_retrieval_n8_$ex.json 2> gpurun_out/bench_retrieval_n8_$ex.err; echo "bench $ex exit $?"
  tail -1 gpurun_out/bench_retrieval_n8_$ex.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['e2e'], d['clocks'])" || tail -5 gpurun_out/bench_retrieval_n8_$ex.err
done
timeout 300 $T bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_v4.json 2> gpurun_out/bench_n8_v4.err; echo "eval n8 exit $?"
tail -1 gpurun_out/bench_n8_v4.json | cut -c1-330
