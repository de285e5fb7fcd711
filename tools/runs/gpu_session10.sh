#!/bin/bash
# 2 GPUs: distributed retrieval correctness, then the retrieval bench with both exchanges, then the eval bench
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/dist_retrieval_check.py > gpurun_out/dist_retrieval_check.json 2> gpurun_out/dist_retrieval_check.err; echo "check exit $?"
cat gpurun_out/dist_retrieval_check.json; tail -3 gpurun_out/dist_retrieval_check.err
for ex in all_gather all_to_all; do
  timeout 300 $T bench.py --gpus 2 --mode retrieval --steps 5 --warmup 3 --exchange $ex > gpurun_out/bench_retrieval_n2_$ex.json 2> gpurun_out/bench_retrieval_n2_$ex.err; echo "bench $ex exit $?"
  tail -1 gpurun_out/bench_retrieval_n2_$ex.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['e2e']['value'])"
done
timeout 300 $T bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_v3.json 2> gpurun_out/bench_n2_v3.err; echo "eval n2 exit $?"
tail -1 gpurun_out/bench_n2_v3.json | cut -c1-400
