#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 300 $T tools/dist_retrieval_check.py > gpurun_out/dist_retrieval_check_p2p.json 2> gpurun_out/dist_retrieval_check_p2p.err; echo "check exit $?"
tail -1 gpurun_out/dist_retrieval_check_p2p.json; grep -n "Error\|error" gpurun_out/dist_retrieval_check_p2p.err | head -5
for ex in all_gather p2p; do
  timeout 300 $T bench.py --gpus 2 --mode retrieval --steps 5 --warmup 3 --exchange $ex > gpurun_out/bench_retrieval_n2_$ex.json 2> gpurun_out/bench_retrieval_n2_$ex.err; echo "bench $ex exit $?"
  tail -1 gpurun_out/bench_retrieval_n2_$ex.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['e2e']['value'])" || tail -5 gpurun_out/bench_retrieval_n2_$ex.err
done
