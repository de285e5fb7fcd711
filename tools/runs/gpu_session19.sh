#!/bin/bash
# BASELINE.json configs[2]: MIND-large shape (2.37 M impressions, 161 k news), one set sharded over 8 GPUs (strong scaling)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $T bench.py --gpus 8 --workload large --shard --steps 20 --warmup 3 > gpurun_out/bench_large_shard_n8.json 2> gpurun_out/bench_large_shard_n8.err; echo "exit $?"
tail -1 gpurun_out/bench_large_shard_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['scaling'], d['roofline']['kernel_ms'], d['e2e'], d['check'], d['impressions_per_rank'])" || tail -5 gpurun_out/bench_large_shard_n8.err
