#!/usr/bin/env python
"""Multi-GPU check of ROW-SHARDED embedding tables (run under torchrun, one rank per GPU): every rank holds 1/R of the rows
of each table, opens the peers' shards over CUDA IPC (dist.share_table_shards) and the fused kernel reads remote rows
directly over NVLink.  Scores, metric sums and the pooled AUROC must be bit-identical to the replicated-table run, because
the same rows go through the same arithmetic.  Also times both layouts on a MIND-small-shaped shard.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dist_sharded_table_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from manner_b200 import data as mdata
from manner_b200 import dist as mdist
from manner_b200.evaluator import ScoreEvaluator


def shard_of(table: torch.Tensor, rank: int, rows: int, dev: torch.device) -> torch.Tensor:
    out = torch.zeros(rows, table.shape[1], dtype=table.dtype, device=dev)
    part = table[rank * rows : (rank + 1) * rows]
    out[: part.shape[0]] = part.to(dev)
    return out


def main() -> None:
    rank, local_rank, world = mdist.init_from_env("nccl")
    dev = torch.device(f"cuda:{local_rank}")
    out = {"world": world}
    for name, n_news, n_impr in (("small", 5000, 20011), ("mind_small_shape", 65238, 73152)):
        bhv = mdata.synth_behaviours(n_news, n_impr, seed=3, cand_window=min(6000, n_news // 2))
        tables = [mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS[:2]]
        rows = 1 << int(np.ceil(np.log2(-(-n_news // world))))
        local = [shard_of(t, rank, rows, dev) for t in tables]
        shards = [mdist.share_table_shards(t) for t in local]
        mine = mdist.shard_for_rank(bhv, rank, world)
        pos_cap = mdist.agree_pos_cap(int(mine.labels.sum()), dev)
        kw = dict(weights=[[1.0, 0.4]], zscore=True, pooled_auc=True, want_scores=True, distributed=True)
        rep = ScoreEvaluator(tables, dev)
        shd = ScoreEvaluator([], dev, table_shards=shards, n_news=n_news)
        dev_bhv = rep.upload(mine, pos_cap=pos_cap)
        a = rep.evaluate(dev_bhv, **kw)
        b = shd.evaluate(dev_bhv, **kw)
        ok = bool(torch.equal(a.scores, b.scores)) and np.array_equal(a.sums, b.sums) and a.auc == b.auc
        t = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out[name + "_bit_identical"] = bool(t.item())
        times = {}
        for label, ev in (("replicated", rep), ("row_sharded", shd)):
            for it in range(6):
                dist.barrier()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ev.launch(dev_bhv, weights=[[1.0, 0.4]], zscore=True)
                e1.record()
                torch.cuda.synchronize(dev)
                if it >= 3:
                    times.setdefault(label, []).append(e0.elapsed_time(e1))
        ms = torch.tensor([np.mean(times["replicated"]), np.mean(times["row_sharded"])], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name + "_ms"] = {"replicated": round(float(ms[0]), 3), "row_sharded": round(float(ms[1]), 3), "impressions_per_rank": mine.n_impressions,
                             "remote_row_fraction": round(1 - 1 / world, 3)}
        del shd, shards, local
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    if not all(v for k, v in out.items() if k.endswith("bit_identical")):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
