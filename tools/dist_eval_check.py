#!/usr/bin/env python
"""Multi-GPU check of the evaluation path (run under torchrun, one rank per GPU): one behaviour set sharded over the ranks
(step-aligned) must give the single-GPU numbers -- metric sums within 1e-12 relative, pooled AUROC exactly, test/loss within
fp64 regrouping -- for late fusion + ensemble and for early fusion + SupCon / cross-entropy loss.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_eval_check.py
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


def main() -> None:
    rank, local_rank, world = mdist.init_from_env("nccl", always=True)
    dev = torch.device(f"cuda:{local_rank}")
    n_news, dim = 5000, 768
    bhv = mdata.synth_behaviours(n_news, 20011, seed=3, cand_window=1500)
    tables = [mdata.synth_table(n_news, dim, s) for s in mdata.TABLE_SEEDS[:2]]
    g = torch.Generator().manual_seed(5)
    att = (torch.randn(200, dim, generator=g) * dim ** -0.5, torch.randn(200, generator=g) * 0.1, torch.rand(200, generator=g) * 0.2 - 0.1)
    shard = mdist.shard_for_rank(bhv, rank, world, align=8)
    pos_cap = mdist.agree_pos_cap(int(shard.labels.sum()), dev)
    out = {"world": world}
    cases = {
        "ensemble": (dict(tables=tables), dict(weights=[[1.0, 0.4]], zscore=True, pooled_auc=True)),
        "early_fusion_supcon": (dict(tables=tables[:1], attention=[att]), dict(pooled_auc=True, loss="supcon", temperature=0.36)),
        "late_fusion_ce": (dict(tables=tables[:1]), dict(pooled_auc=True, loss="ce")),
    }
    for name, (ctor, kw) in cases.items():
        for exchange in ("p2p", "nccl"):  # fused stores into the peers' mailboxes (default) / the three NCCL collectives
            ev = ScoreEvaluator(ctor["tables"], dev, attention=ctor.get("attention"), exchange=exchange)
            one = ev.evaluate(ev.upload(bhv, step_batch=8), **kw)  # the whole set on this GPU
            dev_shard = ev.upload(shard, pos_cap=pos_cap, step_batch=8)
            ok = True
            for rep in range(3):  # three passes: both mailbox parities are reused
                many = ev.evaluate(dev_shard, distributed=True, **kw)
                ok = ok and many.n_impressions == one.n_impressions and np.allclose(many.sums, one.sums, rtol=1e-12, atol=1e-9) and many.auc == one.auc
                ok = ok and many.auc_counts == one.auc_counts
                if one.loss is not None:
                    ok = ok and abs(many.loss - one.loss) <= 1e-12 * abs(one.loss)
            if exchange == "p2p":
                # the same sharded pass as a prepared pass (pipelined upload + fused exchange, manner_b200/prepared.py), interleaved with
                # the generic path on the same mailboxes
                pp = ev.prepare(shard, distributed=True, pos_cap=pos_cap, step_batch=8, **kw)
                for rep in range(3):
                    got = pp.run()
                    ok = ok and got.n_impressions == one.n_impressions and np.allclose(got.sums, one.sums, rtol=1e-12, atol=1e-9) and got.auc == one.auc
                    ok = ok and got.auc_counts == one.auc_counts
                    if one.loss is not None:
                        ok = ok and abs(got.loss - one.loss) <= 1e-12 * abs(one.loss)
                many = ev.evaluate(dev_shard, distributed=True, **kw)
                ok = ok and many.auc == one.auc
            # scores in the unit interval: AUROC without the sigmoid (the rule is decided over ALL ranks' scores)
            if name == "ensemble":
                unit = ScoreEvaluator([t * 0.02 + 0.03 for t in tables[:1]], dev, exchange=exchange)
                u_one = unit.evaluate(unit.upload(bhv), pooled_auc=True)
                u_many = unit.evaluate(unit.upload(shard, pos_cap=pos_cap), pooled_auc=True, distributed=True)
                ok = ok and u_many.auc == u_one.auc and not (u_one.flags & 4) and np.allclose(u_many.sums, u_one.sums, rtol=1e-12, atol=1e-9)
            t = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            out[f"{name}_{exchange}"] = bool(t.item())
        out[name + "_metrics"] = {k: round(v, 6) for k, v in many.metrics().items()}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    if not all(v for k, v in out.items() if isinstance(v, bool)):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
