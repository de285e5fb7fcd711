#!/usr/bin/env python
"""Tiny shapes through every kernel of libmanner_b200.so, as the target of

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_target.py --part eval|auc|exchange|retrieval|all

(VERDICT r1 item 9).  Every part prints "part <name> ok"; the sanitizer's own summary line follows in the log.
Shapes are tiny because the sanitizer slows kernels down 10-100x; they still take every code path: ragged tails, several
modules, aspects, the lane-per-weighting sweep, early fusion, both losses, bf16 rows, a non-reference width, the pipelined upload
with its gating word, the dynamic chunk hand-out, rank search with and without splitters, both retrieval pipelines.
Peer-GPU paths (row-sharded tables, p2p retrieval exchange, exchange with more than one rank) need two processes and are
covered functionally by tests/test_gpu_multigpu.py instead.
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def part_eval(dev) -> None:
    from manner_b200 import data as mdata
    from manner_b200 import ops
    from manner_b200.evaluator import ScoreEvaluator

    n_news = 384
    bhv = mdata.synth_behaviours(n_news, 96, seed=3, cand_window=200)
    aspects = mdata.synth_aspects(n_news)
    for dtype in (torch.float32, torch.bfloat16):
        tables = [mdata.synth_table(n_news, 768, s, dtype) for s in mdata.TABLE_SEEDS]
        ev = ScoreEvaluator(tables, dev, news_category=aspects["category"], news_sentiment=aspects["sentiment"])
        d = ev.upload(bhv)
        ev.evaluate(d, weights=[[1.0, 0.4, 0.2], [1.0, 0.0, 0.5]], zscore=True, want_scores=True, want_per_impression=True, pooled_auc=True)
        plain = ScoreEvaluator(tables, dev)
        grid = torch.tensor([[1.0, a / 4.0, b / 4.0] for a in range(5) for b in range(5)], dtype=torch.float32, device=dev)
        plain.evaluate(plain.upload(bhv), weights=grid, zscore=True)  # lane-per-weighting sweep (W = 25 >= 16)
        for schedule in (1, 2):
            ops.set_tuning(static_chunks=schedule, chunks_per_warp=3)
            plain.evaluate(plain.upload(bhv), weights=[[1.0, 0.4, 0.0]], zscore=True, pooled_auc=True)
        ops.set_tuning(static_chunks=0, chunks_per_warp=0)
    # non-reference width (predicated kernel), early fusion, both losses
    t128 = mdata.synth_table(n_news, 136, 5)
    g = torch.Generator().manual_seed(3)
    att = (torch.randn(24, 136, generator=g) * 0.1, torch.randn(24, generator=g) * 0.1, torch.rand(24, generator=g) * 0.2 - 0.1)
    ef = ScoreEvaluator([t128], dev, attention=[att])
    ef.evaluate(ef.upload(bhv, step_batch=8), want_scores=True, loss="ce", pooled_auc=True)
    lf = ScoreEvaluator([mdata.synth_table(n_news, 768, 6)], dev)
    lf.evaluate(lf.upload(bhv, step_batch=8), loss="supcon", temperature=0.36)
    # pipelined upload: gating word + dynamic hand-out, several passes over recycled buffers
    big = mdata.synth_behaviours(n_news, 1500, seed=4, cand_window=200)
    pinned = lf.pin(big)
    for seg in (2, 5):
        for _ in range(2):
            lf.evaluate(lf.upload(big, pinned, pipelined=True, segments=seg), pooled_auc=True)
    torch.cuda.synchronize()
    print("part eval ok", flush=True)


def part_auc(dev) -> None:
    from manner_b200 import metrics, ops

    g = np.random.default_rng(5)
    for n in (1, 37, 3000, 20000):  # below / above the splitter threshold of the rank search
        preds = torch.from_numpy(g.standard_normal(n).astype(np.float32) * 3).to(dev)
        labels = torch.from_numpy((g.random(n) < 0.2).astype(np.uint8)).to(dev)
        flags = torch.tensor([4], dtype=torch.int32, device=dev)
        torch.ops.manner_b200.pooled_auc(preds, labels, 2, flags)
        ties = torch.from_numpy((g.integers(0, 7, n) / 7.0).astype(np.float32)).to(dev)
        torch.ops.manner_b200.pooled_auc(ties, labels, 0, None)
        sk, pk, npos = ops.auc_build_and_sort(preds, labels, 1, None)
        s2 = torch.zeros(1, dtype=torch.int64, device=dev)
        ops.auc_rank_sum(sk, npos, pk, npos, s2)
    sizes = torch.from_numpy(g.integers(2, 40, 200))
    off = torch.zeros(201, dtype=torch.int32)
    off[1:] = torch.cumsum(sizes, 0)
    n = int(off[-1])
    preds = torch.from_numpy(g.standard_normal(n).astype(np.float32)).to(dev)
    labels = torch.from_numpy((g.random(n) < 0.2).astype(np.uint8)).to(dev)
    metrics.rank_metrics(preds, labels, off.to(dev), int(sizes.max()), want_per_impression=True)
    torch.cuda.synchronize()
    print("part auc ok", flush=True)


def part_exchange(dev) -> None:
    """The fused exchange kernels with one rank (mailbox in local memory): post, sigmoid keys, finish."""
    from manner_b200 import data as mdata
    from manner_b200.evaluator import ScoreEvaluator

    n_news = 384
    bhv = mdata.synth_behaviours(n_news, 400, seed=8, cand_window=200)
    ev = ScoreEvaluator([mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS[:2]], dev, exchange="p2p")
    for _ in range(3):
        ev.evaluate(ev.upload(bhv, pos_cap=int(bhv.labels.sum())), weights=[[1.0, 0.4]], zscore=True, pooled_auc=True, distributed=True)
    torch.cuda.synchronize()
    print("part exchange ok", flush=True)


def part_retrieval(dev) -> None:
    from manner_b200 import data as mdata
    from manner_b200 import ops
    from manner_b200 import retrieval as rt

    n_news = 700
    table = mdata.synth_table(n_news, 768, 9).to(dev)
    bhv = mdata.synth_behaviours(n_news, 300, seed=9, cand_window=300)
    users = rt.pool_users(table, torch.from_numpy(bhv.hist_offsets).to(dev), torch.from_numpy(bhv.hist_ids).to(dev))
    catalog = table.to(torch.bfloat16)
    for pair in (0, 1):
        ops.set_tuning(retrieval_pair=pair)
        s, i, _ = torch.ops.manner_b200.retrieve_topk(users, catalog, 10, 0, False)
        torch.cuda.synchronize()
    ops.set_tuning(retrieval_pair=1)
    rt.merge_topk(torch.stack([s, s]), torch.stack([i, i + n_news]))
    torch.cuda.synchronize()
    print("part retrieval ok", flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--part", default="all")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    parts = {"eval": part_eval, "auc": part_auc, "exchange": part_exchange, "retrieval": part_retrieval}
    for name, fn in parts.items():
        if args.part in ("all", name):
            fn(dev)


if __name__ == "__main__":
    main()
