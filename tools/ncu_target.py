#!/usr/bin/env python
"""Minimal target for `ncu --set full`: loads one workload, launches the fused pass a few times, exits.

    python tools/ncu_target.py --case zipf|uniform|bf16|bf16_uniform|m1|m3 [--aspects] [--variant V] [--launches 5]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from manner_b200 import data as mdata  # noqa: E402
from manner_b200 import ops  # noqa: E402
from manner_b200.evaluator import ScoreEvaluator  # noqa: E402
from tools.variant_bench import CASES  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="zipf")
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--launches", type=int, default=5)
    ap.add_argument("--aspects", action="store_true", help="with the category / sentiment labels: Diversity / Personalization metrics, full ranking")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n_mod, uniform, dtype = CASES[args.case]
    tables, bhv = mdata.synth_workload("small", n_modules=n_mod, uniform_ids=uniform, dtype=dtype)
    asp = mdata.synth_aspects(tables[0].shape[0]) if args.aspects else None
    ev = ScoreEvaluator(tables, dev, news_category=None if asp is None else asp["category"], news_sentiment=None if asp is None else asp["sentiment"])
    d = ev.upload(bhv)
    w = torch.tensor([[1.0, 0.4, 0.2][:n_mod]], dtype=torch.float32, device=dev)
    ops.set_tuning(variant=args.variant)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.launches):
        flush.fill_(1)
        res = ev.evaluate(d, weights=w, zscore=True, pooled_auc=True)
    print("ok", round(res.metrics()["test/ndcg@10"], 6))


if __name__ == "__main__":
    main()
