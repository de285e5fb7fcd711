#!/usr/bin/env python
"""A/B of the fused kernel's chunk schedule and of the pipelined upload, in ONE process on one GPU (numbers from different gpurun
boxes differ by a few percent, which is the size of the effects looked at here).

    python tools/schedule_bench.py [--steps 20] [--table-dtype f32|bf16] [--uniform-ids]

Prints one JSON line per configuration:
  device: kernel ms / step ms with the behaviours resident, for static round-robin vs dynamic hand-out x chunks per warp
  e2e:    upload + pass + read-back ms for upload segments x chunks per warp
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--table-dtype", default="f32")
    ap.add_argument("--uniform-ids", action="store_true")
    ap.add_argument("--chunks", default="1,4,8,16,32")
    ap.add_argument("--segments", default="1,4,8,16")
    args = ap.parse_args()

    from manner_b200 import data as mdata
    from manner_b200 import ops
    from manner_b200.evaluator import ScoreEvaluator

    dev = torch.device("cuda:0")
    tdtype = torch.bfloat16 if args.table_dtype == "bf16" else torch.float32
    tables, bhv = mdata.synth_workload("small", n_modules=2, uniform_ids=args.uniform_ids, dtype=tdtype)
    ev = ScoreEvaluator(tables, dev)
    pinned = ev.pin(bhv)
    dev_bhv = ev.upload(bhv, pinned)
    w = torch.tensor([[1.0, 0.4]], dtype=torch.float32, device=dev)
    kw = dict(weights=w, zscore=True, pooled_auc=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops.set_tuning(time_kernel=1)
    ref = None

    def device_pass(static: int, cpw: int):
        nonlocal ref
        ops.set_tuning(static_chunks=static, chunks_per_warp=cpw)
        for _ in range(3):
            res = ev.finish(ev.launch(dev_bhv, **kw))
        if ref is None:
            ref = res
        same = bool(abs(res.sums - ref.sums).max() <= 1e-9 * abs(ref.sums).max()) and res.auc == ref.auc
        step_ms, kern_ms = [], []
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pending = ev.launch(dev_bhv, **kw)
            e1.record()
            kern_ms.append(ops.last_score_kernel_ms())
            torch.cuda.synchronize()
            step_ms.append(e0.elapsed_time(e1))
        ev.finish(pending)
        return statistics.mean(kern_ms), min(kern_ms), statistics.mean(step_ms), same

    chunks = [int(c) for c in args.chunks.split(",")]
    for rnd in range(args.rounds):
        for static in (1, 2):
            for cpw in chunks:
                k, kmin, st, same = device_pass(static, cpw)
                print(json.dumps({"kind": "device", "round": rnd, "schedule": "static" if static == 1 else "dynamic", "chunks_per_warp": cpw,
                                  "kernel_ms": round(k, 4), "kernel_ms_min": round(kmin, 4), "step_ms": round(st, 4), "same_results": same}), flush=True)

    # pooled AUROC: full sort of all keys vs sort of the positives + streaming pass over the negatives, on this set's scores
    ops.set_tuning(static_chunks=0, chunks_per_warp=0)
    res = ev.evaluate(dev_bhv, want_scores=True, **kw)
    flags = torch.tensor([4], dtype=torch.int32, device=dev)
    for name, fn in (("full_sort", lambda: torch.ops.manner_b200.pooled_auc(res.scores, dev_bhv.labels, 2, flags)),
                     ("sort_positives", lambda: torch.ops.manner_b200.pooled_auc_bounded(res.scores, dev_bhv.labels, 2, flags, dev_bhv.n_pos))):
        for _ in range(3):
            out = fn()
        ms = []
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        print(json.dumps({"kind": "pooled_auc", "path": name, "ms": round(statistics.mean(ms), 4), "ms_min": round(min(ms), 4), "auc": float(out[0]), "sum2": float(out[3])}), flush=True)

    for rnd in range(args.rounds):
        for cpw, deferred in ((16, False), (8, False)):
            for seg in [int(s) for s in args.segments.split(",")]:
                ops.set_tuning(static_chunks=0, chunks_per_warp=cpw)
                ms = []
                for i in range(args.steps + 3):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    d = ev.upload(bhv, pinned, pipelined=seg > 1, segments=seg)
                    r = ev.evaluate(d, **kw)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ms.append(e0.elapsed_time(e1))
                same = bool(abs(r.sums - ref.sums).max() <= 1e-9 * abs(ref.sums).max()) and r.auc == ref.auc
                print(json.dumps({"kind": "e2e", "round": rnd, "chunks_per_warp": cpw, "segments": seg, "deferred": deferred, "e2e_ms": round(statistics.mean(ms), 4),
                                  "e2e_ms_median": round(statistics.median(ms), 4), "e2e_ms_max": round(max(ms), 4), "same_results": same}), flush=True)
    ops.set_tuning(static_chunks=0, chunks_per_warp=0)


if __name__ == "__main__":
    main()
