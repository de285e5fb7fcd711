#!/usr/bin/env python
"""Where an end-to-end pass spends its time: host microseconds in upload() / launch() / finish(), and how long the GPU stream
idled between the step's first event and the start of the fused kernel (mb200_last_score_kernel_begin_after).

    python tools/host_overhead.py [--steps 30]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--segments", type=int, default=8)
    ap.add_argument("--worker-segments", type=int, default=0)
    ap.add_argument("--prepared", action="store_true", help="time ScoreEvaluator.prepare(...).run() instead of upload() + launch() + finish()")
    args = ap.parse_args()
    from manner_b200 import _native as nat
    from manner_b200 import data as mdata
    from manner_b200 import ops
    from manner_b200.evaluator import ScoreEvaluator

    dev = torch.device("cuda:0")
    tables, bhv = mdata.synth_workload("small", n_modules=2)
    ev = ScoreEvaluator(tables, dev)
    pinned = ev.pin(bhv)
    w = torch.tensor([[1.0, 0.4]], dtype=torch.float32, device=dev)
    kw = dict(weights=w, zscore=True, pooled_auc=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops.set_tuning(time_kernel=1)
    lib = nat.lib()
    rec = {k: [] for k in ("upload_us", "launch_us", "finish_us", "gpu_wait_before_kernel_ms", "kernel_ms", "e2e_ms")}
    pp = ev.prepare(bhv, pinned, segments=args.segments, worker_segments=args.worker_segments, **kw) if args.prepared else None
    for i in range(args.steps + 5):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        if pp is not None:
            res = pp.run()
            t1 = t2 = t3 = time.perf_counter()
        else:
            d = ev.upload(bhv, pinned, pipelined=args.segments > 1, segments=args.segments)
            t1 = time.perf_counter()
            pending = ev.launch(d, **kw)
            t2 = time.perf_counter()
            res = ev.finish(pending)
            t3 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            rec["upload_us"].append((t1 - t0) * 1e6), rec["launch_us"].append((t2 - t1) * 1e6), rec["finish_us"].append((t3 - t2) * 1e6)
            rec["gpu_wait_before_kernel_ms"].append(float(lib.mb200_last_score_kernel_begin_after(e0.cuda_event)))
            rec["kernel_ms"].append(ops.last_score_kernel_ms())
            rec["e2e_ms"].append(e0.elapsed_time(e1))
    print(json.dumps({k: round(statistics.median(v), 4) for k, v in rec.items()} | {"segments": args.segments, "worker_segments": args.worker_segments, "prepared": args.prepared, "auc": res.auc}))


if __name__ == "__main__":
    main()
