#!/usr/bin/env python
"""Kernel-variant comparison on one B200 (builder tool, not the judged bench): loads each workload once and times the
fused pass for several tuning variants -- device-timed step (L2 flushed between steps) and the fused kernel alone.

    python tools/variant_bench.py [--steps 10] [--cases zipf,uniform,bf16,m1,m3] [--variants 2,7,8]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from manner_b200 import data as mdata  # noqa: E402
from manner_b200 import ops  # noqa: E402
from manner_b200.evaluator import ScoreEvaluator  # noqa: E402

CASES = {
    # name: (n_modules, uniform_ids, dtype)
    "zipf": (2, False, torch.float32),
    "uniform": (2, True, torch.float32),
    "bf16": (2, False, torch.bfloat16),
    "bf16_uniform": (2, True, torch.bfloat16),
    "m1": (1, False, torch.float32),
    "m3": (3, False, torch.float32),
}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cases", default="zipf,uniform,bf16,m1")
    ap.add_argument("--variants", default="2,7,8")
    ap.add_argument("--chunks-per-warp", default="1")
    ap.add_argument("--hot-kb", default="0", help="comma list of hot-row cache caps in KB (variants 8 / 9)")
    ap.add_argument("--shape", default="small")
    ap.add_argument("--probe", action="store_true", help="also run the read-bandwidth probe (L2-resident and HBM-sized buffers)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops.set_tuning(time_kernel=1)
    out = []
    if args.probe:
        for mib in (24, 48, 96, 1024, 4096):
            pr = ops.read_bandwidth_probe(dev, mib << 20, max(1, (8192 // mib)))
            rec = {"probe_buffer_MiB": pr["buffer_bytes"] >> 20, "read_GBps": pr["GBps"], "by_mode": pr["by_mode"]}
            print(json.dumps(rec), flush=True)
            out.append(rec)
    for case in args.cases.split(","):
        n_mod, uniform, dtype = CASES[case]
        tables, bhv = mdata.synth_workload(args.shape, n_modules=n_mod, uniform_ids=uniform, dtype=dtype)
        ev = ScoreEvaluator(tables, dev)
        d = ev.upload(bhv)
        w = torch.tensor([[1.0, 0.4, 0.2][:n_mod]], dtype=torch.float32, device=dev)
        algo = bhv.algorithmic_bytes(n_mod, 768, tables[0].element_size(), True)
        ref_sums = None
        for variant in [int(v) for v in args.variants.split(",")]:
            for cpw, hot_kb in [(int(c), int(h)) for c in args.chunks_per_warp.split(",") for h in (args.hot_kb.split(",") if variant in (8, 9) else ["0"])]:
                ops.set_tuning(variant=variant, chunks_per_warp=cpw, hot_kb_cap=hot_kb)
                for _ in range(3):
                    res = ev.evaluate(d, weights=w, zscore=True, pooled_auc=True)
                step_ms, kern_ms = [], []
                for _ in range(args.steps):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    p = ev.launch(d, weights=w, zscore=True, pooled_auc=True)
                    e1.record()
                    kern_ms.append(ops.last_score_kernel_ms())
                    torch.cuda.synchronize()
                    step_ms.append(e0.elapsed_time(e1))
                res = ev.finish(p)
                if ref_sums is None:
                    ref_sums = res.sums.copy()
                same = bool(abs(res.sums - ref_sums).max() <= 1e-9 * max(1.0, abs(ref_sums).max()))
                k = sum(kern_ms) / len(kern_ms)
                rec = {"case": case, "variant": variant, "chunks_per_warp": cpw, "hot_kb": hot_kb, "hot": ops.last_hot_stats() if variant in (8, 9) else None, "kernel_ms": round(k, 4), "kernel_ms_min": round(min(kern_ms), 4),
                       "step_ms": round(sum(step_ms) / len(step_ms), 4), "algo_TBps": round(algo / k / 1e9, 2), "sums_match_first_variant": same,
                       "ndcg10": round(res.metrics()["test/ndcg@10"], 6)}
                print(json.dumps(rec), flush=True)
                out.append(rec)
        del ev, d, tables
        torch.cuda.empty_cache()
    ops.set_tuning(variant=-1, chunks_per_warp=1, hot_kb_cap=0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "variant_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
