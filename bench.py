#!/usr/bin/env python
"""Benchmark of the MANNeR scoring / ensemble / metrics hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload ...]

Metric (BASELINE.json): impressions/sec for pool + score + ensemble + metrics.  A *step* is one pass
of the hot path over one MIND-small-shaped set of impressions (73 152 impressions, 65 238 news,
768-d fp32): CR-Module + category A-Module ensemble (z-score, aspect weight), nDCG@5/10 + MRR +
gAUC per impression and the pooled AUROC -- BASELINE.json configs[1].

Prints ONE JSON line.  ``value`` is device-timed (CUDA events, max over ranks) with the inputs resident
in HBM; ``e2e`` is the same pass through ScoreEvaluator.upload + evaluate from pinned host buffers,
copies inside the timed region.  ``--impl reference`` times the oracle port of the reference's own CPU
path (steps of 8 impressions, per-row Python loops, torchmetrics-style group loop run twice) on the
host cores and prints the same line with "impl": "reference".

Other workloads of BASELINE.json (not the default line):
  --workload large --shard          configs[2]: MIND-large shape, one set sharded over the ranks (strong scaling)
  --modules 3 --sweep 121           configs[3]: aspect-weight sweep, 121 weightings from one gather
  --mode retrieval [--users U --catalog-per-gpu N --exchange all_gather|all_to_all|p2p]
                                    configs[4]: users x catalogue bf16 GEMM on tcgen05 + fused top-100, catalogue row-sharded
  --early-fusion / --loss ce|supcon late_fusion=False pooling and the reference's test/loss on device
  --uniform-ids / --table-dtype bf16  the HBM-bound id distribution / bf16 table storage
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

METRIC = "impressions/sec (pool+score+ensemble+metrics)"
UNIT = "impressions/s"
CATEG_WEIGHT = 0.4  # model.categ_weight (the YAMLs ship 0 and are swept by CLI override; 0 would not load the A-Module)
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md
FALLBACK_BF16_TFLOPS = 1590.0
FALLBACK_BF16_TFLOPS_SUSTAINED = 1400.0


def workload_config(args) -> dict:
    return {
        "workload": f"MIND-{args.workload}-shaped synthetic impressions, CR + category A-Module z-score ensemble, {getattr(args, 'table_dtype', 'f32').replace('f32', 'fp32')} 768-d tables",
        "shape": args.workload,
        "n_modules": args.modules,
        "categ_weight": CATEG_WEIGHT,
        "metrics": "ndcg@5 ndcg@10 mrr gauc + pooled auroc",
        "ids": "uniform" if args.uniform_ids else "zipf(1.05)",
        "fusion": "early (additive attention, cached logits)" if getattr(args, "early_fusion", False) else "late (mean)",
        "loss": getattr(args, "loss", None),
        "weightings": (max(2, int(round(args.sweep ** 0.5))) ** 2 if args.sweep else 1),
        "l2": "flushed between timed steps (256 MiB write); table 200 MB/module > 126 MB L2",
    }


def ncu_traffic(shape: str, n_modules: int, uniform_ids: bool, table_dtype: str):
    """(DRAM bytes per launch, L2 counters) of the ncu capture of exactly this workload kept in profiles/traffic.json, or (None, None)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            for t in json.load(f).get("entries", []):
                if (t.get("shape") == shape and t.get("n_modules") == n_modules and bool(t.get("uniform_ids")) == bool(uniform_ids)
                        and t.get("table_dtype", "f32") == table_dtype):
                    return t.get("dram_bytes_per_launch"), t.get("l2")
    return None, None


def peaks() -> tuple:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])), mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def cpu_reference_pass(tables, bhv, n_sample: int, threads: int, keep: bool = False) -> tuple:
    """Times the oracle port of the reference's CPU path on the first ``n_sample`` impressions.  ``keep`` also returns what it
    computed (flat scores + logged metrics) so the GPU arm can be checked against it on the same prefix."""
    from oracle import manner_oracle as mo

    torch.set_num_threads(threads)
    head = bhv.slice(0, min(n_sample, bhv.n_impressions))
    ob = mo.Behaviours(head.hist_offsets, head.hist_ids, head.cand_offsets, head.cand_ids, head.labels)
    weights = [1.0, CATEG_WEIGHT] + [0.0] * (len(tables) - 2)
    t0 = time.perf_counter()
    out = mo.ensemble_eval_epoch([t.float() for t in tables], weights[: len(tables)], ob, double_compute=True, reference_only=True)
    dt = time.perf_counter() - t0
    return head.n_impressions / dt, dt, head.n_impressions, (out if keep else None)


def parity_vs_cpu_prefix(ev, bhv, n: int, kw: dict, cpu_out: dict) -> dict:
    """The GPU path on the SAME impression prefix the CPU leg just evaluated: scores within 2e-5 (z-scores
    are O(1): the 1e-5-relative bar on either side), nDCG@5/10 within 1e-6 -- plus 1/B for every rank flip, and every flip
    must be a near-tie inside the score tolerance.  Raises when the parity bar is missed."""
    import numpy as np

    from oracle import manner_oracle as mo

    head = bhv.slice(0, n)
    kw = dict(kw, distributed=False, want_scores=True)
    res = ev.evaluate(ev.upload(head), **kw)
    got, ref = res.scores.cpu().numpy().astype(np.float64), cpu_out["scores"].astype(np.float64)
    slack = 2e-5 * np.maximum(1.0, np.abs(ref))
    err = np.abs(got - ref)
    flips, unexplained, gap = mo.unexplained_rank_flips(got, ref, slack, head.cand_offsets)
    m = res.metrics()
    deltas = {k: abs(m["test/" + k] - cpu_out["metrics"]["test/" + k]) for k in ("ndcg@5", "ndcg@10")}
    ok = bool(np.all(err <= slack)) and unexplained == 0 and all(d <= 1e-6 + flips / head.n_impressions for d in deltas.values())
    rec = {"parity_vs_cpu_prefix": ok, "prefix_impressions": head.n_impressions, "max_abs_score_diff": float(err.max()),
           "rank_flips": flips, "rank_flips_not_near_ties": unexplained, "widest_flipped_gap": gap,
           "ndcg_abs_diff": {k: float(v) for k, v in deltas.items()}}
    if not ok:
        raise SystemExit("bench.py: GPU result does not match the CPU reference leg on the same prefix: " + json.dumps(rec))
    return rec


def shard_invariance_probe(ev, kw: dict, rank: int, world: int, dev) -> dict:
    """A FIXED probe set (the first 8 192 impressions of the seed-42 MIND-small-shaped behaviours) evaluated whole on every rank
    and sharded ``world`` ways through the distributed path: the reduced sums must agree to 1e-12 and the pooled AUROC exactly
    (VERDICT r1 item 2).  With one rank the probe is split three ways and the parts are added up."""
    import numpy as np

    from manner_b200 import data as mdata
    from manner_b200 import dist as mdist

    n_news = mdata.SHAPES["small"][0]
    probe = mdata.synth_behaviours(n_news, 8192, 42)
    kw = {k: v for k, v in kw.items() if k != "distributed"}
    one = ev.evaluate(ev.upload(probe), **kw)
    if world > 1:
        shard = mdist.shard_for_rank(probe, rank, world)
        cap = mdist.agree_pos_cap(int(shard.labels.sum()), dev)
        many = ev.evaluate(ev.upload(shard, pos_cap=cap), distributed=True, **kw)
        sums, auc, n = many.sums, many.auc, many.n_impressions
    else:
        bounds = mdata.balanced_shard_bounds(probe, 3)
        parts = [ev.evaluate(ev.upload(probe.slice(int(bounds[r]), int(bounds[r + 1]))), **{k: v for k, v in kw.items() if k != "pooled_auc"}) for r in range(3)]
        sums, auc, n = sum(p.sums for p in parts), one.auc, sum(p.n_impressions for p in parts)
    ok = n == one.n_impressions and bool(np.allclose(sums, one.sums, rtol=1e-12, atol=1e-9)) and auc == one.auc
    t = torch.tensor([int(ok)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
    return {"shard_invariant": bool(t.item()), "probe_impressions": probe.n_impressions, "ways": max(world, 3),
            "probe_ndcg@10": round(one.metrics()["test/ndcg@10"], 9), "probe_auc": one.auc}


def timed_passes(ev, dev_bhv, kw: dict, steps: int, flush, rank: int = 0) -> tuple:
    """(mean device-timed ms per pass, mean fused-kernel ms) over ``steps`` passes after 3 warm-up passes, L2 flushed between."""
    from manner_b200 import ops

    for _ in range(3):
        ev.finish(ev.launch(dev_bhv, **kw))
    step_ms, kernel_ms = [], []
    for _ in range(steps):
        flush.fill_(rank + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending = ev.launch(dev_bhv, **kw)
        e1.record()
        kernel_ms.append(ops.last_score_kernel_ms())
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    ev.finish(pending)
    return sum(step_ms) / len(step_ms), sum(kernel_ms) / len(kernel_ms)


def shared_synth_behaviours(args, rank: int, world: int):
    """The one synthetic behaviour set of a strong-scaling run.  Generating the MIND-large shape takes minutes of numpy time, so
    rank 0 does it once per box and leaves the arrays in /dev/shm; the other ranks (and later runs on the same box) load them."""
    import numpy as np
    import torch.distributed as dist

    from manner_b200 import data as mdata

    n_news, n_impr, seed = mdata.SHAPES[args.workload]
    path = f"/dev/shm/mb200_bhv_{args.workload}_{seed}_{int(args.uniform_ids)}.npz"
    names = ("hist_offsets", "hist_ids", "cand_offsets", "cand_ids", "labels")
    if rank == 0 and not os.path.exists(path):
        b = mdata.synth_behaviours(n_news, n_impr, seed, uniform_ids=args.uniform_ids)
        try:
            np.savez(path + ".tmp.npz", **{k: getattr(b, k) for k in names})
            os.replace(path + ".tmp.npz", path)
        except OSError:
            pass  # no shared memory file system: every rank generates its own copy below
    if world > 1:
        dist.barrier()
    if os.path.exists(path):
        z = np.load(path)
        return mdata.Behaviours(*[z[k] for k in names])
    return mdata.synth_behaviours(n_news, n_impr, seed, uniform_ids=args.uniform_ids)


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    from manner_b200 import data as mdata

    threads = os.cpu_count() or 1
    tables, bhv = mdata.synth_workload(args.workload, n_modules=args.modules, uniform_ids=args.uniform_ids)
    # bound the whole run to a few minutes whatever K and W the driver passes (~550 impressions/s on 8 cores)
    budget_s = 150.0 / (args.steps + 0.25 * args.warmup)
    n_sample = int(min(args.cpu_sample, max(256, budget_s * 500))) // 8 * 8
    for _ in range(args.warmup):
        cpu_reference_pass(tables, bhv, max(64, n_sample // 16), threads)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt, n, _ = cpu_reference_pass(tables, bhv, n_sample, threads)
        rates.append(r), times.append(dt)
    value = n_sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n_sample} impressions of the workload per step; oracle port of EnsembleModule.test_step (batches of 8, "
                      "per-row loops) + torchmetrics-style nDCG@5/10 group loop run twice, torch.set_num_threads(cores)",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_gpu_arm(args) -> None:
    import torch.distributed as dist

    from manner_b200 import data as mdata
    from manner_b200 import dist as mdist
    from manner_b200 import ops
    from manner_b200.evaluator import ScoreEvaluator

    os.environ.setdefault("NCCL_DEBUG", "WARN")  # a driver that sets NCCL_DEBUG=INFO keeps it; the JSON line is printed last
    rank, local_rank, world = mdist.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device(f"cuda:{local_rank}")
    distributed = world > 1

    tdtype = torch.bfloat16 if args.table_dtype == "bf16" else torch.float32
    if args.shard:
        # strong scaling (BASELINE.json configs[2]): ONE set of impressions, sharded over the ranks by rows gathered
        full = shared_synth_behaviours(args, rank, world)  # rank 0 generates (minutes for MIND-large), the others map its arrays
        n_news = mdata.SHAPES[args.workload][0]
        tables = [mdata.synth_table(n_news, 768, mdata.TABLE_SEEDS[m], tdtype) for m in range(args.modules)]
        bhv = mdist.shard_for_rank(full, rank, world)
        del full
    else:
        # weak scaling: every rank scores its own MIND-small-shaped shard (different behaviour seed), tables replicated
        tables, bhv = mdata.synth_workload(args.workload, n_modules=args.modules, seed_offset=rank, uniform_ids=args.uniform_ids, dtype=tdtype)
    attention = None
    if args.early_fusion:
        # late_fusion=False (configs/experiment/cr_module_mind_all_scl_ef.yaml): additive attention with query_vector_dim 200 on the CR
        # module; the per-news logits are computed once here (ops.attention_logits), like the embedding table
        ga = torch.Generator().manual_seed(77)
        dim = tables[0].shape[1]
        attention = [(torch.randn(200, dim, generator=ga) * dim ** -0.5, torch.randn(200, generator=ga) * 0.1, torch.rand(200, generator=ga) * 0.2 - 0.1)]
        attention += [None] * (len(tables) - 1)
    step_batch = 8 if (args.early_fusion or args.loss) else None  # configs/data/mind_rec.yaml:51
    ev = ScoreEvaluator(tables, dev, attention=attention, exchange=args.eval_exchange)
    pinned = ev.pin(bhv, step_batch)
    # multi-GPU pooled AUC: agree once (outside the timed loop) on the largest per-rank positive count
    pos_cap = mdist.agree_pos_cap(int(bhv.labels.sum()), dev) if distributed else None
    dev_bhv = ev.upload(bhv, pinned, pos_cap)
    weights = [[1.0, CATEG_WEIGHT] + [0.0] * (args.modules - 2)][0][: args.modules]
    w_dev = torch.tensor([weights], dtype=torch.float32, device=dev)
    kw = dict(weights=w_dev, zscore=True, pooled_auc=True, distributed=distributed)
    if args.loss:
        kw.update(loss=args.loss, temperature=0.36)
    if args.sweep:
        # BASELINE.json configs[3]: aspect-weight sweep, every weighting re-scored from the one gather
        side = max(2, int(round(args.sweep ** 0.5)))
        grid = [[1.0] + ([a / (side - 1)] if args.modules > 1 else []) + ([b / (side - 1)] if args.modules > 2 else []) + [0.0] * max(0, args.modules - 3)
                for a in range(side) for b in range(side)]
        w_dev = torch.tensor(grid, dtype=torch.float32, device=dev)
        kw = dict(weights=w_dev, zscore=True, pooled_auc=False, distributed=distributed)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops.set_tuning(time_kernel=1)
    if args.variant is not None:
        ops.set_tuning(variant=args.variant)
    if args.chunks_per_warp is not None:
        ops.set_tuning(chunks_per_warp=args.chunks_per_warp)
    if args.ctas_per_sm is not None:
        ops.set_tuning(ctas_per_sm=args.ctas_per_sm)

    def barrier() -> None:
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # Both timed loops go through ScoreEvaluator.prepare(...): the public call for a split that is evaluated again and again (after
    # every epoch) -- buffers, descriptors and workspaces are set up once, a pass is a handful of C-ABI calls, so the stream does not
    # wait for Python between the kernels of a 1.7 ms step.  --generic times launch() / upload() + evaluate() instead.
    use_prepared = not args.generic and (not distributed or args.eval_exchange == "p2p")
    prep_kw = dict(weights=kw["weights"], zscore=True, pooled_auc=kw.get("pooled_auc", False), loss=kw.get("loss"),
                   temperature=kw.get("temperature", 0.1), step_batch=step_batch, distributed=distributed, pos_cap=pos_cap)
    resident_pass = ev.prepare(bhv, pinned, resident=True, **prep_kw) if use_prepared else None

    def device_step(timed: bool):
        flush.fill_(rank + 1)  # evict L2 between steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending = resident_pass.launch() if resident_pass is not None else ev.launch(dev_bhv, **kw)
        e1.record()
        return e0, e1, pending

    def read_back(pending):
        return resident_pass.read() if resident_pass is not None else ev.finish(pending)

    for _ in range(max(args.warmup, 3)):
        _, _, pending = device_step(False)
    res = read_back(pending)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ops.launch_counts()
    barrier()
    wall0 = time.perf_counter()
    events, kernel_ms = [], []
    for _ in range(args.steps):
        e0, e1, pending = device_step(True)
        events.append((e0, e1))
        kernel_ms.append(ops.last_score_kernel_ms())  # waits for the fused kernel of this step only
    barrier()
    wall = time.perf_counter() - wall0
    launches1 = ops.launch_counts()
    step_ms = [a.elapsed_time(b) for a, b in events]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    res = read_back(pending)
    # per-rank view of the same steps: which part of a multi-GPU step is the fused kernel on the slowest GPU, which part the rendezvous
    per_rank = None
    if distributed:
        mine = torch.tensor([sum(step_ms) / len(step_ms), sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device=dev)
        allr = torch.zeros(world * 2, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, 2).cpu().tolist()
        per_rank = {"step_ms": [round(a, 4) for a, _ in allr], "kernel_ms": [round(b, 4) for _, b in allr]}

    # the clocks are sampled during the device-timed region only: nvidia-smi polling takes driver locks that can stall
    # the host-side calls of the end-to-end loop below for a whole polling period
    clocks = sampler.stop() if sampler is not None else None

    # end to end: pinned host CSR -> device, pass, metric sums back to the host, every step
    barrier()
    gc.collect()
    gc.disable()  # a generation-2 collection of the interpreter (tens of ms) inside a 2 ms step is host noise, not the path
    e2e_events = []
    e2e_warm = max(args.warmup, 3)  # the end-to-end path gets its own warm-up steps (first uploads, allocator growth, sampler teardown)
    # end to end: run() = pipelined upload from the pinned host CSR (offsets first, the fused kernel starts on the first segment while
    # the copy stream brings the rest) + pass + one pinned read-back
    prepared = ev.prepare(bhv, pinned, segments=max(args.upload_segments, 1), **prep_kw) if use_prepared else None
    for i in range(max(args.steps, 10) + e2e_warm):
        flush.fill_(rank + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if prepared is not None:
            r = prepared.run()  # upload + pass + device -> host read of sums / AUC statistics
        else:
            step_bhv = ev.upload(bhv, pinned, pos_cap, pipelined=args.upload_segments > 1, segments=args.upload_segments)
            r = ev.evaluate(step_bhv, **kw)  # includes the device -> host read of sums / AUC statistics
        e1.record()
        torch.cuda.synchronize(dev)
        if i >= e2e_warm:
            e2e_events.append(e0.elapsed_time(e1))
    gc.enable()
    e2e_median_ms = statistics.median(e2e_events)
    e2e_ms = torch.tensor([sum(e2e_events) / len(e2e_events)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    n_impr_rank = bhv.n_impressions
    n_impr_total = torch.tensor([n_impr_rank], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(n_impr_total, op=dist.ReduceOp.SUM)
    n_impr_total = float(n_impr_total.item())

    # multi-GPU: a fixed probe set sharded `world` ways must reproduce the single-GPU numbers (all ranks take part)
    shard_check = None
    if args.mode == "eval" and not args.sweep and not args.loss and not args.early_fusion and not args.no_checks:
        shard_check = shard_invariance_probe(ev, kw, rank, world, dev)

    if rank != 0:
        if distributed:
            dist.barrier()
            dist.destroy_process_group()
        return

    hbm_peak, peak_kind = peaks()
    algo_bytes = bhv.algorithmic_bytes(args.modules, tables[0].shape[1], tables[0].element_size(), scores_written=True)
    k_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic, l2_note = (None, None)
    if not args.sweep and not args.early_fusion:  # ncu captures of this exact workload with the default kernel
        traffic, l2_note = ncu_traffic(args.workload, args.modules, args.uniform_ids, args.table_dtype)

    # The roof that binds.  Zipf-shaped ids: ~83 % of the gathered sectors hit the 126 MB L2, the rows come
    # over the L2 -> SM crossbar, so the denominator is the L2-resident read bandwidth of the same access shape, measured here
    # (mb200_read_probe on a 48 MiB buffer).  Uniform ids: HBM (MEASURED_PEAKS.json copy bandwidth; the read-only probe on a
    # 6 GiB buffer is reported beside it).
    l2_probe = ops.read_bandwidth_probe(dev, 48 << 20, 160)
    hbm_probe = ops.read_bandwidth_probe(dev, 6 << 30, 2)
    bound = "hbm" if args.uniform_ids else "l2"
    peak = hbm_peak if bound == "hbm" else l2_probe["GBps"]
    roofline = {
        "bound": bound, "kernel": "score_eval_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_kind": (peak_kind + " HBM copy bandwidth (MEASURED_PEAKS.json)") if bound == "hbm" else
                     "measured in this run: L2-resident row-gather read bandwidth (mb200_read_probe, 48 MiB buffer, best of 3 launch shapes)",
        "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": k_ms,
        "kernel_share_of_step": k_ms * args.steps / total_ms if total_ms > 0 else None,
        "l2_read_probe": l2_probe, "hbm_read_probe": hbm_probe, "hbm_copy_peak": hbm_peak, "l2": l2_note,
        # what actually crossed the HBM interface per second (ncu DRAM bytes of the same launch / live kernel time)
        "dram_achieved": (traffic / (k_ms * 1e-3) / 1e9) if traffic else None,
        "dram_frac": (traffic / (k_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
    }

    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, dt, n, cpu_out = cpu_reference_pass(tables, bhv, args.cpu_sample, threads, keep=True)
        cpu = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} impressions of the workload ({dt:.1f} s); oracle port of EnsembleModule.test_step (batches of 8, per-row loops) "
                      "+ torchmetrics-style nDCG@5/10 group loop run twice",
        }
        if not args.sweep and not args.loss and not args.early_fusion and not args.no_checks:
            parity = parity_vs_cpu_prefix(ev, bhv, n, kw, cpu_out)  # raises when the GPU result misses the parity bar

    extra = None
    if world == 1 and args.extra and not (args.sweep or args.loss or args.early_fusion or args.uniform_ids or args.table_dtype != "f32" or args.shard):
        extra = extra_results(args, ev, tables, bhv, dev, flush, hbm_peak, l2_probe["GBps"])

    check = {k: round(v, 6) for k, v in res.metrics().items()}
    if shard_check:
        check.update(shard_check)
    if parity:
        check.update(parity)
    line = {
        "metric": METRIC, "value": n_impr_total * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.shard else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),  # arithmetic is fp32 for either table dtype
        "e2e": {
            "value": n_impr_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": dev_bhv.h2d_bytes,
            "d2h_bytes_per_step": r.d2h_bytes, "ms_per_step": e2e_ms, "ms_per_step_median_rank0": e2e_median_ms, "steps": len(e2e_events),
            "ms_each_rank0": [round(x, 3) for x in e2e_events],
            "api": "ScoreEvaluator.prepare(...).run()" if prepared is not None else "ScoreEvaluator.upload(...) + evaluate(...)",
            "upload": (f"pipelined: offsets, then {args.upload_segments} segments of geometrically growing size on a copy stream overlapped with the fused kernel "
                       "(mb200_upload_begin / _finish)") if args.upload_segments > 1 else "whole set copied in front of the pass",
        },
        "api": "ScoreEvaluator.prepare(resident=True).launch()" if resident_pass is not None else "ScoreEvaluator.launch()",
        "gpu_launches": launches1[0] - launches0[0],
        "library_launches": launches1[1] - launches0[1],
        "exchange": ev.exchange if distributed else "none",
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "wall_s_timed_region": wall,
        "impressions_per_rank": n_impr_rank,
        "per_rank": per_rank,
        "check": check,
        "extra": extra,
    }
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
        time.sleep(0.5)  # the other ranks tear NCCL down at the same moment: let whatever NCCL_DEBUG makes them say come first
    sys.stdout.flush()
    print(json.dumps(line), flush=True)  # last line of stdout, after NCCL has said whatever NCCL_DEBUG asked it to say


def extra_results(args, ev, tables, bhv, dev, flush, hbm_peak: float, l2_peak: float) -> dict:
    """Compact records of the other BASELINE.json configurations in the default line, a few hundred ms of GPU
    time each: the HBM-bound id distribution, bf16 tables, the 121-weighting sweep over three modules, full-catalogue retrieval."""
    from manner_b200 import data as mdata
    from manner_b200.evaluator import ScoreEvaluator

    steps = max(5, min(args.steps, 10))
    n_news = tables[0].shape[0]
    w2 = torch.tensor([[1.0, CATEG_WEIGHT]], dtype=torch.float32, device=dev)
    out = {}

    def record(name, evaluator, behaviours, kw, n_mod, elem, peak, bound, extra_fields=None, traffic=None):
        d = evaluator.upload(behaviours)
        step_ms, k_ms = timed_passes(evaluator, d, kw, steps, flush)
        algo = behaviours.algorithmic_bytes(n_mod, 768, elem, scores_written=True)
        rec = {"impressions_per_s": behaviours.n_impressions / (step_ms * 1e-3), "ms_per_step": step_ms, "kernel_ms": k_ms,
               "roofline": {"bound": bound, "achieved": algo / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": algo / (k_ms * 1e-3) / 1e9 / peak,
                            "algorithmic_bytes_per_launch": algo, "traffic": traffic,
                            # algorithmic bytes exceed what crosses the HBM interface (part of the rows hit the 126 MB L2 even with uniform
                            # ids): DRAM bytes of the ncu capture / live kernel time is the HBM-side utilisation
                            "dram_frac": (traffic / (k_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                            "note": ("algorithmic bytes / time exceeds the HBM copy peak because part of the gathered rows hit the 126 MB L2 "
                                     "even with uniform ids (table 2 x 200 MB); dram_frac is the HBM-side utilisation") if bound == "hbm" else None}}
        rec.update(extra_fields or {})
        out[name] = rec
        del d

    # (a) uniform ids: nothing to reuse in the L2, the gather runs at HBM speed -- the HBM-bound leg of the roofline
    uni = mdata.synth_behaviours(n_news, bhv.n_impressions, mdata.SHAPES[args.workload][2], uniform_ids=True)
    record("uniform_ids", ev, uni, dict(weights=w2, zscore=True, pooled_auc=True), 2, 4, hbm_peak, "hbm",
           traffic=ncu_traffic(args.workload, 2, True, "f32")[0])
    # (b) bf16 tables (fp32 arithmetic): half the row bytes
    ev16 = ScoreEvaluator([t.to(torch.bfloat16) for t in tables], dev)
    record("bf16_tables", ev16, bhv, dict(weights=w2, zscore=True, pooled_auc=True), 2, 2, l2_peak, "l2",
           traffic=ncu_traffic(args.workload, 2, False, "bf16")[0])
    record("bf16_tables_uniform_ids", ev16, uni, dict(weights=w2, zscore=True, pooled_auc=True), 2, 2, hbm_peak, "hbm")
    del ev16, uni
    # (b2) the ensemble epoch as the reference's EnsembleModule.on_test_epoch_end logs it: + Diversity / Personalization @5/10 of
    # category and sentiment (ensemble_module.py:56-84,214-238) -- the full ranking instead of the positives' ranks only
    asp = mdata.synth_aspects(n_news)
    ev_asp = ScoreEvaluator(tables, dev, news_category=asp["category"], news_sentiment=asp["sentiment"])
    record("ensemble_with_aspect_metrics", ev_asp, bhv, dict(weights=w2, zscore=True, pooled_auc=True), 2, tables[0].element_size(), l2_peak, "l2")
    del ev_asp
    # (c) BASELINE.json configs[3]: CR + category + sentiment, 121 weightings re-scored from one gather
    t3 = list(tables) + [mdata.synth_table(n_news, 768, mdata.TABLE_SEEDS[2])]
    ev3 = ScoreEvaluator(t3, dev)
    grid = torch.tensor([[1.0, a / 10.0, b / 10.0] for a in range(11) for b in range(11)], dtype=torch.float32, device=dev)
    record("sweep121_m3", ev3, bhv, dict(weights=grid, zscore=True, pooled_auc=False), 3, 4, l2_peak, "l2",
           {"weightings": 121, "weighting_impressions_per_s": None})
    out["sweep121_m3"]["weighting_impressions_per_s"] = 121 * out["sweep121_m3"]["impressions_per_s"]
    del ev3, t3
    torch.cuda.empty_cache()
    # (d) BASELINE.json configs[4] on one GPU: 37 888 users x 1.25 M news (1/8 of the 10 M catalogue), top-100
    out["retrieval"] = retrieval_extra(dev, flush, steps=3)
    return out


def retrieval_extra(dev, flush, steps: int, n_users: int = 37888, n_shard: int = 1_250_000) -> dict:
    from manner_b200 import retrieval as rt

    dim, k = 768, 100
    g = torch.Generator(device=dev).manual_seed(4321)
    catalog = (torch.randn(n_shard, dim, generator=g, device=dev) * dim ** -0.5).to(torch.bfloat16)
    users = (torch.randn(n_users, dim, generator=g, device=dev) * dim ** -0.5).to(torch.bfloat16)
    r = rt.CatalogRetriever(catalog, k=k)
    for _ in range(3):
        r.local_topk(users)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index or 0)
    events = []
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.local_topk(users)
        e1.record()
        events.append((e0, e1))
    torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in events) / steps
    burst, sustained, kind = FALLBACK_BF16_TFLOPS, FALLBACK_BF16_TFLOPS_SUSTAINED, "fallback"
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            mp = json.load(f)
        burst, sustained, kind = float(mp["bf16_tflops"]), float(mp.get("bf16_tflops_sustained", mp["bf16_tflops"])), "measured"
    tf = 2.0 * n_users * n_shard * dim / (ms * 1e-3) / 1e12
    return {"users_per_s": n_users / (ms * 1e-3), "ms_per_step": ms, "users": n_users, "catalog_rows": n_shard, "k": k, "dtype": "bf16",
            "roofline": {"bound": "tensor", "kernel": "retrieve_topk_kernel", "achieved": tf, "peak": sustained, "unit": "TFLOP/s", "frac": tf / sustained,
                         "peak_kind": kind + " sustained cuBLAS bf16 (a ~55 ms kernel launched back to back under the power cap)",
                         "peak_burst": burst, "frac_of_burst": tf / burst},
            "clocks": clocks}


def run_retrieval_arm(args) -> None:
    """BASELINE.json configs[4] (SURVEY 8(d) mode R): users x catalogue bf16 contraction on tcgen05 with the fused
    per-user top-100, catalogue row-sharded over the ranks (1.25 M rows per GPU = 10 M over 8), NCCL exchange of
    the per-shard top-k + merge kernel.  A step = one block of ``--users`` users against the whole catalogue."""
    import torch.distributed as dist

    from manner_b200 import dist as mdist
    from manner_b200 import ops
    from manner_b200 import retrieval as rt

    os.environ.setdefault("NCCL_DEBUG", "WARN")
    rank, local_rank, world = mdist.init_from_env("nccl")
    dev = torch.device(f"cuda:{local_rank}")
    distributed = world > 1
    dim, k = 768, 100
    n_shard, n_users = args.catalog_per_gpu, args.users
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    catalog = (torch.randn(n_shard, dim, generator=g, device=dev) * dim ** -0.5).to(torch.bfloat16)  # this rank's rows of the catalogue
    gu = torch.Generator().manual_seed(99)
    users_host = (torch.randn(n_users, dim, generator=gu) * dim ** -0.5).to(torch.bfloat16).pin_memory()  # same users on every rank
    users = users_host.to(dev)
    r = rt.CatalogRetriever(catalog, k=k, catalog_id_offset=rank * n_shard, distributed=distributed, exchange=args.exchange, user_block=n_users)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if args.retrieval_diag:
        ops.set_tuning(retrieval_diag=args.retrieval_diag)
    if args.retrieval_pair is not None:
        ops.set_tuning(retrieval_pair=args.retrieval_pair)
    if args.retrieval_window is not None:
        ops.set_tuning(retrieval_window=args.retrieval_window)

    def barrier() -> None:
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(e2e: bool):
        flush.fill_(rank + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        u = users_host.to(dev, non_blocking=True) if e2e else users
        s, i = r.retrieve(u)
        if e2e:
            s, i = s.cpu(), i.cpu()
        e1.record()
        return e0, e1, s, i

    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ops.launch_counts()
    barrier()
    events = [step(False)[:2] for _ in range(args.steps)]
    barrier()
    launches1 = ops.launch_counts()
    # the GEMM + top-k kernel alone (no exchange), for the roofline
    r.local_topk(users)  # first use of this entry point when the exchange is fused into the kernel: keep it out of the timing
    torch.cuda.synchronize(dev)
    k_events = []
    for _ in range(args.steps):
        flush.fill_(rank + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.local_topk(users)
        e1.record()
        k_events.append((e0, e1))
    barrier()
    clocks = sampler.stop() if sampler is not None else None  # sampled during the device-timed regions only
    e2e_events = [step(True)[:2] for _ in range(args.steps + 3)][3:]  # 3 warm-up steps of the end-to-end path
    barrier()

    def max_over_ranks(ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in events))
    kern_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in k_events)) / args.steps
    e2e_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in e2e_events) / len(e2e_events))
    if rank == 0:
        # the kernel runs for ~60 ms per launch, launches back to back, SM clock ~1.4 GHz under sw_power_cap: the regime of
        # the SUSTAINED cuBLAS figure (B200_PROFILING.md); the burst figure is reported beside it
        peak, peak_kind, burst = FALLBACK_BF16_TFLOPS_SUSTAINED, "fallback sustained", FALLBACK_BF16_TFLOPS
        path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(path):
            with open(path) as f:
                mp = json.load(f)
            burst = float(mp["bf16_tflops"])
            peak, peak_kind = float(mp.get("bf16_tflops_sustained", burst)), "measured sustained (long kernel under the power cap; see frac_of_burst)"
        flops = 2.0 * n_users * n_shard * dim  # per launch, per GPU
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_retrieval.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                t = json.load(f)
            if t.get("users") == n_users and t.get("catalog_per_gpu") == n_shard and t.get("dim") == dim:
                traffic = t.get("dram_bytes_per_launch")
        achieved = flops / (kern_ms * 1e-3) / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import manner_oracle as mo

            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            nu, nc = min(n_users, 16384), min(n_shard, 262144)  # ~10 s of host work
            t0 = time.perf_counter()
            sc = mo.retrieval_scores(users_host[:nu], catalog[:nc].cpu())
            torch.topk(sc, k, dim=1)
            dt = time.perf_counter() - t0
            cpu = {"value": nu / dt * (nc / (n_shard * world)), "unit": "users/s", "cores": threads, "kind": "port",
                   "sample": f"{nu} users x {nc} catalogue rows ({dt:.1f} s) fp32 matmul + topk on the host, scaled linearly to the full catalogue"}
        line = {
            "metric": "users/sec (full-catalog retrieval, top-100)", "value": n_users * args.steps / (total_ms * 1e-3), "unit": "users/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"full-catalog retrieval: {n_users} users x {n_shard * world} news ({n_shard} rows per GPU), 768-d bf16, top-100",
                       "exchange": args.exchange if distributed else "none", "l2": "flushed between timed steps (256 MiB write); catalogue shard > L2"},
            "e2e": {"value": n_users / (e2e_ms * 1e-3), "unit": "users/s", "h2d_bytes_per_step": n_users * dim * 2,
                    "d2h_bytes_per_step": (n_users if (not distributed or args.exchange in ("all_gather", "p2p")) else -(-n_users // world)) * k * 12, "ms_per_step": e2e_ms},
            "gpu_launches": launches1[0] - launches0[0],
            "roofline": {"bound": "tensor", "kernel": "retrieve_topk_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_kind": peak_kind, "peak_burst": burst, "frac_of_burst": achieved / burst, "traffic": traffic, "flops_per_launch": flops, "kernel_ms": kern_ms,
                         "kernel_share_of_step": kern_ms * args.steps / total_ms if total_ms > 0 else None},
            "cpu_baseline": cpu, "clocks": clocks,
        }
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
        time.sleep(0.5)
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)  # last line of stdout


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="small", choices=["tiny", "mini", "small", "large"])
    ap.add_argument("--modules", type=int, default=2)
    ap.add_argument("--shard", action="store_true", help="strong scaling: shard one workload over the ranks instead of one workload per rank")
    ap.add_argument("--sweep", type=int, default=0, help="aspect-weight sweep with about this many weightings (configs[3]); no pooled AUC")
    ap.add_argument("--early-fusion", action="store_true", help="CR module with late_fusion=False: additive-attention pooling from cached per-news logits")
    ap.add_argument("--loss", default=None, choices=["ce", "supcon"], help="also compute the reference's test/loss on device")
    ap.add_argument("--table-dtype", default="f32", choices=["f32", "bf16"], help="storage type of the embedding tables (arithmetic stays fp32)")
    ap.add_argument("--uniform-ids", action="store_true", help="draw ids uniformly over the catalogue (no L2-friendly head)")
    ap.add_argument("--cpu-sample", type=int, default=24576, help="impressions of the workload the CPU baseline is timed on (~12 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-checks", action="store_true", help="skip the in-bench parity / shard-invariance assertions (kernel tuning runs)")
    ap.add_argument("--no-extra", dest="extra", action="store_false", help="skip the compact sub-results of the other configurations (uniform ids, bf16, sweep, retrieval)")
    ap.add_argument("--mode", default="eval", choices=["eval", "retrieval"], help="retrieval: BASELINE.json configs[4] (tcgen05 GEMM + fused top-100)")
    ap.add_argument("--users", type=int, default=37888, help="retrieval mode: users per step (37 888 = 2 full waves of 148 CTAs x 128 rows)")
    ap.add_argument("--catalog-per-gpu", type=int, default=1_250_000, help="retrieval mode: catalogue rows per GPU (10 M over 8)")
    ap.add_argument("--exchange", default="all_gather", choices=["all_gather", "all_to_all", "p2p"])
    ap.add_argument("--eval-exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU evaluation: fused stores into the peers' mailboxes over NVLink (default) or the three NCCL collectives")
    ap.add_argument("--retrieval-pair", type=int, default=None, help="retrieval kernel: 1 = CTA pairs (cta_group::2), 0 = one CTA per tile")
    ap.add_argument("--retrieval-window", type=int, default=None, help="retrieval kernel: catalogue tiles a CTA may run ahead of the slowest (0 = unthrottled sweep)")
    ap.add_argument("--retrieval-diag", type=int, default=0, help="DIAGNOSTIC: 1/2 disable parts of the retrieval epilogue (results invalid)")
    ap.add_argument("--upload-segments", type=int, default=5,
                    help="end-to-end pass: segments of the pipelined host -> device upload (1 = copy everything in front of the pass)")
    ap.add_argument("--generic", action="store_true", help="time launch() / upload() + evaluate() instead of prepared passes")
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--chunks-per-warp", type=int, default=None)
    ap.add_argument("--ctas-per-sm", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.mode == "retrieval":
        run_retrieval_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
