#!/usr/bin/env python
"""Per-stage device times of the fused multi-GPU exchange (run under torchrun, one rank per GPU): key build + sort, post (+ sigmoid keys),
finish, with all ranks entering each repetition together -- what the rendezvous itself costs once nobody has to wait for a slower
kernel.  Prints one JSON line per rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/dist_exchange_timing.py
"""
import ctypes
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from manner_b200 import _native as nat
from manner_b200 import data as mdata
from manner_b200 import dist as mdist
from manner_b200 import ops
from manner_b200.evaluator import ScoreEvaluator


def main() -> None:
    rank, local_rank, world = mdist.init_from_env("nccl", always=True)
    dev = torch.device(f"cuda:{local_rank}")
    tables, bhv = mdata.synth_workload("small", n_modules=2, seed_offset=rank)
    ev = ScoreEvaluator(tables, dev)
    cap = mdist.agree_pos_cap(int(bhv.labels.sum()), dev)
    d = ev.upload(bhv, pos_cap=cap)
    w = torch.tensor([[1.0, 0.4]], dtype=torch.float32, device=dev)
    res = ev.evaluate(d, weights=w, zscore=True, pooled_auc=True, distributed=True, want_scores=True)  # creates the mailboxes
    p2p, scores, labels = ev._p2p, res.scores, d.labels
    payload = torch.zeros(p2p.n_payload, dtype=torch.float64, device=dev)
    payload[15 + 1 + 2] = 1.0  # "a score was outside [0,1]": the sigmoid path, as in the bench
    lib = nat.lib()
    rec = {"build_sort_ms": [], "post_ms": [], "finish_ms": [], "total_ms": []}
    for it in range(25):
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        ev_ = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev_[0].record()
        sorted_keys, pos_keys, n_pos = ops.auc_build_and_sort(scores, labels, 0, None)
        ev_[1].record()
        # the two halves of P2PExchange.run, with an event in between
        p2p.epoch += 1
        out = torch.empty(p2p.n_payload + 4, dtype=torch.float64, device=dev)
        tail = out[p2p.n_payload:].view(torch.int64)
        x = nat.ExchangeDesc()
        x.struct_size = ctypes.sizeof(nat.ExchangeDesc)
        x.n_ranks, x.my_rank, x.epoch = p2p.world, p2p.rank, p2p.epoch
        x.n_payload, x.outside_index, x.pos_capacity = p2p.n_payload, 15 + 1 + 2, p2p.pos_cap
        for r, t in enumerate(p2p.peers):
            x.mailbox[r] = t.data_ptr()
        x.payload, x.pos_keys, x.n_pos, x.sorted_neg, x.n_rows = payload.data_ptr(), pos_keys.data_ptr(), n_pos.data_ptr(), sorted_keys.data_ptr(), sorted_keys.numel()
        need = int(lib.mb200_exchange_workspace_bytes(x.n_rows))
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        x.out_payload, x.out_stats, x.flags = out.data_ptr(), tail.data_ptr(), tail[3:].data_ptr()
        x.workspace, x.workspace_bytes = ws.data_ptr(), need
        stream = torch.cuda.current_stream(dev).cuda_stream
        nat.check(lib.mb200_exchange_post(ctypes.byref(x), stream), "post")
        ev_[2].record()
        nat.check(lib.mb200_exchange_finish(ctypes.byref(x), stream), "finish")
        ev_[3].record()
        torch.cuda.synchronize(dev)
        if it >= 5:
            rec["build_sort_ms"].append(ev_[0].elapsed_time(ev_[1])), rec["post_ms"].append(ev_[1].elapsed_time(ev_[2]))
            rec["finish_ms"].append(ev_[2].elapsed_time(ev_[3])), rec["total_ms"].append(ev_[0].elapsed_time(ev_[3]))
    # the single-GPU chain on the same scores, for comparison
    one = []
    flags = torch.tensor([4], dtype=torch.int32, device=dev)
    for it in range(25):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.ops.manner_b200.pooled_auc(scores, labels, 2, flags)
        e1.record()
        torch.cuda.synchronize(dev)
        if it >= 5:
            one.append(e0.elapsed_time(e1))
    line = {"rank": rank, "world": world} | {k: round(statistics.median(v), 4) for k, v in rec.items()} | {"single_gpu_pooled_auc_ms": round(statistics.median(one), 4)}
    for r in range(world):
        if r == rank:
            print(json.dumps(line), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
