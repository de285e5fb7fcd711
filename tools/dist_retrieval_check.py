#!/usr/bin/env python
"""Multi-GPU check of retrieval mode (run under torchrun, one rank per GPU):
the row-sharded catalogue + NCCL top-k exchange + merge kernel must return exactly what ONE GPU returns on the whole
catalogue -- ids bit-identical, scores bit-identical (each score is computed from the same bf16 rows by the same kernel).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_retrieval_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from manner_b200 import dist as mdist
from manner_b200 import retrieval as rt


def main() -> None:
    rank, local_rank, world = mdist.init_from_env("nccl")
    dev = torch.device(f"cuda:{local_rank}")
    g = torch.Generator().manual_seed(77)
    n_users, n_catalog, dim, k = 5000, 200_003, 768, 100
    users = (torch.randn(n_users, dim, generator=g) * dim ** -0.5).to(torch.bfloat16).to(dev)
    catalog = (torch.randn(n_catalog, dim, generator=g) * dim ** -0.5).to(torch.bfloat16)
    catalog[150_000:150_500] = catalog[:500]  # cross-shard ties
    lo, hi = rt.catalog_shard_bounds(n_catalog, world)[rank]
    shard = catalog[lo:hi].to(dev).contiguous()
    whole = rt.CatalogRetriever(catalog.to(dev), k=k)
    ref_s, ref_i = whole.retrieve(users)
    out = {"world": world, "n_users": n_users, "n_catalog": n_catalog}
    for exchange in ("all_gather", "all_to_all", "p2p"):
        r = rt.CatalogRetriever(shard, k=k, catalog_id_offset=lo, distributed=True, exchange=exchange, user_block=2048)
        s, i = r.retrieve(users)
        if exchange in ("all_gather", "p2p"):
            ok = bool(torch.equal(i, ref_i) and torch.equal(s, ref_s))
        else:
            idx = torch.tensor(rt.CatalogRetriever.user_slice(n_users, 2048, rank, world), device=dev)
            ok = bool(torch.equal(i, ref_i[idx]) and torch.equal(s, ref_s[idx]))
        if exchange == "p2p":
            # consecutive single-block calls with DIFFERENT users (the double buffer must alternate across calls, or a
            # fast rank's next call overwrites the gather buffer a slow rank is still merging)
            single = rt.CatalogRetriever(shard, k=k, catalog_id_offset=lo, distributed=True, exchange="p2p", user_block=2048)
            for rep in range(6):
                sl = slice(rep * 700, rep * 700 + 1500 + 37 * rep)
                s2, i2 = single.retrieve(users[sl])
                ok = ok and bool(torch.equal(i2, ref_i[sl]) and torch.equal(s2, ref_s[sl]))
        t = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out[exchange] = bool(t.item())
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    if not (out["all_gather"] and out["all_to_all"] and out["p2p"]):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
