"""Multi-GPU plumbing: one process per GPU, impressions sharded, tables replicated.

The reference is single-device (every trainer config has ``devices: 1``, configs/trainer/default.yaml:9)
so there is no reference behaviour to match beyond "same numbers as one GPU".  Impressions are
independent, hence no data-path collective.  What does cross NVLink, once per evaluation:

* one all-reduce (sum) of the fp64 metric sums + impression count + flag bits  -- (W*13 + 5) doubles;
* pooled AUROC only (it is not a sum of per-impression terms): an all-gather of the POSITIVE keys
  (about 4 % of the rows; the negatives never move) and an all-reduce of three int64
  (rank statistic, positives, negatives).

Every function takes the process group, so the same code runs on NCCL (product) and on gloo with CPU
tensors (tests/test_dist_gloo.py, world_size 2).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

from . import _native as nat
from .data import Behaviours, balanced_shard_bounds

N_FLAG_BITS = 4


def shard_for_rank(bhv: Behaviours, rank: int, world_size: int) -> Behaviours:
    """This rank's contiguous impression range, balanced by rows gathered (not by impression count),
    offsets rebased to 0."""
    bounds = balanced_shard_bounds(bhv, world_size)
    return bhv.slice(int(bounds[rank]), int(bounds[rank + 1]))


def pack_metric_payload(sums: Tensor, flags: Tensor, n_impressions: int) -> Tensor:
    """[W, 13] sums, the impression count and the flag word's bits as one fp64 vector (all additive)."""
    bits = ((flags.to(torch.int64).reshape(1) >> torch.arange(N_FLAG_BITS, device=flags.device)) & 1).to(torch.float64)
    count = torch.full((1,), float(n_impressions), dtype=torch.float64, device=sums.device)
    return torch.cat([sums.reshape(-1), count, bits])


def unpack_metric_payload(payload: Tensor, shape: torch.Size) -> Tuple[Tensor, Tensor, int]:
    n = shape.numel()
    sums = payload[:n].reshape(shape)
    bits = (payload[n + 1 : n + 1 + N_FLAG_BITS] > 0).to(torch.int32)
    flags = (bits << torch.arange(N_FLAG_BITS, device=payload.device, dtype=torch.int32)).sum().to(torch.int32).reshape(1)
    return sums, flags, int(round(float(payload[n].item())))


def unpack_metric_payload_device(payload: Tensor, shape: torch.Size) -> Tuple[Tensor, Tensor, Tensor]:
    """Like ``unpack_metric_payload`` but without a host read: the count stays a device scalar."""
    n = shape.numel()
    bits = (payload[n + 1 : n + 1 + N_FLAG_BITS] > 0).to(torch.int32)
    flags = (bits << torch.arange(N_FLAG_BITS, device=payload.device, dtype=torch.int32)).sum().to(torch.int32).reshape(1)
    return payload[:n].reshape(shape), flags, payload[n]


def reduce_metric_sums(sums: Tensor, flags: Tensor, n_impressions: int, group: Optional[dist.ProcessGroup] = None) -> Tuple[Tensor, Tensor, int]:
    """One all-reduce (sum) over the ranks.  Returns (global sums, global flags, global impression count)."""
    payload = pack_metric_payload(sums, flags, n_impressions)
    dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=group)
    return unpack_metric_payload(payload, sums.shape)


def _cuda_build_and_sort(preds: Tensor, labels: Tensor, flags: Tensor):
    from . import ops

    return ops.auc_build_and_sort(preds, labels, 2, flags)


def _cuda_rank_sum(sorted_keys: Tensor, n_pos_local: Tensor, pos_keys: Tensor, n_pos: Tensor, sum2: Tensor) -> None:
    from . import ops

    ops.auc_rank_sum(sorted_keys, n_pos_local, pos_keys, n_pos, sum2)


def pooled_auc_distributed(
    preds: Tensor,
    labels: Tensor,
    flags: Tensor,
    group: Optional[dist.ProcessGroup] = None,
    build_and_sort: Callable = _cuda_build_and_sort,
    rank_sum: Callable = _cuda_rank_sum,
) -> Tensor:
    """Pooled AUROC over all ranks' rows.  ``flags`` must already be the GLOBAL flag word (the sigmoid
    decision of torchmetrics' AUROC is taken over the whole epoch, cr_module.py:273).

    Each rank sorts only its own negatives.  The positives of all ranks are all-gathered; every rank
    counts, for every positive, the negatives below / not above it among ITS negatives; those counts
    are additive over ranks:  auc = sum_ranks sum_pos (lb + ub) / (2 P N).
    Returns fp64 [4] = (auc, P, N, sum2) on the inputs' device."""
    world = dist.get_world_size(group)
    sorted_keys, pos_keys, n_pos = build_and_sort(preds, labels, flags)
    counts = torch.zeros(world, dtype=torch.int64, device=preds.device)
    dist.all_gather_into_tensor(counts, n_pos.reshape(1), group=group)
    counts_h = counts.cpu()
    cap = max(int(counts_h.max().item()), 1)
    mine = torch.zeros(cap, dtype=pos_keys.dtype, device=preds.device)
    k = int(counts_h[dist.get_rank(group)].item())
    mine[:k] = pos_keys[:k]
    gathered = torch.empty(world * cap, dtype=pos_keys.dtype, device=preds.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    sum2 = torch.zeros(1, dtype=torch.int64, device=preds.device)
    for r in range(world):
        if int(counts_h[r].item()) > 0:
            rank_sum(sorted_keys, n_pos, gathered[r * cap : (r + 1) * cap], counts[r : r + 1], sum2)
    stats = torch.stack([sum2.reshape(()), n_pos.reshape(()), torch.tensor(preds.numel(), device=preds.device) - n_pos.reshape(())])
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    s2, p, n = stats[0].double(), stats[1].double(), stats[2].double()
    auc = torch.where((p > 0) & (n > 0), s2 / (2.0 * p * n).clamp_min(1.0), torch.zeros((), dtype=torch.float64, device=preds.device))
    return torch.stack([auc, p, n, s2])


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from torchrun's environment; initialises the default group when
    WORLD_SIZE > 1 and binds this process to its GPU."""
    import os

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        kwargs = {}
        if backend == "nccl" and torch.cuda.is_available():
            kwargs["device_id"] = torch.device(f"cuda:{local_rank}")
        dist.init_process_group(backend=backend, **kwargs)
    return rank, local_rank, world
