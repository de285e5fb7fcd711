"""Multi-GPU plumbing: one process per GPU, impressions sharded, tables replicated.

The reference is single-device (every trainer config has ``devices: 1``, configs/trainer/default.yaml:9)
so there is no reference behaviour to match beyond "same numbers as one GPU".  Impressions are
independent, hence no data-path collective.  What does cross NVLink, once per evaluation:

* one all-reduce (sum) of the fp64 metric sums + impression count + flag bits  -- (W*NUM_METRICS + 5) doubles;
* pooled AUROC only (it is not a sum of per-impression terms): an all-gather of the POSITIVE keys
  (about 4 % of the rows; the negatives never move) and an all-reduce of three int64
  (rank statistic, positives, negatives).

Every function takes the process group, so the same code runs on NCCL (product) and on gloo with CPU
tensors (tests/test_dist_gloo.py, world_size 2).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .data import Behaviours, balanced_shard_bounds

FLAG_BITS = (1, 2, 4, 8, 64)  # flag-word bits a packed payload carries, in order (_native.PAYLOAD_FLAG_BITS; MB200_PAYLOAD_TAIL - 1 of them)
N_FLAG_BITS = len(FLAG_BITS)


def flags_from_payload_tail(tail) -> int:
    """Flag word from the 0 / >0 doubles behind the impression count of a (reduced) packed payload."""
    return sum(bit for bit, v in zip(FLAG_BITS, tail) if v > 0)


def shard_for_rank(bhv: Behaviours, rank: int, world_size: int, align: int = 1) -> Behaviours:
    """This rank's contiguous impression range, balanced by rows gathered (not by impression count),
    offsets rebased to 0.  Pass ``align=step_batch`` when early fusion or a loss is evaluated (see
    ``balanced_shard_bounds``)."""
    bounds = balanced_shard_bounds(bhv, world_size, align)
    return bhv.slice(int(bounds[rank]), int(bounds[rank + 1]))


_ipc_bases: dict = {}  # (IPC handle, local device) -> [base address of the mapping in this process, users]


class _PeerMemory:
    """A peer GPU's memory mapped into this process, presented to torch through the CUDA array interface (zero copy)."""

    def __init__(self, ptr: int, shape, dtype: torch.dtype) -> None:
        typestr = {torch.float32: "<f4", torch.bfloat16: "<u2", torch.float16: "<f2", torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2, "strides": None}


class SharedShards(list):
    """What ``share_table_shards`` returns: the list of all ranks' tensors, plus the peer mappings behind them.  ``close()`` gives
    the mappings back (the last user of a peer allocation unmaps it with mb200_ipc_close); the peer tensors are dead afterwards.
    Not closing is harmless until the process exits -- but a peer that frees and re-creates its buffer may get the same IPC handle
    bytes again, and a mapping kept for the old allocation would then point at stale memory: owners (P2PExchange,
    CatalogRetriever) close what they opened."""

    _keys: list

    def close(self) -> None:
        from . import _native as nat

        keys, self._keys = getattr(self, "_keys", []), []
        for key in keys:
            entry = _ipc_bases.get(key)
            if entry is None:
                continue
            entry[1] -= 1
            if entry[1] <= 0:
                del _ipc_bases[key]
                nat.lib().mb200_ipc_close(entry[0], key[1])  # best effort: the process may already be tearing CUDA down
        del self[:]


def share_table_shards(local_shard: Tensor, group: Optional[dist.ProcessGroup] = None) -> "SharedShards":
    """(Also used for any other buffer the peers' kernels read or write directly, e.g. the retrieval gather buffers.)
    Row-sharded embedding table over the GPUs of one box: every rank contributes the shard it holds ([2**s, dim],
    contiguous, on its GPU) and gets back the list of ALL shards as tensors whose memory its own GPU's kernels can read --
    the peers' shards are mapped through CUDA IPC with peer access (mb200_ipc_export / mb200_ipc_open), so the fused kernel
    loads remote rows directly over NVLink / NVSwitch.  No data moves here; keep ``local_shard`` alive while any rank uses it,
    and ``close()`` the result when the peers' buffers are no longer used."""
    import ctypes
    import os

    from . import _native as nat

    lib = nat.lib()
    if "expandable_segments:true" in os.environ.get("PYTORCH_CUDA_ALLOC_CONF", "").replace(" ", "").lower():
        raise RuntimeError("share_table_shards exports torch allocations with legacy CUDA IPC (cudaIpcGetMemHandle), which cannot export "
                           "expandable segments: unset PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True for processes that share buffers")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if not local_shard.is_cuda or not local_shard.is_contiguous():
        raise ValueError("the local shard must be a contiguous CUDA tensor")
    handle = ctypes.create_string_buffer(64)
    offset = ctypes.c_int64(0)
    nat.check(lib.mb200_ipc_export(local_shard.data_ptr(), handle, ctypes.byref(offset)), "mb200_ipc_export")
    mine = (bytes(handle.raw), int(offset.value), local_shard.device.index, tuple(local_shard.shape), local_shard.dtype)
    gathered: list = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    shards = SharedShards()
    shards._keys = []
    for r, (h, off, peer_dev, shape, dtype) in enumerate(gathered):
        if r == rank:
            shards.append(local_shard)
            continue
        nat.check(lib.mb200_enable_peer_access(local_shard.device.index, peer_dev), "mb200_enable_peer_access")
        # several tensors of a peer can live in one allocation (torch's caching allocator): every allocation is mapped once
        key = (h, local_shard.device.index)
        if key not in _ipc_bases:
            base = ctypes.c_void_p()
            nat.check(lib.mb200_ipc_open(h, 0, local_shard.device.index, ctypes.byref(base)), "mb200_ipc_open")
            _ipc_bases[key] = [int(base.value), 0]
        _ipc_bases[key][1] += 1
        shards._keys.append(key)
        t = torch.as_tensor(_PeerMemory(_ipc_bases[key][0] + off, shape, dtype), device=torch.device("cuda", peer_dev))
        shards.append(t.view(torch.bfloat16) if dtype == torch.bfloat16 else t)
    torch.cuda.synchronize(local_shard.device)
    dist.barrier(group=group)
    return shards


class P2PExchange:
    """The fused exchange of one evaluation (mb200_exchange_post / mb200_exchange_finish): every rank's metric payload and
    positive keys are stored straight into every other rank's mailbox over NVLink peer memory, the AUROC statistics follow
    the same way -- no NCCL call on the evaluation path.  Construction is collective (the mailboxes are mapped into every
    process with CUDA IPC); ``run`` must then be called by all ranks in the same order.  ``n_payload`` doubles per payload,
    ``pos_cap`` = agreed upper bound on any rank's positives (``agree_pos_cap``)."""

    def __init__(self, device: torch.device, n_payload: int, pos_cap: int, group: Optional[dist.ProcessGroup] = None) -> None:
        from . import _native as nat

        lib = nat.lib()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_payload, self.pos_cap = int(n_payload), max(int(pos_cap), 1)
        nbytes = int(lib.mb200_exchange_mailbox_bytes(self.world, self.n_payload, self.pos_cap))
        if nbytes == 0:
            raise ValueError("unsupported exchange shape (1..8 ranks of one NVSwitch box)")
        self.mailbox = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.peers = share_table_shards(self.mailbox, group) if self.world > 1 else [self.mailbox]
        self.workspace: Optional[Tensor] = None  # sigmoid keys of the sorted negatives, grown on demand
        self.epoch = 0
        self._empty_i32 = torch.zeros(2, dtype=torch.int32, device=device)
        self._zero_i64 = torch.zeros(1, dtype=torch.int64, device=device)

    def fits(self, n_payload: int, pos_cap: int) -> bool:
        return int(n_payload) == self.n_payload and int(pos_cap) <= self.pos_cap

    def close(self) -> None:
        """Collective in spirit: call it on every rank once no exchange is in flight; unmaps the peers' mailboxes."""
        if isinstance(self.peers, SharedShards):
            torch.cuda.synchronize(self.mailbox.device)
            self.peers.close()
        self.peers = []

    def run(self, payload: Tensor, outside_index: int, sorted_keys: Optional[Tensor] = None, pos_keys: Optional[Tensor] = None,
            n_pos: Optional[Tensor] = None) -> Tensor:
        """Enqueues both kernels on the current stream.  Returns fp64 [n_payload + 4]: the payload summed over the ranks, then
        (as int64 bit patterns) sum2, P, N of the pooled AUROC and the exchange's flag word."""
        import ctypes

        from . import _native as nat

        lib = nat.lib()
        dev = payload.device
        if payload.dtype != torch.float64 or payload.numel() != self.n_payload or not payload.is_contiguous():
            raise ValueError("payload must be the contiguous fp64 vector this exchange was sized for")
        self.epoch += 1
        out = torch.empty(self.n_payload + 4, dtype=torch.float64, device=dev)
        tail = out[self.n_payload:].view(torch.int64)
        d = nat.ExchangeDesc()
        d.struct_size = ctypes.sizeof(nat.ExchangeDesc)
        d.n_ranks, d.my_rank, d.epoch = self.world, self.rank, self.epoch
        d.n_payload, d.outside_index, d.pos_capacity = self.n_payload, int(outside_index), self.pos_cap
        for r, t in enumerate(self.peers):
            d.mailbox[r] = t.data_ptr()
        d.payload = payload.data_ptr()
        if sorted_keys is not None:
            d.pos_keys, d.n_pos, d.sorted_neg, d.n_rows = pos_keys.data_ptr(), n_pos.data_ptr(), sorted_keys.data_ptr(), sorted_keys.numel()
        else:  # no pooled AUROC wanted: an empty key set
            d.pos_keys, d.n_pos, d.sorted_neg, d.n_rows = self._empty_i32.data_ptr(), self._zero_i64.data_ptr(), self._empty_i32.data_ptr(), 0
        need = int(lib.mb200_exchange_workspace_bytes(d.n_rows))
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        d.out_payload, d.out_stats, d.flags = out.data_ptr(), tail.data_ptr(), tail[3:].data_ptr()
        d.workspace, d.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            nat.check(lib.mb200_exchange_post(ctypes.byref(d), stream), "mb200_exchange_post")
            nat.check(lib.mb200_exchange_finish(ctypes.byref(d), stream), "mb200_exchange_finish")
        return out


def pack_metric_payload(sums: Tensor, flags: Tensor, n_impressions: int) -> Tensor:
    """[W, NUM_METRICS] sums, the impression count and the flag word's bits as one fp64 vector (all additive)."""
    shifts = torch.tensor([b.bit_length() - 1 for b in FLAG_BITS], device=flags.device)
    bits = ((flags.to(torch.int64).reshape(1) >> shifts) & 1).to(torch.float64)
    count = torch.full((1,), float(n_impressions), dtype=torch.float64, device=sums.device)
    return torch.cat([sums.reshape(-1), count, bits])


def unpack_metric_payload(payload: Tensor, shape: torch.Size) -> Tuple[Tensor, Tensor, int]:
    n = shape.numel()
    sums = payload[:n].reshape(shape)
    bits = (payload[n + 1 : n + 1 + N_FLAG_BITS] > 0).to(torch.int32)
    flags = (bits * torch.tensor(FLAG_BITS, device=payload.device, dtype=torch.int32)).sum().to(torch.int32).reshape(1)
    return sums, flags, int(round(float(payload[n].item())))


def unpack_metric_payload_device(payload: Tensor, shape: torch.Size) -> Tuple[Tensor, Tensor, Tensor]:
    """Like ``unpack_metric_payload`` but without a host read: the count stays a device scalar."""
    n = shape.numel()
    bits = (payload[n + 1 : n + 1 + N_FLAG_BITS] > 0).to(torch.int32)
    flags = (bits * torch.tensor(FLAG_BITS, device=payload.device, dtype=torch.int32)).sum().to(torch.int32).reshape(1)
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
    pos_cap: Optional[int] = None,
    build_and_sort: Callable = _cuda_build_and_sort,
    rank_sum: Callable = _cuda_rank_sum,
) -> Tensor:
    """Pooled AUROC over all ranks' rows.  ``flags`` must already be the GLOBAL flag word (the sigmoid
    decision of torchmetrics' AUROC is taken over the whole epoch, cr_module.py:273).

    Each rank sorts only its own negatives.  The positives of all ranks are all-gathered; every rank
    counts, for every positive, the negatives below / not above it among ITS negatives; those counts
    are additive over ranks:  auc = sum_ranks sum_pos (lb + ub) / (2 P N).

    ``pos_cap`` is an upper bound, agreed by all ranks beforehand, on any rank's number of positives
    (``agree_pos_cap``); without it the ranks first exchange their counts, which costs one more collective
    and a host synchronisation.  Returns int64 [3] = (sum2, P, N) on the inputs' device; auc =
    sum2 / (2 P N) (0 when P or N is 0, as torchmetrics returns)."""
    world = dist.get_world_size(group)
    dev = preds.device
    sorted_keys, pos_keys, n_pos = build_and_sort(preds, labels, flags)
    if pos_cap is None:
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, n_pos.reshape(1), group=group)
        pos_cap = int(counts.max().item())
    cap = max(2, (int(pos_cap) + 1) // 2 * 2)  # even: every rank's segment then starts 8-byte aligned
    # segment = [count as int64 (two int32 words)] [cap positive keys]; one all-gather moves both
    seg = torch.empty(cap + 2, dtype=torch.int32, device=dev)
    seg[:2].view(torch.int64).copy_(n_pos)
    k = min(cap, pos_keys.numel())
    seg[2 : 2 + k].copy_(pos_keys[:k])
    flat = torch.empty(world * (cap + 2), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(flat, seg, group=group)
    gathered = flat.view(world, cap + 2)
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    for r in range(world):
        rank_sum(sorted_keys, n_pos, gathered[r, 2:], gathered[r, :2].view(torch.int64), stats[0:1])
    stats[1:2].copy_(n_pos)
    stats[2:3].copy_(preds.numel() - n_pos)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def auc_from_stats(stats: Tensor) -> Tuple[float, int, int]:
    s2, p, n = (int(x) for x in stats.cpu().tolist())
    return (s2 / (2.0 * p * n) if p > 0 and n > 0 else 0.0), p, n


def agree_pos_cap(n_pos_local: int, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> int:
    """One-time agreement (per uploaded shard, outside the hot loop) on the largest per-rank positive count."""
    t = torch.tensor([int(n_pos_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def init_from_env(backend: str = "nccl", always: bool = False) -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from torchrun's environment; initialises the default group when
    WORLD_SIZE > 1 (or ``always``, under torchrun) and binds this process to its GPU."""
    import os

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if (world > 1 or (always and "MASTER_ADDR" in os.environ)) and not dist.is_initialized():
        kwargs = {}
        if backend == "nccl" and torch.cuda.is_available():
            kwargs["device_id"] = torch.device(f"cuda:{local_rank}")
        dist.init_process_group(backend=backend, **kwargs)
    return rank, local_rank, world
