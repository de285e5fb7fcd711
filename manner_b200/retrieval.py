"""Full-catalog retrieval (BASELINE.json configs[4], SURVEY 8(d) mode R): every user's late-fusion vector
against every news row of the catalogue, top-k per user.

The reference has no counterpart -- it scores only the ~37 candidates of an impression
(manner/models/cr_module.py:105-131) -- so the semantics are fixed here and restated by
``oracle/manner_oracle.py:retrieval_topk``:

  * user vector  = mean of the history rows (cr_module.py:116-123, true division), rounded to bf16;
  * score(u, n)  = sum_d bf16(user[u, d]) * bf16(catalog[n, d]) accumulated in fp32 (tcgen05, TMEM);
  * result       = the k best (score, id) per user, score descending, catalogue id ascending on ties;
                   unused slots (k > catalogue size) hold (-inf, -1).

Multi-GPU: the catalogue is row-sharded, every rank scores all users against its
shard, and the per-shard top-k lists meet in one NCCL exchange per user block followed by a k-way
merge kernel:

  * ``exchange="all_gather"``  every rank ends with the merged lists of all users (the north star's wording);
  * ``exchange="all_to_all"``  rank r ends with the merged lists of its 1/R slice of every user block
                               (1/R of the NVLink bytes and of the merge work);
  * ``exchange="p2p"``         the exchange is fused into the kernel: every finished list is stored straight into all
                               GPUs' gather buffers over NVLink / NVSwitch peer memory (CUDA IPC) while the kernel keeps
                               scoring; the only collective left is a 4-byte completion signal.

torch is plumbing (memory, streams, NCCL); the arithmetic is in manner_b200/csrc/retrieval.cu.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native as nat
from .ops import _require_cuda, _workspace

MAX_K = 128


def pool_users(table: Tensor, hist_offsets: Tensor, hist_ids: Tensor, flags: Optional[Tensor] = None) -> Tensor:
    """bf16 [n_users, dim]: mean over each user's history rows of ``table`` (fp32 or bf16), the
    late-fusion user vector of cr_module.py:116-123 (mb200_pool_users)."""
    lib = nat.lib()
    if not table.is_cuda:
        raise RuntimeError("manner_b200: `table` must be a CUDA tensor (there is no CPU path)")
    if table.dtype not in (torch.float32, torch.bfloat16) or table.dim() != 2 or table.stride(1) != 1:
        raise TypeError("table must be [n_news, dim] float32 or bfloat16 with unit inner stride")
    _require_cuda("hist_offsets", hist_offsets, torch.int32)
    _require_cuda("hist_ids", hist_ids, torch.int32)
    n_users = hist_offsets.numel() - 1
    dev = table.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty((n_users, table.shape[1]), dtype=torch.bfloat16, device=dev)
        if n_users == 0:
            return out
        nat.check(
            lib.mb200_pool_users(table.data_ptr(), nat.F32 if table.dtype == torch.float32 else nat.BF16, table.shape[1], table.stride(0),
                                 table.shape[0], hist_offsets.data_ptr(), hist_ids.data_ptr(), n_users, out.data_ptr(),
                                 None if flags is None else flags.data_ptr(), stream),
            "mb200_pool_users",
        )
    return out


@torch.library.custom_op("manner_b200::retrieve_topk", mutates_args=())
def retrieve_topk(users: Tensor, catalog: Tensor, k: int, catalog_id_offset: int = 0, want_scores_matrix: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """(scores fp32 [U, k], ids int64 [U, k], full score matrix fp32 [U, N] or empty) -- mb200_retrieve_topk.
    ``users`` and ``catalog`` are bf16 [*, dim] on the same device, dim a multiple of 64."""
    lib = nat.lib()
    _require_cuda("users", users, torch.bfloat16)
    _require_cuda("catalog", catalog, torch.bfloat16)
    if users.dim() != 2 or catalog.dim() != 2 or users.shape[1] != catalog.shape[1]:
        raise ValueError("users [U, dim] and catalog [N, dim] expected")
    if users.device != catalog.device:
        raise ValueError("users and catalog must live on the same device")
    dev = users.device
    n_users, n_catalog = users.shape[0], catalog.shape[0]
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out_s = torch.empty((n_users, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((n_users, k), dtype=torch.int64, device=dev)
        full = torch.empty((n_users, n_catalog) if want_scores_matrix else (0,), dtype=torch.float32, device=dev)
        d = nat.RetrievalDesc()
        d.struct_size = ctypes.sizeof(nat.RetrievalDesc)
        d.dim, d.k = users.shape[1], k
        d.n_users, d.n_catalog, d.catalog_id_offset = n_users, n_catalog, catalog_id_offset
        d.users, d.catalog = users.data_ptr(), catalog.data_ptr()
        d.out_scores, d.out_ids = out_s.data_ptr(), out_i.data_ptr()
        d.debug_scores = full.data_ptr() if want_scores_matrix else None
        need = lib.mb200_retrieval_workspace_bytes(ctypes.byref(d))
        ws = _workspace(dev, stream, "retrieval", max(need, 256))
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        nat.check(lib.mb200_retrieve_topk(ctypes.byref(d), stream), "mb200_retrieve_topk")
    return out_s, out_i, full


@retrieve_topk.register_fake
def _(users, catalog, k, catalog_id_offset=0, want_scores_matrix=False):
    dev = users.device
    return (
        torch.empty((users.shape[0], k), dtype=torch.float32, device=dev),
        torch.empty((users.shape[0], k), dtype=torch.int64, device=dev),
        torch.empty((users.shape[0], catalog.shape[0]) if want_scores_matrix else (0,), dtype=torch.float32, device=dev),
    )


def retrieve_topk_p2p(users: Tensor, catalog: Tensor, k: int, catalog_id_offset: int, gather_scores: Sequence[Tensor],
                      gather_ids: Sequence[Tensor], my_rank: int) -> None:
    """mb200_retrieve_topk with the fused exchange: ``gather_scores[r]`` / ``gather_ids[r]`` are GPU r's gather buffers
    [R, rows, k] (fp32 / int64) as this process sees them (``dist.share_table_shards``); the kernel writes this rank's lists
    into slot ``my_rank`` of ALL of them -- its own through ordinary stores, the peers' over NVLink -- while it is still
    scoring.  Nothing is returned: after a completion signal on the stream the caller merges its own buffer."""
    lib = nat.lib()
    _require_cuda("users", users, torch.bfloat16)
    _require_cuda("catalog", catalog, torch.bfloat16)
    world = len(gather_scores)
    mine_s, mine_i = gather_scores[my_rank], gather_ids[my_rank]
    rows = mine_s.shape[1]
    if mine_s.shape != (world, rows, k) or mine_i.shape != (world, rows, k) or users.shape[0] > rows:
        raise ValueError("gather buffers must be [world, rows >= n_users, k]")
    dev = users.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        d = nat.RetrievalDesc()
        d.struct_size = ctypes.sizeof(nat.RetrievalDesc)
        d.dim, d.k = users.shape[1], k
        d.n_users, d.n_catalog, d.catalog_id_offset = users.shape[0], catalog.shape[0], catalog_id_offset
        d.users, d.catalog = users.data_ptr(), catalog.data_ptr()
        d.out_scores, d.out_ids = mine_s[my_rank].data_ptr(), mine_i[my_rank].data_ptr()
        d.n_peers, d.my_rank, d.peer_rows = world, my_rank, rows
        for r in range(world):
            d.peer_scores[r], d.peer_ids[r] = gather_scores[r].data_ptr(), gather_ids[r].data_ptr()
        need = lib.mb200_retrieval_workspace_bytes(ctypes.byref(d))
        ws = _workspace(dev, stream, "retrieval", max(need, 256))
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        nat.check(lib.mb200_retrieve_topk(ctypes.byref(d), stream), "mb200_retrieve_topk")


def merge_topk(scores: Tensor, ids: Tensor) -> Tuple[Tensor, Tensor]:
    """Merges per-shard sorted lists [R, U, k] into the global top-k [U, k] (mb200_merge_topk)."""
    lib = nat.lib()
    _require_cuda("scores", scores, torch.float32)
    _require_cuda("ids", ids, torch.int64)
    if scores.dim() != 3 or scores.shape != ids.shape:
        raise ValueError("scores and ids must both be [shards, n_users, k]")
    shards, n_users, k = scores.shape
    dev = scores.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out_s = torch.empty((n_users, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((n_users, k), dtype=torch.int64, device=dev)
        if n_users:
            nat.check(lib.mb200_merge_topk(scores.data_ptr(), ids.data_ptr(), shards, n_users, k, out_s.data_ptr(), out_i.data_ptr(), stream),
                      "mb200_merge_topk")
    return out_s, out_i


def catalog_shard_bounds(n_catalog: int, world_size: int, align: int = 256) -> list:
    """Row ranges [lo, hi) of the catalogue per rank: equal multiples of the kernel's 256-row tile, the
    remainder on the last ranks that still have rows (host logic, covered by the gloo test)."""
    tiles = (n_catalog + align - 1) // align
    base, extra = divmod(tiles, world_size)
    bounds, lo = [], 0
    for r in range(world_size):
        hi = min(n_catalog, lo + (base + (1 if r < extra else 0)) * align)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def exchange_topk(s: Tensor, i: Tensor, group: Optional["torch.distributed.ProcessGroup"] = None, exchange: str = "all_gather",
                  merge: Callable[[Tensor, Tensor], Tuple[Tensor, Tensor]] = merge_topk) -> Tuple[Tensor, Tensor]:
    """The one collective of retrieval mode: this rank's per-shard lists ``s``/``i`` [n, k] (global ids) meet
    the other ranks' and are merged.  ``all_gather``: returns the merged lists of all n users on every rank;
    ``all_to_all``: returns those of this rank's slice [rank * ceil(n / R), ...) of the n users.  ``merge`` is
    the CUDA merge kernel; the gloo test passes a numpy stand-in with the same contract."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, k = s.shape
    if exchange == "all_gather":
        gs = torch.empty((world * n, k), dtype=s.dtype, device=s.device)
        gi = torch.empty((world * n, k), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(gs, s.contiguous(), group=group)
        dist.all_gather_into_tensor(gi, i.contiguous(), group=group)
        gs, gi = gs.view(world, n, k), gi.view(world, n, k)
    else:
        per = (n + world - 1) // world  # pad to a multiple of the world size so every rank owns an equal slice
        if per * world != n:
            pad = per * world - n
            s = torch.cat([s, torch.full((pad, k), float("-inf"), dtype=s.dtype, device=s.device)])
            i = torch.cat([i, torch.full((pad, k), -1, dtype=i.dtype, device=i.device)])
        gs, gi = torch.empty_like(s), torch.empty_like(i)
        dist.all_to_all_single(gs, s.contiguous(), group=group)
        dist.all_to_all_single(gi, i.contiguous(), group=group)
        keep = max(0, min(per, n - rank * per))
        gs, gi = gs.view(world, per, k)[:, :keep].contiguous(), gi.view(world, per, k)[:, :keep].contiguous()
    return merge(gs, gi)


class CatalogRetriever:
    """One rank's shard of the bf16 catalogue on its GPU + the top-k exchange.

    ``catalog`` is this rank's rows [lo, hi) of the global catalogue (``catalog_id_offset = lo``); with
    ``group=None`` and world size 1 it is the whole catalogue."""

    def __init__(self, catalog: Tensor, k: int = 100, catalog_id_offset: int = 0, distributed: bool = False,
                 group: Optional["torch.distributed.ProcessGroup"] = None, exchange: str = "all_gather", user_block: int = 65536) -> None:
        nat.lib()
        if not 1 <= k <= MAX_K:
            raise ValueError(f"k must be in 1..{MAX_K}")
        if exchange not in ("all_gather", "all_to_all", "p2p"):
            raise ValueError("exchange must be 'all_gather', 'all_to_all' or 'p2p'")
        _require_cuda("catalog", catalog, torch.bfloat16)
        self.catalog, self.k, self.offset = catalog, int(k), int(catalog_id_offset)
        self.distributed, self.group, self.exchange = bool(distributed), group, exchange
        self.user_block = int(user_block)
        self._gather = None  # exchange == "p2p": two sets of gather buffers shared over CUDA IPC (double buffered over user blocks)
        self._blk = 0  # blocks exchanged so far, over ALL retrieve() calls: consecutive blocks must alternate buffers across calls too

    def _p2p_buffers(self):
        import torch.distributed as dist

        from . import dist as mdist

        if self._gather is None:
            world = dist.get_world_size(self.group)
            dev = self.catalog.device
            self._gather = []
            for _ in range(2):
                s = torch.full((world, self.user_block, self.k), float("-inf"), dtype=torch.float32, device=dev)
                i = torch.full((world, self.user_block, self.k), -1, dtype=torch.int64, device=dev)
                self._gather.append((mdist.share_table_shards(s, self.group), mdist.share_table_shards(i, self.group)))
            self._token = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._gather

    def close(self) -> None:
        """Unmaps the peers' gather buffers of the fused exchange (call on every rank once no retrieve() is in flight)."""
        if self._gather is not None:
            torch.cuda.synchronize(self.catalog.device)
            for gs, gi in self._gather:
                gs.close(), gi.close()
            self._gather = None

    def local_topk(self, users: Tensor) -> Tuple[Tensor, Tensor]:
        """Top-k of ``users`` against this rank's shard only (global ids)."""
        s, i, _ = torch.ops.manner_b200.retrieve_topk(users, self.catalog, self.k, self.offset, False)
        return s, i

    def retrieve(self, users: Tensor) -> Tuple[Tensor, Tensor]:
        """Global top-k.  Single rank: one kernel.  Distributed: per block of ``user_block`` users, local
        top-k -> NCCL exchange -> merge kernel; with ``all_to_all`` the result covers this rank's slice
        ``user_slice(n_users)`` of every block, concatenated in user order."""
        if not self.distributed:
            return self.local_topk(users)
        out_s, out_i = [], []
        if self.exchange == "p2p":
            # fused exchange: the kernel stores every finished list into all GPUs' gather buffers over NVLink while it is
            # still scoring; the only collective is a 4-byte all-reduce that tells every rank the others' kernels are done
            import torch.distributed as dist

            rank = dist.get_rank(self.group)
            bufs = self._p2p_buffers()
            for lo in range(0, users.shape[0], self.user_block):
                blk = users[lo : lo + self.user_block]
                # the buffer index persists across calls: a peer may still be merging the previous call's LAST block (its token
                # all-reduce only proves it enqueued that merge), so this call's first block must go to the other buffer; the
                # buffer written two blocks ago is safe because the all-reduce in between is ordered after that merge on every rank
                gs, gi = bufs[self._blk % 2]
                self._blk += 1
                retrieve_topk_p2p(blk, self.catalog, self.k, self.offset, gs, gi, rank)
                dist.all_reduce(self._token, group=self.group)
                n = blk.shape[0]
                ms, mi = merge_topk(gs[rank][:, :n].contiguous(), gi[rank][:, :n].contiguous())
                out_s.append(ms), out_i.append(mi)
            return torch.cat(out_s), torch.cat(out_i)
        for lo in range(0, users.shape[0], self.user_block):
            s, i = self.local_topk(users[lo : lo + self.user_block])
            ms, mi = exchange_topk(s, i, self.group, self.exchange)
            out_s.append(ms), out_i.append(mi)
        return torch.cat(out_s), torch.cat(out_i)

    @staticmethod
    def user_slice(n_users: int, user_block: int, rank: int, world: int) -> list:
        """Global user indices a rank ends up with under ``exchange='all_to_all'`` (host logic)."""
        idx = []
        for lo in range(0, n_users, user_block):
            n = min(user_block, n_users - lo)
            per = (n + world - 1) // world
            a, b = min(n, rank * per), min(n, (rank + 1) * per)
            idx.extend(range(lo + a, lo + b))
        return idx
