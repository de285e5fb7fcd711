"""torch custom ops over the C ABI: ``torch.ops.manner_b200.score_eval`` and
``torch.ops.manner_b200.pooled_auc``.

torch is plumbing here -- device memory, the current stream, output allocation.  The arithmetic is
in libmanner_b200.so (manner_b200/csrc/*.cu), reached through ctypes with raw pointers.  There is no
fallback: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native as nat

_workspaces: Dict[Tuple[int, int, str], Tensor] = {}


def _workspace(device: torch.device, stream: int, kind: str, nbytes: int) -> Tensor:
    """Per-(device, stream) scratch cache, grown on demand; 256-byte aligned by the caching allocator."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream, kind)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(name: str, t: Tensor, dtype: torch.dtype) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"manner_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"manner_b200: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"manner_b200: `{name}` must be contiguous")


@torch.library.custom_op("manner_b200::score_eval", mutates_args=())
def score_eval(
    tables: Sequence[Tensor],
    hist_offsets: Tensor,
    hist_ids: Tensor,
    cand_offsets: Tensor,
    cand_ids: Tensor,
    labels: Tensor,
    weights: Optional[Tensor],
    zscore: bool,
    max_cand: int,
    active_mask: int,
    k0: int,
    k1: int,
    want_scores: bool,
    scores_weighting: int,
    want_per_impression: bool,
    news_category: Optional[Tensor],
    news_sentiment: Optional[Tensor],
    num_categ_classes: int,
    num_sent_classes: int,
    attn_logits: Sequence[Optional[Tensor]],
    pack_payload: bool = False,
    hist_pad: Optional[Tensor] = None,
    loss_kind: int = 0,
    loss_temperature: float = 1.0,
    cand_pad: Optional[Tensor] = None,
    n_table_shards: int = 1,
    table_shard_shift: int = 0,
    n_news: int = 0,
    ready: Optional[Tensor] = None,
    ready_segments: int = 0,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Fused gather / pool / score / z-score / ensemble / per-impression metrics (include/manner_b200.h,
    mb200_score_eval).  Returns (scores fp32 [sum C] or empty, per_impression fp32 [W, B, NUM_METRICS] or empty,
    sums fp64 [W, NUM_METRICS], flags int32 [1], loss_per_impression fp32 [B] or empty).  With ``pack_payload``
    the sums come back flat, fp64 [W*NUM_METRICS + 5], with the impression count and the flag bits appended: the
    buffer a multi-GPU caller all-reduces.

    Early fusion (cr_module.py:124-125): ``attn_logits`` is ``[]`` (late fusion everywhere) or one entry per table: ``attn_logits[m]`` = the [n_news + 1] logits of ``attention_logits`` for
    module m (None keeps module m on late fusion) and ``hist_pad`` [B] int32 = zero rows the reference's
    step batch pads impression i's history with.  ``loss_kind`` (nat.LOSS_CE / LOSS_SUPCON) adds the per-impression
    loss of cr_module.py:140-171 (``cand_pad`` [B]: padded candidate columns, cross entropy only).

    Row-sharded tables (``n_table_shards`` = R > 1): ``tables`` then holds R tensors per module, module-major, each
    [1 << table_shard_shift, dim] -- this GPU's shard and the peers' shards opened over CUDA IPC (dist.share_table_shards);
    ``n_news`` is the size of the whole catalogue.  The kernel reads remote rows directly over NVLink.

    Pipelined upload (ScoreEvaluator.upload(pipelined=True) -> mb200_upload_begin): ``ready`` int32 [1] is the device word the
    copy stream raises to the number of leading impressions whose ids / labels have arrived, ``ready_segments`` the number of
    segments; the kernel starts on the first segment while the others are still being copied."""
    lib = nat.lib()
    n_modules = len(tables) // max(n_table_shards, 1)
    if n_table_shards > 1 and (n_table_shards > nat.MAX_TABLE_SHARDS or n_modules * n_table_shards != len(tables) or n_news <= 0):
        raise ValueError("row-sharded tables: pass n_modules x n_table_shards tensors (module-major) and the catalogue size n_news")
    if n_modules < 1 or n_modules > nat.MAX_MODULES:
        raise ValueError(f"1..{nat.MAX_MODULES} embedding tables expected")
    t0 = tables[0]
    if t0.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("embedding tables must be float32 or bfloat16")
    for m, t in enumerate(tables):
        if not t.is_cuda:
            raise RuntimeError("manner_b200: embedding tables must be CUDA tensors (there is no CPU path)")
        if t.dtype != t0.dtype or t.shape != t0.shape or t.dim() != 2 or t.stride(1) != 1 or t.stride(0) != t0.stride(0):
            raise ValueError(f"table {m}: all tables must share dtype, shape [n_news, dim] and row stride")
    dev = hist_offsets.device if n_table_shards > 1 else t0.device  # sharded: tables[0] may be a peer GPU's shard
    for name, t, dt in (
        ("hist_offsets", hist_offsets, torch.int32), ("hist_ids", hist_ids, torch.int32),
        ("cand_offsets", cand_offsets, torch.int32), ("cand_ids", cand_ids, torch.int32), ("labels", labels, torch.uint8),
    ):
        _require_cuda(name, t, dt)
    n_impr = hist_offsets.numel() - 1
    if cand_offsets.numel() != n_impr + 1 or labels.numel() != cand_ids.numel():
        raise ValueError("inconsistent CSR arrays")
    n_w = 1
    if weights is not None:
        _require_cuda("weights", weights, torch.float32)
        if weights.dim() != 2 or weights.shape[1] != n_modules:
            raise ValueError("weights must be [n_weightings, n_modules]")
        n_w = weights.shape[0]
    if (news_category is None) != (news_sentiment is None):
        raise ValueError("news_category and news_sentiment must be given together")
    if news_category is not None:
        _require_cuda("news_category", news_category, torch.int32)
        _require_cuda("news_sentiment", news_sentiment, torch.int32)
    if len(attn_logits):
        if len(attn_logits) != n_modules:
            raise ValueError("attn_logits needs one entry per table (None = late fusion for that module)")
        for a in attn_logits:
            if a is not None:
                _require_cuda("attn_logits", a, torch.float32)
                if a.numel() != t0.shape[0] + 1:
                    raise ValueError("attn_logits[m] must hold n_news + 1 floats (ops.attention_logits)")
    for name, t in (("hist_pad", hist_pad), ("cand_pad", cand_pad)):
        if t is not None:
            _require_cuda(name, t, torch.int32)
            if t.numel() != n_impr:
                raise ValueError(f"{name} must have one entry per impression")

    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        scores = torch.empty(cand_ids.numel() if want_scores else 0, dtype=torch.float32, device=dev)
        per_impr = torch.empty((n_w, n_impr, nat.NUM_METRICS) if want_per_impression else (0,), dtype=torch.float32, device=dev)
        sums = torch.empty(n_w * nat.NUM_METRICS + nat.PAYLOAD_TAIL if pack_payload else (n_w, nat.NUM_METRICS), dtype=torch.float64, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        loss_out = torch.empty(n_impr if loss_kind else 0, dtype=torch.float32, device=dev)

        d = nat.EvalDesc()
        for m, a in enumerate(attn_logits):
            d.attn_logits[m] = _ptr(a)
        d.hist_pad, d.cand_pad = _ptr(hist_pad), _ptr(cand_pad)
        d.loss_kind, d.loss_temperature = int(loss_kind), float(loss_temperature)
        d.loss_per_impression = loss_out.data_ptr() if loss_kind else None
        d.pack_payload = int(pack_payload)
        d.struct_size = ctypes.sizeof(nat.EvalDesc)
        d.n_modules = n_modules
        d.dtype = nat.F32 if t0.dtype == torch.float32 else nat.BF16
        d.dim = t0.shape[1]
        d.active_modules_mask = active_mask
        d.row_stride = t0.stride(0)
        if n_table_shards > 1:
            if t0.shape[0] != (1 << table_shard_shift):
                raise ValueError("every table shard must hold 1 << table_shard_shift rows")
            d.n_news, d.n_table_shards, d.table_shard_shift = n_news, n_table_shards, table_shard_shift
            for m in range(n_modules):
                for sh in range(n_table_shards):
                    d.table_shards[m][sh] = tables[m * n_table_shards + sh].data_ptr()
        else:
            d.n_news = t0.shape[0]
            for m, t in enumerate(tables):
                d.tables[m] = t.data_ptr()
        d.n_impressions = n_impr
        d.hist_offsets, d.hist_ids = hist_offsets.data_ptr(), hist_ids.data_ptr()
        d.cand_offsets, d.cand_ids, d.labels = cand_offsets.data_ptr(), cand_ids.data_ptr(), labels.data_ptr()
        d.max_cand = max_cand
        d.zscore = int(zscore)
        d.n_weightings = n_w
        d.weights = _ptr(weights)
        d.k0, d.k1 = k0, k1
        d.news_category, d.news_sentiment = _ptr(news_category), _ptr(news_sentiment)
        d.num_categ_classes, d.num_sent_classes = num_categ_classes, num_sent_classes
        d.scores = scores.data_ptr() if want_scores else None
        d.scores_weighting = scores_weighting
        d.per_impression = per_impr.data_ptr() if want_per_impression else None
        d.sums = sums.data_ptr()
        d.flags = flags.data_ptr()
        if ready is not None:
            _require_cuda("ready", ready, torch.int32)
            d.ready, d.ready_segments = ready.data_ptr(), int(ready_segments)
        need = lib.mb200_eval_workspace_bytes(ctypes.byref(d))
        if need == 0:
            # the size query runs the same validation as the call: report the precise status
            d.workspace, d.workspace_bytes = None, 0
            nat.check(lib.mb200_score_eval(ctypes.byref(d), stream), "mb200_score_eval")
            raise nat.NativeError("mb200_eval_workspace_bytes returned 0")
        ws = _workspace(dev, stream, "eval", need)
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        nat.check(lib.mb200_score_eval(ctypes.byref(d), stream), "mb200_score_eval")
    return scores, per_impr, sums, flags, loss_out


@score_eval.register_fake
def _(tables, hist_offsets, hist_ids, cand_offsets, cand_ids, labels, weights, zscore, max_cand, active_mask, k0, k1,
      want_scores, scores_weighting, want_per_impression, news_category, news_sentiment, num_categ_classes, num_sent_classes,
      attn_logits, pack_payload=False, hist_pad=None, loss_kind=0, loss_temperature=1.0, cand_pad=None, n_table_shards=1,
      table_shard_shift=0, n_news=0, ready=None, ready_segments=0):
    dev = tables[0].device
    n_w = 1 if weights is None else weights.shape[0]
    n_impr = hist_offsets.numel() - 1
    return (
        torch.empty(cand_ids.numel() if want_scores else 0, dtype=torch.float32, device=dev),
        torch.empty((n_w, n_impr, nat.NUM_METRICS) if want_per_impression else (0,), dtype=torch.float32, device=dev),
        torch.empty(n_w * nat.NUM_METRICS + nat.PAYLOAD_TAIL if pack_payload else (n_w, nat.NUM_METRICS), dtype=torch.float64, device=dev),
        torch.empty(1, dtype=torch.int32, device=dev),
        torch.empty(n_impr if loss_kind else 0, dtype=torch.float32, device=dev),
    )


def attention_logits(table: Tensor, weight: Tensor, bias: Tensor, query: Tensor) -> Tensor:
    """fp32 [n_rows + 1]: query . tanh(weight x_n + bias) for every row of ``table`` and, last, for an all-zero row
    (mb200_attention_logits) -- the per-news part of NAMLUserEncoder's additive attention (attention.py:20-24)."""
    lib = nat.lib()
    if not table.is_cuda:
        raise RuntimeError("manner_b200: `table` must be a CUDA tensor (there is no CPU path)")
    if table.dtype not in (torch.float32, torch.bfloat16) or table.dim() != 2 or table.stride(1) != 1:
        raise TypeError("table must be [n_rows, dim] float32 or bfloat16 with unit inner stride")
    dev = table.device
    w = weight.detach().to(dev, torch.float32).contiguous()
    b = bias.detach().to(dev, torch.float32).contiguous()
    q = query.detach().to(dev, torch.float32).contiguous()
    if w.dim() != 2 or w.shape[1] != table.shape[1] or b.numel() != w.shape[0] or q.numel() != w.shape[0]:
        raise ValueError("weight [Q, dim], bias [Q], query [Q] expected")
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(table.shape[0] + 1, dtype=torch.float32, device=dev)
        nat.check(
            lib.mb200_attention_logits(table.data_ptr(), nat.F32 if table.dtype == torch.float32 else nat.BF16, table.shape[1], table.stride(0),
                                       table.shape[0], w.data_ptr(), b.data_ptr(), q.data_ptr(), w.shape[0], out.data_ptr(), stream),
            "mb200_attention_logits",
        )
    return out


def step_loss(loss_per_impression: Tensor, step: int, loss_kind: int, cand_offsets: Optional[Tensor] = None, labels: Optional[Tensor] = None) -> Tensor:
    """fp64 [2] = (sum over the reference's steps of the step loss, number of steps): what MeanMetric averages into
    test/loss (cr_module.py:253-259); mb200_step_loss.  ``cand_offsets`` / ``labels`` enable SupCon's step-level guards
    (components/losses.py:15-16,22)."""
    lib = nat.lib()
    _require_cuda("loss_per_impression", loss_per_impression, torch.float32)
    if (cand_offsets is None) != (labels is None):
        raise ValueError("give cand_offsets and labels together")
    if cand_offsets is not None:
        _require_cuda("cand_offsets", cand_offsets, torch.int32)
        _require_cuda("labels", labels, torch.uint8)
    dev = loss_per_impression.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(2, dtype=torch.float64, device=dev)
        nat.check(lib.mb200_step_loss(loss_per_impression.data_ptr(), loss_per_impression.numel(), int(step), int(loss_kind),
                                      _ptr(cand_offsets), _ptr(labels), out.data_ptr(), stream), "mb200_step_loss")
    return out


@torch.library.custom_op("manner_b200::pooled_auc", mutates_args=())
def pooled_auc(preds: Tensor, labels: Tensor, sigmoid_mode: int, flags: Optional[Tensor]) -> Tensor:
    """Pooled AUROC of torchmetrics' ``AUROC(task="binary")`` (cr_module.py:81,273) on one device.
    Returns fp64 [4] = (auc, P, N, sum2).  sigmoid_mode 0 never / 1 always / 2 from ``flags``."""
    lib = nat.lib()
    _require_cuda("preds", preds, torch.float32)
    _require_cuda("labels", labels, torch.uint8)
    if flags is not None:
        _require_cuda("flags", flags, torch.int32)
    dev = preds.device
    n = preds.numel()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(4, dtype=torch.float64, device=dev)
        ws = _workspace(dev, stream, "auc", lib.mb200_pooled_auc_workspace_bytes(n))
        nat.check(
            lib.mb200_pooled_auc(preds.data_ptr(), labels.data_ptr(), n, sigmoid_mode, _ptr(flags), ws.data_ptr(), ws.numel(),
                                 out.data_ptr(), stream),
            "mb200_pooled_auc",
        )
    return out


@pooled_auc.register_fake
def _(preds, labels, sigmoid_mode, flags):
    return torch.empty(4, dtype=torch.float64, device=preds.device)


@torch.library.custom_op("manner_b200::pooled_auc_bounded", mutates_args=())
def pooled_auc_bounded(preds: Tensor, labels: Tensor, sigmoid_mode: int, flags: Optional[Tensor], pos_capacity: int) -> Tensor:
    """``pooled_auc`` for callers that know an upper bound on the number of positives (``pos_capacity``; e.g. counted on the host
    labels before the upload): only the positives are sorted, the negatives are ranked against them in one streaming pass
    (mb200_pooled_auc_bounded).  Same fp64 [4]; auc is NaN if there were more positives than ``pos_capacity``.  When the
    positives are not few (2 * pos_capacity > n) the full sort of ``pooled_auc`` is used."""
    n = preds.numel()
    if 2 * int(pos_capacity) > n:
        return pooled_auc(preds, labels, sigmoid_mode, flags)
    lib = nat.lib()
    _require_cuda("preds", preds, torch.float32)
    _require_cuda("labels", labels, torch.uint8)
    if flags is not None:
        _require_cuda("flags", flags, torch.int32)
    dev = preds.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        out = torch.empty(4, dtype=torch.float64, device=dev)
        ws = _workspace(dev, stream, "auc_bounded", lib.mb200_pooled_auc_bounded_workspace_bytes(n, int(pos_capacity)))
        nat.check(
            lib.mb200_pooled_auc_bounded(preds.data_ptr(), labels.data_ptr(), n, int(pos_capacity), sigmoid_mode, _ptr(flags), ws.data_ptr(), ws.numel(),
                                         out.data_ptr(), stream),
            "mb200_pooled_auc_bounded",
        )
    return out


@pooled_auc_bounded.register_fake
def _(preds, labels, sigmoid_mode, flags, pos_capacity):
    return torch.empty(4, dtype=torch.float64, device=preds.device)


# ---- staged pooled AUC (multi-GPU: positives are exchanged between stage 2 and 3; see dist.py) ----------


def auc_build_and_sort(preds: Tensor, labels: Tensor, sigmoid_mode: int, flags: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """Stages 1+2 on this rank's rows: returns (sorted keys uint32-as-int32 [n], positive keys [n] with the
    first n_pos entries valid, n_pos int64 [1])."""
    lib = nat.lib()
    _require_cuda("preds", preds, torch.float32)
    _require_cuda("labels", labels, torch.uint8)
    dev, n = preds.device, preds.numel()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        keys = torch.empty(n, dtype=torch.int32, device=dev)
        sorted_keys = torch.empty(n, dtype=torch.int32, device=dev)
        pos_keys = torch.empty(n, dtype=torch.int32, device=dev)
        n_pos = torch.zeros(1, dtype=torch.int64, device=dev)
        nat.check(
            lib.mb200_auc_build_keys(preds.data_ptr(), labels.data_ptr(), n, sigmoid_mode, _ptr(flags), keys.data_ptr(), pos_keys.data_ptr(),
                                     n_pos.data_ptr(), stream),
            "mb200_auc_build_keys",
        )
        ws = _workspace(dev, stream, "sort", lib.mb200_auc_sort_workspace_bytes(n))
        nat.check(lib.mb200_auc_sort_keys(keys.data_ptr(), sorted_keys.data_ptr(), n, ws.data_ptr(), ws.numel(), stream), "mb200_auc_sort_keys")
    return sorted_keys, pos_keys, n_pos


def auc_build_keys(preds: Tensor, labels: Tensor, sigmoid_mode: int, flags: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """Stage 1 alone: (keys uint32-as-int32 [n] with the positives marked, positive keys [n] with the first n_pos entries valid,
    n_pos int64 [1]) -- what the fused multi-GPU exchange consumes (dist.P2PExchange)."""
    lib = nat.lib()
    _require_cuda("preds", preds, torch.float32)
    _require_cuda("labels", labels, torch.uint8)
    dev, n = preds.device, preds.numel()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        keys = torch.empty(n, dtype=torch.int32, device=dev)
        pos_keys = torch.empty(n, dtype=torch.int32, device=dev)
        n_pos = torch.empty(1, dtype=torch.int64, device=dev)
        nat.check(lib.mb200_auc_build_keys(preds.data_ptr(), labels.data_ptr(), n, sigmoid_mode, _ptr(flags), keys.data_ptr(), pos_keys.data_ptr(),
                                           n_pos.data_ptr(), stream), "mb200_auc_build_keys")
    return keys, pos_keys, n_pos


def auc_rank_sum(sorted_keys: Tensor, n_pos_local: Tensor, pos_keys: Tensor, n_pos: Tensor, sum2: Tensor) -> None:
    """Stage 3: adds sum over ``pos_keys[:n_pos]`` of (lower_bound + upper_bound) in this rank's sorted
    negatives into ``sum2`` (int64 [1], holds a uint64)."""
    lib = nat.lib()
    dev = sorted_keys.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        nat.check(
            lib.mb200_auc_rank_sum(sorted_keys.data_ptr(), sorted_keys.numel(), n_pos_local.data_ptr(), pos_keys.data_ptr(), pos_keys.numel(),
                                   n_pos.data_ptr(), sum2.data_ptr(), stream),
            "mb200_auc_rank_sum",
        )


def read_bandwidth_probe(device: torch.device, buffer_bytes: int, repeats: int, rounds: int = 3) -> dict:
    """GB/s a warp-per-row gather of 3 KB rows can read from a buffer of about ``buffer_bytes`` on ``device`` (mb200_read_probe,
    best of ``rounds`` per launch shape, CUDA events).  48 MiB stays in the 126 MB L2 after the first pass: the L2 -> SM roof
    of the scoring kernel on Zipf-shaped ids; a buffer several times the L2 measures the HBM roof with the same access shape.
    Returns {"GBps": best, "by_mode": [...], "buffer_bytes": actual}."""
    lib = nat.lib()
    rows = 1 << max(int(buffer_bytes) // 3072, 1).bit_length() - 1  # power of two
    with torch.cuda.device(device):
        buf = torch.randint(0, 2**31 - 1, (rows * 768,), dtype=torch.int32, device=device)
        sink = torch.zeros(1, dtype=torch.int32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        by_mode = []
        for mode in (0, 1, 2):
            nat.check(lib.mb200_read_probe(buf.data_ptr(), rows * 3072, 1, mode, sink.data_ptr(), stream), "mb200_read_probe")  # warm: fills the L2
            best = 0.0
            for _ in range(rounds):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                nat.check(lib.mb200_read_probe(buf.data_ptr(), rows * 3072, repeats, mode, sink.data_ptr(), stream), "mb200_read_probe")
                e1.record()
                e1.synchronize()
                best = max(best, rows * 3072 * repeats / (e0.elapsed_time(e1) * 1e-3) / 1e9)
            by_mode.append(round(best, 1))
    return {"GBps": max(by_mode), "by_mode": by_mode, "buffer_bytes": rows * 3072}


def launch_counts() -> Tuple[int, int]:
    """(kernels of this library launched so far, CUB sort invocations so far) in this process."""
    lib = nat.lib()
    return int(lib.mb200_launch_count()), int(lib.mb200_library_launch_count())


def last_score_kernel_ms() -> float:
    """Duration of the most recent fused kernel launched with ``set_tuning(time_kernel=1)`` (synchronises)."""
    return float(nat.lib().mb200_last_score_kernel_ms())


def last_hot_stats() -> dict:
    """Hot-row cache of the most recent fused launch: slots used / available per module and the sampled fraction of row reads
    it serves from shared memory (mb200_last_hot_stats; synchronises)."""
    out = (ctypes.c_int32 * 4)()
    nat.check(nat.lib().mb200_last_hot_stats(out), "mb200_last_hot_stats")
    n_hot, total, covered, cap = (int(v) for v in out)
    return {"rows_cached_per_module": n_hot, "slots_per_module": cap, "sampled_row_reads": total,
            "hit_fraction": (covered / total) if (total > 0 and n_hot > 0) else 0.0}


def set_tuning(chunks_per_warp: Optional[int] = None, variant: Optional[int] = None, ctas_per_sm: Optional[int] = None,
               time_kernel: Optional[int] = None, retrieval_diag: Optional[int] = None, retrieval_pair: Optional[int] = None,
               hot_kb_cap: Optional[int] = None, static_chunks: Optional[int] = None, retrieval_window: Optional[int] = None) -> None:
    lib = nat.lib()
    for key, val in ((0, chunks_per_warp), (1, variant), (2, ctas_per_sm), (3, time_kernel), (4, retrieval_diag), (5, retrieval_pair), (6, hot_kb_cap),
                     (7, static_chunks), (8, retrieval_window)):
        if val is not None:
            lib.mb200_set_tuning(key, int(val))
