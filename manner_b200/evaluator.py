"""ScoreEvaluator: the host side of the B200 scoring / ensemble / metrics path.

It owns what the reference's LightningModules own between "news vectors" and "logged metrics":

  reference (file:line under manner/)                         here
  ---------------------------------------------------------  --------------------------------------
  news_encoder(...) per batch      cr_module.py:107,113       cached embedding tables on the device
  MINDRecBatch segment ids         mind_rec_dataset.py:114    CSR behaviours (data.Behaviours)
  CRModule.forward / model_step    cr_module.py:105-184       torch.ops.manner_b200.score_eval
  EnsembleModule.forward           ensemble_module.py:95-151  same op, zscore=True + weights
  on_test_epoch_end + log_dict     cr_module.py:266-274       EvalResult.metrics() (same log keys)
  AUROC(task="binary")             cr_module.py:81            torch.ops.manner_b200.pooled_auc

One ``evaluate`` call is one pass over all impressions handed to it (an epoch, or a rank's shard of
it) -- the reference also computes its metrics once per epoch.  With a process group the additive
results are all-reduced over NCCL (dist.py); nothing else crosses GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import _native as nat
from . import dist as mdist
from . import ops
from .data import NUM_CATEG_CLASSES, NUM_SENT_CLASSES, Behaviours, step_pads

# log keys of the reference (cr_module.py:79-89 prefix "test/"; ensemble_module.py:50-84) by metric slot
SLOT_KEYS = {
    nat.M_MRR: "mrr",
    nat.M_NDCG_K0: "ndcg@{k0}",
    nat.M_NDCG_K1: "ndcg@{k1}",
    nat.M_CATEG_DIV_K0: "categ_div@{k0}",
    nat.M_CATEG_DIV_K1: "categ_div@{k1}",
    nat.M_SENT_DIV_K0: "sent_div@{k0}",
    nat.M_SENT_DIV_K1: "sent_div@{k1}",
    nat.M_CATEG_PERS_K0: "categ_pers@{k0}",
    nat.M_CATEG_PERS_K1: "categ_pers@{k1}",
    nat.M_SENT_PERS_K0: "sent_pers@{k0}",
    nat.M_SENT_PERS_K1: "sent_pers@{k1}",
}


@dataclass
class DeviceBehaviours:
    """CSR behaviours resident on the device (+ the two host-side facts the launch needs)."""

    hist_offsets: Tensor
    hist_ids: Tensor
    cand_offsets: Tensor
    cand_ids: Tensor
    labels: Tensor
    n_impressions: int
    max_cand: int
    h2d_bytes: int = 0
    n_pos: int = 0  # positives in this shard (host-known from the labels)
    pos_cap: Optional[int] = None  # agreed upper bound on any rank's positives (multi-GPU pooled AUC)
    # only for early fusion / the loss: the reference's step structure (configs/data/mind_rec.yaml:51) and the zero
    # rows / columns its dense batches append per impression
    step_batch: int = 8
    hist_pad: Optional[Tensor] = None
    cand_pad: Optional[Tensor] = None
    # pipelined upload (upload(..., pipelined=True)): the device word the copy stream raises as segments arrive, the number of
    # segments, and the call that enqueues the segments still missing (made by launch() right behind the fused kernel)
    ready: Optional[Tensor] = None
    ready_segments: int = 0
    _finish_upload: Optional[object] = None

    def finish_upload(self) -> None:
        """Waits until the library's thread has queued every segment copy of a pipelined upload and raises if one failed
        (idempotent; launch() calls it behind the fused kernel's launch)."""
        if self._finish_upload is not None:
            fin, self._finish_upload = self._finish_upload, None
            fin()

    def __del__(self) -> None:
        try:  # the pinned host arrays the copies read from must outlive the queueing
            self.finish_upload()
        except Exception:
            pass


@dataclass
class EvalResult:
    sums: np.ndarray  # fp64 [W, NUM_METRICS], summed over all ranks
    n_impressions: int  # over all ranks
    flags: int
    ks: Tuple[int, int]
    has_aspects: bool
    auc: Optional[float] = None  # pooled AUROC (reference "auc"), weighting `scores_weighting`
    auc_counts: Optional[Tuple[int, int]] = None  # (positives, negatives)
    loss: Optional[float] = None  # the reference's test/loss: MeanMetric over its steps (cr_module.py:253-259)
    scores: Optional[Tensor] = None  # device fp32 [sum C] of this rank
    per_impression: Optional[Tensor] = None  # device fp32 [W, B, NUM_METRICS] of this rank
    d2h_bytes: int = 0

    def metrics(self, weighting: int = 0, prefix: str = "test/") -> Dict[str, float]:
        """Means under the reference's log keys (+ ``gauc``, which the reference does not have)."""
        s = self.sums[weighting]
        n = max(self.n_impressions, 1)
        out: Dict[str, float] = {}
        for slot, key in SLOT_KEYS.items():
            if slot >= nat.M_CATEG_DIV_K0 and not self.has_aspects:
                continue
            out[prefix + key.format(k0=self.ks[0], k1=self.ks[1])] = float(s[slot] / n)
        out[prefix + "gauc"] = float(s[nat.M_GAUC] / s[nat.M_GAUC_VALID]) if s[nat.M_GAUC_VALID] > 0 else 0.0
        if self.auc is not None:
            out[prefix + "auc"] = self.auc
        if self.loss is not None:
            out[prefix + "loss"] = self.loss
        return out


@dataclass
class PendingEval:
    """Device-side results of an enqueued pass (no host synchronisation yet)."""

    sums: Tensor  # [W, NUM_METRICS], or the packed all-reduced payload [W*NUM_METRICS + 5] when distributed
    flags: Tensor
    n_weightings: int
    n_impressions: int  # of this rank
    distributed: bool
    auc_stats: Optional[Tensor]
    scores: Optional[Tensor]
    per_impression: Optional[Tensor]
    loss_stats: Optional[Tensor] = None  # fp64 [2]: sum of step losses, number of steps (all ranks)
    n_pos_bound: int = 0
    fused: bool = False  # distributed through the fused exchange: `sums` = [payload, sum2, P, N, exchange flags] (one buffer, one read)
    has_auc: bool = False


class ScoreEvaluator:
    """Cached embedding tables + aspect labels on one device; ``evaluate`` runs the fused path.

    tables[0] is the CR-Module's table, tables[1:] the A-Modules' (category, sentiment, ...), all
    [n_news, dim] fp32 or bf16 with the same shape (the table is the new boundary, filled
    once by the PyTorch news encoders)."""

    def __init__(
        self,
        tables: Sequence[Tensor],
        device: Union[str, torch.device, None] = None,
        news_category: Optional[Union[Tensor, np.ndarray]] = None,
        news_sentiment: Optional[Union[Tensor, np.ndarray]] = None,
        num_categ_classes: int = NUM_CATEG_CLASSES,
        num_sent_classes: int = NUM_SENT_CLASSES,
        ks: Tuple[int, int] = (5, 10),
        attention: Optional[Sequence[Optional[Tuple[Tensor, Tensor, Tensor]]]] = None,
        table_shards: Optional[Sequence[Sequence[Tensor]]] = None,
        n_news: Optional[int] = None,
        exchange: str = "p2p",
    ) -> None:
        """``table_shards`` (instead of ``tables``): row-sharded tables for catalogues too large to replicate --
        ``table_shards[m]`` = the R shards of module m, each [2**s, dim] on ITS OWN GPU (this rank's shard plus the peers'
        opened with ``dist.share_table_shards``), ``n_news`` the catalogue size; row n lives in shard n >> s.  The kernel
        reads remote rows over NVLink.  Late fusion, reference width."""
        nat.lib()  # fail now, loudly, if the CUDA library is not built
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' (fused: stores into the peers' mailboxes over NVLink) or 'nccl' (three collectives)")
        self.exchange = exchange  # how a distributed evaluation meets the other ranks (dist.P2PExchange / dist.pooled_auc_distributed)
        self._p2p: Optional[mdist.P2PExchange] = None
        self._copy_stream: Optional[torch.cuda.Stream] = None  # pipelined uploads run here
        if not torch.cuda.is_available():
            raise RuntimeError("manner_b200.ScoreEvaluator needs a CUDA device; there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_table_shards, self.table_shard_shift = 1, 0
        if table_shards is not None:
            if attention is not None or n_news is None:
                raise ValueError("row-sharded tables need n_news and support late fusion only")
            r, rows = len(table_shards[0]), table_shards[0][0].shape[0]
            if rows & (rows - 1) or any(len(s) != r for s in table_shards) or not 1 <= len(table_shards) <= nat.MAX_MODULES:
                raise ValueError("every module needs the same number of shards of 2**s rows each")
            self.n_modules = len(table_shards)
            self.tables = [t for shards in table_shards for t in shards]  # module-major, as ops.score_eval takes them
            self.n_table_shards, self.table_shard_shift = r, rows.bit_length() - 1
            self.n_news, self.dim = int(n_news), self.tables[0].shape[1]
        else:
            if not 1 <= len(tables) <= nat.MAX_MODULES:
                raise ValueError(f"1..{nat.MAX_MODULES} tables expected")
            self.tables: List[Tensor] = [t.to(self.device, non_blocking=True).contiguous() for t in tables]
            self.n_modules = len(self.tables)
            self.n_news, self.dim = self.tables[0].shape
        self.ks = (int(ks[0]), int(ks[1]))
        self.num_categ_classes, self.num_sent_classes = int(num_categ_classes), int(num_sent_classes)
        self.news_category = self._aspect(news_category)
        self.news_sentiment = self._aspect(news_sentiment)
        if (self.news_category is None) != (self.news_sentiment is None):
            raise ValueError("give both aspect label arrays or neither")
        # early fusion (late_fusion=False, cr_module.py:63-68,124-125): attention[m] = (linear.weight [Q, D], linear.bias [Q],
        # query [Q]) of module m's NAMLUserEncoder.additive_attention, or None for late fusion.  The per-news logits are
        # computed once here and cached next to the table.
        self.attn_logits: Optional[List[Optional[Tensor]]] = None
        if attention is not None and any(a is not None for a in attention):
            if len(attention) != len(self.tables):
                raise ValueError("attention needs one entry per table")
            self.attn_logits = [None if a is None else ops.attention_logits(t, *a) for t, a in zip(self.tables, attention)]

    def _aspect(self, a: Optional[Union[Tensor, np.ndarray]]) -> Optional[Tensor]:
        if a is None:
            return None
        t = torch.as_tensor(np.asarray(a) if not isinstance(a, Tensor) else a).to(torch.int32)
        if t.numel() != self.n_news:
            raise ValueError("aspect labels must have one entry per news row")
        return t.to(self.device).contiguous()

    # -- inputs ----------------------------------------------------------------------------------------------
    def upload(self, bhv: Behaviours, pinned: Optional[Dict[str, object]] = None, pos_cap: Optional[int] = None,
               step_batch: Optional[int] = None, pipelined: bool = False, segments: int = 5,
               worker_segments: int = 0) -> DeviceBehaviours:
        """Host CSR -> device (asynchronous on the current stream).  ``pinned`` lets a caller reuse
        page-locked staging tensors (see ``pin``); ``pos_cap`` is the multi-GPU bound of
        ``dist.agree_pos_cap`` when the caller already has it.  ``step_batch`` (the reference's eval batch size)
        additionally uploads the per-impression pad counts early fusion and the cross-entropy loss need.

        ``pipelined``: the copy overlaps the pass instead of preceding it (mb200_upload_begin / _finish): the offsets go first,
        the id / label arrays follow in ``segments`` segments of geometrically growing size on a copy stream, and the fused kernel -- launched
        by the next ``launch`` / ``evaluate`` with the returned object -- starts on the first segment while the others are in
        flight.  The returned arrays must not be read by anything else before that launch.  ``worker_segments`` = k lets a thread
        of the library queue the last k segment copies while this thread goes on to launch the kernel (see prepared.py: opt-in,
        tools that serialise CUDA calls behind a running kernel can starve those copies)."""
        src = pinned if pinned is not None else self.pin(bhv, step_batch)
        if pos_cap is not None and pos_cap < src["n_pos"]:
            raise ValueError(f"pos_cap {pos_cap} is below this shard's {src['n_pos']} positives: the pooled AUROC would silently drop keys (dist.agree_pos_cap)")
        nbytes = sum(v.numel() * v.element_size() for k, v in src.items() if isinstance(v, Tensor) and k != "marks")
        if pipelined and bhv.n_impressions >= 64 * segments and self.n_table_shards == 1:
            return self._upload_pipelined(bhv, src, pos_cap, nbytes, int(segments), int(worker_segments))
        dev = {k: v.to(self.device, non_blocking=True) for k, v in src.items() if isinstance(v, Tensor) and k != "marks"}
        return DeviceBehaviours(
            dev["hist_offsets"], dev["hist_ids"], dev["cand_offsets"], dev["cand_ids"], dev["labels"],
            bhv.n_impressions, src["max_cand"], nbytes, src["n_pos"], pos_cap,
            src.get("step_batch", 8), dev.get("hist_pad"), dev.get("cand_pad"),
        )

    def _upload_pipelined(self, bhv: Behaviours, src: Dict[str, object], pos_cap: Optional[int], nbytes: int, segments: int,
                          worker_segments: int = 0) -> DeviceBehaviours:
        import ctypes

        lib = nat.lib()
        segments = max(1, min(segments, nat.MAX_UPLOAD_SEGMENTS))
        if "marks" not in src:
            src["marks"] = torch.zeros(nat.MAX_UPLOAD_SEGMENTS, dtype=torch.int32).pin_memory()
        with torch.cuda.device(self.device):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            compute = torch.cuda.current_stream(self.device)
            dev = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in src.items() if isinstance(v, Tensor) and k != "marks"}
            for t in dev.values():
                t.record_stream(self._copy_stream)  # written there: the allocator must not recycle the block under the copies
            ready = torch.empty(1, dtype=torch.int32, device=self.device)
            d = nat.UploadDesc()
            d.struct_size = ctypes.sizeof(nat.UploadDesc)
            d.n_segments, d.segments_first, d.n_impressions = segments, max(1, segments - max(0, worker_segments)), bhv.n_impressions
            for name in ("hist_offsets", "hist_ids", "cand_offsets", "cand_ids", "labels"):
                setattr(d, "h_" + name, src[name].data_ptr())
                setattr(d, "d_" + name, dev[name].data_ptr())
            for name in ("hist_pad", "cand_pad"):
                if name in dev:
                    setattr(d, "h_" + name, src[name].data_ptr())
                    setattr(d, "d_" + name, dev[name].data_ptr())
            d.ready, d.h_marks, d.copy_stream = ready.data_ptr(), src["marks"].data_ptr(), self._copy_stream.cuda_stream
            nat.check(lib.mb200_upload_begin(ctypes.byref(d), compute.cuda_stream), "mb200_upload_begin")

        def finish(d=d, keep=(src, dev, ready)) -> None:
            nat.check(lib.mb200_upload_finish(ctypes.byref(d)), "mb200_upload_finish")

        out = DeviceBehaviours(
            dev["hist_offsets"], dev["hist_ids"], dev["cand_offsets"], dev["cand_ids"], dev["labels"],
            bhv.n_impressions, src["max_cand"], nbytes, src["n_pos"], pos_cap,
            src.get("step_batch", 8), dev.get("hist_pad"), dev.get("cand_pad"), ready, segments, finish,
        )
        return out

    @staticmethod
    def pin(bhv: Behaviours, step_batch: Optional[int] = None) -> Dict[str, object]:
        """Page-locked staging copies of the CSR arrays + the two host-side facts a launch needs
        (largest candidate list, number of positives), computed once here rather than per upload."""
        arrays = dict(hist_offsets=bhv.hist_offsets, hist_ids=bhv.hist_ids, cand_offsets=bhv.cand_offsets, cand_ids=bhv.cand_ids, labels=bhv.labels)
        if step_batch is not None:
            arrays["hist_pad"] = step_pads(bhv.hist_offsets, step_batch)
            arrays["cand_pad"] = step_pads(bhv.cand_offsets, step_batch)
        out: Dict[str, object] = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in arrays.items()}
        out["max_cand"] = max(bhv.max_cand, 1)
        out["n_pos"] = int(bhv.labels.sum())
        if step_batch is not None:
            out["step_batch"] = int(step_batch)
        return out

    # -- the hot path ----------------------------------------------------------------------------------------
    def launch(
        self,
        bhv: DeviceBehaviours,
        weights: Optional[Union[Sequence[Sequence[float]], Tensor]] = None,
        zscore: bool = False,
        pooled_auc: bool = False,
        want_scores: bool = False,
        want_per_impression: bool = False,
        scores_weighting: int = 0,
        group: Optional["torch.distributed.ProcessGroup"] = None,
        distributed: bool = False,
        loss: Optional[str] = None,
        temperature: float = 0.1,
    ) -> "PendingEval":
        """Enqueue one pass on the current stream and return device-side results without waiting for
        them.  ``weights`` may be a device fp32 tensor [W, n_modules] (then every module is gathered).
        ``loss`` = "ce" | "supcon" adds the reference's test/loss (cr_module.py:140-171,253-259; ``temperature`` for
        SupCon); it and early fusion need ``upload(..., step_batch=...)``."""
        if (loss is not None or self.attn_logits is not None) and bhv.hist_pad is None:
            raise ValueError("early fusion / the losses depend on the reference's step structure: upload(..., step_batch=8)")
        n_mod = self.n_modules
        w_dev: Optional[Tensor] = None
        active = (1 << n_mod) - 1
        if isinstance(weights, Tensor) and weights.is_cuda:
            w_dev = weights.float().reshape(-1, n_mod).contiguous()
        elif weights is not None:
            w_host = torch.as_tensor(weights, dtype=torch.float32).reshape(-1, n_mod)
            # a module whose weight is 0 everywhere is never gathered (ensemble_module.py:37-46,100-107)
            active = 1
            for m in range(1, n_mod):
                if bool((w_host[:, m] != 0).any()):
                    active |= 1 << m
            w_dev = w_host.to(self.device, non_blocking=True).contiguous()
        need_scores = want_scores or pooled_auc
        loss_kind = {None: nat.LOSS_NONE, "ce": nat.LOSS_CE, "supcon": nat.LOSS_SUPCON}[loss]
        if loss is not None and w_dev is not None and w_dev.shape[0] >= 16 and self.news_category is None:
            # >= 16 weightings without aspects take the lane-per-weighting sweep path, which fills the ranking slots only
            raise ValueError("the loss is not computed in the aspect-weight sweep mode: evaluate it with a single weighting")
        scores, per_impr, sums, flags, loss_per_impr = torch.ops.manner_b200.score_eval(
            self.tables, bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, bhv.cand_ids, bhv.labels, w_dev, zscore,
            bhv.max_cand, active, self.ks[0], self.ks[1], need_scores, scores_weighting, want_per_impression,
            self.news_category, self.news_sentiment, self.num_categ_classes, self.num_sent_classes,
            self.attn_logits if self.attn_logits is not None else [], distributed,
            bhv.hist_pad if self.attn_logits is not None else None, loss_kind, float(temperature),
            bhv.cand_pad if loss is not None else None, self.n_table_shards, self.table_shard_shift, self.n_news,
            bhv.ready, bhv.ready_segments,
        )
        bhv.finish_upload()  # pipelined upload: by now the library's thread has queued the segment copies; collect its status
        loss_stats: Optional[Tensor] = None
        if loss is not None:
            # MeanMetric over the reference's steps (cr_module.py:253-259): (sum of step losses, number of steps)
            loss_stats = ops.step_loss(loss_per_impr, bhv.step_batch, loss_kind, bhv.cand_offsets, bhv.labels)
            if distributed:
                torch.distributed.all_reduce(loss_stats, op=torch.distributed.ReduceOp.SUM, group=group)
        n_w = 1 if w_dev is None else w_dev.shape[0]
        auc_stats: Optional[Tensor] = None
        fused = False
        if distributed and self.exchange == "p2p" and (group is None or torch.distributed.get_backend(group) == "nccl"):
            # fused exchange: payload + positive keys stored into every rank's mailbox over NVLink, AUROC statistics likewise;
            # two kernels, no collective.  Creating / growing the mailboxes is collective and happens outside the hot loop
            # (every rank sees the same n_payload and the agreed pos_cap, so all ranks do it in the same call).
            cap = bhv.pos_cap if bhv.pos_cap is not None else mdist.agree_pos_cap(bhv.n_pos, self.device, group)
            if self._p2p is None or not self._p2p.fits(sums.numel(), cap):
                if self._p2p is not None:
                    self._p2p.close()  # unmap the peers' old mailboxes before they are replaced (every rank is here at the same call)
                self._p2p = mdist.P2PExchange(self.device, sums.numel(), max(cap, 1), group)
            outside = n_w * nat.NUM_METRICS + 1 + 2  # tail entry of the MB200_FLAG_OUTSIDE_UNIT bit
            if pooled_auc:
                sorted_keys, pos_keys, n_pos = ops.auc_build_and_sort(scores, bhv.labels, 0, None)  # raw score order; the sigmoid rule is applied after the exchange
                sums = self._p2p.run(sums, outside, sorted_keys, pos_keys, n_pos)
            else:
                sums = self._p2p.run(sums, -1)
            fused = True
        elif distributed:
            # `sums` is the packed payload [W*NUM_METRICS sums, impression count, 4 flag bits]: one NCCL all-reduce
            torch.distributed.all_reduce(sums, op=torch.distributed.ReduceOp.SUM, group=group)
            if pooled_auc:
                outside = n_w * nat.NUM_METRICS + 1 + 2  # bit index of MB200_FLAG_OUTSIDE_UNIT in the tail
                gflags = (sums[outside : outside + 1] > 0).to(torch.int32) * nat.FLAG_OUTSIDE_UNIT
                auc_stats = mdist.pooled_auc_distributed(scores, bhv.labels, gflags, group, pos_cap=bhv.pos_cap)
        elif pooled_auc:
            # (mb200_pooled_auc_bounded -- sort only the positives, stream the negatives against them -- gives the same integers but
            # measured slower on click-log shapes: 0.201 vs 0.185 ms, profiles/r2_n_schedule.log; the full sort stays the default)
            auc_stats = torch.ops.manner_b200.pooled_auc(scores, bhv.labels, 2, flags)
        return PendingEval(sums, flags, n_w, bhv.n_impressions, distributed, auc_stats,
                           scores if want_scores else None, per_impr if want_per_impression else None, loss_stats, bhv.n_pos, fused, pooled_auc)

    def finish(self, pending: "PendingEval") -> EvalResult:
        """The one device -> host read of a pass: metric sums, flag word, AUC statistics."""
        n_block = pending.n_weightings * nat.NUM_METRICS
        auc = counts = None
        if pending.fused:
            raw = pending.sums.cpu()  # the ONE device -> host read: reduced payload + AUROC integers + exchange flags
            packed, tail = raw[: n_block + nat.PAYLOAD_TAIL].numpy(), raw[n_block + nat.PAYLOAD_TAIL :].view(torch.int64).tolist()
            d2h = raw.numel() * 8
            if tail[3] & nat.FLAG_EXCHANGE_TIMEOUT:
                raise nat.NativeError("fused multi-GPU exchange: a peer GPU's stores did not arrive within 4 s (a rank died or skipped the call)")
            if tail[3] & nat.FLAG_POS_OVERFLOW:
                raise nat.NativeError("fused multi-GPU exchange: a rank had more positives than the agreed pos_cap (dist.agree_pos_cap)")
            sums_h, n_total = packed[:n_block].reshape(pending.n_weightings, nat.NUM_METRICS), int(round(packed[n_block]))
            flags_h = mdist.flags_from_payload_tail(packed[n_block + 1 : n_block + 1 + mdist.N_FLAG_BITS])
            if pending.has_auc:
                s2, p, n = tail[0], tail[1], tail[2]
                auc, counts = (s2 / (2.0 * p * n) if p > 0 and n > 0 else 0.0), (p, n)
        elif pending.distributed:
            packed = pending.sums.cpu().numpy()
            sums_h, n_total = packed[:n_block].reshape(pending.n_weightings, nat.NUM_METRICS), int(round(packed[n_block]))
            flags_h = mdist.flags_from_payload_tail(packed[n_block + 1 : n_block + 1 + mdist.N_FLAG_BITS])
            d2h = packed.size * 8
            if pending.auc_stats is not None:
                auc, p, n = mdist.auc_from_stats(pending.auc_stats)
                counts = (p, n)
                d2h += 24
        else:
            parts = [pending.sums.reshape(-1), pending.flags.to(torch.float64)]
            if pending.auc_stats is not None:
                parts.append(pending.auc_stats)
            packed = torch.cat(parts).cpu().numpy()  # one read: sums, flag word, AUROC statistics
            sums_h, flags_h, n_total = packed[:n_block].reshape(pending.n_weightings, nat.NUM_METRICS), int(packed[n_block]), pending.n_impressions
            d2h = packed.size * 8
            if pending.auc_stats is not None:
                a = packed[n_block + 1 :]
                auc, counts = float(a[0]), (int(a[1]), int(a[2]))
                if auc != auc:
                    raise nat.NativeError(f"pooled AUROC: the rows hold {counts[0]} positives, more than the {pending.n_pos_bound} the upload counted")
        loss_value = None
        if pending.loss_stats is not None:
            ls = pending.loss_stats.cpu().numpy()
            d2h += 16
            loss_value = float(ls[0] / ls[1]) if ls[1] > 0 else 0.0
        if flags_h & nat.FLAG_UPLOAD_TIMEOUT:
            raise nat.NativeError("pipelined upload: a segment of the behaviour set did not reach the device within 4 s (copy stream stalled?)")
        if flags_h & (nat.FLAG_BAD_ID | nat.FLAG_CAND_OVERFLOW | nat.FLAG_BAD_ASPECT):
            raise nat.NativeError(
                f"manner_b200 kernels flagged bad input (flags={flags_h}): "
                "1=row id outside the table, 2=impression longer than max_cand, 8=aspect label outside [0, num_classes)"
            )
        return EvalResult(
            sums=sums_h, n_impressions=int(n_total), flags=flags_h, ks=self.ks, has_aspects=self.news_category is not None,
            auc=auc, auc_counts=counts, scores=pending.scores, per_impression=pending.per_impression, d2h_bytes=d2h, loss=loss_value,
        )

    def prepare(self, bhv: Behaviours, pinned: Optional[Dict[str, object]] = None, **kwargs):
        """A ``PreparedPass`` over a fixed behaviour set (the validation / test split evaluated after every epoch): buffers,
        descriptors and workspaces are set up once, ``run()`` is then upload (overlapped) + pass + one pinned read-back in a
        handful of C-ABI calls -- the host no longer keeps the stream waiting.  Keywords: weights, zscore, pooled_auc, loss,
        temperature, step_batch, segments, distributed, group, pos_cap, want_scores (manner_b200/prepared.py)."""
        from .prepared import PreparedPass

        return PreparedPass(self, bhv, pinned, **kwargs)

    def evaluate(self, bhv: DeviceBehaviours, **kwargs) -> EvalResult:
        """One pass: scores (+ z-score ensemble for every row of ``weights`` [W, n_modules]) and metric
        means.  ``zscore=False, weights=None`` is the CRModule evaluation (cr_module.py:105-131,266-274);
        ``zscore=True`` with ``weights=[[1, categ_weight, sent_weight]]`` is the EnsembleModule
        (ensemble_module.py:95-151,214-238).  With ``distributed=True`` the sums (and the pooled-AUC rank
        statistic) are reduced over ``group`` with NCCL.  See ``launch`` for the arguments."""
        return self.finish(self.launch(bhv, **kwargs))
