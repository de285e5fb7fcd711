"""Drop-in metric objects for the reference's epoch-end seam: ``metrics(preds, targets, indexes=...)`` then ``compute()``.

Every LightningModule of the reference -- CRModule, EnsembleModule and the nine baseline recommenders -- finishes an epoch by
handing the flat prediction vector to a torchmetrics ``MetricCollection`` (cr_module.py:79-89,266-274;
nrms_plm_module.py:58-68,275-313) and, for the ensemble, to ``Diversity`` / ``Personalization`` objects
(ensemble_module.py:56-84,214-238; manner/metrics/*.py).  ``RetrievalMetricsB200`` has the same call protocol and log keys and
runs on the B200 kernels (``mb200_rank_metrics`` + ``mb200_pooled_auc``): one warp per impression instead of torchmetrics'
Python loop over every impression.  No CPU fallback: tensors must be on the GPU.

    metrics = RetrievalMetricsB200(prefix="test/")                       # auc, mrr, ndcg@5, ndcg@10 (+ gauc)
    metrics.update(preds, targets, indexes=indexes)                      # any number of times
    self.log_dict(metrics.compute()); metrics.reset()
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _native as nat
from .evaluator import SLOT_KEYS
from .ops import _ptr, _require_cuda, _workspace


def rank_metrics(
    preds: Tensor,
    labels: Tensor,
    cand_offsets: Tensor,
    max_cand: int,
    ks: Tuple[int, int] = (5, 10),
    cand_category: Optional[Tensor] = None,
    cand_sentiment: Optional[Tensor] = None,
    hist_offsets: Optional[Tensor] = None,
    hist_category: Optional[Tensor] = None,
    hist_sentiment: Optional[Tensor] = None,
    num_categ_classes: int = 19,
    num_sent_classes: int = 4,
    want_per_impression: bool = False,
) -> Tuple[Tensor, Tensor, Tensor]:
    """(sums fp64 [NUM_METRICS], per_impression fp32 [B, NUM_METRICS] or empty, flags int32 [1]) -- mb200_rank_metrics on flat
    predictions whose impressions are contiguous (``cand_offsets`` int32 [B + 1])."""
    lib = nat.lib()
    _require_cuda("preds", preds, torch.float32)
    _require_cuda("labels", labels, torch.uint8)
    _require_cuda("cand_offsets", cand_offsets, torch.int32)
    aspect = (cand_category, cand_sentiment, hist_offsets, hist_category, hist_sentiment)
    if any(a is not None for a in aspect):
        if any(a is None for a in aspect):
            raise ValueError("aspect metrics need cand_category, cand_sentiment, hist_offsets, hist_category and hist_sentiment together")
        for name, a in zip(("cand_category", "cand_sentiment", "hist_offsets", "hist_category", "hist_sentiment"), aspect):
            _require_cuda(name, a, torch.int32)
    n_impr = cand_offsets.numel() - 1
    if labels.numel() != preds.numel():
        raise ValueError("preds and labels must have one entry per candidate row")
    dev = preds.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        sums = torch.empty(nat.NUM_METRICS, dtype=torch.float64, device=dev)
        per = torch.empty((n_impr, nat.NUM_METRICS) if want_per_impression else (0,), dtype=torch.float32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        d = nat.MetricsDesc()
        d.struct_size = ctypes.sizeof(nat.MetricsDesc)
        d.k0, d.k1, d.max_cand, d.n_impressions = int(ks[0]), int(ks[1]), int(max(max_cand, 1)), n_impr
        d.preds, d.labels, d.cand_offsets = preds.data_ptr(), labels.data_ptr(), cand_offsets.data_ptr()
        d.cand_category, d.cand_sentiment = _ptr(cand_category), _ptr(cand_sentiment)
        d.hist_offsets, d.hist_category, d.hist_sentiment = _ptr(hist_offsets), _ptr(hist_category), _ptr(hist_sentiment)
        d.num_categ_classes, d.num_sent_classes = int(num_categ_classes), int(num_sent_classes)
        d.per_impression = per.data_ptr() if want_per_impression else None
        d.sums, d.flags = sums.data_ptr(), flags.data_ptr()
        need = lib.mb200_metrics_workspace_bytes(ctypes.byref(d))
        ws = _workspace(dev, stream, "metrics", max(need, 256))
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        nat.check(lib.mb200_rank_metrics(ctypes.byref(d), stream), "mb200_rank_metrics")
    return sums, per, flags


def _offsets_from_indexes(indexes: Tensor) -> Tuple[Tensor, Optional[Tensor], int]:
    """Group ids (torchmetrics ``indexes``) -> (int32 CSR offsets over the groups in ascending id order, permutation that makes
    the groups contiguous or None when they already are, largest group).  Mirrors ``RetrievalMetric.compute``: sort by index,
    ``_flexible_bincount``, split (metrics/base.py:99-111)."""
    idx = indexes.long().flatten()
    perm: Optional[Tensor] = None
    if idx.numel() > 1 and bool((idx[1:] < idx[:-1]).any()):
        idx, perm = torch.sort(idx, stable=True)
    _, counts = torch.unique_consecutive(idx, return_counts=True)
    off = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=idx.device)
    off[1:] = torch.cumsum(counts, 0)
    return off, perm, int(counts.max().item()) if counts.numel() else 0


class RetrievalMetricsB200:
    """``MetricCollection({"auc": AUROC(task="binary"), "mrr": RetrievalMRR(), "ndcg@5": ..., "ndcg@10": ...})`` of
    cr_module.py:79-89 -- and, when aspect labels are passed to ``update``, the Diversity / Personalization collections of
    ensemble_module.py:56-84 -- as one object with the torchmetrics protocol (``update`` / ``__call__`` / ``compute`` / ``reset``)."""

    def __init__(self, prefix: str = "", ks: Tuple[int, int] = (5, 10), with_auc: bool = True, num_categ_classes: int = 19,
                 num_sent_classes: int = 4) -> None:
        nat.lib()
        self.prefix, self.ks, self.with_auc = prefix, (int(ks[0]), int(ks[1])), with_auc
        self.num_categ_classes, self.num_sent_classes = int(num_categ_classes), int(num_sent_classes)
        self.reset()

    def reset(self) -> None:
        self._preds: List[Tensor] = []
        self._targets: List[Tensor] = []
        self._indexes: List[Tensor] = []
        self._aspects: Dict[str, List[Tensor]] = {k: [] for k in ("tc", "ts", "hc", "hs", "hi")}

    def update(self, preds: Tensor, target: Tensor, indexes: Tensor, target_categories: Optional[Tensor] = None,
               target_sentiments: Optional[Tensor] = None, hist_categories: Optional[Tensor] = None, hist_sentiments: Optional[Tensor] = None,
               hist_indexes: Optional[Tensor] = None) -> None:
        if indexes is None:
            raise ValueError("Argument `indexes` cannot be None")  # torchmetrics' message
        if not preds.is_cuda:
            raise RuntimeError("manner_b200: `preds` must be a CUDA tensor (there is no CPU path)")
        # like torchmetrics, rows of different updates that carry the same index form one group
        self._preds.append(preds.detach().float().flatten())
        self._targets.append(target.detach().flatten())
        self._indexes.append(indexes.detach().flatten().to(preds.device))
        given = [target_categories, target_sentiments, hist_categories, hist_sentiments, hist_indexes]
        if any(g is not None for g in given):
            if any(g is None for g in given):
                raise ValueError("aspect metrics need target_categories, target_sentiments, hist_categories, hist_sentiments and hist_indexes")
            for key, g in zip(("tc", "ts", "hc", "hs", "hi"), given):
                self._aspects[key].append(g.detach().flatten().to(preds.device))

    def __call__(self, preds: Tensor, target: Tensor, **kwargs: Tensor) -> Dict[str, float]:
        """torchmetrics' forward: accumulate, and return the value of the accumulated state."""
        self.update(preds, target, **kwargs)
        return self.compute()

    def compute(self) -> Dict[str, float]:
        if not self._preds:
            return {}
        preds, target, indexes = torch.cat(self._preds), torch.cat(self._targets), torch.cat(self._indexes)
        off, perm, max_cand = _offsets_from_indexes(indexes)
        if perm is not None:
            preds, target = preds[perm], target[perm]
        labels = (target != 0).to(torch.uint8).contiguous()
        preds = preds.contiguous()
        kw = {}
        has_aspects = bool(self._aspects["tc"])
        if has_aspects:
            tc, ts = torch.cat(self._aspects["tc"]), torch.cat(self._aspects["ts"])
            if perm is not None:
                tc, ts = tc[perm], ts[perm]
            hoff, hperm, _ = _offsets_from_indexes(torch.cat(self._aspects["hi"]))
            hc, hs = torch.cat(self._aspects["hc"]), torch.cat(self._aspects["hs"])
            if hperm is not None:
                hc, hs = hc[hperm], hs[hperm]
            if hoff.numel() != off.numel():
                raise ValueError("`indexes` and `hist_indexes` must name the same impressions")
            kw = dict(cand_category=tc.to(torch.int32).contiguous(), cand_sentiment=ts.to(torch.int32).contiguous(), hist_offsets=hoff,
                      hist_category=hc.to(torch.int32).contiguous(), hist_sentiment=hs.to(torch.int32).contiguous(),
                      num_categ_classes=self.num_categ_classes, num_sent_classes=self.num_sent_classes)
        sums, _, flags = rank_metrics(preds, labels, off, max_cand, self.ks, **kw)
        auc_stats = None
        if self.with_auc:
            outside = (~((preds >= 0) & (preds <= 1)).all()).to(torch.int32).reshape(1) * nat.FLAG_OUTSIDE_UNIT  # AUROC's sigmoid rule
            auc_stats = torch.ops.manner_b200.pooled_auc(preds, labels, 2, outside)
        s = sums.cpu().numpy()
        f = int(flags.cpu().item())
        if f & (nat.FLAG_CAND_OVERFLOW | nat.FLAG_BAD_ASPECT):
            raise nat.NativeError(f"manner_b200 kernels flagged bad input (flags={f}): 8 = aspect label outside [0, num_classes)")
        n = max(off.numel() - 1, 1)
        out: Dict[str, float] = {}
        for slot, key in SLOT_KEYS.items():
            if slot >= nat.M_CATEG_DIV_K0 and not has_aspects:
                continue
            out[self.prefix + key.format(k0=self.ks[0], k1=self.ks[1])] = float(s[slot] / n)
        out[self.prefix + "gauc"] = float(s[nat.M_GAUC] / s[nat.M_GAUC_VALID]) if s[nat.M_GAUC_VALID] > 0 else 0.0
        if auc_stats is not None:
            out[self.prefix + "auc"] = float(auc_stats.cpu()[0])
        return out
