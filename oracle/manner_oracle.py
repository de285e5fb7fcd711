"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MANNeR scoring / ensemble / metrics hot path.

A plain-torch restatement of what andreeaiana/manner computes between "news embeddings of the
batch" and "logged test metrics", with the PLM news encoder replaced by a lookup in a cached
embedding table (the cache is the new boundary; parity is defined given an identical
table).  Each function cites the reference lines it follows.

Pinning status
--------------
* In-repo arithmetic (CRModule.forward / model_step / on_test_epoch_end, EnsembleModule.forward /
  _submodel_forward / model_step / on_test_epoch_end, DotProduct, metrics/functional.py,
  Diversity / Personalization / CustomRetrievalMetric.compute): PINNED -- tests/test_oracle.py
  checks this restatement against golden vectors produced by executing the reference's own
  unmodified code in this container (tests/golden/make_golden.py, via oracle/ref_stubs.py).
* Third-party arithmetic (torchmetrics 0.11.4, pyg to_dense_batch): PARITY UNPINNED -- the
  reference ships no tests or fixtures and the packages are absent; oracle/thirdparty.py
  restates their published algorithms and is cross-checked against sklearn / scipy / hand-computed
  known answers only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; it is the checker, never the product.  manner_b200/ must not import it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import thirdparty as tp

STEP_BATCH = 8  # configs/data/mind_rec.yaml:51 (eval batch_size)


@dataclass
class Behaviours:
    """Ragged impressions in CSR form (what MINDCollate's segment-id vectors encode,
    mind_rec_dataset.py:114-132,171-174).  All arrays are host numpy."""

    hist_offsets: np.ndarray  # int32 [B+1]
    hist_ids: np.ndarray  # int32 [sum H]   row of the embedding table
    cand_offsets: np.ndarray  # int32 [B+1]
    cand_ids: np.ndarray  # int32 [sum C]
    labels: np.ndarray  # uint8 [sum C]

    @property
    def n_impressions(self) -> int:
        return int(self.hist_offsets.shape[0] - 1)


def step_batch(bhv: Behaviours, lo: int, hi: int, aspects: Optional[Dict[str, np.ndarray]] = None) -> Dict:
    """Impressions [lo, hi) as the reference's MINDRecBatch dict (mind_batch.py:6-12).  ``x_hist`` /
    ``x_cand`` carry the table row ids under "news_row" instead of tokenised text."""
    h0, h1 = int(bhv.hist_offsets[lo]), int(bhv.hist_offsets[hi])
    c0, c1 = int(bhv.cand_offsets[lo]), int(bhv.cand_offsets[hi])
    hs = torch.from_numpy(np.diff(bhv.hist_offsets[lo : hi + 1]).astype(np.int64))
    cs = torch.from_numpy(np.diff(bhv.cand_offsets[lo : hi + 1]).astype(np.int64))
    hist_rows = torch.from_numpy(bhv.hist_ids[h0:h1].astype(np.int64))
    cand_rows = torch.from_numpy(bhv.cand_ids[c0:c1].astype(np.int64))
    x_hist: Dict[str, Tensor] = {"news_row": hist_rows}
    x_cand: Dict[str, Tensor] = {"news_row": cand_rows}
    if aspects is not None:
        for key, per_news in aspects.items():  # "category" / "sentiment" int64 label per table row
            lab = torch.from_numpy(np.asarray(per_news).astype(np.int64))
            x_hist[key], x_cand[key] = lab[hist_rows], lab[cand_rows]
    return {
        "batch_hist": torch.repeat_interleave(torch.arange(hi - lo), hs),  # _make_batch_assignees :171-174
        "batch_cand": torch.repeat_interleave(torch.arange(hi - lo), cs),
        "x_hist": x_hist,
        "x_cand": x_cand,
        "labels": torch.from_numpy(bhv.labels[c0:c1].astype(np.float32)),  # MINDCollate :122 (.float())
        "users": torch.arange(lo, hi),
    }


# ----------------------------------------------------------------------------------------------
# scoring
# ----------------------------------------------------------------------------------------------


def dot_product(user: Tensor, cand: Tensor) -> Tensor:
    """click_predictors.py:9-12."""
    return torch.bmm(user, cand).squeeze(1)


def late_fusion_scores(table: Tensor, batch: Dict) -> Tuple[Tensor, Tensor]:
    """cr_module.py:105-131 (late_fusion=True) == ensemble_module.py:114-135: gather, pad, count the
    history with a per-row loop, mean-pool by true division, bmm.  Returns scores [B, Cmax], mask."""
    hist = table[batch["x_hist"]["news_row"]]
    hist_dense, mask_hist = tp.to_dense_batch(hist, batch["batch_hist"])
    cand = table[batch["x_cand"]["news_row"]]
    cand_dense, mask_cand = tp.to_dense_batch(cand, batch["batch_cand"])
    hist_size = torch.tensor([torch.where(mask_hist[i])[0].shape[0] for i in range(mask_hist.shape[0])])
    user = torch.div(hist_dense.sum(dim=1), hist_size.unsqueeze(dim=-1))
    scores = dot_product(user.unsqueeze(dim=1), cand_dense.permute(0, 2, 1))
    return scores, mask_cand


@dataclass
class Attention:
    """Parameters of the reference's early-fusion user encoder: NAMLUserEncoder -> AdditiveAttention
    (user_encoder.py:9-21, attention.py:6-13): linear.weight [Q, D], linear.bias [Q], query [Q]."""

    weight: Tensor
    bias: Tensor
    query: Tensor


def additive_attention(att: Attention, x: Tensor) -> Tensor:
    """attention.py:15-29 on a padded batch [B, Hmax, D].  NOTE the reference does not mask the padded
    (all-zero) history rows: they enter the softmax with the logit query . tanh(bias)."""
    a = torch.tanh(torch.nn.functional.linear(x, att.weight, att.bias))
    w = torch.nn.functional.softmax(torch.matmul(a, att.query), dim=1)
    return torch.bmm(w.unsqueeze(dim=1), x).squeeze(dim=1)


def early_fusion_scores(table: Tensor, att: Attention, batch: Dict) -> Tuple[Tensor, Tensor]:
    """cr_module.py:105-131 with late_fusion=False: user_vector = self.user_encoder(clicked_news_vector_agg)."""
    hist = table[batch["x_hist"]["news_row"]]
    hist_dense, _ = tp.to_dense_batch(hist, batch["batch_hist"])
    cand = table[batch["x_cand"]["news_row"]]
    cand_dense, mask_cand = tp.to_dense_batch(cand, batch["batch_cand"])
    user = additive_attention(att, hist_dense)
    scores = dot_product(user.unsqueeze(dim=1), cand_dense.permute(0, 2, 1))
    return scores, mask_cand


def attention_logits(att: Attention, table: Tensor) -> Tensor:
    """Per-news additive-attention logit  query . tanh(W x_n + b)  -- the quantity the B200 path caches next
    to the embedding table (it depends on the news row only, not on the user)."""
    return torch.matmul(torch.tanh(torch.nn.functional.linear(table, att.weight, att.bias)), att.query)


def ce_step_loss(scores: Tensor, batch: Dict) -> Tensor:
    """cr_module.py:140-142,171: CrossEntropyLoss()(scores [B, Cmax], y_true [B, Cmax]) with the 0/1 label
    matrix as class probabilities -- padded columns (score exactly 0) are part of the log-softmax -- mean over
    the step's impressions."""
    y_true, _ = tp.to_dense_batch(batch["labels"], batch["batch_cand"])
    return torch.nn.functional.cross_entropy(scores, y_true)


def supcon_step_loss(scores: Tensor, batch: Dict, temperature: float) -> Tensor:
    """cr_module.py:144-169 + components/losses.py:6-40 on pytorch_metric_learning 2.1.1's SupConLoss: per impression with
    >= 1 positive, -mean over positives of (s_p / T - logsumexp over the impression's real candidates of s / T) (the row
    maximum that is subtracted first runs over the DENSE row, pads included); AvgNonZeroReducer averages the impressions whose
    loss is > 0.  Step-level guards: at most one positive and at most one negative pair in the whole step
    (`all(len(x) <= 1 for x in indices_tuple)`, losses.py:15-16), or no positive / no negative anywhere (:22), give 0.
    Pinned on tests/golden/cr_supcon_*.npz, which the reference's own SupConLoss produced (oracle/ref_stubs.py)."""
    y_true, mask_cand = tp.to_dense_batch(batch["labels"], batch["batch_cand"])
    pos = (y_true > 0) & mask_cand
    neg = (y_true == 0) & mask_cand
    n_pos, n_neg = int(pos.sum()), int(neg.sum())
    if (n_pos <= 1 and n_neg <= 1) or n_pos == 0 or n_neg == 0:
        return torch.zeros(())
    mat = scores / temperature
    mat = mat - mat.max(dim=1, keepdim=True)[0]
    denom = torch.logsumexp(mat.masked_fill(~mask_cand, torch.finfo(mat.dtype).min), dim=1, keepdim=True)
    log_prob = mat - denom
    mean_log_prob_pos = (pos * log_prob).sum(dim=1) / (pos.sum(dim=1) + torch.finfo(torch.float32).tiny)
    losses = -mean_log_prob_pos
    nz = losses > 0
    return losses[nz].mean() if bool(nz.any()) else torch.zeros(())


def zscore(scores: Tensor, mask_cand: Tensor) -> Tensor:
    """ensemble_module.py:137-149: unbiased std over the valid columns, mean = row sum over ALL
    (padded) columns / candidate count."""
    cand_size = torch.tensor([torch.where(mask_cand[i])[0].shape[0] for i in range(mask_cand.shape[0])])
    std = torch.stack([torch.std(scores[i][mask_cand[i]]) for i in range(mask_cand.shape[0])]).unsqueeze(-1)
    mean = torch.div(torch.sum(scores, dim=1), cand_size).unsqueeze(-1).expand_as(scores)
    return torch.div(scores - mean, std)


def ensemble_scores(tables: Sequence[Tensor], weights: Sequence[float], batch: Dict) -> Tuple[Tensor, Tensor]:
    """ensemble_module.py:95-109.  tables[0] is the CR-Module table (weight fixed at 1); tables[m>=1]
    are A-Module tables combined as ``scores += w * z`` in order, skipped when w == 0."""
    s, mask = late_fusion_scores(tables[0], batch)
    scores = zscore(s, mask)
    for m in range(1, len(tables)):
        w = weights[m]
        if w != 0:
            sm, mk = late_fusion_scores(tables[m], batch)
            scores += w * zscore(sm, mk)
    return scores, mask


def flatten_for_metrics(scores: Tensor, batch: Dict) -> Tuple[Tensor, Tensor, Tensor]:
    """cr_module.py:142,173-182 == ensemble_module.py:155,166-175."""
    y_true, mask_cand = tp.to_dense_batch(batch["labels"], batch["batch_cand"])
    n = mask_cand.shape[0]
    preds = torch.cat([scores[i][mask_cand[i]] for i in range(n)], dim=0).detach()
    targets = torch.cat([y_true[i][mask_cand[i]] for i in range(n)], dim=0).long()
    cand_news_size = torch.tensor([torch.where(mask_cand[i])[0].shape[0] for i in range(n)])
    return preds, targets, cand_news_size


# ----------------------------------------------------------------------------------------------
# in-repo metrics (manner/metrics/functional.py)
# ----------------------------------------------------------------------------------------------


def diversity(preds: Tensor, target: Tensor, num_classes: int, k: Optional[int] = None) -> Tensor:
    """metrics/functional.py:8-28: normalised entropy of the top-k aspect histogram."""
    preds, target = tp._check_retrieval_functional_inputs(preds, target, allow_non_binary_target=True)
    k = preds.shape[-1] if k is None else k
    top = target[tp.stable_desc_argsort(preds)][:k]
    count = torch.bincount(top)
    count = torch.nn.functional.pad(count, pad=(0, num_classes - count.shape[0]))
    dist = torch.distributions.Categorical(count / count.shape[0])
    return torch.div(dist.entropy(), torch.log(torch.tensor(num_classes)))


def generalized_jaccard(pred: Tensor, target: Tensor) -> Tensor:
    """metrics/functional.py:65-70."""
    return torch.min(pred, target).sum(dim=0) / torch.max(pred, target).sum(dim=0)


def personalization(preds: Tensor, cand_aspects: Tensor, hist_aspects: Tensor, num_classes: int, k: Optional[int] = None) -> Tensor:
    """metrics/functional.py:31-62: generalised Jaccard of top-k candidate aspects vs history aspects."""
    preds, cand_aspects = tp._check_retrieval_functional_inputs(preds, cand_aspects, allow_non_binary_target=True)
    k = preds.shape[-1] if k is None else k
    top = cand_aspects[tp.stable_desc_argsort(preds)][:k]
    pc = torch.bincount(top)
    pc = torch.nn.functional.pad(pc, pad=(0, num_classes - pc.shape[0]))
    hc = torch.bincount(hist_aspects)
    hc = torch.nn.functional.pad(hc, pad=(0, num_classes - hc.shape[0]))
    return generalized_jaccard(pc, hc)


def _group_mean(values: List[Tensor], like: Tensor) -> Tensor:
    """metrics/base.py:127-129: fp32 mean of the per-impression values."""
    return torch.stack([v.to(like) for v in values]).mean() if values else torch.tensor(0.0).to(like)


def diversity_epoch(preds: Tensor, aspects: Tensor, sizes: Sequence[int], num_classes: int, k: int) -> Tensor:
    """metrics/diversity.py:8-34 on top of RetrievalMetric.compute: impressions whose aspect labels
    sum to 0 score 0.0 (empty_target_action='neg')."""
    out = []
    for p, a in zip(torch.split(preds.float(), list(sizes)), torch.split(aspects.long(), list(sizes))):
        out.append(torch.tensor(0.0) if not a.sum() else diversity(p, a, num_classes, k))
    return _group_mean(out, preds.float())


def personalization_epoch(
    preds: Tensor, cand_aspects: Tensor, hist_aspects: Tensor, cand_sizes: Sequence[int], hist_sizes: Sequence[int], num_classes: int, k: int
) -> Tensor:
    """metrics/personalization.py:8-38 on top of metrics/base.py:92-129."""
    out = []
    for p, a, h in zip(
        torch.split(preds.float(), list(cand_sizes)),
        torch.split(cand_aspects.long(), list(cand_sizes)),
        torch.split(hist_aspects.long(), list(hist_sizes)),
    ):
        out.append(torch.tensor(0.0) if not a.sum() else personalization(p, a, h, num_classes, k))
    return _group_mean(out, preds.float())


# ----------------------------------------------------------------------------------------------
# gAUC -- NOT in the reference; this definition is the spec for the new metric.
# ----------------------------------------------------------------------------------------------


def gauc_per_impression(preds: np.ndarray, labels: np.ndarray) -> Tuple[float, bool]:
    pos = preds[labels != 0]
    neg = preds[labels == 0]
    if pos.size == 0 or neg.size == 0:
        return 0.0, False
    lt = (neg[None, :] < pos[:, None]).sum()
    eq = (neg[None, :] == pos[:, None]).sum()
    return float((np.float64(lt) + 0.5 * np.float64(eq)) / (np.float64(pos.size) * np.float64(neg.size))), True


# ----------------------------------------------------------------------------------------------
# epoch drivers (reference-faithful: steps of 8, per-row loops, metric objects)
# ----------------------------------------------------------------------------------------------


def _recommendation_metrics(with_auc_mrr: bool) -> tp.MetricCollection:
    metrics: Dict[str, tp.Metric] = {}
    if with_auc_mrr:  # cr_module.py:79-86
        metrics["auc"] = tp.AUROC(task="binary", num_classes=2)
        metrics["mrr"] = tp.RetrievalMRR()
    metrics["ndcg@5"] = tp.RetrievalNormalizedDCG(k=5)  # ensemble_module.py:50-55 has only these two
    metrics["ndcg@10"] = tp.RetrievalNormalizedDCG(k=10)
    return tp.MetricCollection(metrics).clone(prefix="test/")


def cr_eval_epoch(table: Tensor, bhv: Behaviours, step: int = STEP_BATCH, double_compute: bool = False, attention: Optional[Attention] = None,
                  supcon_temperature: Optional[float] = None) -> Dict:
    """CRModule test epoch: test_step per batch (cr_module.py:253-264) then on_test_epoch_end
    (:266-274).  ``double_compute`` repeats the metric pass the way Lightning does (forward +
    compute at log time) -- timing only, the values are identical.  ``attention`` switches to early fusion
    (late_fusion=False); ``test/loss`` is the MeanMetric over the steps' losses (cr_module.py:255-259): cross
    entropy, or the SupCon restatement when ``supcon_temperature`` is given."""
    preds_l, targets_l, sizes_l, losses = [], [], [], []
    for lo in range(0, bhv.n_impressions, step):
        batch = step_batch(bhv, lo, min(lo + step, bhv.n_impressions))
        scores, _ = late_fusion_scores(table, batch) if attention is None else early_fusion_scores(table, attention, batch)
        losses.append(ce_step_loss(scores, batch) if supcon_temperature is None else supcon_step_loss(scores, batch, supcon_temperature))
        p, t, s = flatten_for_metrics(scores, batch)
        preds_l.append(p), targets_l.append(t), sizes_l.append(s)
    preds, targets, sizes = torch.cat(preds_l), torch.cat(targets_l), torch.cat(sizes_l)
    indexes = torch.arange(sizes.shape[0]).repeat_interleave(sizes)
    coll = _recommendation_metrics(with_auc_mrr=True)
    values = coll(preds, targets, **{"indexes": indexes})
    if double_compute:
        values = coll.compute()
    out = {k: float(v) for k, v in values.items()}
    out.update(gauc_epoch(preds.numpy(), targets.numpy(), sizes.numpy()))
    mean_loss = tp.MeanMetric()
    for l in losses:
        mean_loss.update(l)
    out["test/loss"] = float(mean_loss.compute())
    return {"scores": preds.numpy(), "targets": targets.numpy(), "cand_news_size": sizes.numpy(), "metrics": out,
            "step_losses": np.asarray([float(l) for l in losses], dtype=np.float32)}


def ensemble_eval_epoch(
    tables: Sequence[Tensor],
    weights: Sequence[float],
    bhv: Behaviours,
    aspects: Optional[Dict[str, np.ndarray]] = None,
    num_classes: Optional[Dict[str, int]] = None,
    step: int = STEP_BATCH,
    double_compute: bool = False,
    reference_only: bool = False,
) -> Dict:
    """EnsembleModule test epoch (ensemble_module.py:202-256).  ``reference_only`` skips the two extras
    the reference does not compute (mrr, gauc) -- used when this function is the timed CPU baseline.  ``aspects`` maps "category" /
    "sentiment" to per-news int labels; when given, the diversity / personalization keys of
    ensemble_module.py:231-238 are produced as well."""
    acc: Dict[str, List[Tensor]] = {k: [] for k in ("preds", "targets", "cs", "hs", "tc", "ts", "hc", "hsent")}
    for lo in range(0, bhv.n_impressions, step):
        batch = step_batch(bhv, lo, min(lo + step, bhv.n_impressions), aspects)
        scores, _ = ensemble_scores(tables, weights, batch)
        p, t, s = flatten_for_metrics(scores, batch)
        acc["preds"].append(p), acc["targets"].append(t), acc["cs"].append(s)
        _, mask_hist = tp.to_dense_batch(batch["x_hist"]["news_row"], batch["batch_hist"])
        acc["hs"].append(torch.tensor([torch.where(mask_hist[n])[0].shape[0] for n in range(mask_hist.shape[0])]))
        if aspects is not None:
            # ensemble_module.py:157-192 flattens what it just padded: the values are the inputs
            acc["tc"].append(batch["x_cand"]["category"].long()), acc["ts"].append(batch["x_cand"]["sentiment"].long())
            acc["hc"].append(batch["x_hist"]["category"].long()), acc["hsent"].append(batch["x_hist"]["sentiment"].long())
    preds, targets = torch.cat(acc["preds"]), torch.cat(acc["targets"])
    cs, hs = torch.cat(acc["cs"]), torch.cat(acc["hs"])
    indexes = torch.arange(cs.shape[0]).repeat_interleave(cs)
    coll = _recommendation_metrics(with_auc_mrr=False)
    values = coll(preds, targets, **{"indexes": indexes})
    if double_compute:
        values = coll.compute()
    out = {k: float(v) for k, v in values.items()}
    if aspects is not None:
        nc = num_classes or {"category": 19, "sentiment": 4}  # configs/model/ensemble_module.yaml:8-9
        tc, ts, hc, hsent = (torch.cat(acc[k]) for k in ("tc", "ts", "hc", "hsent"))
        for k in (5, 10):
            out[f"test/categ_div@{k}"] = float(diversity_epoch(preds, tc, cs.tolist(), nc["category"], k))
            out[f"test/sent_div@{k}"] = float(diversity_epoch(preds, ts, cs.tolist(), nc["sentiment"], k))
            out[f"test/categ_pers@{k}"] = float(personalization_epoch(preds, tc, hc, cs.tolist(), hs.tolist(), nc["category"], k))
            out[f"test/sent_pers@{k}"] = float(personalization_epoch(preds, ts, hsent, cs.tolist(), hs.tolist(), nc["sentiment"], k))
    if reference_only:
        return {"scores": preds.numpy(), "targets": targets.numpy(), "cand_news_size": cs.numpy(), "metrics": out}
    # extras the reference's EnsembleModule does not log but the B200 path reports for every call
    extra = tp.MetricCollection({"mrr": tp.RetrievalMRR()}).clone(prefix="test/")
    out.update({k: float(v) for k, v in extra(preds, targets, **{"indexes": indexes}).items()})
    out.update(gauc_epoch(preds.numpy(), targets.numpy(), cs.numpy()))
    return {"scores": preds.numpy(), "targets": targets.numpy(), "cand_news_size": cs.numpy(), "metrics": out}


def gauc_epoch(preds: np.ndarray, targets: np.ndarray, sizes: np.ndarray) -> Dict[str, float]:
    total, n, start = 0.0, 0, 0
    for c in sizes.tolist():
        v, ok = gauc_per_impression(preds[start : start + c], targets[start : start + c])
        start += c
        if ok:
            total, n = total + v, n + 1
    return {"test/gauc": total / n if n else 0.0, "gauc_impressions": float(n)}


# ----------------------------------------------------------------------------------------------
# metrics from flat scores, vectorised (for the full-size GPU parity tests)
# ----------------------------------------------------------------------------------------------


def stable_ranks(scores: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """1-based rank of every candidate inside its impression under the canonical rule A3
    (descending score, lower position first on ties; NaN sorts first as in torch)."""
    scores = np.asarray(scores, dtype=np.float32)
    n = scores.shape[0]
    seg = np.repeat(np.arange(offsets.shape[0] - 1), np.diff(offsets))
    key = -scores.astype(np.float64)
    key[np.isnan(scores)] = -np.inf
    order = np.lexsort((np.arange(n), key, seg))  # primary seg, then key asc (= score desc), then position
    ranks = np.empty(n, dtype=np.int64)
    ranks[order] = np.arange(n) - offsets[:-1].astype(np.int64)[seg[order]] + 1
    return ranks


def aten_sum_f32(x: np.ndarray) -> np.ndarray:
    """Row sums of fp32 ``x`` [n, L] in the order ATen's CPU sum kernel uses for a contiguous inner
    reduction (SumKernel.cpp: 8-lane vectors with 4-way ILP over vectors, scalar tail first, then the
    lanes in order; L < 8 takes the scalar 4-way ILP path).  Probed bit-exact against ``Tensor.sum``
    of this image's torch for L in 1..300; lets the vectorised metrics reproduce ``_dcg`` exactly."""
    x = np.asarray(x, dtype=np.float32)
    n, ell = x.shape
    f = np.float32

    def ilp4(rows):  # rows: list of [n, w] arrays, summed over the list
        r = len(rows) // 4
        p = [np.zeros_like(rows[0]) if rows else np.zeros((n, 1), dtype=f) for _ in range(4)]
        for i in range(r):
            for k in range(4):
                p[k] = (p[k] + rows[4 * i + k]).astype(f)
        for i in range(4 * r, len(rows)):
            p[0] = (p[0] + rows[i]).astype(f)
        for k in range(1, 4):
            p[0] = (p[0] + p[k]).astype(f)
        return p[0]

    if ell == 0:
        return np.zeros(n, dtype=f)
    if ell < 8:
        return ilp4([x[:, j : j + 1] for j in range(ell)])[:, 0]
    vs = ell // 8
    acc = ilp4([x[:, 8 * i : 8 * i + 8] for i in range(vs)])
    fin = np.zeros(n, dtype=f)
    for j in range(8 * vs, ell):
        fin = (fin + x[:, j]).astype(f)
    for lane in range(8):
        fin = (fin + acc[:, lane]).astype(f)
    return fin


_DISC = torch.log2(torch.arange(64) + 2.0).numpy()  # fp32, the denominators torchmetrics' _dcg uses


def per_impression_metrics(scores: np.ndarray, labels: np.ndarray, offsets: np.ndarray, ks: Tuple[int, int] = (5, 10)) -> np.ndarray:
    """[B, 5] fp32: mrr, ndcg@k0, ndcg@k1, gauc, gauc_valid -- same definitions as the metric
    objects above, evaluated without the per-impression Python loop.  Checked against the faithful
    loop in tests/test_oracle.py."""
    offsets = offsets.astype(np.int64)
    nb = offsets.shape[0] - 1
    ranks = stable_ranks(scores, offsets)
    seg = np.repeat(np.arange(nb), np.diff(offsets))
    lab = labels.astype(np.int64)
    out = np.zeros((nb, 5), dtype=np.float32)
    npos = np.bincount(seg, weights=lab, minlength=nb).astype(np.int64)
    ncand = np.diff(offsets)
    posmask = lab != 0
    # mrr: first hit
    best = np.full(nb, np.iinfo(np.int64).max)
    np.minimum.at(best, seg[posmask], ranks[posmask])
    has = npos > 0
    out[has, 0] = (np.float32(1.0) / best[has].astype(np.float32)).astype(np.float32)
    # ndcg: (target / denom).sum() over the first L = min(k, C) ranks, in ATen's fp32 summation order
    for col, k in ((1, ks[0]), (2, ks[1])):
        hit = np.zeros((nb, k), dtype=bool)
        sel = posmask & (ranks <= k)
        hit[seg[sel], ranks[sel] - 1] = True
        inv = (np.float32(1.0) / _DISC[:k]).astype(np.float32)
        terms = np.where(hit, inv[None, :], np.float32(0.0)).astype(np.float32)
        ideal = np.where(np.arange(k)[None, :] < npos[:, None], inv[None, :], np.float32(0.0)).astype(np.float32)
        length = np.minimum(k, ncand)
        dcg, idcg = np.zeros(nb, dtype=np.float32), np.zeros(nb, dtype=np.float32)
        for ell in np.unique(length):
            rows = length == ell
            dcg[rows] = aten_sum_f32(terms[rows][:, :ell])
            idcg[rows] = aten_sum_f32(ideal[rows][:, :ell])
        ok = idcg > 0
        out[ok, col] = (dcg[ok] / idcg[ok]).astype(np.float32)
    # gauc
    for i in np.nonzero((npos > 0) & (npos < ncand))[0]:
        v, _ = gauc_per_impression(np.asarray(scores[offsets[i] : offsets[i + 1]], dtype=np.float32), lab[offsets[i] : offsets[i + 1]])
        out[i, 3], out[i, 4] = np.float32(v), 1.0
    return out


def pooled_auc_exact(preds: np.ndarray, labels: np.ndarray, sigmoid: Optional[bool] = None) -> float:
    """Rank-statistic form of A6 in fp64 on the fp32 keys torchmetrics would sort: the value the
    fp32 trapezoid approximates.  ``sigmoid=None`` applies the reference's any-outside-[0,1] rule."""
    p = torch.from_numpy(np.asarray(preds, dtype=np.float32))
    if sigmoid is None:
        sigmoid = not bool(torch.all((p >= 0) * (p <= 1)))
    if sigmoid:
        p = p.sigmoid()
    p = p.numpy()
    pos = np.sort(p[labels != 0])
    neg = np.sort(p[labels == 0])
    if pos.size == 0 or neg.size == 0:
        return 0.0
    lo = np.searchsorted(neg, pos, side="left").astype(np.float64)
    hi = np.searchsorted(neg, pos, side="right").astype(np.float64)
    return float((lo + hi).sum() / (2.0 * pos.size * neg.size))


# ----------------------------------------------------------------------------------------------
# full-size score check: the same arithmetic in fp64, vectorised over impressions (no per-step padding -- the
# padded cells of the reference's dense batches contribute exact zeros to every sum it takes)
# ----------------------------------------------------------------------------------------------


def ensemble_truth_f64(tables: Sequence[Tensor], weights: Sequence[float], bhv: Behaviours, zscore_modules: bool = True,
                       rtol: float = 1e-5, chunk: int = 4096) -> Tuple[np.ndarray, np.ndarray]:
    """(scores fp64 [sum C], tol fp64 [sum C]): cr_module.py:105-131 / ensemble_module.py:95-151 evaluated in fp64, and the
    tolerance of the stated parity bar: raw dot products |ds| <= rtol * sum_i |u_i c_i| (condition-aware);
    z-scored module scores 2 rtol (1 + |z|); the weighting adds them up.  An fp32 evaluation in any summation order (the
    reference's bmm, this library's warp reduction) must lie within ``tol`` of these scores."""
    ho, co = bhv.hist_offsets.astype(np.int64), bhv.cand_offsets.astype(np.int64)
    n_impr = ho.shape[0] - 1
    out = np.zeros(int(co[-1]), dtype=np.float64)
    tol = np.zeros(int(co[-1]), dtype=np.float64)
    for lo in range(0, n_impr, chunk):
        hi = min(lo + chunk, n_impr)
        hids = torch.from_numpy(bhv.hist_ids[ho[lo]:ho[hi]].astype(np.int64))
        cids = torch.from_numpy(bhv.cand_ids[co[lo]:co[hi]].astype(np.int64))
        hseg = torch.from_numpy(np.repeat(np.arange(hi - lo), np.diff(ho[lo:hi + 1])))
        cseg = torch.from_numpy(np.repeat(np.arange(hi - lo), np.diff(co[lo:hi + 1])))
        hlen = torch.from_numpy(np.diff(ho[lo:hi + 1]).astype(np.float64))
        clen = torch.from_numpy(np.diff(co[lo:hi + 1]).astype(np.float64))
        total = torch.zeros(cids.shape[0], dtype=torch.float64)
        total_tol = torch.zeros(cids.shape[0], dtype=torch.float64)
        for m, (table, w) in enumerate(zip(tables, weights)):
            if m > 0 and w == 0:
                continue
            t = table.float()
            hrows = t[hids].double()
            u = torch.zeros(hi - lo, t.shape[1], dtype=torch.float64).index_add_(0, hseg, hrows) / hlen[:, None]
            ua = torch.zeros(hi - lo, t.shape[1], dtype=torch.float64).index_add_(0, hseg, hrows.abs()) / hlen[:, None]
            crows = t[cids].double()
            sc = (crows * u[cseg]).sum(1)
            st = rtol * (crows.abs() * ua[cseg]).sum(1) + 1e-30
            if zscore_modules:
                mean = torch.zeros(hi - lo, dtype=torch.float64).index_add_(0, cseg, sc) / clen
                var = torch.zeros(hi - lo, dtype=torch.float64).index_add_(0, cseg, (sc - mean[cseg]) ** 2) / (clen - 1)
                std = var.sqrt()
                z = (sc - mean[cseg]) / std[cseg]
                st = 2.0 * rtol * (1.0 + z.abs())  # z-scores are O(1): the stated 1e-5 relative bar on each side (the propagated bound is ~100x looser)
                sc = z
            total += float(w) * sc if m > 0 else (sc if w == 1 else float(w) * sc)
            total_tol += abs(float(w)) * st
        out[co[lo]:co[hi]] = total.numpy()
        tol[co[lo]:co[hi]] = total_tol.numpy()
    return out, tol


def unexplained_rank_flips(got: np.ndarray, ref: np.ndarray, slack: np.ndarray, offsets: np.ndarray) -> Tuple[int, int, float]:
    """Candidates ranked differently under ``got`` and ``ref`` scores: (flips, unexplained, largest gap among the explained).
    A flip of candidate j is *explained* when another candidate of its impression lies within slack_j + slack_k of it in
    ``ref`` -- a near-tie that two correct fp32 evaluations may order differently; anything else is a real ranking error."""
    ra, rb = stable_ranks(got, offsets), stable_ranks(ref, offsets)
    bad = np.nonzero(ra != rb)[0]
    if bad.size == 0:
        return 0, 0, 0.0
    seg = np.searchsorted(offsets.astype(np.int64), bad, side="right") - 1
    unexplained, worst = 0, 0.0
    for j, i in zip(bad.tolist(), seg.tolist()):
        lo, hi = int(offsets[i]), int(offsets[i + 1])
        gap = np.abs(ref[lo:hi] - ref[j]) - (slack[lo:hi] + slack[j])
        gap[j - lo] = np.inf
        k = int(np.argmin(gap))
        if gap[k] <= 0:
            worst = max(worst, float(abs(ref[lo + k] - ref[j])))
        else:
            unexplained += 1
    return int(bad.size), unexplained, worst


# ---- full-catalogue retrieval (BASELINE.json configs[4]; no reference counterpart) -----------------------


def pooled_users(table: Tensor, hist_offsets: np.ndarray, hist_ids: np.ndarray) -> Tensor:
    """fp32 [U, D]: late-fusion user vectors, `sum(dim=1) / hist_size` of cr_module.py:116-123 (history rows
    added in history order, true division)."""
    t = table.float()
    out = torch.zeros((len(hist_offsets) - 1, t.shape[1]), dtype=torch.float32)
    for u in range(out.shape[0]):
        rows = t[torch.from_numpy(np.asarray(hist_ids[hist_offsets[u] : hist_offsets[u + 1]], dtype=np.int64))]
        acc = torch.zeros(t.shape[1], dtype=torch.float32)
        for r in rows:
            acc = acc + r
        out[u] = acc / float(rows.shape[0])
    return out


def retrieval_scores(users_bf16: Tensor, catalog_bf16: Tensor) -> Tensor:
    """fp32 [U, N] = bf16 x bf16 products (exact in fp32) accumulated in fp32 -- what a bf16 tensor-core
    contraction with fp32 accumulation computes, up to summation order (DotProduct, click_predictors.py:9-12,
    applied to every catalogue row instead of an impression's candidates)."""
    return users_bf16.float() @ catalog_bf16.float().T


def topk_select(scores: np.ndarray, k: int, id_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Exact selection rule of the retrieval path on a given fp32 score matrix [U, N]: the k best per row,
    score descending, id ascending on ties (the A3 ranking rule); unused slots (k > N) = (-inf, -1)."""
    u, n = scores.shape
    order = np.argsort(-scores.astype(np.float64), axis=1, kind="stable")[:, :k]  # stable: lower id first on ties
    top_s = np.full((u, k), -np.inf, dtype=np.float32)
    top_i = np.full((u, k), -1, dtype=np.int64)
    kk = min(k, n)
    top_s[:, :kk] = np.take_along_axis(scores, order, axis=1)[:, :kk]
    top_i[:, :kk] = order[:, :kk] + id_offset
    return top_s, top_i


def rank_key(scores: np.ndarray) -> np.ndarray:
    """The total order the device code ranks candidates by (csrc/score_eval.cu: rank_key), restated: an unsigned key per fp32 score
    with  key(a) > key(b)  <=>  a ranks before b in ``torch.sort(descending=True)``  and  key(a) == key(b)  <=>  they tie there
    (every NaN above everything and equal to every other NaN, -0 == +0).  ``np.lexsort((index, -key))`` must therefore be
    ``torch.argsort(scores, descending=True, stable=True)`` -- the order behind MRR / nDCG / Diversity / Personalization
    (torchmetrics 0.11.4 retrieval metrics; metrics/functional.py:17,47)."""
    x = np.asarray(scores, dtype=np.float32).copy()
    x[x == 0.0] = 0.0  # -0 -> +0
    b = x.view(np.uint32).astype(np.uint64)
    key = np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)
    key[np.isnan(x)] = 0xFFFFFFFF
    return key.astype(np.uint64)
