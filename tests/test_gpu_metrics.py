"""GPU parity tests of the standalone metric objects (manner_b200/metrics.py, mb200_rank_metrics): the torchmetrics
`update(preds, target, indexes)` / `compute()` seam every module of the reference ends its epoch with, against the values the
reference's own CRModule / EnsembleModule logged (tests/golden) and against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import manner_oracle as mo  # noqa: E402  (checker only)

from manner_b200 import _native as nat  # noqa: E402

METRIC_ATOL = 1e-6


@pytest.fixture(scope="module")
def metrics_mod():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200 import metrics

    return metrics


@pytest.mark.parametrize("name", ["cr_d128", "cr_d768", "cr_ties", "cr_ef_d128"])
def test_metric_collection_on_the_references_preds(golden_dir, metrics_mod, name):
    """preds / targets / cand_news_size exactly as CRModule.on_test_epoch_end assembles them (cr_module.py:266-271) -> the
    values its MetricCollection logged."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    preds, targets = torch.from_numpy(z["preds"]).cuda(), torch.from_numpy(z["targets"]).cuda()
    sizes = torch.from_numpy(z["cand_news_size"])
    indexes = torch.arange(sizes.shape[0]).repeat_interleave(sizes).cuda()
    m = metrics_mod.RetrievalMetricsB200(prefix="test/")
    out = m(preds, targets, indexes=indexes)
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(out["test/" + k] - float(z["test_" + k])) <= METRIC_ATOL, (k, out["test/" + k], float(z["test_" + k]))
    # per-impression values bit-exact against the oracle on the same predictions
    off = torch.zeros(sizes.shape[0] + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(sizes, 0)
    _, per, _ = metrics_mod.rank_metrics(preds, (targets != 0).to(torch.uint8), off.cuda(), int(sizes.max()), want_per_impression=True)
    ref = mo.per_impression_metrics(z["preds"], z["targets"].astype(np.uint8), off.numpy())
    np.testing.assert_array_equal(per.cpu().numpy()[:, :3], ref[:, :3])
    # updates in several pieces and in shuffled row order give the same result (torchmetrics sorts by index)
    m.reset()
    g = torch.Generator().manual_seed(1)
    perm = torch.randperm(preds.numel(), generator=g).cuda()
    half = preds.numel() // 2
    m.update(preds[perm[:half]], targets[perm[:half]], indexes=indexes[perm[:half]])
    m.update(preds[perm[half:]], targets[perm[half:]], indexes=indexes[perm[half:]])
    out2 = m.compute()
    if name != "cr_ties":  # a shuffle changes which of two tied rows comes first; the reference has the same dependence
        for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
            assert abs(out2["test/" + k] - float(z["test_" + k])) <= METRIC_ATOL, k


def test_aspect_metrics_on_the_references_ensemble_preds(golden_dir, metrics_mod):
    """EnsembleModule.on_test_epoch_end (ensemble_module.py:214-238): Diversity / Personalization of category and sentiment."""
    z = np.load(os.path.join(golden_dir, "ensemble_d128.npz"))
    co, ho = z["cand_offsets"], z["hist_offsets"]
    n = len(co) - 1
    cand_idx = torch.arange(n).repeat_interleave(torch.from_numpy(np.diff(co).astype(np.int64))).cuda()
    hist_idx = torch.arange(n).repeat_interleave(torch.from_numpy(np.diff(ho).astype(np.int64))).cuda()
    tc, ts = torch.from_numpy(z["category"][z["cand_ids"]]).cuda(), torch.from_numpy(z["sentiment"][z["cand_ids"]]).cuda()
    hc, hs = torch.from_numpy(z["category"][z["hist_ids"]]).cuda(), torch.from_numpy(z["sentiment"][z["hist_ids"]]).cuda()
    targets = torch.from_numpy(z["labels"].astype(np.int64)).cuda()
    for w in range(len(z["weightings"])):
        m = metrics_mod.RetrievalMetricsB200(prefix="test/", with_auc=False)
        m.update(torch.from_numpy(z[f"w{w}_preds"]).cuda(), targets, indexes=cand_idx, target_categories=tc, target_sentiments=ts,
                 hist_categories=hc, hist_sentiments=hs, hist_indexes=hist_idx)
        out = m.compute()
        for k in ("ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
                  "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
            assert abs(out["test/" + k] - float(z[f"w{w}_test_{k}"])) <= METRIC_ATOL, (w, k)


def test_metric_object_edge_cases(metrics_mod):
    m = metrics_mod.RetrievalMetricsB200()
    assert m.compute() == {}
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.update(torch.rand(4), torch.zeros(4), indexes=torch.zeros(4, dtype=torch.long))
    with pytest.raises(ValueError, match="indexes"):
        m.update(torch.rand(4).cuda(), torch.zeros(4).cuda(), indexes=None)
    # one impression without a positive, one with only positives, non-consecutive index values
    preds = torch.tensor([0.3, 0.2, 0.9, 0.8, 0.1], device="cuda")
    target = torch.tensor([0, 0, 1, 1, 1], device="cuda")
    out = m(preds, target, indexes=torch.tensor([7, 7, 42, 42, 42], device="cuda"))
    assert out["mrr"] == pytest.approx(0.5) and out["ndcg@5"] == pytest.approx(0.5) and out["auc"] == pytest.approx(2 / 3)


def test_aspect_metrics_top_k_under_ties_nan_and_signed_zero(metrics_mod):
    """The top-k whose aspect labels Diversity / Personalization look at is picked on the device by k rounds of a warp arg-max; the
    reference takes `argsort(preds, descending=True, stable=True)[:k]` (torch.sort: NaN above everything, equal values in position
    order, -0 == +0).  Quantised predictions (long runs of ties), NaNs, both zeros and +-inf, impressions from 1 to 300 candidates:
    every per-impression value must equal the oracle's."""
    g = np.random.default_rng(17)
    sizes = np.concatenate([[1, 2, 3, 31, 32, 33, 64, 65, 300], g.integers(2, 120, 150)])
    hsizes = g.integers(1, 50, sizes.shape[0])
    co = np.zeros(sizes.shape[0] + 1, dtype=np.int32)
    co[1:] = np.cumsum(sizes)
    ho = np.zeros(sizes.shape[0] + 1, dtype=np.int32)
    ho[1:] = np.cumsum(hsizes)
    n, nh = int(co[-1]), int(ho[-1])
    preds = (g.integers(-3, 4, n) / 2.0).astype(np.float32)  # 7 distinct values -> ties everywhere
    special = g.random(n)
    preds[special < 0.03] = np.nan
    preds[(special >= 0.03) & (special < 0.05)] = -0.0
    preds[(special >= 0.05) & (special < 0.06)] = np.inf
    preds[(special >= 0.06) & (special < 0.07)] = -np.inf
    labels = (g.random(n) < 0.1).astype(np.uint8)
    ccat, csent = g.integers(1, 19, n).astype(np.int32), g.integers(1, 4, n).astype(np.int32)
    hcat, hsent = g.integers(1, 19, nh).astype(np.int32), g.integers(1, 4, nh).astype(np.int32)
    cu = lambda a: torch.from_numpy(a).cuda()
    _, per, flags = metrics_mod.rank_metrics(cu(preds), cu(labels), cu(co), int(sizes.max()), cand_category=cu(ccat), cand_sentiment=cu(csent),
                                             hist_offsets=cu(ho), hist_category=cu(hcat), hist_sentiment=cu(hsent), want_per_impression=True)
    per = per.cpu().numpy()
    assert int(flags.item()) & ~nat.FLAG_OUTSIDE_UNIT == 0
    slots = {(nat.M_CATEG_DIV_K0, "div", "c", 5), (nat.M_CATEG_DIV_K1, "div", "c", 10), (nat.M_SENT_DIV_K0, "div", "s", 5), (nat.M_SENT_DIV_K1, "div", "s", 10),
             (nat.M_CATEG_PERS_K0, "pers", "c", 5), (nat.M_CATEG_PERS_K1, "pers", "c", 10), (nat.M_SENT_PERS_K0, "pers", "s", 5), (nat.M_SENT_PERS_K1, "pers", "s", 10)}
    for i in range(sizes.shape[0]):
        p = torch.from_numpy(preds[co[i]:co[i + 1]])
        for slot, kind, asp, k in slots:
            ca = torch.from_numpy((ccat if asp == "c" else csent)[co[i]:co[i + 1]].astype(np.int64))
            ha = torch.from_numpy((hcat if asp == "c" else hsent)[ho[i]:ho[i + 1]].astype(np.int64))
            ncls = 19 if asp == "c" else 4
            kk = min(k, int(sizes[i]))
            want = float(mo.diversity(p, ca, ncls, kk)) if kind == "div" else float(mo.personalization(p, ca, ha, ncls, kk))
            assert abs(per[i, slot] - want) <= 2e-6, (i, int(sizes[i]), kind, asp, k, per[i, slot], want)
