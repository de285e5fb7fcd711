"""CPU: the oracle restatement vs (a) golden vectors produced by the reference's own code
(tests/golden/make_golden.py), (b) hand-computed known answers, (c) sklearn / scipy."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import manner_oracle as mo
from oracle import thirdparty as tp


def _bhv(z):
    return mo.Behaviours(z["hist_offsets"], z["hist_ids"], z["cand_offsets"], z["cand_ids"], z["labels"])


@pytest.mark.parametrize("name", ["cr_d128", "cr_d768", "cr_ties"])
def test_cr_epoch_matches_reference_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    out = mo.cr_eval_epoch(torch.from_numpy(z["table"]), _bhv(z))
    # same torch ops in the same order as cr_module.py:105-131 -> bit-identical on this machine class;
    # allow 1 ulp-scale slack for a different BLAS
    np.testing.assert_allclose(out["scores"], z["preds"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(out["targets"], z["targets"])
    np.testing.assert_array_equal(out["cand_news_size"], z["cand_news_size"])
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(out["metrics"]["test/" + k] - float(z["test_" + k])) <= 1e-7, k
    # CrossEntropyLoss per step (cr_module.py:171) and its MeanMetric over the steps (:255-259)
    np.testing.assert_allclose(out["step_losses"], z["step_losses"], rtol=1e-6)
    assert abs(out["metrics"]["test/loss"] - float(z["test_loss"])) <= 1e-6 * abs(float(z["test_loss"]))


def _attention(z):
    return mo.Attention(torch.from_numpy(z["att_weight"]), torch.from_numpy(z["att_bias"]), torch.from_numpy(z["att_query"]))


@pytest.mark.parametrize("name", ["cr_ef_d128", "cr_ef_d768"])
def test_early_fusion_epoch_matches_reference_golden(golden_dir, name):
    """late_fusion=False (cr_module.py:124-125): NAMLUserEncoder's additive attention over the PADDED history."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    table, bhv, att = torch.from_numpy(z["table"]), _bhv(z), _attention(z)
    out = mo.cr_eval_epoch(table, bhv, attention=att)
    np.testing.assert_allclose(out["scores"], z["preds"], rtol=1e-6, atol=1e-7)
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(out["metrics"]["test/" + k] - float(z["test_" + k])) <= 1e-7, k
    np.testing.assert_allclose(out["step_losses"], z["step_losses"], rtol=1e-6)
    # the cached-logit formulation the CUDA path uses (softmax over per-news logits + pad logits) is the same function
    logits = mo.attention_logits(att, table)
    pad_logit = float(torch.dot(torch.tanh(att.bias), att.query))
    H = np.diff(bhv.hist_offsets)
    for i in range(bhv.n_impressions):
        lo = i // 8 * 8
        n_pad = int(H[lo : lo + 8].max() - H[i])
        ids = torch.from_numpy(bhv.hist_ids[bhv.hist_offsets[i] : bhv.hist_offsets[i + 1]].astype(np.int64))
        l = torch.cat([logits[ids], torch.full((n_pad,), pad_logit)])
        w = torch.softmax(l, dim=0)[: len(ids)]
        user = (w.unsqueeze(1) * table[ids]).sum(0)
        cids = torch.from_numpy(bhv.cand_ids[bhv.cand_offsets[i] : bhv.cand_offsets[i + 1]].astype(np.int64))
        np.testing.assert_allclose((table[cids] @ user).numpy(), z["preds"][bhv.cand_offsets[i] : bhv.cand_offsets[i + 1]], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["cr_supcon_d128", "cr_supcon_ef_d768"])
def test_supcon_loss_matches_reference_golden(golden_dir, name):
    """supcon_loss=True (the reference default): the goldens hold what the reference's OWN SupConLoss.compute_loss / _compute_loss
    (components/losses.py:6-40) returned per step, including a step without any positive (losses.py:22) and a one-impression
    step with one positive and one negative (the `all(len(x) <= 1 ...)` guard, :15-16)."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    bhv = mo.Behaviours(z["hist_offsets"], z["hist_ids"], z["cand_offsets"], z["cand_ids"], z["labels"])
    att = mo.Attention(torch.from_numpy(z["att_weight"]), torch.from_numpy(z["att_bias"]), torch.from_numpy(z["att_query"])) if "att_weight" in z else None
    ref = mo.cr_eval_epoch(torch.from_numpy(z["table"]), bhv, attention=att, supcon_temperature=float(z["temperature"]))
    np.testing.assert_allclose(ref["step_losses"], z["step_losses"], rtol=1e-6, atol=0)
    assert ref["metrics"]["test/loss"] == pytest.approx(float(z["test_loss"]), rel=1e-6)
    if name == "cr_supcon_d128":
        assert z["step_losses"][1] == 0.0 and z["step_losses"][-1] == 0.0  # the two step-level guards fired in the reference


def test_supcon_restatement_known_answers():
    """Hand-computed values of the SupCon step loss (the restatement is pinned on the reference's own loss by the goldens above)."""
    bhv = mo.Behaviours(np.array([0, 1, 2, 3], np.int32), np.zeros(3, np.int32), np.array([0, 3, 5, 7], np.int32), np.zeros(7, np.int32),
                        np.array([1, 0, 0, 1, 1, 0, 0], np.uint8))
    batch = mo.step_batch(bhv, 0, 3)
    scores = torch.tensor([[1.0, 0.0, -1.0], [0.5, 0.5, 0.0], [2.0, 1.0, 0.0]])  # third column of rows 1, 2 is padding
    T = 0.5
    l0 = -(2.0 - math.log(math.exp(2.0) + 1.0 + math.exp(-2.0)))
    l1 = -(1.0 - math.log(2 * math.exp(1.0)))  # two positives, no negative: -mean(log 1/2) = log 2
    # row 2 has no positive: contributes 0 and is dropped by the AvgNonZero reduction
    want = (l0 + math.log(2.0)) / 2
    assert l1 == pytest.approx(math.log(2.0))
    assert float(mo.supcon_step_loss(scores, batch, T)) == pytest.approx(want, rel=1e-6)
    # a step without negatives (or without positives) gives 0
    bhv2 = mo.Behaviours(np.array([0, 1], np.int32), np.zeros(1, np.int32), np.array([0, 2], np.int32), np.zeros(2, np.int32), np.array([1, 1], np.uint8))
    assert float(mo.supcon_step_loss(torch.tensor([[0.3, 0.1]]), mo.step_batch(bhv2, 0, 1), T)) == 0.0


@pytest.mark.parametrize("name", ["ensemble_d128", "ensemble_d768"])
def test_ensemble_epoch_matches_reference_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    tabs = [torch.from_numpy(z[f"table{m}"]) for m in range(3)]
    aspects = {"category": z["category"], "sentiment": z["sentiment"]}
    for w, (wc, ws) in enumerate(z["weightings"].tolist()):
        out = mo.ensemble_eval_epoch(tabs, [1.0, wc, ws], _bhv(z), aspects)
        np.testing.assert_allclose(out["scores"], z[f"w{w}_preds"], rtol=1e-6, atol=1e-7)
        for k in ("ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
                  "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
            assert abs(out["metrics"]["test/" + k] - float(z[f"w{w}_test_{k}"])) <= 1e-7, (w, k)


def test_functional_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "functional.npz"))
    got = mo.dot_product(torch.from_numpy(z["user"]), torch.from_numpy(z["cand"]))
    np.testing.assert_allclose(got.numpy(), z["dot_product"], rtol=1e-6)
    p, c, h = torch.from_numpy(z["preds"]), torch.from_numpy(z["cats"]), torch.from_numpy(z["hist_cats"])
    assert float(mo.diversity(p, c, 19, 5)) == pytest.approx(float(z["div5"]), abs=1e-7)
    assert float(mo.diversity(p, c, 19, 10)) == pytest.approx(float(z["div10"]), abs=1e-7)
    assert float(mo.diversity(p[:4], torch.full((4,), 7), 19, 5)) == pytest.approx(float(z["div_onehot"]), abs=1e-9)
    assert float(mo.personalization(p, c, h, 19, 5)) == pytest.approx(float(z["pers5"]), abs=1e-7)
    assert float(mo.personalization(p, c, h, 19, 10)) == pytest.approx(float(z["pers10"]), abs=1e-7)
    assert float(mo.generalized_jaccard(torch.tensor([3, 0, 2, 1]), torch.tensor([1, 1, 2, 0]))) == pytest.approx(float(z["jaccard"]))


# ---- hand-computed known answers ---------------------------------------------------


def test_kat_three_candidates():
    s, y = torch.tensor([0.9, 0.1, 0.5]), torch.tensor([0, 1, 1])
    assert float(tp.retrieval_reciprocal_rank(s, y)) == pytest.approx(0.5)
    want = (1 / math.log2(3) + 1 / math.log2(4)) / (1 + 1 / math.log2(3))
    assert float(tp.retrieval_normalized_dcg(s, y, k=5)) == pytest.approx(want, abs=1e-6)
    assert mo.gauc_per_impression(s.numpy(), y.numpy()) == (0.0, True)


def test_kat_tie_is_stable_and_half_credit():
    s, y = torch.tensor([0.5, 0.5]), torch.tensor([0, 1])
    assert float(tp.retrieval_reciprocal_rank(s, y)) == pytest.approx(0.5)  # position 0 ranks first
    assert float(tp.binary_auroc(s, y)) == pytest.approx(0.5)
    assert mo.gauc_per_impression(s.numpy(), y.numpy())[0] == pytest.approx(0.5)


def test_kat_no_positive_and_short_lists():
    idx = torch.tensor([0, 0, 0, 1, 1])
    s = torch.tensor([0.3, 0.2, 0.1, 0.7, 0.6])
    y = torch.tensor([0, 0, 0, 0, 1])
    mrr, nd = tp.RetrievalMRR(), tp.RetrievalNormalizedDCG(k=5)
    mrr.update(s, y, idx), nd.update(s, y, idx)
    assert float(mrr.compute()) == pytest.approx((0.0 + 0.5) / 2)
    assert float(nd.compute()) == pytest.approx((0.0 + 1 / math.log2(3)) / 2, abs=1e-6)


def test_many_ties_descending_is_stable():
    s = torch.tensor([1.0] * 40 + [2.0] * 3)
    order = tp.stable_desc_argsort(s)
    assert order.tolist() == [40, 41, 42] + list(range(40))


def test_zscore_formula_by_hand():
    scores = torch.tensor([[1.0, 2.0, 4.0, 0.0], [3.0, 5.0, 0.0, 0.0]])
    mask = torch.tensor([[True, True, True, False], [True, True, False, False]])
    z = mo.zscore(scores, mask)
    m0, sd0 = 7.0 / 3.0, math.sqrt(((1 - 7 / 3) ** 2 + (2 - 7 / 3) ** 2 + (4 - 7 / 3) ** 2) / 2)
    np.testing.assert_allclose(z[0, :3].numpy(), [(1 - m0) / sd0, (2 - m0) / sd0, (4 - m0) / sd0], rtol=1e-6)
    np.testing.assert_allclose(z[1, :2].numpy(), [(3 - 4) / math.sqrt(2), (5 - 4) / math.sqrt(2)], rtol=1e-6)


def test_to_dense_batch_layout():
    x = torch.arange(10.0).view(5, 2)
    dense, mask = tp.to_dense_batch(x, torch.tensor([0, 0, 0, 2, 2]))
    assert dense.shape == (3, 3, 2) and mask.tolist() == [[True] * 3, [False] * 3, [True, True, False]]
    assert dense[2, 1].tolist() == [8.0, 9.0] and dense[1].abs().sum() == 0


# ---- independent cross-checks -----------------------------------------------------------------------


def test_pooled_auc_against_sklearn_and_exact_form():
    from sklearn.metrics import roc_auc_score

    rng = np.random.default_rng(0)
    s = (rng.standard_normal(20000) * 2).astype(np.float32)
    s[::7] = s[3]  # heavy ties
    y = (rng.random(20000) < 0.05).astype(np.int64)
    got = float(tp.binary_auroc(torch.from_numpy(s), torch.from_numpy(y)))
    sig = torch.from_numpy(s).sigmoid().numpy()
    assert got == pytest.approx(roc_auc_score(y, sig), abs=1e-6)
    assert mo.pooled_auc_exact(s, y) == pytest.approx(roc_auc_score(y, sig), abs=1e-9)
    # inside [0,1] the reference does not apply the sigmoid
    u = rng.random(5000).astype(np.float32)
    yu = (rng.random(5000) < 0.2).astype(np.int64)
    assert float(tp.binary_auroc(torch.from_numpy(u), torch.from_numpy(yu))) == pytest.approx(roc_auc_score(yu, u), abs=1e-6)
    assert mo.pooled_auc_exact(u, yu) == pytest.approx(roc_auc_score(yu, u), abs=1e-9)


def test_ndcg_against_sklearn_tie_free():
    from sklearn.metrics import ndcg_score

    rng = np.random.default_rng(1)
    for c in (2, 5, 9, 37):
        s = rng.standard_normal(c).astype(np.float32)
        y = np.zeros(c, dtype=np.int64)
        y[rng.permutation(c)[: max(1, c // 4)]] = 1
        for k in (5, 10):
            got = float(tp.retrieval_normalized_dcg(torch.from_numpy(s), torch.from_numpy(y), k=k))
            assert got == pytest.approx(ndcg_score(y[None], s[None], k=k, ignore_ties=True), abs=1e-6)


def test_vectorised_metrics_equal_the_faithful_loop(golden_dir):
    for name in ("cr_d128", "cr_ties"):
        z = np.load(os.path.join(golden_dir, name + ".npz"))
        per = mo.per_impression_metrics(z["preds"], z["labels"], z["cand_offsets"])
        assert float(per[:, 0].astype(np.float64).mean()) == pytest.approx(float(z["test_mrr"]), abs=1e-7)
        assert float(per[:, 1].astype(np.float64).mean()) == pytest.approx(float(z["test_ndcg@5"]), abs=1e-7)
        assert float(per[:, 2].astype(np.float64).mean()) == pytest.approx(float(z["test_ndcg@10"]), abs=1e-7)
        sizes = np.diff(z["cand_offsets"])
        g = mo.gauc_epoch(z["preds"], z["labels"].astype(np.int64), sizes)
        valid = per[:, 4] > 0
        assert float(per[valid, 3].astype(np.float64).mean()) == pytest.approx(g["test/gauc"], abs=1e-7)
        # per-impression values are bit-identical to the metric objects' values
        start = 0
        for i, c in enumerate(sizes.tolist()):
            p, t = torch.from_numpy(z["preds"][start : start + c]), torch.from_numpy(z["labels"][start : start + c].astype(np.int64))
            start += c
            if t.sum() == 0:
                assert per[i, 0] == 0 and per[i, 1] == 0 and per[i, 2] == 0
                continue
            assert per[i, 0] == float(tp.retrieval_reciprocal_rank(p, t))
            assert per[i, 1] == float(tp.retrieval_normalized_dcg(p, t, k=5))
            assert per[i, 2] == float(tp.retrieval_normalized_dcg(p, t, k=10))
