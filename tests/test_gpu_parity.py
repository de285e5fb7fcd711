"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle / the golden vectors the
reference's own code produced.  Run with ``pytest -m gpu`` on a B200."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import manner_oracle as mo  # noqa: E402  (checker only)
from oracle import thirdparty as tp  # noqa: E402

import manner_b200  # noqa: E402
from manner_b200 import _native as nat  # noqa: E402
from manner_b200 import data as mdata  # noqa: E402

SCORE_RTOL = 1e-5  # BASELINE.json: scores within 1e-5 relative in fp32
METRIC_ATOL = 1e-6  # AUC / MRR / nDCG within 1e-6 absolute


@pytest.fixture(scope="module")
def evaluator_cls():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200.evaluator import ScoreEvaluator

    return ScoreEvaluator


def _bhv(z):
    return mdata.Behaviours(z["hist_offsets"].astype(np.int32), z["hist_ids"].astype(np.int32), z["cand_offsets"].astype(np.int32),
                            z["cand_ids"].astype(np.int32), z["labels"].astype(np.uint8))


def _score_tol(table: torch.Tensor, bhv, scores_ref: np.ndarray) -> np.ndarray:
    """Condition-aware tolerance: |ds| <= rtol * sum_i |u_i c_i| per candidate -- a dot product
    can cancel to ~0, so the error scales with the magnitude of its terms, not of its value."""
    tol = np.empty_like(scores_ref, dtype=np.float64)
    t = table.double().abs()
    for i in range(bhv.n_impressions):
        h = bhv.hist_ids[bhv.hist_offsets[i]:bhv.hist_offsets[i + 1]]
        c = bhv.cand_ids[bhv.cand_offsets[i]:bhv.cand_offsets[i + 1]]
        u = t[torch.from_numpy(h.astype(np.int64))].mean(0)
        tol[bhv.cand_offsets[i]:bhv.cand_offsets[i + 1]] = (t[torch.from_numpy(c.astype(np.int64))] @ u).numpy()
    return SCORE_RTOL * tol + 1e-30


@pytest.mark.parametrize("name", ["cr_d128", "cr_d768", "cr_ties"])
def test_cr_eval_matches_reference_golden(golden_dir, evaluator_cls, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    table, bhv = torch.from_numpy(z["table"]), _bhv(z)
    ev = evaluator_cls([table])
    res = ev.evaluate(ev.upload(bhv), pooled_auc=True, want_scores=True, want_per_impression=True)
    scores = res.scores.cpu().numpy()
    # scores: within 1e-5 relative (condition-aware) of what the reference's CRModule produced
    assert np.all(np.abs(scores.astype(np.float64) - z["preds"].astype(np.float64)) <= _score_tol(table, bhv, z["preds"]))
    # rankings + per-impression metrics: bit-exact on identical scores (the device's own scores)
    per_dev = res.per_impression.cpu().numpy()[0]
    per_ref = mo.per_impression_metrics(scores, bhv.labels, bhv.cand_offsets)
    np.testing.assert_array_equal(per_dev[:, :3], per_ref[:, :3])
    np.testing.assert_allclose(per_dev[:, 3:5], per_ref[:, 3:5], atol=1e-7)
    # epoch metrics vs the reference's logged values
    m = res.metrics()
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(m["test/" + k] - float(z["test_" + k])) <= METRIC_ATOL, (k, m["test/" + k], float(z["test_" + k]))
    assert res.n_impressions == bhv.n_impressions


def test_ensemble_matches_reference_golden(golden_dir, evaluator_cls):
    z = np.load(os.path.join(golden_dir, "ensemble_d128.npz"))
    tabs = [torch.from_numpy(z[f"table{m}"]) for m in range(3)]
    bhv = _bhv(z)
    ev = evaluator_cls(tabs, news_category=z["category"], news_sentiment=z["sentiment"])
    dev_bhv = ev.upload(bhv)
    weightings = [[1.0, wc, ws] for wc, ws in z["weightings"].tolist()]
    # (a) one call per weighting, like one EnsembleModule run each
    for w, wt in enumerate(weightings):
        res = ev.evaluate(dev_bhv, weights=[wt], zscore=True, want_scores=True)
        got = res.scores.cpu().numpy().astype(np.float64)
        ref = z[f"w{w}_preds"].astype(np.float64)
        assert np.all(np.abs(got - ref) <= 2e-5 * np.maximum(1.0, np.abs(ref))), w  # z-scores are O(1)
        m = res.metrics()
        for k in ("ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
                  "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
            assert abs(m["test/" + k] - float(z[f"w{w}_test_{k}"])) <= METRIC_ATOL, (w, k, m["test/" + k], float(z[f"w{w}_test_{k}"]))
    # (b) the sweep: all weightings from ONE gather give the same sums as the separate calls
    sweep = ev.evaluate(dev_bhv, weights=weightings, zscore=True)
    for w, wt in enumerate(weightings):
        single = ev.evaluate(dev_bhv, weights=[wt], zscore=True)
        np.testing.assert_allclose(sweep.sums[w], single.sums[0], rtol=0, atol=1e-9)


def test_ensemble_at_reference_width_matches_reference_golden(golden_dir, evaluator_cls):
    """D = 768 (text_embedding_dim of the reference's configs): the exact-width kernels against what the reference's own
    EnsembleModule logged.  Scores within 2e-5 (z-scores are O(1)); metrics within 1e-6 unless a near-tie (|ds| below the score
    tolerance) ranks differently on the two sides, which is reported and bounded instead of hidden."""
    z = np.load(os.path.join(golden_dir, "ensemble_d768.npz"))
    tabs = [torch.from_numpy(z[f"table{m}"]) for m in range(3)]
    bhv = _bhv(z)
    ev = evaluator_cls(tabs, news_category=z["category"], news_sentiment=z["sentiment"])
    dev_bhv = ev.upload(bhv)
    for w, (wc, ws) in enumerate(z["weightings"].tolist()):
        res = ev.evaluate(dev_bhv, weights=[[1.0, wc, ws]], zscore=True, want_scores=True)
        got, ref = res.scores.cpu().numpy(), z[f"w{w}_preds"]
        assert np.all(np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= 2e-5 * np.maximum(1.0, np.abs(ref))), w
        flips = int((mo.stable_ranks(got, bhv.cand_offsets) != mo.stable_ranks(ref, bhv.cand_offsets)).sum())
        m = res.metrics()
        for k in ("ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
                  "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
            tol = METRIC_ATOL if flips == 0 else 2.0 / bhv.n_impressions
            assert abs(m["test/" + k] - float(z[f"w{w}_test_{k}"])) <= tol, (w, k, flips)


def test_weight_sweep_path_equals_single_weighting_calls(golden_dir, evaluator_cls):
    """BASELINE.json configs[3]: >= 16 weightings take the lane-per-weighting path; it must give exactly
    what one call per weighting gives (and therefore what the reference logs per weighting)."""
    z = np.load(os.path.join(golden_dir, "ensemble_d128.npz"))
    tabs = [torch.from_numpy(z[f"table{m}"]) for m in range(3)]
    bhv = _bhv(z)
    ev = evaluator_cls(tabs)  # no aspects: nDCG / MRR / gAUC only
    dev_bhv = ev.upload(bhv)
    grid = [[1.0, wc, ws] for wc in (0.0, 0.1, 0.25, 0.5, 1.0) for ws in (0.0, 0.2, 0.7, 1.0)] + [[1.0, wc, ws] for wc, ws in z["weightings"].tolist()]
    assert len(grid) >= 16
    sweep = ev.evaluate(dev_bhv, weights=grid, zscore=True, want_per_impression=True, want_scores=True, scores_weighting=23)
    per_sweep = sweep.per_impression.cpu().numpy()
    for w, wt in enumerate(grid):
        single = ev.evaluate(dev_bhv, weights=[wt], zscore=True, want_per_impression=True, want_scores=(w == 23))
        np.testing.assert_array_equal(per_sweep[w], single.per_impression.cpu().numpy()[0])
        np.testing.assert_allclose(sweep.sums[w], single.sums[0], rtol=0, atol=1e-9)
        if w == 23:
            np.testing.assert_array_equal(sweep.scores.cpu().numpy(), single.scores.cpu().numpy())
    # the last four weightings are the ones the reference was run with
    for w in range(4):
        m = sweep.metrics(weighting=len(grid) - 4 + w)
        for k in ("ndcg@5", "ndcg@10"):
            assert abs(m["test/" + k] - float(z[f"w{w}_test_{k}"])) <= METRIC_ATOL, (w, k)


def test_bf16_table_matches_oracle_on_rounded_table(golden_dir, evaluator_cls):
    z = np.load(os.path.join(golden_dir, "cr_d768.npz"))
    table, bhv = torch.from_numpy(z["table"]).bfloat16(), _bhv(z)
    ev = evaluator_cls([table])
    res = ev.evaluate(ev.upload(bhv), want_scores=True)
    ref = mo.cr_eval_epoch(table.float(), mo.Behaviours(bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, bhv.cand_ids, bhv.labels))
    got = res.scores.cpu().numpy().astype(np.float64)
    # stated bf16 tolerance: the table is rounded to bf16 on both sides, arithmetic is fp32 on both sides
    assert np.all(np.abs(got - ref["scores"]) <= _score_tol(table.float(), bhv, ref["scores"]))


def test_bf16_tensor_core_path_against_the_fma_kernels_and_the_oracle(evaluator_cls):
    """Tuning variant 11: bf16 rows of the reference width score their candidates on the tensor cores (mma.sync, u = hi + mid + lo
    split exactly into three bf16 vectors, fp32 accumulation); variant 3 / default is the FMA kernel.  Same contract -- bf16
    storage, fp32 arithmetic -- in another summation order: both within the condition-aware 1e-5 bar of the fp64 evaluation on
    the rounded table, every rank flip between them a near-tie.  (Measured slower than the FMA kernel: kept as an experiment.)"""
    from manner_b200 import ops

    n_news = 3000
    tables = [mdata.synth_table(n_news, 768, s, torch.bfloat16) for s in mdata.TABLE_SEEDS[:2]]
    bhv = mdata.synth_behaviours(n_news, 4000, seed=5, cand_window=800)
    res = {}
    try:
        for variant in (3, 11, -1):
            ops.set_tuning(variant=variant)
            ev = evaluator_cls(tables)
            res[variant] = ev.evaluate(ev.upload(bhv), weights=[[1.0, 0.4]], zscore=True, pooled_auc=True, want_scores=True, want_per_impression=True)
            ev1 = evaluator_cls(tables[:1])
            res[(variant, "cr")] = ev1.evaluate(ev1.upload(bhv), pooled_auc=True, want_scores=True)
    finally:
        ops.set_tuning(variant=-1)
    np.testing.assert_array_equal(res[3].scores.cpu().numpy(), res[-1].scores.cpu().numpy())  # the default is the FMA kernel
    for tabs, w, zs, key in ((tables, [1.0, 0.4], True, None), (tables[:1], [1.0], False, "cr")):
        truth, tol = mo.ensemble_truth_f64([t.float() for t in tabs], w, bhv, zscore_modules=zs)
        for variant in (3, 11):
            got = res[variant if key is None else (variant, key)].scores.cpu().numpy().astype(np.float64)
            assert np.all(np.abs(got - truth) <= tol), (variant, key, float((np.abs(got - truth) / tol).max()))
            flips, unexplained, _ = mo.unexplained_rank_flips(got, truth, tol, bhv.cand_offsets)
            assert unexplained == 0, (variant, key, flips)
    a, b = res[3], res[11]
    flips, unexplained, gap = mo.unexplained_rank_flips(a.scores.cpu().numpy().astype(np.float64), b.scores.cpu().numpy().astype(np.float64),
                                                        2e-5 * (1 + np.abs(b.scores.cpu().numpy().astype(np.float64))), bhv.cand_offsets)
    assert unexplained == 0
    for k in ("test/ndcg@5", "test/ndcg@10", "test/mrr", "test/auc"):
        assert abs(a.metrics()[k] - b.metrics()[k]) <= 1e-6 + flips / bhv.n_impressions, (k, flips)
    # per-impression metrics are bit-exact on the path's own scores
    per_ref = mo.per_impression_metrics(b.scores.cpu().numpy(), bhv.labels, bhv.cand_offsets)
    np.testing.assert_array_equal(b.per_impression.cpu().numpy()[0][:, :3], per_ref[:, :3])


def test_bf16_tables_stated_drift_against_the_fp32_reference(evaluator_cls):
    """The stated bf16 tolerance (BASELINE.json: "a stated bf16 tolerance otherwise"), measured against the
    fp32 tables on a MIND-small-shaped sample: rounding the tables to bf16 (8 significand bits) moves a z-scored ensemble score
    by at most 0.05 (typically 4e-3), flips a few % of the candidate ranks, and moves each epoch metric by < 5e-3 absolute.
    The numbers are printed so the evidence log keeps them."""
    tables, bhv = mdata.synth_workload("small", n_modules=2)
    bhv = bhv.slice(0, 20_000)
    ev32 = evaluator_cls(tables)
    ev16 = evaluator_cls([t.to(torch.bfloat16) for t in tables])
    kw = dict(weights=[[1.0, 0.4]], zscore=True, pooled_auc=True, want_scores=True)
    a, b = ev32.evaluate(ev32.upload(bhv), **kw), ev16.evaluate(ev16.upload(bhv), **kw)
    sa, sb = a.scores.cpu().numpy().astype(np.float64), b.scores.cpu().numpy().astype(np.float64)
    ds = np.abs(sa - sb)
    flips = int((mo.stable_ranks(sa, bhv.cand_offsets) != mo.stable_ranks(sb, bhv.cand_offsets)).sum())
    ma, mb = a.metrics(), b.metrics()
    dm = {k: abs(ma[k] - mb[k]) for k in ("test/auc", "test/mrr", "test/ndcg@5", "test/ndcg@10", "test/gauc")}
    print(f"bf16 vs fp32 tables, {bhv.n_impressions} impressions: max |ds| = {ds.max():.4f}, mean |ds| = {ds.mean():.5f}, "
          f"rank flips = {flips} of {sa.size} ({100.0 * flips / sa.size:.2f} %), metric deltas = { {k: round(v, 6) for k, v in dm.items()} }")
    assert ds.max() <= 0.05 and ds.mean() <= 6e-3
    assert flips <= 0.06 * sa.size
    assert all(v <= 5e-3 for v in dm.values()), dm


@pytest.mark.parametrize("dim", [64, 100, 256, 400, 1024])
def test_other_widths_against_oracle(evaluator_cls, dim):
    n_news = 300
    bhv = mdata.synth_behaviours(n_news, 48, seed=dim, cand_window=200)
    table = mdata.synth_table(n_news, dim, 77)
    ev = evaluator_cls([table])
    res = ev.evaluate(ev.upload(bhv), pooled_auc=True, want_scores=True)
    ref = mo.cr_eval_epoch(table, mo.Behaviours(bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, bhv.cand_ids, bhv.labels))
    got = res.scores.cpu().numpy()
    assert np.all(np.abs(got.astype(np.float64) - ref["scores"]) <= _score_tol(table, bhv, ref["scores"]))
    per = mo.per_impression_metrics(got, bhv.labels, bhv.cand_offsets)
    n = bhv.n_impressions
    assert abs(res.sums[0, nat.M_MRR] / n - per[:, 0].astype(np.float64).mean()) < 1e-9
    assert abs(res.sums[0, nat.M_NDCG_K1] / n - per[:, 2].astype(np.float64).mean()) < 1e-9
    assert abs(res.auc - mo.pooled_auc_exact(got, bhv.labels)) < METRIC_ATOL
    assert abs(res.auc - float(tp.binary_auroc(torch.from_numpy(got), torch.from_numpy(bhv.labels.astype(np.int64))))) < METRIC_ATOL


def test_pooled_auc_kernels_known_answers(evaluator_cls):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    # heavy ties + saturating scores + in-unit-interval scores (no sigmoid) + one-class edge cases
    cases = []
    s = (rng.standard_normal(50_000) * 3).astype(np.float32)
    s[::5] = s[7]
    cases.append((s, (rng.random(50_000) < 0.04).astype(np.uint8), 1))
    cases.append(((rng.standard_normal(20_000) * 40).astype(np.float32), (rng.random(20_000) < 0.3).astype(np.uint8), 1))
    cases.append((rng.random(30_000).astype(np.float32), (rng.random(30_000) < 0.1).astype(np.uint8), 0))
    cases.append((np.array([0.5, 0.5], np.float32), np.array([0, 1], np.uint8), 0))
    for preds, labels, mode in cases:
        out = torch.ops.manner_b200.pooled_auc(torch.from_numpy(preds).to(dev), torch.from_numpy(labels).to(dev), mode, None).cpu().numpy()
        want = float(tp.binary_auroc(torch.from_numpy(preds), torch.from_numpy(labels.astype(np.int64))))
        assert abs(out[0] - want) < METRIC_ATOL
        assert out[1] == labels.sum() and out[2] == labels.size - labels.sum()
    for labels in (np.zeros(100, np.uint8), np.ones(100, np.uint8)):
        out = torch.ops.manner_b200.pooled_auc(torch.rand(100, device=dev), torch.from_numpy(labels).to(dev), 0, None).cpu().numpy()
        assert out[0] == 0.0  # torchmetrics: no positive or no negative -> 0 (with a warning)


def test_edge_cases_and_error_reporting(evaluator_cls):
    table = mdata.synth_table(64, 128, 3)
    ev = evaluator_cls([table])
    # C = 1 with z-score -> NaN scores like torch.std of one element; rank is still 1
    bhv = mdata.Behaviours(np.array([0, 2, 3], np.int32), np.array([1, 2, 3], np.int32), np.array([0, 1, 4], np.int32),
                           np.array([5, 6, 7, 8], np.int32), np.array([1, 0, 1, 0], np.uint8))
    res = ev.evaluate(ev.upload(bhv), weights=[[1.0]], zscore=True, want_scores=True, want_per_impression=True)
    sc = res.scores.cpu().numpy()
    assert np.isnan(sc[0]) and not np.isnan(sc[1:]).any()
    per = res.per_impression.cpu().numpy()[0]
    assert per[0, nat.M_MRR] == 1.0 and per[0, nat.M_NDCG_K0] == 1.0
    # a row id outside the table is flagged, not read
    bad = mdata.Behaviours(bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, np.array([5, 6, 7, 64], np.int32), bhv.labels)
    with pytest.raises(nat.NativeError):
        ev.evaluate(ev.upload(bad))
    # CPU tensors are refused: there is no fallback path
    with pytest.raises(RuntimeError):
        torch.ops.manner_b200.pooled_auc(torch.rand(4), torch.zeros(4, dtype=torch.uint8), 0, None)


def test_empty_and_extreme_inputs(evaluator_cls):
    """Empty call, the largest impression shapes the reference's data can produce, and all-positive / no-positive lists."""
    table = mdata.synth_table(512, 768, 3)
    ev = evaluator_cls([table])
    empty = mdata.Behaviours(np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8))
    res = ev.evaluate(ev.upload(empty), pooled_auc=True, want_scores=True, want_per_impression=True)
    assert res.n_impressions == 0 and not res.sums.any() and res.scores.numel() == 0 and res.auc == 0.0
    assert all(v == 0.0 for v in res.metrics().values())
    # H = 50 (the reference's truncation, mind_rec_dataset.py:92), C = 300 candidates, every label pattern
    rng = np.random.default_rng(9)
    hs, cs = [50, 1, 50, 7], [300, 300, 2, 299]
    labels = [np.ones(300, np.uint8), np.zeros(300, np.uint8), np.array([0, 1], np.uint8), (rng.random(299) < 0.5).astype(np.uint8)]
    off = lambda xs: np.concatenate([[0], np.cumsum(xs)]).astype(np.int32)
    bhv = mdata.Behaviours(off(hs), rng.integers(0, 512, sum(hs)).astype(np.int32), off(cs),
                           np.concatenate([rng.permutation(512)[:c] for c in cs]).astype(np.int32), np.concatenate(labels))
    res = ev.evaluate(ev.upload(bhv), pooled_auc=True, want_scores=True, want_per_impression=True)
    ob = mo.Behaviours(bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, bhv.cand_ids, bhv.labels)
    ref = mo.cr_eval_epoch(table, ob, step=4)
    scores = res.scores.cpu().numpy()
    assert np.all(np.abs(scores.astype(np.float64) - ref["scores"]) <= _score_tol(table, bhv, ref["scores"]))
    per = mo.per_impression_metrics(scores, bhv.labels, bhv.cand_offsets)
    np.testing.assert_array_equal(res.per_impression.cpu().numpy()[0][:, :3], per[:, :3])
    assert per[1, 0] == 0.0 and per[1, 1] == 0.0  # no positive: MRR = nDCG = 0 (empty_target_action="neg")
    assert per[0, 0] == 1.0 and per[0, 1] == 1.0  # all positive
    # an impression longer than the declared max_cand is skipped and flagged, never read out of bounds
    dev_bhv = ev.upload(bhv)
    dev_bhv.max_cand = 100
    with pytest.raises(nat.NativeError, match="flagged bad input"):
        ev.evaluate(dev_bhv)


def test_small_shape_properties_full_size(evaluator_cls):
    """BASELINE.json's MIND-small shape (73 152 impressions, 768-d): size-independent properties."""
    tables, bhv = mdata.synth_workload("small", n_modules=1)
    ev = evaluator_cls(tables)
    dev_bhv = ev.upload(bhv)
    full = ev.evaluate(dev_bhv, pooled_auc=True, want_scores=True, want_per_impression=True)
    scores = full.scores.cpu().numpy()
    # per-impression metrics bit-exact against the oracle on the device's scores, at full size
    per_ref = mo.per_impression_metrics(scores, bhv.labels, bhv.cand_offsets)
    per_dev = full.per_impression.cpu().numpy()[0]
    np.testing.assert_array_equal(per_dev[:, :3], per_ref[:, :3])
    np.testing.assert_allclose(per_dev[:, 3:5], per_ref[:, 3:5], atol=1e-7)
    assert abs(full.auc - mo.pooled_auc_exact(scores, bhv.labels)) < 1e-9
    # run-to-run determinism (fixed-order reduction): bit-identical sums
    again = ev.evaluate(dev_bhv, pooled_auc=True)
    np.testing.assert_array_equal(full.sums, again.sums)
    assert full.auc == again.auc
    # sharding invariance: metric sums over any partition equal the whole (<= 1e-12 relative)
    bounds = mdata.balanced_shard_bounds(bhv, 3)
    acc = np.zeros_like(full.sums)
    for r in range(3):
        part = ev.evaluate(ev.upload(bhv.slice(int(bounds[r]), int(bounds[r + 1]))))
        acc += part.sums
    np.testing.assert_allclose(acc, full.sums, rtol=1e-12, atol=1e-9)
    # a 2 000-impression prefix against the reference-faithful oracle loop (steps of 8, padding, metric objects)
    head = bhv.slice(0, 2000)
    ref = mo.cr_eval_epoch(tables[0], mo.Behaviours(head.hist_offsets, head.hist_ids, head.cand_offsets, head.cand_ids, head.labels))
    n = head.n_cand
    assert np.all(np.abs(scores[:n].astype(np.float64) - ref["scores"]) <= _score_tol(tables[0], head, ref["scores"]))
    flips = int((mo.stable_ranks(scores[:n], head.cand_offsets) != mo.stable_ranks(ref["scores"], head.cand_offsets)).sum())
    head_res = ev.evaluate(ev.upload(head), pooled_auc=True)
    m = head_res.metrics()
    if flips == 0:  # near-tie flips (|ds| below tolerance) are reported, not hidden: they may move a metric by 1/B
        for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
            assert abs(m["test/" + k] - ref["metrics"]["test/" + k]) <= METRIC_ATOL, k
    else:
        print(f"near-tie rank flips between device and oracle scores: {flips}")
        assert flips <= 4
