"""Parity of the configurations bench.py actually times, at their own size:

* BASELINE.json configs[1]: MIND-small shape (73 152 impressions, 768-d fp32), CR + category A-Module z-score ensemble,
  categ_weight 0.4 -- against the reference-faithful oracle loop on a prefix and, on ALL impressions, against the same
  arithmetic in fp64 (scores) and the vectorised metric oracle (per-impression values, bit-exact on the device's scores);
* configs[3]: 121 weightings x 3 modules at D = 768 from one gather against one call per weighting and the oracle;
* configs[2]: a 100 k-impression sample of the MIND-large shape.

Every rank flip between the device's and the oracle's scores must be a near-tie inside the stated score tolerance
(``mo.unexplained_rank_flips``); a metric may then differ by that many 1/B, never silently.
Also: the streaming kernel (hot-row cache in shared memory, rotating row pipeline) is bit-identical to the register-batch
kernels it replaced, with the cache on and off.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import manner_oracle as mo  # noqa: E402  (checker only)

from manner_b200 import _native as nat  # noqa: E402
from manner_b200 import data as mdata  # noqa: E402
from manner_b200 import ops  # noqa: E402

METRIC_ATOL = 1e-6
CATEG_WEIGHT = 0.4


@pytest.fixture(scope="module")
def evaluator_cls():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200.evaluator import ScoreEvaluator

    return ScoreEvaluator


def _ob(b):
    return mo.Behaviours(b.hist_offsets, b.hist_ids, b.cand_offsets, b.cand_ids, b.labels)


def _check_scores_and_flips(got: np.ndarray, truth: np.ndarray, tol: np.ndarray, offsets: np.ndarray, what: str) -> int:
    err = np.abs(got.astype(np.float64) - truth)
    worst = int(np.argmax(err / tol))
    assert np.all(err <= tol), f"{what}: score {worst} off by {err[worst]:.3e} > tol {tol[worst]:.3e}"
    flips, unexplained, gap = mo.unexplained_rank_flips(got.astype(np.float64), truth, tol, offsets)
    print(f"{what}: max |ds|/tol = {float((err / tol).max()):.4f}; rank flips vs fp64 = {flips} (all near-ties, widest gap {gap:.3e})")
    assert unexplained == 0, f"{what}: {unexplained} rank flips that are not near-ties"
    return flips


def test_headline_ensemble_full_size(evaluator_cls):
    """configs[1] as bench.py runs it: M = 2, z-score, weights [1, 0.4], all 73 152 impressions."""
    tables, bhv = mdata.synth_workload("small", n_modules=2)
    weights = [1.0, CATEG_WEIGHT]
    ev = evaluator_cls(tables)
    dev_bhv = ev.upload(bhv)
    res = ev.evaluate(dev_bhv, weights=[weights], zscore=True, pooled_auc=True, want_scores=True, want_per_impression=True)
    scores = res.scores.cpu().numpy()

    # (a) all impressions: scores against the fp64 evaluation of ensemble_module.py:95-151, flips bounded by near-ties
    truth, tol = mo.ensemble_truth_f64(tables, weights, bhv)
    _check_scores_and_flips(scores, truth, tol, bhv.cand_offsets, "S ensemble, full size")
    # (b) all impressions: rankings / MRR / nDCG@5/10 bit-exact on the device's own scores, gAUC to 1e-7, AUROC exact
    per_ref = mo.per_impression_metrics(scores, bhv.labels, bhv.cand_offsets)
    per_dev = res.per_impression.cpu().numpy()[0]
    np.testing.assert_array_equal(per_dev[:, :3], per_ref[:, :3])
    np.testing.assert_allclose(per_dev[:, 3:5], per_ref[:, 3:5], atol=1e-7)
    n = bhv.n_impressions
    for slot in (nat.M_MRR, nat.M_NDCG_K0, nat.M_NDCG_K1):
        assert abs(res.sums[0, slot] / n - per_ref[:, slot].astype(np.float64).mean()) < 1e-9
    assert abs(res.auc - mo.pooled_auc_exact(scores, bhv.labels)) < 1e-9

    # (c) a 2 048-impression prefix against the reference-faithful loop (steps of 8, dense padding, per-row loops, metric objects)
    head = bhv.slice(0, 2048)
    ref = mo.ensemble_eval_epoch(tables, weights, _ob(head))
    nh = head.n_cand
    err = np.abs(scores[:nh].astype(np.float64) - ref["scores"].astype(np.float64))
    assert np.all(err <= 2e-5 * np.maximum(1.0, np.abs(ref["scores"]))), float(err.max())
    flips, unexplained, _ = mo.unexplained_rank_flips(scores[:nh].astype(np.float64), ref["scores"].astype(np.float64), tol[:nh], head.cand_offsets)
    assert unexplained == 0
    m = ev.evaluate(ev.upload(head), weights=[weights], zscore=True, pooled_auc=True).metrics()
    for k in ("ndcg@5", "ndcg@10", "mrr", "gauc"):
        # each near-tie flip may move one impression's value by at most 1 -> 1/B on the mean; no flips: the 1e-6 bar
        bound = METRIC_ATOL + flips / head.n_impressions
        assert abs(m["test/" + k] - ref["metrics"]["test/" + k]) <= bound, (k, m["test/" + k], ref["metrics"]["test/" + k], flips)


def test_sweep121_three_modules_reference_width(evaluator_cls):
    """configs[3]: W = 121 weightings, CR + category + sentiment tables, D = 768, on a prefix of the S shape: the
    lane-per-weighting sweep against one call per weighting (bit-identical) and against the oracle loop."""
    tables, bhv = mdata.synth_workload("small", n_modules=3)
    head = bhv.slice(0, 1536)
    grid = [[1.0, a / 10.0, b / 10.0] for a in range(11) for b in range(11)]
    ev = evaluator_cls(tables)
    dev_bhv = ev.upload(head)
    sweep = ev.evaluate(dev_bhv, weights=grid, zscore=True, want_per_impression=True)
    per_sweep = sweep.per_impression.cpu().numpy()
    for w in (0, 1, 11, 37, 60, 93, 120):
        single = ev.evaluate(dev_bhv, weights=[grid[w]], zscore=True, want_per_impression=True, want_scores=True)
        np.testing.assert_array_equal(per_sweep[w][:, :5], single.per_impression.cpu().numpy()[0][:, :5])
        np.testing.assert_allclose(sweep.sums[w][:5], single.sums[0][:5], rtol=0, atol=1e-9)
        got = single.scores.cpu().numpy()
        truth, tol = mo.ensemble_truth_f64(tables, grid[w], head)
        flips = _check_scores_and_flips(got, truth, tol, head.cand_offsets, f"sweep weighting {w}")
        if w in (37, 120):  # the reference-faithful loop is slow: two weightings
            ref = mo.ensemble_eval_epoch(tables, grid[w], _ob(head))
            m = sweep.metrics(weighting=w)
            fl, unexplained, _ = mo.unexplained_rank_flips(got.astype(np.float64), ref["scores"].astype(np.float64), tol, head.cand_offsets)
            assert unexplained == 0
            for k in ("ndcg@5", "ndcg@10"):
                assert abs(m["test/" + k] - ref["metrics"]["test/" + k]) <= METRIC_ATOL + fl / head.n_impressions, (w, k)


def test_large_shape_sample(evaluator_cls):
    """configs[2]: MIND-large catalogue (161 013 news), a 100 k-impression sample, M = 2 z-score ensemble."""
    n_news = mdata.SHAPES["large"][0]
    tables = [mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS[:2]]
    bhv = mdata.synth_behaviours(n_news, 100_000, mdata.SHAPES["large"][2])
    weights = [1.0, CATEG_WEIGHT]
    ev = evaluator_cls(tables)
    res = ev.evaluate(ev.upload(bhv), weights=[weights], zscore=True, pooled_auc=True, want_scores=True, want_per_impression=True)
    scores = res.scores.cpu().numpy()
    truth, tol = mo.ensemble_truth_f64(tables, weights, bhv)
    _check_scores_and_flips(scores, truth, tol, bhv.cand_offsets, "L sample")
    per_ref = mo.per_impression_metrics(scores, bhv.labels, bhv.cand_offsets)
    np.testing.assert_array_equal(res.per_impression.cpu().numpy()[0][:, :3], per_ref[:, :3])
    assert abs(res.auc - mo.pooled_auc_exact(scores, bhv.labels)) < 1e-9
    head = bhv.slice(0, 1024)
    ref = mo.ensemble_eval_epoch(tables, weights, _ob(head))
    nh = head.n_cand
    assert np.all(np.abs(scores[:nh].astype(np.float64) - ref["scores"]) <= 2e-5 * np.maximum(1.0, np.abs(ref["scores"])))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stream_kernel_bit_identical_to_register_batch_kernels(evaluator_cls, dtype):
    """Tuning variant 2 = the round-1 register-batch kernel (4-warp CTAs); 9 / default = one 16-warp CTA per SM with the
    hot-row cache in shared memory, 10 = that shape without the cache; 7 / 8 = the rotating-pipeline experiment without / with
    the cache: where a row is read from (HBM, L2, shared memory) must not change a single bit of any output."""
    n_news = 4096
    tables = [mdata.synth_table(n_news, 768, s, dtype) for s in mdata.TABLE_SEEDS]
    aspects = mdata.synth_aspects(n_news)
    bhv = mdata.synth_behaviours(n_news, 3000, seed=11, cand_window=600)
    grid = [[1.0, a / 4.0, b / 4.0] for a in range(5) for b in range(5)]
    outs = {}
    try:
        for variant in (2, 7, 8, 9, 10, -1):
            ops.set_tuning(variant=variant)
            runs = []
            ev = evaluator_cls(tables[:2])
            d = ev.upload(bhv)
            runs.append(ev.evaluate(d, weights=[[1.0, CATEG_WEIGHT]], zscore=True, pooled_auc=True, want_scores=True, want_per_impression=True))
            ev1 = evaluator_cls(tables[:1])
            runs.append(ev1.evaluate(ev1.upload(bhv), pooled_auc=True, want_scores=True, want_per_impression=True))
            ev3 = evaluator_cls(tables, news_category=aspects["category"], news_sentiment=aspects["sentiment"])
            runs.append(ev3.evaluate(ev3.upload(bhv), weights=[[1.0, 0.3, 0.5], [1.0, 0.0, 0.2]], zscore=True, want_scores=True, want_per_impression=True, scores_weighting=1))
            ev3b = evaluator_cls(tables)
            runs.append(ev3b.evaluate(ev3b.upload(bhv), weights=grid, zscore=True, want_scores=True, want_per_impression=True, scores_weighting=7))
            outs[variant] = runs
    finally:
        ops.set_tuning(variant=-1)
    base = outs[2]
    for variant in (7, 8, 9, 10, -1):
        for a, b in zip(base, outs[variant]):
            np.testing.assert_array_equal(a.scores.cpu().numpy().view(np.uint32), b.scores.cpu().numpy().view(np.uint32))
            np.testing.assert_array_equal(a.per_impression.cpu().numpy().view(np.uint32), b.per_impression.cpu().numpy().view(np.uint32))
            np.testing.assert_allclose(a.sums, b.sums, rtol=1e-13, atol=1e-9)  # same values, other partial-sum grouping (warps per grid differ)
            assert a.auc == b.auc and a.flags == b.flags


def test_stream_kernel_edge_shapes(evaluator_cls):
    """H and C around the pipeline depth and the 32-id window, the reference's limits (H = 50, C = 300), bad ids."""
    n_news = 700
    table = mdata.synth_table(n_news, 768, 5)
    rng = np.random.default_rng(3)
    hs = [1, 2, 3, 4, 5, 6, 29, 30, 31, 32, 33, 34, 50, 50, 1, 50]
    cs = [1, 2, 3, 4, 31, 32, 33, 2, 1, 64, 65, 95, 300, 1, 300, 299]
    off = lambda xs: np.concatenate([[0], np.cumsum(xs)]).astype(np.int32)
    labels = np.concatenate([(rng.random(c) < 0.3).astype(np.uint8) for c in cs])
    # hot ids on purpose: the first 20 news rows are gathered again and again so the cache is on
    hist = np.concatenate([np.where(rng.random(h) < 0.5, rng.integers(0, 20, h), rng.integers(0, n_news, h)) for h in hs]).astype(np.int32)
    cand = np.concatenate([rng.permutation(n_news)[:c] for c in cs]).astype(np.int32)
    bhv = mdata.Behaviours(off(hs), hist, off(cs), cand, labels)
    ev = evaluator_cls([table])
    outs = {}
    try:
        for variant in (2, 7, 8, 9, 10):
            ops.set_tuning(variant=variant)
            outs[variant] = ev.evaluate(ev.upload(bhv), pooled_auc=True, want_scores=True, want_per_impression=True)
    finally:
        ops.set_tuning(variant=-1)
    for variant in (7, 8, 9, 10):
        np.testing.assert_array_equal(outs[2].scores.cpu().numpy().view(np.uint32), outs[variant].scores.cpu().numpy().view(np.uint32))
        np.testing.assert_array_equal(outs[2].per_impression.cpu().numpy(), outs[variant].per_impression.cpu().numpy())
    ref = mo.cr_eval_epoch(table, _ob(bhv))
    got = outs[9].scores.cpu().numpy()
    truth, tol = mo.ensemble_truth_f64([table], [1.0], bhv, zscore_modules=False)
    _check_scores_and_flips(got, truth, tol, bhv.cand_offsets, "edge shapes")
    assert np.allclose(got, ref["scores"], rtol=2e-5, atol=2e-6)
    # a row id outside the table is flagged by the streaming kernel too
    bad = mdata.Behaviours(bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, np.where(np.arange(cand.size) == 40, n_news, cand).astype(np.int32), labels)
    with pytest.raises(nat.NativeError):
        ev.evaluate(ev.upload(bad))
