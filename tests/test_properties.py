"""Property tests (hypothesis): facts the domain guarantees whatever the inputs.

CPU: the oracle itself (metric definitions) -- invariance of an impression's metrics under the step it is batched in (padding),
under the order of the impressions, and of a tie-free impression under a permutation of its candidates; shard additivity.
GPU (``-m gpu``): the same properties on the CUDA path, plus kernel == oracle on drawn ragged behaviours."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from manner_b200 import data as mdata
from oracle import manner_oracle as mo

N_NEWS, DIM = 97, 32
TABLE = mdata.synth_table(N_NEWS, DIM, 5)


@st.composite
def behaviours(draw, max_impr=12, max_h=9, max_c=14):
    n = draw(st.integers(1, max_impr))
    hs = draw(st.lists(st.integers(1, max_h), min_size=n, max_size=n))
    cs = draw(st.lists(st.integers(1, max_c), min_size=n, max_size=n))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    hist = np.concatenate([rng.integers(0, N_NEWS, h) for h in hs]).astype(np.int32)
    cand = np.concatenate([rng.permutation(N_NEWS)[:c] for c in cs]).astype(np.int32)  # no duplicate candidate: no exact ties
    labels = np.concatenate([(rng.random(c) < 0.35).astype(np.uint8) for c in cs])
    off = lambda xs: np.concatenate([[0], np.cumsum(xs)]).astype(np.int32)
    return mdata.Behaviours(off(hs), hist, off(cs), cand, labels)


def _ob(b):
    return mo.Behaviours(b.hist_offsets, b.hist_ids, b.cand_offsets, b.cand_ids, b.labels)


def _permute_impressions(b, perm):
    parts = [b.slice(int(i), int(i) + 1) for i in perm]
    off = lambda arrs: np.concatenate([[0], np.cumsum([a.shape[0] for a in arrs])]).astype(np.int32)
    return mdata.Behaviours(off([p.hist_ids for p in parts]), np.concatenate([p.hist_ids for p in parts]), off([p.cand_ids for p in parts]),
                            np.concatenate([p.cand_ids for p in parts]), np.concatenate([p.labels for p in parts]))


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(behaviours(), st.integers(1, 8), st.integers(1, 8))
def test_oracle_padding_invariance(bhv, step_a, step_b):
    """The dense batches pad every impression to its step's longest (to_dense_batch); the late-fusion scores and the metrics
    must not depend on which impressions share a step (cr_module.py:108-131 adds exact zeros)."""
    a = mo.cr_eval_epoch(TABLE, _ob(bhv), step=step_a)
    b = mo.cr_eval_epoch(TABLE, _ob(bhv), step=step_b)
    np.testing.assert_allclose(a["scores"], b["scores"], rtol=2e-5, atol=2e-6)  # bmm may sum in another order for another padded shape
    for k in ("test/mrr", "test/ndcg@5", "test/ndcg@10"):
        if np.array_equal(mo.stable_ranks(a["scores"], bhv.cand_offsets), mo.stable_ranks(b["scores"], bhv.cand_offsets)):
            assert a["metrics"][k] == pytest.approx(b["metrics"][k], abs=1e-7)


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(behaviours(), st.randoms(use_true_random=False))
def test_oracle_impression_order_and_shard_additivity(bhv, rnd):
    scores = mo.cr_eval_epoch(TABLE, _ob(bhv))["scores"]
    per = mo.per_impression_metrics(scores, bhv.labels, bhv.cand_offsets)
    perm = list(range(bhv.n_impressions))
    rnd.shuffle(perm)
    shuffled = _permute_impressions(bhv, perm)
    s2 = mo.cr_eval_epoch(TABLE, _ob(shuffled), step=1)["scores"]
    per2 = mo.per_impression_metrics(s2, shuffled.labels, shuffled.cand_offsets)
    np.testing.assert_allclose(per2[:, :3], per[perm][:, :3], atol=1e-6)
    cut = bhv.n_impressions // 2
    if 0 < cut < bhv.n_impressions:  # metric sums are additive over any split
        head, tail = bhv.slice(0, cut), bhv.slice(cut, bhv.n_impressions)
        ph = mo.per_impression_metrics(scores[: head.n_cand], head.labels, head.cand_offsets)
        pt = mo.per_impression_metrics(scores[head.n_cand :], tail.labels, tail.cand_offsets)
        np.testing.assert_array_equal(np.concatenate([ph, pt]), per)


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 40), st.integers(0, 2**31 - 1))
def test_oracle_candidate_permutation_invariance_without_ties(c, seed):
    rng = np.random.default_rng(seed)
    scores = rng.permutation(c).astype(np.float32) * 0.37 - 3.0  # distinct values: no ties
    labels = (rng.random(c) < 0.4).astype(np.uint8)
    off = np.array([0, c], np.int32)
    base = mo.per_impression_metrics(scores, labels, off)
    p = rng.permutation(c)
    np.testing.assert_array_equal(mo.per_impression_metrics(scores[p], labels[p], off), base)
    assert mo.pooled_auc_exact(scores[p], labels[p]) == mo.pooled_auc_exact(scores, labels)


_SPECIAL = st.sampled_from([float("nan"), -float("nan"), 0.0, -0.0, float("inf"), -float("inf"), 1.0, -1.0, 0.5, 1e-45, -1e-45, 3.4e38])


@settings(max_examples=200, deadline=None)
@given(st.lists(st.one_of(_SPECIAL, st.floats(width=32, allow_nan=True, allow_infinity=True)), min_size=1, max_size=80))
def test_rank_key_is_the_stable_descending_sort_order(values):
    """The integer key the kernels rank by == torch's descending stable sort (NaN first, ties in position order, -0 == +0)."""
    x = np.asarray(values, dtype=np.float32)
    key = mo.rank_key(x).astype(np.int64)
    mine = np.lexsort((np.arange(x.shape[0]), -key))
    want = torch.argsort(torch.from_numpy(x), descending=True, stable=True).numpy()
    np.testing.assert_array_equal(mine, want)


# ---- the same properties on the CUDA path ----------------------------------------------------------------------------------


@pytest.fixture(scope="module")
def evaluator():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200.evaluator import ScoreEvaluator

    return ScoreEvaluator([TABLE]), ScoreEvaluator([mdata.synth_table(N_NEWS, 768, 6), mdata.synth_table(N_NEWS, 768, 7)])


@pytest.mark.gpu
@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(bhv=behaviours(), rnd=st.randoms(use_true_random=False))
def test_gpu_matches_oracle_and_is_order_and_shard_invariant(evaluator, bhv, rnd):
    ev, ev2 = evaluator
    res = ev.evaluate(ev.upload(bhv), pooled_auc=True, want_scores=True, want_per_impression=True)
    ref = mo.cr_eval_epoch(TABLE, _ob(bhv))
    got = res.scores.cpu().numpy()
    np.testing.assert_allclose(got, ref["scores"], rtol=2e-5, atol=2e-6)
    per_dev = res.per_impression.cpu().numpy()[0]
    np.testing.assert_array_equal(per_dev[:, :3], mo.per_impression_metrics(got, bhv.labels, bhv.cand_offsets)[:, :3])
    assert abs(res.auc - mo.pooled_auc_exact(got, bhv.labels)) < 1e-9
    # impression order: each impression's row of results moves with it, bit for bit
    perm = list(range(bhv.n_impressions))
    rnd.shuffle(perm)
    shuffled = _permute_impressions(bhv, perm)
    res2 = ev.evaluate(ev.upload(shuffled), pooled_auc=True, want_per_impression=True)
    np.testing.assert_array_equal(res2.per_impression.cpu().numpy()[0], per_dev[perm])
    assert res2.auc == res.auc
    np.testing.assert_allclose(res2.sums, res.sums, rtol=1e-12, atol=1e-9)
    # shards add up (z-scored two-module ensemble at the reference width, the kernel bench.py times)
    whole = ev2.evaluate(ev2.upload(bhv), weights=[[1.0, 0.4]], zscore=True)
    cut = bhv.n_impressions // 2
    if 0 < cut:
        parts = [ev2.evaluate(ev2.upload(p), weights=[[1.0, 0.4]], zscore=True) for p in (bhv.slice(0, cut), bhv.slice(cut, bhv.n_impressions))]
        np.testing.assert_allclose(parts[0].sums + parts[1].sums, whole.sums, rtol=1e-12, atol=1e-9)
