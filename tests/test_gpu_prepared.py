"""GPU tests of ScoreEvaluator.prepare(...).run() (manner_b200/prepared.py): the pre-built pass queues the same kernels as
upload() + evaluate(), so every number must agree -- sums to 1e-12 (the chunk schedule differs), the pooled AUROC, the loss and
the flat scores exactly -- and repeated runs must be bit-identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from manner_b200 import data as mdata  # noqa: E402


@pytest.fixture(scope="module")
def evaluator_cls():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200.evaluator import ScoreEvaluator

    return ScoreEvaluator


def _agree(a, b, scores=True):
    np.testing.assert_allclose(a.sums, b.sums, rtol=1e-12, atol=1e-9)
    assert a.auc == b.auc and a.auc_counts == b.auc_counts and a.n_impressions == b.n_impressions and a.flags == b.flags
    assert a.loss == b.loss
    if scores and a.scores is not None and b.scores is not None:
        np.testing.assert_array_equal(a.scores.cpu().numpy(), b.scores.cpu().numpy())
    assert a.metrics().keys() == b.metrics().keys()


@pytest.mark.parametrize("segments", [1, 8])
def test_prepared_ensemble_pass_equals_evaluate(evaluator_cls, segments):
    n_news = 4096
    tables = [mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS]
    aspects = mdata.synth_aspects(n_news)
    bhv = mdata.synth_behaviours(n_news, 5000, seed=9)
    ev = evaluator_cls(tables, "cuda:0", news_category=aspects["category"], news_sentiment=aspects["sentiment"])
    kw = dict(weights=[[1.0, 0.4, 0.0], [1.0, 0.2, 0.7]], zscore=True, pooled_auc=True)
    ref = ev.evaluate(ev.upload(bhv), want_scores=True, **kw)
    pp = ev.prepare(bhv, segments=segments, want_scores=True, **kw)
    first = pp.run()
    _agree(first, ref)
    # resident form: behaviours copied once at prepare(), passes queued back to back without a host read in between
    rp = ev.prepare(bhv, resident=True, want_scores=True, **kw)
    for _ in range(3):
        rp.launch()
    res = rp.read()
    _agree(res, ref)
    np.testing.assert_array_equal(res.sums, ref.sums)  # resident sets run the static schedule of evaluate(): bit-identical
    for _ in range(3):
        again = pp.run()
        np.testing.assert_array_equal(again.sums, first.sums)  # same chunks, same slots, same order: bit-identical
        assert again.auc == first.auc


def test_prepared_early_fusion_with_losses(evaluator_cls):
    n_news = 2048
    table = mdata.synth_table(n_news, 768, 1234)
    g = torch.Generator().manual_seed(3)
    att = (torch.randn(40, 768, generator=g) * 768 ** -0.5, torch.randn(40, generator=g) * 0.1, torch.rand(40, generator=g) * 0.2 - 0.1)
    ev = evaluator_cls([table], "cuda:0", attention=[att])
    bhv = mdata.synth_behaviours(n_news, 3001, seed=2)
    for loss in ("ce", "supcon"):
        ref = ev.evaluate(ev.upload(bhv, step_batch=8), pooled_auc=True, loss=loss, temperature=0.36, want_scores=True)
        got = ev.prepare(bhv, pooled_auc=True, loss=loss, temperature=0.36, want_scores=True).run()
        _agree(got, ref)
        assert got.loss is not None and got.loss > 0


def test_prepared_sweep_and_small_set(evaluator_cls):
    n_news = 1024
    tables = [mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS]
    ev = evaluator_cls(tables, "cuda:0")
    grid = torch.tensor([[1.0, a / 4.0, b / 4.0] for a in range(5) for b in range(5)], dtype=torch.float32, device="cuda:0")
    bhv = mdata.synth_behaviours(n_news, 900, seed=4, cand_window=500)
    _agree(ev.prepare(bhv, weights=grid, zscore=True).run(), ev.evaluate(ev.upload(bhv), weights=grid, zscore=True))
    tiny = mdata.synth_behaviours(n_news, 40, seed=5, cand_window=500)  # too small to split: one segment
    _agree(ev.prepare(tiny, weights=[[1.0, 0.4, 0.2]], zscore=True, pooled_auc=True).run(),
           ev.evaluate(ev.upload(tiny), weights=[[1.0, 0.4, 0.2]], zscore=True, pooled_auc=True))


def test_prepared_pass_through_the_fused_exchange_with_one_rank(evaluator_cls):
    """distributed=True with a single rank: payload packing, key build, sort and both exchange kernels (mailbox in local memory)."""
    n_news = 2048
    tables = [mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS[:2]]
    ev = evaluator_cls(tables, "cuda:0", exchange="p2p")
    bhv = mdata.synth_behaviours(n_news, 4000, seed=6)
    kw = dict(weights=[[1.0, 0.4]], zscore=True, pooled_auc=True)
    ref = ev.evaluate(ev.upload(bhv), **kw)
    pp = ev.prepare(bhv, distributed=True, pos_cap=int(bhv.labels.sum()), **kw)
    for _ in range(3):
        _agree(pp.run(), ref, scores=False)
    # and interleaved with the generic distributed path on the same mailboxes (the epoch counter is shared)
    gen = ev.evaluate(ev.upload(bhv, pos_cap=int(bhv.labels.sum())), distributed=True, **kw)
    _agree(gen, ref, scores=False)
    _agree(pp.run(), ref, scores=False)
