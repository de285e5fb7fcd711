"""GPU tests of the pipelined behaviour upload (mb200_upload_begin / _finish + mb200_eval_desc.ready, manner_b200/csrc/upload.cu):
the fused kernel starts on the first segment while the others are still being copied.  Nothing about the arithmetic changes,
so everything per impression must be bit-identical to the plain "copy, then launch" path; the fp64 sums only differ in the
order the per-warp partials were accumulated in.  Also: the pooled-AUROC rank search (shared-memory splitters, uniform
windows, galloping upper bound) against the oracle on tie-heavy and saturating inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import manner_oracle as mo  # noqa: E402  (checker only)

from manner_b200 import _native as nat  # noqa: E402
from manner_b200 import data as mdata  # noqa: E402


@pytest.fixture(scope="module")
def evaluator_cls():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200.evaluator import ScoreEvaluator

    return ScoreEvaluator


def _same(a, b):
    np.testing.assert_array_equal(a.scores.cpu().numpy(), b.scores.cpu().numpy())
    np.testing.assert_array_equal(a.per_impression.cpu().numpy(), b.per_impression.cpu().numpy())
    np.testing.assert_allclose(a.sums, b.sums, rtol=1e-12, atol=1e-9)
    assert a.auc == b.auc and a.auc_counts == b.auc_counts and a.n_impressions == b.n_impressions and a.flags == b.flags


@pytest.mark.parametrize("segments", [1, 3, 8, 32])
def test_pipelined_upload_is_bit_identical(evaluator_cls, segments):
    n_news = 4096
    tables = [mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS[:2]]
    bhv = mdata.synth_behaviours(n_news, 6000, seed=5)
    ev = evaluator_cls(tables, "cuda:0")
    kw = dict(weights=[[1.0, 0.4]], zscore=True, pooled_auc=True, want_scores=True, want_per_impression=True)
    plain = ev.evaluate(ev.upload(bhv), **kw)
    pinned = ev.pin(bhv)
    for it in range(4):  # repeated passes recycle the device buffers: the copies must wait for the previous pass
        d = ev.upload(bhv, pinned, pipelined=True, segments=segments, worker_segments=(it % 3))  # 0, 1, 2 segments by the library's thread
        assert d.ready is not None and d.ready_segments == segments
        _same(ev.evaluate(d, **kw), plain)


def test_pipelined_upload_changing_sets_early_fusion_and_loss(evaluator_cls):
    """Different behaviour sets back to back through the same evaluator (the recycled buffers hold the previous set's ids
    until the copies land), with the pads of early fusion / cross entropy travelling with the offsets."""
    n_news = 2048
    table = mdata.synth_table(n_news, 768, 1234)
    g = torch.Generator().manual_seed(3)
    att = (torch.randn(40, 768, generator=g) * 768 ** -0.5, torch.randn(40, generator=g) * 0.1, torch.rand(40, generator=g) * 0.2 - 0.1)
    ev = evaluator_cls([table], "cuda:0", attention=[att])
    kw = dict(want_scores=True, want_per_impression=True, loss="ce", pooled_auc=True)
    for seed in (1, 2, 3, 4):
        bhv = mdata.synth_behaviours(n_news, 3000 + 500 * seed, seed=seed)
        plain = ev.evaluate(ev.upload(bhv, step_batch=8), **kw)
        piped = ev.evaluate(ev.upload(bhv, step_batch=8, pipelined=True, segments=4), **kw)
        _same(piped, plain)
        assert piped.loss == plain.loss


def test_small_sets_fall_back_to_the_plain_copy(evaluator_cls):
    table = mdata.synth_table(512, 768, 1234)
    bhv = mdata.synth_behaviours(512, 64, seed=7, cand_window=300)
    ev = evaluator_cls([table], "cuda:0")
    d = ev.upload(bhv, pipelined=True)
    assert d.ready is None  # too few impressions to split
    ev.evaluate(d)


def test_upload_that_never_arrives_is_reported_not_hung(evaluator_cls):
    """A `ready` word nobody raises: the kernel gives up after its bounded wait and the host raises."""
    table = mdata.synth_table(512, 768, 1234)
    bhv = mdata.synth_behaviours(512, 64, seed=7, cand_window=300)
    ev = evaluator_cls([table], "cuda:0")
    d = ev.upload(bhv)
    d.ready, d.ready_segments = torch.zeros(1, dtype=torch.int32, device="cuda:0"), 2
    with pytest.raises(nat.NativeError, match="pipelined upload"):
        ev.evaluate(d)


@pytest.mark.parametrize("case", ["ties", "saturated", "few_negatives", "plain"])
def test_pooled_auc_rank_search_against_the_oracle(evaluator_cls, case):
    from manner_b200 import ops

    g = np.random.default_rng(11)
    n = 300_000
    if case == "ties":
        preds = g.integers(0, 50, n).astype(np.float32) / 50.0  # 50 distinct values in [0, 1): long runs of equal keys, no sigmoid
    elif case == "saturated":
        preds = (g.standard_normal(n) * 30).astype(np.float32)  # sigmoid saturates to exactly 0 / 1 for most rows
    elif case == "few_negatives":
        n = 5000
        preds = g.standard_normal(n).astype(np.float32)
    else:
        preds = g.standard_normal(n).astype(np.float32)
    labels = (g.random(n) < (0.9 if case == "few_negatives" else 0.05)).astype(np.uint8)
    p, lab = torch.from_numpy(preds).cuda(), torch.from_numpy(labels).cuda()
    flags = torch.tensor([0 if case == "ties" else nat.FLAG_OUTSIDE_UNIT], dtype=torch.int32, device="cuda:0")
    out = torch.ops.manner_b200.pooled_auc(p, lab, 2, flags).cpu().numpy()
    want = mo.pooled_auc_exact(preds, labels)
    # without the sigmoid the statistic is exact integer arithmetic; with it, torch's vectorised fp32 sigmoid and the kernel's
    # (fp64 exp, rounded once) may differ in the last place and turn a few near-equal pairs into ties or back
    assert abs(out[0] - want) < (1e-12 if case == "ties" else 1e-7), (case, out[0], want)
    assert int(out[1]) == int(labels.sum()) and int(out[2]) == n - int(labels.sum())
    # the staged form the multi-GPU paths use: build + sort, then rank every positive
    sorted_keys, pos_keys, n_pos = ops.auc_build_and_sort(p, lab, 2, flags)
    s2 = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    ops.auc_rank_sum(sorted_keys, n_pos, pos_keys, n_pos, s2)
    assert int(s2.item()) == int(out[3])
    # the bounded form (sort the positives, stream the negatives): the same integer statistic, bit for bit
    n_pos_host = int(labels.sum())
    for cap in (n_pos_host, n_pos_host + 1000):
        b = torch.ops.manner_b200.pooled_auc_bounded(p, lab, 2, flags, cap).cpu().numpy()
        assert b.tolist() == out.tolist(), (case, cap, b, out)
    if 2 * (n_pos_host - 1) <= n and n_pos_host > 1:
        under = torch.ops.manner_b200.pooled_auc_bounded(p, lab, 2, flags, n_pos_host - 1).cpu().numpy()
        assert np.isnan(under[0]) and int(under[1]) == n_pos_host  # more positives than promised: reported, not silently wrong
