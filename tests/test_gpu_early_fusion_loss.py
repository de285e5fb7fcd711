"""GPU parity tests of the two SURVEY 8(f) rows next to the scoring path: early fusion (late_fusion=False: additive
attention over the padded history, cr_module.py:63-68,124-125) and the step losses behind test/loss / val/loss
(cr_module.py:140-171,253-259) -- the CUDA path through the C ABI against golden vectors produced by the reference's
own CRModule(late_fusion=False) and CrossEntropyLoss, and against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import manner_oracle as mo  # noqa: E402  (checker only)

from manner_b200 import _native as nat  # noqa: E402
from manner_b200 import data as mdata  # noqa: E402

SCORE_RTOL = 1e-5
METRIC_ATOL = 1e-6
LOSS_RTOL = 2e-6  # fp32 log-softmax on both sides; ours accumulates in fp64


@pytest.fixture(scope="module")
def evaluator_cls():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200.evaluator import ScoreEvaluator

    return ScoreEvaluator


def _bhv(z):
    return mdata.Behaviours(z["hist_offsets"].astype(np.int32), z["hist_ids"].astype(np.int32), z["cand_offsets"].astype(np.int32),
                            z["cand_ids"].astype(np.int32), z["labels"].astype(np.uint8))


def _obhv(b):
    return mo.Behaviours(b.hist_offsets, b.hist_ids, b.cand_offsets, b.cand_ids, b.labels)


def _att(z):
    return (torch.from_numpy(z["att_weight"]), torch.from_numpy(z["att_bias"]), torch.from_numpy(z["att_query"]))


def _cond_tol(table, bhv):
    """|ds| <= rtol * sum_d |u_d||c_d| with |u| bounded by the largest |row| of the history (softmax weights sum to <= 1)."""
    t = table.double().abs()
    tol = np.empty(bhv.n_cand, dtype=np.float64)
    for i in range(bhv.n_impressions):
        h = bhv.hist_ids[bhv.hist_offsets[i]:bhv.hist_offsets[i + 1]].astype(np.int64)
        c = bhv.cand_ids[bhv.cand_offsets[i]:bhv.cand_offsets[i + 1]].astype(np.int64)
        u = t[torch.from_numpy(h)].max(0).values
        tol[bhv.cand_offsets[i]:bhv.cand_offsets[i + 1]] = (t[torch.from_numpy(c)] @ u).numpy()
    return SCORE_RTOL * tol + 1e-30


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_logits_kernel(dtype):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200 import ops

    g = torch.Generator().manual_seed(3)
    for n, dim, q in ((1, 64, 8), (37, 128, 200), (300, 768, 200), (50, 1024, 33)):
        table = (torch.randn(n, dim, generator=g) * 2 / dim ** 0.5).to(dtype)
        w, b, qv = torch.randn(q, dim, generator=g) * dim ** -0.5, torch.randn(q, generator=g) * 0.1, torch.rand(q, generator=g) * 0.2 - 0.1
        got = ops.attention_logits(table.cuda(), w, b, qv).cpu()
        att = mo.Attention(w, b, qv)
        want = torch.cat([mo.attention_logits(att, table.float()), torch.dot(torch.tanh(b), qv).reshape(1)])
        # |d logit| <= sum_q |query_q| * |d pre-activation_q| (tanh is 1-Lipschitz), pre-activation error ~ rtol * sum_d |w||x|
        cond = (qv.abs() @ (w.abs() @ table.float().abs().T + b.abs().unsqueeze(1))).numpy()
        cond = np.concatenate([cond, [float(qv.abs() @ b.abs())]])
        assert got.shape == (n + 1,)
        assert np.all(np.abs(got.numpy().astype(np.float64) - want.numpy()) <= 1e-5 * cond + 1e-7), (n, dim, q)
        # deterministic
        assert torch.equal(got, ops.attention_logits(table.cuda(), w, b, qv).cpu())


@pytest.mark.parametrize("name", ["cr_ef_d128", "cr_ef_d768"])
def test_early_fusion_matches_reference_golden(golden_dir, evaluator_cls, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    table, bhv = torch.from_numpy(z["table"]), _bhv(z)
    ev = evaluator_cls([table], attention=[_att(z)])
    dev_bhv = ev.upload(bhv, step_batch=8)
    res = ev.evaluate(dev_bhv, pooled_auc=True, want_scores=True, want_per_impression=True, loss="ce")
    scores = res.scores.cpu().numpy()
    assert np.all(np.abs(scores.astype(np.float64) - z["preds"].astype(np.float64)) <= _cond_tol(table, bhv))
    per_dev = res.per_impression.cpu().numpy()[0]
    per_ref = mo.per_impression_metrics(scores, bhv.labels, bhv.cand_offsets)
    np.testing.assert_array_equal(per_dev[:, :3], per_ref[:, :3])  # rankings / MRR / nDCG bit-exact on the device's own scores
    m = res.metrics()
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(m["test/" + k] - float(z["test_" + k])) <= METRIC_ATOL, (k, m["test/" + k], float(z["test_" + k]))
    assert abs(m["test/loss"] - float(z["test_loss"])) <= 1e-5 * abs(float(z["test_loss"]))  # loss of scores that differ by 1e-5
    # without the step padding (hist_pad = 0) the scores are different: the pads really carry softmax mass
    with pytest.raises(ValueError, match="step_batch"):
        ev.evaluate(ev.upload(bhv), want_scores=True)


def test_early_fusion_other_widths_and_step_sizes(evaluator_cls):
    g = torch.Generator().manual_seed(8)
    for dim, step in ((64, 1), (100, 3), (400, 8), (768, 5)):
        n_news = 200
        bhv = mdata.synth_behaviours(n_news, 45, seed=dim, cand_window=150)
        table = mdata.synth_table(n_news, dim, 7)
        att = (torch.randn(50, dim, generator=g) * dim ** -0.5, torch.randn(50, generator=g) * 0.1, torch.rand(50, generator=g) * 0.2 - 0.1)
        ev = evaluator_cls([table], attention=[att])
        res = ev.evaluate(ev.upload(bhv, step_batch=step), want_scores=True, loss="ce")
        ref = mo.cr_eval_epoch(table, _obhv(bhv), step=step, attention=mo.Attention(*att))
        assert np.all(np.abs(res.scores.cpu().numpy().astype(np.float64) - ref["scores"]) <= _cond_tol(table, bhv)), (dim, step)
        assert abs(res.loss - ref["metrics"]["test/loss"]) <= 1e-5 * abs(ref["metrics"]["test/loss"])


@pytest.mark.parametrize("name", ["cr_d128", "cr_d768", "cr_ties"])
def test_cross_entropy_loss_matches_reference_golden(golden_dir, evaluator_cls, name):
    """test/loss of the late-fusion CRModule (supcon_loss=False): per-step CrossEntropyLoss over the padded score matrix,
    MeanMetric over the steps."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    table, bhv = torch.from_numpy(z["table"]), _bhv(z)
    ev = evaluator_cls([table])
    res = ev.evaluate(ev.upload(bhv, step_batch=8), loss="ce", want_per_impression=True)
    want = float(z["test_loss"])
    assert abs(res.loss - want) <= 1e-5 * abs(want), (res.loss, want)
    # per step: mean of the per-impression losses == the reference's step losses
    per = res.per_impression.cpu().numpy()[0][:, nat.M_LOSS].astype(np.float64)
    steps = np.array([per[lo:lo + 8].mean() for lo in range(0, bhv.n_impressions, 8)])
    np.testing.assert_allclose(steps, z["step_losses"], rtol=1e-5)
    assert abs(res.sums[0][nat.M_LOSS] - per.sum()) <= 1e-9 * abs(per.sum())
    # the loss does not disturb the metrics
    base = ev.evaluate(ev.upload(bhv))
    np.testing.assert_array_equal(res.sums[0][:5], base.sums[0][:5])


@pytest.mark.parametrize("name", ["cr_supcon_d128", "cr_supcon_ef_d768"])
def test_supcon_loss_matches_reference_golden(golden_dir, evaluator_cls, name):
    """test/loss with supcon_loss=True (the reference default, what ModelCheckpoint(monitor="val/loss") watches): against the
    per-step values of the reference's OWN SupConLoss (components/losses.py:6-40 run on the pytorch_metric_learning base
    restated in oracle/ref_stubs.py), late fusion and early fusion, including both step-level guards."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    table, bhv, T = torch.from_numpy(z["table"]), _bhv(z), float(z["temperature"])
    att = [(torch.from_numpy(z["att_weight"]), torch.from_numpy(z["att_bias"]), torch.from_numpy(z["att_query"]))] if "att_weight" in z else None
    ev = evaluator_cls([table], attention=att)
    res = ev.evaluate(ev.upload(bhv, step_batch=8), loss="supcon", temperature=T, want_per_impression=True, want_scores=True)
    want = float(z["test_loss"])
    assert abs(res.loss - want) <= 1e-5 * abs(want), (res.loss, want)
    # per step: mean of the non-zero per-impression losses == the reference's step losses, 0 where a step-level guard fired
    per = res.per_impression.cpu().numpy()[0][:, nat.M_LOSS].astype(np.float64)
    lab, off = bhv.labels, bhv.cand_offsets
    for s_i, lo in enumerate(range(0, bhv.n_impressions, 8)):
        hi = min(lo + 8, bhv.n_impressions)
        pos = int(lab[off[lo]:off[hi]].sum())
        neg = int(off[hi] - off[lo]) - pos
        guard = (pos <= 1 and neg <= 1) or pos == 0 or neg == 0
        nz = per[lo:hi][per[lo:hi] > 0]
        mine = 0.0 if (guard or nz.size == 0) else nz.mean()
        assert abs(mine - float(z["step_losses"][s_i])) <= 1e-5 * max(1.0, abs(float(z["step_losses"][s_i]))), (s_i, mine, float(z["step_losses"][s_i]))
    np.testing.assert_allclose(res.scores.cpu().numpy(), z["preds"], rtol=2e-5, atol=2e-6)


def test_supcon_loss_larger_set_against_oracle(evaluator_cls):
    """A larger behaviour set against the oracle (itself pinned on the reference's SupConLoss by tests/test_oracle.py)."""
    n_news, dim = 300, 128
    bhv = mdata.synth_behaviours(n_news, 83, seed=21, cand_window=200)
    table = mdata.synth_table(n_news, dim, 5)
    ev = evaluator_cls([table])
    for T in (0.1, 0.36):
        res = ev.evaluate(ev.upload(bhv, step_batch=8), loss="supcon", temperature=T)
        ref = mo.cr_eval_epoch(table, _obhv(bhv), supcon_temperature=T)
        assert abs(res.loss - ref["metrics"]["test/loss"]) <= 2e-5 * abs(ref["metrics"]["test/loss"]), (T, res.loss, ref["metrics"]["test/loss"])


def test_step_loss_kernel_known_answers():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200 import ops

    loss = torch.tensor([1.0, 3.0, 0.0, 2.0, 0.0, 0.0, 5.0], device="cuda")
    ce = ops.step_loss(loss, 3, nat.LOSS_CE).cpu().numpy()  # steps: [1,3,0] [2,0,0] [5] -> 4/3 + 2/3 + 5
    assert ce[1] == 3 and abs(ce[0] - (np.float32(4 / 3) + np.float64(np.float32(2 / 3)) + 5)) < 1e-12
    sc = ops.step_loss(loss, 3, nat.LOSS_SUPCON).cpu().numpy()  # mean of the non-zero losses: 2 + 2 + 5
    assert sc[1] == 3 and sc[0] == 9.0
    z = ops.step_loss(torch.zeros(4, device="cuda"), 2, nat.LOSS_SUPCON).cpu().numpy()
    assert z[0] == 0.0 and z[1] == 2
    # step-level guards of components/losses.py:15-16,22 (need the labels): steps of 2 impressions
    off = torch.tensor([0, 1, 2, 5, 8, 10, 12, 13], dtype=torch.int32, device="cuda")
    lab = torch.tensor([1, 0,  1, 0, 0, 1, 0, 0,  1, 1, 1, 1,  1], dtype=torch.uint8, device="cuda")
    g = ops.step_loss(loss, 2, nat.LOSS_SUPCON, off, lab).cpu().numpy()
    # step 0: one positive + one negative in total -> 0; step 1: (3 + 0 nonzero -> 3)... values [1,3][0,2][0,0][5]
    # step 0 = impressions 0,1 (labels [1],[0]) -> guard 1; step 1 = impressions 2,3 -> mean of {2} = 2; step 2 = impressions 4,5 all
    # positive -> no negative -> 0; step 3 = impression 6, a single positive -> 0
    assert g[1] == 4 and g[0] == 2.0, g
