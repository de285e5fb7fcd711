"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE (run in this container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only, never copied)

What runs unmodified from /root/reference/manner:
  CRModule.__init__/forward/model_step/test_step/on_test_epoch_end      (models/cr_module.py)
  EnsembleModule.__init__/forward/_submodel_forward/model_step/test_step/on_test_epoch_end
  DotProduct.forward                                                    (components/click_predictors.py)
  diversity / personalization / generalized_jaccard                     (metrics/functional.py)
  Diversity / Personalization / CustomRetrievalMetric.compute           (metrics/*.py)

What is substituted, and why:
  * the PLM ``MannerNewsEncoder`` -> a lookup into a seeded embedding table (the cached
    table is the boundary of the accelerated path);
  * ``load_from_checkpoint`` -> returns such table-backed sub-modules (no checkpoints exist here);
  * lightning / torch_geometric / torchmetrics / ... imports -> oracle/ref_stubs.py (absent packages;
    their arithmetic is the restatement in oracle/thirdparty.py, "parity unpinned");
  * ``torch.argsort(..., descending=True)`` defaults to ``stable=True`` while the reference code runs:
    the reference's era (torch 2.0/2.1 CPU) sorted stably, this image's torch 2.11 AVX-512 sort does
    not.  On the tie-free cases the generator asserts that the unpatched run gives
    identical numbers.

Every fixture stores its inputs (tables, CSR arrays, aspect labels, weights) next to the reference's
outputs, so tests never need the generator or /root/reference again.
"""
from __future__ import annotations

import os
import sys
from contextlib import contextmanager
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_stubs  # noqa: E402

ref_stubs.install("/root/reference")

from oracle import manner_oracle as mo  # noqa: E402  (only for Behaviours / step_batch input plumbing)

import manner.models.cr_module as ref_cr  # noqa: E402
import manner.models.a_module as ref_a  # noqa: E402
import manner.models.ensemble_module as ref_ens  # noqa: E402
from manner.metrics import functional as ref_F  # noqa: E402
from manner.models.components.click_predictors import DotProduct  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
_orig_argsort = torch.argsort


@contextmanager
def stable_argsort(enabled: bool = True):
    def patched(input, dim=-1, descending=False, stable=True):
        return _orig_argsort(input, dim=dim, descending=descending, stable=stable)

    if enabled:
        torch.argsort = patched
    try:
        yield
    finally:
        torch.argsort = _orig_argsort


class TableEncoder(torch.nn.Module):
    """Stands in for MannerNewsEncoder: ``x`` is the batch's x_hist / x_cand dict."""

    def __init__(self, table: Optional[torch.Tensor] = None, **ignored) -> None:
        super().__init__()
        self.table = table

    def forward(self, x: Dict) -> torch.Tensor:
        return self.table[x["news_row"]]


class _Sub(torch.nn.Module):
    def __init__(self, table: torch.Tensor) -> None:
        super().__init__()
        self.news_encoder = TableEncoder(table)


def make_behaviours(rng: np.random.Generator, n_news: int, hs: Sequence[int], cs: Sequence[int], ps: Sequence[int], dup_cands: bool = False) -> mo.Behaviours:
    hist_ids, cand_ids, labels = [], [], []
    for h, c, p in zip(hs, cs, ps):
        hist_ids.append(rng.integers(0, n_news, h))
        if dup_cands:  # few distinct rows -> exact score ties inside the impression
            cand_ids.append(rng.integers(0, max(2, c // 3), c))
        else:
            cand_ids.append(rng.permutation(n_news)[:c])
        lab = np.zeros(c, dtype=np.uint8)
        lab[rng.permutation(c)[:p]] = 1
        labels.append(lab)
    off = lambda xs: np.concatenate([[0], np.cumsum(xs)]).astype(np.int32)
    return mo.Behaviours(off(hs), np.concatenate(hist_ids).astype(np.int32), off(cs), np.concatenate(cand_ids).astype(np.int32), np.concatenate(labels))


def ragged_sizes(rng: np.random.Generator, b: int, cmax: int = 60):
    hs = np.clip(np.rint(rng.lognormal(2.0, 1.0, b)), 1, 50).astype(int)
    cs = np.clip(np.rint(rng.lognormal(2.6, 0.9, b)), 2, cmax).astype(int)
    ps = np.clip(1 + rng.poisson(0.5, b), 1, cs - 1).astype(int)
    # edge cases the domain has: H=1, H=50, C=2, C<5, no positive, all positive, many positives
    hs[0], hs[1] = 1, 50
    cs[2], ps[2] = 2, 1
    cs[3], ps[3] = 3, 1
    ps[4] = 0
    ps[5] = cs[5]
    ps[6] = max(1, cs[6] // 2)
    cs[7], ps[7] = cmax, 3
    return hs.tolist(), cs.tolist(), ps.tolist()


def table(n_news: int, dim: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_news, dim, generator=g) * (2.0 / dim**0.5)


def run_reference_cr(tab: torch.Tensor, bhv: mo.Behaviours, step: int = 8, late_fusion: bool = True, query_vector_dim: int = 200,
                     supcon_loss: bool = False, temperature: float = 0.1) -> Dict[str, np.ndarray]:
    ref_cr.MannerNewsEncoder = TableEncoder  # the only substitution inside CRModule.__init__
    model = ref_cr.CRModule(
        supcon_loss=supcon_loss, late_fusion=late_fusion, temperature=temperature, plm_model="", frozen_layers=[], dropout_probability=0.2,
        use_entities=False, pretrained_entity_embeddings_path="", entity_embedding_dim=100, num_attention_heads=10,
        query_vector_dim=query_vector_dim, text_embedding_dim=tab.shape[1], optimizer=None,
    )
    model.news_encoder.table = tab
    model.eval()
    step_losses = []
    with torch.no_grad():
        for i, lo in enumerate(range(0, bhv.n_impressions, step)):
            batch = mo.step_batch(bhv, lo, min(lo + step, bhv.n_impressions))
            step_losses.append(float(model.model_step(batch)[0]))  # CrossEntropyLoss()(scores, y_true) of cr_module.py:171
            model.test_step(batch, i)
        preds = torch.cat(model.test_step_outputs["preds"]).numpy().copy()
        targets = torch.cat(model.test_step_outputs["targets"]).numpy().copy()
        sizes = torch.cat(model.test_step_outputs["cand_news_size"]).numpy().copy()
        model.on_test_epoch_end()
    out = {"preds": preds, "targets": targets, "cand_news_size": sizes, "step_losses": np.asarray(step_losses, dtype=np.float32),
           "test/loss": np.float32(float(model.test_loss.compute()))}  # MeanMetric over the steps (cr_module.py:255-259)
    for k in ("test/auc", "test/mrr", "test/ndcg@5", "test/ndcg@10"):
        out[k] = np.float32(float(model.logged[k]))
    if not late_fusion:
        att = model.user_encoder.additive_attention  # NAMLUserEncoder -> AdditiveAttention (user_encoder.py:9-21, attention.py:6-29)
        out.update(att_weight=att.linear.weight.detach().numpy().copy(), att_bias=att.linear.bias.detach().numpy().copy(),
                   att_query=att.query.detach().numpy().copy())
    return out


def run_reference_ensemble(tabs: List[torch.Tensor], wc: float, ws: float, bhv: mo.Behaviours, aspects: Dict[str, np.ndarray], step: int = 8) -> Dict[str, np.ndarray]:
    ref_cr.CRModule.load_from_checkpoint = classmethod(lambda cls, checkpoint_path, **kw: _Sub(tabs[0]))
    by_path = {"categ": 1, "sent": 2}
    ref_a.AModule.load_from_checkpoint = classmethod(lambda cls, checkpoint_path, **kw: _Sub(tabs[by_path[checkpoint_path]]))
    model = ref_ens.EnsembleModule(
        cr_module_module_ckpt="cr", a_module_categ_ckpt="categ", a_module_sent_ckpt="sent",
        categ_weight=wc, sent_weight=ws, num_categ_classes=19, num_sent_classes=4,
    )
    model.eval()
    with torch.no_grad():
        for i, lo in enumerate(range(0, bhv.n_impressions, step)):
            model.test_step(mo.step_batch(bhv, lo, min(lo + step, bhv.n_impressions), aspects), i)
        preds = torch.cat(model.test_step_outputs["preds"]).numpy().copy()
        model.on_test_epoch_end()
    out = {"preds": preds}
    for k, v in model.logged.items():
        out[k] = np.float32(float(v))
    return out


def run_reference_baseline_epoch_end(bhv: mo.Behaviours, preds: np.ndarray, aspects: Dict[str, np.ndarray], step: int = 8) -> Dict[str, np.ndarray]:
    """The epoch end every BASELINE recommender of the reference shares (nrms_plm_module.py:275-313), run on the reference's own
    NRMSPLMModule: `test_step_outputs` is filled per step exactly as its test_step does (:258-272) from given flat predictions,
    then the unmodified on_test_epoch_end feeds the five metric collections."""
    import manner.models.baselines.nrms_plm_module as ref_nrms

    class _Inert(torch.nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__()

    ref_nrms.NewsEncoder = _Inert
    ref_nrms.UserEncoder = _Inert
    model = ref_nrms.NRMSPLMModule(plm_model="", frozen_layers=[], dropout_probability=0.2, text_embedding_dim=16, num_attention_heads=2,
                                   query_vector_dim=8, num_categ_classes=19, num_sent_classes=4, optimizer=None)
    ho, co = bhv.hist_offsets, bhv.cand_offsets
    for lo in range(0, bhv.n_impressions, step):
        hi = min(lo + step, bhv.n_impressions)
        c_ids, h_ids = bhv.cand_ids[co[lo]:co[hi]], bhv.hist_ids[ho[lo]:ho[hi]]
        out = model.test_step_outputs
        out["preds"].append(torch.from_numpy(preds[co[lo]:co[hi]]))
        out["targets"].append(torch.from_numpy(bhv.labels[co[lo]:co[hi]].astype(np.int64)))
        out["cand_news_size"].append(torch.from_numpy(np.diff(co[lo:hi + 1]).astype(np.int64)))
        out["hist_news_size"].append(torch.from_numpy(np.diff(ho[lo:hi + 1]).astype(np.int64)))
        out["target_categories"].append(torch.from_numpy(aspects["category"][c_ids].astype(np.int64)))
        out["target_sentiments"].append(torch.from_numpy(aspects["sentiment"][c_ids].astype(np.int64)))
        out["hist_categories"].append(torch.from_numpy(aspects["category"][h_ids].astype(np.int64)))
        out["hist_sentiments"].append(torch.from_numpy(aspects["sentiment"][h_ids].astype(np.int64)))
    with torch.no_grad():
        model.on_test_epoch_end()
    assert all(len(v) == 0 for v in model.test_step_outputs.values())  # the reference clears its buffers
    return {k: np.float32(float(v)) for k, v in model.logged.items()}


def save(name: str, **arrays) -> None:
    path = os.path.join(OUT, name + ".npz")
    np.savez(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1e3:.0f} kB)")


def bhv_arrays(bhv: mo.Behaviours) -> Dict[str, np.ndarray]:
    return dict(hist_offsets=bhv.hist_offsets, hist_ids=bhv.hist_ids, cand_offsets=bhv.cand_offsets, cand_ids=bhv.cand_ids, labels=bhv.labels)


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(1)

    # ---- CR eval, D=128, 40 ragged impressions (5 steps of 8), tie-free -------------------------
    rng = np.random.default_rng(2024)
    n_news, dim = 256, 128
    hs, cs, ps = ragged_sizes(rng, 40)
    bhv = make_behaviours(rng, n_news, hs, cs, ps)
    tab = table(n_news, dim, 1234)
    with stable_argsort():
        ref = run_reference_cr(tab, bhv)
    raw = run_reference_cr(tab, bhv)  # unpatched torch.argsort: must agree when there are no ties
    for k in ("test/mrr", "test/ndcg@5", "test/ndcg@10", "test/auc"):
        assert ref[k] == raw[k], (k, ref[k], raw[k])
    save("cr_d128", table=tab.numpy(), **bhv_arrays(bhv), **{k.replace("/", "_"): v for k, v in ref.items()})

    # ---- CR eval, D=768 (the reference's text_embedding_dim), 24 impressions ----------------------
    rng = np.random.default_rng(2025)
    n_news, dim = 96, 768
    hs, cs, ps = ragged_sizes(rng, 24, cmax=40)
    bhv = make_behaviours(rng, n_news, hs, cs, ps)
    tab = table(n_news, dim, 1234)
    with stable_argsort():
        ref = run_reference_cr(tab, bhv)
    save("cr_d768", table=tab.numpy(), **bhv_arrays(bhv), **{k.replace("/", "_"): v for k, v in ref.items()})

    # ---- CR eval with exact score ties (duplicate candidate rows): canonical stable rule -----------
    rng = np.random.default_rng(2026)
    n_news, dim = 64, 128
    hs, cs, ps = ragged_sizes(rng, 32, cmax=48)
    bhv = make_behaviours(rng, n_news, hs, cs, ps, dup_cands=True)
    tab = table(n_news, dim, 1234)
    with stable_argsort():
        ref = run_reference_cr(tab, bhv)
    save("cr_ties", table=tab.numpy(), **bhv_arrays(bhv), **{k.replace("/", "_"): v for k, v in ref.items()})

    # ---- CR eval, EARLY fusion (late_fusion=False: NAMLUserEncoder additive attention over the PADDED history) + CE loss ----
    for name, dim, qdim, n_news, b, seed in (("cr_ef_d128", 128, 200, 256, 40, 2028), ("cr_ef_d768", 768, 200, 96, 20, 2029)):
        rng = np.random.default_rng(seed)
        hs, cs, ps = ragged_sizes(rng, b, cmax=40)
        bhv = make_behaviours(rng, n_news, hs, cs, ps)
        tab = table(n_news, dim, 1234)
        torch.manual_seed(seed)  # the attention parameters are random-initialised by the reference's own constructors
        with stable_argsort():
            ref = run_reference_cr(tab, bhv, late_fusion=False, query_vector_dim=qdim)
        save(name, table=tab.numpy(), **bhv_arrays(bhv), **{k.replace("/", "_"): v for k, v in ref.items()})

    # ---- CR eval with the reference's OWN SupConLoss (supcon_loss=True, the reference default; components/losses.py on the
    #      pytorch_metric_learning base restated in oracle/ref_stubs.py), late and early fusion; the last step of the first set
    #      is a single impression with one positive and one negative -> the `all(len(x) <= 1 ...)` guard (losses.py:15-16) ----
    for name, dim, late, n_news, b, seed, T in (("cr_supcon_d128", 128, True, 256, 41, 2031, 0.36), ("cr_supcon_ef_d768", 768, False, 96, 22, 2032, 0.1)):
        rng = np.random.default_rng(seed)
        hs, cs, ps = ragged_sizes(rng, b, cmax=40)
        if name == "cr_supcon_d128":
            cs[-1], ps[-1] = 2, 1  # the 41st impression is alone in its step
            cs[8:16] = [3] * 8  # a step without any positive -> `pos_mask.any()` guard (losses.py:22)
            ps[8:16] = [0] * 8
        bhv = make_behaviours(rng, n_news, hs, cs, ps)
        tab = table(n_news, dim, 1234)
        torch.manual_seed(seed)
        with stable_argsort():
            ref = run_reference_cr(tab, bhv, late_fusion=late, supcon_loss=True, temperature=T)
        assert ref["step_losses"][-1] == 0.0 or name != "cr_supcon_d128"
        save(name, table=tab.numpy(), temperature=np.float32(T), **bhv_arrays(bhv), **{k.replace("/", "_"): v for k, v in ref.items()})

    # ---- ensemble: CR + category + sentiment A-Modules, 4 weightings, with aspects ------------------
    rng = np.random.default_rng(2027)
    n_news, dim = 200, 128
    hs, cs, ps = ragged_sizes(rng, 32, cmax=50)
    cs = [max(c, 2) for c in cs]
    bhv = make_behaviours(rng, n_news, hs, cs, ps)
    tabs = [table(n_news, dim, s) for s in (1234, 1235, 1236)]
    aspects = {
        "category": rng.integers(1, 19, n_news).astype(np.int32),
        "sentiment": rng.integers(1, 4, n_news).astype(np.int32),
    }
    # impression 9's candidates all get aspect label 0 -> the "neg" empty-target branch (base.py:114-122)
    c0, c1 = bhv.cand_offsets[9], bhv.cand_offsets[10]
    aspects["category"][bhv.cand_ids[c0:c1]] = 0
    weightings = [(0.0, 0.0), (0.3, 0.0), (0.0, 0.5), (0.2, 0.7)]
    fixture = dict(bhv_arrays(bhv), weightings=np.asarray(weightings, dtype=np.float64), category=aspects["category"], sentiment=aspects["sentiment"])
    for m, t in enumerate(tabs):
        fixture[f"table{m}"] = t.numpy()
    for w, (wc, ws) in enumerate(weightings):
        with stable_argsort():
            ref = run_reference_ensemble(tabs, wc, ws, bhv, aspects)
        for k, v in ref.items():
            fixture[f"w{w}_" + k.replace("/", "_")] = v
    save("ensemble_d128", **fixture)

    # ---- ensemble at the reference's width (D = 768: the exact-width kernels), 2 weightings, with aspects -----------
    rng = np.random.default_rng(2030)
    n_news, dim = 120, 768
    hs, cs, ps = ragged_sizes(rng, 24, cmax=40)
    cs = [max(c, 2) for c in cs]
    bhv = make_behaviours(rng, n_news, hs, cs, ps)
    tabs = [table(n_news, dim, s) for s in (1234, 1235, 1236)]
    aspects = {
        "category": rng.integers(1, 19, n_news).astype(np.int32),
        "sentiment": rng.integers(1, 4, n_news).astype(np.int32),
    }
    weightings = [(0.4, 0.0), (0.25, 0.6)]
    fixture = dict(bhv_arrays(bhv), weightings=np.asarray(weightings, dtype=np.float64), category=aspects["category"], sentiment=aspects["sentiment"])
    for m, t in enumerate(tabs):
        fixture[f"table{m}"] = t.numpy()
    for w, (wc, ws) in enumerate(weightings):
        with stable_argsort():
            ref = run_reference_ensemble(tabs, wc, ws, bhv, aspects)
        for k, v in ref.items():
            fixture[f"w{w}_" + k.replace("/", "_")] = v
    save("ensemble_d768", **fixture)

    # ---- the baselines' shared epoch end (nrms_plm_module.py:275-313) on the reference's own NRMSPLMModule --------------------
    rng = np.random.default_rng(2033)
    n_news = 300
    hs, cs, ps = ragged_sizes(rng, 48, cmax=45)
    bhv = make_behaviours(rng, n_news, hs, cs, ps)
    aspects = {"category": rng.integers(1, 19, n_news).astype(np.int32), "sentiment": rng.integers(1, 4, n_news).astype(np.int32)}
    c0, c1 = bhv.cand_offsets[11], bhv.cand_offsets[12]
    aspects["sentiment"][bhv.cand_ids[c0:c1]] = 0  # an impression whose sentiment labels sum to 0 -> the "neg" branch (base.py:114-122)
    preds = rng.standard_normal(int(bhv.cand_offsets[-1])).astype(np.float32)  # a baseline's click scores: any real numbers
    with stable_argsort():
        ref = run_reference_baseline_epoch_end(bhv, preds, aspects)
    save("baseline_epoch_end", preds=preds, category=aspects["category"], sentiment=aspects["sentiment"], **bhv_arrays(bhv),
         **{k.replace("/", "_"): v for k, v in ref.items()})

    # ---- functional level ------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    user = torch.randn(6, 1, 64, generator=g)
    cand = torch.randn(6, 64, 11, generator=g)
    dp = DotProduct()(user, cand)
    preds = torch.rand(23, generator=g)
    cats = torch.randint(1, 19, (23,), generator=g)
    hist_cats = torch.randint(1, 19, (17,), generator=g)
    with stable_argsort():
        div5 = ref_F.diversity(preds, cats, 19, k=5)
        div10 = ref_F.diversity(preds, cats, 19, k=10)
        div_onehot = ref_F.diversity(preds[:4], torch.full((4,), 7), 19, k=5)
        pers5 = ref_F.personalization(preds, cats, hist_cats, 19, k=5)
        pers10 = ref_F.personalization(preds, cats, hist_cats, 19, k=10)
    jac = ref_F.generalized_jaccard(torch.tensor([3, 0, 2, 1]), torch.tensor([1, 1, 2, 0]))
    save(
        "functional", user=user.numpy(), cand=cand.numpy(), dot_product=dp.numpy(), preds=preds.numpy(), cats=cats.numpy(),
        hist_cats=hist_cats.numpy(), div5=div5.numpy(), div10=div10.numpy(), div_onehot=div_onehot.numpy(),
        pers5=pers5.numpy(), pers10=pers10.numpy(), jaccard=jac.numpy(),
    )


if __name__ == "__main__":
    main()
