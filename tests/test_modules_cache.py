"""Host mirror of the reference's module interface (manner_b200/modules.py) and the cache builder
(manner_b200/cache.py).  CPU tests cover the host logic; the GPU tests drive the step-mode drop-in with the
reference's own batches and compare with the values the reference logged (tests/golden)."""
import os
import types

import numpy as np
import pytest
import torch

from manner_b200 import cache as mcache
from manner_b200 import data as mdata
from oracle import manner_oracle as mo


class TableEncoder(torch.nn.Module):
    """Stands in for MannerNewsEncoder: looks the batch's news rows up in a table."""

    def __init__(self, table):
        super().__init__()
        self.register_buffer("table", table)

    def forward(self, x):
        return self.table[x["news_row"].to(self.table.device)]


def _bhv(z):
    return mo.Behaviours(z["hist_offsets"], z["hist_ids"], z["cand_offsets"], z["cand_ids"], z["labels"])


# ---- CPU -----------------------------------------------------------------------------------------------


def test_modules_import_without_the_reference_and_fail_loudly():
    from manner_b200 import modules

    assert not modules.HAVE_REFERENCE  # lightning / torchmetrics / torch_geometric are absent in this image
    with pytest.raises(ImportError, match="reference package"):
        modules.CRModuleB200(supcon_loss=False)
    seg = torch.tensor([0, 0, 2, 2, 2])
    assert modules._segment_offsets(seg, 3).tolist() == [0, 2, 2, 5]


def test_parsed_behaviours_to_csr_follows_the_reference_format():
    news_ids = [f"N{i}" for i in range(10)]
    nid2row = mcache.news_row_map(news_ids)
    # cells exactly as to_tsv() writes python lists (mind_dataframe.py:360-366) and the converters read them back
    hist = ["['N1', 'N2', 'N3']", "['N4']"]
    cand = ["['N5', 'N6']", "['N7', 'N8', 'N9']"]
    labs = ["[0, 1]", "[1, 0, 0]"]
    frame = types.SimpleNamespace()
    import pandas as pd

    frame = pd.DataFrame({"history": hist, "candidates": cand, "labels": labs})
    bhv = mcache.behaviours_frame_to_csr(frame, nid2row, max_history_length=2)
    assert bhv.hist_offsets.tolist() == [0, 2, 3] and bhv.hist_ids.tolist() == [1, 2, 4]  # first 2 clicks kept
    assert bhv.cand_offsets.tolist() == [0, 2, 5] and bhv.cand_ids.tolist() == [5, 6, 7, 8, 9]
    assert bhv.labels.tolist() == [0, 1, 1, 0, 0] and bhv.labels.dtype == np.uint8
    with pytest.raises(KeyError):
        mcache.behaviours_to_csr([["N1"]], [["N99"]], [[1]], nid2row)
    with pytest.raises(ValueError):
        mcache.behaviours_to_csr([[]], [["N1"]], [[1]], nid2row)


def test_read_parsed_behaviors_file(tmp_path):
    """The reference's cached parsed_behaviors.tsv (pandas ``to_csv(sep="\\t")`` of list-valued cells) -> CSR."""
    import pandas as pd

    beh = pd.DataFrame({"user": [3, 7], "history": [["N1", "N2", "N3"], ["N4"]], "candidates": [["N5", "N6"], ["N7", "N8", "N9"]],
                        "labels": [[0, 1], [1, 0, 0]]})
    path = tmp_path / "parsed_behaviors.tsv"
    beh.to_csv(path, sep="\t", index=False)  # what mind_dataframe.py's to_tsv writes
    nid2row = mcache.news_row_map([f"N{i}" for i in range(10)])
    bhv = mcache.read_parsed_behaviors(str(path), nid2row, max_history_length=2)
    same = mcache.behaviours_frame_to_csr(pd.read_table(path), nid2row, max_history_length=2)
    for a, b in ((bhv.hist_offsets, same.hist_offsets), (bhv.hist_ids, same.hist_ids), (bhv.cand_offsets, same.cand_offsets),
                 (bhv.cand_ids, same.cand_ids), (bhv.labels, same.labels)):
        np.testing.assert_array_equal(a, b)
    assert bhv.hist_ids.tolist() == [1, 2, 4] and bhv.cand_ids.tolist() == [5, 6, 7, 8, 9] and bhv.labels.tolist() == [0, 1, 1, 0, 0]


def test_build_embedding_table_on_cpu_plumbing():
    table = torch.randn(10, 16)
    enc = TableEncoder(table)
    batches = [{"news_row": torch.arange(0, 4)}, {"news_row": torch.arange(4, 10)}]
    out = mcache.build_embedding_table(enc, batches, 10, 16, torch.device("cpu"))
    assert torch.equal(out, table)
    with pytest.raises(ValueError):
        mcache.build_embedding_table(enc, batches[:1], 10, 16, torch.device("cpu"))


# ---- GPU: the drop-in against what the reference logged ---------------------------------------------------


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cr_d128", "cr_ties"])
def test_cr_step_mode_dropin_logs_the_reference_values(golden_dir, name):
    from manner_b200.modules import B200EvalMixin

    class FakeCR(B200EvalMixin, torch.nn.Module):
        def __init__(self, table):
            super().__init__()
            self.news_encoder = TableEncoder(table)
            self.logged = {}

        def _b200_encoders(self):
            return [self.news_encoder]

        def log_dict(self, values, **kw):
            self.logged.update(values)

    z = np.load(os.path.join(golden_dir, name + ".npz"))
    bhv = _bhv(z)
    model = FakeCR(torch.from_numpy(z["table"])).cuda()
    for i, lo in enumerate(range(0, bhv.n_impressions, 8)):  # configs/data/mind_rec.yaml:51 -> steps of 8
        model.test_step(mo.step_batch(bhv, lo, min(lo + 8, bhv.n_impressions)), i)
    model.on_test_epoch_end()
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(model.logged["test/" + k] - float(z["test_" + k])) <= 1e-6, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cr_ef_d128", "cr_d768"])
def test_cr_step_mode_early_fusion_and_loss(golden_dir, name):
    """Step mode with late_fusion=False (attention parameters taken from the module) and test/loss: the values the
    reference's CRModule logged."""
    from manner_b200.modules import B200EvalMixin

    z = np.load(os.path.join(golden_dir, name + ".npz"))
    early = "att_weight" in z.files

    class FakeCR(B200EvalMixin, torch.nn.Module):
        def __init__(self, table):
            super().__init__()
            self.news_encoder = TableEncoder(table)
            if early:
                self.att = torch.nn.ParameterList([torch.nn.Parameter(torch.from_numpy(z[k])) for k in ("att_weight", "att_bias", "att_query")])
            self.logged = {}

        def _b200_encoders(self):
            return [self.news_encoder]

        def _b200_loss(self):
            return "ce"

        def _b200_attention(self):
            return [tuple(self.att)] if early else None

        def log_dict(self, values, **kw):
            self.logged.update(values)

    bhv = _bhv(z)
    model = FakeCR(torch.from_numpy(z["table"])).cuda()
    for i, lo in enumerate(range(0, bhv.n_impressions, 8)):
        model.test_step(mo.step_batch(bhv, lo, min(lo + 8, bhv.n_impressions)), i)
    model.on_test_epoch_end()
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(model.logged["test/" + k] - float(z["test_" + k])) <= 1e-6, k
    assert abs(model.logged["test/loss"] - float(z["test_loss"])) <= 1e-5 * abs(float(z["test_loss"]))
    # the validation path (cr_module.py:214-251) logs the same numbers under val/ plus the best val/loss so far
    for epoch, n_steps in enumerate((2, None)):  # a short first "epoch", then the full one
        for i, lo in enumerate(range(0, bhv.n_impressions, 8)):
            if n_steps is not None and i >= n_steps:
                break
            model.validation_step(mo.step_batch(bhv, lo, min(lo + 8, bhv.n_impressions)), i)
        model.on_validation_epoch_end()
        if epoch == 0:
            first = model.logged["val/loss"]
            assert abs(first - float(np.mean(z["step_losses"][:2]))) <= 1e-5 * abs(first)
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(model.logged["val/" + k] - float(z["test_" + k])) <= 1e-6, k
    assert abs(model.logged["val/loss"] - float(z["test_loss"])) <= 1e-5 * abs(float(z["test_loss"]))
    assert model.logged["val/loss_best"] == min(first, model.logged["val/loss"])


@pytest.mark.gpu
def test_ensemble_step_mode_dropin_logs_the_reference_values(golden_dir):
    from manner_b200.modules import B200EvalMixin

    class FakeEnsemble(B200EvalMixin, torch.nn.Module):
        _b200_zscore = True
        _b200_with_auc = False

        def __init__(self, tables, weights):
            super().__init__()
            self.encs = torch.nn.ModuleList([TableEncoder(t) for t in tables])
            self.w, self.logged = weights, {}

        def _b200_encoders(self):
            return [e for e, w in zip(self.encs, self.w) if w != 0]

        def _b200_weights(self):
            return [w for w in self.w if w != 0]

        def log_dict(self, values, **kw):
            self.logged.update(values)

    z = np.load(os.path.join(golden_dir, "ensemble_d128.npz"))
    bhv = _bhv(z)
    aspects = {"category": z["category"], "sentiment": z["sentiment"]}
    tabs = [torch.from_numpy(z[f"table{m}"]) for m in range(3)]
    for w, (wc, ws) in enumerate(z["weightings"].tolist()):
        model = FakeEnsemble(tabs, [1.0, wc, ws]).cuda()
        for i, lo in enumerate(range(0, bhv.n_impressions, 8)):
            model.test_step(mo.step_batch(bhv, lo, min(lo + 8, bhv.n_impressions), aspects), i)
        model.on_test_epoch_end()
        for k in ("ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
                  "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
            assert abs(model.logged["test/" + k] - float(z[f"w{w}_test_{k}"])) <= 1e-6, (w, k)


@pytest.mark.gpu
def test_ensemble_aspect_weight_sweep_logs_every_weighting(golden_dir):
    """`model.aspect_weights` (EnsembleModuleB200(aspect_weights=[[wc, ws], ...])): all weightings of the golden from ONE pass per
    step, logged as test/<metric>/w<i> -- each equal to what the reference's EnsembleModule logged when it was run with that
    (categ_weight, sent_weight) pair."""
    from manner_b200.modules import B200EvalMixin

    class FakeSweep(B200EvalMixin, torch.nn.Module):
        _b200_zscore = True
        _b200_with_auc = False

        def __init__(self, tables, grid):
            super().__init__()
            self.encs = torch.nn.ModuleList([TableEncoder(t) for t in tables])
            self.grid, self.logged = grid, {}

        def _b200_encoders(self):
            return list(self.encs)

        def _b200_weights(self):
            return [[1.0, wc, ws] for wc, ws in self.grid]  # what EnsembleModuleB200._b200_weights returns for a sweep

        def log_dict(self, values, **kw):
            self.logged.update(values)

    z = np.load(os.path.join(golden_dir, "ensemble_d128.npz"))
    bhv = _bhv(z)
    aspects = {"category": z["category"], "sentiment": z["sentiment"]}
    tabs = [torch.from_numpy(z[f"table{m}"]) for m in range(3)]
    grid = z["weightings"].tolist()
    model = FakeSweep(tabs, grid).cuda()
    for i, lo in enumerate(range(0, bhv.n_impressions, 8)):
        model.test_step(mo.step_batch(bhv, lo, min(lo + 8, bhv.n_impressions), aspects), i)
    model.on_test_epoch_end()
    for w in range(len(grid)):
        for k in ("ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
                  "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
            assert abs(model.logged[f"test/{k}/w{w}"] - float(z[f"w{w}_test_{k}"])) <= 1e-6, (w, k)
    assert "test/ndcg@5" not in model.logged  # a sweep logs per-weighting keys only


@pytest.mark.gpu
def test_baseline_metrics_mixin_logs_what_the_reference_baseline_logged(golden_dir):
    """B200MetricsMixin on the epoch-end seam of the reference's nine baselines: the golden holds what the reference's own
    NRMSPLMModule.on_test_epoch_end (nrms_plm_module.py:275-313) logged for these step outputs."""
    from manner_b200.modules import B200MetricsMixin

    class FakeBaseline(B200MetricsMixin, torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.keys = ["preds", "targets", "cand_news_size", "hist_news_size", "target_categories", "target_sentiments",
                         "hist_categories", "hist_sentiments"]  # nrms_plm_module.py:105-111
            self.test_step_outputs = {k: [] for k in self.keys}
            self.hparams = {"num_categ_classes": 19, "num_sent_classes": 4}
            self.logged = {}

        def log_dict(self, values, **kw):
            self.logged.update(values)

    z = np.load(os.path.join(golden_dir, "baseline_epoch_end.npz"))
    bhv = _bhv(z)
    model = FakeBaseline()
    ho, co = bhv.hist_offsets, bhv.cand_offsets
    for lo in range(0, bhv.n_impressions, 8):  # what the baseline's test_step appends (:258-272): preds on the device, sizes on the host
        hi = min(lo + 8, bhv.n_impressions)
        c_ids, h_ids = bhv.cand_ids[co[lo]:co[hi]], bhv.hist_ids[ho[lo]:ho[hi]]
        out = model.test_step_outputs
        out["preds"].append(torch.from_numpy(z["preds"][co[lo]:co[hi]]).cuda())
        out["targets"].append(torch.from_numpy(bhv.labels[co[lo]:co[hi]].astype(np.int64)).cuda())
        out["cand_news_size"].append(torch.from_numpy(np.diff(co[lo:hi + 1]).astype(np.int64)))
        out["hist_news_size"].append(torch.from_numpy(np.diff(ho[lo:hi + 1]).astype(np.int64)))
        out["target_categories"].append(torch.from_numpy(z["category"][c_ids].astype(np.int64)).cuda())
        out["target_sentiments"].append(torch.from_numpy(z["sentiment"][c_ids].astype(np.int64)).cuda())
        out["hist_categories"].append(torch.from_numpy(z["category"][h_ids].astype(np.int64)).cuda())
        out["hist_sentiments"].append(torch.from_numpy(z["sentiment"][h_ids].astype(np.int64)).cuda())
    model.on_test_epoch_end()
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
              "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
        assert abs(model.logged["test/" + k] - float(z["test_" + k])) <= 1e-6, (k, model.logged["test/" + k], float(z["test_" + k]))
    assert all(len(v) == 0 for v in model.test_step_outputs.values())  # buffers cleared like the reference does


@pytest.mark.gpu
def test_cached_mode_table_builder_feeds_the_evaluator(golden_dir):
    from manner_b200.evaluator import ScoreEvaluator

    z = np.load(os.path.join(golden_dir, "cr_d768.npz"))
    table = torch.from_numpy(z["table"])
    n_news, dim = table.shape
    enc = TableEncoder(table).cuda()
    batches = [{"news_row": torch.arange(lo, min(lo + 32, n_news))} for lo in range(0, n_news, 32)]
    built = mcache.build_embedding_table(enc, batches, n_news, dim, torch.device("cuda:0"))
    assert torch.equal(built.cpu(), table)
    # behaviours in the reference's textual form -> CSR -> one evaluation call for the whole epoch
    ids = [f"N{i}" for i in range(n_news)]
    nid2row = mcache.news_row_map(ids)
    bhv = _bhv(z)
    hist = [[ids[j] for j in bhv.hist_ids[bhv.hist_offsets[i]:bhv.hist_offsets[i + 1]]] for i in range(bhv.n_impressions)]
    cand = [[ids[j] for j in bhv.cand_ids[bhv.cand_offsets[i]:bhv.cand_offsets[i + 1]]] for i in range(bhv.n_impressions)]
    labs = [bhv.labels[bhv.cand_offsets[i]:bhv.cand_offsets[i + 1]].tolist() for i in range(bhv.n_impressions)]
    csr = mcache.behaviours_to_csr(hist, cand, labs, nid2row)
    ev = ScoreEvaluator([built])
    m = ev.evaluate(ev.upload(csr), pooled_auc=True).metrics()
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(m["test/" + k] - float(z["test_" + k])) <= 1e-6, k


class _FakeLoader:
    def __init__(self, dataset, collate_fn, batch_size):
        self.dataset, self.collate_fn, self.batch_size = dataset, collate_fn, batch_size


def _fake_datamodule(z, with_aspects=False):
    """The parts of the reference's datamodule the cached mode touches: the test dataset's ``news`` / ``behaviors`` frames and
    ``max_history_length`` (mind_rec_dataset.py:81-99) and the collate's ``_tokenize_df`` (:146-168)."""
    import pandas as pd

    n_news = z["table0" if "table0" in z.files else "table"].shape[0]
    ids = [f"N{i}" for i in range(n_news)]
    news = pd.DataFrame({"row": np.arange(n_news)}, index=ids)
    if with_aspects:
        news["category_label"], news["sentiment_label"] = z["category"], z["sentiment"]
    ho, co = z["hist_offsets"], z["cand_offsets"]
    beh = pd.DataFrame({
        "history": [[ids[j] for j in z["hist_ids"][ho[i]:ho[i + 1]]] for i in range(len(ho) - 1)],
        "candidates": [[ids[j] for j in z["cand_ids"][co[i]:co[i + 1]]] for i in range(len(co) - 1)],
        "labels": [z["labels"][co[i]:co[i + 1]].tolist() for i in range(len(co) - 1)],
    })
    ds = types.SimpleNamespace(news=news, behaviors=beh, max_history_length=50)
    collate = types.SimpleNamespace(_tokenize_df=lambda df: {"news_row": torch.from_numpy(df["row"].to_numpy().copy())})
    loader = _FakeLoader(ds, collate, 8)
    return types.SimpleNamespace(test_dataloader=lambda: loader, val_dataloader=lambda: loader)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cr_d768", "cr_ef_d128"])
def test_cached_mode_scores_the_epoch_in_one_call(golden_dir, name):
    """scorer=b200_cached: table built from the unique news in on_test_start, behaviours frame -> CSR, one evaluation call;
    test_step is a no-op.  Logs the values the reference logged."""
    from manner_b200.modules import B200EvalMixin

    z = np.load(os.path.join(golden_dir, name + ".npz"))
    early = "att_weight" in z.files
    calls = {"encoder": 0}

    class CountingEncoder(TableEncoder):
        def forward(self, x):
            calls["encoder"] += 1
            return super().forward(x)

    class FakeCR(B200EvalMixin, torch.nn.Module):
        _b200_cached = True
        _b200_news_batch = 32

        def __init__(self, table):
            super().__init__()
            self.news_encoder = CountingEncoder(table)
            if early:
                self.att = torch.nn.ParameterList([torch.nn.Parameter(torch.from_numpy(z[k])) for k in ("att_weight", "att_bias", "att_query")])
            self.logged = {}

        def _b200_encoders(self):
            return [self.news_encoder]

        def _b200_loss(self):
            return "ce"

        def _b200_attention(self):
            return [tuple(self.att)] if early else None

        def log_dict(self, values, **kw):
            self.logged.update(values)

    model = FakeCR(torch.from_numpy(z["table"])).cuda()
    model.trainer = types.SimpleNamespace(datamodule=_fake_datamodule(z))
    model.on_test_start()
    n_encoder_calls = calls["encoder"]
    model.test_step({"never": "looked at"}, 0)
    model.on_test_epoch_end()
    assert calls["encoder"] == n_encoder_calls  # test_step did not touch the encoder
    for k in ("auc", "mrr", "ndcg@5", "ndcg@10"):
        assert abs(model.logged["test/" + k] - float(z["test_" + k])) <= 1e-6, k
    assert abs(model.logged["test/loss"] - float(z["test_loss"])) <= 1e-5 * abs(float(z["test_loss"]))
    # validation in cached mode: same numbers under val/, val/loss_best tracked across epochs
    model.on_validation_start()
    model.validation_step({"never": "looked at"}, 0)
    model.on_validation_epoch_end()
    assert abs(model.logged["val/ndcg@10"] - float(z["test_ndcg@10"])) <= 1e-6
    assert model.logged["val/loss_best"] == model.logged["val/loss"]


@pytest.mark.gpu
def test_cached_mode_ensemble_with_aspects(golden_dir):
    from manner_b200.modules import B200EvalMixin

    z = np.load(os.path.join(golden_dir, "ensemble_d128.npz"))

    class FakeEnsemble(B200EvalMixin, torch.nn.Module):
        _b200_zscore = True
        _b200_with_auc = False
        _b200_cached = True

        def __init__(self, tables, weights):
            super().__init__()
            self.encs = torch.nn.ModuleList([TableEncoder(t) for t in tables])
            self.w, self.logged = weights, {}

        def _b200_encoders(self):
            return [e for e, w in zip(self.encs, self.w) if w != 0]

        def _b200_weights(self):
            return [w for w in self.w if w != 0]

        def log_dict(self, values, **kw):
            self.logged.update(values)

    tabs = [torch.from_numpy(z[f"table{m}"]) for m in range(3)]
    for w, (wc, ws) in enumerate(z["weightings"].tolist()):
        model = FakeEnsemble(tabs, [1.0, wc, ws]).cuda()
        model.trainer = types.SimpleNamespace(datamodule=_fake_datamodule(z, with_aspects=True))
        model.on_test_start()
        model.on_test_epoch_end()
        for k in ("ndcg@5", "ndcg@10", "categ_div@5", "categ_div@10", "sent_div@5", "sent_div@10",
                  "categ_pers@5", "categ_pers@10", "sent_pers@5", "sent_pers@10"):
            assert abs(model.logged["test/" + k] - float(z[f"w{w}_test_{k}"])) <= 1e-6, (w, k)


def test_unique_news_ids_follow_the_history_truncation():
    import pandas as pd

    beh = pd.DataFrame({"history": [["N3", "N1", "N9"], "['N1', 'N4']"], "candidates": [["N5"], "['N3', 'N6']"]})
    assert mcache.unique_news_ids(beh, max_history_length=2) == ["N3", "N1", "N5", "N4", "N6"]  # N9 is cut off
