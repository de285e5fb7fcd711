"""Drop-in LightningModules for the reference's Hydra seam.

    # configs/model/cr_module_b200.yaml        _target_: manner_b200.modules.CRModuleB200
    # configs/model/ensemble_module_b200.yaml  _target_: manner_b200.modules.EnsembleModuleB200

``CRModuleB200`` / ``EnsembleModuleB200`` subclass the reference's ``CRModule`` / ``EnsembleModule``
(manner/models/cr_module.py:19, ensemble_module.py:17) -- same constructor keywords (plus defaulted
ones; the YAMLs inherit the reference's model configs through Hydra's `defaults` list), same ``news_encoder.*`` parameters, so existing experiment YAMLs and Lightning checkpoints
load unchanged -- and replace only the test path: ``test_step`` / ``on_test_epoch_end`` (and
``validation_step`` metrics for the CR module) hand the batch's news vectors to the fused sm_100a
kernel instead of ``to_dense_batch`` + per-row Python loops + ``bmm`` + torchmetrics' group loops.

The evaluation logic lives in ``B200EvalMixin`` and ``StepScorer``, which do not need Lightning: the
reference package and its dependencies (lightning, torchmetrics, torch_geometric, ...) are optional
imports, so this file is importable -- and testable -- without them.

Two ways to feed the kernel (constructor keyword / YAML key ``scorer``):
  * step mode (``scorer: b200``, default): the unchanged DataModule delivers ``MINDRecBatch`` es (mind_batch.py:6-12);
    the module runs its PLM ``news_encoder`` on ``x_hist`` / ``x_cand`` exactly as the reference does
    (cr_module.py:107,113) and uses the resulting [sum H + sum C, D] matrix as the step's table;
  * cached mode (``scorer: b200_cached``): in ``on_test_start`` the module runs its news encoder(s) once over the UNIQUE
    news of the split (``manner_b200.cache``), converts the datamodule's behaviours frame to CSR and scores the whole epoch
    with one ``ScoreEvaluator.evaluate`` call; ``test_step`` then does nothing -- the PLM leaves the loop (65 k encoder
    rows instead of 4.3 M on MIND-small).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import _native as nat
from .evaluator import SLOT_KEYS

try:  # the reference package is optional: it needs lightning, torchmetrics, torch_geometric, ...
    from manner.models.cr_module import CRModule as _RefCRModule  # type: ignore
    from manner.models.ensemble_module import EnsembleModule as _RefEnsembleModule  # type: ignore

    HAVE_REFERENCE = True
except Exception:  # pragma: no cover - depends on the environment
    HAVE_REFERENCE = False

    class _Unavailable(torch.nn.Module):
        def __init__(self, *args: Any, **kwargs: Any) -> None:
            raise ImportError(
                "manner_b200.modules needs the reference package `manner` (andreeaiana/manner) and its dependencies "
                "(lightning, torchmetrics, torch_geometric, pytorch_metric_learning) on PYTHONPATH; "
                "use manner_b200.ScoreEvaluator directly when they are not installed"
            )

    _RefCRModule = _RefEnsembleModule = _Unavailable  # type: ignore


def _segment_offsets(seg: Tensor, n: int) -> Tensor:
    """Sorted segment ids (MINDCollate's ``repeat_interleave(arange(B), sizes)``,
    mind_rec_dataset.py:171-174) -> int32 CSR offsets [n + 1], on the ids' device, no host sync."""
    counts = torch.bincount(seg, minlength=n)
    off = torch.zeros(n + 1, dtype=torch.int32, device=seg.device)
    off[1:] = torch.cumsum(counts, 0)
    return off


class StepScorer:
    """Scores one MINDRecBatch from already-encoded news vectors and accumulates epoch statistics.

    Replaces, per step, cr_module.py:108-131,173-182 / ensemble_module.py:116-149,155-192 and, at the
    end, cr_module.py:266-274 / ensemble_module.py:214-238."""

    def __init__(self, zscore: bool, ks=(5, 10), with_auc: bool = True, num_categ_classes: int = 19, num_sent_classes: int = 4,
                 loss: Optional[str] = None, temperature: float = 0.1) -> None:
        nat.lib()
        self.zscore, self.ks, self.with_auc = zscore, (int(ks[0]), int(ks[1])), with_auc
        self.num_categ_classes, self.num_sent_classes = num_categ_classes, num_sent_classes
        self.loss_kind = {None: nat.LOSS_NONE, "ce": nat.LOSS_CE, "supcon": nat.LOSS_SUPCON}[loss]
        self.temperature = float(temperature)
        self.reset()

    def reset(self) -> None:
        self.loss_stats: Optional[Tensor] = None  # fp64 [2]: sum of the step losses, number of steps
        self.sums: Optional[Tensor] = None
        self.n_impressions = 0
        self.flags: Optional[Tensor] = None
        self.preds: List[Tensor] = []
        self.targets: List[Tensor] = []
        self.has_aspects = False

    @torch.no_grad()
    def step(self, hist_vecs: Sequence[Tensor], cand_vecs: Sequence[Tensor], batch: Dict[str, Any], weights: Optional[Sequence[float]] = None,
             attention: Optional[Sequence[Optional[Sequence[Tensor]]]] = None) -> Tensor:
        """``hist_vecs[m]`` / ``cand_vecs[m]``: module m's news vectors for ``batch['x_hist']`` / ``['x_cand']``
        ([sum H, D] / [sum C, D]).  ``attention[m]`` = (linear.weight, linear.bias, query) of module m's additive
        attention for early fusion (cr_module.py:124-125), None for late fusion.  Returns the flat combined scores
        [sum C] (the reference's ``preds``)."""
        dev = cand_vecs[0].device
        batch_hist, batch_cand = batch["batch_hist"].to(dev), batch["batch_cand"].to(dev)
        n_hist, n_cand = hist_vecs[0].shape[0], cand_vecs[0].shape[0]
        # B = batch.max() + 1 as to_dense_batch derives it, and the longest candidate list (it sizes the kernel's per-warp shared
        # memory: the step's TOTAL candidate count would cap eval batches at ~3 000 candidates) -- the one host read of the step
        n_impr, max_cand = (int(v) for v in torch.stack([torch.maximum(batch_hist.max(), batch_cand.max()) + 1, torch.bincount(batch_cand).max()]).tolist())
        hist_off, cand_off = _segment_offsets(batch_hist, n_impr), _segment_offsets(batch_cand, n_impr)
        # the step's table is [history rows; candidate rows]: ids are just positions
        tables = [torch.cat([h, c]).float().contiguous() for h, c in zip(hist_vecs, cand_vecs)]
        hist_ids = torch.arange(n_hist, dtype=torch.int32, device=dev)
        cand_ids = torch.arange(n_hist, n_hist + n_cand, dtype=torch.int32, device=dev)
        labels = (batch["labels"].to(dev) != 0).to(torch.uint8)
        w_dev = None
        if weights is not None:  # one weighting [n_modules] or a sweep [[...], ...] re-scored from the one gather
            w_dev = torch.tensor(weights, dtype=torch.float32, device=dev).reshape(-1, len(tables))
        active = (1 << len(tables)) - 1
        cat = sent = None
        if "category" in batch["x_cand"] and "sentiment" in batch["x_cand"] and self.zscore:
            cat = torch.cat([batch["x_hist"]["category"], batch["x_cand"]["category"]]).to(dev, torch.int32)
            sent = torch.cat([batch["x_hist"]["sentiment"], batch["x_cand"]["sentiment"]]).to(dev, torch.int32)
            self.has_aspects = True
        attn_logits: List[Optional[Tensor]] = []
        hist_pad = cand_pad = None
        if attention is not None and any(a is not None for a in attention):
            from . import ops

            attn_logits = [None if a is None else ops.attention_logits(t, *a) for t, a in zip(tables, attention)]
        if attn_logits or self.loss_kind != nat.LOSS_NONE:
            # the step IS the reference's dense batch: every impression is padded to the step's longest (to_dense_batch)
            h_len, c_len = hist_off[1:] - hist_off[:-1], cand_off[1:] - cand_off[:-1]
            hist_pad, cand_pad = (h_len.max() - h_len).to(torch.int32), (c_len.max() - c_len).to(torch.int32)
        scores, _, sums, flags, loss_per_impr = torch.ops.manner_b200.score_eval(
            tables, hist_off, hist_ids, cand_off, cand_ids, labels, w_dev, self.zscore, max(max_cand, 1), active,
            self.ks[0], self.ks[1], True, 0, False, cat, sent, self.num_categ_classes, self.num_sent_classes, attn_logits,
            False, hist_pad if attn_logits else None, self.loss_kind, self.temperature, cand_pad if self.loss_kind != nat.LOSS_NONE else None,
        )
        if self.loss_kind != nat.LOSS_NONE:
            from . import ops

            st = ops.step_loss(loss_per_impr, n_impr, self.loss_kind, cand_off, labels)  # this step's value for the MeanMetric (cr_module.py:255-259)
            self.loss_stats = st if self.loss_stats is None else self.loss_stats + st
        self.sums = sums if self.sums is None else self.sums + sums
        self.flags = flags if self.flags is None else self.flags | flags
        self.n_impressions += n_impr
        if self.with_auc:
            self.preds.append(scores), self.targets.append(labels)
        return scores

    def compute(self, prefix: str = "test/") -> Dict[str, float]:
        """Epoch means under the reference's log keys (cr_module.py:79-89, ensemble_module.py:50-84)."""
        if self.sums is None:
            return {}
        out: Dict[str, float] = {}
        auc_stats = None
        if self.with_auc and self.preds:
            # AUROC's any-outside-[0,1] sigmoid rule is decided over the whole epoch (cr_module.py:273)
            auc_stats = torch.ops.manner_b200.pooled_auc(torch.cat(self.preds), torch.cat(self.targets), 2, self.flags)
        sums = self.sums.cpu().numpy()
        flags = int(self.flags.cpu().item())
        if flags & (nat.FLAG_BAD_ID | nat.FLAG_CAND_OVERFLOW | nat.FLAG_BAD_ASPECT):
            raise nat.NativeError(f"manner_b200 kernels flagged bad input (flags={flags})")
        n = max(self.n_impressions, 1)
        for w in range(sums.shape[0]):
            # one weighting: the reference's keys; a sweep (EnsembleModuleB200(aspect_weights=...)): `<key>/w<i>` per weighting
            suffix = "" if sums.shape[0] == 1 else f"/w{w}"
            s = sums[w]
            for slot, key in SLOT_KEYS.items():
                if slot >= nat.M_CATEG_DIV_K0 and not self.has_aspects:
                    continue
                out[prefix + key.format(k0=self.ks[0], k1=self.ks[1]) + suffix] = float(s[slot] / n)
            out[prefix + "gauc" + suffix] = float(s[nat.M_GAUC] / s[nat.M_GAUC_VALID]) if s[nat.M_GAUC_VALID] > 0 else 0.0
        if auc_stats is not None:
            out[prefix + "auc"] = float(auc_stats.cpu()[0])
        if self.loss_stats is not None:
            ls = self.loss_stats.cpu().numpy()
            out[prefix + "loss"] = float(ls[0] / ls[1]) if ls[1] > 0 else 0.0
        return out


class B200EvalMixin:
    """``test_step`` / ``on_test_epoch_end`` on the fused kernel.  Mixed in before the reference module
    (or any module exposing the same attributes) in the MRO."""

    _b200_zscore = False
    _b200_with_auc = True

    def _b200_scorer(self, stage: str = "test") -> StepScorer:
        scorers = self.__dict__.setdefault("_b200_step_scorers", {})
        if stage not in scorers:
            hp = getattr(self, "hparams", {})
            scorers[stage] = StepScorer(
                self._b200_zscore, with_auc=self._b200_with_auc,
                num_categ_classes=int(hp.get("num_categ_classes", 19)) if hasattr(hp, "get") else 19,
                num_sent_classes=int(hp.get("num_sent_classes", 4)) if hasattr(hp, "get") else 4,
                loss=self._b200_loss(), temperature=float(hp.get("temperature", 0.1)) if hasattr(hp, "get") else 0.1,
            )
        return scorers[stage]

    def _b200_loss(self) -> Optional[str]:
        """"ce" | "supcon" | None: the step loss logged as test/loss (cr_module.py:72-76,253-259)."""
        return None

    def _b200_attention(self) -> Optional[List[Optional[Sequence[Tensor]]]]:
        """Per module: (linear.weight, linear.bias, query) of its early-fusion additive attention, or None."""
        return None

    def _b200_encoders(self) -> List[torch.nn.Module]:  # pragma: no cover - overridden
        raise NotImplementedError

    def _b200_weights(self) -> Optional[List[float]]:
        return None

    def _b200_step(self, stage: str, batch: Dict[str, Any]) -> None:
        encoders = self._b200_encoders()
        hist = [enc(batch["x_hist"]) for enc in encoders]
        cand = [enc(batch["x_cand"]) for enc in encoders]
        self._b200_scorer(stage).step(hist, cand, batch, self._b200_weights(), self._b200_attention())

    def _b200_epoch_end(self, stage: str) -> Dict[str, float]:
        scorer = self._b200_scorer(stage)
        values = scorer.compute(prefix=stage + "/")
        scorer.reset()
        return values

    # ---- cached mode: the PLM runs once per unique news, the whole epoch is ONE evaluation call ----
    _b200_cached = False       # scorer == "b200_cached"
    _b200_news_batch = 256     # news per encoder call while the table is built

    def _b200_evaluate_cached(self, stage: str) -> Dict[str, float]:
        """Builds the embedding table(s) with the module's own news encoders over the UNIQUE news of the split, converts
        the split's behaviours frame to CSR and scores the whole epoch with one ``ScoreEvaluator.evaluate`` call.  Uses
        only what the reference's datamodule already has: the dataset's ``news`` / ``behaviors`` frames and the collate's
        tokeniser (mind_rec_dataset.py:81-99,114-168)."""
        from . import cache
        from .evaluator import ScoreEvaluator

        dm = self.trainer.datamodule
        loader = dm.test_dataloader() if stage == "test" else dm.val_dataloader()
        ds, collate = loader.dataset, loader.collate_fn
        step = int(getattr(loader, "batch_size", 8) or 8)
        device = getattr(self, "device", None)  # LightningModule.device; plain modules: wherever their tensors live
        if not isinstance(device, torch.device):
            import itertools

            device = next(itertools.chain(self.parameters(), self.buffers())).device
        encoders = self._b200_encoders()
        # on-disk cache of the tables / id map / CSR (test stage only: during training the encoder weights change between
        # validations).  The key digests what the tables depend on; a cache built from other checkpoints is ignored.
        cache_dir = getattr(self, "_b200_cache_dir", None) if stage == "test" else None
        cache_key = cache.fingerprint(*self._b200_cache_sources(), len(ds.behaviors), ds.max_history_length, len(encoders)) if cache_dir else ""
        cached = cache.load_cache(cache_dir, cache_key, device) if cache_dir else None
        if cached is not None:
            tables, news_ids, bhv, asp = cached
            return self._b200_evaluate_tables(stage, tables, bhv, asp, step, device)
        news_ids = cache.unique_news_ids(ds.behaviors, ds.max_history_length)
        nid2row = cache.news_row_map(news_ids)

        def news_batches():
            for lo in range(0, len(news_ids), self._b200_news_batch):
                yield collate._tokenize_df(ds.news.loc[news_ids[lo : lo + self._b200_news_batch]])

        tables = []
        for enc in encoders:
            first = next(iter(news_batches()))
            with torch.no_grad():
                was = enc.training
                enc.eval()
                dim = int(enc(cache._to_device(first, device)).shape[1])
                enc.train(was)
            tables.append(cache.build_embedding_table(enc, news_batches(), len(news_ids), dim, device))
        bhv = cache.behaviours_frame_to_csr(ds.behaviors, nid2row, ds.max_history_length)
        asp = None
        if self._b200_zscore and "category_label" in ds.news.columns and "sentiment_label" in ds.news.columns:
            asp = cache.aspect_arrays(ds.news, nid2row)
        if cache_dir:
            cache.save_cache(cache_dir, tables, news_ids, bhv, cache_key, asp, ds.max_history_length)
        return self._b200_evaluate_tables(stage, tables, bhv, asp, step, device)

    def _b200_cache_sources(self) -> list:
        """What the cached tables were computed from (digested into the cache key): checkpoint paths when the module has them."""
        hp = getattr(self, "hparams", {})
        get = hp.get if hasattr(hp, "get") else (lambda k, d=None: d)
        return [get(k) for k in ("cr_module_module_ckpt", "a_module_categ_ckpt", "a_module_sent_ckpt", "plm_model")]

    def _b200_evaluate_tables(self, stage: str, tables, bhv, asp, step: int, device) -> Dict[str, float]:
        from .evaluator import ScoreEvaluator

        aspects = {}
        if asp is not None:
            hp = getattr(self, "hparams", {})
            aspects = dict(news_category=asp[0], news_sentiment=asp[1],
                           num_categ_classes=int(hp.get("num_categ_classes", 19)) if hasattr(hp, "get") else 19,
                           num_sent_classes=int(hp.get("num_sent_classes", 4)) if hasattr(hp, "get") else 4)
        ev = ScoreEvaluator(tables, device, attention=self._b200_attention(), **aspects)
        weights = self._b200_weights()
        hp = getattr(self, "hparams", {})
        sweep = weights is not None and len(weights) > 0 and isinstance(weights[0], (list, tuple))
        res = ev.evaluate(
            ev.upload(bhv, step_batch=step, pipelined=True), weights=None if weights is None else (weights if sweep else [weights]), zscore=self._b200_zscore,
            pooled_auc=self._b200_with_auc, loss=None if self._b200_zscore else self._b200_loss(),
            temperature=float(hp.get("temperature", 0.1)) if hasattr(hp, "get") else 0.1,
        )
        if not sweep:
            return res.metrics(prefix=stage + "/")
        values: Dict[str, float] = {}
        for w in range(len(weights)):
            values.update({k + f"/w{w}": v for k, v in res.metrics(weighting=w, prefix=stage + "/").items()})
        return values

    def on_test_start(self) -> None:
        if self._b200_cached:
            self.__dict__["_b200_cached_values"] = self._b200_evaluate_cached("test")

    def test_step(self, batch: Dict[str, Any], batch_idx: int) -> None:
        if self._b200_cached:
            return  # the epoch was scored in on_test_start; the batch (and its tokenisation) is not needed
        self._b200_step("test", batch)

    def on_test_epoch_end(self) -> None:
        values = self.__dict__.pop("_b200_cached_values", None) if self._b200_cached else self._b200_epoch_end("test")
        self.log_dict(values or {}, on_step=False, on_epoch=True, prog_bar=True, logger=True)

    def on_validation_start(self) -> None:
        if self._b200_cached:  # the encoder weights changed since the last validation: the table is rebuilt every time
            self.__dict__["_b200_cached_val_values"] = self._b200_evaluate_cached("val")

    def validation_step(self, batch: Dict[str, Any], batch_idx: int) -> None:
        """cr_module.py:214-225: val/loss is what `ModelCheckpoint(monitor="val/loss")` watches (configs/callbacks/default.yaml)."""
        if self._b200_cached:
            return
        self._b200_step("val", batch)

    def on_validation_epoch_end(self) -> None:
        """cr_module.py:227-251: val/loss, the best val/loss so far (MinMetric), then the validation metrics."""
        values = (self.__dict__.pop("_b200_cached_val_values", None) or {}) if self._b200_cached else self._b200_epoch_end("val")
        if "val/loss" in values:
            best = min(self.__dict__.get("_b200_val_loss_best", float("inf")), values["val/loss"])
            self.__dict__["_b200_val_loss_best"] = best
            values["val/loss_best"] = best
        self.log_dict(values, on_step=False, on_epoch=True, prog_bar=True, logger=True)

    def on_train_start(self) -> None:
        """cr_module.py:133-138: forget what Lightning's sanity-check validation steps accumulated."""
        self.__dict__.pop("_b200_val_loss_best", None)
        if "val" in self.__dict__.get("_b200_step_scorers", {}):
            self._b200_scorer("val").reset()
        parent = getattr(super(), "on_train_start", None)
        if callable(parent):
            parent()


class CRModuleB200(B200EvalMixin, _RefCRModule):
    """CRModule (late or early fusion, CE or SupCon loss) with the B200 test path.  Extra keyword: ``scorer``
    ("b200" | "reference") to fall back to the reference's own test path for A/B comparison."""

    def __init__(self, *args: Any, scorer: str = "b200", cache_dir: Optional[str] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if scorer not in ("b200", "b200_cached", "reference"):
            raise ValueError("scorer must be 'b200' (per step), 'b200_cached' (table built once, one call per epoch) or 'reference'")
        self._b200_enabled = scorer != "reference"
        self._b200_cached = scorer == "b200_cached"
        self._b200_cache_dir = cache_dir  # cached mode: keep tables + nid -> row map + CSR on disk between runs

    def _b200_encoders(self) -> List[torch.nn.Module]:
        return [self.news_encoder]

    def _b200_loss(self) -> Optional[str]:
        return "supcon" if self.hparams.supcon_loss else "ce"

    def _b200_attention(self) -> Optional[List[Optional[Sequence[Tensor]]]]:
        if self.hparams.late_fusion:
            return None
        att = self.user_encoder.additive_attention  # NAMLUserEncoder (user_encoder.py:9-21) -> AdditiveAttention (attention.py:6-13)
        return [(att.linear.weight, att.linear.bias, att.query)]

    def test_step(self, batch: Dict[str, Any], batch_idx: int) -> None:
        if not self._b200_enabled:
            return _RefCRModule.test_step(self, batch, batch_idx)
        return B200EvalMixin.test_step(self, batch, batch_idx)

    def on_test_epoch_end(self) -> None:
        if not self._b200_enabled:
            return _RefCRModule.on_test_epoch_end(self)
        return B200EvalMixin.on_test_epoch_end(self)

    def validation_step(self, batch: Dict[str, Any], batch_idx: int) -> None:
        if not self._b200_enabled:
            return _RefCRModule.validation_step(self, batch, batch_idx)
        with torch.no_grad():
            return B200EvalMixin.validation_step(self, batch, batch_idx)

    def on_validation_epoch_end(self) -> None:
        if not self._b200_enabled:
            return _RefCRModule.on_validation_epoch_end(self)
        return B200EvalMixin.on_validation_epoch_end(self)


class EnsembleModuleB200(B200EvalMixin, _RefEnsembleModule):
    """EnsembleModule (CR + category / sentiment A-Modules, z-score, aspect weights) with the B200 test
    path.  Logs the reference's keys (ndcg, *_div, *_pers) plus mrr / gauc.

    Extra keywords (defaulted, so the reference's YAMLs and `load_from_checkpoint` work unchanged):
      ``scorer``         "b200" (per step) | "b200_cached" (tables built once, one call per epoch)
      ``aspect_weights`` ``[[categ_weight, sent_weight], ...]`` (YAML key ``model.aspect_weights``): an aspect-weight SWEEP --
                         every weighting is re-scored from ONE gather of the step / epoch (BASELINE.json configs[3]) instead of
                         one `python manner/train.py ... model.categ_weight=...` run per weighting; metrics are logged as
                         ``test/<metric>/w<i>``.  A-Modules a weighting needs are loaded even when ``categ_weight`` /
                         ``sent_weight`` are 0 (ensemble_module.py:37-46 loads them only for non-zero weights)."""

    _b200_zscore = True
    _b200_with_auc = False  # the reference's EnsembleModule has no AUROC (ensemble_module.py:50-55)

    def __init__(self, *args: Any, scorer: str = "b200", aspect_weights: Optional[Sequence[Sequence[float]]] = None,
                 cache_dir: Optional[str] = None, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        if scorer not in ("b200", "b200_cached"):
            raise ValueError("scorer must be 'b200' (per step) or 'b200_cached' (tables built once, one call per epoch)")
        self._b200_cached = scorer == "b200_cached"
        self._b200_cache_dir = cache_dir
        self._b200_aspect_weights: Optional[List[List[float]]] = None
        if aspect_weights is not None:
            grid = [[float(w) for w in pair] for pair in aspect_weights]
            if not grid or any(len(pair) != 2 for pair in grid):
                raise ValueError("aspect_weights must be a non-empty list of [categ_weight, sent_weight] pairs")
            self._b200_aspect_weights = grid
            # the sub-modules the sweep needs, loaded exactly as ensemble_module.py:37-46 loads them
            if any(wc != 0 for wc, _ in grid) and not hasattr(self, "a_module_categ"):
                from manner.models.a_module import AModule  # type: ignore

                assert isinstance(self.hparams.a_module_categ_ckpt, str), "aspect_weights with a category weight needs a_module_categ_ckpt"
                self.a_module_categ = AModule.load_from_checkpoint(checkpoint_path=self.hparams.a_module_categ_ckpt)
            if any(ws != 0 for _, ws in grid) and not hasattr(self, "a_module_sent"):
                from manner.models.a_module import AModule  # type: ignore

                assert isinstance(self.hparams.a_module_sent_ckpt, str), "aspect_weights with a sentiment weight needs a_module_sent_ckpt"
                self.a_module_sent = AModule.load_from_checkpoint(checkpoint_path=self.hparams.a_module_sent_ckpt)

    def _b200_use(self) -> tuple:
        """(category A-Module takes part, sentiment A-Module takes part)"""
        if self._b200_aspect_weights is not None:
            return any(wc != 0 for wc, _ in self._b200_aspect_weights), any(ws != 0 for _, ws in self._b200_aspect_weights)
        return self.hparams.categ_weight != 0, self.hparams.sent_weight != 0

    def _b200_encoders(self) -> List[torch.nn.Module]:
        use_c, use_s = self._b200_use()
        encs = [self.cr_module.news_encoder]
        if use_c:
            encs.append(self.a_module_categ.news_encoder)
        if use_s:
            encs.append(self.a_module_sent.news_encoder)
        return encs

    def validation_step(self, batch: Dict[str, Any], batch_idx: int) -> None:  # ensemble_module.py:199-200: the ensemble is never validated
        pass

    def on_validation_epoch_end(self) -> None:
        pass

    def _b200_weights(self):
        use_c, use_s = self._b200_use()
        if self._b200_aspect_weights is not None:
            return [[1.0] + ([wc] if use_c else []) + ([ws] if use_s else []) for wc, ws in self._b200_aspect_weights]
        w = [1.0]
        if use_c:
            w.append(float(self.hparams.categ_weight))
        if use_s:
            w.append(float(self.hparams.sent_weight))
        return w


class B200MetricsMixin:
    """Epoch-end metrics of the reference's nine BASELINE recommenders (manner/models/baselines/*_module.py) on the B200 kernels.

    The baselines keep their own model code; what they share with CRModule / EnsembleModule is the epoch end: `test_step`
    collects `preds`, `targets`, `cand_news_size`, `hist_news_size` and the aspect labels in `self.test_step_outputs`, and
    `on_test_epoch_end` feeds five torchmetrics collections (nrms_plm_module.py:275-313): auc / mrr / ndcg@5/10,
    categ_div / sent_div @5/10, categ_pers / sent_pers @5/10.  Mixed in BEFORE the baseline class

        class NRMSModuleB200(B200MetricsMixin, NRMSModule): pass

    this replaces only `on_test_epoch_end`: one `mb200_rank_metrics` launch + `mb200_pooled_auc` instead of torchmetrics'
    Python loop over every impression, same log keys (plus ``test/gauc``)."""

    def on_test_epoch_end(self) -> None:
        from .metrics import RetrievalMetricsB200

        outs = self.test_step_outputs
        cat = lambda key: torch.cat([o for o in outs[key]])  # noqa: E731  (nrms_plm_module.py:276-286)
        preds, targets = cat("preds"), cat("targets")
        cand_news_size, hist_news_size = cat("cand_news_size"), cat("hist_news_size")
        dev = preds.device
        if not preds.is_cuda:
            raise RuntimeError("B200MetricsMixin needs the module on a CUDA device (there is no CPU path)")
        cand_indexes = torch.arange(cand_news_size.shape[0], device=dev).repeat_interleave(cand_news_size.to(dev))
        hist_indexes = torch.arange(hist_news_size.shape[0], device=dev).repeat_interleave(hist_news_size.to(dev))
        hp = getattr(self, "hparams", {})
        metrics = RetrievalMetricsB200(
            prefix="test/",
            num_categ_classes=int(hp.get("num_categ_classes", 19)) if hasattr(hp, "get") else 19,
            num_sent_classes=int(hp.get("num_sent_classes", 4)) if hasattr(hp, "get") else 4,
        )
        metrics.update(preds, targets, indexes=cand_indexes, target_categories=cat("target_categories"), target_sentiments=cat("target_sentiments"),
                       hist_categories=cat("hist_categories"), hist_sentiments=cat("hist_sentiments"), hist_indexes=hist_indexes)
        self.log_dict(metrics.compute(), on_step=False, on_epoch=True, prog_bar=True, logger=True)
        for key in self.keys:  # clean memory for the next epoch (nrms_plm_module.py:311-313)
            outs[key].clear()
