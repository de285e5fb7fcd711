"""TEST INFRASTRUCTURE ONLY -- import shims that let the reference's OWN in-repo code execute here.

/root/reference (andreeaiana/manner) imports lightning, torch_geometric, torchmetrics,
pytorch_metric_learning, MulticoreTSNE, seaborn, colorcet and matplotlib at module top
(manner/models/cr_module.py:5-10, ensemble_module.py:4-7, a_module.py:3-12); none of them is
installed in this image and there is no network.  ``install()`` registers placeholder modules for
them in ``sys.modules``:

* the arithmetic the hot path needs from them (``to_dense_batch``, the torchmetrics metric classes
  and helpers) comes from the restatement in oracle/thirdparty.py;
* pytorch_metric_learning: the base class and helpers the reference's own SupConLoss (components/losses.py) builds on are
  restated from the published 2.1.1 sources, so `val/loss` / `test/loss` with supcon_loss=True come from the reference's code;
* everything off the path (tSNE, plotting) is an inert placeholder.

With the shims in place tests/golden/make_golden.py runs the reference's unmodified
``CRModule.forward/model_step/test_step/on_test_epoch_end``, ``EnsembleModule.*``, ``DotProduct``,
``manner.metrics.functional.*`` and ``Diversity``/``Personalization`` on table-lookup "news
encoders" and stores the outputs as golden vectors.  It is only ever used in this container (the
GPU box has no /root/reference); nothing under manner_b200/ imports it.
"""
from __future__ import annotations

import inspect
import sys
import types
from typing import Any, Dict

import torch

from . import thirdparty as tp


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class LightningModule(torch.nn.Module):
    """Inert stand-in for lightning.LightningModule: hparams capture, CPU device, log capture."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__()
        self.logged: Dict[str, Any] = {}

    @property
    def device(self) -> torch.device:
        return torch.device("cpu")

    @property
    def hparams(self) -> _AttrDict:
        if "_hparams" not in self.__dict__:
            self.__dict__["_hparams"] = _AttrDict()
        return self.__dict__["_hparams"]

    def save_hyperparameters(self, *args: Any, logger: bool = True, **kwargs: Any) -> None:
        frame = inspect.currentframe().f_back
        local = frame.f_locals
        # like Lightning: the arguments of the __init__ whose frame called us (the class that owns the frame), so a
        # subclass with an (*args, **kwargs) constructor still records the reference module's own keywords
        owner = local.get("__class__", type(self))
        names = [n for n in inspect.signature(owner.__init__).parameters if n != "self"]
        self.hparams.update({n: local[n] for n in names if n in local})

    def log(self, name: str, value: Any, *args: Any, **kwargs: Any) -> None:
        self.__dict__.setdefault("logged", {})[name] = value

    def log_dict(self, dictionary: Any, *args: Any, **kwargs: Any) -> None:
        values = dictionary.compute() if hasattr(dictionary, "compute") else dictionary
        for k, v in values.items():
            self.log(k, v)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, **kwargs: Any):  # pragma: no cover
        raise RuntimeError("checkpoints are not available here; make_golden.py patches this")


def _module(name: str, **attrs: Any) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Inert:
    def __init__(self, *args: Any, **kwargs: Any) -> None:
        pass

    def add_to_recordable_attributes(self, *args: Any, **kwargs: Any) -> None:
        pass


# ---- pytorch_metric_learning 2.1.1 (pinned `>=2.1.1`, requirements.txt): the pieces the reference's own SupConLoss subclass
# (manner/models/components/losses.py:6-40) inherits or calls, restated from the published 2.1.1 sources so that the reference's
# OWN compute_loss / _compute_loss run unmodified on top of them.  TEST INFRASTRUCTURE (golden generation only).


class PmlSupConLossBase(torch.nn.Module):
    """losses.SupConLoss as the reference subclasses it: BaseMetricLossFunction.forward (compute_loss -> reducer),
    GenericPairLoss.mat_based_loss (pos / neg masks from the index tuple), zero_losses, and the class default reducer
    AvgNonZeroReducer (mean of the element losses that are > 0, else 0)."""

    def __init__(self, temperature: float = 0.1, **kwargs: Any) -> None:
        super().__init__()
        self.temperature = temperature

    def add_to_recordable_attributes(self, *args: Any, **kwargs: Any) -> None:
        pass

    def zero_losses(self) -> Dict[str, Any]:
        return {"loss": {"losses": 0, "indices": None, "reduction_type": "already_reduced"}}

    def loss_method(self, mat: torch.Tensor, indices_tuple) -> Dict[str, Any]:  # GenericPairLoss.mat_based_loss
        a1, p, a2, n = indices_tuple
        pos_mask, neg_mask = torch.zeros_like(mat), torch.zeros_like(mat)
        pos_mask[a1, p] = 1
        neg_mask[a2, n] = 1
        return self._compute_loss(mat, pos_mask, neg_mask)

    def forward(self, embeddings, labels=None, indices_tuple=None, ref_emb=None, ref_labels=None) -> torch.Tensor:
        loss_dict = self.compute_loss(embeddings, labels, indices_tuple, ref_emb, ref_labels)
        item = loss_dict["loss"]
        if item["reduction_type"] == "already_reduced" or not torch.is_tensor(item["losses"]):
            return torch.sum(embeddings * 0)  # BaseReducer.zero_loss
        losses = item["losses"]
        keep = losses > 0  # AvgNonZeroReducer = ThresholdReducer(low=0)
        return torch.mean(losses[keep]) if bool(keep.any()) else torch.sum(embeddings * 0)


def _pml_logsumexp(x: torch.Tensor, keep_mask=None, add_one: bool = True, dim: int = 1) -> torch.Tensor:
    """loss_and_miner_utils.logsumexp."""
    if keep_mask is not None:
        x = x.masked_fill(~keep_mask, torch.finfo(x.dtype).min)
    if add_one:
        zeros = torch.zeros(x.size(dim - 1), dtype=x.dtype, device=x.device).unsqueeze(dim)
        x = torch.cat([x, zeros], dim=dim)
    output = torch.logsumexp(x, dim=dim, keepdim=True)
    if keep_mask is not None:
        output = output.masked_fill(~torch.any(keep_mask, dim=dim, keepdim=True), 0)
    return output


def install(reference_root: str = "/root/reference") -> None:
    """Register the shims and put the reference on sys.path (idempotent)."""
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    if "lightning" in sys.modules and getattr(sys.modules["lightning"], "_mb200_stub", False):
        return
    _module("lightning", LightningModule=LightningModule, _mb200_stub=True)
    pyg = _module("torch_geometric")
    pyg.utils = _module("torch_geometric.utils", to_dense_batch=tp.to_dense_batch)

    tm = _module(
        "torchmetrics",
        Metric=tp.Metric,
        MetricCollection=tp.MetricCollection,
        MeanMetric=tp.MeanMetric,
        MinMetric=tp.MinMetric,
    )
    tm.classification = _module("torchmetrics.classification", AUROC=tp.AUROC)
    tm.retrieval = _module(
        "torchmetrics.retrieval", RetrievalMRR=tp.RetrievalMRR, RetrievalNormalizedDCG=tp.RetrievalNormalizedDCG
    )
    tm.retrieval.base = _module("torchmetrics.retrieval.base", RetrievalMetric=tp.RetrievalMetric)
    tm.utilities = _module("torchmetrics.utilities")
    tm.utilities.checks = _module(
        "torchmetrics.utilities.checks",
        _check_retrieval_inputs=tp._check_retrieval_inputs,
        _check_retrieval_functional_inputs=tp._check_retrieval_functional_inputs,
    )
    tm.utilities.data = _module(
        "torchmetrics.utilities.data", _flexible_bincount=tp._flexible_bincount, dim_zero_cat=tp.dim_zero_cat
    )

    pml = _module("pytorch_metric_learning")
    pml.losses = _module("pytorch_metric_learning.losses", SupConLoss=PmlSupConLossBase)
    pml.distances = _module("pytorch_metric_learning.distances", DotProductSimilarity=_Inert)
    pml.utils = _module("pytorch_metric_learning.utils")
    pml.utils.common_functions = _module(
        "pytorch_metric_learning.utils.common_functions",
        small_val=lambda dtype: torch.finfo(dtype).tiny,
        torch_arange_from_size=lambda t, size_dim=0: torch.arange(t.size(size_dim), device=t.device),
    )
    pml.utils.loss_and_miner_utils = _module("pytorch_metric_learning.utils.loss_and_miner_utils", logsumexp=_pml_logsumexp)
    _module("MulticoreTSNE", MulticoreTSNE=_Inert)
    _module("seaborn")
    _module("colorcet")
    mpl = _module("matplotlib")
    mpl.pyplot = _module("matplotlib.pyplot")
