"""TEST INFRASTRUCTURE ONLY -- restatement of the THIRD-PARTY arithmetic on the MANNeR hot path.

The reference (andreeaiana/manner) delegates part of its evaluation arithmetic to two dependencies
whose sources are NOT under /root/reference and are not installable here (no network):

* torch_geometric  (requirements.txt:7 ``pyg>=2.3.1``)  -- ``to_dense_batch``; call sites
  manner/models/cr_module.py:108,114,142 and manner/models/ensemble_module.py:116,122,155-163.
* torchmetrics     (requirements.txt:5 ``>=0.11.4`` and environment.yaml:31 ``0.*`` => 0.11.4) --
  ``AUROC(task="binary")``, ``RetrievalMRR``, ``RetrievalNormalizedDCG``, ``MetricCollection``,
  ``RetrievalMetric.compute``, ``_check_retrieval_inputs``; call sites cr_module.py:79-89,273 and
  ensemble_module.py:50-84,230-238; the in-repo fork manner/metrics/base.py:92-129 corroborates the
  group-by ``compute`` loop.

This file restates the *published algorithms* of those pinned versions on plain torch CPU tensors.
Nothing pins them (the reference ships no tests): PARITY UNPINNED for the third-party part; it is
cross-checked against sklearn / scipy in tests/test_oracle.py.  The same classes double as import
stubs when tests/golden/make_golden.py executes the reference's own in-repo code (see
oracle/ref_stubs.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product path (manner_b200/) never does.
"""
from __future__ import annotations

import inspect
from copy import deepcopy
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import Tensor

# --------------------------------------------------------------------------------------------
# torch_geometric.utils.to_dense_batch (pyg 2.3)
# --------------------------------------------------------------------------------------------


def to_dense_batch(
    x: Tensor,
    batch: Optional[Tensor] = None,
    fill_value: float = 0.0,
    max_num_nodes: Optional[int] = None,
    batch_size: Optional[int] = None,
) -> Tuple[Tensor, Tensor]:
    """Left-aligned padded [B, Nmax, ...] tensor + bool mask from sorted segment ids."""
    n = x.size(0)
    if batch is None and max_num_nodes is None:
        return x.unsqueeze(0), torch.ones(1, n, dtype=torch.bool, device=x.device)
    if batch is None:
        batch = x.new_zeros(n, dtype=torch.long)
    if batch_size is None:
        batch_size = int(batch.max()) + 1
    counts = torch.zeros(batch_size, dtype=torch.long, device=x.device)
    counts.scatter_add_(0, batch, torch.ones_like(batch))
    starts = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    if max_num_nodes is None:
        max_num_nodes = int(counts.max())
    within = torch.arange(n, device=x.device) - starts[batch]
    flat = within + batch * max_num_nodes
    keep = within < max_num_nodes
    if not bool(keep.all()):
        x, flat = x[keep], flat[keep]
    dense = x.new_full([batch_size * max_num_nodes] + list(x.shape[1:]), fill_value)
    dense[flat] = x
    dense = dense.view([batch_size, max_num_nodes] + list(x.shape[1:]))
    mask = torch.zeros(batch_size * max_num_nodes, dtype=torch.bool, device=x.device)
    mask[flat] = True
    return dense, mask.view(batch_size, max_num_nodes)


# --------------------------------------------------------------------------------------------
# torchmetrics.utilities
# --------------------------------------------------------------------------------------------


def dim_zero_cat(x: Any) -> Tensor:
    if isinstance(x, Tensor):
        return x
    x = [y.unsqueeze(0) if y.numel() == 1 and y.ndim == 0 else y for y in x]
    if not x:
        raise ValueError("No samples to concatenate")
    return torch.cat(x, dim=0)


def _flexible_bincount(x: Tensor) -> Tensor:
    x = x - x.min()
    uniq = torch.unique(x)
    return torch.bincount(x, minlength=int(uniq.max()) + 1)[uniq]


def _retrieval_types(preds: Tensor, target: Tensor, allow_non_binary_target: bool = False):
    if target.dtype not in (torch.bool, torch.long, torch.int) and not target.is_floating_point():
        raise ValueError("`target` must be a tensor of booleans, integers or floats")
    if not preds.is_floating_point():
        raise ValueError("`preds` must be a tensor of floats")
    if not allow_non_binary_target and (target.max() > 1 or target.min() < 0):
        raise ValueError("`target` must contain `binary` values")
    target = target.float().flatten() if target.is_floating_point() else target.long().flatten()
    return preds.float().flatten(), target


def _check_retrieval_functional_inputs(preds: Tensor, target: Tensor, allow_non_binary_target: bool = False):
    if preds.shape != target.shape:
        raise ValueError("`preds` and `target` must be of the same shape")
    if not preds.numel() or not preds.size():
        raise ValueError("`preds` and `target` must be non-empty and non-scalar tensors")
    return _retrieval_types(preds, target, allow_non_binary_target)


def _check_retrieval_inputs(
    indexes: Tensor,
    preds: Tensor,
    target: Tensor,
    allow_non_binary_target: bool = False,
    ignore_index: Optional[int] = None,
):
    if indexes.shape != preds.shape or preds.shape != target.shape:
        raise ValueError("`indexes`, `preds` and `target` must be of the same shape")
    if ignore_index is not None:
        valid = target != ignore_index
        indexes, preds, target = indexes[valid], preds[valid], target[valid]
    if not indexes.numel() or not indexes.size():
        raise ValueError("`indexes`, `preds` and `target` must be non-empty and non-scalar tensors")
    if indexes.dtype is not torch.long:
        raise ValueError("`indexes` must be a tensor of long integers")
    preds, target = _retrieval_types(preds, target, allow_non_binary_target)
    return indexes.long().flatten(), preds, target


# --------------------------------------------------------------------------------------------
# torchmetrics.functional.retrieval (0.11.4)
# --------------------------------------------------------------------------------------------


def stable_desc_argsort(preds: Tensor) -> Tensor:
    """``argsort(preds, descending=True)`` with the canonical tie rule: among
    equal scores the lower position ranks first.  torch<=2.1 CPU argsort behaved this way; the
    torch 2.11 AVX-512 sort in this image does not, hence the explicit ``stable=True``."""
    return torch.argsort(preds, dim=-1, descending=True, stable=True)


def retrieval_reciprocal_rank(preds: Tensor, target: Tensor) -> Tensor:
    preds, target = _check_retrieval_functional_inputs(preds, target)
    if not target.sum():
        return torch.tensor(0.0, device=preds.device)
    ordered = target[stable_desc_argsort(preds)]
    first = torch.nonzero(ordered).view(-1)
    return 1.0 / (first[0] + 1.0)


def _dcg(target: Tensor) -> Tensor:
    denom = torch.log2(torch.arange(target.shape[-1], device=target.device) + 2.0)
    return (target / denom).sum(dim=-1)


def retrieval_normalized_dcg(preds: Tensor, target: Tensor, k: Optional[int] = None) -> Tensor:
    preds, target = _check_retrieval_functional_inputs(preds, target, allow_non_binary_target=True)
    k = preds.shape[-1] if k is None else k
    if not (isinstance(k, int) and k > 0):
        raise ValueError("`k` has to be a positive integer or None")
    got = target[stable_desc_argsort(preds)][:k]
    best = torch.sort(target, descending=True)[0][:k]
    ideal_dcg = _dcg(best)
    target_dcg = _dcg(got)
    irrelevant = ideal_dcg == 0
    target_dcg[irrelevant] = 0
    target_dcg[~irrelevant] /= ideal_dcg[~irrelevant]
    return target_dcg.mean()


# --------------------------------------------------------------------------------------------
# torchmetrics.functional.classification  BinaryAUROC pieces (0.11.4, thresholds=None)
# --------------------------------------------------------------------------------------------


def _binary_format(preds: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    preds, target = preds.flatten(), target.flatten()
    if not torch.all((preds >= 0) * (preds <= 1)):
        preds = preds.sigmoid()
    return preds, target


def _binary_clf_curve(preds: Tensor, target: Tensor, pos_label: int = 1):
    order = torch.argsort(preds, descending=True)
    preds, target = preds[order], target[order]
    distinct = torch.where(preds[1:] - preds[:-1])[0]
    cut = torch.nn.functional.pad(distinct, [0, 1], value=target.size(0) - 1)
    target = (target == pos_label).to(torch.long)
    tps = torch.cumsum(target, dim=0)[cut]
    fps = 1 + cut - tps
    return fps, tps, preds[cut]


def binary_auroc(preds: Tensor, target: Tensor) -> Tensor:
    """Pooled AUROC exactly as ``AUROC(task="binary")`` evaluates it."""
    preds, target = _binary_format(preds.float(), target.long())
    fps, tps, thr = _binary_clf_curve(preds, target)
    tps = torch.cat([torch.zeros(1, dtype=tps.dtype), tps])
    fps = torch.cat([torch.zeros(1, dtype=fps.dtype), fps])
    fpr = torch.zeros(fps.shape, dtype=torch.float32) if fps[-1] <= 0 else fps / fps[-1]
    tpr = torch.zeros(tps.shape, dtype=torch.float32) if tps[-1] <= 0 else tps / tps[-1]
    return torch.trapz(tpr, fpr)


# --------------------------------------------------------------------------------------------
# torchmetrics Metric / MetricCollection protocol (just enough to run the reference's classes)
# --------------------------------------------------------------------------------------------


class Metric(torch.nn.Module):
    is_differentiable: Optional[bool] = None
    higher_is_better: Optional[bool] = None
    full_state_update: Optional[bool] = None

    def __init__(self, **kwargs: Any) -> None:
        super().__init__()
        self._defaults: Dict[str, Any] = {}

    def add_state(self, name: str, default: Any, dist_reduce_fx: Any = None, persistent: bool = False) -> None:
        self._defaults[name] = deepcopy(default)
        setattr(self, name, deepcopy(default))

    def reset(self) -> None:
        for name, default in self._defaults.items():
            setattr(self, name, deepcopy(default))

    def update(self, *args: Any, **kwargs: Any) -> None:  # pragma: no cover - abstract
        raise NotImplementedError

    def compute(self) -> Any:  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        # 0.11.4 `_forward_reduce_state_update`: batch value from a fresh state, then merge.
        saved = {n: getattr(self, n) for n in self._defaults}
        self.reset()
        self.update(*args, **kwargs)
        batch_val = self.compute()
        for n, old in saved.items():
            new = getattr(self, n)
            setattr(self, n, (old + new) if isinstance(old, list) else old + new)
        return batch_val

    def clone(self) -> "Metric":
        return deepcopy(self)

    def _filter_kwargs(self, **kwargs: Any) -> Dict[str, Any]:
        params = inspect.signature(self.update).parameters
        if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in params.values()):
            return kwargs
        named = {
            k for k, p in params.items() if p.kind not in (inspect.Parameter.VAR_POSITIONAL, inspect.Parameter.VAR_KEYWORD)
        }
        return {k: v for k, v in kwargs.items() if k in named}


class MetricCollection(torch.nn.ModuleDict):
    def __init__(self, metrics: Dict[str, Metric], prefix: Optional[str] = None, postfix: Optional[str] = None):
        super().__init__()
        self.prefix, self.postfix = prefix, postfix
        for name, metric in metrics.items():
            self[name] = metric

    def _name(self, base: str) -> str:
        return f"{self.prefix or ''}{base}{self.postfix or ''}"

    def forward(self, *args: Any, **kwargs: Any) -> Dict[str, Any]:
        return {self._name(k): m(*args, **m._filter_kwargs(**kwargs)) for k, m in self.items()}

    def update(self, *args: Any, **kwargs: Any) -> None:
        for m in self.values():
            m.update(*args, **m._filter_kwargs(**kwargs))

    def compute(self) -> Dict[str, Any]:
        return {self._name(k): m.compute() for k, m in self.items()}

    def reset(self) -> None:
        for m in self.values():
            m.reset()

    def clone(self, prefix: Optional[str] = None, postfix: Optional[str] = None) -> "MetricCollection":
        mc = deepcopy(self)
        if prefix is not None:
            mc.prefix = prefix
        if postfix is not None:
            mc.postfix = postfix
        return mc


class BinaryAUROC(Metric):
    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state("preds", [])
        self.add_state("target", [])

    def update(self, preds: Tensor, target: Tensor) -> None:
        preds, target = _binary_format(preds.float(), target)
        self.preds.append(preds)
        self.target.append(target)

    def compute(self) -> Tensor:
        preds, target = dim_zero_cat(self.preds), dim_zero_cat(self.target)
        fps, tps, _ = _binary_clf_curve(preds, target)
        tps = torch.cat([torch.zeros(1, dtype=tps.dtype), tps])
        fps = torch.cat([torch.zeros(1, dtype=fps.dtype), fps])
        fpr = torch.zeros(fps.shape, dtype=torch.float32) if fps[-1] <= 0 else fps / fps[-1]
        tpr = torch.zeros(tps.shape, dtype=torch.float32) if tps[-1] <= 0 else tps / tps[-1]
        return torch.trapz(tpr, fpr)


def AUROC(task: str = "binary", num_classes: Optional[int] = None, **kwargs: Any) -> Metric:
    if task != "binary":
        raise NotImplementedError("only the reference's task='binary' call (cr_module.py:81) is restated")
    return BinaryAUROC(**kwargs)


class RetrievalMetric(Metric):
    def __init__(self, empty_target_action: str = "neg", ignore_index: Optional[int] = None, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.allow_non_binary_target = False
        if empty_target_action not in ("error", "skip", "neg", "pos"):
            raise ValueError(f"Argument `empty_target_action` received a wrong value `{empty_target_action}`.")
        self.empty_target_action = empty_target_action
        self.ignore_index = ignore_index
        self.add_state("indexes", [])
        self.add_state("preds", [])
        self.add_state("target", [])

    def update(self, preds: Tensor, target: Tensor, indexes: Tensor) -> None:
        if indexes is None:
            raise ValueError("Argument `indexes` cannot be None")
        indexes, preds, target = _check_retrieval_inputs(
            indexes, preds, target, allow_non_binary_target=self.allow_non_binary_target, ignore_index=self.ignore_index
        )
        self.indexes.append(indexes)
        self.preds.append(preds)
        self.target.append(target)

    def compute(self) -> Tensor:
        indexes, preds, target = dim_zero_cat(self.indexes), dim_zero_cat(self.preds), dim_zero_cat(self.target)
        # canonical rule: keep the within-impression order (stable), as torch<=2.1 CPU sort did
        indexes, order = torch.sort(indexes, stable=True)
        preds, target = preds[order], target[order]
        sizes = _flexible_bincount(indexes).detach().cpu().tolist()
        res: List[Tensor] = []
        for p, t in zip(torch.split(preds, sizes, dim=0), torch.split(target, sizes, dim=0)):
            if not t.sum():
                if self.empty_target_action == "error":
                    raise ValueError("`compute` method was provided with a query with no positive target.")
                if self.empty_target_action == "pos":
                    res.append(torch.tensor(1.0))
                elif self.empty_target_action == "neg":
                    res.append(torch.tensor(0.0))
            else:
                res.append(self._metric(p, t))
        return torch.stack([x.to(preds) for x in res]).mean() if res else torch.tensor(0.0).to(preds)

    def _metric(self, preds: Tensor, target: Tensor) -> Tensor:  # pragma: no cover - abstract
        raise NotImplementedError


class RetrievalMRR(RetrievalMetric):
    def _metric(self, preds: Tensor, target: Tensor) -> Tensor:
        return retrieval_reciprocal_rank(preds, target)


class RetrievalNormalizedDCG(RetrievalMetric):
    def __init__(self, empty_target_action: str = "neg", ignore_index: Optional[int] = None, k: Optional[int] = None, **kw: Any):
        super().__init__(empty_target_action=empty_target_action, ignore_index=ignore_index, **kw)
        if k is not None and not (isinstance(k, int) and k > 0):
            raise ValueError("`k` has to be a positive integer or None")
        self.k = k
        self.allow_non_binary_target = True

    def _metric(self, preds: Tensor, target: Tensor) -> Tensor:
        return retrieval_normalized_dcg(preds, target, k=self.k)


class MeanMetric(Metric):
    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state("total", torch.tensor(0.0))
        self.add_state("weight", torch.tensor(0.0))

    def update(self, value: Any, weight: float = 1.0) -> None:
        value = torch.as_tensor(value, dtype=torch.float32)
        self.total = self.total + (value * weight).sum()
        self.weight = self.weight + weight * value.numel()

    def compute(self) -> Tensor:
        return self.total / self.weight


class MinMetric(Metric):
    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state("value", torch.tensor(float("inf")))

    def update(self, value: Any) -> None:
        self.value = torch.minimum(self.value, torch.as_tensor(value, dtype=torch.float32).min())

    def compute(self) -> Tensor:
        return self.value
