"""Behaviours in CSR form + the seeded synthetic MIND-shaped generator.

The reference hands the model a ragged batch as *sorted segment-id vectors*
(``batch_hist`` / ``batch_cand``, manner/data/components/mind_rec_dataset.py:114-132,171-174) next to
per-row inputs and ``labels``.  The B200 path consumes the equivalent CSR arrays directly -- int32
offsets + int32 table-row ids + uint8 labels -- so nothing is ever padded (``to_dense_batch`` of
cr_module.py:108-114 disappears).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

MAX_HISTORY = 50  # configs/data/mind_rec.yaml:43 -- the reference keeps the FIRST 50 clicks (mind_rec_dataset.py:92)

SHAPES = {
    # name: (n_news, n_impressions, behaviour seed)
    "tiny": (512, 64, 7),
    "mini": (4096, 1000, 11),
    "small": (65_238, 73_152, 42),  # MIND-small dev shape
    "large": (161_013, 2_370_727, 43),  # MIND-large test shape
}
TABLE_SEEDS = (1234, 1235, 1236)  # CR, category A-Module, sentiment A-Module
NUM_CATEG_CLASSES = 19  # configs/model/ensemble_module.yaml:8 (18 + 1)
NUM_SENT_CLASSES = 4  # configs/model/ensemble_module.yaml:9 (3 + 1)


@dataclass
class Behaviours:
    """CSR impressions on the host (numpy).  ``hist_ids`` / ``cand_ids`` index embedding-table rows."""

    hist_offsets: np.ndarray  # int32 [B+1]
    hist_ids: np.ndarray  # int32 [sum H]
    cand_offsets: np.ndarray  # int32 [B+1]
    cand_ids: np.ndarray  # int32 [sum C]
    labels: np.ndarray  # uint8 [sum C]

    @property
    def n_impressions(self) -> int:
        return int(self.hist_offsets.shape[0] - 1)

    @property
    def n_hist(self) -> int:
        return int(self.hist_ids.shape[0])

    @property
    def n_cand(self) -> int:
        return int(self.cand_ids.shape[0])

    @property
    def max_cand(self) -> int:
        return int(np.diff(self.cand_offsets).max()) if self.n_impressions else 0

    @property
    def max_hist(self) -> int:
        return int(np.diff(self.hist_offsets).max()) if self.n_impressions else 0

    def validate(self, n_news: Optional[int] = None) -> None:
        for name, off, ids in (("hist", self.hist_offsets, self.hist_ids), ("cand", self.cand_offsets, self.cand_ids)):
            if off.dtype != np.int32 or ids.dtype != np.int32:
                raise TypeError(f"{name}: offsets and ids must be int32")
            if off.ndim != 1 or off.shape[0] < 1 or off[0] != 0 or off[-1] != ids.shape[0]:
                raise ValueError(f"{name}: offsets must start at 0 and end at len(ids)")
            if np.any(np.diff(off) < 0):
                raise ValueError(f"{name}: offsets must be non-decreasing")
            if n_news is not None and ids.size and (ids.min() < 0 or ids.max() >= n_news):
                raise ValueError(f"{name}: ids out of range [0, {n_news})")
        if self.hist_offsets.shape != self.cand_offsets.shape:
            raise ValueError("hist and cand offsets must describe the same impressions")
        if self.labels.dtype != np.uint8 or self.labels.shape != self.cand_ids.shape:
            raise TypeError("labels must be uint8 with one entry per candidate")
        if self.labels.size and self.labels.max() > 1:
            raise ValueError("labels must be binary")  # torchmetrics RetrievalMRR rejects anything else
        if self.n_impressions and np.diff(self.hist_offsets).min() < 1:
            # users with an empty history are dropped at parse time (mind_dataframe.py:311-315)
            raise ValueError("every impression needs at least one history row")
        if self.n_impressions and np.diff(self.cand_offsets).min() < 1:
            raise ValueError("every impression needs at least one candidate")

    def slice(self, lo: int, hi: int) -> "Behaviours":
        """Impressions [lo, hi) with offsets rebased to 0 (what a rank's shard looks like)."""
        h0, h1 = int(self.hist_offsets[lo]), int(self.hist_offsets[hi])
        c0, c1 = int(self.cand_offsets[lo]), int(self.cand_offsets[hi])
        return Behaviours(
            (self.hist_offsets[lo : hi + 1] - h0).astype(np.int32),
            self.hist_ids[h0:h1],
            (self.cand_offsets[lo : hi + 1] - c0).astype(np.int32),
            self.cand_ids[c0:c1],
            self.labels[c0:c1],
        )

    def algorithmic_bytes(self, n_modules: int, dim: int, elem_bytes: int = 4, scores_written: bool = True) -> int:
        """row reads + ids + labels + offsets (+ scores)."""
        rows = self.n_hist + self.n_cand
        b = n_modules * elem_bytes * dim * rows + 4 * rows + self.n_cand + 2 * 4 * (self.n_impressions + 1)
        return b + (4 * self.n_cand if scores_written else 0)


def step_pads(offsets: np.ndarray, step: int) -> np.ndarray:
    """int32 [B]: how many zero rows / columns the reference's dense batch appends to each impression when the
    impressions are processed in steps of ``step`` (configs/data/mind_rec.yaml:51; ``to_dense_batch`` pads every
    segment of a step to the step's longest, cr_module.py:108-114,142)."""
    sizes = np.diff(np.asarray(offsets, dtype=np.int64))
    out = np.empty(sizes.shape[0], dtype=np.int32)
    for lo in range(0, sizes.shape[0], step):
        seg = sizes[lo : lo + step]
        out[lo : lo + step] = seg.max() - seg
    return out


def from_segment_ids(batch_hist: torch.Tensor, hist_rows: torch.Tensor, batch_cand: torch.Tensor, cand_rows: torch.Tensor, labels: torch.Tensor) -> Behaviours:
    """Convert the reference's MINDRecBatch segment-id contract (mind_batch.py:6-12;
    ``batch = repeat_interleave(arange(B), sizes)``, sorted) to CSR.  ``B = batch.max() + 1`` exactly
    as ``to_dense_batch`` derives it."""
    nb = int(max(int(batch_hist.max()), int(batch_cand.max()))) + 1
    for seg in (batch_hist, batch_cand):
        if seg.numel() > 1 and bool((seg[1:] < seg[:-1]).any()):
            raise ValueError("segment ids must be sorted (MINDCollate emits them sorted)")

    def offs(seg: torch.Tensor) -> np.ndarray:
        counts = torch.bincount(seg.long(), minlength=nb)
        return np.concatenate([[0], np.cumsum(counts.numpy())]).astype(np.int32)

    return Behaviours(
        offs(batch_hist),
        hist_rows.numpy().astype(np.int32),
        offs(batch_cand),
        cand_rows.numpy().astype(np.int32),
        (labels.numpy() != 0).astype(np.uint8),
    )


def _zipf_cdf(n: int, alpha: float) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), alpha)
    c = np.cumsum(w)
    return c / c[-1]


def _dedupe_within_segments(seg: np.ndarray, ids: np.ndarray, draw, rounds: int = 64) -> np.ndarray:
    """Re-draw duplicate ids inside each segment until none is left (no accidental score ties)."""
    ids = ids.copy()
    live = np.arange(ids.shape[0])  # positions whose segment may still hold a duplicate
    for it in range(rounds):
        key = seg[live].astype(np.int64) * (int(ids.max()) + 1) + ids[live]
        order = np.argsort(key, kind="stable")
        k = key[order]
        dup_sorted = np.zeros(live.shape[0], dtype=bool)
        dup_sorted[1:] = k[1:] == k[:-1]
        dup = live[order[dup_sorted]]
        if dup.size == 0:
            return ids
        ids[dup] = draw(dup.size, uniform=it >= rounds // 2)
        live = live[np.isin(seg[live], np.unique(seg[dup]))]
    # a few segments nearly as long as the pool they draw from: finish them exactly
    pool = np.unique(draw(max(4096, 64 * ids.shape[0] // max(1, np.unique(seg).size)), uniform=True))
    for sgm in np.unique(seg[live]):
        where = np.nonzero(seg == sgm)[0]
        _, first = np.unique(ids[where], return_index=True)
        dup_pos = np.setdiff1d(np.arange(where.size), first)
        free = np.setdiff1d(pool, ids[where])
        if free.size < dup_pos.size:
            raise RuntimeError("could not draw duplicate-free candidate lists")
        ids[where[dup_pos]] = free[: dup_pos.size]
    return ids


def synth_behaviours(
    n_news: int,
    n_impressions: int,
    seed: int,
    cand_window: int = 6000,
    zipf_alpha: float = 1.05,
    uniform_ids: bool = False,
) -> Behaviours:
    """MIND-shaped ragged impressions: H ~ clip(round(LogNormal(2.9, 1.0)), 1, 50),
    C ~ clip(round(LogNormal(3.2, 0.9)), 2, 300), positives clip(1 + Poisson(0.5), 1, C-1) at random
    positions, history ids Zipf(1.05) over the catalogue, candidates Zipf(1.05) over a recent window,
    no duplicate candidate inside an impression.  ``uniform_ids`` draws every id uniformly over the
    whole catalogue instead (no L2-friendly head: the HBM stress variant)."""
    rng = np.random.default_rng(seed)
    b = n_impressions
    h = np.clip(np.rint(rng.lognormal(2.9, 1.0, b)), 1, MAX_HISTORY).astype(np.int64)
    c = np.clip(np.rint(rng.lognormal(3.2, 0.9, b)), 2, 300).astype(np.int64)
    window = min(cand_window, n_news)
    c = np.minimum(c, window)
    p = np.clip(1 + rng.poisson(0.5, b), 1, c - 1).astype(np.int64)
    hist_offsets = np.concatenate([[0], np.cumsum(h)])
    cand_offsets = np.concatenate([[0], np.cumsum(c)])
    nh, ncand = int(hist_offsets[-1]), int(cand_offsets[-1])
    if max(nh, ncand) >= 2**31:
        raise ValueError("CSR offsets are int32; shard the behaviours")

    perm = rng.permutation(n_news)  # popularity rank -> table row, so hot rows are scattered in memory
    if uniform_ids:
        hist_ids = rng.integers(0, n_news, nh)
    else:
        hist_ids = perm[np.searchsorted(_zipf_cdf(n_news, zipf_alpha), rng.random(nh))]
    wcdf = _zipf_cdf(window, zipf_alpha)
    recent = perm[rng.permutation(n_news)[:window]] if not uniform_ids else None

    def draw(n: int, uniform: bool = False) -> np.ndarray:
        if uniform_ids:
            return rng.integers(0, n_news, n)
        if uniform:
            return recent[rng.integers(0, window, n)]
        return recent[np.searchsorted(wcdf, rng.random(n))]

    seg = np.repeat(np.arange(b), c)
    cand_ids = _dedupe_within_segments(seg, draw(ncand), draw)

    # positives: the p smallest of C random keys per impression
    keys = rng.random(ncand)
    order = np.lexsort((keys, seg))
    rank_in_seg = np.empty(ncand, dtype=np.int64)
    rank_in_seg[order] = np.arange(ncand) - cand_offsets[:-1][seg[order]]
    labels = (rank_in_seg < p[seg]).astype(np.uint8)

    return Behaviours(
        hist_offsets.astype(np.int32),
        hist_ids.astype(np.int32),
        cand_offsets.astype(np.int32),
        cand_ids.astype(np.int32),
        labels,
    )


def synth_table(n_news: int, dim: int, seed: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """E ~ N(0, 1) * 2/sqrt(D): scores O(1), sigmoid unsaturated."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(n_news, dim, generator=g, dtype=torch.float32) * (2.0 / dim**0.5)
    return t.to(dtype)


def synth_aspects(n_news: int, seed: int = 99) -> Dict[str, np.ndarray]:
    """category ~ U{1..18}, sentiment ~ U{1..3} per news (label 0 is the reference's padding class)."""
    rng = np.random.default_rng(seed)
    return {
        "category": rng.integers(1, NUM_CATEG_CLASSES, n_news).astype(np.int32),
        "sentiment": rng.integers(1, NUM_SENT_CLASSES, n_news).astype(np.int32),
    }


def synth_workload(shape: str, n_modules: int = 1, dim: int = 768, dtype: torch.dtype = torch.float32, seed_offset: int = 0, uniform_ids: bool = False) -> Tuple[list, Behaviours]:
    n_news, n_impr, seed = SHAPES[shape]
    tables = [synth_table(n_news, dim, TABLE_SEEDS[m], dtype) for m in range(n_modules)]
    return tables, synth_behaviours(n_news, n_impr, seed + seed_offset, uniform_ids=uniform_ids)


def balanced_shard_bounds(bhv: Behaviours, world_size: int, align: int = 1) -> np.ndarray:
    """Contiguous impression ranges per rank balanced by rows gathered, sum(H_i + C_i), not by count
    (SURVEY 8(e)).  Returns int64 [world_size + 1] impression boundaries.  ``align`` > 1 rounds the inner boundaries
    to multiples of the reference's step size: early fusion and the losses depend on which impressions share a step
    (pads, MeanMetric over steps), so a shard must start on a step boundary to reproduce the single-GPU numbers."""
    work = bhv.hist_offsets.astype(np.int64) + bhv.cand_offsets.astype(np.int64)
    targets = work[-1] * np.arange(1, world_size, dtype=np.float64) / world_size
    cuts = np.searchsorted(work, targets, side="left")
    if align > 1:
        cuts = np.minimum((cuts + align // 2) // align * align, bhv.n_impressions)
        cuts = np.maximum.accumulate(cuts)
    return np.concatenate([[0], cuts, [bhv.n_impressions]]).astype(np.int64)
