"""The step BEFORE the hot path: fill the embedding tables once and turn the
reference's parsed behaviours into CSR, so the PLM leaves the evaluation loop.

The reference re-encodes every history and candidate row of every batch with the PLM
(cr_module.py:107,113; ensemble_module.py:115,121).  Here each module's ``news_encoder`` runs once
over the unique news of the split (the set ``MINDNewsDataset`` builds, mind_news_dataset.py:16-25)
and its output is written to a [n_news, D] table; behaviours become int32 CSR over table rows with
the reference's history truncation (first ``max_history_length`` clicks, mind_rec_dataset.py:92).

Everything here is host-side plumbing around torch modules and pandas frames; no arithmetic of the
scoring path is done on the CPU.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from .data import MAX_HISTORY, Behaviours


def news_row_map(news_ids: Iterable[str]) -> Dict[str, int]:
    """news id (the DataFrame index of the reference's ``news`` frame) -> table row."""
    return {nid: row for row, nid in enumerate(news_ids)}


def unique_news_ids(behaviors: Any, max_history_length: int = MAX_HISTORY) -> List[str]:
    """The news a split's behaviours refer to (history cut to ``max_history_length`` like the dataset does,
    mind_rec_dataset.py:92), in first-appearance order -- the set ``MINDNewsDataset`` builds (mind_news_dataset.py:16-25)."""
    seen: Dict[str, None] = {}
    for h, c in zip(behaviors["history"].tolist(), behaviors["candidates"].tolist()):
        h = parse_id_list(h) if isinstance(h, str) else list(h)
        c = parse_id_list(c) if isinstance(c, str) else list(c)
        for nid in h[:max_history_length]:
            seen.setdefault(nid, None)
        for nid in c:
            seen.setdefault(nid, None)
    return list(seen)


def parse_id_list(text: str) -> List[str]:
    """One ``history`` / ``candidates`` cell of ``parsed_behaviors.tsv`` -- a stringified Python list --
    exactly as the reference's converters read it (mind_dataframe.py:283-286)."""
    return text.strip("[]").replace("'", "").split(", ")


def parse_label_list(text: str) -> List[int]:
    return list(map(int, text.strip("[]").split(", ")))


def behaviours_to_csr(
    histories: Sequence[Sequence[str]],
    candidates: Sequence[Sequence[str]],
    labels: Sequence[Sequence[int]],
    nid2row: Dict[str, int],
    max_history_length: int = MAX_HISTORY,
) -> Behaviours:
    """Rows of the reference's behaviours frame (columns ``history``, ``candidates``, ``labels``;
    mind_dataframe.py:360) -> CSR.  Keeps the FIRST ``max_history_length`` clicks, like
    ``MINDRecDatasetTest.__getitem__`` (mind_rec_dataset.py:92).  Unknown news ids raise KeyError, as
    ``news.loc[...]`` does in the reference (mind_rec_dataset.py:96-97)."""
    if not (len(histories) == len(candidates) == len(labels)):
        raise ValueError("histories, candidates and labels must have one entry per impression")
    hist_rows: List[int] = []
    cand_rows: List[int] = []
    labs: List[int] = []
    hist_off, cand_off = [0], [0]
    for h, c, y in zip(histories, candidates, labels):
        if len(c) != len(y):
            raise ValueError("one label per candidate expected")
        h = list(h)[:max_history_length]
        if len(h) == 0:
            # the reference drops such users when it parses behaviors.tsv (mind_dataframe.py:311-315)
            raise ValueError("impression with an empty history (drop it, as the reference does at parse time)")
        hist_rows.extend(nid2row[n] for n in h)
        cand_rows.extend(nid2row[n] for n in c)
        labs.extend(int(v) for v in y)
        hist_off.append(len(hist_rows))
        cand_off.append(len(cand_rows))
    bhv = Behaviours(
        np.asarray(hist_off, dtype=np.int32), np.asarray(hist_rows, dtype=np.int32), np.asarray(cand_off, dtype=np.int32),
        np.asarray(cand_rows, dtype=np.int32), (np.asarray(labs) != 0).astype(np.uint8),
    )
    bhv.validate(len(nid2row))
    return bhv


def behaviours_frame_to_csr(behaviors: Any, nid2row: Dict[str, int], max_history_length: int = MAX_HISTORY) -> Behaviours:
    """Same, from the pandas frame the reference's ``MINDDataFrame._load_behaviors`` returns (or from the
    raw text columns of ``parsed_behaviors.tsv``)."""
    def col(name: str, parse: Callable[[str], list]) -> list:
        values = behaviors[name].tolist()
        return [parse(v) if isinstance(v, str) else list(v) for v in values]

    return behaviours_to_csr(col("history", parse_id_list), col("candidates", parse_id_list), col("labels", parse_label_list), nid2row, max_history_length)


def read_parsed_behaviors(path: str, nid2row: Dict[str, int], max_history_length: int = MAX_HISTORY) -> Behaviours:
    """``parsed_behaviors.tsv`` (the file the reference caches with ``to_tsv`` and reloads in
    ``MINDDataFrame._load_behaviors``, mind_dataframe.py:278-288,360-366: tab separated, columns ``user``, ``history``,
    ``candidates``, ``labels`` with stringified Python lists) -> CSR, without going through pandas objects per cell."""
    import csv

    histories: List[List[str]] = []
    candidates: List[List[str]] = []
    labels: List[List[int]] = []
    with open(path, newline="") as f:
        reader = csv.reader(f, delimiter="\t")
        header = next(reader)
        col = {name: header.index(name) for name in ("history", "candidates", "labels")}
        for row in reader:
            if not row:
                continue
            histories.append(parse_id_list(row[col["history"]]))
            candidates.append(parse_id_list(row[col["candidates"]]))
            labels.append(parse_label_list(row[col["labels"]]))
    return behaviours_to_csr(histories, candidates, labels, nid2row, max_history_length)


@torch.no_grad()
def build_embedding_table(
    news_encoder: torch.nn.Module,
    news_batches: Iterable[Any],
    n_news: int,
    dim: int,
    device: torch.device,
    dtype: torch.dtype = torch.float32,
) -> Tensor:
    """Runs ``news_encoder`` (eval mode, the module's own PyTorch forward: news_encoder.py:75-129) over
    ``news_batches`` -- an iterable of already tokenised news inputs in table-row order, e.g. the
    ``x`` dicts ``MINDCollate._tokenize_df`` produces (mind_rec_dataset.py:146-168) -- and writes the
    vectors into a [n_news, dim] table on ``device``."""
    was_training = news_encoder.training
    news_encoder.eval()
    table = torch.empty(n_news, dim, dtype=dtype, device=device)
    row = 0
    for x in news_batches:
        vec = news_encoder(_to_device(x, device))
        if vec.dim() != 2 or vec.shape[1] != dim:
            raise ValueError(f"news_encoder returned {tuple(vec.shape)}, expected [batch, {dim}]")
        table[row : row + vec.shape[0]] = vec.to(dtype)
        row += vec.shape[0]
    if row != n_news:
        raise ValueError(f"news_batches covered {row} news, table has {n_news} rows")
    news_encoder.train(was_training)
    return table


def _to_device(x: Any, device: torch.device) -> Any:
    if isinstance(x, Tensor):
        return x.to(device, non_blocking=True)
    if isinstance(x, dict):
        return {k: _to_device(v, device) for k, v in x.items()}
    if hasattr(x, "to") and not isinstance(x, (str, bytes)):
        try:
            return x.to(device)  # transformers.BatchEncoding
        except TypeError:
            return x
    return x


def aspect_arrays(news_frame: Any, nid2row: Dict[str, int]) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row ``category_label`` / ``sentiment_label`` (the columns MINDCollate reads,
    mind_rec_dataset.py:164-165) in table-row order."""
    order = sorted(nid2row, key=nid2row.get)
    sub = news_frame.loc[order]
    return sub["category_label"].to_numpy().astype(np.int32), sub["sentiment_label"].to_numpy().astype(np.int32)


# ---- persistence: tables + id map + CSR on disk, so cached mode does not re-encode per run ------------------
# The reference writes its parsed frames once and reloads them (mind_dataframe.py:278-288,360-366: `parsed_behaviors.tsv`; the
# news frame likewise); the embedding tables are the equivalent artefact of the accelerated path.  One directory per
# (split, checkpoint): `tables.npy` [M, n_news, D] (np.save: memory-mappable, fp32 or bf16-as-uint16), `news_ids.txt` (row order =
# the nid -> row map), `behaviours.npz` (CSR), optional `aspects.npz`, and `meta.json` with a fingerprint of whatever produced the
# table (checkpoint path + mtime / size, dtype, max history) so a stale cache is refused rather than silently used.

CACHE_FORMAT = 1


def fingerprint(*sources: Any) -> str:
    """A short stable digest of what the cached tables depend on: file paths are digested as (path, size, mtime), anything else by repr."""
    import hashlib
    import os

    h = hashlib.sha256()
    for s in sources:
        if isinstance(s, str) and os.path.exists(s):
            stt = os.stat(s)
            h.update(f"{os.path.abspath(s)}:{stt.st_size}:{int(stt.st_mtime)}".encode())
        else:
            h.update(repr(s).encode())
    return h.hexdigest()[:16]


def save_cache(directory: str, tables: Sequence[Tensor], news_ids: Sequence[str], bhv: Behaviours, key: str,
               aspects: Tuple[np.ndarray, np.ndarray] = None, max_history_length: int = MAX_HISTORY) -> None:
    """Writes the cache directory atomically (temporary names, then rename): a reader never sees a half-written cache."""
    import json
    import os

    os.makedirs(directory, exist_ok=True)
    if len(news_ids) != tables[0].shape[0] or any(t.shape != tables[0].shape or t.dtype != tables[0].dtype for t in tables):
        raise ValueError("tables must share shape / dtype and have one row per news id")
    stack = torch.stack([t.detach().cpu() for t in tables])
    is_bf16 = stack.dtype == torch.bfloat16
    arr = stack.view(torch.int16).numpy().view(np.uint16) if is_bf16 else stack.float().numpy()
    tmp = lambda name: os.path.join(directory, "." + name + ".tmp")
    final = lambda name: os.path.join(directory, name)
    with open(tmp("tables.npy"), "wb") as f:
        np.save(f, arr)
    with open(tmp("news_ids.txt"), "w") as f:
        f.write("\n".join(news_ids) + "\n")
    with open(tmp("behaviours.npz"), "wb") as f:
        np.savez(f, hist_offsets=bhv.hist_offsets, hist_ids=bhv.hist_ids, cand_offsets=bhv.cand_offsets, cand_ids=bhv.cand_ids, labels=bhv.labels)
    names = ["tables.npy", "news_ids.txt", "behaviours.npz"]
    if aspects is not None:
        with open(tmp("aspects.npz"), "wb") as f:
            np.savez(f, category=np.asarray(aspects[0], dtype=np.int32), sentiment=np.asarray(aspects[1], dtype=np.int32))
        names.append("aspects.npz")
    meta = {"format": CACHE_FORMAT, "key": key, "n_modules": len(tables), "n_news": int(tables[0].shape[0]), "dim": int(tables[0].shape[1]),
            "dtype": "bf16" if is_bf16 else "f32", "n_impressions": bhv.n_impressions, "max_history_length": int(max_history_length),
            "has_aspects": aspects is not None}
    for name in names:
        os.replace(tmp(name), final(name))
    with open(tmp("meta.json"), "w") as f:
        json.dump(meta, f)
    os.replace(tmp("meta.json"), final("meta.json"))  # last: its presence marks the cache complete


def load_cache(directory: str, key: str, device: torch.device = None, mmap: bool = True):
    """(tables [list of Tensor], news_ids, Behaviours, aspects or None) from ``save_cache``'s directory, or None when there is no
    complete cache or it was built from something else (``key`` differs) -- the caller then rebuilds.  ``mmap`` maps the table
    file instead of reading it through Python (a 161 k x 768 x 3 fp32 set is 1.5 GB); ``device`` uploads the tables."""
    import json
    import os

    meta_path = os.path.join(directory, "meta.json")
    if not os.path.exists(meta_path):
        return None
    with open(meta_path) as f:
        meta = json.load(f)
    if meta.get("format") != CACHE_FORMAT or meta.get("key") != key:
        return None
    arr = np.load(os.path.join(directory, "tables.npy"), mmap_mode="r" if mmap else None)
    if arr.shape != (meta["n_modules"], meta["n_news"], meta["dim"]):
        return None
    tables = []
    for m in range(arr.shape[0]):
        t = torch.from_numpy(np.ascontiguousarray(arr[m]).view(np.int16) if meta["dtype"] == "bf16" else np.ascontiguousarray(arr[m]))
        t = t.view(torch.bfloat16) if meta["dtype"] == "bf16" else t
        tables.append(t.to(device) if device is not None else t)
    with open(os.path.join(directory, "news_ids.txt")) as f:
        news_ids = f.read().split("\n")[:-1]
    z = np.load(os.path.join(directory, "behaviours.npz"))
    bhv = Behaviours(z["hist_offsets"], z["hist_ids"], z["cand_offsets"], z["cand_ids"], z["labels"])
    bhv.validate(meta["n_news"])
    aspects = None
    if meta.get("has_aspects"):
        a = np.load(os.path.join(directory, "aspects.npz"))
        aspects = (a["category"], a["sentiment"])
    if len(news_ids) != meta["n_news"] or bhv.n_impressions != meta["n_impressions"]:
        return None
    return tables, news_ids, bhv, aspects
