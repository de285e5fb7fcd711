"""CPU, world_size 2 over gloo: the N>1 host logic of manner_b200/dist.py -- shard by rows, one
all-reduce of the additive metric payload, and the pooled-AUC exchange (all-gather of positive keys,
per-rank counting against the local sorted negatives, all-reduce of three integers).

The compute callables passed in here are numpy stand-ins for the CUDA stages (same contracts as
mb200_auc_build_keys / sort_keys / rank_sum); what is under test is the protocol around them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manner_b200 import data as mdata
from manner_b200 import dist as mdist
from oracle import manner_oracle as mo

WORLD = 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _orderable(x: np.ndarray) -> np.ndarray:
    x = np.where(x == 0, np.float32(0), x).astype(np.float32)
    b = x.view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)


def _np_build_and_sort(preds, labels, flags):
    """numpy stand-in for stages 1+2 (keys as int32 bit patterns, like the CUDA path hands torch)."""
    p = preds.numpy().astype(np.float32)
    if int(flags.item()) & 4:
        p = torch.from_numpy(p).sigmoid().numpy()
    keys = _orderable(p)
    pos = labels.numpy() != 0
    neg_keys = np.where(pos, np.uint32(0xFFFFFFFF), keys)
    pos_keys = np.zeros(p.size, dtype=np.uint32)
    pos_keys[: pos.sum()] = keys[pos]
    return (torch.from_numpy(np.sort(neg_keys).view(np.int32).copy()), torch.from_numpy(pos_keys.view(np.int32).copy()),
            torch.tensor([int(pos.sum())], dtype=torch.int64))


def _np_rank_sum(sorted_keys, n_pos_local, pos_keys, n_pos, sum2):
    s = sorted_keys.numpy().view(np.uint32)[: sorted_keys.numel() - int(n_pos_local.item())]
    k = pos_keys.numpy().view(np.uint32)[: int(n_pos.item())]
    sum2 += int(np.searchsorted(s, k, "left").sum() + np.searchsorted(s, k, "right").sum())


def _worker(rank: int, port: int, tmp: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        n_news = 300
        table = mdata.synth_table(n_news, 64, 5)
        bhv = mdata.synth_behaviours(n_news, 120, seed=9, cand_window=200)
        shard = mdist.shard_for_rank(bhv, rank, WORLD)
        ob = mo.Behaviours(shard.hist_offsets, shard.hist_ids, shard.cand_offsets, shard.cand_ids, shard.labels)
        local = mo.cr_eval_epoch(table, ob)  # the checker plays the part of the per-rank CUDA pass
        per = mo.per_impression_metrics(local["scores"], shard.labels, shard.cand_offsets)
        sums = torch.zeros(1, 13, dtype=torch.float64)
        sums[0, :5] = torch.from_numpy(per.astype(np.float64).sum(0))
        outside = int(not ((local["scores"] >= 0) & (local["scores"] <= 1)).all())
        flags = torch.tensor([4 * outside + (8 if rank == 1 else 0)], dtype=torch.int32)  # rank 1 also raises a made-up bit
        g_sums, g_flags, g_n = mdist.reduce_metric_sums(sums, flags, shard.n_impressions)
        preds_t, labels_t = torch.from_numpy(local["scores"]), torch.from_numpy(shard.labels)
        # (a) counts exchanged inside the call; (b) with a bound agreed beforehand (no count exchange)
        stats_a = mdist.pooled_auc_distributed(preds_t, labels_t, g_flags, build_and_sort=_np_build_and_sort, rank_sum=_np_rank_sum)
        cap = mdist.agree_pos_cap(int(shard.labels.sum()), torch.device("cpu"))
        stats_b = mdist.pooled_auc_distributed(preds_t, labels_t, g_flags, pos_cap=cap, build_and_sort=_np_build_and_sort, rank_sum=_np_rank_sum)
        assert torch.equal(stats_a, stats_b)
        auc, p, n = mdist.auc_from_stats(stats_b)
        if rank == 0:
            np.savez(os.path.join(tmp, "out.npz"), sums=g_sums.numpy(), flags=g_flags.numpy(), n=g_n, stats=np.array([auc, p, n]), cap=cap)
    finally:
        dist.destroy_process_group()


def test_two_rank_reduction_and_pooled_auc(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    z = np.load(tmp_path / "out.npz")
    n_news = 300
    table = mdata.synth_table(n_news, 64, 5)
    bhv = mdata.synth_behaviours(n_news, 120, seed=9, cand_window=200)
    whole = mo.cr_eval_epoch(table, mo.Behaviours(bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, bhv.cand_ids, bhv.labels))
    per = mo.per_impression_metrics(whole["scores"], bhv.labels, bhv.cand_offsets)
    assert int(z["n"]) == bhv.n_impressions
    np.testing.assert_allclose(z["sums"][0, :5], per.astype(np.float64).sum(0), rtol=1e-12)
    assert int(z["flags"][0]) == 4 + 8  # OR of the ranks' flag words survives the sum-reduction
    auc, p, n = z["stats"]
    bounds = mdata.balanced_shard_bounds(bhv, WORLD)
    assert int(z["cap"]) == max(int(bhv.slice(int(bounds[r]), int(bounds[r + 1])).labels.sum()) for r in range(WORLD))
    assert p == bhv.labels.sum() and n == bhv.labels.size - bhv.labels.sum()
    assert abs(auc - mo.pooled_auc_exact(whole["scores"], bhv.labels)) < 1e-12
    assert abs(auc - whole["metrics"]["test/auc"]) < 1e-6


def test_payload_round_trip():
    sums = torch.arange(26, dtype=torch.float64).reshape(2, 13)
    payload = mdist.pack_metric_payload(sums, torch.tensor([5], dtype=torch.int32), 1234)
    assert payload.numel() == 26 + 1 + mdist.N_FLAG_BITS
    s, f, n = mdist.unpack_metric_payload(payload * 1.0, sums.shape)
    assert torch.equal(s, sums) and int(f.item()) == 5 and n == 1234
    s2, f2, n2 = mdist.unpack_metric_payload_device(payload + payload, sums.shape)  # two identical ranks
    assert torch.equal(s2, 2 * sums) and int(f2.item()) == 5 and float(n2) == 2468


# ---- retrieval mode: the top-k exchange (SURVEY 8(e) collective 3) ----------------------------------------


def _np_merge(scores, ids):
    """numpy stand-in for mb200_merge_topk: [R, U, k] sorted lists -> global top-k (score desc, id asc)."""
    r, u, k = scores.shape
    s = scores.numpy().transpose(1, 0, 2).reshape(u, r * k).astype(np.float64)
    i = ids.numpy().transpose(1, 0, 2).reshape(u, r * k)
    s = np.where(i < 0, -np.inf, s)
    key_i = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((key_i, -s), axis=1)[:, :k]
    out_s, out_i = np.take_along_axis(s, order, 1).astype(np.float32), np.take_along_axis(i, order, 1)
    out_i = np.where(np.isneginf(out_s), -1, out_i)
    return torch.from_numpy(out_s), torch.from_numpy(out_i)


def _retrieval_worker(rank: int, port: int, tmp: str) -> None:
    from manner_b200 import retrieval as rt

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        g = torch.Generator().manual_seed(3)
        users = torch.randn(37, 64, generator=g).to(torch.bfloat16)
        catalog = torch.randn(700, 64, generator=g).to(torch.bfloat16)
        catalog[350:] = catalog[:350]  # cross-shard ties: the lower global id must win after the merge
        k = 20
        lo, hi = rt.catalog_shard_bounds(700, WORLD)[rank]
        s, i = mo.topk_select(mo.retrieval_scores(users, catalog[lo:hi]).numpy(), k, lo)  # the checker plays the per-rank kernel
        s, i = torch.from_numpy(s), torch.from_numpy(i)
        ag = rt.exchange_topk(s, i, None, "all_gather", merge=_np_merge)
        a2a = rt.exchange_topk(s, i, None, "all_to_all", merge=_np_merge)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), ag_s=ag[0].numpy(), ag_i=ag[1].numpy(), a2a_s=a2a[0].numpy(), a2a_i=a2a[1].numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_topk_exchange(tmp_path):
    from manner_b200 import retrieval as rt

    port = _free_port()
    mp.spawn(_retrieval_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    g = torch.Generator().manual_seed(3)
    users = torch.randn(37, 64, generator=g).to(torch.bfloat16)
    catalog = torch.randn(700, 64, generator=g).to(torch.bfloat16)
    catalog[350:] = catalog[:350]
    # per-shard scores are computed on row slices: restate the global matrix the same way so ties are exact
    bounds = rt.catalog_shard_bounds(700, WORLD)
    full = np.concatenate([mo.retrieval_scores(users, catalog[lo:hi]).numpy() for lo, hi in bounds], axis=1)
    want_s, want_i = mo.topk_select(full, 20)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(WORLD)]
    for z in outs:  # all_gather: every rank holds the full merged result
        np.testing.assert_array_equal(z["ag_i"], want_i)
        np.testing.assert_array_equal(z["ag_s"], want_s)
    # all_to_all: rank r holds its slice; concatenated in rank order they are the full result
    for r, z in enumerate(outs):
        idx = rt.CatalogRetriever.user_slice(37, 37, r, WORLD)
        np.testing.assert_array_equal(z["a2a_i"], want_i[idx])
        np.testing.assert_array_equal(z["a2a_s"], want_s[idx])
    assert sorted(sum((rt.CatalogRetriever.user_slice(37, 16, r, WORLD) for r in range(WORLD)), [])) == list(range(37))


def test_payload_flag_bits_round_trip():
    """The flag word travels in a packed payload as one 0 / 1 double per bit, in the order of MB200_PAYLOAD_TAIL (count, then bits
    1, 2, 4, 8 and 64 = upload time-out); host packing, host unpacking and the decode of a reduced tail must agree."""
    from manner_b200 import _native as nat

    assert mdist.FLAG_BITS == nat.PAYLOAD_FLAG_BITS and nat.PAYLOAD_TAIL == 1 + len(mdist.FLAG_BITS)
    for flags in (0, 1, 2 | 8, 64, 1 | 4 | 64, 1 | 2 | 4 | 8 | 64):
        payload = mdist.pack_metric_payload(torch.zeros(2, nat.NUM_METRICS, dtype=torch.float64), torch.tensor([flags], dtype=torch.int32), 7)
        assert payload.numel() == 2 * nat.NUM_METRICS + nat.PAYLOAD_TAIL
        _, f, n = mdist.unpack_metric_payload(payload, torch.Size([2, nat.NUM_METRICS]))
        assert int(f.item()) == flags and n == 7
        tail = payload[2 * nat.NUM_METRICS + 1 :].tolist()
        assert mdist.flags_from_payload_tail(tail) == flags
        # after a sum over ranks an entry is "how many ranks had the bit": still decodes to the OR
        assert mdist.flags_from_payload_tail([3.0 * v for v in tail]) == flags
