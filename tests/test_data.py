"""CPU: host logic -- CSR contract, segment-id conversion, generator shape, sharding."""
import numpy as np
import pytest
import torch

from manner_b200 import data as mdata
from oracle import manner_oracle as mo


def test_generator_matches_the_mind_shape_and_is_seeded():
    a = mdata.synth_behaviours(4096, 2000, seed=11)
    b = mdata.synth_behaviours(4096, 2000, seed=11)
    for f in ("hist_offsets", "hist_ids", "cand_offsets", "cand_ids", "labels"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
    a.validate(4096)
    h, c = np.diff(a.hist_offsets), np.diff(a.cand_offsets)
    assert h.min() >= 1 and h.max() <= mdata.MAX_HISTORY and c.min() >= 2 and c.max() <= 300
    assert 15 < h.mean() < 30 and 28 < c.mean() < 45
    seg = np.repeat(np.arange(a.n_impressions), c)
    pos = np.bincount(seg, weights=a.labels, minlength=a.n_impressions)
    assert pos.min() >= 1 and np.all(pos < c)
    key = seg.astype(np.int64) * 4096 + a.cand_ids
    assert np.unique(key).size == key.size  # no duplicate candidate inside an impression


def test_segment_ids_round_trip_to_csr():
    bhv = mdata.synth_behaviours(512, 40, seed=3, cand_window=300)
    ob = mo.Behaviours(bhv.hist_offsets, bhv.hist_ids, bhv.cand_offsets, bhv.cand_ids, bhv.labels)
    batch = mo.step_batch(ob, 8, 24)  # the reference's MINDRecBatch for impressions 8..23
    back = mdata.from_segment_ids(batch["batch_hist"], batch["x_hist"]["news_row"], batch["batch_cand"], batch["x_cand"]["news_row"], batch["labels"])
    ref = bhv.slice(8, 24)
    for f in ("hist_offsets", "hist_ids", "cand_offsets", "cand_ids", "labels"):
        np.testing.assert_array_equal(getattr(back, f), getattr(ref, f))
    with pytest.raises(ValueError):
        mdata.from_segment_ids(torch.tensor([1, 0]), torch.tensor([0, 0]), torch.tensor([0, 1]), torch.tensor([0, 0]), torch.tensor([0.0, 1.0]))


def test_validate_rejects_bad_behaviours():
    good = mdata.synth_behaviours(128, 10, seed=1, cand_window=100)
    good.validate(128)
    with pytest.raises(ValueError):
        good.validate(16)  # ids out of range
    bad = mdata.Behaviours(good.hist_offsets, good.hist_ids, good.cand_offsets, good.cand_ids, (good.labels * 2).astype(np.uint8))
    with pytest.raises(ValueError):
        bad.validate(128)
    empty_hist = mdata.Behaviours(np.array([0, 0, 1], np.int32), np.array([1], np.int32), np.array([0, 1, 2], np.int32),
                                  np.array([1, 2], np.int32), np.array([1, 0], np.uint8))
    with pytest.raises(ValueError):
        empty_hist.validate(8)  # the reference drops users with an empty history at parse time


def test_shards_are_balanced_by_rows_and_cover_everything():
    bhv = mdata.synth_behaviours(4096, 3000, seed=5)
    for world in (1, 2, 3, 8):
        b = mdata.balanced_shard_bounds(bhv, world)
        assert b[0] == 0 and b[-1] == bhv.n_impressions and np.all(np.diff(b) > 0)
        work = bhv.hist_offsets.astype(np.int64) + bhv.cand_offsets
        per = np.diff(work[b])
        assert per.max() - per.min() <= 2 * (mdata.MAX_HISTORY + 300)
        parts = [bhv.slice(int(b[r]), int(b[r + 1])) for r in range(world)]
        assert sum(p.n_impressions for p in parts) == bhv.n_impressions
        np.testing.assert_array_equal(np.concatenate([p.cand_ids for p in parts]), bhv.cand_ids)
        np.testing.assert_array_equal(np.concatenate([p.labels for p in parts]), bhv.labels)
        for p in parts:
            p.validate(4096)


def test_step_aligned_shards_keep_the_references_step_structure():
    """Early fusion / losses: pads and step means of a shard equal the corresponding slice of the whole set's only when
    the shard starts on a step boundary."""
    bhv = mdata.synth_behaviours(2048, 1003, seed=6)
    whole_h, whole_c = mdata.step_pads(bhv.hist_offsets, 8), mdata.step_pads(bhv.cand_offsets, 8)
    for world in (2, 3, 5):
        b = mdata.balanced_shard_bounds(bhv, world, align=8)
        assert b[0] == 0 and b[-1] == bhv.n_impressions and np.all(np.diff(b) >= 0) and np.all(b[1:-1] % 8 == 0)
        for r in range(world):
            part = bhv.slice(int(b[r]), int(b[r + 1]))
            np.testing.assert_array_equal(mdata.step_pads(part.hist_offsets, 8), whole_h[b[r]:b[r + 1]])
            np.testing.assert_array_equal(mdata.step_pads(part.cand_offsets, 8), whole_c[b[r]:b[r + 1]])
    assert mdata.step_pads(np.array([0, 3, 4, 9, 9]), 2).tolist() == [0, 2, 0, 5]


def test_algorithmic_bytes_formula():
    bhv = mdata.synth_behaviours(1024, 100, seed=2)
    rows = bhv.n_hist + bhv.n_cand
    want = 2 * 4 * 768 * rows + 4 * rows + bhv.n_cand + 8 * 101 + 4 * bhv.n_cand
    assert bhv.algorithmic_bytes(2, 768, 4, True) == want
    assert bhv.algorithmic_bytes(2, 768, 4, False) == want - 4 * bhv.n_cand
