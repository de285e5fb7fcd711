"""GPU parity tests of the full-catalogue retrieval mode (BASELINE.json configs[4]): tcgen05 GEMM + fused
per-user top-k through the C ABI, against oracle/manner_oracle.py (retrieval_scores / topk_select /
pooled_users).  The reference has no counterpart (SURVEY 8(d) mode R: parity vs matmul + topk).

Bars: the score matrix within 1e-5 of the fp32 contraction, condition-aware (|ds| <= 1e-5 * sum_d |u_d c_d|);
the top-k selection BIT-EXACT on the kernel's own scores (score desc, id asc on ties); ids against the
oracle's fp32 matmul equal wherever the oracle's k-th / (k+1)-th gap exceeds the score tolerance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import manner_oracle as mo  # noqa: E402  (checker only)

from manner_b200 import data as mdata  # noqa: E402

RTOL = 1e-5


@pytest.fixture(scope="module")
def rt():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from manner_b200 import retrieval

    return retrieval


@pytest.fixture(autouse=True, params=[0, 1], ids=["cta_group1", "cta_pair"])
def pair_mode(request, rt):
    """Every test runs on both GEMM pipelines: one CTA per user tile (tcgen05 cta_group::1) and CTA pairs (cta_group::2,
    UMMA M = 256, operands split across the two CTAs' shared memory)."""
    from manner_b200 import ops

    ops.set_tuning(retrieval_pair=request.param)
    yield request.param
    ops.set_tuning(retrieval_pair=1)  # the library default


def _rand_bf16(n, d, seed, scale=None):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, d, generator=g) * (scale if scale is not None else d ** -0.5)).to(torch.bfloat16)


def _check(rt, users, catalog, k, id_offset=0):
    s, i, full = torch.ops.manner_b200.retrieve_topk(users.cuda(), catalog.cuda(), k, id_offset, True)
    torch.cuda.synchronize()
    s, i, full = s.cpu().numpy(), i.cpu().numpy(), full.cpu().numpy()
    # 1. the GEMM: condition-aware 1e-5 against the fp32 contraction of the same bf16 inputs
    ref = mo.retrieval_scores(users, catalog).numpy()
    cond = (users.float().abs() @ catalog.float().abs().T).numpy()
    assert np.all(np.abs(full.astype(np.float64) - ref) <= RTOL * cond + 1e-30), float(np.max(np.abs(full - ref) / (cond + 1e-30)))
    # 2. the selection: bit-exact on the kernel's own score matrix
    want_s, want_i = mo.topk_select(full, k, id_offset)
    np.testing.assert_array_equal(i, want_i)
    np.testing.assert_array_equal(s, want_s)
    return s, i, full


@pytest.mark.parametrize(
    "n_users,n_catalog,dim,k",
    [
        (1, 1, 64, 1),  # smallest legal problem
        (3, 5, 64, 8),  # k > catalogue: unused slots (-inf, -1)
        (130, 300, 128, 10),  # ragged in both tile dimensions
        (128, 256, 768, 100),  # exactly one tile
        (257, 1000, 768, 100),  # 3 user tiles, 4 catalogue tiles, last ones partial
        (64, 5000, 768, 128),  # k at the limit, many compactions per row
    ],
)
def test_retrieval_small_shapes(rt, n_users, n_catalog, dim, k):
    _check(rt, _rand_bf16(n_users, dim, 1), _rand_bf16(n_catalog, dim, 2), k)


def test_retrieval_ties_break_by_lower_id(rt):
    """Duplicate catalogue rows score identically: the lower catalogue id must come first, and an id at the
    k-boundary tie must be the lower one."""
    base = _rand_bf16(40, 128, 3)
    catalog = base.repeat(8, 1)  # row j == row j + 40 == ... : every score appears 8 times
    users = _rand_bf16(70, 128, 4)
    s, i, _ = _check(rt, users, catalog, 12)
    assert np.all(i[:, 0] < 40) and np.all(i[:, 1] == i[:, 0] + 40)
    # the same across many catalogue tiles (the two epilogue warpgroups own alternating tiles and exchange thresholds):
    # every score appears 8 times, 300 rows apart
    catalog2 = _rand_bf16(300, 128, 5).repeat(8, 1)
    s, i, _ = _check(rt, _rand_bf16(200, 128, 6), catalog2, 20)
    assert np.all(i[:, 0] < 300) and np.all(i[:, 1] == i[:, 0] + 300) and np.all(i[:, 7] == i[:, 0] + 2100)
    # all-equal scores (zero users): ids 0..k-1 in order
    s, i, _ = _check(rt, torch.zeros(5, 128, dtype=torch.bfloat16), catalog, 12)
    np.testing.assert_array_equal(i, np.tile(np.arange(12), (5, 1)))
    s, i, _ = _check(rt, torch.zeros(5, 128, dtype=torch.bfloat16), catalog2, 100)
    np.testing.assert_array_equal(i, np.tile(np.arange(100), (5, 1)))


def test_retrieval_slice_4096_by_65536(rt):
    """parity on a 4 096 x 65 536 slice (768-d, top-100) vs matmul + topk."""
    users, catalog = _rand_bf16(4096, 768, 5), _rand_bf16(65536, 768, 6)
    s, i, full = torch.ops.manner_b200.retrieve_topk(users.cuda(), catalog.cuda(), 100, 0, True)
    ref = users.cuda().float() @ catalog.cuda().float().T  # fp32 (no TF32: torch default for matmul is off)
    cond = users.cuda().float().abs() @ catalog.cuda().float().abs().T
    assert bool(torch.all((full.double() - ref.double()).abs() <= RTOL * cond.double()))
    # selection bit-exact on the kernel's scores: (score desc, id asc) == stable descending sort
    order = torch.sort(full, dim=1, descending=True, stable=True)
    assert torch.equal(i, order.indices[:, :100])
    assert torch.equal(s, order.values[:, :100])
    # against the independent fp32 matmul: same ids wherever the reference's scores are separated by more than the tolerance
    ref_sorted = torch.sort(ref, dim=1, descending=True, stable=True)
    same = i == ref_sorted.indices[:, :100]
    gap_ok = (ref_sorted.values[:, :101].diff(dim=1).abs() > 4 * RTOL * cond.max()).all(dim=1)
    assert bool(same[gap_ok].all())
    assert float(same.float().mean()) > 0.999
    # the production call (no score matrix) returns the same lists
    s2, i2, _ = torch.ops.manner_b200.retrieve_topk(users.cuda(), catalog.cuda(), 100, 0, False)
    assert torch.equal(s, s2) and torch.equal(i, i2)


def test_retrieval_is_deterministic_and_persistent_grid_covers_all_tiles(rt):
    """More user tiles than SMs (the persistent loop wraps) and run-to-run identical results."""
    users, catalog = _rand_bf16(128 * 150 + 7, 64, 7), _rand_bf16(700, 64, 8)
    a = torch.ops.manner_b200.retrieve_topk(users.cuda(), catalog.cuda(), 20, 0, False)
    b = torch.ops.manner_b200.retrieve_topk(users.cuda(), catalog.cuda(), 20, 0, False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    full = mo.retrieval_scores(users, catalog).numpy()
    want_s, want_i = mo.topk_select(full, 20)
    got_i = a[1].cpu().numpy()
    # D = 64: fp32 accumulation order differences are tiny; compare where the oracle's boundary gap is clear
    srt = -np.sort(-full, axis=1)
    clear = np.min(np.abs(np.diff(srt[:, :21], axis=1)), axis=1) > 1e-5
    np.testing.assert_array_equal(got_i[clear], want_i[clear])
    assert clear.mean() > 0.9


def test_sharded_catalogue_merge_equals_single_pass(rt):
    """Row-sharded catalogue: per-shard top-k with id offsets + merge kernel == one pass."""
    users, catalog = _rand_bf16(300, 256, 9).cuda(), _rand_bf16(3000, 256, 10).cuda()
    one_s, one_i, _ = torch.ops.manner_b200.retrieve_topk(users, catalog, 50, 0, False)
    bounds = rt.catalog_shard_bounds(3000, 4)
    parts = [torch.ops.manner_b200.retrieve_topk(users, catalog[lo:hi].contiguous(), 50, lo, False) for lo, hi in bounds if hi > lo]
    ms, mi = rt.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi, one_i) and torch.equal(ms, one_s)
    # a shard smaller than k contributes its (-inf, -1) padding without disturbing the merge
    tiny = torch.ops.manner_b200.retrieve_topk(users, catalog[:7].contiguous(), 50, 0, False)
    rest = torch.ops.manner_b200.retrieve_topk(users, catalog[7:].contiguous(), 50, 7, False)
    ms, mi = rt.merge_topk(torch.stack([tiny[0], rest[0]]), torch.stack([tiny[1], rest[1]]))
    assert torch.equal(mi, one_i) and torch.equal(ms, one_s)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pool_users_matches_late_fusion(rt, dtype):
    """mb200_pool_users == bf16(sum of the history rows / H) of cr_module.py:116-123."""
    n_news, dim = 400, 768
    bhv = mdata.synth_behaviours(n_news, 97, seed=11, cand_window=200)
    table = mdata.synth_table(n_news, dim, 1234).to(dtype)
    got = rt.pool_users(table.cuda(), torch.from_numpy(bhv.hist_offsets).cuda(), torch.from_numpy(bhv.hist_ids).cuda()).cpu()
    want = mo.pooled_users(table, bhv.hist_offsets, bhv.hist_ids)
    # same summation order (history order, fp32) and IEEE division: identical before the bf16 rounding
    assert torch.equal(got, want.to(torch.bfloat16))


def test_retriever_end_to_end_from_behaviours(rt):
    """pool_users -> CatalogRetriever.retrieve on one GPU: the user's own history rows rank at the top."""
    n_news, dim = 2048, 128
    g = torch.Generator().manual_seed(12)
    table = torch.randn(n_news, dim, generator=g)
    bhv = mdata.synth_behaviours(n_news, 50, seed=13, cand_window=500)
    users = rt.pool_users(table.cuda(), torch.from_numpy(bhv.hist_offsets).cuda(), torch.from_numpy(bhv.hist_ids).cuda())
    r = rt.CatalogRetriever(table.to(torch.bfloat16).cuda(), k=100)
    s, i = r.retrieve(users)
    full = mo.retrieval_scores(users.cpu(), table.to(torch.bfloat16)).numpy()
    want_s, want_i = mo.topk_select(full, 100)
    assert (i.cpu().numpy() == want_i).mean() > 0.995
    np.testing.assert_allclose(s.cpu().numpy(), want_s, rtol=1e-4, atol=1e-4)


def test_retrieval_rejects_bad_arguments(rt):
    u, c = _rand_bf16(4, 64, 1).cuda(), _rand_bf16(9, 64, 2).cuda()
    from manner_b200 import _native as nat

    with pytest.raises(nat.NativeError, match="unsupported"):
        torch.ops.manner_b200.retrieve_topk(u, c, 129, 0, False)  # k > 128
    with pytest.raises(nat.NativeError, match="unsupported"):
        torch.ops.manner_b200.retrieve_topk(_rand_bf16(4, 96, 1).cuda(), _rand_bf16(9, 96, 2).cuda(), 4, 0, False)  # dim % 64
    with pytest.raises(RuntimeError, match="no CPU path"):
        torch.ops.manner_b200.retrieve_topk(u.cpu(), c.cpu(), 4, 0, False)
