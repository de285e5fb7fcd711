"""What can be checked about memory safety without compute-sanitizer (closed on this GPU pool, profiles/r2_l_compute_sanitizer_refused.log):
every caller-owned OUTPUT and WORKSPACE buffer of the C ABI is handed over at exactly the size the header asks for, with canary
bands on both sides that must be intact afterwards (out-of-bounds stores), and the workspace is poisoned with two different
patterns between which every output must stay bit-identical (reads of memory the call did not write first)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from manner_b200 import _native as nat  # noqa: E402
from manner_b200 import data as mdata  # noqa: E402

GUARD = 4096
CANARY = 0xA5


class Guarded:
    """`nbytes` of device memory between two canary bands (256-byte aligned interior)."""

    def __init__(self, nbytes: int, fill: int = 0) -> None:
        self.n = int(nbytes)
        pad = (256 - self.n % 256) % 256
        self.buf = torch.full((GUARD + self.n + pad + GUARD,), CANARY, dtype=torch.uint8, device="cuda:0")
        self.inner = self.buf[GUARD : GUARD + self.n]
        self.inner.fill_(fill)
        assert self.inner.data_ptr() % 256 == 0
        self.ptr = self.inner.data_ptr()

    def intact(self) -> bool:
        return bool((self.buf[:GUARD] == CANARY).all()) and bool((self.buf[GUARD + self.n :] == CANARY).all())

    def bytes(self) -> bytes:
        return self.inner.cpu().numpy().tobytes()


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return nat.lib()


def _eval_once(lib, tables, bhv, aspects, weights, poison, dim, loss_kind=0):
    dev = "cuda:0"
    t = [x.to(dev).contiguous() for x in tables]
    ho, hi = torch.from_numpy(bhv.hist_offsets).to(dev), torch.from_numpy(bhv.hist_ids).to(dev)
    co, ci = torch.from_numpy(bhv.cand_offsets).to(dev), torch.from_numpy(bhv.cand_ids).to(dev)
    lab = torch.from_numpy(bhv.labels).to(dev)
    w = torch.tensor(weights, dtype=torch.float32, device=dev)
    n_w, n_impr, n_cand = w.shape[0], bhv.n_impressions, bhv.n_cand
    cat = torch.from_numpy(aspects["category"]).to(dev) if aspects else None
    sent = torch.from_numpy(aspects["sentiment"]).to(dev) if aspects else None
    pads = None
    if loss_kind:
        pads = (torch.from_numpy(mdata.step_pads(bhv.hist_offsets, 8)).to(dev), torch.from_numpy(mdata.step_pads(bhv.cand_offsets, 8)).to(dev))
    out = {
        "scores": Guarded(n_cand * 4), "per_impr": Guarded(n_w * n_impr * nat.NUM_METRICS * 4), "sums": Guarded(n_w * nat.NUM_METRICS * 8),
        "flags": Guarded(4), "loss": Guarded(n_impr * 4),
    }
    d = nat.EvalDesc()
    d.struct_size = ctypes.sizeof(nat.EvalDesc)
    d.n_modules, d.dtype, d.dim, d.active_modules_mask = len(t), (nat.F32 if t[0].dtype == torch.float32 else nat.BF16), dim, (1 << len(t)) - 1
    d.n_news, d.row_stride = t[0].shape[0], t[0].stride(0)
    for m, x in enumerate(t):
        d.tables[m] = x.data_ptr()
    d.n_impressions = n_impr
    d.hist_offsets, d.hist_ids, d.cand_offsets, d.cand_ids, d.labels = ho.data_ptr(), hi.data_ptr(), co.data_ptr(), ci.data_ptr(), lab.data_ptr()
    d.max_cand, d.zscore, d.n_weightings, d.weights, d.k0, d.k1 = bhv.max_cand, 1, n_w, w.data_ptr(), 5, 10
    if aspects:
        d.news_category, d.news_sentiment, d.num_categ_classes, d.num_sent_classes = cat.data_ptr(), sent.data_ptr(), 19, 4
    d.scores, d.per_impression, d.sums, d.flags = out["scores"].ptr, out["per_impr"].ptr, out["sums"].ptr, out["flags"].ptr
    if loss_kind:
        d.loss_kind, d.loss_temperature, d.loss_per_impression = loss_kind, 0.36, out["loss"].ptr
        d.cand_pad = pads[1].data_ptr()
    need = lib.mb200_eval_workspace_bytes(ctypes.byref(d))
    assert need > 0
    ws = Guarded(need, fill=poison)
    d.workspace, d.workspace_bytes = ws.ptr, need
    nat.check(lib.mb200_score_eval(ctypes.byref(d), torch.cuda.current_stream().cuda_stream), "mb200_score_eval")
    torch.cuda.synchronize()
    for name, g in list(out.items()) + [("workspace", ws)]:
        assert g.intact(), f"{name}: a canary band was overwritten"
    assert int(out["flags"].inner.view(torch.int32).item()) & ~nat.FLAG_OUTSIDE_UNIT == 0
    keep = ["scores", "per_impr", "sums"] + (["loss"] if loss_kind else [])
    return {k: out[k].bytes() for k in keep}


@pytest.mark.parametrize("case", ["f32_aspects_w2", "bf16_sweep", "dim136_loss"])
def test_score_eval_stays_inside_its_buffers_and_ignores_workspace_contents(lib, case):
    n_news = 600
    bhv = mdata.synth_behaviours(n_news, 700, seed=21, cand_window=400)
    if case == "f32_aspects_w2":
        tables = [mdata.synth_table(n_news, 768, s) for s in mdata.TABLE_SEEDS]
        args = (tables, bhv, mdata.synth_aspects(n_news), [[1.0, 0.4, 0.2], [1.0, 0.0, 0.7]])
        kw = dict(dim=768)
    elif case == "bf16_sweep":
        tables = [mdata.synth_table(n_news, 768, s, torch.bfloat16) for s in mdata.TABLE_SEEDS]
        args = (tables, bhv, None, [[1.0, a / 4.0, b / 4.0] for a in range(5) for b in range(5)])
        kw = dict(dim=768)
    else:
        args = ([mdata.synth_table(n_news, 136, 5)], bhv, None, [[1.0]])
        kw = dict(dim=136, loss_kind=nat.LOSS_CE)
    a = _eval_once(lib, *args, poison=0xFF, **kw)
    b = _eval_once(lib, *args, poison=0x00, **kw)
    c = _eval_once(lib, *args, poison=0x7F, **kw)
    assert a == b == c, "an output depends on what the workspace held before the call"


@pytest.mark.parametrize("bounded", [False, True])
def test_pooled_auc_stays_inside_its_buffers(lib, bounded):
    g = np.random.default_rng(3)
    n = 50_000
    preds = torch.from_numpy((g.standard_normal(n) * 2).astype(np.float32)).cuda()
    labels_h = (g.random(n) < 0.05).astype(np.uint8)
    labels = torch.from_numpy(labels_h).cuda()
    flags = torch.tensor([nat.FLAG_OUTSIDE_UNIT], dtype=torch.int32, device="cuda:0")
    cap = int(labels_h.sum())
    results = []
    for poison in (0xFF, 0x00):
        need = lib.mb200_pooled_auc_bounded_workspace_bytes(n, cap) if bounded else lib.mb200_pooled_auc_workspace_bytes(n)
        ws, out = Guarded(need, fill=poison), Guarded(32, fill=poison)
        stream = torch.cuda.current_stream().cuda_stream
        if bounded:
            nat.check(lib.mb200_pooled_auc_bounded(preds.data_ptr(), labels.data_ptr(), n, cap, 2, flags.data_ptr(), ws.ptr, need, out.ptr, stream), "bounded")
        else:
            nat.check(lib.mb200_pooled_auc(preds.data_ptr(), labels.data_ptr(), n, 2, flags.data_ptr(), ws.ptr, need, out.ptr, stream), "pooled_auc")
        torch.cuda.synchronize()
        assert ws.intact() and out.intact()
        results.append(out.bytes())
    assert results[0] == results[1]


def test_rank_metrics_stays_inside_its_buffers(lib):
    g = np.random.default_rng(4)
    sizes = g.integers(2, 60, 500)
    off = np.zeros(501, dtype=np.int32)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    preds = torch.from_numpy(g.standard_normal(n).astype(np.float32)).cuda()
    labels = torch.from_numpy((g.random(n) < 0.2).astype(np.uint8)).cuda()
    offs = torch.from_numpy(off).cuda()
    results = []
    for poison in (0xFF, 0x00):
        d = nat.MetricsDesc()
        d.struct_size = ctypes.sizeof(nat.MetricsDesc)
        d.k0, d.k1, d.max_cand, d.n_impressions = 5, 10, int(sizes.max()), 500
        d.preds, d.labels, d.cand_offsets = preds.data_ptr(), labels.data_ptr(), offs.data_ptr()
        per, sums, flags = Guarded(500 * nat.NUM_METRICS * 4), Guarded(nat.NUM_METRICS * 8), Guarded(4)
        d.per_impression, d.sums, d.flags = per.ptr, sums.ptr, flags.ptr
        need = lib.mb200_metrics_workspace_bytes(ctypes.byref(d))
        ws = Guarded(need, fill=poison)
        d.workspace, d.workspace_bytes = ws.ptr, need
        nat.check(lib.mb200_rank_metrics(ctypes.byref(d), torch.cuda.current_stream().cuda_stream), "mb200_rank_metrics")
        torch.cuda.synchronize()
        assert per.intact() and sums.intact() and flags.intact() and ws.intact()
        results.append((per.bytes(), sums.bytes()))
    assert results[0] == results[1]
