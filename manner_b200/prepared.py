"""PreparedPass: one evaluation pass over a FIXED behaviour set with everything that does not change from pass to pass done
ahead of time -- device buffers, descriptors, workspaces, the pinned result buffer -- so that a pass is a handful of C-ABI calls.

Why: the reference evaluates the same validation / test split after every epoch (cr_module.py:214-274); here such a pass takes
1.7 ms on the GPU, and the generic path (``ScoreEvaluator.upload`` + ``evaluate``: tensor allocations, custom-op dispatch,
argument checks, descriptor filling) kept the stream waiting ~0.2 ms for the host before the fused kernel was even launched
(tools/host_overhead.py).  ``ScoreEvaluator.prepare(...)`` does that work once; ``run()`` then queues

    mb200_upload_begin  ->  mb200_score_eval  ->  [mb200_step_loss]  ->  mb200_pooled_auc | (build keys, sort, exchange post / finish)

and reads the result back with ONE pinned device -> host copy.  Same kernels, same numbers as ``evaluate``
(tests/test_gpu_prepared.py)."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch
from torch import Tensor

from . import _native as nat
from . import dist as mdist
from . import ops
from .data import Behaviours

_LOSS = {None: nat.LOSS_NONE, "ce": nat.LOSS_CE, "supcon": nat.LOSS_SUPCON}


class PreparedPass:
    def __init__(self, ev, bhv: Behaviours, pinned: Optional[Dict[str, object]] = None, weights=None, zscore: bool = False,
                 pooled_auc: bool = False, loss: Optional[str] = None, temperature: float = 0.1, step_batch: Optional[int] = None,
                 segments: int = 5, distributed: bool = False, group=None, pos_cap: Optional[int] = None, want_scores: bool = False,
                 resident: bool = False, worker_segments: int = 0) -> None:
        from .evaluator import EvalResult  # noqa: F401  (cycle-free import at call time)

        lib = nat.lib()
        if ev.n_table_shards != 1:
            raise ValueError("prepared passes use replicated tables")
        if (loss is not None or ev.attn_logits is not None) and step_batch is None:
            step_batch = 8
        self.ev, self.dev = ev, ev.device
        self.src = pinned if pinned is not None else ev.pin(bhv, step_batch)
        if (loss is not None or ev.attn_logits is not None) and "hist_pad" not in self.src:
            raise ValueError("early fusion / the losses need the step pads: pin(bhv, step_batch)")
        # page-locked scratch the upload writes its segment marks to: this pass's own (passes over one pinned set may be in flight together)
        self.marks = torch.zeros(nat.MAX_UPLOAD_SEGMENTS, dtype=torch.int32).pin_memory()
        self.n_impr, self.n_cand = bhv.n_impressions, bhv.n_cand
        if self.n_impr < 1:
            raise ValueError("empty behaviour set")
        self.pooled_auc, self.loss = bool(pooled_auc), loss
        self.distributed = bool(distributed)
        # resident: the behaviours are copied to the device once, here, and every pass runs on them (no upload in run())
        self.resident = bool(resident)
        if self.distributed and not (ev.exchange == "p2p" and (group is None or torch.distributed.get_backend(group) == "nccl")):
            raise ValueError("prepared multi-GPU passes use the fused exchange (ScoreEvaluator(exchange='p2p') on NCCL ranks)")
        n_mod, dev = ev.n_modules, self.dev
        segments = max(1, min(int(segments), nat.MAX_UPLOAD_SEGMENTS))
        if self.n_impr < 64 * segments:
            segments = 1

        with torch.cuda.device(dev):
            self.copy_stream = ev._copy_stream or torch.cuda.Stream(dev)
            ev._copy_stream = self.copy_stream
            stream = torch.cuda.current_stream(dev).cuda_stream
            # ---- inputs on the device + the upload descriptor
            self.d = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in self.src.items() if isinstance(v, Tensor) and k != "marks"}
            if self.resident:
                for k, t in self.d.items():
                    t.copy_(self.src[k], non_blocking=True)
            self.ready = torch.zeros(1, dtype=torch.int32, device=dev)
            up = nat.UploadDesc()
            up.struct_size = ctypes.sizeof(nat.UploadDesc)
            # Every segment copy is queued by this thread, in front of the kernel launch (~2 us per copy).  ``worker_segments`` = k hands
            # the LAST k segments to the library's thread instead, which queues them while this thread is already launching the kernel
            # (the launch reaches the GPU ~20 us earlier) -- but a tool that serialises CUDA API calls behind a running kernel (ncu)
            # can then keep those copies from ever being issued while the kernel waits for them, so it is opt-in.
            up.n_segments, up.segments_first, up.n_impressions = segments, max(1, segments - max(0, int(worker_segments))), self.n_impr
            for name in ("hist_offsets", "hist_ids", "cand_offsets", "cand_ids", "labels"):
                setattr(up, "h_" + name, self.src[name].data_ptr())
                setattr(up, "d_" + name, self.d[name].data_ptr())
            for name in ("hist_pad", "cand_pad"):
                if name in self.d:
                    setattr(up, "h_" + name, self.src[name].data_ptr())
                    setattr(up, "d_" + name, self.d[name].data_ptr())
            up.ready, up.h_marks, up.copy_stream = self.ready.data_ptr(), self.marks.data_ptr(), self.copy_stream.cuda_stream
            self.up = up
            self.h2d_bytes = sum(v.numel() * v.element_size() for k, v in self.src.items() if isinstance(v, Tensor) and k != "marks")

            # ---- weights, active modules (a module whose weight is 0 everywhere is never gathered: ensemble_module.py:37-46,100-107)
            active = (1 << n_mod) - 1
            self.w = None
            if isinstance(weights, Tensor) and weights.is_cuda:
                self.w = weights.float().reshape(-1, n_mod).contiguous()
            elif weights is not None:
                wh = torch.as_tensor(weights, dtype=torch.float32).reshape(-1, n_mod)
                active = 1
                for m in range(1, n_mod):
                    if bool((wh[:, m] != 0).any()):
                        active |= 1 << m
                self.w = wh.to(dev).contiguous()
            self.n_w = 1 if self.w is None else self.w.shape[0]
            n_block = self.n_w * nat.NUM_METRICS
            self.n_block = n_block
            loss_kind = _LOSS[loss]
            if loss is not None and self.n_w >= 16 and ev.news_category is None:
                raise ValueError("the loss is not computed in the aspect-weight sweep mode: evaluate it with a single weighting")

            # ---- outputs: everything the host reads lives in ONE fp64 buffer
            #   single GPU : [sums W*15][auc, P, N, sum2][flags word][loss sum, steps]
            #   fused multi: [payload W*15 + 6 (reduced)][sum2, P, N, exchange flags (int64 bit patterns)][loss sum, steps]
            self.scores = torch.empty(self.n_cand if (pooled_auc or want_scores) else 0, dtype=torch.float32, device=dev)
            n_payload = n_block + nat.PAYLOAD_TAIL
            n_res = (n_payload + 4 + 2) if self.distributed else (n_block + 4 + 1 + 2)
            self.result = torch.zeros(n_res, dtype=torch.float64, device=dev)
            self.host = torch.zeros(n_res, dtype=torch.float64).pin_memory()
            base = self.result.data_ptr()
            self.loss_per = torch.empty(self.n_impr if loss_kind else 0, dtype=torch.float32, device=dev)
            if self.distributed:
                self.payload = torch.zeros(n_payload, dtype=torch.float64, device=dev)
                self.flags = torch.zeros(1, dtype=torch.int32, device=dev)
                sums_ptr, flags_ptr = self.payload.data_ptr(), self.flags.data_ptr()
                self.off_loss = n_payload + 4
            else:
                sums_ptr, flags_ptr = base, base + 8 * (n_block + 4)
                self.off_loss = n_block + 5
            self.loss_ptr = base + 8 * self.off_loss

            # ---- the evaluation descriptor
            e = nat.EvalDesc()
            e.struct_size = ctypes.sizeof(nat.EvalDesc)
            t0 = ev.tables[0]
            e.n_modules, e.dtype, e.dim, e.active_modules_mask = n_mod, (nat.F32 if t0.dtype == torch.float32 else nat.BF16), t0.shape[1], active
            e.n_news, e.row_stride = t0.shape[0], t0.stride(0)
            for m, t in enumerate(ev.tables):
                e.tables[m] = t.data_ptr()
            e.n_impressions = self.n_impr
            e.hist_offsets, e.hist_ids = self.d["hist_offsets"].data_ptr(), self.d["hist_ids"].data_ptr()
            e.cand_offsets, e.cand_ids, e.labels = self.d["cand_offsets"].data_ptr(), self.d["cand_ids"].data_ptr(), self.d["labels"].data_ptr()
            e.max_cand, e.zscore, e.n_weightings = self.src["max_cand"], int(zscore), self.n_w
            e.weights = None if self.w is None else self.w.data_ptr()
            e.k0, e.k1 = ev.ks
            if ev.news_category is not None:
                e.news_category, e.news_sentiment = ev.news_category.data_ptr(), ev.news_sentiment.data_ptr()
                e.num_categ_classes, e.num_sent_classes = ev.num_categ_classes, ev.num_sent_classes
            e.scores = self.scores.data_ptr() if self.scores.numel() else None
            e.scores_weighting, e.pack_payload = 0, int(self.distributed)
            e.sums, e.flags, e.zero_flags = sums_ptr, flags_ptr, 1
            if ev.attn_logits is not None:
                for m, a in enumerate(ev.attn_logits):
                    e.attn_logits[m] = None if a is None else a.data_ptr()
                e.hist_pad = self.d["hist_pad"].data_ptr()
            if loss_kind:
                e.loss_kind, e.loss_temperature = loss_kind, float(temperature)
                e.cand_pad, e.loss_per_impression = self.d["cand_pad"].data_ptr(), self.loss_per.data_ptr()
            if not self.resident:
                e.ready, e.ready_segments = self.ready.data_ptr(), segments
            need = lib.mb200_eval_workspace_bytes(ctypes.byref(e))
            if need == 0:
                e.workspace, e.workspace_bytes = None, 0
                nat.check(lib.mb200_score_eval(ctypes.byref(e), stream), "mb200_score_eval")
                raise nat.NativeError("mb200_eval_workspace_bytes returned 0")
            self.ws_eval = torch.empty(need, dtype=torch.uint8, device=dev)
            e.workspace, e.workspace_bytes = self.ws_eval.data_ptr(), need
            self.e = e
            self.loss_kind, self.step_batch = loss_kind, int(self.src.get("step_batch", 8))

            # ---- pooled AUROC
            n = self.n_cand
            self.xd = None
            if self.distributed:
                cap = pos_cap if pos_cap is not None else mdist.agree_pos_cap(self.src["n_pos"], dev, group)
                if cap < self.src["n_pos"]:
                    raise ValueError("pos_cap is below this shard's number of positives")
                if ev._p2p is None or not ev._p2p.fits(n_payload, cap):
                    if ev._p2p is not None:
                        ev._p2p.close()
                    ev._p2p = mdist.P2PExchange(dev, n_payload, max(cap, 1), group)
                self.p2p = ev._p2p
                x = nat.ExchangeDesc()
                x.struct_size = ctypes.sizeof(nat.ExchangeDesc)
                x.n_ranks, x.my_rank = self.p2p.world, self.p2p.rank
                x.n_payload, x.pos_capacity = n_payload, self.p2p.pos_cap
                x.outside_index = (n_block + 1 + 2) if pooled_auc else -1
                for r, t in enumerate(self.p2p.peers):
                    x.mailbox[r] = t.data_ptr()
                x.payload = self.payload.data_ptr()
                if pooled_auc:
                    self.keys = torch.empty(n, dtype=torch.int32, device=dev)
                    self.sorted_keys = torch.empty(n, dtype=torch.int32, device=dev)
                    self.pos_keys = torch.empty(n, dtype=torch.int32, device=dev)
                    self.n_pos_dev = torch.zeros(1, dtype=torch.int64, device=dev)
                    self.ws_sort = torch.empty(lib.mb200_auc_sort_workspace_bytes(n), dtype=torch.uint8, device=dev)
                    x.pos_keys, x.n_pos, x.sorted_neg, x.n_rows = self.pos_keys.data_ptr(), self.n_pos_dev.data_ptr(), self.sorted_keys.data_ptr(), n
                else:
                    self.empty_i32 = torch.zeros(2, dtype=torch.int32, device=dev)
                    self.n_pos_dev = torch.zeros(1, dtype=torch.int64, device=dev)
                    x.pos_keys, x.n_pos, x.sorted_neg, x.n_rows = self.empty_i32.data_ptr(), self.n_pos_dev.data_ptr(), self.empty_i32.data_ptr(), 0
                self.ws_x = torch.empty(lib.mb200_exchange_workspace_bytes(x.n_rows), dtype=torch.uint8, device=dev)
                x.workspace, x.workspace_bytes = self.ws_x.data_ptr(), self.ws_x.numel()
                x.out_payload, x.out_stats, x.flags = base, base + 8 * n_payload, base + 8 * (n_payload + 3)
                self.xd = x
                self.group = group
            elif pooled_auc:
                self.ws_auc = torch.empty(lib.mb200_pooled_auc_workspace_bytes(n), dtype=torch.uint8, device=dev)
                self.auc_ptr = base + 8 * n_block
            self.lib = lib
            self.d2h_bytes = n_res * 8

    # ------------------------------------------------------------------------------------------------------
    def run(self):
        """Upload (overlapped) + pass + read-back.  Returns an ``EvalResult``."""
        self.launch()
        return self.read()

    def launch(self) -> None:
        """Queues one pass on the current stream (with the pipelined upload in front unless the pass is ``resident``); no host
        synchronisation.  ``read()`` fetches the result of the most recent launch."""
        lib, e = self.lib, self.e
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        if not self.resident:
            nat.check(lib.mb200_upload_begin(ctypes.byref(self.up), stream), "mb200_upload_begin")
        nat.check(lib.mb200_score_eval(ctypes.byref(e), stream), "mb200_score_eval")
        if not self.resident:
            nat.check(lib.mb200_upload_finish(ctypes.byref(self.up)), "mb200_upload_finish")  # every copy is queued from here on
        if self.loss_kind:
            nat.check(lib.mb200_step_loss(self.loss_per.data_ptr(), self.n_impr, self.step_batch, self.loss_kind, self.d["cand_offsets"].data_ptr(),
                                          self.d["labels"].data_ptr(), self.loss_ptr, stream), "mb200_step_loss")
        if self.xd is not None:
            x = self.xd
            if self.pooled_auc:
                n = self.n_cand
                nat.check(lib.mb200_auc_build_keys(self.scores.data_ptr(), self.d["labels"].data_ptr(), n, 0, None, self.keys.data_ptr(),
                                                   self.pos_keys.data_ptr(), self.n_pos_dev.data_ptr(), stream), "mb200_auc_build_keys")
                nat.check(lib.mb200_auc_sort_keys(self.keys.data_ptr(), self.sorted_keys.data_ptr(), n, self.ws_sort.data_ptr(), self.ws_sort.numel(), stream),
                          "mb200_auc_sort_keys")
            self.p2p.epoch += 1
            x.epoch = self.p2p.epoch
            nat.check(lib.mb200_exchange_post(ctypes.byref(x), stream), "mb200_exchange_post")
            nat.check(lib.mb200_exchange_finish(ctypes.byref(x), stream), "mb200_exchange_finish")
        elif self.pooled_auc:
            nat.check(lib.mb200_pooled_auc(self.scores.data_ptr(), self.d["labels"].data_ptr(), self.n_cand, 2, e.flags, self.ws_auc.data_ptr(),
                                           self.ws_auc.numel(), self.auc_ptr, stream), "mb200_pooled_auc")
        if self.loss_kind and self.distributed:
            torch.distributed.all_reduce(self.result[self.off_loss : self.off_loss + 2], op=torch.distributed.ReduceOp.SUM, group=self.group)

    def read(self):
        """The ONE device -> host read of a pass (pinned buffer), decoded into an ``EvalResult``."""
        from .evaluator import EvalResult

        cur = torch.cuda.current_stream(self.dev)
        self.host.copy_(self.result, non_blocking=True)
        cur.synchronize()
        return self._decode(EvalResult)

    def _decode(self, EvalResult):
        h = self.host.numpy()
        nb = self.n_block
        auc = counts = None
        if self.distributed:
            tail = self.host[nb + nat.PAYLOAD_TAIL : nb + nat.PAYLOAD_TAIL + 4].view(torch.int64).tolist()
            if tail[3] & nat.FLAG_EXCHANGE_TIMEOUT:
                raise nat.NativeError("fused multi-GPU exchange: a peer GPU's stores did not arrive within 4 s (a rank died or skipped the call)")
            if tail[3] & nat.FLAG_POS_OVERFLOW:
                raise nat.NativeError("fused multi-GPU exchange: a rank had more positives than the agreed pos_cap (dist.agree_pos_cap)")
            sums = h[:nb].reshape(self.n_w, nat.NUM_METRICS).copy()
            n_total = int(round(h[nb]))
            flags = mdist.flags_from_payload_tail(h[nb + 1 : nb + 1 + mdist.N_FLAG_BITS])
            if self.pooled_auc:
                s2, p, n = tail[0], tail[1], tail[2]
                auc, counts = (s2 / (2.0 * p * n) if p > 0 and n > 0 else 0.0), (p, n)
        else:
            sums = h[:nb].reshape(self.n_w, nat.NUM_METRICS).copy()
            n_total = self.n_impr
            flags = int(self.host[nb + 4 : nb + 5].view(torch.int32)[0])
            if self.pooled_auc:
                auc, counts = float(h[nb]), (int(h[nb + 1]), int(h[nb + 2]))
        loss_value = None
        if self.loss_kind:
            ls = h[self.off_loss : self.off_loss + 2]
            loss_value = float(ls[0] / ls[1]) if ls[1] > 0 else 0.0
        if flags & nat.FLAG_UPLOAD_TIMEOUT:
            raise nat.NativeError("pipelined upload: a segment of the behaviour set did not reach the device within 4 s (copy stream stalled?)")
        if flags & (nat.FLAG_BAD_ID | nat.FLAG_CAND_OVERFLOW | nat.FLAG_BAD_ASPECT):
            raise nat.NativeError(f"manner_b200 kernels flagged bad input (flags={flags}): "
                                  "1=row id outside the table, 2=impression longer than max_cand, 8=aspect label outside [0, num_classes)")
        return EvalResult(sums=sums, n_impressions=n_total, flags=flags, ks=self.ev.ks, has_aspects=self.ev.news_category is not None,
                          auc=auc, auc_counts=counts, scores=self.scores if self.scores.numel() else None, per_impression=None,
                          d2h_bytes=self.d2h_bytes, loss=loss_value)
