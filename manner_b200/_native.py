"""ctypes binding of libmanner_b200.so (include/manner_b200.h).  No torch types cross this boundary:
raw device pointers, sizes and the stream handle only.

There is NO CPU fallback: if the shared library is missing, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p
from typing import Optional

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmanner_b200.so")

ABI_VERSION = 4
MAX_MODULES = 4
MAX_TABLE_SHARDS = 8
MAX_UPLOAD_SEGMENTS = 32
MAX_K = 31
MAX_CLASSES = 64
NUM_METRICS = 15
PAYLOAD_TAIL = 6

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, 1, 2, 3, 4
FLAG_BAD_ID, FLAG_CAND_OVERFLOW, FLAG_OUTSIDE_UNIT, FLAG_BAD_ASPECT = 1, 2, 4, 8
FLAG_EXCHANGE_TIMEOUT, FLAG_POS_OVERFLOW, FLAG_UPLOAD_TIMEOUT = 16, 32, 64
# flag bits a packed payload carries behind the impression count (mb200_eval_desc.pack_payload), in this order
PAYLOAD_FLAG_BITS = (1, 2, 4, 8, 64)
F32, BF16 = 0, 1

# metric slots
M_MRR, M_NDCG_K0, M_NDCG_K1, M_GAUC, M_GAUC_VALID = 0, 1, 2, 3, 4
M_CATEG_DIV_K0, M_CATEG_DIV_K1, M_SENT_DIV_K0, M_SENT_DIV_K1 = 5, 6, 7, 8
M_CATEG_PERS_K0, M_CATEG_PERS_K1, M_SENT_PERS_K0, M_SENT_PERS_K1 = 9, 10, 11, 12
M_LOSS, M_LOSS_NONZERO = 13, 14
LOSS_NONE, LOSS_CE, LOSS_SUPCON = 0, 1, 2


class EvalDesc(Structure):
    """mb200_eval_desc, field for field."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("n_modules", c_int32),
        ("dtype", c_int32),
        ("dim", c_int32),
        ("active_modules_mask", c_int32),
        ("n_news", c_int64),
        ("row_stride", c_int64),
        ("tables", c_void_p * MAX_MODULES),
        ("n_impressions", c_int64),
        ("hist_offsets", c_void_p),
        ("hist_ids", c_void_p),
        ("cand_offsets", c_void_p),
        ("cand_ids", c_void_p),
        ("labels", c_void_p),
        ("max_cand", c_int32),
        ("zscore", c_int32),
        ("n_weightings", c_int32),
        ("weights", c_void_p),
        ("k0", c_int32),
        ("k1", c_int32),
        ("news_category", c_void_p),
        ("news_sentiment", c_void_p),
        ("num_categ_classes", c_int32),
        ("num_sent_classes", c_int32),
        ("scores", c_void_p),
        ("scores_weighting", c_int32),
        ("pack_payload", c_int32),
        ("per_impression", c_void_p),
        ("sums", c_void_p),
        ("flags", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
        ("attn_logits", c_void_p * MAX_MODULES),
        ("hist_pad", c_void_p),
        ("loss_kind", c_int32),
        ("loss_temperature", c_float),
        ("cand_pad", c_void_p),
        ("loss_per_impression", c_void_p),
        ("n_table_shards", c_int32),
        ("table_shard_shift", c_int32),
        ("table_shards", (c_void_p * MAX_TABLE_SHARDS) * MAX_MODULES),
        ("ready", c_void_p),
        ("ready_segments", c_int32),
        ("zero_flags", c_int32),
    ]


class UploadDesc(Structure):
    """mb200_upload_desc, field for field."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("n_segments", c_int32),
        ("segments_first", c_int32),
        ("reserved", c_int32),
        ("n_impressions", c_int64),
        ("h_hist_offsets", c_void_p),
        ("h_hist_ids", c_void_p),
        ("h_cand_offsets", c_void_p),
        ("h_cand_ids", c_void_p),
        ("h_labels", c_void_p),
        ("d_hist_offsets", c_void_p),
        ("d_hist_ids", c_void_p),
        ("d_cand_offsets", c_void_p),
        ("d_cand_ids", c_void_p),
        ("d_labels", c_void_p),
        ("h_hist_pad", c_void_p),
        ("d_hist_pad", c_void_p),
        ("h_cand_pad", c_void_p),
        ("d_cand_pad", c_void_p),
        ("ready", c_void_p),
        ("h_marks", c_void_p),
        ("copy_stream", c_void_p),
    ]


class MetricsDesc(Structure):
    """mb200_metrics_desc, field for field."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("k0", c_int32),
        ("k1", c_int32),
        ("max_cand", c_int32),
        ("n_impressions", c_int64),
        ("preds", c_void_p),
        ("labels", c_void_p),
        ("cand_offsets", c_void_p),
        ("cand_category", c_void_p),
        ("cand_sentiment", c_void_p),
        ("hist_offsets", c_void_p),
        ("hist_category", c_void_p),
        ("hist_sentiment", c_void_p),
        ("num_categ_classes", c_int32),
        ("num_sent_classes", c_int32),
        ("per_impression", c_void_p),
        ("sums", c_void_p),
        ("flags", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
    ]


class RetrievalDesc(Structure):
    """mb200_retrieval_desc, field for field."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("dim", c_int32),
        ("k", c_int32),
        ("reserved", c_int32),
        ("n_users", c_int64),
        ("n_catalog", c_int64),
        ("catalog_id_offset", c_int64),
        ("users", c_void_p),
        ("catalog", c_void_p),
        ("out_scores", c_void_p),
        ("out_ids", c_void_p),
        ("debug_scores", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
        ("n_peers", c_int32),
        ("my_rank", c_int32),
        ("peer_rows", c_int64),
        ("peer_scores", c_void_p * MAX_TABLE_SHARDS),
        ("peer_ids", c_void_p * MAX_TABLE_SHARDS),
    ]


class ExchangeDesc(Structure):
    """mb200_exchange_desc, field for field."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("n_ranks", c_int32),
        ("my_rank", c_int32),
        ("epoch", c_uint32),
        ("n_payload", c_int32),
        ("outside_index", c_int32),
        ("pos_capacity", c_int64),
        ("mailbox", c_void_p * MAX_TABLE_SHARDS),
        ("payload", c_void_p),
        ("pos_keys", c_void_p),
        ("n_pos", c_void_p),
        ("sorted_neg", c_void_p),
        ("n_rows", c_int64),
        ("out_payload", c_void_p),
        ("out_stats", c_void_p),
        ("flags", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
    ]


# every symbol include/manner_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "mb200_abi_version": (c_int, []),
    "mb200_status_str": (c_char_p, [c_int]),
    "mb200_last_cuda_error": (c_char_p, []),
    "mb200_eval_workspace_bytes": (c_size_t, [POINTER(EvalDesc)]),
    "mb200_score_eval": (c_int, [POINTER(EvalDesc), c_void_p]),
    "mb200_upload_begin": (c_int, [POINTER(UploadDesc), c_void_p]),
    "mb200_upload_finish": (c_int, [POINTER(UploadDesc)]),
    "mb200_auc_build_keys": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mb200_auc_sort_workspace_bytes": (c_size_t, [c_int64]),
    "mb200_auc_sort_keys": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "mb200_auc_rank_sum": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "mb200_pooled_auc_workspace_bytes": (c_size_t, [c_int64]),
    "mb200_pooled_auc": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "mb200_pooled_auc_bounded_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "mb200_pooled_auc_bounded": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "mb200_retrieval_workspace_bytes": (c_size_t, [POINTER(RetrievalDesc)]),
    "mb200_retrieve_topk": (c_int, [POINTER(RetrievalDesc), c_void_p]),
    "mb200_pool_users": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "mb200_merge_topk": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "mb200_attention_logits": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mb200_step_loss": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mb200_metrics_workspace_bytes": (c_size_t, [POINTER(MetricsDesc)]),
    "mb200_rank_metrics": (c_int, [POINTER(MetricsDesc), c_void_p]),
    "mb200_exchange_mailbox_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "mb200_exchange_workspace_bytes": (c_size_t, [c_int64]),
    "mb200_exchange_post": (c_int, [POINTER(ExchangeDesc), c_void_p]),
    "mb200_exchange_finish": (c_int, [POINTER(ExchangeDesc), c_void_p]),
    "mb200_read_probe": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p]),
    "mb200_enable_peer_access": (c_int, [c_int, c_int]),
    "mb200_ipc_export": (c_int, [c_void_p, c_char_p, POINTER(c_int64)]),
    "mb200_ipc_open": (c_int, [c_char_p, c_int64, c_int, POINTER(c_void_p)]),
    "mb200_ipc_close": (c_int, [c_void_p, c_int]),
    "mb200_dcg_discount": (c_float, [c_int]),
    "mb200_launch_count": (c_int64, []),
    "mb200_library_launch_count": (c_int64, []),
    "mb200_set_tuning": (c_int, [c_int, c_int]),
    "mb200_last_score_kernel_ms": (c_float, []),
    "mb200_last_score_kernel_begin_after": (c_float, [c_void_p]),
    "mb200_last_hot_stats": (c_int, [POINTER(c_int32)]),
}

_lib: Optional[ctypes.CDLL] = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """The loaded shared library.  Raises (loudly) when it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m manner_b200.build` (needs nvcc). "
                "manner_b200 has no CPU or PyTorch fallback for its CUDA path."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            if not hasattr(handle, name):
                raise NativeLibraryMissing(f"{LIB_PATH} does not export {name}; rebuild it")
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        if handle.mb200_abi_version() != ABI_VERSION:
            raise NativeLibraryMissing(f"{LIB_PATH} has ABI {handle.mb200_abi_version()}, python side expects {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


class NativeError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != OK:
        l = lib()
        detail = l.mb200_status_str(status).decode()
        if status == ERR_CUDA:
            detail += ": " + l.mb200_last_cuda_error().decode()
        raise NativeError(f"{what} failed: {detail} (status {status})")
