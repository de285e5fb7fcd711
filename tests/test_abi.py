"""CPU: the C-ABI shared library loads without a GPU and exports exactly what include/manner_b200.h
declares; host-only entry points behave; the product refuses to compute without CUDA."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from manner_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "manner_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MB200_API[^;(]*?\b(mb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nat.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(nat.SIGNATURES) == declared, "python binding and header disagree on the symbol list"
    exported = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True).stdout
    public = sorted(set(re.findall(r" T (mb200_\w+)", exported)))
    assert public == declared, "the library exports symbols the header does not declare (or the reverse)"


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "manner_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu",'
                   "sizeof(mb200_eval_desc),offsetof(mb200_eval_desc,tables),offsetof(mb200_eval_desc,weights),"
                   "offsetof(mb200_eval_desc,scores),offsetof(mb200_eval_desc,workspace_bytes));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)  # the header is plain C
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    E = nat.EvalDesc
    assert got == [ctypes.sizeof(E), E.tables.offset, E.weights.offset, E.scores.offset, E.workspace_bytes.offset]


def test_upload_desc_layout_and_validation_without_a_gpu(tmp_path):
    """mb200_upload_desc (pipelined host -> device upload) and the `ready` tail of mb200_eval_desc."""
    src = tmp_path / "su.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "manner_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %d",'
                   "sizeof(mb200_upload_desc),offsetof(mb200_upload_desc,n_impressions),offsetof(mb200_upload_desc,d_hist_offsets),"
                   "offsetof(mb200_upload_desc,ready),offsetof(mb200_upload_desc,copy_stream),offsetof(mb200_eval_desc,ready),"
                   "MB200_PAYLOAD_TAIL);return 0;}\n")
    exe = tmp_path / "su"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    U = nat.UploadDesc
    assert got == [ctypes.sizeof(U), U.n_impressions.offset, U.d_hist_offsets.offset, U.ready.offset, U.copy_stream.offset,
                   nat.EvalDesc.ready.offset, nat.PAYLOAD_TAIL]
    assert nat.PAYLOAD_TAIL == 1 + len(nat.PAYLOAD_FLAG_BITS)
    lib = nat.lib()
    d = U()
    assert lib.mb200_upload_begin(ctypes.byref(d), None) == nat.ERR_INVALID_ARG  # struct_size unset
    assert lib.mb200_upload_finish(None) == nat.ERR_INVALID_ARG
    d.struct_size, d.n_segments, d.segments_first, d.n_impressions = ctypes.sizeof(U), 99, 1, 10
    assert lib.mb200_upload_begin(ctypes.byref(d), None) == nat.ERR_INVALID_ARG  # more than MB200_MAX_UPLOAD_SEGMENTS


def test_host_only_entry_points():
    lib = nat.lib()
    assert lib.mb200_abi_version() == nat.ABI_VERSION
    assert lib.mb200_status_str(nat.OK) == b"ok" and b"workspace" in lib.mb200_status_str(nat.ERR_WORKSPACE)
    assert lib.mb200_pooled_auc_workspace_bytes(1000) >= 3 * 4000
    assert lib.mb200_launch_count() == 0  # nothing has been launched in a CPU-only process
    # the DCG discount table the kernels use == what torchmetrics' _dcg computes on the reference's CPU path
    want = (torch.ones(nat.MAX_K) / torch.log2(torch.arange(nat.MAX_K) + 2.0)).numpy()
    got = np.array([lib.mb200_dcg_discount(r) for r in range(1, nat.MAX_K + 1)], dtype=np.float32)
    np.testing.assert_array_equal(got, want)
    assert lib.mb200_dcg_discount(0) == 0.0 and lib.mb200_dcg_discount(nat.MAX_K + 1) == 0.0


def test_descriptor_validation_without_a_gpu():
    lib = nat.lib()
    d = nat.EvalDesc()
    assert lib.mb200_eval_workspace_bytes(ctypes.byref(d)) == 0  # struct_size unset -> invalid
    assert lib.mb200_score_eval(ctypes.byref(d), None) == nat.ERR_INVALID_ARG
    assert lib.mb200_score_eval(None, None) == nat.ERR_INVALID_ARG
    assert lib.mb200_pooled_auc(None, None, -1, 0, None, None, 0, None, None) == nat.ERR_INVALID_ARG


def test_product_refuses_to_run_without_cuda():
    """No CPU fallback anywhere: CPU tensors raise, and the evaluator needs a device."""
    import manner_b200.ops  # noqa: F401  registers the custom ops

    with pytest.raises(RuntimeError, match="no CPU path"):
        torch.ops.manner_b200.pooled_auc(torch.rand(4), torch.zeros(4, dtype=torch.uint8), 0, None)
    t = torch.zeros(8, 128)
    i32 = torch.zeros(2, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        torch.ops.manner_b200.score_eval([t], i32, i32, i32, i32, torch.zeros(2, dtype=torch.uint8), None, False, 4, 1, 5, 10,
                                         False, 0, False, None, None, 19, 4, [])
    if not torch.cuda.is_available():
        from manner_b200.evaluator import ScoreEvaluator

        with pytest.raises(RuntimeError, match="CUDA"):
            ScoreEvaluator([t])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "manner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(nat.NativeLibraryMissing, match="no CPU or PyTorch fallback"):
        nat.lib()


def test_retrieval_desc_layout_and_validation_without_a_gpu(tmp_path):
    """mb200_retrieval_desc: the ctypes mirror matches the C struct, and the entry points reject bad
    descriptors on the host (no GPU needed to get a status code back)."""
    src = tmp_path / "rsz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "manner_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu",'
                   "sizeof(mb200_retrieval_desc),offsetof(mb200_retrieval_desc,n_users),offsetof(mb200_retrieval_desc,users),"
                   "offsetof(mb200_retrieval_desc,debug_scores),offsetof(mb200_retrieval_desc,workspace_bytes));return 0;}\n")
    exe = tmp_path / "rsz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    R = nat.RetrievalDesc
    assert got == [ctypes.sizeof(R), R.n_users.offset, R.users.offset, R.debug_scores.offset, R.workspace_bytes.offset]

    lib = nat.lib()
    d = R()
    assert lib.mb200_retrieval_workspace_bytes(ctypes.byref(d)) == 0  # struct_size unset
    assert lib.mb200_retrieve_topk(ctypes.byref(d), None) == nat.ERR_INVALID_ARG
    assert lib.mb200_retrieve_topk(None, None) == nat.ERR_INVALID_ARG
    d.struct_size, d.n_users, d.n_catalog, d.dim, d.k = ctypes.sizeof(R), 1000, 5000, 768, 100
    ws = lib.mb200_retrieval_workspace_bytes(ctypes.byref(d))
    assert ws >= 8 * 128 * 256 * 8  # 8 user tiles x 128 rows x 256 candidate slots x (score + id)
    assert lib.mb200_retrieve_topk(ctypes.byref(d), None) == nat.ERR_INVALID_ARG  # null pointers
    assert lib.mb200_merge_topk(None, None, 2, 4, 4, None, None, None) == nat.ERR_INVALID_ARG
    assert lib.mb200_pool_users(None, nat.F32, 768, 768, 10, None, None, 4, None, None, None) == nat.ERR_INVALID_ARG


def test_retrieval_refuses_cpu_tensors():
    import manner_b200.retrieval as rt

    u, c = torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(9, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        torch.ops.manner_b200.retrieve_topk(u, c, 4, 0, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rt.pool_users(torch.zeros(4, 64), torch.zeros(2, dtype=torch.int32), torch.zeros(1, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        rt.merge_topk(torch.zeros(2, 3, 4), torch.zeros(2, 3, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        rt.CatalogRetriever(c, k=4)
    assert rt.catalog_shard_bounds(1000, 3) == [(0, 512), (512, 768), (768, 1000)]
    assert rt.catalog_shard_bounds(100, 4) == [(0, 100), (100, 100), (100, 100), (100, 100)]
    assert rt.catalog_shard_bounds(10_000_000, 8)[-1][1] == 10_000_000


def test_metrics_desc_layout_and_validation_without_a_gpu(tmp_path):
    src = tmp_path / "msz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "manner_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu",'
                   "sizeof(mb200_metrics_desc),offsetof(mb200_metrics_desc,preds),offsetof(mb200_metrics_desc,hist_offsets),"
                   "offsetof(mb200_metrics_desc,num_categ_classes),offsetof(mb200_metrics_desc,workspace_bytes));return 0;}\n")
    exe = tmp_path / "msz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    M = nat.MetricsDesc
    assert got == [ctypes.sizeof(M), M.preds.offset, M.hist_offsets.offset, M.num_categ_classes.offset, M.workspace_bytes.offset]
    lib = nat.lib()
    d = M()
    assert lib.mb200_metrics_workspace_bytes(ctypes.byref(d)) == 0 and lib.mb200_rank_metrics(ctypes.byref(d), None) == nat.ERR_INVALID_ARG
    assert lib.mb200_rank_metrics(None, None) == nat.ERR_INVALID_ARG
    from manner_b200 import metrics

    with pytest.raises(RuntimeError, match="no CPU path"):
        metrics.rank_metrics(torch.rand(4), torch.zeros(4, dtype=torch.uint8), torch.tensor([0, 4], dtype=torch.int32), 4)
    # the host-side grouping mirrors torchmetrics: sort by index, group sizes in ascending index order
    off, perm, mx = metrics._offsets_from_indexes(torch.tensor([5, 5, 2, 9, 2, 2]))
    assert off.tolist() == [0, 3, 5, 6] and perm.tolist() == [2, 4, 5, 0, 1, 3] and mx == 3
    off, perm, mx = metrics._offsets_from_indexes(torch.tensor([0, 0, 1, 3, 3]))
    assert off.tolist() == [0, 2, 3, 5] and perm is None and mx == 2


def test_ipc_and_sharded_table_entry_points_reject_bad_arguments_without_a_gpu():
    lib = nat.lib()
    buf = ctypes.create_string_buffer(64)
    off = ctypes.c_int64(0)
    out = ctypes.c_void_p()
    assert lib.mb200_ipc_export(None, buf, ctypes.byref(off)) == nat.ERR_INVALID_ARG
    assert lib.mb200_ipc_open(None, 0, 0, ctypes.byref(out)) == nat.ERR_INVALID_ARG
    assert lib.mb200_ipc_open(buf.raw, -1, 0, ctypes.byref(out)) == nat.ERR_INVALID_ARG
    # descriptor validation of row-sharded tables is host-side
    d = nat.EvalDesc()
    d.struct_size = ctypes.sizeof(nat.EvalDesc)
    d.n_modules, d.dtype, d.dim, d.active_modules_mask, d.n_news, d.row_stride = 1, nat.F32, 768, 1, 5000, 768
    d.n_impressions, d.max_cand, d.n_weightings, d.k0, d.k1 = 4, 10, 1, 5, 10
    for name in ("hist_offsets", "hist_ids", "cand_offsets", "cand_ids", "labels", "sums"):
        setattr(d, name, 0x1000)
    d.n_table_shards, d.table_shard_shift = 2, 12
    assert lib.mb200_eval_workspace_bytes(ctypes.byref(d)) == 0  # shard pointers missing
    d.table_shards[0][0], d.table_shards[0][1] = 0x10000, 0x20000
    assert lib.mb200_eval_workspace_bytes(ctypes.byref(d)) > 0
    d.table_shard_shift = 11  # 2 x 2048 rows < 5000 news
    assert lib.mb200_eval_workspace_bytes(ctypes.byref(d)) == 0
    d.table_shard_shift, d.n_table_shards = 12, 9
    assert lib.mb200_eval_workspace_bytes(ctypes.byref(d)) == 0
