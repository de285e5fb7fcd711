"""Multi-process checks of everything that crosses GPUs, run under torchrun from pytest (one rank per GPU, NCCL only for the
rendezvous): a behaviour set sharded over the ranks gives the single-GPU numbers through the fused exchange
(mb200_exchange_post / _finish: stores into the peers' mailboxes over NVLink) and through the NCCL fallback; row-sharded
tables are bit-identical to replicated ones; retrieval's row-sharded catalogue returns the single-GPU lists.

World size 1 runs on any GPU box (the exchange kernels with one rank); world size 2 needs two devices and is additionally
marked ``multigpu`` (skipped where there is one)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script: str, world: int, port: int, timeout: int = 900) -> dict:
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    env = dict(os.environ, NCCL_DEBUG=os.environ.get("NCCL_DEBUG", "WARN"))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert lines, res.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("world", [1, pytest.param(2, marks=pytest.mark.multigpu)])
def test_sharded_evaluation_equals_single_gpu(world):
    out = _torchrun("dist_eval_check.py", world, 29541 + world)
    assert out["world"] == world
    for case in ("ensemble", "early_fusion_supcon", "late_fusion_ce"):
        for exchange in ("p2p", "nccl"):
            assert out[f"{case}_{exchange}"] is True, (case, exchange, out)


@pytest.mark.multigpu
def test_row_sharded_tables_bit_identical_to_replicated():
    out = _torchrun("dist_sharded_table_check.py", 2, 29551)
    assert out["small_bit_identical"] is True and out["mind_small_shape_bit_identical"] is True, out


@pytest.mark.multigpu
def test_row_sharded_catalogue_retrieval_equals_single_gpu():
    out = _torchrun("dist_retrieval_check.py", 2, 29511)
    assert out["all_gather"] and out["all_to_all"] and out["p2p"], out
