"""Builds manner_b200/lib/libmanner_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m manner_b200.build [--force] [--verbose]

One object per .cu (so touching the fused kernel does not recompile CUB), linked into one shared
library with the CUDA runtime linked statically -- the .so has no dependency on torch.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmanner_b200.so")
SOURCES = ["api.cu", "score_eval.cu", "pooled_auc.cu", "retrieval.cu", "attention.cu", "upload.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(HERE), "include", "manner_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmanner_b200.so cannot be built")


def _stale(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    jobs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
            jobs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, cmd, proc in jobs:  # the translation units compile concurrently
        out, _ = proc.communicate()
        log = os.path.join(BUILD, src + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + out)
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} (see {log})")
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-Xlinker", "--no-undefined"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libmanner_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
