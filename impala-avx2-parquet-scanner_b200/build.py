"""Builds libfle_b200.so (nvcc, sm_100a only) in-tree, next to this file.

The .so is git-ignored but travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfle_b200.so")
SOURCES = ["fle_api.cu", "fle_eval.cu", "fle_eval1.cu", "fle_eval_in.cu", "fle_codec.cu", "fle_bitmap.cu", "fle_scan.cu", "fle_plain.cu", "fle_dict.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


OBJ_DIR = os.path.join(CSRC, "_obj")
HEADERS = ["fle_device.cuh", "fle_internal.h", "fle_eval_common.cuh"]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Incremental: one object per .cu (rebuilt when it or a shared header changed), then link."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = max([_mtime(os.path.join(CSRC, h)) for h in HEADERS] +
                   [_mtime(os.path.join(HERE, "..", "include", "fle_b200.h")), _mtime(__file__)])
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _mtime(obj) < max(_mtime(os.path.join(CSRC, src)), hdr_time):
            jobs.append([nvcc] + compile_flags + ["-c", "-o", obj, os.path.join(CSRC, src)])

    def run(cmd):
        print("[fle_b200] " + " ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2)) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB) or any(_mtime(o) > _mtime(LIB) for o in objs):
        run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
