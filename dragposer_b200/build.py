"""Builds the engine's shared libraries in-tree with nvcc for sm_100a.

`libdp_engine.so` (CUDA kernels + the dp_engine_* C ABI, include/dp_engine.h) and
`libDragPoserDLL.so` (the reference's exportFunc.h C ABI on top of it).  The .so
files are git-ignored but travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ENGINE_SO = os.path.join(HERE, "libdp_engine.so")
DLL_SO = os.path.join(HERE, "libDragPoserDLL.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def engine_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dp_engine.h")]
    return srcs, deps


def build_engine(force=False, verbose=False):
    srcs, deps = engine_sources()
    if not force and not _stale(ENGINE_SO, deps):
        return ENGINE_SO
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [_nvcc(), *flags, "-shared", "-o", ENGINE_SO, *srcs, "-Xcompiler", "-fvisibility=default"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return ENGINE_SO


def build_dll(force=False, verbose=False):
    """libDragPoserDLL.so: the reference's exportFunc.h C ABI (host C++ only) linked against libdp_engine.so."""
    build_engine(force=force, verbose=verbose)
    src = os.path.join(HERE, "csrc_dll", "exportFunc.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "exportFunc.h"), os.path.join(HERE, "..", "include", "dp_engine.h"), ENGINE_SO]
    if not force and not _stale(DLL_SO, deps):
        return DLL_SO
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-fvisibility=hidden", "-o", DLL_SO, src, "-L" + HERE, "-ldp_engine",
           "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return DLL_SO


def build_all(force=False, verbose=False):
    return [build_engine(force=force, verbose=verbose), build_dll(force=force, verbose=verbose)]


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
