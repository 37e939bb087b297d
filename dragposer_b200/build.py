"""Builds the engine's shared libraries in-tree with nvcc for sm_100a.

`libdp_engine.so` (CUDA kernels + the dp_engine_* C ABI, include/dp_engine.h) and
`libDragPoserDLL.so` (the reference's exportFunc.h C ABI on top of it).  The .so
files are git-ignored but travel to the GPU box with the repo snapshot.

Every translation unit is compiled to its own object (in parallel) and keyed by a
SHA-256 of its preprocessing inputs (the source, every header under csrc/ and
include/, the compiler flags): `build_all()` recompiles exactly the objects whose
key changed and relinks when any object did, so a stale or foreign .so that
travelled with a snapshot is never trusted on file times alone.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
ENGINE_SO = os.path.join(HERE, "libdp_engine.so")
DLL_SO = os.path.join(HERE, "libDragPoserDLL.so")
STAMP = os.path.join(OBJ, "stamp.json")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]
DLL_FLAGS = ["-std=c++17", "-O2", "-shared", "-fPIC", "-fvisibility=hidden"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sha(paths, extra=()):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as fh:
            h.update(fh.read())
    for e in extra:
        h.update(str(e).encode())
    return h.hexdigest()


def engine_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "dp_engine.h"))
    return srcs, hdrs


def _load_stamp():
    try:
        with open(STAMP) as fh:
            return json.load(fh)
    except Exception:
        return {}


def _save_stamp(st):
    os.makedirs(OBJ, exist_ok=True)
    with open(STAMP, "w") as fh:
        json.dump(st, fh, indent=1, sort_keys=True)


def _file_sha(path):
    return _sha([path]) if os.path.exists(path) else None


def build_engine(force=False, verbose=False):
    srcs, hdrs = engine_sources()
    os.makedirs(OBJ, exist_ok=True)
    stamp = _load_stamp()
    nvcc = _nvcc()
    jobs, objs = [], []
    for src in srcs:
        name = os.path.basename(src)
        obj = os.path.join(OBJ, name[:-3] + ".o")
        key = _sha([src] + hdrs, NVCC_FLAGS)
        objs.append(obj)
        if force or stamp.get(name) != key or not os.path.exists(obj):
            cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", obj, src]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((name, key, cmd))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            def run(job):
                name, key, cmd = job
                r = subprocess.run(cmd, capture_output=True, text=True)
                return name, key, r
            for name, key, r in pool.map(run, jobs):
                if verbose or r.returncode:
                    sys.stderr.write(f"--- {name}\n{r.stdout}{r.stderr}")
                if r.returncode:
                    stamp.pop(name, None)
                    _save_stamp(stamp)
                    raise RuntimeError(f"nvcc failed on {name}")
                stamp[name] = key
    link_key = _sha(objs, ["link"])
    if force or jobs or stamp.get("libdp_engine.link") != link_key or stamp.get("libdp_engine.so") != _file_sha(ENGINE_SO):
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", ENGINE_SO, *objs], check=True)
        stamp["libdp_engine.link"] = link_key
        stamp["libdp_engine.so"] = _file_sha(ENGINE_SO)
    for gone in [k for k in stamp if k.endswith(".cu") and k not in {os.path.basename(s) for s in srcs}]:
        stamp.pop(gone)
        o = os.path.join(OBJ, gone[:-3] + ".o")
        if os.path.exists(o):
            os.remove(o)
    _save_stamp(stamp)
    return ENGINE_SO


def build_dll(force=False, verbose=False):
    """libDragPoserDLL.so: the reference's exportFunc.h C ABI (host C++ only) linked against libdp_engine.so."""
    build_engine(force=force, verbose=verbose)
    src = os.path.join(HERE, "csrc_dll", "exportFunc.cpp")
    deps = [src, os.path.join(ROOT, "include", "exportFunc.h"), os.path.join(ROOT, "include", "dp_engine.h")]
    stamp = _load_stamp()
    key = _sha(deps, DLL_FLAGS + [stamp.get("libdp_engine.so")])
    if not force and stamp.get("libDragPoserDLL.key") == key and stamp.get("libDragPoserDLL.so") == _file_sha(DLL_SO):
        return DLL_SO
    cmd = ["g++", *DLL_FLAGS, "-o", DLL_SO, src, "-L" + HERE, "-ldp_engine", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    stamp["libDragPoserDLL.key"] = key
    stamp["libDragPoserDLL.so"] = _file_sha(DLL_SO)
    _save_stamp(stamp)
    return DLL_SO


def build_all(force=False, verbose=False):
    return [build_engine(force=force, verbose=verbose), build_dll(force=force, verbose=verbose)]


def build_probes(verbose=False):
    """scripts/probes/libdp_probe.so: the tcgen05 descriptor / rate probes (measurement tools, not part of the product library)."""
    src = os.path.join(ROOT, "scripts", "probes", "dp_selftest.cu")
    out = os.path.join(ROOT, "scripts", "probes", "libdp_probe.so")
    cmd = [_nvcc(), *NVCC_FLAGS, "-I" + CSRC, "-shared", "-o", out, src, "-Xcompiler", "-fvisibility=default"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--probes" in sys.argv:
        print(build_probes(verbose="-v" in sys.argv))
