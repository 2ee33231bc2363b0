"""Build the native libraries IN-TREE for sm_100a (nvcc cross-compiles without a GPU).

  libb200sdr.so         kernels + the C-ABI (include/gsdr/*.h, include/b200sdr/b200sdr.h)
  libgpusdrpipeline.so  C++ host framework mirroring the reference interface (getFactoriesSingleton, ...)

The built .so files are git-ignored but travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(ROOT, "build", "obj")

LIB_KERNELS = os.path.join(PKG, "libb200sdr.so")
LIB_HOST = os.path.join(PKG, "libgpusdrpipeline.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = "/usr/bin/g++"  # the image's $CXX is a trimmed wrapper; use the system compiler
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-ccbin", CXX,
                     "-I" + INCLUDE, "-I" + CSRC] + os.environ.get("NVCC_EXTRA", "").split()


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        raise RuntimeError("native build failed: " + os.path.basename(cmd[-1]))


def _headers(*dirs: str) -> list[str]:
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files if f.endswith((".h", ".cuh", ".hpp"))]
    return out


def build_kernels(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = _headers(CSRC, INCLUDE)
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        if force or _newer(obj, [src] + deps):
            jobs.append((src, obj))
    if verbose and jobs:
        print("nvcc:", ", ".join(os.path.basename(s) for s, _ in jobs), flush=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(lambda j: _run([NVCC] + NVCC_FLAGS + ["-c", j[0], "-o", j[1]]), jobs))
    objs = [os.path.join(OBJ, os.path.basename(s) + ".o") for s in sources]
    if force or jobs or _newer(LIB_KERNELS, objs):
        _run([NVCC] + ARCH + ["-shared", "-ccbin", CXX, "-o", LIB_KERNELS] + objs)
    return LIB_KERNELS


def build_host(force: bool = False, verbose: bool = False) -> str | None:
    if not os.path.isdir(HOST):
        return None
    sources = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    if not sources:
        return None
    os.makedirs(OBJ, exist_ok=True)
    deps = _headers(HOST, INCLUDE)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    flags = ["-O2", "-std=c++20", "-fPIC", "-fvisibility=hidden", "-Wall", "-I" + INCLUDE, "-I" + HOST,
             "-I" + os.path.join(cuda_home, "include")]
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, "host_" + os.path.basename(src) + ".o")
        if force or _newer(obj, [src] + deps):
            jobs.append((src, obj))
    if verbose and jobs:
        print("g++:", ", ".join(os.path.basename(s) for s, _ in jobs), flush=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(lambda j: _run([CXX] + flags + ["-c", j[0], "-o", j[1]]), jobs))
    objs = [os.path.join(OBJ, "host_" + os.path.basename(s) + ".o") for s in sources]
    if force or jobs or _newer(LIB_HOST, objs + [LIB_KERNELS]):
        _run([CXX, "-shared", "-o", LIB_HOST] + objs +
             ["-L" + PKG, "-lb200sdr", "-Wl,-rpath,$ORIGIN", "-L" + os.path.join(cuda_home, "lib64"), "-lcudart_static",
              "-ldl", "-lrt", "-lpthread"])
    return LIB_HOST


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_kernels(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built", LIB_KERNELS, "and", LIB_HOST if os.path.exists(LIB_HOST) else "(no host library yet)")
