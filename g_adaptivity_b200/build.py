"""Ahead-of-time build of the C-ABI CUDA library (sm_100a only), in-tree.

    python -m g_adaptivity_b200.build [--force] [--verbose]

Produces `g_adaptivity_b200/libgadapt_b200.so` from `g_adaptivity_b200/csrc/*.cu` with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`.  No torch headers are involved: the
library is plain CUDA behind `include/gadapt.h` and is loaded with ctypes (`_lib.py`).  nvcc
cross-compiles without a GPU, so this runs in the build container; the .so travels to the GPU box
with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libgadapt_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC]
# --use_fast_math is NOT enabled: expf/logf/division stay IEEE-accurate (parity bar 1e-5).
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the gadapt CUDA library cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "gadapt.h"), __file__]
    return max(os.path.getmtime(p) for p in paths)


def up_to_date() -> bool:
    return os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc()] + ARCH + NVCC_FLAGS + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(BUILD, os.path.basename(src)[:-3] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    cmd = [nvcc()] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
