"""Ahead-of-time build of the C-ABI CUDA library (sm_100a only), in-tree.

    python -m g_adaptivity_b200.build [--force] [--verbose] [--out NAME.so] [-D MACRO[=V] ...]

Produces `g_adaptivity_b200/libgadapt_b200.so` from `g_adaptivity_b200/csrc/*.cu` with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`.  No torch headers are involved: the
library is plain CUDA behind `include/gadapt.h` and is loaded with ctypes (`_lib.py`).  nvcc
cross-compiles without a GPU, so this runs in the build container; the .so travels to the GPU box
with the repo snapshot (it is git-ignored, not gpurun-ignored).

`--out` / `-D` build a kernel variant next to the default library (selected at run time with the
GAD_LIB environment variable; used by scripts/kbench.py to compare variants on the GPU box).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libgadapt_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# no --use_fast_math: expf / logf / division stay accurate unless a kernel asks for an intrinsic
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC]


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


def up_to_date(lib: str = LIB) -> bool:
    return os.path.exists(lib) and os.path.getmtime(lib) >= _deps_mtime()


def _compile(src: str, objdir: str, defines, verbose: bool) -> str:
    base = os.path.basename(src)[:-3]
    obj = os.path.join(objdir, base + ".o")
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(INCLUDE, "gadapt.h")]
    newest = max(os.path.getmtime(p) for p in [src, __file__] + hdrs)
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [nvcc()] + ARCH + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(objdir, base + ".ptxas.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, out: str = LIB, defines=()) -> str:
    defines = list(defines)
    if not os.path.isabs(out):
        out = os.path.join(PKG, out)
    if not force and up_to_date(out):
        return out
    tag = hashlib.md5(" ".join(sorted(defines)).encode()).hexdigest()[:8] if defines else "default"
    objdir = os.path.join(BUILD, tag)
    if force and os.path.isdir(objdir):
        shutil.rmtree(objdir)
    os.makedirs(objdir, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, objdir, defines, verbose), srcs))
    cmd = [nvcc()] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", out] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--out", default=LIB)
    ap.add_argument("-D", dest="defines", action="append", default=[])
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, out=a.out, defines=a.defines))
