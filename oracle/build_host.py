"""Build the host (CPU) harnesses of oracle/ with g++: `python -m oracle.build_host`.  TEST INFRASTRUCTURE ONLY.
Output goes to oracle/_build/ (git-ignored; it travels to the GPU box like the other built libraries)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libfem2d_host.so")
SRC = os.path.join(HERE, "fem2d_host.cpp")
DEP = os.path.join(os.path.dirname(HERE), "g_adaptivity_b200", "csrc", "fem2d_math.cuh")


EMU_LIB = os.path.join(OUT, "libfem2d_emu.so")
EMU_SRC = os.path.join(HERE, "fem2d_emu.cpp")
EMU_DEPS = [EMU_SRC, os.path.join(HERE, "cuda_emu.h"), DEP, os.path.join(os.path.dirname(DEP), "fem2d.cu")]


def build_emu(force: bool = False) -> str:
    """csrc/fem2d.cu compiled for the CPU (kernels on std::threads, oracle/cuda_emu.h)."""
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(EMU_LIB) and os.path.getmtime(EMU_LIB) >= max(os.path.getmtime(p) for p in EMU_DEPS):
        return EMU_LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found: the fem2d kernel emulation cannot be built")
    cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-x", "c++", EMU_SRC, "-o", EMU_LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return EMU_LIB


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(DEP)):
        return LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found: the fem2d host harness cannot be built")
    # -ffp-contract=off: the edge tests and barycentric values must round every product and sum separately
    cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
    print(build_emu(force=True))
