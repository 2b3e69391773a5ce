"""In-tree build of librl_engine.so (hand-written sm_100a CUDA + the C-ABI).

nvcc cross-compiles without a GPU, so this runs in the CPU-only dev container; the built
.so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "librl_engine.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "128,177",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: limitador_b200 needs the CUDA toolkit to build librl_engine.so")


def sources():
    return [os.path.join(CSRC, f) for f in ("rl_engine.cu", "rl_front.cu", "rl_match.cpp")]


def _deps():
    out = [os.path.join(_ROOT, "include", "rl_engine.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def build_engine(force: bool = False, verbose: bool = False) -> str:
    """Compile librl_engine.so if missing or stale; returns its path."""
    if not force and os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(d) <= t for d in _deps()):
            return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", os.path.join(_ROOT, "include"), "-o", LIB_PATH, *sources()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


def build_variant(name: str, defines) -> str:
    """An A/B build with extra -D flags into limitador_b200/variants/ (load it with RL_ENGINE_LIB=<path>);
    the product library is always the plain build above."""
    out_dir = os.path.join(_PKG, "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"librl_engine_{name}.so")
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", os.path.join(_ROOT, "include"), "-o", out, *sources()]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_engine(force=True, verbose=True))
