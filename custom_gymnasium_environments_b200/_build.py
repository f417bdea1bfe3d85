"""Build libbeng.so (the sm_100a CUDA engine) in-tree with nvcc.

The library is built IN the package directory so that it travels to the GPU box with the repo
snapshot; it is git-ignored (*.so).  nvcc cross-compiles for sm_100a without a GPU.
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libbeng.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-DBENG_ARCH=100", "--expt-relaxed-constexpr", "--fmad=false",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))


HASH_PATH = LIB_PATH + ".srchash"


def source_hash() -> str:
    """sha256 over the compile flags and every source the library is built from (content, not mtimes: the snapshot that
    carries the library to the GPU box does not preserve them)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sorted(_deps()):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_stale() -> bool:
    """True when libbeng.so is missing or was built from different sources / flags than the ones in the tree."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libbeng.so")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH + ".tmp", *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash() + "\n")
    if verbose:
        print(res.stderr)
    return LIB_PATH
