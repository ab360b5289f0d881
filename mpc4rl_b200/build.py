"""Compile the CUDA library in-tree: nvcc -> mpc4rl_b200/librlmpc_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.path.join(PKG_DIR, "librlmpc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources():
    out = [os.path.join(CSRC, "rlmpc_b200.cu")]
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    deps += [os.path.join(CSRC, "models", f) for f in sorted(os.listdir(os.path.join(CSRC, "models")))]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "rlmpc_b200.h"))
    return out, deps


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    srcs, deps = _sources()
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in srcs + deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs, _ = _sources()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB
