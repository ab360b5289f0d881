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
    out = [os.path.join(CSRC, "rlmpc_b200.cu"), os.path.join(CSRC, "rlmpc_chain.cu")]
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    deps += [os.path.join(CSRC, "chain", f) for f in sorted(os.listdir(os.path.join(CSRC, "chain")))]
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
    """Every translation unit is compiled to an object file under csrc/_build (in parallel, only the stale ones),
    then linked into the shared library."""
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs, deps = _sources()
    bdir = os.path.join(CSRC, "_build")
    os.makedirs(bdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    newest_dep = max(os.path.getmtime(f) for f in deps)
    procs, objs = [], []
    for src in srcs:
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_dep):
            continue
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = ""
    for cmd, p in procs:
        out, _ = p.communicate()
        log += out
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(log)
    return LIB
