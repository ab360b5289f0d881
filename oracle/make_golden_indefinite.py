"""Fixture for sensitivities at iterates whose EXACT Lagrangian Hessian is indefinite (TEST INFRASTRUCTURE; oracle outputs).

update_nlp solves the KKT system with a general sparse LU (rlmpc/mpc/nlp.py:1413-1424): it needs the system
nonsingular, not the reduced Hessian positive definite.  In closed loop, while the pole falls through the lower half
plane, the RTI iterate is far from a minimiser and the exact Hessian is indefinite on a large part of the batch (40 % of
the environments around step 48 of the swing-up).  Up to round 2 the engine's Riccati factorisation rejected those
samples (status 4); now it accepts pivots of either sign.

States: 64 environments from the hanging position, 48 closed-loop steps of the host build of the engine with
exploration noise; per sample the primal iterate (U, X) BEFORE step 49, the state, and the dense oracle's result of ONE
SQP step from that iterate (QP at the tau-central point) followed by the restated update_nlp.  `indefinite` marks the
samples on which a positive-definiteness test of the Riccati pivots fails (host build with -DRLMPC_SENS_REQUIRE_PD).

    python -m oracle.make_golden_indefinite [n_keep]      # tests/golden/cartpole_original_indefinite.npz
"""
from __future__ import annotations

import ctypes as C
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle(args):
    i, x, U, X = args
    import torch

    torch.set_num_threads(1)
    from .problems import make_cartpole
    from .solver import DenseSolver

    r, ru = DenseSolver(make_cartpole("original")).unit(x, init=(U, X), max_iter=1, polish=False)
    print(f"[indefinite {i}] u0'={r.U[0, 0]:.6f} kkt'={r.kkt:.2e}", flush=True)
    return dict(u1=r.U[0], V1=r.cost, dV1=ru["dL_dp"][0][:3], dpi1=ru["dpi_dp"][:, :3])


def main(n_keep=32, n_env=64, steps=48, seed=0):
    sys.path.insert(0, ROOT)
    from bench import env_step_np
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec
    from oracle import cpu_port as cp

    spec = cartpole_spec(cartpole_original_config())
    N = spec.N
    pd = cp.make_pd(N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-6, warm_ipm=1)
    rng = np.random.default_rng(seed)
    x = np.array([0.0, 0.0, np.pi, 0.0]) + rng.uniform(-0.05, 0.05, size=(n_env, 4))
    o = cp.unit(1, pd, 0, 30, spec.p_nominal, x, do_sens=False)
    it = o["iterate"]
    for _ in range(steps):
        o = cp.unit(1, pd, 0, 1, spec.p_nominal, x, iterate=it, do_sens=False)
        it = o["iterate"]
        u = o["u0"][:, 0] + 0.05 * 80.0 * rng.standard_normal(n_env)
        x = env_step_np(x, np.clip(u, -80.0, 80.0))
    before = it.copy()
    # label: does a positive-definiteness test of the pivots fail?  (variant build of the host port)
    with tempfile.TemporaryDirectory() as td:
        so = os.path.join(td, "libcpu_port_pd.so")
        subprocess.run(["g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-pthread", "-fPIC", "-x", "c++", "-DRLMPC_SENS_REQUIRE_PD",
                        "-shared", "-o", so, os.path.join(ROOT, "oracle", "cpu_port", "cpu_port.cpp")], check=True)
        keep = cp._lib
        cp._lib = C.CDLL(so)
        flag = cp.unit(1, pd, 0, 1, spec.p_nominal, x, iterate=before.copy(), do_sens=True)["status"] == 4
        cp._lib = keep
    print("indefinite on", int(flag.sum()), "of", n_env)
    sel = np.concatenate([np.where(flag)[0][: n_keep // 2], np.where(~flag)[0][: n_keep - n_keep // 2]])
    Xs = np.stack([before[:(N + 1) * 4, i].reshape(N + 1, 4) for i in sel])
    Us = np.stack([before[(N + 1) * 4:(N + 1) * 4 + N, i].reshape(N, 1) for i in sel])
    with mp.get_context("spawn").Pool(4) as pool:
        res = pool.map(_oracle, [(j, x[i], Us[j], Xs[j]) for j, i in enumerate(sel)], chunksize=1)
    out = {k: np.array([r[k] for r in res]) for k in res[0]}
    path = os.path.join(ROOT, "tests", "golden", "cartpole_original_indefinite.npz")
    np.savez_compressed(path, x1=x[sel], U=Us, X=Xs, indefinite=flag[sel], theta=spec.p_nominal, **out)
    print("wrote", path)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32)
