"""Large golden sets (TEST INFRASTRUCTURE): >= 256 converged samples per cart-pole / linear-system configuration and 32 for
the evaporation process, produced by the dense oracle in a process pool.  Same caveat as oracle/make_golden.py:
outputs of the restatement, not of the reference ("parity unpinned").  Only what the parity tests compare is stored
(V, u0, dV/dtheta, dpi/dtheta, Q, dQ/dtheta, status; no trajectories), so the fixtures stay small.

    python -m oracle.make_golden_large cartpole_original 256 [workers]     # about 12 minutes with 4 workers
    python -m oracle.make_golden_large cartpole_default 64
    python -m oracle.make_golden_large linear_system 256
    python -m oracle.make_golden_large evaporation 32
    python -m oracle.make_golden_large evaporation_n100 8        # full horizon, about 40 minutes with 4 workers
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_S = {}


def _problem(name):
    from .problems import make_cartpole, make_evaporation, make_linear_system

    if name == "cartpole_original":
        return make_cartpole("original")
    if name == "cartpole_default":
        return make_cartpole("default")
    if name == "linear_system":
        return make_linear_system(gamma=0.9)
    if name == "evaporation":
        return make_evaporation(gamma=0.95, N=40)
    if name == "evaporation_n100":  # the reference's full horizon: ~10 minutes of dense oracle per solve
        return make_evaporation(gamma=0.95, N=100)
    raise ValueError(name)


def _states(name, n, seed=4321):
    from .make_golden import sample_states

    rng = np.random.default_rng(seed)
    if name == "cartpole_original":
        return sample_states(n, seed, "original")
    if name == "cartpole_default":
        return sample_states(n, seed, "default")
    if name == "linear_system":
        return rng.uniform([0.0, -1.0], [1.0, 1.0], size=(n, 2)), rng.uniform(-1.0, 1.0, size=(n, 1))
    if name == "evaporation_n100":
        rng = np.random.default_rng(seed + 1)
    x = rng.uniform([25.0, 49.7], [40.0, 70.0], size=(n, 2))
    return x, np.column_stack([rng.uniform(150.0, 350.0, size=(n, 2)), np.full(n, 0.5)])


def _one(args):
    name, i, x0, a = args
    import torch

    torch.set_num_threads(1)
    from .solver import DenseSolver

    if name not in _S:
        _S[name] = DenseSolver(_problem(name))
    s = _S[name]
    pb = s.pb if hasattr(s, "pb") else _problem(name)
    nth, nu = len(pb.p_nominal), pb.nu
    res = []
    for u0 in (None, a):
        try:  # a failed oracle solve is recorded (status -1), never repaired
            sol, upd = s.unit(x0, u0=u0, tol=1e-10 if not name.startswith("evaporation") else 1e-9)
            sl = max(float(np.max(sol.slbx, initial=0.0)), float(np.max(sol.subx, initial=0.0))) if hasattr(sol, "slbx") else 0.0
            res.append((sol.status, sol.cost, np.array(sol.U[0]), upd["dL_dp"][0], upd["dpi_dp"], sl))
        except Exception as e:  # noqa: BLE001
            print(f"[{name} {i}] oracle failed: {type(e).__name__}: {e}", flush=True)
            res.append((-1, np.nan, np.full(nu, np.nan), np.full(nth, np.nan), np.full((nu, nth), np.nan), 0.0))
    (sv, V, u, dV, dpi, sl), (sq, Q, _, dQ, _, slq) = res
    print(f"[{name} {i}] V={V:.6f} st={sv} | Q={Q:.6f} st={sq}", flush=True)
    return i, sv, sq, V, u, dV, dpi, Q, dQ, max(sl, slq)


def main(name, n, workers):
    import multiprocessing as mp

    pb = _problem(name)
    x0s, acts = _states(name, n)
    with mp.get_context("fork").Pool(workers) as pool:
        out = pool.map(_one, [(name, i, x0s[i], acts[i]) for i in range(n)], chunksize=1)
    out.sort(key=lambda r: r[0])
    # the gradient columns that are structurally zero (cost parameters of a model-parameter-only problem) are dropped
    dV = np.array([r[5] for r in out]); dpi = np.array([r[6] for r in out]); dQ = np.array([r[8] for r in out])
    live = np.where((np.nan_to_num(np.abs(dV)).max(0) > 0) | (np.nan_to_num(np.abs(dQ)).max(0) > 0)
                    | (np.nan_to_num(np.abs(dpi)).max((0, 1)) > 0))[0]
    path = os.path.join(ROOT, "tests", "golden", f"{name}_{n}.npz")
    np.savez_compressed(path, x0=x0s, a=acts, theta=pb.p_nominal, status=np.array([[r[1], r[2]] for r in out]),
                        V=np.array([r[3] for r in out]), u0=np.array([r[4] for r in out]), Q=np.array([r[7] for r in out]),
                        cols=live, dV=dV[:, live], dpi=dpi[:, :, live], dQ=dQ[:, live],
                        slmax=np.array([r[9] for r in out]))  # largest slack of a softened bound (0: none active)
    print("wrote", path, "live gradient columns:", len(live))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 4)
