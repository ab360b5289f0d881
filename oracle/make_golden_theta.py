"""Golden fixture AWAY from the nominal parameters (TEST INFRASTRUCTURE; oracle outputs, not acados outputs): every
sample has its own theta -- what MPC.set_p / set_parameter do between learning steps (rlmpc/mpc/common/mpc.py:137-149,
212-257) and what a batched TD step with per-sample parameters needs.  cartpole_original.yaml; the model parameters
(M, m, l) of sample i are the nominal ones times U(0.6, 1.4) each, the cost parameters stay nominal; V(x0) and Q(x0, a)
to convergence + restated update_nlp at that theta.

    python -m oracle.make_golden_theta [n_samples] [n_procs]      # tests/golden/cartpole_original_theta.npz
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_S = {}


def _one(args):
    i, x0, a, th = args
    import torch

    torch.set_num_threads(1)
    from .problems import make_cartpole
    from .solver import DenseSolver

    if "s" not in _S:
        _S["s"] = DenseSolver(make_cartpole("original"))
    s = _S["s"]
    sol, upd = s.unit(x0, p=th, tol=1e-10)
    solq, updq = s.unit(x0, u0=a, p=th, tol=1e-10)
    print(f"[theta {i}] (M, m, l)={th[:3]} V={sol.cost:.6f} u0={sol.U[0]} st={sol.status} | Q={solq.cost:.6f} st={solq.status}", flush=True)
    return dict(status=np.array([sol.status, solq.status]), V=sol.cost, u0=sol.U[0], dV=upd["dL_dp"][0][:3], dpi=upd["dpi_dp"][:, :3],
                Q=solq.cost, dQ=updq["dL_dp"][0][:3])


def main(n=16, procs=8, seed=2468):
    from .make_golden import sample_states
    from .problems import make_cartpole

    pb = make_cartpole("original")
    x0s, acts = sample_states(n, seed, "original")
    rng = np.random.default_rng(seed + 1)
    th = np.tile(pb.p_nominal, (n, 1))
    th[:, :3] *= rng.uniform(0.6, 1.4, size=(n, 3))
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_one, [(i, x0s[i], acts[i], th[i]) for i in range(n)], chunksize=1)
    out = {k: np.array([r[k] for r in res]) for k in res[0]}
    out.update(x0=x0s, a=acts, theta=th)
    path = os.path.join(ROOT, "tests", "golden", "cartpole_original_theta.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "converged (V, Q):", (out["status"] == 0).sum(0))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 8)
