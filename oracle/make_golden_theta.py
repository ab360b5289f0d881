"""Golden fixture AWAY from the nominal parameters (TEST INFRASTRUCTURE; oracle outputs, not acados outputs): every
sample has its own theta -- what MPC.set_p / set_parameter do between learning steps (rlmpc/mpc/common/mpc.py:137-149,
212-257) and what a batched TD step with per-sample parameters needs.  cartpole_original.yaml; the model parameters
(M, m, l) of sample i are the nominal ones times U(0.6, 1.4) each, the cost parameters stay nominal; V(x0) and Q(x0, a)
to convergence + restated update_nlp at that theta.

Also: the linear system with its 12 learnable parameters (A, B, b, f, V_0) moved by N(0, 0.03^2) each -- what the
Q-learning example does to them (examples/linear_system_mpc_qlearning.py:203-205) -- and the evaporation process (N = 40)
with tracking weights W_0, W scaled as D W D, D = diag(U(0.8, 1.2)), and references yref_0, yref times U(0.97, 1.03)
(scripts/evaporation_process_mpc_qlearning.py updates exactly these).

    python -m oracle.make_golden_theta [n_samples] [n_procs] [cartpole_original|linear_system|evaporation]
                                                                  # tests/golden/<name>_theta.npz
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_S = {}


def _one(args):
    name, i, x0, a, th = args
    import torch

    torch.set_num_threads(1)
    from .make_golden_large import _problem
    from .solver import DenseSolver

    if name not in _S:
        _S[name] = DenseSolver(_problem(name))
    s = _S[name]
    ng = 3 if name == "cartpole_original" else len(th)  # gradient columns kept (cart-pole: the model parameters)
    tol = 1e-9 if name == "evaporation" else 1e-10
    try:  # a failed oracle solve is recorded (status -1), never repaired
        sol, upd = s.unit(x0, p=th, tol=tol)
        solq, updq = s.unit(x0, u0=a, p=th, tol=tol)
    except Exception as e:  # noqa: BLE001
        print(f"[{name} theta {i}] oracle failed: {type(e).__name__}: {e}", flush=True)
        nu = s.pb.nu
        return dict(status=np.array([-1, -1]), V=np.nan, u0=np.full(nu, np.nan), dV=np.full(ng, np.nan), dpi=np.full((nu, ng), np.nan),
                    Q=np.nan, dQ=np.full(ng, np.nan), slmax=0.0)
    sl = max(float(np.max(sol.slbx, initial=0.0)), float(np.max(sol.subx, initial=0.0))) if hasattr(sol, "slbx") else 0.0
    print(f"[{name} theta {i}] V={sol.cost:.6f} u0={sol.U[0]} st={sol.status} | Q={solq.cost:.6f} st={solq.status}", flush=True)
    return dict(status=np.array([sol.status, solq.status]), V=sol.cost, u0=sol.U[0], dV=upd["dL_dp"][0][:ng], dpi=upd["dpi_dp"][:, :ng],
                Q=solq.cost, dQ=updq["dL_dp"][0][:ng], slmax=sl)


def main(n=16, procs=8, name="cartpole_original", seed=2468):
    from .make_golden_large import _problem, _states

    pb = _problem(name)
    x0s, acts = _states(name, n, seed=seed)
    rng = np.random.default_rng(seed + 1)
    th = np.tile(pb.p_nominal, (n, 1))
    if name == "cartpole_original":
        th[:, :3] *= rng.uniform(0.6, 1.4, size=(n, 3))
    elif name == "linear_system":
        th += 0.03 * rng.standard_normal(th.shape)
    else:  # evaporation: theta = [W_0 (5 x 5), W (5 x 5), yref_0 (5), yref (5)]
        for i in range(n):
            for o in (0, 25):
                D = np.diag(rng.uniform(0.8, 1.2, size=5))
                th[i, o:o + 25] = (D @ th[i, o:o + 25].reshape(5, 5) @ D).T.ravel()
            th[i, 50:] *= rng.uniform(0.97, 1.03, size=10)
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_one, [(name, i, x0s[i], acts[i], th[i]) for i in range(n)], chunksize=1)
    out = {k: np.array([r[k] for r in res]) for k in res[0]}
    out.update(x0=x0s, a=acts, theta=th)
    path = os.path.join(ROOT, "tests", "golden", f"{name}_theta.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "converged (V, Q):", (out["status"] == 0).sum(0))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 8,
         sys.argv[3] if len(sys.argv) > 3 else "cartpole_original")
