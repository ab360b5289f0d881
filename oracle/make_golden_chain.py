"""Golden fixtures of the chain-of-masses problem (TEST INFRASTRUCTURE; oracle outputs, not acados outputs).

rlmpc/mpc/chain_mass/ocp_utils.py:42-56, 59-147, 195-316, 353-371 as restated in oracle/problems.py.  Per sample:
  * V(x0) and Q(x0, a) solved to convergence from the MPC.reset guess, restated update_nlp at the solution
    (x0 = define_x0 + N(0, 1e-2), rlmpc/examples/chain_mass.py:17-25, perturb_scale of ocp_utils.py:334, seed 50)
  * one SQP-RTI step from the converged V iterate after the state moved by one closed-loop step of the nominal
    model plus a small disturbance, x0' = f_disc(x0, u0*) + N(0, 1e-3), and update_nlp at the resulting iterate.

    python -m oracle.make_golden_chain <n_mass> <n_samples> [n_procs] [tag]   # tests/golden/chain_mass_<n_mass><tag>.npz
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _one(args):
    n_mass, i, x0, a, noise = args
    import torch

    torch.set_num_threads(1)
    from .problems import F64, make_chain_mass
    from .solver import DenseSolver

    pb = make_chain_mass(n_mass)
    s = DenseSolver(pb)
    sol, upd = s.unit(x0, tol=1e-10)
    solq, updq = s.unit(x0, u0=a, tol=1e-10)
    T = lambda v: torch.as_tensor(v, dtype=F64)
    x1 = pb.f_disc(T(x0), T(sol.U[0]), T(pb.p_nominal)).numpy() + noise
    r, ru = s.unit(x1, init=(sol.U, sol.X), max_iter=1, polish=False)
    print(f"[chain_mass_{n_mass} {i}] V={sol.cost:.8f} u0={sol.U[0]} it={sol.sqp_iter} st={sol.status} | Q={solq.cost:.8f} st={solq.status}"
          f" | rti u0'={r.U[0]} kkt'={r.kkt:.2e}", flush=True)
    return dict(x0=x0, a=a, status=np.array([sol.status, solq.status]), V=sol.cost, u0=sol.U[0], dV=upd["dL_dp"][0], dpi=upd["dpi_dp"],
                Q=solq.cost, dQ=updq["dL_dp"][0], U=sol.U, X=sol.X, pi=sol.pi, lam=sol.lam, t=sol.t,
                x1=x1, V1=r.cost, u1=r.U[0], dV1=ru["dL_dp"][0], dpi1=ru["dpi_dp"], kkt1=r.kkt)


def main(n_mass=3, n=4, procs=4, seed=50, tag=""):
    from .problems import make_chain_mass

    pb = make_chain_mass(n_mass)
    rng = np.random.default_rng(seed)
    x0s = np.vstack([pb.x0_example] + [pb.x0_example + 1e-2 * rng.standard_normal(pb.nx) for _ in range(n - 1)])
    acts = rng.uniform(-0.8, 0.8, size=(n, 3))
    noise = 1e-3 * rng.standard_normal((n, pb.nx))
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_one, [(n_mass, i, x0s[i], acts[i], noise[i]) for i in range(n)], chunksize=1)
    out = {k: np.array([r[k] for r in res]) for k in res[0]}
    out["theta"] = pb.p_nominal; out["x_ss"] = pb.x_ss; out["n_mass"] = n_mass
    path = os.path.join(ROOT, "tests", "golden", f"chain_mass_{n_mass}{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    # optional 4th argument: file-name tag of an additional (larger) set, e.g. "_64" -> chain_mass_3_64.npz
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 4, tag=sys.argv[4] if len(sys.argv) > 4 else "")
