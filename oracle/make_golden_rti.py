"""Golden fixture for the path bench.py times: ONE SQP-RTI step from a stored iterate, then the restated
update_nlp at the iterate the step produced (TEST INFRASTRUCTURE; oracle outputs, not acados outputs).

Per sample (cartpole_original.yaml, SURVEY.md 8(d) config 2 state distribution):
  1. the dense oracle solves V(x0) to convergence from the MPC.reset guess  -> stored iterate (U, X, pi, lam, t)
  2. the state moves by one environment step under the MPC policy, x0' = env.step(x0, u0*) with the
     reference's ContinuousCartPole dynamics (rlmpc/gym/continuous_cartpole/environment.py:104-131, explicit
     Euler, tau = 0.02, force = u0 clipped to +-80 N; note polemass_length = m*l, quirk Q9)
  3. ONE full SQP step from the stored iterate with x_0 := x0' (QP solved to the tau-central point by the dense
     interior-point method of oracle/solver.py), V-mode; the same in Q-mode with a random action
  4. restated update_nlp (dense dR/dz + SuperLU) at the new, in general NOT converged, iterate.

    python -m oracle.make_golden_rti [n_samples] [n_procs]     # tests/golden/cartpole_original_rti.npz
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ENV = dict(gravity=9.8, masscart=1.0, masspole=0.1, length=0.5, force_mag=80.0, tau=0.02)


def env_step(x, force):
    """continuous_cartpole/environment.py:104-131 (euler branch); force in newtons."""
    g, mc, mp_, ln, tau = ENV["gravity"], ENV["masscart"], ENV["masspole"], ENV["length"], ENV["tau"]
    total, pml = mc + mp_, mp_ * ln
    s, sd, th, thd = x
    f = float(np.clip(force, -ENV["force_mag"], ENV["force_mag"]))
    ct, sn = np.cos(th), np.sin(th)
    temp = (f + pml * thd**2 * sn) / total
    thacc = (g * sn - ct * temp) / (ln * (4.0 / 3.0 - mp_ * ct**2 / total))
    xacc = temp - pml * thacc * ct / total
    return np.array([s + tau * sd, sd + tau * xacc, th + tau * thd, thd + tau * thacc])


def _one(args):
    i, x0, a = args
    import torch

    torch.set_num_threads(1)
    from .problems import make_cartpole
    from .solver import DenseSolver

    s = DenseSolver(make_cartpole("original"))
    sol, upd = s.unit(x0, tol=1e-10)
    x1 = env_step(x0, sol.U[0, 0])
    r, ru = s.unit(x1, init=(sol.U, sol.X), max_iter=1, polish=False)
    rq, rqu = s.unit(x1, u0=a, init=(sol.U, sol.X), max_iter=1, polish=False)
    print(f"[rti {i}] st={sol.status} it={sol.sqp_iter} u0={sol.U[0, 0]:.6f} -> u0'={r.U[0, 0]:.6f} kkt'={r.kkt:.2e}", flush=True)
    return dict(
        x0=x0, a=a, status=sol.status, V=sol.cost, u0=sol.U[0], dV=upd["dL_dp"][0], dpi=upd["dpi_dp"],
        U=sol.U, X=sol.X, pi=sol.pi, lam=sol.lam, t=sol.t,
        x1=x1, V1=r.cost, u1=r.U[0], dV1=ru["dL_dp"][0], dpi1=ru["dpi_dp"], kkt1=r.kkt, U1=r.U, X1=r.X,
        Q1=rq.cost, dQ1=rqu["dL_dp"][0], kktq1=rq.kkt,
    )


def main(n=256, procs=8, seed=4321):
    rng = np.random.default_rng(seed)
    lo = np.array([-1.0, -2.0, -np.pi, -4.0])
    x0s = rng.uniform(lo, -lo, size=(n, 4))
    acts = rng.uniform(-80.0, 80.0, size=(n, 1))
    x0s[0] = [0.0, 0.0, np.pi / 2, 0.0]; acts[0] = [-30.0]  # scripts/cartpole_mpc_sensitivities.py:80-81
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_one, [(i, x0s[i], acts[i]) for i in range(n)], chunksize=4)
    out = {k: np.array([r[k] for r in res]) for k in res[0]}
    path = os.path.join(ROOT, "tests", "golden", "cartpole_original_rti.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 8)
