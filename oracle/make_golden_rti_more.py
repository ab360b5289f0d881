"""RTI-step fixtures for the secondary configurations (TEST INFRASTRUCTURE; oracle outputs, not acados outputs):
what oracle/make_golden_rti.py does for cartpole_original, for config/cartpole.yaml (state bounds), the linear system
(softened bound, EXACT Hessian) and the evaporation process (general rows, EXACT Hessian, N = 40).  Per sample:
  1. the dense oracle solves V(x0) to convergence                       -> stored iterate (U, X, pi, lam)
  2. the state moves: cart-pole by one environment step under the MPC policy (continuous_cartpole/environment.py:104-131,
     force_mag = 30 as config/cartpole.yaml's |u| <= 30), the linear system by one step of LinearSystemEnv
     (gym/linear_system/environment.py:15,32: A = [[.9,.35],[0,1.1]], B = [.0813,.2], noise U(-.1,.1) on x_1, clipped
     into the state box), the evaporation process by N(0, 0.05^2) on both states (5 x the bench's perturbation)
  3. ONE full SQP step from the stored iterate INCLUDING its multipliers (the exact Hessian of the first step needs pi)
     with x_0 := x1, QP at the tau-central point, V-mode and Q-mode
  4. restated update_nlp at the new, not converged, iterate.

    python -m oracle.make_golden_rti_more <cartpole_default|linear_system|evaporation|evaporation_n100> <n_samples> [n_procs]
(evaporation_n100: the reference's full horizon, about 3 minutes per 8 samples on 8 processes)
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_S = {}


def _move(name, x0, u0, rng):
    if name == "cartpole_default":
        from . import make_golden_rti as r

        old = r.ENV["force_mag"]
        r.ENV["force_mag"] = 30.0
        try:
            return r.env_step(x0, u0[0])
        finally:
            r.ENV["force_mag"] = old
    if name == "linear_system":
        A = np.array([[0.9, 0.35], [0.0, 1.1]]); B = np.array([0.0813, 0.2])
        x1 = A @ x0 + B * u0[0] + np.array([rng.uniform(-0.1, 0.1), 0.0])
        return np.clip(x1, [0.02, -0.98], [0.98, 0.98])
    return x0 + 0.05 * rng.standard_normal(2)


def _one(args):
    name, i, x0, a, seed = args
    import torch

    torch.set_num_threads(1)
    from .make_golden_large import _problem
    from .solver import DenseSolver

    if name not in _S:
        _S[name] = DenseSolver(_problem(name))
    s = _S[name]
    pb = s.pb
    nth, nu = len(pb.p_nominal), pb.nu
    rng = np.random.default_rng(seed)
    tol = 1e-9 if name.startswith("evaporation") else 1e-10
    nan = dict(x1=np.full(pb.nx, np.nan), V1=np.nan, u1=np.full(nu, np.nan), dV1=np.full(nth, np.nan), dpi1=np.full((nu, nth), np.nan),
               kkt1=np.nan, sl1=0.0, Q1=np.nan, dQ1=np.full(nth, np.nan))
    try:
        sol, _ = s.unit(x0, tol=tol)
    except Exception as e:  # noqa: BLE001
        print(f"[{name} {i}] oracle failed: {type(e).__name__}: {e}", flush=True)
        return dict(i=i, status=-1, ok1=False, **nan)
    out = dict(i=i, status=sol.status, ok1=False, **nan)
    if sol.status != 0:
        return out
    x1 = _move(name, np.asarray(x0, float), sol.U[0], rng)
    out["x1"] = x1
    try:
        init = (sol.U, sol.X, sol.pi, sol.lam)
        r, ru = s.unit(x1, init=init, max_iter=1, polish=False)
        rq, rqu = s.unit(x1, u0=a, init=init, max_iter=1, polish=False)
        sl = max(float(np.max(r.slbx, initial=0.0)), float(np.max(r.subx, initial=0.0))) if hasattr(r, "slbx") else 0.0
        out.update(ok1=bool(np.isfinite(r.cost) and np.isfinite(rq.cost)), V1=r.cost, u1=np.array(r.U[0]), dV1=ru["dL_dp"][0],
                   dpi1=ru["dpi_dp"], kkt1=r.kkt, sl1=sl, Q1=rq.cost, dQ1=rqu["dL_dp"][0])
        print(f"[{name} {i}] u0={sol.U[0]} -> u0'={r.U[0]} kkt'={r.kkt:.2e} V'={r.cost:.6f} Q'={rq.cost:.6f} sl'={sl:.1e}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[{name} {i}] RTI step failed: {type(e).__name__}: {e}", flush=True)
    return out


def main(name, n, workers):
    from .make_golden_large import _problem, _states

    pb = _problem(name)
    x0s, acts = _states(name, n, seed=9876)
    with mp.get_context("fork").Pool(workers) as pool:
        res = pool.map(_one, [(name, i, x0s[i], acts[i], 555 + i) for i in range(n)], chunksize=1)
    res.sort(key=lambda r: r["i"])
    A = lambda k: np.array([r[k] for r in res])
    dV, dpi, dQ = A("dV1"), A("dpi1"), A("dQ1")
    live = np.where((np.nan_to_num(np.abs(dV)).max(0) > 0) | (np.nan_to_num(np.abs(dQ)).max(0) > 0)
                    | (np.nan_to_num(np.abs(dpi)).max((0, 1)) > 0))[0]
    path = os.path.join(ROOT, "tests", "golden", f"{name}_rti.npz")
    np.savez_compressed(path, x0=x0s, a=acts, theta=pb.p_nominal, status=A("status"), ok1=A("ok1"), x1=A("x1"), V1=A("V1"), u1=A("u1"),
                        cols=live, dV1=dV[:, live], dpi1=dpi[:, :, live], kkt1=A("kkt1"), sl1=A("sl1"), Q1=A("Q1"), dQ1=dQ[:, live])
    print("wrote", path, "usable:", int(A("ok1").sum()), "of", n, "live gradient columns:", len(live))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 4)
