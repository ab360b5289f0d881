"""Golden fixture for cartpole with parameterize_tracking_cost=True (TEST INFRASTRUCTURE; oracle outputs): the cost weights
and references W_0, W, W_e, yref_0, yref, yref_e are part of theta with non-zero gradient (nlp.py:1057-1074), so dL/dp and
dpi/dp have 83 live columns instead of 3.

    python -m oracle.make_golden_paramcost      # tests/golden/cartpole_paramcost.npz
"""
from __future__ import annotations

import os

import numpy as np

from .problems import make_cartpole
from .solver import DenseSolver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(n=6, seed=77):
    pb = make_cartpole("original")
    pb.parameterize_tracking_cost = True
    s = DenseSolver(pb)
    rng = np.random.default_rng(seed)
    lo = np.array([-1.0, -2.0, -np.pi, -4.0])
    x0s = rng.uniform(lo, -lo, size=(n, 4))
    x0s[0] = [0.0, 0.0, np.pi / 2, 0.0]
    acts = rng.uniform(-80.0, 80.0, size=(n, 1))
    # a theta with a non-symmetric W and non-zero references, so that every column is exercised
    p = pb.p_nominal.copy()
    p[3:3 + 66] += 2e-3 * rng.standard_normal(66)  # keeps W positive definite (smallest diagonal entry: 0.01)
    p[69:] = 0.1 * rng.standard_normal(14)
    out = {k: [] for k in ("V", "u0", "dV", "dpi", "Q", "dQ", "status")}
    for i in range(n):
        sol, upd = s.unit(x0s[i], p=p, tol=1e-10)
        solq, updq = s.unit(x0s[i], u0=acts[i], p=p, tol=1e-10)
        print(f"[paramcost {i}] V={sol.cost:.6f} u0={sol.U[0]} it={sol.sqp_iter} st={sol.status} | Q={solq.cost:.6f} st={solq.status}", flush=True)
        out["status"].append([sol.status, solq.status])
        out["V"].append(sol.cost); out["u0"].append(sol.U[0]); out["dV"].append(upd["dL_dp"][0]); out["dpi"].append(upd["dpi_dp"])
        out["Q"].append(solq.cost); out["dQ"].append(updq["dL_dp"][0])
    out = {k: np.array(v) for k, v in out.items()}
    out["x0"] = x0s; out["a"] = acts; out["theta"] = p
    path = os.path.join(ROOT, "tests", "golden", "cartpole_paramcost.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
