"""Generate the golden fixtures under tests/golden/ with the dense oracle (TEST INFRASTRUCTURE).

NOT reference outputs: the reference cannot run here (no acados/CasADi) and ships
no golden vectors; these are outputs of the float64 restatement in oracle/ (cold-started dense
SQP+IPM, then restated update_nlp with dense dR/dz + SuperLU).  "Parity unpinned" applies.

    python -m oracle.make_golden            # writes tests/golden/*.npz (about 3 minutes)
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .problems import make_cartpole
from .solver import DenseSolver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sample_states(n, seed, variant="original"):
    """'original': SURVEY.md 8(d) config 2 distribution.  'default' (config/cartpole.yaml, state
    bounds, |u| <= 30): states around the hanging position the YAML starts from (x0 = [0,0,3.14,0]);
    far-away starts make the first linearised QP infeasible there (acados: status 4)."""
    rng = np.random.default_rng(seed)
    if variant == "original":
        lo = np.array([-1.0, -2.0, -np.pi, -4.0]); hi = -lo
        x = rng.uniform(lo, hi, size=(n, 4))
        a = rng.uniform(-80.0, 80.0, size=(n, 1))
    else:
        x = np.array([0.0, 0.0, np.pi, 0.0]) + rng.uniform(-1.0, 1.0, size=(n, 4)) * np.array([1.0, 1.0, 0.5, 1.0])
        a = rng.uniform(-30.0, 30.0, size=(n, 1))
    return x, a


def cartpole_golden(variant="original", n=20, seed=1234):
    pb = make_cartpole(variant)
    s = DenseSolver(pb)
    x0s, acts = sample_states(n, seed, variant)
    if variant == "original":
        # the reference's own test point (scripts/cartpole_mpc_sensitivities.py:80-81)
        x0s[0] = [0.0, 0.0, np.pi / 2, 0.0]; acts[0] = [-30.0]
    else:
        x0s[0] = [0.0, 0.0, 3.14, 0.0]; acts[0] = [-30.0]  # config/cartpole.yaml:80, u0 on the bound
    out = {k: [] for k in ("V", "u0", "dV", "dpi", "Q", "dQ", "U", "X", "pi", "lam", "t", "UQ", "XQ", "status")}
    for i in range(n):
        sol, upd = s.unit(x0s[i], tol=1e-10)
        solq, updq = s.unit(x0s[i], u0=acts[i], tol=1e-10)
        print(f"[{variant} {i}] V={sol.cost:.6f} u0={sol.U[0]} it={sol.sqp_iter} kkt={sol.kkt:.1e} | Q={solq.cost:.6f} it={solq.sqp_iter}",
              flush=True)
        out["status"].append([sol.status, solq.status])
        out["V"].append(sol.cost); out["u0"].append(sol.U[0]); out["dV"].append(upd["dL_dp"][0]); out["dpi"].append(upd["dpi_dp"])
        out["Q"].append(solq.cost); out["dQ"].append(updq["dL_dp"][0])
        out["U"].append(sol.U); out["X"].append(sol.X); out["pi"].append(sol.pi); out["lam"].append(sol.lam); out["t"].append(sol.t)
        out["UQ"].append(solq.U); out["XQ"].append(solq.X)
    out = {k: np.array(v) for k, v in out.items()}
    out["x0"] = x0s; out["a"] = acts; out["theta"] = pb.p_nominal
    path = os.path.join(ROOT, "tests", "golden", f"cartpole_{variant}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def linear_system_golden(gamma=0.9, seed=7):
    """Linear system (rlmpc/mpc/linear_system/acados.py) at the example's discount factor.  States
    are drawn from the box the environment visits; samples whose optimum violates the soft bound on
    x[0] exercise the slack rows.  Also the LQR known answers of SURVEY.md 8(c) (widened bounds)."""
    from .problems import make_linear_system

    pb = make_linear_system(gamma=gamma)
    s = DenseSolver(pb)
    rng = np.random.default_rng(seed)
    x0s = np.vstack([[0.5, 0.5], [0.2, 0.2], [0.5, -0.2], [0.9, 0.8], [0.05, -0.6],
                     rng.uniform([0.0, -1.0], [1.0, 1.0], size=(7, 2))])
    acts = rng.uniform(-1.0, 1.0, size=(len(x0s), 1))
    out = {k: [] for k in ("V", "u0", "dV", "dpi", "Q", "dQ", "U", "X", "lam", "t", "slmax", "status")}
    for i in range(len(x0s)):
        sol, upd = s.unit(x0s[i], tol=1e-10)
        solq, updq = s.unit(x0s[i], u0=acts[i], tol=1e-10)
        print(f"[linear {i}] V={sol.cost:.6f} u0={sol.U[0]} it={sol.sqp_iter} | Q={solq.cost:.6f} slack={max(sol.slbx.max(), sol.subx.max()):.3f}", flush=True)
        out["status"].append([sol.status, solq.status])
        out["V"].append(sol.cost); out["u0"].append(sol.U[0]); out["dV"].append(upd["dL_dp"][0]); out["dpi"].append(upd["dpi_dp"])
        out["Q"].append(solq.cost); out["dQ"].append(updq["dL_dp"][0])
        out["U"].append(sol.U); out["X"].append(sol.X); out["lam"].append(sol.lam); out["t"].append(sol.t)
        out["slmax"].append(max(sol.slbx.max(), sol.subx.max(), solq.slbx.max(), solq.subx.max()))
    out = {k: np.array(v) for k, v in out.items()}
    out["x0"] = x0s; out["a"] = acts; out["theta"] = pb.p_nominal; out["gamma"] = gamma
    path = os.path.join(ROOT, "tests", "golden", "linear_system.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def evaporation_golden(gamma=0.95, N=40, seed=7):
    """Evaporation process (rlmpc/mpc/evaporation_process/acados.py): affine h rows shared with the slack
    input, discounted tracking cost, theta = (W_0, W, yref_0, yref) = 60 entries.  Horizon N=40 instead of
    the reference's 100: the dense oracle needs ~30 s per solve at N=40 and ~10 min at N=100 (one N=100
    sample is added by ``evaporation_golden_full``)."""
    from .problems import make_evaporation

    pb = make_evaporation(gamma=gamma, N=N)
    s = DenseSolver(pb)
    rng = np.random.default_rng(seed)
    # SURVEY.md 8(d) config 4 distribution X_2~U(25,40), P_2~U(49.7,70), plus one state below the soft bound
    x0s = np.vstack([[30.0, 60.0], [25.5, 50.0], [24.0, 55.0], rng.uniform([25.0, 49.7], [40.0, 70.0], size=(2, 2))])
    acts = np.column_stack([rng.uniform(150.0, 350.0, size=(len(x0s), 2)), np.full(len(x0s), 0.5)])
    out = {k: [] for k in ("V", "u0", "dV", "dpi", "Q", "dQ", "U", "X", "status")}
    nan = lambda *shape: np.full(shape, np.nan)
    for i in range(len(x0s)):
        res = []
        for u0 in (None, acts[i]):
            try:  # the dense IPM has no safeguards: a failed solve is recorded as status -1, not fixed up
                sol, upd = s.unit(x0s[i], u0=u0, tol=1e-9)
                res.append((sol.status, sol.cost, sol.U, sol.X, upd["dL_dp"][0], upd["dpi_dp"]))
            except Exception as e:  # noqa: BLE001
                print(f"[evaporation {i}] oracle failed: {type(e).__name__}: {e}", flush=True)
                res.append((-1, np.nan, nan(N, 3), nan(N + 1, 2), nan(60), nan(3, 60)))
        (sv, V, U, X, dV, dpi), (sq, Q, _, _, dQ, _) = res
        print(f"[evaporation {i}] V={V:.6f} u0={U[0]} st={sv} | Q={Q:.6f} st={sq}", flush=True)
        out["status"].append([sv, sq])
        out["V"].append(V); out["u0"].append(U[0]); out["dV"].append(dV); out["dpi"].append(dpi)
        out["Q"].append(Q); out["dQ"].append(dQ); out["U"].append(U); out["X"].append(X)
    out = {k: np.array(v) for k, v in out.items()}
    out["x0"] = x0s; out["a"] = acts; out["theta"] = pb.p_nominal; out["gamma"] = gamma; out["N"] = N
    path = os.path.join(ROOT, "tests", "golden", "evaporation.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def evaporation_golden_full(gamma=0.95):
    """One V-mode sample at the reference's full horizon N=100."""
    from .problems import make_evaporation

    pb = make_evaporation(gamma=gamma, N=100)
    x0 = np.array([30.0, 60.0])
    sol, upd = DenseSolver(pb).unit(x0, tol=1e-9)
    print(f"[evaporation N=100] V={sol.cost:.6f} u0={sol.U[0]} it={sol.sqp_iter} st={sol.status}", flush=True)
    path = os.path.join(ROOT, "tests", "golden", "evaporation_n100.npz")
    np.savez_compressed(path, x0=x0, V=sol.cost, u0=sol.U[0], dV=upd["dL_dp"][0], dpi=upd["dpi_dp"], X=sol.X, U=sol.U,
                        status=sol.status, gamma=gamma, theta=pb.p_nominal)
    print("wrote", path)


def chain_mass_golden(n_mass=3, n=4, seed=50):
    """Chain of masses (rlmpc/mpc/chain_mass/ocp_utils.py) at n_mass = 3 (nx = 9, nu = 3, ntheta = 113; what
    rlmpc/examples/chain_mass.py:__main__ runs): define_x0 and perturbations of it (perturb_scale = 1e-2, seed 50 as
    in get_chain_params), V-mode and Q-mode with |u0| <= 1.  The specification the CUDA path for this problem
     has to meet."""
    from .problems import make_chain_mass

    pb = make_chain_mass(n_mass)
    s = DenseSolver(pb)
    rng = np.random.default_rng(seed)
    x0s = np.vstack([pb.x0_example] + [pb.x0_example + 1e-2 * rng.standard_normal(pb.nx) for _ in range(n - 1)])
    acts = rng.uniform(-0.8, 0.8, size=(n, 3))
    out = {k: [] for k in ("V", "u0", "dV", "dpi", "Q", "dQ", "U", "X", "pi", "status")}
    for i in range(n):
        sol, upd = s.unit(x0s[i], tol=1e-10)
        solq, updq = s.unit(x0s[i], u0=acts[i], tol=1e-10)
        print(f"[chain_mass_{n_mass} {i}] V={sol.cost:.8f} u0={sol.U[0]} it={sol.sqp_iter} st={sol.status} | Q={solq.cost:.8f} st={solq.status}", flush=True)
        out["status"].append([sol.status, solq.status])
        out["V"].append(sol.cost); out["u0"].append(sol.U[0]); out["dV"].append(upd["dL_dp"][0]); out["dpi"].append(upd["dpi_dp"])
        out["Q"].append(solq.cost); out["dQ"].append(updq["dL_dp"][0]); out["U"].append(sol.U); out["X"].append(sol.X); out["pi"].append(sol.pi)
    out = {k: np.array(v) for k, v in out.items()}
    out["x0"] = x0s; out["a"] = acts; out["theta"] = pb.p_nominal; out["x_ss"] = pb.x_ss; out["n_mass"] = n_mass
    path = os.path.join(ROOT, "tests", "golden", f"chain_mass_{n_mass}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def cartpole_free_g_golden(n=6, seed=1234):
    """cartpole_original with g un-fixed as scripts/cartpole_mpc_qlearning.py:184-187 does: theta = [M, m, l, g | W_0 ...]
    (84 entries); the states of ``cartpole_golden('original')``'s first samples."""
    pb = make_cartpole("original", free_g=True)
    s = DenseSolver(pb)
    x0s, acts = sample_states(n, seed, "original")
    x0s[0] = [0.0, 0.0, np.pi / 2, 0.0]
    out = {k: [] for k in ("V", "u0", "dV", "dpi", "Q", "dQ", "status")}
    for i in range(n):
        sol, upd = s.unit(x0s[i], tol=1e-10)
        solq, updq = s.unit(x0s[i], u0=acts[i], tol=1e-10)
        print(f"[free g {i}] V={sol.cost:.6f} u0={sol.U[0]} st={sol.status} | Q={solq.cost:.6f} st={solq.status}", flush=True)
        out["status"].append([sol.status, solq.status])
        out["V"].append(sol.cost); out["u0"].append(sol.U[0]); out["dV"].append(upd["dL_dp"][0]); out["dpi"].append(upd["dpi_dp"])
        out["Q"].append(solq.cost); out["dQ"].append(updq["dL_dp"][0])
    out = {k: np.array(v) for k, v in out.items()}
    out["x0"] = x0s; out["a"] = acts; out["theta"] = pb.p_nominal
    path = os.path.join(ROOT, "tests", "golden", "cartpole_free_g.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    which = sys.argv[1:] or ["original"]
    for v in which:
        if v == "free_g":
            cartpole_free_g_golden()
        elif v == "evaporation":
            evaporation_golden()
        elif v == "evaporation_full":
            evaporation_golden_full()
        elif v == "linear":
            linear_system_golden()
        elif v == "chain_mass":
            chain_mass_golden()
        else:
            cartpole_golden(v, n=20 if v == "original" else 12)
