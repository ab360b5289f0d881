"""Chain-of-masses MPC on the warp-cooperative CUDA engine against the dense oracle.

Fixtures: tests/golden/chain_mass_{3,5,6,3_64,5_64}.npz (oracle/make_golden_chain.py; oracle outputs, parity unpinned vs acados).
Tolerances vs the restated oracle: |u0| 1e-8, V/Q 1e-9 rel, dL/dtheta 1e-6 rel, dpi/dtheta 1e-5 rel.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _engine(n_mass, B, tol=1e-9):
    from mpc4rl_b200 import BatchedMPC
    from mpc4rl_b200.problems import chain_mass_spec, get_chain_params

    cp = get_chain_params()
    cp["n_mass"] = n_mass
    spec = chain_mass_spec(cp)
    mpc = BatchedMPC(spec, max_batch=B, device=0)
    mpc.set_option("tol", tol)
    return spec, mpc


def _T(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")


def _fixture(name):
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}.npz not generated")
    g = np.load(path)
    return g, int(g["n_mass"])


FIXTURES = ["chain_mass_3", "chain_mass_5", "chain_mass_6", "chain_mass_3_64", "chain_mass_5_64"]  # 16, 16, 16, 64 and 64 samples (n_mass 6: nx = 27, 800 parameters)


@pytest.mark.parametrize("fixture", FIXTURES)
def test_converged_solve_and_sensitivities_match_oracle(fixture):
    g, n_mass = _fixture(fixture)
    B = g["x0"].shape[0]
    spec, mpc = _engine(n_mass, B)
    assert spec.ntheta == g["theta"].shape[0] and np.abs(spec.x_ss - g["x_ss"]).max() < 1e-12
    x0 = _T(g["x0"])
    mpc.reset(x0)
    out = mpc.solve_sens(x0, max_sqp=60)
    torch.cuda.synchronize()
    st = out["status"].cpu().numpy()
    assert np.array_equal(st, g["status"][:, 0]), st
    assert np.abs(out["u0"].cpu().numpy() - g["u0"]).max() < 1e-8
    assert np.abs(out["cost"].cpu().numpy() - g["V"]).max() < 1e-9 * np.abs(g["V"]).max()
    assert out["res"].cpu().numpy().max() < 1e-8
    dL, dpi = out["dL"].cpu().numpy(), out["dpi"].cpu().numpy()
    assert dL.shape == (B, spec.ntheta) and dpi.shape == (B, 3, spec.ntheta)
    assert np.abs(dL - g["dV"]).max() < 1e-6 * np.abs(g["dV"]).max()
    assert np.abs(dpi - g["dpi"]).max() < 1e-5 * np.abs(g["dpi"]).max()
    # the iterate itself (what ocp_solver.get(k, "x" / "u" / "pi") returns)
    X = np.stack([mpc.get("x", k, B).cpu().numpy() for k in range(spec.N + 1)], 1)
    PI = np.stack([mpc.get("pi", k, B).cpu().numpy() for k in range(spec.N)], 1)
    assert np.abs(X - g["X"]).max() < 1e-7
    assert np.abs(PI.reshape(B, -1) - g["pi"]).max() < 1e-6 * max(1.0, np.abs(g["pi"]).max())
    # Q-mode: u_0 clamped to the action; dpi/dtheta = 0 by construction (quirk Q7)
    mpc.reset(x0)
    outq = mpc.solve_sens(x0, _T(g["a"]), max_sqp=60)
    assert np.array_equal(outq["status"].cpu().numpy(), g["status"][:, 1])
    assert np.abs(outq["cost"].cpu().numpy() - g["Q"]).max() < 1e-9 * np.abs(g["Q"]).max()
    assert np.abs(outq["dL"].cpu().numpy() - g["dQ"]).max() < 1e-6 * np.abs(g["dQ"]).max()
    assert np.abs(outq["u0"].cpu().numpy() - g["a"]).max() == 0.0
    assert float(outq["dpi"].abs().max()) == 0.0


@pytest.mark.parametrize("fixture", FIXTURES)
def test_rti_step_matches_oracle(fixture):
    """The path bench.py times for this problem: one SQP-RTI step from the stored (converged) iterate after the state
    moved, sensitivities at the resulting, not converged, iterate -- against the oracle's one full SQP step with the
    QP solved to the tau-central point and the restated update_nlp there."""
    g, n_mass = _fixture(fixture)
    B = g["x0"].shape[0]
    spec, mpc = _engine(n_mass, B)
    x0 = _T(g["x0"])
    mpc.reset(x0)
    mpc.solve(x0, max_sqp=60)
    out = mpc.solve_sens(_T(g["x1"]), max_sqp=1)  # default options: warm interior point, comp_accept 0.5
    torch.cuda.synchronize()
    assert np.all(out["status"].cpu().numpy() == 0)
    assert np.abs(out["u0"].cpu().numpy() - g["u1"]).max() < 1e-5
    assert np.abs(out["cost"].cpu().numpy() - g["V1"]).max() < 1e-8 * np.abs(g["V1"]).max()
    assert np.abs(out["dL"].cpu().numpy() - g["dV1"]).max() < 1e-5 * np.abs(g["dV1"]).max()
    assert np.abs(out["dpi"].cpu().numpy() - g["dpi1"]).max() < 1e-4 * np.abs(g["dpi1"]).max()


def test_reference_parameter_sweep_through_the_mirror():
    """tests/test_chain_mass.py -> examples/chain_mass.py:main_nlp(np_test=10): sweep C_{M}_0 from 0.5x to 1.5x,
    set_p / update / update_nlp / get_pi / get_dpi_dp per point (the reference only asserts that it runs; here the
    sensitivity column is also checked against a fine central difference of re-solved optima)."""
    from mpc4rl_b200.mpc.chain_mass.acados import AcadosMPC
    from mpc4rl_b200.mpc.chain_mass.ocp_utils import chain_define_x0, define_param_struct_symSX, find_idx_for_labels, get_chain_params

    cp = get_chain_params()
    mpc = AcadosMPC(cp, 1.0)
    mpc.ocp_solver.engine.set_option("tol", 1e-10)
    M = cp["n_mass"] - 2
    x0 = chain_define_x0(cp)
    p_idx = find_idx_for_labels(define_param_struct_symSX(cp["n_mass"], disturbance=True).cat, f"C_{M}_0")[0]
    p_nom = mpc.nlp.p.val.cat.full().flatten()
    assert p_nom.shape == (499,) and p_idx == 4 + 12 + 12 + 3 * M
    p_var = np.linspace(0.5 * p_nom[p_idx], 1.5 * p_nom[p_idx], 10)
    mpc.reset(x0)
    u_opt, sens_u = [], []
    for v in p_var:
        p = p_nom.copy()
        p[p_idx] = v
        mpc.set_p(p)
        assert mpc.update(x0) == 0
        mpc.update_nlp()
        u_opt.append(mpc.get_pi())
        assert mpc.get_dpi_dp().shape == (3, 499) and mpc.get_dV_dp().shape == (1, 499)
        sens_u.append(mpc.get_dpi_dp()[:, p_idx].flatten())
    u_opt, sens_u = np.vstack(u_opt), np.vstack(sens_u)
    # fine central difference at the middle point
    i, d = 4, 1e-4  # (solver tolerance / d = the noise of the difference quotient)
    up = []
    for s in (+1, -1):
        p = p_nom.copy()
        p[p_idx] = p_var[i] + s * d
        mpc.set_p(p)
        mpc.update(x0)
        up.append(mpc.get_pi())
    fd = (up[0] - up[1]) / (2 * d)
    assert np.abs(fd - sens_u[i]).max() < 2e-5 * max(1.0, np.abs(fd).max())
    # the coarse reconstruction the reference plots: cumulative sum of the gradients follows the optima
    rec = np.cumsum(sens_u, axis=0) * (p_var[1] - p_var[0])
    rec += u_opt[0] - rec[0]
    assert np.abs(rec - u_opt).max() < 0.05 * max(np.ptp(u_opt, axis=0).max(), 1e-3) + 1e-6


def test_batch_is_sample_independent_at_size():
    """BASELINE configs[2] shape (n_mass = 5, batch scaled to what a test may take): replicated samples give bitwise
    identical results wherever they sit in the batch, and every converged sample satisfies the KKT conditions."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "chain_mass_5.npz"))
    B = 2048
    spec, mpc = _engine(5, B, tol=1e-8)
    idx = np.arange(B) % g["x0"].shape[0]
    x0 = _T(g["x0"][idx])
    mpc.reset(x0)
    out = mpc.solve_sens(x0, max_sqp=60)
    torch.cuda.synchronize()
    assert int((out["status"] != 0).sum()) == 0
    assert float(out["res"].max()) < 1e-7
    for k in ("u0", "cost", "dL", "dpi"):
        a = out[k].cpu().numpy()
        assert np.array_equal(a[:16], a[16:32]) and np.array_equal(a[:16], a[B - 16:])
    assert np.abs(out["u0"].cpu().numpy()[:16] - g["u0"]).max() < 1e-8


def test_reference_acados_cross_check_path(tmp_path):
    """examples/chain_mass.py:main_acados with both solvers replaced by the shim: set(stage, "p", ..) on every stage,
    solve_for_x0, store_iterate -> load_iterate into a second solver (acados iterate JSON, lam / t in acados' stage
    layout), eval_solution_sensitivity(0, "params_global") -- the hand-off a machine with a real acados install would
    use to pin parity."""
    from mpc4rl_b200.mpc.chain_mass.ocp_utils import chain_define_x0, define_param_struct_symSX, find_idx_for_labels, get_chain_params
    from mpc4rl_b200.mpc.ocp_solver import OcpSolverShim
    from mpc4rl_b200.problems import chain_mass_spec

    cp = get_chain_params()
    cp["n_mass"] = 3
    spec = chain_mass_spec(cp)
    ocp_solver = OcpSolverShim(spec, max_iter=50, tol=1e-9)
    sens_solver = OcpSolverShim(spec, max_iter=50, tol=1e-9)
    M, x0 = cp["n_mass"] - 2, chain_define_x0(cp)
    p_idx = find_idx_for_labels(define_param_struct_symSX(cp["n_mass"], disturbance=True).cat, f"C_{M}_0")[0]
    p_val = spec.p_nominal.copy()
    fn = str(tmp_path / "iterate.json")
    us, ss = [], []
    for v in np.linspace(0.5, 1.5, 3) * spec.p_nominal[p_idx]:
        p_val[p_idx] = v
        for stage in range(spec.N + 1):
            ocp_solver.set(stage, "p", p_val)
            sens_solver.set(stage, "p", p_val)
        us.append(ocp_solver.solve_for_x0(x0))
        assert ocp_solver.status == 0
        ocp_solver.store_iterate(filename=fn, overwrite=True, verbose=False)
        import json
        d = json.load(open(fn))
        assert len(d["lam_0"]) == 2 * (3 + spec.nx) and len(d["lam_1"]) == 6 and len(d[f"lam_{spec.N}"]) == 0  # acados layout
        sens_solver.load_iterate(filename=fn, verbose=False)
        u2 = sens_solver.solve_for_x0(x0, fail_on_nonzero_status=False, print_stats_on_failure=False)
        assert sens_solver.status == 0 and np.abs(u2 - us[-1]).max() < 1e-8  # the loaded iterate is already the solution
        sens_x, sens_u = sens_solver.eval_solution_sensitivity(0, "params_global")
        assert sens_x.shape == (spec.nx, spec.ntheta) and sens_u.shape == (3, spec.ntheta) and not sens_x.any()
        ss.append(sens_u[:, p_idx])
    us, ss = np.array(us), np.array(ss)
    # secant slopes of the optima bracket the sensitivities (u0 is smooth and monotone in C over this range)
    h = 0.5 * spec.p_nominal[p_idx]
    sec = (us[1:] - us[:-1]) / h
    free = np.abs(ss).max(0) > 1e-8
    assert free.any()
    for j in np.where(free)[0]:
        assert min(ss[0, j], ss[1, j]) - 1e-6 <= sec[0, j] <= max(ss[0, j], ss[1, j]) + 1e-6 or abs(sec[0, j] - 0.5 * (ss[0, j] + ss[1, j])) < 0.2 * abs(sec[0, j])
