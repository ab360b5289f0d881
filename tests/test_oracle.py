"""Pins of the oracle itself (CPU).  The reference ships no golden vectors and cannot run here
(SURVEY.md 8(c)); what it does have are relational checks, reproduced here on the restatement:
  * update_nlp's assertions (rlmpc/mpc/nlp.py:1445-1537): KKT self-consistency at the solution
  * finite-difference agreement of dV/dp (scripts/linear_system_mpc_nlp.py:43-49), here at 1e-5
    instead of the reference's atol=1e-1
plus consistency of the committed golden fixtures and of the host port of the engine."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cartpole():
    from oracle.problems import make_cartpole
    from oracle.solver import DenseSolver

    return DenseSolver(make_cartpole("original"))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cartpole_original.npz"))


def test_layout_matches_reference_counts(cartpole):
    nlp = cartpole.nlp
    # cartpole_original N=40: nw=204, npi=160, nlam=88, nz=540, ntheta=83
    assert (nlp.nw, nlp.npi, nlp.nlam, nlp.nz, cartpole.pb.ntheta) == (204, 160, 88, 540, 83)
    kinds = [r.kind for r in nlp.rows if r.stage == 0]
    assert kinds == ["lbu"] + ["lbx"] * 4 + ["ubu"] + ["ubx"] * 4  # acados order, stage 0
    assert [r.kind for r in nlp.rows if r.stage == 5] == ["lbu", "ubu"]


def test_update_nlp_assertions_hold_at_golden_solution(cartpole, golden):
    """nlp.py:1445-1537 on a stored solution: g~0, h<=0, complementarity, stationarity, R~0."""
    from oracle.nlp import Bounds

    pb, nlp = cartpole.pb, cartpole.nlp
    for i in (0, 2):
        N = pb.N
        b = Bounds(lbu=np.tile(pb.lbu, (N, 1)), ubu=np.tile(pb.ubu, (N, 1)), lbx0=golden["x0"][i], ubx0=golden["x0"][i],
                   slbx=np.zeros((N + 1, 0)), subx=np.zeros((N + 1, 0)))
        upd = nlp.update(golden["U"][i], golden["X"][i], golden["pi"][i], golden["lam"][i], golden["t"][i], pb.p_nominal, b)
        assert abs(upd["cost"] - golden["V"][i]) < 1e-9 * abs(golden["V"][i])
        assert np.abs(upd["g"]).max() <= 1e-4 and np.abs(upd["g"]).max() < 1e-10
        assert (upd["h"] < 1e-6).all()
        assert np.abs(golden["lam"][i] * upd["h"]).max() <= 1e-5
        nu = pb.nu * N
        assert np.abs(upd["dL_dw"][:nu]).max() < 1e-3 and np.abs(upd["dL_dw"][:nu]).max() < 1e-8  # dL/du
        assert np.abs(upd["dL_dw"][nu + pb.nx:]).max() < 1e-8  # dL/dx (x_0 rows hold the equality multipliers)
        assert np.abs(upd["R"]).max() < 1e-6  # nlp.assert_kkt_residual (nlp.py:1295-1299)
        assert np.allclose(upd["dL_dp"][0], golden["dV"][i], rtol=1e-9, atol=1e-12)
        assert np.allclose(upd["dpi_dp"], golden["dpi"][i], rtol=1e-7, atol=1e-9)


def test_value_gradient_matches_finite_differences(cartpole, golden):
    """dV/dp = dL/dp (envelope theorem) against central differences of the re-solved V."""
    pb = cartpole.pb
    i = 1
    x0 = golden["x0"][i]
    init = (golden["U"][i], golden["X"][i])
    d = 1e-5
    for j in range(3):
        p = pb.p_nominal.copy(); p[j] += d
        vp = cartpole.solve(x0, p=p, init=init, tol=1e-11)
        p[j] -= 2 * d
        vm = cartpole.solve(x0, p=p, init=init, tol=1e-11)
        fd = (vp.cost - vm.cost) / (2 * d)
        assert abs(fd - golden["dV"][i][j]) < 1e-5 * max(1.0, abs(fd))
        fdu = (vp.U[0, 0] - vm.U[0, 0]) / (2 * d)
        assert abs(fdu - golden["dpi"][i][0, j]) < 1e-4 * max(1.0, abs(fdu))


def test_scipy_cross_check_of_the_minimiser(cartpole, golden):
    """Independent NLP solve (scipy SLSQP on the restated cost/constraints) reaches the same optimum."""
    import torch
    from scipy.optimize import minimize
    from torch.func import grad, jacrev

    from oracle.nlp import Bounds
    from oracle.problems import F64

    pb, nlp = cartpole.pb, cartpole.nlp
    i = 4
    x0 = golden["x0"][i]
    N = pb.N
    b = Bounds(lbu=np.tile(pb.lbu, (N, 1)), ubu=np.tile(pb.ubu, (N, 1)), lbx0=x0, ubx0=x0, slbx=np.zeros((N + 1, 0)),
               subx=np.zeros((N + 1, 0)))
    p = torch.as_tensor(pb.p_nominal, dtype=F64)
    T = lambda w: torch.as_tensor(w, dtype=F64)
    f = lambda w: float(nlp.cost(T(w), p, b))
    df = lambda w: grad(lambda w_: nlp.cost(w_, p, b))(T(w)).numpy()
    nu = N * pb.nu
    geq = lambda w: np.concatenate([nlp.g(T(w), p).numpy(), w[nu:nu + pb.nx] - x0])
    E0 = np.zeros((pb.nx, nlp.nw)); E0[:, nu:nu + pb.nx] = np.eye(pb.nx)
    dgeq = lambda w: np.vstack([jacrev(lambda w_: nlp.g(w_, p))(T(w)).numpy(), E0])
    w0 = nlp.pack_w(golden["U"][i] * 0.9, golden["X"][i]).numpy()
    bounds = [(-80.0, 80.0)] * nu + [(None, None)] * (nlp.nw - nu)
    r = minimize(f, w0, jac=df, bounds=bounds, constraints=[{"type": "eq", "fun": geq, "jac": dgeq}], method="SLSQP",
                 options={"maxiter": 300, "ftol": 1e-14})
    assert abs(r.fun - golden["V"][i]) < 1e-6 * abs(golden["V"][i])
    assert abs(r.x[0] - golden["u0"][i][0]) < 1e-3


def test_q_mode_fixes_first_input(golden):
    assert np.array_equal(golden["UQ"][:, 0, :], golden["a"])
    ok = (golden["status"] == 0).all(axis=1)
    assert (golden["Q"][ok] >= golden["V"][ok] - 1e-9).all()  # Q(s,a) >= V(s) = min_a Q(s,a)
    # full-step SQP (acados' default, no globalisation) 2-cycles on samples 3 and 18 of this seed
    assert ok.sum() == 18 and not ok[3] and not ok[18]


def test_host_port_of_engine_matches_oracle(golden):
    """The Riccati/IPM/adjoint engine (host build of the CUDA templates) against the dense oracle."""
    from oracle import cpu_port as cp
    from oracle.problems import make_cartpole

    pb = make_cartpole("original")
    pd = cp.make_pd(pb.N, [pb.stage_scale(k) for k in range(pb.N + 1)], pb.lbu, pb.ubu, [pb.tf / pb.N / 4, 9.8], tol=1e-10)
    o = cp.unit(1, pd, 0, 200, pb.p_nominal, golden["x0"])
    ok = golden["status"][:, 0] == 0
    assert np.array_equal(o["status"] == 0, ok)  # the same samples 2-cycle in both implementations
    assert np.abs(o["u0"] - golden["u0"])[ok].max() < 1e-6
    assert np.abs(o["cost"] - golden["V"])[ok].max() < 1e-9 * np.abs(golden["V"]).max()
    assert np.abs(o["dL"][:, :3] - golden["dV"][:, :3])[ok].max() < 1e-6 * np.abs(golden["dV"]).max()
    assert np.abs(o["dpi"][:, :, :3] - golden["dpi"][:, :, :3])[ok].max() < 1e-5 * np.abs(golden["dpi"]).max()
    q = cp.unit(1, pd, 1, 200, pb.p_nominal, golden["x0"], u0=golden["a"])
    okq = golden["status"][:, 1] == 0
    assert np.array_equal(q["status"] == 0, okq)
    assert np.abs(q["cost"] - golden["Q"])[okq].max() < 1e-9 * np.abs(golden["Q"]).max()
    assert np.abs(q["dL"][:, :3] - golden["dQ"][:, :3])[okq].max() < 1e-6 * np.abs(golden["dQ"]).max()
    assert np.count_nonzero(q["dpi"]) == 0


# ------------------------------------------------------------------------------------------------
# other problem definitions: state-bounded cartpole (config/cartpole.yaml) and the linear system
# ------------------------------------------------------------------------------------------------
def test_host_port_with_state_bounds_matches_oracle_fixtures():
    from oracle import cpu_port as cp
    from oracle.problems import make_cartpole

    pb = make_cartpole("default")
    g = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_default.npz"))
    scale = np.array([pb.stage_scale(k) for k in range(pb.N + 1)])
    pd = cp.make_pd(pb.N, scale, pb.lbu, pb.ubu, [pb.tf / pb.N / 4, 9.8], tol=1e-10, warm_ipm=1,
                    lbx=pb.lbx, ubx=pb.ubx, lbx_e=pb.lbx_e, ubx_e=pb.ubx_e)
    o = cp.unit(2, pd, 0, 300, pb.p_nominal, g["x0"])
    ok = (g["status"][:, 0] == 0) & (o["status"] == 0)
    assert ok.sum() >= len(ok) - 1
    assert np.abs(o["u0"] - g["u0"])[ok].max() < 1e-8
    assert (np.abs(o["cost"] - g["V"]) / np.abs(g["V"]))[ok].max() < 1e-10
    assert np.abs(o["dL"] - g["dV"][:, :3])[ok].max() < 1e-8 * np.abs(g["dV"]).max()
    assert np.abs(o["dpi"] - g["dpi"][:, :, :3])[ok].max() < 1e-6 * np.abs(g["dpi"]).max()
    assert min((g["lam"][i] > 1e-6).sum() for i in range(len(ok))) >= 15  # many active rows
    assert abs(np.abs(g["X"][7, :, 0]).max() - 2.4) < 1e-4  # sample 7 rides the cart-position bound


def test_linear_system_oracle_lqr_and_host_port():
    """Known answer: with gamma = 1 and the bounds out of reach the MPC is the LQR with
    the DARE terminal cost -- for the dense oracle and for the host port of the engine; then the
    committed fixtures (soft bound active in two of them) against the host port."""
    from scipy.linalg import solve_discrete_are

    from oracle import cpu_port as cp
    from oracle.problems import linear_system_param_nominal, make_linear_system
    from oracle.solver import DenseSolver

    par = linear_system_param_nominal()
    P = solve_discrete_are(par["A"], par["B"], par["Q"], par["R"])
    K = np.linalg.solve(par["R"] + par["B"].T @ P @ par["B"], par["B"].T @ P @ par["A"])
    assert np.abs(K.ravel() - [0.805778325799, 1.503607449412]).max() < 1e-10
    pbw = make_linear_system(gamma=1.0, lbx=(-1e3, -1e3), ubx=(1e3, 1e3))
    pbw.lbu = np.array([-1e3]); pbw.ubu = np.array([1e3])
    x0 = np.array([0.2, 0.2])
    sol, upd = DenseSolver(pbw).unit(x0, tol=1e-10)
    assert abs(sol.U[0, 0] + 0.461877155042) < 1e-8
    # the tau-central slacks of the 2 x 39 soft rows cost z * tau / z each: 78e-8 above the closed form
    assert abs(sol.cost - 1e-3 - 0.460894064457 - 78e-8) < 1e-9
    assert abs(upd["dL_dp"][0, 8] - 1.0) < 1e-12  # dV/dV_0
    mc = [P[0, 0], P[0, 1], P[1, 1]]
    pdw = cp.make_pd(40, np.ones(41), [-1e3], [1e3], mc, tol=1e-10, warm_ipm=1, lbx=[-1e3, -1e3], ubx=[1e3, 1e3],
                     zl=[1e2], zu=[1e2])
    o = cp.unit(3, pdw, 0, 20, pbw.p_nominal, x0[None, :], nx=2, nu=1)
    assert o["status"][0] == 0 and abs(o["u0"][0, 0] + 0.461877155042) < 1e-9
    # fixtures
    g = np.load(os.path.join(ROOT, "tests", "golden", "linear_system.npz"))
    pb = make_linear_system(gamma=float(g["gamma"]))
    scale = np.array([pb.stage_scale(k) for k in range(pb.N + 1)])
    pd = cp.make_pd(pb.N, scale, pb.lbu, pb.ubu, mc, tol=1e-10, warm_ipm=1, lbx=pb.lbx, ubx=pb.ubx, zl=pb.zl, zu=pb.zu)
    o = cp.unit(3, pd, 0, 100, pb.p_nominal, g["x0"], nx=2, nu=1)
    assert (o["status"] == 0).all() and (g["status"][:, 0] == 0).all()
    assert np.abs(o["u0"] - g["u0"]).max() < 1e-8
    assert (np.abs(o["cost"] - g["V"]) / np.abs(g["V"])).max() < 1e-9
    assert np.abs(o["dL"] - g["dV"]).max() < 1e-6 * np.abs(g["dV"]).max()
    soft = g["slmax"] > 1e-6
    assert soft.sum() >= 2 and np.abs(o["dpi"] - g["dpi"])[~soft].max() < 1e-5 * np.abs(g["dpi"]).max()


def test_evaporation_host_port_matches_oracle_fixtures():
    """Evaporation process: host port of the engine against the dense-oracle fixtures (N=40 set and the one
    full-horizon N=100 sample; u0, V, dV/dtheta and dpi/dtheta over the 60 tracking-cost parameters), plus
    a live short-horizon oracle solve."""
    from oracle import cpu_port as cp
    from oracle.problems import EVAPORATION_PARAM, make_evaporation
    from oracle.solver import DenseSolver

    def port(pb, x0s, gamma):
        scale = np.array([pb.stage_scale(k) for k in range(pb.N + 1)])
        mc = list(EVAPORATION_PARAM.values()) + [0.25, 4]
        pd = cp.make_pd(pb.N, scale, pb.lbu, pb.ubu, mc, tol=1e-9, warm_ipm=1, lg=pb.lh, ug=pb.uh)
        B, nx, nu, N = len(x0s), 2, 3, pb.N
        it = np.zeros((cp.lib().cpu_port_iterate_size(4, N), B))
        for k in range(N + 1):
            it[k * nx:(k + 1) * nx, :] = pb.x_init[:, None]
        for k in range(N):
            it[(N + 1) * nx + k * nu:(N + 1) * nx + (k + 1) * nu, :] = pb.u_init[:, None]
        return cp.unit(4, pd, 0, 100, pb.p_nominal, x0s, iterate=it, nx=2, nu=3)

    g = np.load(os.path.join(ROOT, "tests", "golden", "evaporation.npz"))
    pb = make_evaporation(gamma=float(g["gamma"]), N=int(g["N"]))
    o = port(pb, g["x0"], float(g["gamma"]))
    ok = (g["status"][:, 0] == 0) & (o["status"] == 0)
    assert ok.sum() >= len(ok) - 1
    assert np.abs(o["u0"] - g["u0"])[ok].max() < 1e-7
    assert (np.abs(o["cost"] - g["V"]) / np.abs(g["V"]))[ok].max() < 1e-10
    assert np.abs(o["dL"] - g["dV"])[ok].max() < 1e-7 * np.abs(g["dV"]).max()
    assert np.abs(o["dpi"] - g["dpi"])[ok].max() < 1e-6 * np.abs(g["dpi"]).max()
    g1 = np.load(os.path.join(ROOT, "tests", "golden", "evaporation_n100.npz"))
    o1 = port(make_evaporation(gamma=float(g1["gamma"]), N=100), g1["x0"][None, :], float(g1["gamma"]))
    assert int(g1["status"]) == 0 and o1["status"][0] == 0
    assert np.abs(o1["u0"][0] - g1["u0"]).max() < 1e-7 and abs(o1["cost"][0] - float(g1["V"])) < 1e-10 * abs(float(g1["V"]))
    assert np.abs(o1["dpi"][0] - g1["dpi"]).max() < 1e-6 * np.abs(g1["dpi"]).max()
    # live: N = 10 horizon, the oracle finishes in seconds
    pbs = make_evaporation(gamma=0.95, N=10)
    x0 = np.array([[30.0, 60.0]])
    sol, upd = DenseSolver(pbs).unit(x0[0], tol=1e-9)
    os_ = port(pbs, x0, 0.95)
    assert sol.status == 0 and os_["status"][0] == 0
    assert np.abs(sol.U[0] - os_["u0"][0]).max() < 1e-8 and abs(sol.cost - os_["cost"][0]) < 1e-9 * abs(sol.cost)
    assert np.abs(upd["dpi_dp"] - os_["dpi"][0]).max() < 1e-6 * np.abs(upd["dpi_dp"]).max()


# ------------------------------------------------------------------------------------------------
# chain of masses: oracle and fixture only -- the CUDA path for nx = 21 / 27 blocks
# is the next round's work; these tests pin the specification it has to meet.
# ------------------------------------------------------------------------------------------------
def test_chain_mass_layout_matches_reference_counts():
    from oracle.nlp import RestatedNLP
    from oracle.problems import chain_param_layout, make_chain_mass

    # define_nx_nu / define_param_struct_symSX (ocp_utils.py:344-371): n_mass = 3, 5, 6
    assert [chain_param_layout(n)[1] for n in (3, 5, 6)] == [113, 499, 800]
    pb = make_chain_mass(5)
    nlp = RestatedNLP(pb)
    # chain n_mass=5: nw=981, npi=840, nlam=282 => nz=2385, ntheta=499
    assert (pb.nx, pb.nu, nlp.nw, nlp.npi, nlp.nlam, nlp.nz, pb.ntheta) == (21, 3, 981, 840, 282, 2385, 499)
    sl, _ = chain_param_layout(5)
    Q = pb.p_nominal[sl["Q"]].reshape(21, 21).T
    assert np.array_equal(np.diag(Q), 2.0 * np.array([1.0] * 9 + [4.0] * 3 + [1.0] * 9))  # ocp_utils.py:263-266
    assert np.array_equal(pb.p_nominal[sl["R"]].reshape(3, 3), 0.02 * np.eye(3))
    # steady state: at rest, last mass at x_end, force balance on the intermediate masses (ocp_utils.py:150-192)
    import torch

    from oracle.problems import chain_ode

    f = chain_ode(torch.as_tensor(pb.x_ss), torch.zeros(3, dtype=torch.float64), torch.as_tensor(pb.p_nominal), 5)
    assert float(f.abs().max()) < 1e-12 and np.allclose(pb.x_ss[9:12], [0.033 * 4 * 6, 0.0, 0.0])


def test_chain_mass_policy_gradient_against_parameter_sweep():
    """What the reference's own test does (tests/test_chain_mass.py -> examples/chain_mass.py:main_nlp): sweep the
    damping C_{M}_0 of the last link, solve from define_x0, and compare d pi / d p from update_nlp with the
    numerical gradient of u_opt along the sweep -- here n_mass = 3 and central differences around the fixture."""
    from oracle.problems import chain_param_layout, make_chain_mass
    from oracle.solver import DenseSolver

    g = np.load(os.path.join(ROOT, "tests", "golden", "chain_mass_3.npz"))
    pb = make_chain_mass(3)
    assert np.array_equal(pb.p_nominal, g["theta"]) and np.allclose(pb.x_ss, g["x_ss"], atol=1e-13)
    s = DenseSolver(pb)
    sl, _ = chain_param_layout(3)
    M = 1
    idx = sl["C"].start + 3 * M + 0  # C_{M}_0
    d = 1e-4 * pb.p_nominal[idx]
    us, Vs = [], []
    for sgn in (+1.0, -1.0):
        p = pb.p_nominal.copy()
        p[idx] += sgn * d
        sol = s.solve(g["x0"][0], p=p, tol=1e-10, init=(g["U"][0], g["X"][0]))
        assert sol.status == 0
        us.append(sol.U[0]); Vs.append(sol.cost)
    fd_pi = (us[0] - us[1]) / (2 * d)
    fd_V = (Vs[0] - Vs[1]) / (2 * d)
    assert abs(g["u0"][0][2] - 1.0) < 1e-6  # third input sits on its bound: its sensitivity is the IPM-smoothed ~0
    assert np.abs(fd_pi - g["dpi"][0][:, idx]).max() < 1e-5 * max(1.0, np.abs(g["dpi"][0][:, idx]).max())
    assert abs(fd_V - g["dV"][0][idx]) < 1e-6 * max(1.0, abs(g["dV"][0][idx]))
    # Q-mode fixes u_0; V <= Q
    assert (g["status"] == 0).all() and (g["Q"] >= g["V"] - 1e-9).all()
