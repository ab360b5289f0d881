"""GPU parity tests for the other problem definitions the engine carries -- through the C ABI:

* cartpole with state bounds (config/cartpole.yaml: N=30, |u|<=30, box on all four states),
* the linear system of the reference's usage example / pytest (rlmpc/mpc/linear_system/acados.py,
  tests/test_linear_example.py): EXTERNAL cost, discount factor, soft state bound with slack rows.

Golden fixtures: oracle/make_golden.py (dense restatement; parity unpinned, see oracle/__init__.py).
Tolerances as in test_gpu_cartpole.py.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device="cuda:0")


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


# ------------------------------------------------------------------------------------------------
# cartpole.yaml (state bounds)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def g_bx():
    return np.load(os.path.join(ROOT, "tests", "golden", "cartpole_default.npz"))


def _bx_engine(B):
    from mpc4rl_b200 import BatchedMPC, cartpole_config, cartpole_spec

    m = BatchedMPC(cartpole_spec(cartpole_config()), max_batch=B, device=0)
    m.set_option("tol", 1e-10)
    return m


def test_state_bounds_v_and_q_match_golden(g_bx):
    x0 = _dev(g_bx["x0"])
    B = x0.shape[0]
    m = _bx_engine(B)
    assert m.nbx == 4 and m.nrows == 10  # [lbu, lbx(4), ubu, ubx(4)]
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=300)
    st = out["status"].cpu().numpy()
    ok = (g_bx["status"][:, 0] == 0) & (st == 0)
    assert ok.sum() >= B - 1
    assert np.abs(out["u0"].cpu().numpy() - g_bx["u0"])[ok].max() < 1e-6
    assert _rel(out["cost"].cpu().numpy()[ok], g_bx["V"][ok]) < 1e-9
    assert _rel(m.full_grad(out["dL"]).cpu().numpy()[ok], g_bx["dV"][ok]) < 1e-6
    assert np.abs(m.full_grad(out["dpi"]).cpu().numpy() - g_bx["dpi"])[ok].max() < 1e-5 * max(1.0, np.abs(g_bx["dpi"]).max())
    assert out["res"][torch.tensor(ok, device="cuda:0")].max().item() < 1e-9
    # 20-35 active rows per solution (mostly the saturated input); sample 7 rides the cart-position bound
    X = np.stack([m.get("x", k, B).cpu().numpy() for k in range(31)], axis=1)
    assert np.abs(X - g_bx["X"])[ok].max() < 1e-6
    assert ok[7] and abs(np.abs(X[7, :, 0]).max() - 2.4) < 1e-4  # tau-central: slack tau/lam below the bound
    # Q-mode
    a = _dev(g_bx["a"])
    m.reset(x0)
    oq = m.solve_sens(x0, a, max_sqp=300)
    okq = (g_bx["status"][:, 1] == 0) & (oq["status"].cpu().numpy() == 0)
    assert okq.sum() >= B - 1
    assert _rel(oq["cost"].cpu().numpy()[okq], g_bx["Q"][okq]) < 1e-9
    assert _rel(m.full_grad(oq["dL"]).cpu().numpy()[okq], g_bx["dQ"][okq]) < 1e-6


def test_state_bounds_through_the_mirrored_api(g_bx):
    from mpc4rl_b200 import cartpole_config
    from mpc4rl_b200.mpc.cartpole.acados import AcadosMPC

    mpc = AcadosMPC(config=cartpole_config(), build=True)
    x0 = g_bx["x0"][0]  # config/cartpole.yaml x0 = [0, 0, 3.14, 0]
    mpc.reset(x0)
    mpc.update(x0)
    mpc.update_nlp()
    assert abs(mpc.get_V() - g_bx["V"][0]) < 1e-6 * abs(g_bx["V"][0])
    assert np.abs(mpc.get_pi() - g_bx["u0"][0]).max() < 1e-5
    lam1 = mpc.ocp_solver.get(1, "lam")  # acados order [lbu, lbx(4), ubu, ubx(4)]
    assert lam1.shape == (10,)
    assert mpc.ocp_solver.get(30, "lam").shape == (8,)  # terminal: [lbx_e(4), ubx_e(4)]
    assert mpc.ocp_solver.get(0, "lam").shape == (10,)  # stage 0: lbu, lbx_0 (all nx), ubu, ubx_0
    mpc.nlp.assert_kkt_residual(1e-6)  # scripts/cartpole_mpc_kkt_conditions.py:74-85
    # infeasible first QP (pendulum horizontal: no input authority in the linearisation) -> acados status 4
    mpc.reset(np.array([0.0, 0.0, np.pi / 2, 0.0]))
    with pytest.raises(RuntimeError, match="status 4"):
        mpc.update(np.array([0.0, 0.0, np.pi / 2, 0.0]))


# ------------------------------------------------------------------------------------------------
# linear system
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def g_lin():
    return np.load(os.path.join(ROOT, "tests", "golden", "linear_system.npz"))


def _lin_engine(B, **kw):
    from mpc4rl_b200 import BatchedMPC, linear_system_spec

    m = BatchedMPC(linear_system_spec(**kw), max_batch=B, device=0)
    m.set_option("tol", 1e-10)
    return m


def test_linear_system_matches_golden(g_lin):
    x0, a = _dev(g_lin["x0"]), _dev(g_lin["a"])
    B = x0.shape[0]
    m = _lin_engine(B, gamma=float(g_lin["gamma"]))
    assert m.ngrad == 12 and m.nrows == 8  # [lbu, lbx(2), ubu, ubx(2), lsbx, usbx]
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=100)
    ok = (g_lin["status"][:, 0] == 0) & (out["status"].cpu().numpy() == 0)
    assert ok.all()
    assert np.abs(out["u0"].cpu().numpy() - g_lin["u0"]).max() < 1e-8
    assert _rel(out["cost"].cpu().numpy(), g_lin["V"]) < 1e-9
    assert _rel(out["dL"].cpu().numpy(), g_lin["dV"]) < 1e-6
    assert np.abs(out["dL"][:, 8].cpu().numpy() - 1.0).max() < 1e-12  # dV/dV_0 = 1 exactly
    # dpi/dp: where a softened bound is violated the reference's formula (slack frozen, quirk Q4; x_0
    # "fixed" by two barrier rows of finite stiffness, quirk Q7) is a ratio of IPM-sized stiffnesses --
    # compared only on the samples without an active slack
    soft = g_lin["slmax"] > 1e-6
    assert soft.sum() >= 2
    assert np.abs(out["dpi"].cpu().numpy() - g_lin["dpi"])[~soft].max() < 1e-5 * np.abs(g_lin["dpi"]).max()
    # slack values of the violated soft bound (ocp_solver.get(stage, "su"))
    T = np.stack([m.get("t", k, B).cpu().numpy() for k in range(1, 40)], axis=1)  # [B, 39, 8]
    assert abs(T[3, :, 7].max() - 0.21875) < 1e-6  # x0 = [0.9, 0.8]: upper bound on x[0] exceeded by 0.21875
    # Q-mode
    m.reset(x0)
    oq = m.solve_sens(x0, a, max_sqp=100)
    okq = (g_lin["status"][:, 1] == 0) & (oq["status"].cpu().numpy() == 0)
    assert okq.sum() >= B - 1  # one (s, a) pair makes the hard bound on x[1] infeasible
    assert _rel(oq["cost"].cpu().numpy()[okq], g_lin["Q"][okq]) < 1e-9
    assert _rel(oq["dL"].cpu().numpy()[okq], g_lin["dQ"][okq]) < 1e-6


def test_linear_system_lqr_known_answer():
    """gamma = 1 and bounds out of reach: the MPC is the LQR (terminal cost = DARE solution), closed form
    from SURVEY.md 8(c): x0=[0.2,0.2] -> u0=-0.461877155042, V-V_0=0.460894064457; [0.5,-0.2] -> ..."""
    m = _lin_engine(2, gamma=1.0, lbx=(-1e3, -1e3), ubx=(1e3, 1e3), lbu=(-1e3,), ubu=(1e3,))
    x0 = _dev([[0.2, 0.2], [0.5, -0.2]])
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=20)
    assert (out["status"] == 0).all()
    assert np.abs(out["u0"].cpu().numpy().ravel() - [-0.461877155042, -0.102167673017]).max() < 1e-9
    # + 78e-8: the tau-central slacks of the 2 x 39 soft rows (z * tau / z each)
    assert np.abs(out["cost"].cpu().numpy() - 1e-3 - 78e-8 - [0.460894064457, 0.680269101675]).max() < 1e-8
    assert np.isfinite(out["dpi"].cpu().numpy()).all()


def test_linear_example_construction_like_the_reference_pytest():
    """tests/test_linear_example.py:8-26 of the reference: construct and touch the attributes."""
    from mpc4rl_b200.mpc.linear_system.acados import AcadosMPC
    from mpc4rl_b200.problems import linear_system_param_nominal

    mpc = AcadosMPC(linear_system_param_nominal(), discount_factor=0.99)
    assert mpc is not None
    assert mpc.ocp_solver is not None
    assert mpc.nlp is not None
    assert mpc.ocp_solver.acados_ocp is not None
    assert mpc.ocp_solver.acados_ocp.model is not None
    assert mpc.ocp_solver.acados_ocp.dims is not None
    assert mpc.ocp_solver.acados_ocp.cost is not None
    assert mpc.ocp_solver.acados_ocp.dims.N == 40 and mpc.get_p().shape == (12,)
    assert mpc.get_parameter_labels()[8] == "V_0"
    x0 = np.array([0.5, 0.5])
    mpc.reset(x0)
    mpc.q_update(x0, np.array([0.3]))
    assert mpc.get_dQ_dp().shape == (1, 12) and abs(mpc.get_dQ_dp()[0, 8] - 1.0) < 1e-12
    mpc.update(x0)
    mpc.update_nlp()
    assert mpc.get_dpi_dp().shape == (1, 12)
    assert mpc.ocp_solver.get(1, "lam").shape == (8,) and mpc.ocp_solver.get(1, "su").shape == (1,)


def test_qlearning_example_batched_equals_per_sample_loop():
    """The usage example of the reference: the batched learning step gives the parameter update of the
    per-sample q_update/update loop (examples/linear_system_mpc_qlearning.py:171-205)."""
    from mpc4rl_b200.examples import linear_system_mpc_qlearning as ex
    from mpc4rl_b200.mpc.linear_system.acados import AcadosMPC
    from mpc4rl_b200.problems import linear_system_param_nominal

    mpc = AcadosMPC(linear_system_param_nominal(), discount_factor=ex.GAMMA)
    env = ex.LinearSystemEnv(mpc.ocp_solver.acados_ocp.constraints.lbx, mpc.ocp_solver.acados_ocp.constraints.ubx, seed=1)
    S, A, C = ex.rollout(mpc, env, episode_length=24)
    assert np.isfinite(C).all() and np.abs(A).max() <= 1.0 + 1e-9
    dp_loop, td_loop, q_loop, v_loop = ex.learn_loop(mpc, S, A, C)
    engine = mpc.batched(max_batch=32)
    engine.set_option("tol", 1e-6)
    dp_b, td_b, q_b, v_b = ex.learn_batched(engine, mpc, S, A, C)
    assert np.abs(q_b - q_loop).max() < 1e-6 * max(1.0, np.abs(q_loop).max())
    assert np.abs(v_b - v_loop).max() < 1e-6 * max(1.0, np.abs(v_loop).max())
    assert np.abs(td_b - td_loop).max() < 1e-5
    assert np.abs(dp_b - dp_loop).max() < 1e-6 * max(1e-6, np.abs(dp_loop).max()) + 1e-12
    p0 = mpc.get_parameter_values().copy()
    mpc.set_parameter(p0 + dp_b)
    assert np.abs(mpc.get_p() - (p0 + dp_b)).max() == 0.0
    log = ex.main(n_episodes=2, episode_length=16, verbose=False)
    assert len(log) == 2 and np.isfinite(log[-1]["td_error"])


# ------------------------------------------------------------------------------------------------
# evaporation process (affine h rows through the slack input, N = 100, theta = (W_0, W, yref_0, yref))
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def g_evap():
    return np.load(os.path.join(ROOT, "tests", "golden", "evaporation.npz"))


def _evap_engine(B, gamma, N=100):
    from mpc4rl_b200 import BatchedMPC, evaporation_spec

    spec = evaporation_spec(gamma=gamma, N=N)
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-9)
    return m, spec


def _evap_guess(m, spec, B):
    """The reference starts every stage at the steady state (evaporation_process/acados.py:104-109)."""
    m.reset(B=B)
    for k in range(spec.N + 1):
        m.put("x", k, _dev(np.tile(spec.x_init, (B, 1))))
    for k in range(spec.N):
        m.put("u", k, _dev(np.tile(spec.u_init, (B, 1))))


def test_evaporation_matches_golden(g_evap):
    x0, a = _dev(g_evap["x0"]), _dev(g_evap["a"])
    B = x0.shape[0]
    m, spec = _evap_engine(B, float(g_evap["gamma"]), int(g_evap["N"]))  # fixtures: N=40 (dense oracle cost)
    assert m.ngrad == 60 and m.nrows == 10 and m.ng == 2 and m.nbx == 0  # [lbu(3), lh(2), ubu(3), uh(2)]
    _evap_guess(m, spec, B)
    out = m.solve_sens(x0, max_sqp=100)
    ok = (g_evap["status"][:, 0] == 0) & (out["status"].cpu().numpy() == 0)
    assert ok.sum() >= B - 1
    assert np.abs(out["u0"].cpu().numpy() - g_evap["u0"])[ok].max() < 1e-6
    assert _rel(out["cost"].cpu().numpy()[ok], g_evap["V"][ok]) < 1e-9
    assert _rel(out["dL"].cpu().numpy()[ok], g_evap["dV"][ok]) < 1e-6     # d/d(W_0, W, yref_0, yref)
    assert _rel(out["dpi"].cpu().numpy()[ok], g_evap["dpi"][ok]) < 1e-5
    X = np.stack([m.get("x", k, B).cpu().numpy() for k in range(spec.N + 1)], axis=1)
    assert np.abs(X - g_evap["X"])[ok].max() < 1e-6
    # the soft constraint x + s >= 25 is active for the state that starts below it
    U = np.stack([m.get("u", k, B).cpu().numpy() for k in range(spec.N)], axis=1)
    assert U[2, 0, 2] > 0.5  # x0 = [24, 55]: slack input used at stage 0
    _evap_guess(m, spec, B)
    oq = m.solve_sens(x0, a, max_sqp=100)
    okq = (g_evap["status"][:, 1] == 0) & (oq["status"].cpu().numpy() == 0)
    assert okq.sum() >= B - 2
    assert _rel(oq["cost"].cpu().numpy()[okq], g_evap["Q"][okq]) < 1e-9
    assert _rel(oq["dL"].cpu().numpy()[okq], g_evap["dQ"][okq]) < 1e-6


def test_evaporation_through_the_mirrored_api():
    """The reference's class at its full horizon N=100 against the one N=100 oracle sample."""
    from mpc4rl_b200.mpc.evaporation_process.acados import AcadosMPC
    from mpc4rl_b200.problems import EVAPORATION_PARAM, H_NOMINAL

    g1 = np.load(os.path.join(ROOT, "tests", "golden", "evaporation_n100.npz"))
    g_evap = {k: (g1[k][None] if g1[k].ndim >= 1 and k != "theta" else g1[k]) for k in g1.files}
    g_evap["V"] = np.array([float(g1["V"])])
    gamma = float(g_evap["gamma"])
    mpc = AcadosMPC(model_param=EVAPORATION_PARAM, cost_param={"H": {"l": H_NOMINAL}}, gamma=gamma)
    assert mpc.get_p().shape == (60,) and mpc.ocp_solver.acados_ocp.dims.N == 100 and mpc.ocp_solver.acados_ocp.dims.nh == 2
    x0 = g_evap["x0"][0]
    u = mpc.get_action(x0)
    assert np.abs(u - g_evap["u0"][0]).max() < 1e-4  # default tol 1e-6
    mpc.reset()
    mpc.update(x0)
    mpc.update_nlp()
    assert abs(mpc.get_V() - g_evap["V"][0]) < 1e-7 * abs(g_evap["V"][0])
    assert mpc.get_dV_dp().shape == (1, 60) and mpc.get_dpi_dp().shape == (3, 60)
    assert _rel(mpc.get_dV_dp()[0], g_evap["dV"][0]) < 1e-5
    assert mpc.ocp_solver.get(1, "lam").shape == (10,) and mpc.ocp_solver.get(0, "lam").shape == (14,)
    assert mpc.ocp_solver.get(100, "lam").shape == (0,)
    # learning W through set_parameter keeps solver and NLP consistent (mpc.py:233-257)
    p = mpc.get_p()
    p[25] *= 1.1  # W[0,0]
    mpc.set_parameter(p)
    mpc.reset()
    mpc.update(x0)
    assert mpc.get_V() > g_evap["V"][0]


# ------------------------------------------------------------------------------------------------
# full-size property tests (BASELINE.json sizes; the oracle is too slow there)
# ------------------------------------------------------------------------------------------------
def test_evaporation_full_batch_properties():
    """BASELINE configs[3] size (32 768 samples, N=100): every sample converges from the steady-state guess,
    KKT residuals at tolerance, replicated inputs give bit-identical outputs, the constraint x + s >= 25 holds
    along the horizon, and dV/dyref agrees with central differences through per-sample theta."""
    from mpc4rl_b200 import BatchedMPC, evaporation_spec

    B = 32768
    spec = evaporation_spec(gamma=0.99)
    g = torch.Generator(device="cpu").manual_seed(7)
    lo, hi = torch.tensor([25.0, 49.7], dtype=torch.float64), torch.tensor([40.0, 70.0], dtype=torch.float64)
    x0 = (lo + (hi - lo) * torch.rand(B, 2, generator=g, dtype=torch.float64)).cuda()
    x0[B // 2:] = x0[: B // 2]
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-8)
    _evap_guess(m, spec, B)
    out = m.solve_sens(x0, max_sqp=100)
    st = out["status"].cpu().numpy()
    assert (st == 0).mean() > 0.999, np.bincount(st)
    okm = out["status"] == 0
    assert out["res"][okm].max().item() < 1e-8 * (1 + 1e-3)
    for k in ("u0", "cost", "dL", "dpi"):
        assert torch.equal(out[k][: B // 2], out[k][B // 2:]), k
    for k in (0, 1, 50, 99):
        x, u = m.get("x", k, B), m.get("u", k, B)
        assert ((x + u[:, 2:3])[okm] >= 25.0 - 1e-6).all()
        assert ((u >= torch.tensor(spec.lbu, device="cuda:0") - 1e-9) & (u <= torch.tensor(spec.ubu, device="cuda:0") + 1e-9))[okm].all()
    # FD of V w.r.t. yref[0] (theta index 55) and W[0,0] (index 25) through per-sample theta
    n, d = 32, 1e-5
    idx = [25, 55]
    th = np.tile(spec.p_nominal, (2 * len(idx) * n, 1))
    for c, j in enumerate(idx):
        th[(2 * c) * n:(2 * c + 1) * n, j] += d
        th[(2 * c + 1) * n:(2 * c + 2) * n, j] -= d
    m2 = BatchedMPC(spec, max_batch=th.shape[0], device=0)
    m2.set_option("tol", 1e-11)
    m2.set_theta(th)
    xx = x0[:n].repeat(2 * len(idx), 1)
    _evap_guess(m2, spec, th.shape[0])
    o2 = m2.solve_sens(xx, max_sqp=200)
    assert (o2["status"] == 0).all()
    V = o2["cost"].cpu().numpy().reshape(2 * len(idx), n)
    an = out["dL"][:n].cpu().numpy()
    for c, j in enumerate(idx):
        fd = (V[2 * c] - V[2 * c + 1]) / (2 * d)
        assert np.abs(fd - an[:, j]).max() < 1e-4 * max(1.0, np.abs(an[:, j]).max()), j


def test_linear_system_batch_properties():
    """4 096 random states of the linear system in one call: converged, |u| <= 1, the hard bound on x[1] holds,
    V >= V_0, V is (weakly) convex along a line of states, dV/dV_0 = 1 everywhere."""
    B = 4096
    m = _lin_engine(B, gamma=0.9)
    m.set_option("tol", 1e-8)
    g = torch.Generator(device="cpu").manual_seed(3)
    x0 = torch.rand(B, 2, generator=g, dtype=torch.float64)
    x0[:, 1] = 1.6 * x0[:, 1] - 0.8
    t = torch.linspace(0.0, 1.0, 33, dtype=torch.float64)
    x0[:33] = torch.stack([0.1 + 0.7 * t, 0.3 - 0.5 * t], dim=1)  # a line segment for the convexity check
    x0 = x0.cuda()
    m.reset(x0)
    out = m.solve_sens(x0, max_sqp=100)
    okm = out["status"] == 0
    assert okm.double().mean().item() > 0.98
    assert out["res"][okm].max().item() < 1e-8 * (1 + 1e-3)
    assert (out["u0"][okm].abs() <= 1.0 + 1e-8).all()
    assert (out["cost"][okm] >= 1e-3 - 1e-12).all()
    assert (out["dL"][okm][:, 8] - 1.0).abs().max().item() < 1e-12
    x1 = m.get("x", 1, B)
    assert (x1[okm][:, 1].abs() <= 1.0 + 1e-7).all()
    V = out["cost"][:33].cpu().numpy()
    assert okm[:33].all() and (V[:-2] + V[2:] - 2 * V[1:-1] >= -1e-9).all()


# ------------------------------------------------------------------------------------------------
# warp-per-sample queue kernel, general version (csrc/coop_general.cuh) against the thread-per-sample kernels
# ------------------------------------------------------------------------------------------------
def _compare_queue_kernels(make, x0, starts, steps, u0=None, min_ok=0.6):
    """Same calls with option coop = 1 (default) and 0; statuses equal, outputs equal to rounding."""
    runs = {}
    for coop in (1, 0):
        m = make()
        m.set_option("tol", 1e-8)
        m.set_option("coop", coop)
        m.set_option("timing", 1)
        starts(m)
        outs = [m.solve_sens(x0, max_sqp=60)]
        used = False
        for dx in steps:
            outs.append(m.solve_sens(x0 + dx, max_sqp=1))
            used = used or m.timings()["queue_ipm_iters"] > 0  # (counted per SQP round; an RTI call has one)
        if u0 is not None:
            outs.append(m.solve_sens(x0, u0=u0, max_sqp=60))
        runs[coop] = ([{k: v.cpu().numpy() for k, v in o.items()} for o in outs], used)
    assert runs[1][1] and not runs[0][1]  # the warp-per-sample kernel ran with coop=1 and only then
    conv = None  # samples converged in the cold solve: only those start the later calls from the same iterate in
    for step, (a, b) in enumerate(zip(runs[1][0], runs[0][0])):  # both runs (an unconverged SQP amplifies rounding)
        same = a["status"] == b["status"]
        assert same.mean() > 0.99, (step, same.mean())
        ok = same & (a["status"] == 0)
        conv = ok if conv is None else conv
        ok = ok & conv
        assert ok.mean() > min_ok, (step, ok.mean())
        assert np.abs(a["u0"] - b["u0"])[ok].max() < 1e-6 * max(1.0, np.abs(b["u0"][ok]).max()), step
        assert _rel(a["cost"][ok], b["cost"][ok]) < 1e-8, step
        assert _rel(a["dL"][ok], b["dL"][ok]) < 1e-5, step


def test_general_warp_per_sample_kernel_state_bounds():
    B = 2048
    g = torch.Generator(device="cpu").manual_seed(5)
    lo = torch.tensor([-1.0, -1.0, -0.5 * np.pi, -2.0], dtype=torch.float64)
    x0 = (lo + (-2 * lo) * torch.rand(B, 4, generator=g, dtype=torch.float64)).cuda()
    dx = [1e-3 * torch.randn(B, 4, generator=g, dtype=torch.float64).cuda() for _ in range(2)]
    a0 = (-30.0 + 60.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)).cuda()
    # random swing-up starts against |u| <= 30 and the state box: about half of them converge (the rest: status 2 / 4, equal in both paths)
    _compare_queue_kernels(lambda: _bx_engine(B), x0, lambda m: m.reset(x0), dx, u0=a0, min_ok=0.3)


def test_general_warp_per_sample_kernel_soft_bounds():
    B = 2048
    g = torch.Generator(device="cpu").manual_seed(6)
    x0 = (torch.tensor([-0.2, -1.0]) + torch.tensor([1.4, 2.0]) * torch.rand(B, 2, generator=g)).double().cuda()
    dx = [1e-2 * torch.randn(B, 2, generator=g, dtype=torch.float64).cuda() for _ in range(2)]
    a0 = (-1.0 + 2.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)).cuda()
    _compare_queue_kernels(lambda: _lin_engine(B), x0, lambda m: m.reset(x0), dx, u0=a0, min_ok=0.9)


def test_general_warp_per_sample_kernel_general_rows():
    B = 1024
    g = torch.Generator(device="cpu").manual_seed(8)
    lo, hi = torch.tensor([25.0, 49.7], dtype=torch.float64), torch.tensor([40.0, 70.0], dtype=torch.float64)
    x0 = (lo + (hi - lo) * torch.rand(B, 2, generator=g, dtype=torch.float64)).cuda()
    dx = [1e-2 * torch.randn(B, 2, generator=g, dtype=torch.float64).cuda() for _ in range(2)]
    spec_box = {}

    def make():
        m, spec = _evap_engine(B, 0.99)
        spec_box["spec"] = spec
        return m

    _compare_queue_kernels(make, x0, lambda m: _evap_guess(m, spec_box["spec"], B), dx, min_ok=0.95)
