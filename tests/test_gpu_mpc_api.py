"""The reference-facing surface (rlmpc.mpc.common.mpc.MPC + AcadosMPC + ocp_solver) on the GPU.
These read like the reference's own drivers: scripts/cartpole_mpc_sensitivities.py:79-98,
scripts/cartpole_mpc_kkt_conditions.py:74-85, scripts/linear_system_mpc_nlp.py:17-106."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mpc():
    from mpc4rl_b200 import cartpole_original_config
    from mpc4rl_b200.mpc.cartpole.acados import AcadosMPC

    return AcadosMPC(cartpole_original_config(), build=True)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cartpole_original.npz"))


def test_sensitivity_script_point(mpc, golden):
    """x0=[0,0,pi/2,0], u0=-30 (cartpole_mpc_sensitivities.py:80-81): golden sample 0."""
    x0, u0 = golden["x0"][0], golden["a"][0]
    mpc.reset(x0)
    assert mpc.q_update(x0=x0, u0=u0) == 0
    Q, dQ = mpc.get_Q(), mpc.get_dQ_dp()
    assert dQ.shape == (1, 83)
    assert abs(Q - golden["Q"][0]) < 1e-6 * abs(golden["Q"][0])
    assert np.allclose(dQ[0], golden["dQ"][0], rtol=1e-4, atol=1e-5)
    assert np.abs(mpc.get_dpi_dp()).max() == 0.0  # Q-mode: u_0 clamped
    assert mpc.update(x0=x0) == 0
    V, pi = mpc.get_V(), mpc.get_pi()
    assert abs(V - golden["V"][0]) < 1e-6 * abs(golden["V"][0])
    assert pi.shape == (1,) and abs(pi[0] - golden["u0"][0, 0]) < 1e-5  # north-star |du0| < 1e-5
    # quirk Q6: update() does not refresh the sensitivities -> still the Q-mode ones
    assert np.array_equal(mpc.get_dV_dp(), dQ)
    mpc.update_nlp()
    assert np.allclose(mpc.get_dV_dp()[0], golden["dV"][0], rtol=1e-4, atol=1e-5)
    dpi = mpc.get_dpi_dp()
    assert dpi.shape == (1, 83) and np.allclose(dpi, golden["dpi"][0], rtol=1e-3, atol=1e-4)
    assert set(mpc.nlp_timing) >= {"dL_dp", "lin_params", "solve_params"}
    assert mpc.nlp.assert_kkt_residual()  # cartpole_mpc_kkt_conditions.py:81
    assert mpc.get_V() <= Q + 1e-9


def test_ocp_solver_shim_fields(mpc, golden):
    x0 = golden["x0"][1]
    mpc.reset(x0)
    mpc.update(x0)
    s = mpc.ocp_solver
    N = s.acados_ocp.dims.N
    X = np.stack([s.get(k, "x") for k in range(N + 1)])
    U = np.stack([s.get(k, "u") for k in range(N)])
    PI = np.concatenate([s.get(k, "pi") for k in range(N)])
    assert np.abs(X - golden["X"][1]).max() < 1e-5 and np.abs(U - golden["U"][1]).max() < 1e-5
    assert np.abs(PI - golden["pi"][1]).max() < 1e-4 * np.abs(golden["pi"][1]).max()
    # multipliers in acados order: stage 0 has 2*(nu+nx) rows, stages 1..N-1 2*nu, stage N none
    mpc.update_nlp()
    assert s.get(0, "lam").shape == (10,) and s.get(1, "lam").shape == (2,) and s.get(N, "lam").shape == (0,)
    lam = np.concatenate([s.get(k, "lam") for k in range(N + 1)])
    assert lam.shape == golden["lam"][1].shape and (lam >= 0).all()
    g = golden["lam"][1].copy()
    eq_rows = np.r_[1:5, 6:10]  # the oracle parks tau/t_eq = 100 on the eliminated x0 rows
    g[eq_rows] -= 100.0
    assert np.allclose(lam, g, rtol=1e-4, atol=1e-6)
    assert np.allclose(s.get_residuals(), 0.0, atol=1e-6)
    assert abs(s.get_cost() - golden["V"][1]) < 1e-6 * golden["V"][1]


def test_get_action_is_scaled_and_labels(mpc, golden):
    a = mpc.get_action(golden["x0"][1])
    assert a.shape == (1,) and abs(a[0] - golden["u0"][1, 0] / 80.0) < 1e-6
    assert np.allclose(mpc.unscale_action(a), golden["u0"][1], atol=1e-4)
    assert mpc.get_parameter_labels() == ["M", "m", "l"]
    assert mpc.get_state_labels() == ["x", "x_dot", "theta", "theta_dot"] and mpc.get_input_labels() == ["F"]
    assert mpc.get_p().shape == (83,) and mpc.ocp_solver.acados_ocp.dims.np == 3


def test_nonzero_status_raises_like_the_reference(golden):
    from mpc4rl_b200 import cartpole_original_config
    from mpc4rl_b200.mpc.cartpole.acados import AcadosMPC

    cfg = cartpole_original_config()
    cfg["ocp_options"]["nlp_solver_max_iter"] = 2
    m = AcadosMPC(cfg)
    x0 = golden["x0"][0]
    m.reset(x0)
    with pytest.raises(RuntimeError, match="Solver failed update with status 2"):
        m.update(x0)
    with pytest.raises(RuntimeError, match="Solver failed q_update with status 2"):
        m.reset(x0)
        m.q_update(x0, golden["a"][0])


def test_value_gradient_vs_parameter_sweep(mpc, golden):
    """test_acados_ocp_nlp of the reference (linear_system_mpc_nlp.py:17-62, run on the cartpole by
    cartpole_mpc_sensitivities.py:232-239): np.gradient(V) over a +-10% sweep vs dV_dp, atol 1e-1."""
    x0 = golden["x0"][1]
    p_nom = mpc.get_p()
    for i_param in range(mpc.ocp_solver.acados_ocp.dims.np):
        sweep = np.linspace(0.9 * p_nom[i_param], 1.1 * p_nom[i_param], 21)
        V, dV = [], []
        mpc.reset(x0)
        for v in sweep:
            p = p_nom.copy(); p[i_param] = v
            mpc.set_p(p)
            mpc.update(x0)
            mpc.update_nlp()
            V.append(mpc.get_V()); dV.append(mpc.get_dV_dp()[0, i_param])
        fd = np.gradient(np.array(V), sweep[1] - sweep[0])
        assert np.allclose(fd[1:-1], np.array(dV)[1:-1], rtol=2e-2, atol=1e-1)
    mpc.set_p(p_nom)


def test_finite_difference_policy_gradient(mpc, golden):
    """get_dpi_dp(finite_differences=True) (mpc.py:353-414) against the adjoint result."""
    x0 = golden["x0"][0]  # interior u0
    mpc.set_p(golden["theta"])
    mpc.reset(x0)
    mpc.update(x0)
    mpc.update_nlp()
    an = mpc.get_dpi_dp().copy()
    for idx in range(3):
        mpc.update(x0)  # like the reference, the FD helper measures from the solver's current solution
        fd = mpc.get_dpi_dp(finite_differences=True, idx=idx)
        assert abs(fd[0, idx] - an[0, idx]) < 1e-2 * max(1.0, abs(an[0, idx]))


def test_store_and_load_iterate(mpc, golden, tmp_path):
    x0 = golden["x0"][1]
    mpc.reset(x0)
    mpc.update(x0)
    u = mpc.get_pi().copy()
    f = str(tmp_path / "iterate.json")
    mpc.ocp_solver.store_iterate(filename=f, overwrite=True, verbose=False)
    mpc.reset(golden["x0"][2])
    mpc.ocp_solver.load_iterate(f)
    assert np.array_equal(mpc.ocp_solver.get(0, "u"), u)
