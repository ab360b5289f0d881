"""Host builds of the engines (oracle/cpu_port, test infrastructure) against the dense oracle's fixtures: the same
templates the CUDA kernels instantiate, run without a GPU.  These pin the MATHS of paths whose GPU parity tests need
a device: the cost-parameter columns of dpi/dtheta (parameterize_tracking_cost) and the warp-cooperative chain-mass
engine under the fiber emulation of a warp."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cartpole_cost_parameter_columns_match_oracle():
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec
    from oracle import cpu_port as cp

    g = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_paramcost.npz"))
    spec = cartpole_spec(cartpole_original_config())
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-10, warm_ipm=1, param_cost=1)
    o = cp.unit(1, pd, 0, 200, g["theta"], g["x0"])
    ok = (o["status"] == 0) & (g["status"][:, 0] == 0)
    assert ok.sum() >= 4
    assert o["dL"].shape[1] == 83 and o["dpi"].shape[2] == 83
    assert np.abs(o["u0"] - g["u0"])[ok].max() < 1e-8
    assert np.abs(o["dL"] - g["dV"])[ok].max() < 1e-6 * np.abs(g["dV"][ok]).max()
    assert np.abs(g["dpi"][ok][:, :, 3:]).max() > 1.0  # the cost columns are live ...
    assert np.abs(o["dpi"] - g["dpi"])[ok].max() < 1e-5 * np.abs(g["dpi"][ok]).max()  # ... and agree


def test_cartpole_free_g_matches_oracle():
    """g un-fixed (scripts/cartpole_mpc_qlearning.py:184-187): theta has 84 entries, 4 of them model parameters."""
    import copy

    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec
    from oracle import cpu_port as cp

    g = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_free_g.npz"))
    cfg = copy.deepcopy(cartpole_original_config())
    cfg["model"]["params"]["g"]["fixed"] = False
    spec = cartpole_spec(cfg)
    assert spec.ntheta == 84 and np.array_equal(spec.p_nominal, g["theta"])
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-10, warm_ipm=1)
    o = cp.unit(5, pd, 0, 200, g["theta"], g["x0"])
    ok = (o["status"] == 0) & (g["status"][:, 0] == 0)
    assert ok.sum() >= 4
    assert o["dL"].shape[1] == 4 and np.abs(g["dV"][ok][:, 3]).max() > 1e-3  # the g column is live
    assert np.abs(o["u0"] - g["u0"])[ok].max() < 1e-8
    assert np.abs(o["dL"] - g["dV"][:, :4])[ok].max() < 1e-6 * np.abs(g["dV"][ok]).max()
    assert np.abs(o["dpi"] - g["dpi"][:, :, :4])[ok].max() < 1e-5 * np.abs(g["dpi"][ok]).max()
    q = cp.unit(5, pd, 1, 200, g["theta"], g["x0"], u0=g["a"])
    okq = (q["status"] == 0) & (g["status"][:, 1] == 0)
    assert okq.sum() >= 4
    assert np.abs(q["cost"] - g["Q"])[okq].max() < 1e-9 * np.abs(g["Q"][okq]).max()
    assert np.abs(q["dL"] - g["dQ"][:, :4])[okq].max() < 1e-6 * np.abs(g["dQ"][okq]).max()


def test_interior_point_pass_state_machine_equals_the_monolithic_loop():
    """Engine::ipm_pass (one interior-point iteration per call, state carried outside: what the pass kernels of option
    ipm_passes run, and what k_qp3's pick-up of k_qp1's first iteration mirrors) against Engine::qp_ipm run to the end in
    one call: same iterates, statuses and outputs over closed-loop RTI steps -- also when the passes run out and the
    monolithic loop takes the sample over."""
    import sys

    sys.path.insert(0, ROOT)
    from bench import env_step_np, synth_states
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec
    from oracle import cpu_port as cp

    spec = cartpole_spec(cartpole_original_config())
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-6, warm_ipm=1)
    x0 = synth_states(256, 1234).numpy()
    runs = {}
    try:
        for passes in (0, 6, 2):
            cp.lib().cpu_port_set_passes(passes)
            o = cp.unit(1, pd, 0, 30, spec.p_nominal, x0)
            it, x1, outs = o["iterate"], x0, []
            for _ in range(3):
                x1 = env_step_np(x1, o["u0"][:, 0])
                o = cp.unit(1, pd, 0, 1, spec.p_nominal, x1, iterate=it)
                it = o["iterate"]
                outs.append(o)
            runs[passes] = outs
    finally:
        cp.lib().cpu_port_set_passes(0)
    assert (runs[0][0]["iters"][:, 1] > 1).mean() > 0.5  # most samples need more than the one Newton iteration of qp_fast
    for passes in (6, 2):
        for a, b in zip(runs[0], runs[passes]):
            assert np.array_equal(a["status"], b["status"])
            assert np.abs(a["u0"] - b["u0"]).max() < 1e-9
            assert np.abs(a["cost"] - b["cost"]).max() < 1e-9 * np.abs(a["cost"]).max()
            assert np.abs(a["dpi"] - b["dpi"]).max() < 1e-6 * np.abs(a["dpi"]).max()


@pytest.mark.parametrize("n_mass,n", [(3, 4), (6, 2)])  # nx = 9 / 113 parameters; nx = 27 / 800 parameters (BASELINE configs[2])
def test_chain_mass_host_run_matches_oracle(n_mass, n):
    from mpc4rl_b200.problems import chain_mass_spec, get_chain_params
    from oracle import cpu_port as cp

    g = np.load(os.path.join(ROOT, "tests", "golden", f"chain_mass_{n_mass}.npz"))
    cpar = get_chain_params()
    cpar["n_mass"] = n_mass
    spec = chain_mass_spec(cpar)
    assert spec.ntheta == g["theta"].shape[0]
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-9, warm_ipm=1)
    o = cp.chain_unit(n_mass, pd, 0, 60, spec.p_nominal, spec.x_ss, g["x0"][:n])
    assert np.all(o["status"] == 0)
    assert np.abs(o["u0"] - g["u0"][:n]).max() < 1e-8
    assert np.abs(o["cost"] - g["V"][:n]).max() < 1e-9 * np.abs(g["V"]).max()
    assert np.abs(o["dL"] - g["dV"][:n]).max() < 1e-6 * np.abs(g["dV"]).max()
    assert np.abs(o["dpi"] - g["dpi"][:n]).max() < 1e-5 * np.abs(g["dpi"]).max()
    # one RTI step from that iterate at the moved state, sensitivities at the new iterate
    r = cp.chain_unit(n_mass, pd, 0, 1, spec.p_nominal, spec.x_ss, g["x1"][:n], iterate=o["iterate"])
    assert np.all(r["status"] == 0)
    assert np.abs(r["u0"] - g["u1"][:n]).max() < 1e-5
    assert np.abs(r["dpi"] - g["dpi1"][:n]).max() < 1e-4 * np.abs(g["dpi1"]).max()
