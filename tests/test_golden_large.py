"""The large golden sets (oracle/make_golden_large.py: 256 converged samples of cartpole_original and of the linear system,
64 of cartpole.yaml) against the host build of the engine (CPU suite) and against the CUDA path through the C ABI
(-m gpu).  Tolerances as in the small sets: |u0| 1e-6 abs, V/Q 1e-9 rel, dV/dtheta 1e-6 rel,
dpi/dtheta 1e-5 rel; samples on which the ORACLE's full-step SQP did not converge are skipped (and counted)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}.npz not generated")
    return np.load(path)


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _check(g, v, q, min_ok, dpi_mask=None, dpi_tol=1e-5):
    """v / q: dicts with status, u0, cost, dL, dpi (engine gradient width) of the V- and the Q-mode run."""
    cols = g["cols"]
    okv = (g["status"][:, 0] == 0) & (v["status"] == 0)
    okq = (g["status"][:, 1] == 0) & (q["status"] == 0)
    assert okv.mean() >= min_ok and okq.mean() >= min_ok, (okv.mean(), okq.mean())
    # the engine may converge where the oracle's undamped SQP cycles, never the other way round on more than a handful
    assert ((g["status"][:, 0] == 0) & (v["status"] != 0)).sum() <= 2
    w = v["dL"].shape[1]
    assert np.array_equal(cols, np.arange(len(cols))) and w >= len(cols)  # live columns = the engine's gradient prefix
    assert np.abs(v["u0"] - g["u0"])[okv].max() < 1e-6
    assert _rel(v["cost"][okv], g["V"][okv]) < 1e-9
    assert _rel(v["dL"][okv][:, :len(cols)], g["dV"][okv]) < 1e-6
    m = okv if dpi_mask is None else okv & dpi_mask
    assert _rel(v["dpi"][m][:, :, :len(cols)], g["dpi"][m]) < dpi_tol
    assert _rel(q["cost"][okq], g["Q"][okq]) < 1e-9
    assert _rel(q["dL"][okq][:, :len(cols)], g["dQ"][okq]) < 1e-6
    return int(okv.sum()), int(okq.sum())


# ---------------------------------------------------------------- host build (CPU suite)
def test_host_port_cartpole_original_256():
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec
    from oracle import cpu_port as cp

    g = _load("cartpole_original_256")
    spec = cartpole_spec(cartpole_original_config())
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-10, warm_ipm=1)
    v = cp.unit(1, pd, 0, 300, g["theta"], g["x0"])
    q = cp.unit(1, pd, 1, 300, g["theta"], g["x0"], u0=g["a"])
    nv, nq = _check(g, v, q, 0.85)
    assert nv >= 220 and nq >= 220


def test_host_port_linear_system_256():
    from scipy.linalg import solve_discrete_are

    from oracle import cpu_port as cp
    from oracle.problems import linear_system_param_nominal, make_linear_system

    g = _load("linear_system_256")
    par = linear_system_param_nominal()
    P = solve_discrete_are(par["A"], par["B"], par["Q"], par["R"])
    pb = make_linear_system(gamma=0.9)
    scale = np.array([pb.stage_scale(k) for k in range(pb.N + 1)])
    pd = cp.make_pd(pb.N, scale, pb.lbu, pb.ubu, [P[0, 0], P[0, 1], P[1, 1]], tol=1e-10, warm_ipm=1, lbx=pb.lbx, ubx=pb.ubx,
                    zl=pb.zl, zu=pb.zu)
    v = cp.unit(3, pd, 0, 100, g["theta"], g["x0"], nx=2, nu=1)
    q = cp.unit(3, pd, 1, 100, g["theta"], g["x0"], u0=g["a"], nx=2, nu=1)
    # dpi/dtheta with an active slack: the reference formula is a ratio of interior-point stiffnesses (quirks Q4 / Q7),
    # compared only where no softened bound is active
    soft = g["slmax"] > 1e-6 if "slmax" in g.files else None
    nv, nq = _check(g, v, q, 0.95, None if soft is None else ~soft)
    assert nv >= 250


@pytest.mark.parametrize("fixture,N", [("evaporation_32", 40), ("evaporation_n100_8", 100)])
def test_host_port_evaporation_sets(fixture, N):
    from oracle import cpu_port as cp
    from oracle.problems import EVAPORATION_PARAM, make_evaporation

    g = _load(fixture)
    pb = make_evaporation(gamma=0.95, N=N)
    scale = np.array([pb.stage_scale(k) for k in range(N + 1)])
    pd = cp.make_pd(N, scale, pb.lbu, pb.ubu, list(EVAPORATION_PARAM.values()) + [0.25, 4], tol=1e-9, warm_ipm=1, lg=pb.lh, ug=pb.uh)
    B, nx, nu = g["x0"].shape[0], 2, 3
    res = []
    for mode, u0 in ((0, None), (1, g["a"])):
        it = np.zeros((cp.lib().cpu_port_iterate_size(4, N), B))  # every stage on the steady state, like the reference
        for k in range(N + 1):
            it[k * nx:(k + 1) * nx, :] = pb.x_init[:, None]
        for k in range(N):
            it[(N + 1) * nx + k * nu:(N + 1) * nx + (k + 1) * nu, :] = pb.u_init[:, None]
        res.append(cp.unit(4, pd, mode, 100, g["theta"], g["x0"], u0=u0, iterate=it, nx=nx, nu=nu))
    _check(g, res[0], res[1], 0.9)


# ---------------------------------------------------------------- CUDA path
def _gpu_run(spec, g, max_sqp):
    import torch

    from mpc4rl_b200 import BatchedMPC

    dev = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device="cuda:0")
    B = g["x0"].shape[0]
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-10)
    x0 = dev(g["x0"])
    res = []
    for u0 in (None, dev(g["a"])):
        m.reset(x0)
        o = m.solve_sens(x0, u0=u0, max_sqp=max_sqp)
        res.append({k: t.cpu().numpy() for k, t in o.items()})
    return res


@pytest.mark.gpu
def test_gpu_cartpole_original_256():
    from mpc4rl_b200 import cartpole_original_config, cartpole_spec

    g = _load("cartpole_original_256")
    v, q = _gpu_run(cartpole_spec(cartpole_original_config()), g, 300)
    nv, nq = _check(g, v, q, 0.85)
    assert nv >= 220 and nq >= 220


@pytest.mark.gpu
def test_gpu_cartpole_default_64():
    from mpc4rl_b200 import cartpole_config, cartpole_spec

    g = _load("cartpole_default_64")
    v, q = _gpu_run(cartpole_spec(cartpole_config()), g, 300)
    _check(g, v, q, 0.7)


@pytest.mark.gpu
def test_gpu_linear_system_256():
    from mpc4rl_b200 import linear_system_spec

    g = _load("linear_system_256")
    v, q = _gpu_run(linear_system_spec(gamma=0.9), g, 100)
    soft = g["slmax"] > 1e-6 if "slmax" in g.files else None
    # 2e-5: x_0 is eliminated here while the reference's formula carries it as two barrier rows of finite stiffness (quirk
    # Q7), an O(tau / stiffness) difference that shows on a few of the 188 slack-free samples (worst: 1.2e-5)
    nv, nq = _check(g, v, q, 0.95, None if soft is None else ~soft, dpi_tol=2e-5)
    assert nv >= 250


# ---------------------------------------------------------------- indefinite exact Hessian (closed-loop RTI iterates)
def _check_indefinite(g, status, u0, cost, dL, dpi):
    assert g["indefinite"].sum() >= 8 and (~g["indefinite"]).sum() >= 8  # the fixture covers both kinds
    assert np.all(status == 0), status  # the reference's sparse LU does not need a definite reduced Hessian
    assert np.abs(u0 - g["u1"]).max() < 1e-5
    assert np.abs(cost - g["V1"]).max() < 1e-8 * np.abs(g["V1"]).max()
    for sel in (g["indefinite"], ~g["indefinite"]):
        assert np.abs(dL - g["dV1"])[sel].max() < 1e-5 * np.abs(g["dV1"][sel]).max()
        rel = np.abs(dpi - g["dpi1"])[sel].max(axis=(1, 2)) / np.abs(g["dpi1"][sel]).max(axis=(1, 2))
        assert rel.max() < 1e-4, rel  # per sample: dpi/dtheta spans 1e-5 .. 1e5 over these iterates


def test_host_port_sensitivities_with_indefinite_exact_hessian():
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec
    from oracle import cpu_port as cp

    g = _load("cartpole_original_indefinite")
    spec = cartpole_spec(cartpole_original_config())
    N, B = spec.N, g["x1"].shape[0]
    pd = cp.make_pd(N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-6, warm_ipm=1)
    it = np.zeros((cp.lib().cpu_port_iterate_size(1, N), B))
    it[:(N + 1) * 4] = g["X"].reshape(B, -1).T
    it[(N + 1) * 4:(N + 1) * 4 + N] = g["U"].reshape(B, -1).T
    o = cp.unit(1, pd, 0, 1, g["theta"], g["x1"], iterate=it)
    _check_indefinite(g, o["status"], o["u0"], o["cost"], o["dL"], o["dpi"])


@pytest.mark.gpu
def test_gpu_sensitivities_with_indefinite_exact_hessian():
    import torch

    from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec

    g = _load("cartpole_original_indefinite")
    spec = cartpole_spec(cartpole_original_config())
    B = g["x1"].shape[0]
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")
    m = BatchedMPC(spec, max_batch=B, device=0)  # default options = the bench's
    m.reset(dev(g["x1"]))
    for k in range(spec.N + 1):
        m.put("x", k, dev(g["X"][:, k]))
    for k in range(spec.N):
        m.put("u", k, dev(g["U"][:, k]))
    o = m.solve_sens(dev(g["x1"]), max_sqp=1)
    _check_indefinite(g, o["status"].cpu().numpy(), o["u0"].cpu().numpy(), o["cost"].cpu().numpy(), o["dL"].cpu().numpy(),
                      o["dpi"].cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,N", [("evaporation_32", 40), ("evaporation_n100_8", 100)])
def test_gpu_evaporation_sets(fixture, N):
    """32 states of the SURVEY.md 8(d) config-4 distribution at the oracle's affordable horizon N = 40 and 8 at the
    reference's full horizon N = 100 (gamma 0.95), V- and Q-mode, all 60 tracking-cost parameter columns."""
    import torch

    from mpc4rl_b200 import BatchedMPC, evaporation_spec

    g = _load(fixture)
    spec = evaporation_spec(gamma=0.95, N=N)
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")
    B = g["x0"].shape[0]
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-9)
    x0 = dev(g["x0"])
    res = []
    for u0 in (None, dev(g["a"])):
        m.reset(B=B)  # every stage on the steady state (evaporation_process/acados.py:104-109)
        for k in range(spec.N + 1):
            m.put("x", k, dev(np.tile(spec.x_init, (B, 1))))
        for k in range(spec.N):
            m.put("u", k, dev(np.tile(spec.u_init, (B, 1))))
        o = m.solve_sens(x0, u0=u0, max_sqp=100)
        res.append({k: t.cpu().numpy() for k, t in o.items()})
    _check(g, res[0], res[1], 0.9)


# ---------------------------------------------------------------- away from the nominal parameters: one theta per sample
def _check_theta(g, v, q):
    okv = (g["status"][:, 0] == 0) & (v["status"] == 0)
    okq = (g["status"][:, 1] == 0) & (q["status"] == 0)
    assert okv.sum() >= 24 and okq.sum() >= 24, (okv.sum(), okq.sum())
    assert np.ptp(g["theta"][:, :3], axis=0).min() > 0.02  # every model parameter really differs between samples
    assert np.abs(v["u0"] - g["u0"])[okv].max() < 1e-6
    assert _rel(v["cost"][okv], g["V"][okv]) < 1e-9
    assert _rel(v["dL"][okv][:, :3], g["dV"][okv]) < 1e-6
    rel = np.abs(v["dpi"][:, :, :3] - g["dpi"])[okv].max(axis=(1, 2)) / np.abs(g["dpi"][okv]).max(axis=(1, 2))
    assert rel.max() < 1e-5, rel  # per sample: the scale of dpi/dtheta changes with theta
    assert _rel(q["cost"][okq], g["Q"][okq]) < 1e-9
    assert _rel(q["dL"][okq][:, :3], g["dQ"][okq]) < 1e-6


def test_host_port_per_sample_parameters_match_oracle():
    """32 cart-pole states, each with its own (M, m, l) = nominal x U(0.6, 1.4) (oracle/make_golden_theta.py): what
    MPC.set_p does between learning steps, and the [B, ntheta] form of rlmpc_set_theta."""
    from mpc4rl_b200.problems import cartpole_original_config, cartpole_spec
    from oracle import cpu_port as cp

    g = _load("cartpole_original_theta")
    spec = cartpole_spec(cartpole_original_config())
    pd = cp.make_pd(spec.N, spec.cost_scaling(), spec.lbu, spec.ubu, spec.model_const, tol=1e-10, warm_ipm=1)
    v = cp.unit(1, pd, 0, 300, g["theta"], g["x0"])
    q = cp.unit(1, pd, 1, 300, g["theta"], g["x0"], u0=g["a"])
    _check_theta(g, v, q)


@pytest.mark.gpu
def test_gpu_per_sample_parameters_match_oracle():
    import torch

    from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec

    g = _load("cartpole_original_theta")
    spec = cartpole_spec(cartpole_original_config())
    dev = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device="cuda:0")
    B = g["x0"].shape[0]
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-10)
    m.set_theta(g["theta"])  # [B, ntheta]: one parameter vector per sample
    x0 = dev(g["x0"])
    res = []
    for u0 in (None, dev(g["a"])):
        m.reset(x0)
        o = m.solve_sens(x0, u0=u0, max_sqp=300)
        res.append({k: t.cpu().numpy() for k, t in o.items()})
    _check_theta(g, res[0], res[1])


def _check_theta_full(g, v, q, min_ok, dpi_tol=1e-5):
    """Linear system / evaporation away from nominal: all parameter columns, per-sample scale for dpi/dtheta."""
    okv = (g["status"][:, 0] == 0) & (v["status"] == 0)
    okq = (g["status"][:, 1] == 0) & (q["status"] == 0)
    assert okv.sum() >= min_ok and okq.sum() >= min_ok, (okv.sum(), okq.sum())
    nc = g["dV"].shape[1]
    assert np.abs(v["u0"] - g["u0"])[okv].max() < 1e-6
    assert _rel(v["cost"][okv], g["V"][okv]) < 1e-9
    assert _rel(v["dL"][okv][:, :nc], g["dV"][okv]) < 1e-6
    m = okv & (g["slmax"] <= 1e-6)  # dpi/dtheta with an active slack: quirks Q4 / Q7 (see _check above)
    assert m.sum() >= min_ok // 2
    rel = np.abs(v["dpi"][:, :, :nc] - g["dpi"])[m].max(axis=(1, 2)) / np.abs(g["dpi"][m]).max(axis=(1, 2))
    assert rel.max() < dpi_tol, rel
    assert _rel(q["cost"][okq], g["Q"][okq]) < 1e-9
    assert _rel(q["dL"][okq][:, :nc], g["dQ"][okq]) < 1e-6


def test_host_port_linear_system_and_evaporation_away_from_nominal():
    """Per-sample theta for the two problems whose learning loops move it: the linear system's 12 model parameters
    (N(0, 0.03^2) around nominal) and the evaporation process' tracking weights / references (N = 40)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_rti_secondary import _host_problem

    from oracle import cpu_port as cp

    for name, min_ok in (("linear_system", 28), ("evaporation", 14)):
        g = _load(name + "_theta")
        pb, pd, model, nx, nu = _host_problem(name)
        B, N = g["x0"].shape[0], pb.N
        res = []
        for mode, u0 in ((0, None), (1, g["a"])):
            it = None
            if name == "evaporation":
                it = np.zeros((cp.lib().cpu_port_iterate_size(model, N), B))
                for k in range(N + 1):
                    it[k * nx:(k + 1) * nx, :] = pb.x_init[:, None]
                for k in range(N):
                    it[(N + 1) * nx + k * nu:(N + 1) * nx + (k + 1) * nu, :] = pb.u_init[:, None]
            res.append(cp.unit(model, pd, mode, 300, g["theta"], g["x0"], u0=u0, iterate=it, nx=nx, nu=nu))
        _check_theta_full(g, res[0], res[1], min_ok, dpi_tol=2e-5 if name == "linear_system" else 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name,min_ok", [("linear_system", 28), ("evaporation", 14)])
def test_gpu_linear_system_and_evaporation_away_from_nominal(name, min_ok):
    import torch

    from mpc4rl_b200 import BatchedMPC, evaporation_spec, linear_system_spec

    g = _load(name + "_theta")
    spec = linear_system_spec(gamma=0.9) if name == "linear_system" else evaporation_spec(gamma=0.95, N=40)
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")
    B = g["x0"].shape[0]
    m = BatchedMPC(spec, max_batch=B, device=0)
    m.set_option("tol", 1e-9 if name == "evaporation" else 1e-10)
    m.set_theta(g["theta"])  # [B, ntheta]
    x0 = dev(g["x0"])
    res = []
    for u0 in (None, dev(g["a"])):
        if name == "evaporation":
            m.reset(B=B)
            for k in range(spec.N + 1):
                m.put("x", k, dev(np.tile(spec.x_init, (B, 1))))
            for k in range(spec.N):
                m.put("u", k, dev(np.tile(spec.u_init, (B, 1))))
        else:
            m.reset(x0)
        o = m.solve_sens(x0, u0=u0, max_sqp=300)
        res.append({k: t.cpu().numpy() for k, t in o.items()})
    _check_theta_full(g, res[0], res[1], min_ok, dpi_tol=2e-5 if name == "linear_system" else 1e-5)
