"""The SQP-RTI step + sensitivities that bench.py times, on the SECONDARY configurations, against the dense oracle:
config/cartpole.yaml (state bounds, one environment step), the linear system (softened bound, one step of
LinearSystemEnv) and the evaporation process (general rows, exact Hessian; N = 40, the oracle's affordable horizon).
Fixtures: tests/golden/{cartpole_default,linear_system,evaporation,evaporation_n100}_rti.npz (oracle/make_golden_rti_more.py: ONE dense
SQP step from the stored converged iterate incl. multipliers, QP at the tau-central point, restated update_nlp at the
new iterate; oracle outputs, parity unpinned vs acados).  Host build of the engine in the CPU suite, CUDA path through
the C ABI under -m gpu.  Tolerances of the timed path (tests/test_gpu_rti_oracle.py): |du0| 1e-5, V 1e-8 rel,
dL/dtheta 1e-5 rel, dpi/dtheta 1e-4 rel."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["cartpole_default", "linear_system", "evaporation", "evaporation_n100"]  # the last: full horizon N = 100, 8 states


def _load(name):
    path = os.path.join(ROOT, "tests", "golden", name + "_rti.npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}_rti.npz not generated")
    return np.load(path)


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _check(name, g, st0, r, rq, min_frac):
    """st0: status of the engine's converged solve; r / rq: V- and Q-mode RTI results at x1."""
    nc = len(g["cols"])
    ok = (g["status"] == 0) & g["ok1"] & (st0 == 0)
    assert ok.mean() >= min_frac, ok.mean()
    assert np.all(r["status"][ok] == 0) and np.all(rq["status"][ok] == 0)
    assert np.abs(r["u0"] - g["u1"])[ok].max() < 1e-5
    assert _rel(r["cost"][ok], g["V1"][ok]) < 1e-8
    assert _rel(r["dL"][ok][:, :nc], g["dV1"][ok]) < 1e-5
    # dpi/dtheta with an active slack is a ratio of interior-point stiffnesses in the reference formula (quirks Q4 / Q7)
    m = ok & (g["sl1"] <= 1e-6)
    assert m.sum() >= 0.5 * ok.sum()
    assert _rel(r["dpi"][m][:, :, :nc], g["dpi1"][m]) < 1e-4
    assert _rel(rq["cost"][ok], g["Q1"][ok]) < 1e-8
    assert _rel(rq["dL"][ok][:, :nc], g["dQ1"][ok]) < 1e-5
    return int(ok.sum())


# ---------------------------------------------------------------- host build (CPU suite)
def _host_problem(name):
    from oracle import cpu_port as cp
    from oracle.make_golden_large import _problem

    pb = _problem(name)
    scale = np.array([pb.stage_scale(k) for k in range(pb.N + 1)])
    if name == "cartpole_default":
        pd = cp.make_pd(pb.N, scale, pb.lbu, pb.ubu, [pb.tf / pb.N / 4, 9.8], tol=1e-10, warm_ipm=1, lbx=pb.lbx, ubx=pb.ubx,
                        lbx_e=pb.lbx_e, ubx_e=pb.ubx_e)
        return pb, pd, 2, 4, 1
    if name == "linear_system":
        from scipy.linalg import solve_discrete_are

        from oracle.problems import linear_system_param_nominal

        par = linear_system_param_nominal()
        P = solve_discrete_are(par["A"], par["B"], par["Q"], par["R"])
        pd = cp.make_pd(pb.N, scale, pb.lbu, pb.ubu, [P[0, 0], P[0, 1], P[1, 1]], tol=1e-10, warm_ipm=1, lbx=pb.lbx, ubx=pb.ubx,
                        zl=pb.zl, zu=pb.zu)
        return pb, pd, 3, 2, 1
    from oracle.problems import EVAPORATION_PARAM

    pd = cp.make_pd(pb.N, scale, pb.lbu, pb.ubu, list(EVAPORATION_PARAM.values()) + [0.25, 4], tol=1e-9, warm_ipm=1, lg=pb.lh, ug=pb.uh)
    return pb, pd, 4, 2, 3


@pytest.mark.parametrize("name", NAMES)
def test_host_port_rti_step_matches_oracle(name):
    from oracle import cpu_port as cp

    g = _load(name)
    pb, pd, model, nx, nu = _host_problem(name)
    B, N = g["x0"].shape[0], pb.N
    it = None
    if name.startswith("evaporation"):  # every stage on the steady state, like the reference
        it = np.zeros((cp.lib().cpu_port_iterate_size(model, N), B))
        for k in range(N + 1):
            it[k * nx:(k + 1) * nx, :] = pb.x_init[:, None]
        for k in range(N):
            it[(N + 1) * nx + k * nu:(N + 1) * nx + (k + 1) * nu, :] = pb.u_init[:, None]
    v = cp.unit(model, pd, 0, 300, g["theta"], g["x0"], iterate=it, nx=nx, nu=nu, do_sens=False)
    x1 = np.where(np.isfinite(g["x1"]), g["x1"], g["x0"])
    pd.tol = 1e-6  # the engine's default; an RTI call is one step whatever the tolerance
    r = cp.unit(model, pd, 0, 1, g["theta"], x1, iterate=v["iterate"].copy(), nx=nx, nu=nu)
    rq = cp.unit(model, pd, 1, 1, g["theta"], x1, u0=g["a"], iterate=v["iterate"].copy(), nx=nx, nu=nu)
    _check(name, g, v["status"], r, rq, 0.7)


# ---------------------------------------------------------------- CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_rti_step_matches_oracle(name):
    import torch

    from mpc4rl_b200 import BatchedMPC, cartpole_config, cartpole_spec, evaporation_spec, linear_system_spec

    g = _load(name)
    spec = {"cartpole_default": lambda: cartpole_spec(cartpole_config()), "linear_system": lambda: linear_system_spec(gamma=0.9),
            "evaporation": lambda: evaporation_spec(gamma=0.95, N=40),
            "evaporation_n100": lambda: evaporation_spec(gamma=0.95, N=100)}[name]()
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:0")
    B = g["x0"].shape[0]
    x0, x1 = dev(g["x0"]), dev(np.where(np.isfinite(g["x1"]), g["x1"], g["x0"]))
    res = []
    for u0 in (None, dev(g["a"])):
        m = BatchedMPC(spec, max_batch=B, device=0)  # default options = the bench's, except the tolerance of the setup solve
        m.set_option("tol", 1e-9 if name.startswith("evaporation") else 1e-10)
        if name.startswith("evaporation"):
            m.reset(B=B)
            for k in range(spec.N + 1):
                m.put("x", k, dev(np.tile(spec.x_init, (B, 1))))
            for k in range(spec.N):
                m.put("u", k, dev(np.tile(spec.u_init, (B, 1))))
        else:
            m.reset(x0)
        _, _, st0 = m.solve(x0, max_sqp=300)
        m.set_option("tol", 1e-6)
        o = m.solve_sens(x1, u0=u0, max_sqp=1)
        res.append({k: t.cpu().numpy() for k, t in o.items()})
    _check(name, g, st0.cpu().numpy(), res[0], res[1], 0.7)
