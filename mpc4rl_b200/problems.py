"""Host-side problem descriptions: what the reference packs into an ``AcadosOcp`` before it
hands it to ``AcadosOcpSolver`` -- dims, cost scaling, bounds, integrator constants and the
parameter vector ``p`` (layout of rlmpc/mpc/nlp.py:970-989).  They are turned into the
``rlmpc_problem_desc`` of the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from . import _cabi

INF = 1e30


@dataclass
class ProblemSpec:
    name: str
    model: int
    N: int
    nx: int
    nu: int
    tf: float
    p_entries: List[Tuple[str, Tuple[int, ...]]]  # name, shape -- column-major flattening
    p_nominal: np.ndarray
    lbu: np.ndarray
    ubu: np.ndarray
    lbx: np.ndarray
    ubx: np.ndarray
    lbx_e: np.ndarray
    ubx_e: np.ndarray
    model_const: np.ndarray
    gamma: float = 1.0
    cost_type: str = "NONLINEAR_LS"
    parameterize_tracking_cost: bool = False
    idxsbx: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=int))  # soft rows: positions within idxbx
    zl: np.ndarray = field(default_factory=lambda: np.zeros(0))
    zu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    lh: np.ndarray = field(default_factory=lambda: np.zeros(0))  # affine general constraints lh <= h(x,u) <= uh
    uh: np.ndarray = field(default_factory=lambda: np.zeros(0))
    x_init: np.ndarray | None = None  # initial guess of every stage where the reference sets one
    u_init: np.ndarray | None = None
    state_labels: List[str] = field(default_factory=list)
    input_labels: List[str] = field(default_factory=list)
    parameter_labels: List[str] = field(default_factory=list)  # labels of p["model"]
    model_vectors: Dict[str, np.ndarray] = field(default_factory=dict)  # model data too large for model_const (chain: x_ss)

    @property
    def ntheta(self) -> int:
        return int(sum(int(np.prod(s)) for _, s in self.p_entries))

    @property
    def np_model(self) -> int:
        return int(np.prod(dict(self.p_entries).get("model", (0,))))

    def p_slices(self) -> Dict[str, Tuple[slice, Tuple[int, ...]]]:
        out, o = {}, 0
        for name, shape in self.p_entries:
            n = int(np.prod(shape))
            out[name] = (slice(o, o + n), shape)
            o += n
        return out

    def cost_scaling(self, gamma: float | None = None) -> np.ndarray:
        """s_k of the reference's NLP cost (nlp.py:1038-1134, quirk Q3)."""
        g = self.gamma if gamma is None else gamma
        N, dT = self.N, self.tf / self.N
        s = np.empty(N + 1)
        if self.cost_type in ("NONLINEAR_LS", "LINEAR_LS") and not self.parameterize_tracking_cost:
            s[:N] = dT; s[N] = 1.0  # nlp.py:1044-1055: no gamma
        else:
            s[0] = dT
            for k in range(1, N):
                s[k] = g**k * dT
            s[N] = g**N
        return s

    def to_desc(self) -> _cabi.ProblemDesc:
        d = _cabi.ProblemDesc()
        d.model, d.N = self.model, self.N
        for i, v in enumerate(self.cost_scaling()):
            d.scale[i] = v
        for i in range(_cabi.MAXD):
            d.lbu[i] = d.lbx[i] = d.lbx_e[i] = d.lg[i] = -INF
            d.ubu[i] = d.ubx[i] = d.ubx_e[i] = d.ug[i] = INF
        for i, v in enumerate(np.asarray(self.lh, dtype=float).ravel()):
            d.lg[i] = v
        for i, v in enumerate(np.asarray(self.uh, dtype=float).ravel()):
            d.ug[i] = v
        for name in ("lbu", "ubu", "lbx", "ubx", "lbx_e", "ubx_e"):
            arr = getattr(d, name)
            for i, v in enumerate(np.asarray(getattr(self, name), dtype=float).ravel()):
                arr[i] = v
        for i, v in enumerate(self.model_const):
            d.model_const[i] = v
        for i, v in enumerate(np.asarray(self.zl, dtype=float).ravel()):
            d.zl[i] = v
        for i, v in enumerate(np.asarray(self.zu, dtype=float).ravel()):
            d.zu[i] = v
        return d


def cartpole_spec(config: dict, gamma: float = 1.0) -> ProblemSpec:
    """From the reference's YAML ``config["mpc"]`` (config/cartpole*.yaml), following
    rlmpc/mpc/cartpole/acados.py:28-203 and cartpole/common.py:349-360."""
    N = int(config["dimensions"]["N"])
    tf = float(config["ocp_options"]["tf"])
    nstg = int(config["ocp_options"].get("sim_method_num_stages", 4))
    params = config["model"]["params"]
    free = [k for k in ("M", "m", "l", "g") if not params[k]["fixed"]]
    if free not in (["M", "m", "l"], ["M", "m", "l", "g"]):
        # (the reference's YAML fixes g; scripts/cartpole_mpc_qlearning.py:184-187 un-fixes everything)
        raise NotImplementedError(f"cartpole device models exist for free parameters (M, m, l) and (M, m, l, g); got {free}")
    cost = config["cost"]
    W_0, W, W_e = (np.array(cost[k], dtype=float) for k in ("W_0", "W", "W_e"))
    yref_0, yref, yref_e = (np.array(cost[k], dtype=float) for k in ("yref_0", "yref", "yref_e"))
    cons = config["constraints"]
    nx, nu = 4, 1
    full = lambda lo, idx, val: _scatter(lo, idx, val, nx)
    lbx = full(-INF, cons.get("idxbx", []), cons.get("lbx", []))
    ubx = full(INF, cons.get("idxbx", []), cons.get("ubx", []))
    lbx_e = full(-INF, cons.get("idxbx_e", []), cons.get("lbx_e", []))
    ubx_e = full(INF, cons.get("idxbx_e", []), cons.get("ubx_e", []))
    p_entries = [("model", (len(free),)), ("W_0", W_0.shape), ("W", W.shape), ("W_e", W_e.shape),
                 ("yref_0", yref_0.shape), ("yref", yref.shape), ("yref_e", yref_e.shape)]
    p_nom = np.concatenate([[params[k]["value"] for k in free], W_0.T.ravel(), W.T.ravel(), W_e.T.ravel(),
                            yref_0, yref, yref_e]).astype(float)
    return ProblemSpec(
        name=config["model"].get("name", "cartpole"), model=_cabi.MODEL_CARTPOLE, N=N, nx=nx, nu=nu, tf=tf,
        p_entries=p_entries, p_nominal=p_nom,
        lbu=np.array(cons["lbu"], dtype=float), ubu=np.array(cons["ubu"], dtype=float),
        lbx=lbx, ubx=ubx, lbx_e=lbx_e, ubx_e=ubx_e,
        # quirk Q1: one RK4 step of dT/4; [2] = 1: g is the fourth model parameter
        model_const=np.array([tf / N / nstg, float(params["g"]["value"]), 1.0 if "g" in free else 0.0]),
        gamma=gamma, cost_type=cost.get("cost_type", "NONLINEAR_LS"),
        state_labels=["x", "x_dot", "theta", "theta_dot"], input_labels=["F"], parameter_labels=free,
    )


def _scatter(fill, idx, val, n):
    out = np.full(n, fill, dtype=float)
    for i, v in zip(idx, val):
        out[int(i)] = float(v)
    return out


def cartpole_original_config() -> dict:
    """config/cartpole_original.yaml of the reference as a dict (``config["mpc"]``): N=40, tf=0.8,
    |u| <= 80, no state bounds -- BASELINE.json configs[1]."""
    W = np.diag([200.0, 0.02, 200.0, 0.02, 0.01]).tolist()
    W_e = np.diag([200.0, 0.02, 200.0, 0.02]).tolist()
    return {
        "id": "cartpole",
        "meta": {"json_file": "config/cartpole_ocp.json", "code_export_dir": "c_generated_code/acados_mpc"},
        "ocp_options": {"tf": 0.8, "qp_solver": "PARTIAL_CONDENSING_HPIPM", "hessian_approx": "GAUSS_NEWTON",
                        "integrator_type": "DISCRETE", "sim_method_num_stages": 4, "nlp_solver_type": "SQP",
                        "nlp_solver_max_iter": 500, "qp_solver_iter_max": 200},
        "dimensions": {"nx": 4, "nu": 1, "N": 40},
        "cost": {"cost_type_0": "NONLINEAR_LS", "cost_type": "NONLINEAR_LS", "cost_type_e": "NONLINEAR_LS",
                 "W_0": W, "W": W, "W_e": W_e, "yref_0": [0.0] * 5, "yref": [0.0] * 5, "yref_e": [0.0] * 4},
        "model": {"name": "cartpole", "params": {"M": {"value": 1.0, "fixed": False}, "m": {"value": 0.1, "fixed": False},
                                                 "l": {"value": 0.5, "fixed": False}, "g": {"value": 9.8, "fixed": True}}},
        "constraints": {"constr_type": "BGH", "x0": [0.0, 0.0, 3.14, 0.0], "idxbu": [0], "lbu": [-80.0], "ubu": [80.0]},
    }


def cartpole_config() -> dict:
    """config/cartpole.yaml of the reference as a dict (``config["mpc"]``): N=30, tf=3.0, |u| <= 30,
    box bounds on all four states on stages 1..N (config/cartpole.yaml:1-93)."""
    c = cartpole_original_config()
    W = np.diag([10.0, 0.1, 10.0, 0.1, 0.01]).tolist()
    W_e = np.diag([10.0, 0.1, 10.0, 0.1]).tolist()
    c["ocp_options"]["tf"] = 3.0
    c["dimensions"]["N"] = 30
    c["cost"].update({"W_0": W, "W": W, "W_e": W_e})
    bx = [2.4, 10.0, 6.28, 10.0]
    c["constraints"] = {"constr_type": "BGH", "x0": [0.0, 0.0, 3.14, 0.0], "idxbu": [0], "lbu": [-30.0], "ubu": [30.0],
                        "idxbx_0": [0, 1, 2, 3], "idxbx": [0, 1, 2, 3], "lbx": [-v for v in bx], "ubx": bx,
                        "idxbx_e": [0, 1, 2, 3], "lbx_e": [-v for v in bx], "ubx_e": bx}
    return c


def linear_system_param_nominal() -> dict:
    """The parameter dict of tests/test_linear_example.py:9-17 / examples/linear_system_mpc_qlearning.py:109-117."""
    return {
        "A": np.array([[1.0, 0.25], [0.0, 1.0]]), "B": np.array([[0.03125], [0.25]]),
        "Q": np.identity(2), "R": np.identity(1), "b": np.array([[0.0], [0.0]]),
        "f": np.array([[0.0], [0.0], [0.0]]), "V_0": np.array([1e-3]),
    }


def linear_system_spec(param: dict | None = None, gamma: float = 0.99, N: int = 40,
                       lbx=(-0.0, -1.0), ubx=(1.0, 1.0), lbu=(-1.0,), ubu=(1.0,)) -> ProblemSpec:
    """rlmpc/mpc/linear_system/acados.py:73-131: x+ = Ax + Bu + b, EXTERNAL cost 1/2 y'y + f'y (+ V_0 at
    stage 0), terminal 1/2 x'Px with P = DARE(A,B,Q,R) of THIS dict (a constant), soft bound on x[0]."""
    from scipy.linalg import solve_discrete_are

    param = linear_system_param_nominal() if param is None else param
    P = solve_discrete_are(param["A"], param["B"], param["Q"], param["R"])
    p_nom = np.concatenate([np.asarray(param[k], dtype=float).T.reshape(-1) for k in ["A", "B", "b", "V_0", "f"]])
    labels = ["A_0", "A_1", "A_2", "A_3", "B_0", "B_1", "b_0", "b_1", "V_0", "f_0", "f_1", "f_2"]
    return ProblemSpec(
        name="lti", model=_cabi.MODEL_LINEAR_SYSTEM, N=N, nx=2, nu=1, tf=float(N),  # tf = N: dT = 1 (acados.py:120)
        p_entries=[("model", (12,))], p_nominal=p_nom,
        lbu=np.array(lbu, dtype=float), ubu=np.array(ubu, dtype=float),
        lbx=np.array(lbx, dtype=float), ubx=np.array(ubx, dtype=float),
        lbx_e=np.full(2, -INF), ubx_e=np.full(2, INF),
        model_const=np.array([P[0, 0], P[0, 1], P[1, 1]]), gamma=gamma, cost_type="EXTERNAL",
        idxsbx=np.array([0]), zl=np.array([1e2]), zu=np.array([1e2]),
        state_labels=["x_0", "x_1"], input_labels=["u_0"], parameter_labels=labels,
    )


EVAPORATION_PARAM = {
    # rlmpc/gym/evaporation_process/environment.py:5-25 (dict order = order of the model constants)
    "a": 0.5616, "b": 0.3126, "c": 48.43, "d": 0.507, "e": 55.0, "f": 0.1538, "g": 90.0, "h": 0.16, "M": 20.0,
    "C": 4.0, "U_A2": 6.84, "C_p": 0.07, "lam": 38.5, "lam_s": 36.6, "F_1": 10.0, "X_1": 5.0, "F_3": 50.0,
    "T_1": 40.0, "T_200": 25.0,
}
H_NOMINAL = np.diag([10.0, 10.0, 0.1, 0.1, 0.1])  # scripts/evaporation_process_mpc.py:33


def evaporation_spec(model_param: dict | None = None, cost_param: dict | None = None, gamma: float = 1.0,
                     H: np.ndarray | None = None, N: int = 100, n_sub: int = 4) -> ProblemSpec:
    """rlmpc/mpc/evaporation_process/acados.py:142-228.  ``cost_param["H"]["l"]`` (or ``H``) is the 5x5
    tracking weight; theta = [W_0, W, yref_0, yref] (parameterize_tracking_cost=True, acados.py:115)."""
    mp = dict(EVAPORATION_PARAM if model_param is None else model_param)
    if H is None:
        H = cost_param["H"]["l"] if cost_param is not None else H_NOMINAL
    H = np.asarray(H, dtype=float)
    x_ss, u_ss = np.array([25.0, 49.743]), np.array([191.713, 215.888, 0.0])
    yref = np.concatenate([x_ss, u_ss])
    p_entries = [("model", (0,)), ("W_0", (5, 5)), ("W", (5, 5)), ("yref_0", (5,)), ("yref", (5,))]
    p_nom = np.concatenate([H.T.ravel(), H.T.ravel(), yref, yref])
    mc = np.array([mp[k] for k in EVAPORATION_PARAM] + [1.0 / n_sub, float(n_sub)])
    return ProblemSpec(
        name="evaporation_process", model=_cabi.MODEL_EVAPORATION, N=N, nx=2, nu=3, tf=float(N),
        p_entries=p_entries, p_nominal=p_nom,
        lbu=np.array([100.0, 100.0, 0.0]), ubu=np.array([400.0, 400.0, 10.0]),
        lbx=np.full(2, -INF), ubx=np.full(2, INF), lbx_e=np.full(2, -INF), ubx_e=np.full(2, INF),
        model_const=mc, gamma=gamma, cost_type="NONLINEAR_LS", parameterize_tracking_cost=True,
        lh=np.array([-1e3, -1e3]), uh=np.array([0.0, 0.0]), x_init=x_ss, u_init=u_ss,
        state_labels=["X_2", "P_2"], input_labels=["P_100", "F_200", "s"], parameter_labels=[],
    )


# --------------------------------------------------------------------------------------------
# chain of masses (rlmpc/mpc/chain_mass/ocp_utils.py)
# --------------------------------------------------------------------------------------------
def get_chain_params() -> dict:
    """ocp_utils.py:319-341 (the entries the OCP uses)."""
    return {"n_mass": 5, "Ts": 0.2, "Tsim": 5, "N": 40, "u_init": np.array([-1, 1, 1]), "with_wall": True, "yPosWall": -0.05,
            "xPosFirstMass": np.zeros(3), "m": 0.033, "D": 1.0, "L": 0.033, "C": 0.1, "perturb_scale": 1e-2,
            "nlp_iter": 50, "seed": 50, "nlp_tol": 1e-5}


def chain_param_layout(n_mass: int):
    """define_param_struct_symSX(disturbance=True) (ocp_utils.py:353-371): [m | D | L | C | Q | R | w]; entries with
    `repeat` are laid out repetition by repetition, matrices column-major.  Returns ({name: slice}, length, labels)."""
    n_link, M = n_mass - 1, n_mass - 2
    nx = (2 * M + 1) * 3
    sizes = [("m", n_link), ("D", 3 * n_link), ("L", 3 * n_link), ("C", 3 * n_link), ("Q", nx * nx), ("R", 9), ("w", 3 * M)]
    out, o = {}, 0
    for name, n in sizes:
        out[name] = slice(o, o + n)
        o += n
    labels = [f"m_{i}" for i in range(n_link)]
    for nm in ("D", "L", "C"):
        labels += [f"{nm}_{i}_{j}" for i in range(n_link) for j in range(3)]
    labels += [f"Q_{i}" for i in range(nx * nx)] + [f"R_{i}" for i in range(9)]
    labels += [f"w_{i}_{j}" for i in range(M) for j in range(3)]
    return out, o, labels


def chain_ode(x: np.ndarray, u: np.ndarray, p: np.ndarray, n_mass: int) -> np.ndarray:
    """f_expl = [xvel ; u ; f] of ocp_utils.py:59-147 (numpy; used on the host for the steady state only)."""
    M = n_mass - 2
    sl, _, _ = chain_param_layout(n_mass)
    m, D, L, C = p[sl["m"]], p[sl["D"]].reshape(M + 1, 3), p[sl["L"]].reshape(M + 1, 3), p[sl["C"]].reshape(M + 1, 3)
    w = p[sl["w"]].reshape(M, 3)
    pos, vel = x[: 3 * (M + 1)].reshape(M + 1, 3), x[3 * (M + 1):].reshape(M, 3)
    f = np.tile(np.array([0.0, 0.0, -9.81]), (M, 1)) + w
    for i in range(M + 1):
        dist = pos[i] - (pos[i - 1] if i > 0 else 0.0)
        vl = vel[0] if i == 0 else (u - vel[M - 1] if i == M else vel[i] - vel[i - 1])
        T = D[i] / m[i] * (1.0 - L[i] / np.linalg.norm(dist)) * dist + C[i] * vl
        if i < M:
            f[i] -= T
        if i > 0:
            f[i - 1] += T
    return np.concatenate([vel.reshape(-1), u, f.reshape(-1)])


def chain_steady_state(n_mass: int, p: np.ndarray, x_end: np.ndarray) -> np.ndarray:
    """compute_parametric_steady_state (ocp_utils.py:150-192) without IPOPT: xdot = 0 with the last mass held at
    x_end and u = 0 -- zero velocities and force balance on the intermediate masses; Newton with a central-difference
    Jacobian on the 3 M unknown positions, started on the straight line like the reference's initial guess."""
    M = n_mass - 2
    nx = (2 * M + 1) * 3

    def resid(q):
        x = np.concatenate([q, x_end, np.zeros(3 * M)])
        return chain_ode(x, np.zeros(3), p, n_mass)[3 * (M + 1):]

    q = np.zeros(3 * M)
    q[0::3] = np.linspace(0.0, float(x_end[0]), M + 2)[1:-1]
    for _ in range(100):
        r = resid(q)
        if np.abs(r).max() < 1e-13:
            break
        J = np.zeros((3 * M, 3 * M))
        for j in range(3 * M):
            e = np.zeros(3 * M)
            e[j] = 1e-6
            J[:, j] = (resid(q + e) - resid(q - e)) / 2e-6
        q = q - np.linalg.solve(J, r)
    x = np.zeros(nx)
    x[: 3 * M] = q
    x[3 * M: 3 * (M + 1)] = x_end
    return x


def chain_define_x0(chain_params: dict) -> np.ndarray:
    """define_x0 (rlmpc/examples/chain_mass.py:17-25): masses on the straight line to x_end, at rest."""
    M = chain_params["n_mass"] - 2
    x0 = np.zeros((2 * M + 1) * 3)
    x0[: 3 * (M + 1): 3] = np.linspace(chain_params["xPosFirstMass"][0], chain_params["L"] * (M + 1) * 6, M + 2)[1:]
    return x0


def chain_mass_spec(chain_params: dict | None = None, gamma: float = 1.0) -> ProblemSpec:
    """export_parametric_ocp(chain_params, integrator_type="DISCRETE") as rlmpc/mpc/chain_mass/acados.py:32-45 builds it:
    disturbance parameters present (w = 0), EXTERNAL cost 1/2 (x - x_ss)'Q(x - x_ss) + 1/2 u'Ru with Q, R part of p
    (ocp_utils.py:266-277), |u| <= 1, ERK4 with two sub-steps of Ts/2 (:42-56), GAUSS_NEWTON, tol 1e-5, 50 SQP
    iterations (:298-312).  theta = p has 113 / 499 / 800 entries for n_mass = 3 / 5 / 6."""
    cp = dict(get_chain_params() if chain_params is None else chain_params)
    n_mass = int(cp["n_mass"])
    M = n_mass - 2
    nx, nu, N, Ts = (2 * M + 1) * 3, 3, int(cp["N"]), float(cp["Ts"])
    sl, nth, labels = chain_param_layout(n_mass)
    p = np.zeros(nth)
    p[sl["m"]] = cp["m"]; p[sl["D"]] = cp["D"]; p[sl["L"]] = cp["L"]; p[sl["C"]] = cp["C"]  # random_scale = 0 (ocp_utils.py:205)
    q_diag = np.ones(nx)
    q_diag[3 * M: 3 * M + 3] = M + 1
    p[sl["Q"]] = (2.0 * np.diag(q_diag)).T.ravel()
    p[sl["R"]] = (2.0 * 1e-2 * np.eye(nu)).T.ravel()
    x_end = np.array([cp["L"] * (M + 1) * 6, 0.0, 0.0])
    x_ss = chain_steady_state(n_mass, p, x_end)
    spec = ProblemSpec(
        name=f"chain_mass_ds_{n_mass}", model=_cabi.MODEL_CHAIN_MASS, N=N, nx=nx, nu=nu, tf=N * Ts,
        p_entries=[("model", (nth,))], p_nominal=p,
        lbu=-np.ones(nu), ubu=np.ones(nu), lbx=np.zeros(0), ubx=np.zeros(0), lbx_e=np.zeros(0), ubx_e=np.zeros(0),
        model_const=np.array([Ts / 2.0, float(n_mass)]), gamma=gamma, cost_type="EXTERNAL",
        state_labels=[f"xpos_{i}" for i in range(3 * (M + 1))] + [f"xvel_{i}" for i in range(3 * M)],
        input_labels=[f"u_{i}" for i in range(nu)], parameter_labels=labels,
    )
    spec.model_vectors = {"x_ss": x_ss}
    spec.x_ss = x_ss
    return spec
