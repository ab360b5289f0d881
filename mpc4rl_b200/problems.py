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
    if free != ["M", "m", "l"]:
        raise NotImplementedError(f"cartpole device model is generated for free parameters (M, m, l); got {free}")
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
    p_entries = [("model", (3,)), ("W_0", W_0.shape), ("W", W.shape), ("W_e", W_e.shape),
                 ("yref_0", yref_0.shape), ("yref", yref.shape), ("yref_e", yref_e.shape)]
    p_nom = np.concatenate([[params[k]["value"] for k in free], W_0.T.ravel(), W.T.ravel(), W_e.T.ravel(),
                            yref_0, yref, yref_e]).astype(float)
    return ProblemSpec(
        name=config["model"].get("name", "cartpole"), model=_cabi.MODEL_CARTPOLE, N=N, nx=nx, nu=nu, tf=tf,
        p_entries=p_entries, p_nominal=p_nom,
        lbu=np.array(cons["lbu"], dtype=float), ubu=np.array(cons["ubu"], dtype=float),
        lbx=lbx, ubx=ubx, lbx_e=lbx_e, ubx_e=ubx_e,
        model_const=np.array([tf / N / nstg, float(params["g"]["value"])]),  # quirk Q1: one RK4 step of dT/4
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
