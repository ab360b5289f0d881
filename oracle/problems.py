"""Restated problem definitions (TEST INFRASTRUCTURE, see oracle/__init__.py).

Each ``Problem`` carries what the reference puts into an ``AcadosOcp``: discrete
dynamics, stage/terminal cost, bounds, the parameter struct ``p`` (layout of
``rlmpc/mpc/nlp.py:970-989``: ``model, W_0, W, W_e, yref_0, yref, yref_e`` -- only the
non-empty ones, matrices flattened column-major), horizon and cost scaling.
Everything is written with torch float64 ops so that every derivative the oracle
needs comes from torch.func, independently of the hand-derived CUDA code.

Reference files followed:
  cartpole       rlmpc/mpc/cartpole/acados.py:28-108, config/cartpole*.yaml,
                 rlmpc/common/integrator.py:6-33
  linear system  rlmpc/mpc/linear_system/acados.py:27-131
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

F64 = torch.float64


def erk4(f: Callable, x, u, pm, h: float):
    """One explicit RK4 step (rlmpc/common/integrator.py:6-33)."""
    k1 = f(x, u, pm)
    k2 = f(x + h / 2 * k1, u, pm)
    k3 = f(x + h / 2 * k2, u, pm)
    k4 = f(x + h * k3, u, pm)
    return x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


@dataclass
class Problem:
    name: str
    N: int
    nx: int
    nu: int
    tf: float
    p_entries: List[Tuple[str, Tuple[int, ...]]]
    p_nominal: np.ndarray
    cost_type: str  # "NLS" | "EXTERNAL"
    hessian_approx: str  # "GAUSS_NEWTON" | "EXACT"
    f_disc: Callable  # (x, u, p_model) -> x_next
    gamma: float = 1.0
    parameterize_tracking_cost: bool = False
    # NLS pieces (y = y_fun(x,u), y_e = y_e_fun(x)); constants used when not parameterised
    y_fun: Optional[Callable] = None
    y_e_fun: Optional[Callable] = None
    W_0: Optional[np.ndarray] = None
    W: Optional[np.ndarray] = None
    W_e: Optional[np.ndarray] = None
    yref_0: Optional[np.ndarray] = None
    yref: Optional[np.ndarray] = None
    yref_e: Optional[np.ndarray] = None
    # EXTERNAL pieces
    ext_cost_0: Optional[Callable] = None  # (x,u,pm)
    ext_cost: Optional[Callable] = None
    ext_cost_e: Optional[Callable] = None  # (x,pm)
    # bounds (acados BGH subset the reference uses)
    idxbu: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=int))
    lbu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    ubu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    idxbx: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=int))
    lbx: np.ndarray = field(default_factory=lambda: np.zeros(0))
    ubx: np.ndarray = field(default_factory=lambda: np.zeros(0))
    idxbx_e: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=int))
    lbx_e: np.ndarray = field(default_factory=lambda: np.zeros(0))
    ubx_e: np.ndarray = field(default_factory=lambda: np.zeros(0))
    idxsbx: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=int))
    zl: np.ndarray = field(default_factory=lambda: np.zeros(0))
    zu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    idxsbx_e: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=int))
    zl_e: np.ndarray = field(default_factory=lambda: np.zeros(0))
    zu_e: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # general constraints lh <= h(x,u) <= uh on stages 0..N-1; the reference's h is affine:
    # h = h0 + Ch [x;u]  (evaporation_process/acados.py:207-210)
    Ch: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))
    h0: np.ndarray = field(default_factory=lambda: np.zeros(0))
    lh: np.ndarray = field(default_factory=lambda: np.zeros(0))
    uh: np.ndarray = field(default_factory=lambda: np.zeros(0))
    # initial guess of every stage (the evaporation model needs a non-zero one, acados.py:104-109)
    x_init: Optional[np.ndarray] = None
    u_init: Optional[np.ndarray] = None

    # ---- parameter struct helpers (nlp.py:970-989; CasADi column-major) ----
    @property
    def ntheta(self) -> int:
        return int(sum(int(np.prod(s)) for _, s in self.p_entries))

    def p_slices(self):
        out, o = {}, 0
        for name, shape in self.p_entries:
            n = int(np.prod(shape))
            out[name] = (slice(o, o + n), shape)
            o += n
        return out

    def p_get(self, p, name):
        sl, shape = self.p_slices()[name]
        v = p[sl]
        if len(shape) == 2 and shape[1] > 1:
            return v.reshape(shape[1], shape[0]).T  # column-major
        return v.reshape(shape[0]) if len(shape) >= 1 else v

    def p_model(self, p):
        sl, _ = self.p_slices()["model"]
        return p[sl]

    @property
    def dT(self) -> float:
        return self.tf / self.N

    # ---- cost exactly as build_nlp assembles it (nlp.py:1038-1134) ----
    def stage_scale(self, k: int) -> float:
        """Multiplier of the stage-k cost term."""
        N, g, dT = self.N, self.gamma, self.dT
        if self.cost_type == "NLS" and not self.parameterize_tracking_cost:
            return dT if k < N else 1.0  # nlp.py:1044-1055 (no gamma)
        # parameterised NLS (1057-1074) and EXTERNAL (1084-1091)
        if k == 0:
            return dT
        if k < N:
            return g**k * dT
        return g**N


def _nls(y, yref, W):
    e = y - yref
    return 0.5 * (e @ (W @ e))


def stage_cost_unscaled(pb: Problem, k: int, x, u, p):
    """l_0 / l / l_e of the reference (unscaled)."""
    if pb.cost_type == "NLS":
        if pb.parameterize_tracking_cost:
            if k == 0:
                return _nls(pb.y_fun(x, u), pb.p_get(p, "yref_0"), pb.p_get(p, "W_0"))
            if k < pb.N:
                return _nls(pb.y_fun(x, u), pb.p_get(p, "yref"), pb.p_get(p, "W"))
            if "W_e" not in dict(pb.p_entries):
                return 0.0 * x.sum()  # no terminal cost (nlp.py:1069-1072)
            return _nls(pb.y_e_fun(x), pb.p_get(p, "yref_e"), pb.p_get(p, "W_e"))
        T = lambda a: torch.as_tensor(a, dtype=F64)
        if k == 0:
            return _nls(pb.y_fun(x, u), T(pb.yref_0), T(pb.W_0))
        if k < pb.N:
            return _nls(pb.y_fun(x, u), T(pb.yref), T(pb.W))
        return _nls(pb.y_e_fun(x), T(pb.yref_e), T(pb.W_e))
    pm = pb.p_model(p)
    if k == 0:
        return pb.ext_cost_0(x, u, pm)
    if k < pb.N:
        return pb.ext_cost(x, u, pm)
    return pb.ext_cost_e(x, pm)


# --------------------------------------------------------------------------------------
# cartpole  (rlmpc/mpc/cartpole/acados.py:28-108)
# --------------------------------------------------------------------------------------
def cartpole_ode(x, u, pm, g: float = 9.8):
    M, m, l = pm[0], pm[1], pm[2]
    if pm.shape[0] == 4:  # g un-fixed (scripts/cartpole_mpc_qlearning.py:184-187): fourth model parameter
        g = pm[3]
    s_dot, theta, theta_dot = x[1], x[2], x[3]
    c, sn = torch.cos(theta), torch.sin(theta)
    temp = (u[0] + m * theta_dot**2 * sn) / (m + M)
    theta_ddot = (g * sn - c * temp) / (l * (4.0 / 3.0 - m * c**2 / (m + M)))
    return torch.stack([s_dot, temp - m * theta_ddot * c / (m + M), theta_dot, theta_ddot])


def make_cartpole(variant: str = "original", gamma: float = 1.0, free_g: bool = False) -> Problem:
    """variant 'original' = config/cartpole_original.yaml (N=40, tf=0.8, |u|<=80, no state
    bounds); 'default' = config/cartpole.yaml (N=30, tf=3.0, |u|<=30, state bounds)."""
    if variant == "original":
        N, tf, umax = 40, 0.8, 80.0
        W = np.diag([200.0, 0.02, 200.0, 0.02, 0.01])
        W_e = np.diag([200.0, 0.02, 200.0, 0.02])
        bx = None
    elif variant == "default":
        N, tf, umax = 30, 3.0, 30.0
        W = np.diag([10.0, 0.1, 10.0, 0.1, 0.01])
        W_e = np.diag([10.0, 0.1, 10.0, 0.1])
        bx = np.array([2.4, 10.0, 6.28, 10.0])
    else:
        raise ValueError(variant)
    h = tf / N / 4  # quirk Q1: ONE RK4 step of dT / sim_method_num_stages (cartpole/acados.py:86-92)
    p_entries = [("model", (4 if free_g else 3,)), ("W_0", (5, 5)), ("W", (5, 5)), ("W_e", (4, 4)),
                 ("yref_0", (5,)), ("yref", (5,)), ("yref_e", (4,))]
    p_nom = np.concatenate([[1.0, 0.1, 0.5] + ([9.8] if free_g else []), W.T.ravel(), W.T.ravel(), W_e.T.ravel(),
                            np.zeros(5), np.zeros(5), np.zeros(4)])
    pb = Problem(
        name=f"cartpole_{variant}", N=N, nx=4, nu=1, tf=tf, p_entries=p_entries, p_nominal=p_nom,
        cost_type="NLS", hessian_approx="GAUSS_NEWTON", gamma=gamma,
        f_disc=lambda x, u, pm: erk4(cartpole_ode, x, u, pm, h),
        y_fun=lambda x, u: torch.cat([x, u]), y_e_fun=lambda x: x,
        W_0=W, W=W, W_e=W_e, yref_0=np.zeros(5), yref=np.zeros(5), yref_e=np.zeros(4),
        idxbu=np.array([0]), lbu=np.array([-umax]), ubu=np.array([umax]),
    )
    if bx is not None:
        pb.idxbx = np.arange(4); pb.lbx = -bx; pb.ubx = bx
        pb.idxbx_e = np.arange(4); pb.lbx_e = -bx; pb.ubx_e = bx
    return pb


# --------------------------------------------------------------------------------------
# linear system  (rlmpc/mpc/linear_system/acados.py:27-131)
# --------------------------------------------------------------------------------------
def linear_system_param_nominal():
    """tests/test_linear_example.py:9-17 / examples/linear_system_mpc_qlearning.py:109-117"""
    return {
        "A": np.array([[1.0, 0.25], [0.0, 1.0]]), "B": np.array([[0.03125], [0.25]]),
        "Q": np.identity(2), "R": np.identity(1), "b": np.array([[0.0], [0.0]]),
        "f": np.array([[0.0], [0.0], [0.0]]), "V_0": np.array([1e-3]),
    }


def make_linear_system(param: Optional[dict] = None, gamma: float = 0.99, N: int = 40,
                       lbx=(-0.0, -1.0), ubx=(1.0, 1.0)) -> Problem:
    from scipy.linalg import solve_discrete_are

    param = linear_system_param_nominal() if param is None else param
    P = solve_discrete_are(param["A"], param["B"], param["Q"], param["R"])  # constant (acados.py:51-57)
    Pt = torch.as_tensor(P, dtype=F64)

    def unpack(pm):
        A = pm[0:4].reshape(2, 2).T  # column-major (acados.py:60-70)
        B = pm[4:6]
        b = pm[6:8]
        V0 = pm[8]
        f = pm[9:12]
        return A, B, b, V0, f

    def f_disc(x, u, pm):
        A, B, b, _, _ = unpack(pm)
        return A @ x + B * u[0] + b

    def ext(x, u, pm):
        _, _, _, _, f = unpack(pm)
        y = torch.cat([x, u])
        return 0.5 * (y @ y) + f @ y

    def ext0(x, u, pm):
        return unpack(pm)[3] + ext(x, u, pm)

    def exte(x, pm):
        return 0.5 * (x @ (Pt @ x))

    # ocp.parameter_values = concat(param[key].T.reshape(-1,1)) (acados.py:90)
    p_nom = np.concatenate([np.asarray(param[k], dtype=float).T.reshape(-1) for k in ["A", "B", "b", "V_0", "f"]])
    return Problem(
        name="linear_system", N=N, nx=2, nu=1, tf=float(N), p_entries=[("model", (12,))], p_nominal=p_nom,
        cost_type="EXTERNAL", hessian_approx="EXACT", gamma=gamma, f_disc=f_disc,
        ext_cost_0=ext0, ext_cost=ext, ext_cost_e=exte,
        idxbu=np.array([0]), lbu=np.array([-1.0]), ubu=np.array([1.0]),
        idxbx=np.array([0, 1]), lbx=np.array(lbx, dtype=float), ubx=np.array(ubx, dtype=float),
        idxsbx=np.array([0]), zl=np.array([1e2]), zu=np.array([1e2]),
    )


# --------------------------------------------------------------------------------------
# evaporation process  (rlmpc/mpc/evaporation_process/acados.py:142-228,
#                       rlmpc/gym/evaporation_process/environment.py:5-25, 50-106)
# --------------------------------------------------------------------------------------
EVAPORATION_PARAM = {
    "a": 0.5616, "b": 0.3126, "c": 48.43, "d": 0.507, "e": 55.0, "f": 0.1538, "g": 90.0, "h": 0.16, "M": 20.0,
    "C": 4.0, "U_A2": 6.84, "C_p": 0.07, "lam": 38.5, "lam_s": 36.6, "F_1": 10.0, "X_1": 5.0, "F_3": 50.0,
    "T_1": 40.0, "T_200": 25.0,
}


def evaporation_ode(x, u, pm):
    """compute_data + f_expl with the plant parameters as constants (model.p is empty)."""
    p = EVAPORATION_PARAM if pm is None or len(pm) == 0 else pm
    X_2, P_2 = x[0], x[1]
    P_100, F_200 = u[0], u[1]
    T_2 = p["a"] * P_2 + p["b"] * X_2 + p["c"]
    T_3 = p["d"] * P_2 + p["e"]
    T_100 = p["f"] * P_100 + p["g"]
    U_A1 = p["h"] * (p["F_1"] + p["F_3"])
    Q_100 = U_A1 * (T_100 - T_2)
    F_4 = (Q_100 - p["F_1"] * p["C_p"] * (T_2 - p["T_1"])) / p["lam"]
    Q_200 = p["U_A2"] * (T_3 - p["T_200"]) / (1 + (p["U_A2"] / (2 * p["C_p"] * F_200)))
    F_5 = Q_200 / p["lam"]
    F_2 = p["F_1"] - F_4
    return torch.stack([(p["F_1"] * p["X_1"] - F_2 * X_2) / p["M"], (F_4 - F_5) / p["C"]])


def make_evaporation(H: Optional[np.ndarray] = None, gamma: float = 0.99, N: int = 100, n_sub: int = 4) -> Problem:
    H = np.diag([10.0, 10.0, 0.1, 0.1, 0.1]) if H is None else np.asarray(H, float)  # H_nominal, scripts/evaporation_process_mpc.py:33
    x_ss, u_ss = np.array([25.0, 49.743]), np.array([191.713, 215.888, 0.0])
    yref = np.concatenate([x_ss, u_ss])
    dt = 1.0 / n_sub

    def f_disc(x, u, pm):
        for _ in range(n_sub):  # acados.py:74-86
            x = erk4(lambda xx, uu, _p: evaporation_ode(xx, uu, None), x, u, None, dt)
        return x

    p_entries = [("model", (0,)), ("W_0", (5, 5)), ("W", (5, 5)), ("yref_0", (5,)), ("yref", (5,))]
    p_nom = np.concatenate([H.T.ravel(), H.T.ravel(), yref, yref])
    Ch = np.array([[-1.0, 0.0, 0.0, 0.0, -1.0], [0.0, -1.0, 0.0, 0.0, -1.0]])  # h = 25 - x - s
    return Problem(
        name="evaporation_process", N=N, nx=2, nu=3, tf=float(N), p_entries=p_entries, p_nominal=p_nom,
        cost_type="NLS", hessian_approx="GAUSS_NEWTON", gamma=gamma, parameterize_tracking_cost=True,
        f_disc=f_disc, y_fun=lambda x, u: torch.cat([x, u]), y_e_fun=lambda x: x,
        idxbu=np.array([0, 1, 2]), lbu=np.array([100.0, 100.0, 0.0]), ubu=np.array([400.0, 400.0, 10.0]),
        Ch=Ch, h0=np.array([25.0, 25.0]), lh=np.array([-1e3, -1e3]), uh=np.array([0.0, 0.0]),
        x_init=x_ss, u_init=u_ss,
    )


# --------------------------------------------------------------------------------------
# chain of masses  (rlmpc/mpc/chain_mass/ocp_utils.py:42-56, 59-147, 195-316, 319-371)
# --------------------------------------------------------------------------------------
def chain_params() -> dict:
    """get_chain_params() (ocp_utils.py:319-341), the entries the OCP uses."""
    return {"n_mass": 5, "Ts": 0.2, "N": 40, "m": 0.033, "D": 1.0, "L": 0.033, "C": 0.1, "xPosFirstMass": np.zeros(3)}


def chain_param_layout(n_mass: int):
    """define_param_struct_symSX(disturbance=True) (ocp_utils.py:353-371): [m | D | L | C | Q | R | w], entries with
    `repeat` laid out repetition by repetition, matrices column-major.  Returns {name: slice} and the total length."""
    n_link, M = n_mass - 1, n_mass - 2
    nx = (2 * M + 1) * 3
    sizes = [("m", n_link), ("D", 3 * n_link), ("L", 3 * n_link), ("C", 3 * n_link), ("Q", nx * nx), ("R", 9), ("w", 3 * M)]
    out, o = {}, 0
    for name, n in sizes:
        out[name] = slice(o, o + n)
        o += n
    return out, o


def chain_ode(x, u, pm, n_mass: int):
    """f_expl = [xvel ; u ; f] (ocp_utils.py:59-147).  Spring force of link i on its masses:
    F_j = D_ij / m_i (1 - L_ij / |dist|) dist_j; damping F_j = C_ij vel_j (not divided by the mass, as in the
    reference); gravity -9.81 on z; disturbance w_i added to the acceleration of intermediate mass i."""
    M = n_mass - 2
    sl, _ = chain_param_layout(n_mass)
    m, D, L, C, w = pm[sl["m"]], pm[sl["D"]].reshape(M + 1, 3), pm[sl["L"]].reshape(M + 1, 3), pm[sl["C"]].reshape(M + 1, 3), pm[sl["w"]].reshape(M, 3)
    xpos = x[: 3 * (M + 1)].reshape(M + 1, 3)
    xvel = x[3 * (M + 1):].reshape(M, 3)
    zero3 = torch.zeros(3, dtype=x.dtype)
    f = [torch.stack([zero3[0], zero3[0], zero3[0] - 9.81]) for _ in range(M)]
    for i in range(M + 1):
        dist = xpos[i] - (xpos[i - 1] if i > 0 else zero3)
        F = D[i] / m[i] * (1.0 - L[i] / torch.linalg.vector_norm(dist)) * dist
        if i < M:
            f[i] = f[i] - F
        if i > 0:
            f[i - 1] = f[i - 1] + F
    for i in range(M + 1):
        if i == 0:
            vel = xvel[0]
        elif i == M:
            vel = u - xvel[M - 1]
        else:
            vel = xvel[i] - xvel[i - 1]
        F = C[i] * vel
        if i < M:
            f[i] = f[i] - F
        if i > 0:
            f[i - 1] = f[i - 1] + F
    f = [f[i] + w[i] for i in range(M)]
    return torch.cat([xvel.reshape(-1), u, torch.cat(f)])


def chain_steady_state(n_mass: int, pm: np.ndarray, x_end: np.ndarray) -> np.ndarray:
    """compute_parametric_steady_state (ocp_utils.py:150-192) without IPOPT: xdot = 0 with the last mass held at
    x_end and u = 0, i.e. zero velocities and force balance on the intermediate masses (Newton on 3M unknowns,
    started on the straight line like the reference's initial guess)."""
    M = n_mass - 2
    nx = (2 * M + 1) * 3
    pmt = torch.as_tensor(pm, dtype=F64)

    def resid(q):  # q: positions of the M intermediate masses
        x = torch.cat([q, torch.as_tensor(x_end, dtype=F64), torch.zeros(3 * M, dtype=F64)])
        return chain_ode(x, torch.zeros(3, dtype=F64), pmt, n_mass)[3 * (M + 1):]

    q = torch.zeros(3 * M, dtype=F64)
    q[0::3] = torch.linspace(0.0, float(x_end[0]), M + 2, dtype=F64)[1:-1]
    for _ in range(50):
        r = resid(q)
        if float(r.abs().max()) < 1e-13:
            break
        J = torch.func.jacrev(resid)(q)
        q = q - torch.linalg.solve(J, r)
    x = np.zeros(nx)
    x[: 3 * M] = q.numpy()
    x[3 * M: 3 * (M + 1)] = x_end
    return x


def make_chain_mass(n_mass: int = 5, gamma: float = 1.0, params: Optional[dict] = None) -> Problem:
    """export_parametric_ocp(chain_params, integrator_type="DISCRETE") as chain_mass/acados.py:32-45 builds it:
    disturbance parameters present (w = 0), EXTERNAL cost 1/2 (x-x_ss)'Q(x-x_ss) + 1/2 u'Ru with Q, R part of p,
    |u| <= 1, ERK4 with two sub-steps of Ts/2, GAUSS_NEWTON (= exact Hessian of the quadratic cost), gamma = 1 in
    examples/chain_mass.py:main_nlp.  theta = p has 113 / 499 / 800 entries for n_mass = 3 / 5 / 6."""
    cp = dict(chain_params() if params is None else params)
    cp["n_mass"] = n_mass
    M, n_link = n_mass - 2, n_mass - 1
    nx, nu, N, Ts = (2 * M + 1) * 3, 3, cp["N"], cp["Ts"]
    sl, nth = chain_param_layout(n_mass)
    p = np.zeros(nth)
    p[sl["m"]] = cp["m"]; p[sl["D"]] = cp["D"]; p[sl["L"]] = cp["L"]; p[sl["C"]] = cp["C"]  # random_scale = 0 (ocp_utils.py:205)
    q_diag = np.ones(nx)
    q_diag[3 * M: 3 * M + 3] = M + 1
    p[sl["Q"]] = (2.0 * np.diag(q_diag)).T.ravel()
    p[sl["R"]] = (2.0 * 1e-2 * np.eye(nu)).T.ravel()
    x_end = np.array([cp["L"] * (M + 1) * 6, 0.0, 0.0])
    x_ss = chain_steady_state(n_mass, p, x_end)
    xss_t = torch.as_tensor(x_ss, dtype=F64)
    h = Ts / 2

    def f_disc(x, u, pm):
        ode = lambda xx, uu, pp: chain_ode(xx, uu, pp, n_mass)
        for _ in range(2):  # export_discrete_erk4_integrator_step, n_stages = 2 (ocp_utils.py:42-56)
            x = erk4(ode, x, u, pm, h)
        return x

    def Qm(pm):
        return pm[sl["Q"]].reshape(nx, nx).T

    def ext(x, u, pm):
        e = x - xss_t
        return 0.5 * (e @ (Qm(pm) @ e) + u @ (pm[sl["R"]].reshape(nu, nu).T @ u))

    def exte(x, pm):
        e = x - xss_t
        return 0.5 * (e @ (Qm(pm) @ e))

    x0 = np.zeros(nx)  # define_x0 (examples/chain_mass.py:17-25): masses on the straight line to x_end, at rest
    x0[: 3 * (M + 1): 3] = np.linspace(cp["xPosFirstMass"][0], x_end[0], M + 2)[1:]
    pb = Problem(
        name=f"chain_mass_{n_mass}", N=N, nx=nx, nu=nu, tf=N * Ts, p_entries=[("model", (nth,))], p_nominal=p,
        cost_type="EXTERNAL", hessian_approx="GAUSS_NEWTON", gamma=gamma, f_disc=f_disc,
        ext_cost_0=ext, ext_cost=ext, ext_cost_e=exte,
        idxbu=np.arange(nu), lbu=-np.ones(nu), ubu=np.ones(nu),
    )
    pb.x_ss = x_ss
    pb.x0_example = x0
    return pb
