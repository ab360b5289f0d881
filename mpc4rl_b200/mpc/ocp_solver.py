"""``AcadosOcpSolver`` look-alike backed by the B200 engine (batch of one sample).

Covers the subset of acados_template.AcadosOcpSolver that the reference uses (census in
SURVEY.md section 1 / 8(b)): set, get, solve, solve_for_x0, get_cost, get_residuals, cost_set,
constraints_set, reset, store_iterate, load_iterate, status, acados_ocp.  One instance owns one
mutable iterate and is not re-entrant, like the original.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np
import torch

from ..batched import BatchedMPC
from ..problems import INF, ProblemSpec


def _finite_idx(lo, hi):
    return np.array([i for i in range(len(lo)) if lo[i] > -INF / 2 or hi[i] < INF / 2], dtype=int)


def make_ocp_view(spec: ProblemSpec) -> SimpleNamespace:
    """The attributes of ``ocp_solver.acados_ocp`` that the reference reads."""
    idxbx, idxbx_e = _finite_idx(spec.lbx, spec.ubx), _finite_idx(spec.lbx_e, spec.ubx_e)
    sl = spec.p_slices()
    pv = lambda k: (spec.p_nominal[sl[k][0]].reshape(sl[k][1][::-1]).T.copy() if len(sl[k][1]) == 2
                    else spec.p_nominal[sl[k][0]].copy()) if k in sl else np.array([])
    cons = SimpleNamespace(
        idxbu=np.arange(spec.nu), lbu=spec.lbu.copy(), ubu=spec.ubu.copy(),
        idxbx=idxbx, lbx=spec.lbx[idxbx].copy(), ubx=spec.ubx[idxbx].copy(),
        idxbx_e=idxbx_e, lbx_e=spec.lbx_e[idxbx_e].copy(), ubx_e=spec.ubx_e[idxbx_e].copy(),
        idxbx_0=np.arange(spec.nx), lbx_0=np.zeros(spec.nx), ubx_0=np.zeros(spec.nx),
        idxsbx=np.asarray(spec.idxsbx, dtype=int), idxsbu=np.array([], dtype=int), idxsh=np.array([], dtype=int),
        idxsbx_e=np.array([], dtype=int), idxsh_e=np.array([], dtype=int),
        lh=np.asarray(spec.lh, dtype=float).copy(), uh=np.asarray(spec.uh, dtype=float).copy(),
        lh_e=np.array([]), uh_e=np.array([]),
    )
    dims = SimpleNamespace(N=spec.N, nx=spec.nx, nu=spec.nu, np=spec.np_model, nbu=spec.nu, nbx=len(idxbx),
                           nbx_0=spec.nx, nbx_e=len(idxbx_e), nh=len(spec.lh), nh_e=0, nsbx=len(spec.idxsbx), nsbu=0, nsh=0, nsbx_e=0, nsh_e=0,
                           ny_0=spec.nx + spec.nu, ny=spec.nx + spec.nu, ny_e=spec.nx)
    cost = SimpleNamespace(cost_type_0=spec.cost_type, cost_type=spec.cost_type, cost_type_e=spec.cost_type,
                           W_0=pv("W_0"), W=pv("W"), W_e=pv("W_e"), yref_0=pv("yref_0"), yref=pv("yref"), yref_e=pv("yref_e"),
                           zl=np.asarray(spec.zl, dtype=float).copy(), zu=np.asarray(spec.zu, dtype=float).copy(),
                           zl_e=np.array([]), zu_e=np.array([]))
    model = SimpleNamespace(name=spec.name, x_labels=list(spec.state_labels), u_labels=list(spec.input_labels),
                            p_labels=list(spec.parameter_labels))
    return SimpleNamespace(dims=dims, constraints=cons, cost=cost, model=model,
                           parameter_values=spec.p_nominal[sl["model"][0]].copy() if "model" in sl else np.array([]),
                           solver_options=SimpleNamespace(tf=spec.tf, nlp_solver_max_iter=100, tol=1e-6))


class OcpSolverShim:
    def __init__(self, spec: ProblemSpec, device: int = 0, max_iter: int = 100, tol: float = 1e-6):
        self.spec = spec
        self.engine = BatchedMPC(spec, max_batch=1, device=device)
        self.acados_ocp = make_ocp_view(spec)
        self.acados_ocp.solver_options.nlp_solver_max_iter = int(max_iter)
        self.acados_ocp.solver_options.tol = float(tol)
        self.engine.set_option("tol", tol)
        self.status = 0
        self.N = spec.N
        self._theta = np.array(spec.p_nominal, dtype=np.float64)
        self._lbx0 = np.zeros(spec.nx)
        self._ubx0 = np.zeros(spec.nx)
        self._lbu0 = spec.lbu.copy()
        self._ubu0 = spec.ubu.copy()
        self._cost = float("nan")
        self._res = np.full(4, np.nan)
        self._dev = self.engine.device
        # acados starts a fresh solver from x_k = constraints.x0 on every stage; here x_0 arrives with the first solve, so
        # an iterate nobody has written yet (set / load_iterate / reset) is initialised there (MPC.reset semantics)
        self._iterate_written = False

    # ---- helpers ----
    def _t(self, v, n):
        a = np.asarray(v, dtype=np.float64).reshape(-1)
        if a.shape[0] != n:
            raise ValueError(f"expected {n} values, got {a.shape[0]}")
        return torch.tensor(a.reshape(1, n), dtype=torch.float64, device=self._dev)

    @property
    def qmode(self) -> bool:
        """u_0 is clamped (q_update: constraints_set(0,'lbu'/'ubu',u0), mpc.py:71-73)."""
        return bool(np.all(self._lbu0 == self._ubu0))

    def _push_theta(self):
        self.engine.set_theta(self._theta)

    # ---- acados API subset ----
    def set(self, stage: int, field: str, value) -> None:
        if field in ("lbx", "ubx"):
            if stage != 0:
                raise NotImplementedError("only the stage-0 state bounds (x0 fixing) can be changed at run time")
            setattr(self, "_lbx0" if field == "lbx" else "_ubx0", np.asarray(value, dtype=np.float64).reshape(-1).copy())
            getattr(self.acados_ocp.constraints, field + "_0")[:] = np.asarray(value, dtype=np.float64).reshape(-1)
        elif field in ("x", "u", "pi"):
            dim = {"x": self.spec.nx, "u": self.spec.nu, "pi": self.spec.nx}[field]
            self.engine.put(field, stage, self._t(value, dim))
            self._iterate_written = True
        elif field == "p":
            sl = self.spec.p_slices()["model"][0]
            self._theta[sl] = np.asarray(value, dtype=np.float64).reshape(-1)
            self.acados_ocp.parameter_values = self._theta[sl].copy()
            self._push_theta()
        else:
            raise NotImplementedError(f"set(stage, {field!r}, ...) is not part of the reference's usage")

    def constraints_set(self, stage: int, field: str, value) -> None:
        v = np.asarray(value, dtype=np.float64).reshape(-1)
        if field in ("lbu", "ubu"):
            if stage == 0:
                setattr(self, "_lbu0" if field == "lbu" else "_ubu0", v.copy())
            else:
                raise NotImplementedError("input bounds of stages > 0 are fixed at construction")
        elif field in ("lbx", "ubx"):
            self.set(stage, field, v)
        else:
            raise NotImplementedError(field)

    def cost_set(self, stage: int, field: str, value, api: str = "warn") -> None:
        key = {"W": "W", "yref": "yref"}.get(field)
        if key is None:
            raise NotImplementedError(field)
        name = key + ("_0" if stage == 0 else "_e" if stage == self.N else "")
        sl, shape = self.spec.p_slices()[name]
        v = np.asarray(value, dtype=np.float64)
        self._theta[sl] = v.T.reshape(-1) if v.ndim == 2 else v.reshape(-1)  # column-major like CasADi
        setattr(self.acados_ocp.cost, name, v.copy())
        self._push_theta()

    def set_theta(self, theta) -> None:
        """Whole parameter vector p at once (what MPC.set_p / set_parameter need)."""
        self._theta = np.array(theta, dtype=np.float64).reshape(-1)
        sl = self.spec.p_slices()
        if "model" in sl:
            self.acados_ocp.parameter_values = self._theta[sl["model"][0]].copy()
        self._push_theta()

    def get_theta(self) -> np.ndarray:
        return self._theta.copy()

    def reset(self) -> None:
        self.engine.reset(B=1)
        self._iterate_written = True

    def solve(self) -> int:
        if not np.array_equal(self._lbx0, self._ubx0):
            raise NotImplementedError("the engine fixes x_0: set(0,'lbx',x0) and set(0,'ubx',x0) must agree")
        x0 = self._t(self._lbx0, self.spec.nx)
        if not self._iterate_written:
            self.engine.reset(x0)
            self._iterate_written = True
        u0 = None
        if self.qmode:
            u0 = self._t(self._lbu0, self.spec.nu)
        elif not (np.array_equal(self._lbu0, self.spec.lbu) and np.array_equal(self._ubu0, self.spec.ubu)):
            raise NotImplementedError("stage-0 input bounds must be nominal (V) or equal (Q)")
        _, cost, status = self.engine.solve(x0, u0, max_sqp=self.acados_ocp.solver_options.nlp_solver_max_iter)
        self._cost = float(cost.item())
        self.status = int(status.item())
        return self.status

    def solve_for_x0(self, x0_bar, fail_on_nonzero_status: bool = True, print_stats_on_failure: bool = True):
        self.set(0, "lbx", x0_bar)
        self.set(0, "ubx", x0_bar)
        status = self.solve()
        if status != 0 and fail_on_nonzero_status:
            raise Exception(f"acados acados_ocp_solver returned status {status}")
        return self.get(0, "u")

    def evaluate(self):
        """update_nlp work at the current iterate: returns (dL_dtheta [ntheta], dpi_dtheta [nu, ntheta])."""
        dL, dpi, cost, res, status = self.engine.sens(1, qmode=self.qmode)
        self._cost = float(cost.item())
        self._res = res[0].cpu().numpy()
        return (self.engine.full_grad(dL)[0].cpu().numpy(), self.engine.full_grad(dpi)[0].cpu().numpy(), int(status.item()))

    def get_cost(self) -> float:
        return self._cost

    def eval_solution_sensitivity(self, stages, with_respect_to: str = "params_global"):
        """acados' ``AcadosOcpSolver.eval_solution_sensitivity(stages, with_respect_to)`` as the reference uses it
        (examples/chain_mass.py:126: ``_, sens_u = sensitivity_solver.eval_solution_sensitivity(0, "params_global")``):
        returns ``(sens_x, sens_u)`` = d x_stage / d p and d u_stage / d p, shapes (nx, np) and (nu, np), at the current
        iterate.  Only what the engine computes is available: stage 0, parameters -- d u_0 / d p is the adjoint KKT
        solve of ``update_nlp``, d x_0 / d p = 0 because x_0 is fixed.  (acados needs a second solver with the EXACT
        Hessian for this; the engine's sensitivities always use the exact Hessian.)"""
        if with_respect_to not in ("params_global", "p_global"):
            raise NotImplementedError(f"eval_solution_sensitivity w.r.t. {with_respect_to!r}: only the parameters are supported")
        single = np.isscalar(stages)
        for st in ([stages] if single else list(stages)):
            if int(st) != 0:
                raise NotImplementedError("eval_solution_sensitivity: only stage 0 (the policy gradient) is available")
        _, dpi, _ = self.evaluate()
        sl = self.spec.p_slices()["model"][0] if "model" in self.spec.p_slices() else slice(0, 0)
        sens_u = np.array(dpi)[:, sl]  # acados differentiates w.r.t. model.p, the "model" part of the NLP's p
        sens_x = np.zeros((self.spec.nx, sens_u.shape[1]))
        return (sens_x, sens_u) if single else (sens_x[None], sens_u[None])

    def get_residuals(self):
        """[stat, eq, ineq, comp] of the last evaluation (acados: get_residuals())."""
        _, _, cost, res, _ = self.engine.sens(1, qmode=self.qmode)
        self._res = res[0].cpu().numpy()
        return self._res.copy()

    def get(self, stage: int, field: str) -> np.ndarray:
        nx, nu, N = self.spec.nx, self.spec.nu, self.N
        if field in ("x", "u", "pi"):
            return self.engine.get(field, stage, 1)[0].cpu().numpy()
        nbx, ns = self.engine.nbx, len(self.spec.idxsbx)
        if field in ("sl", "su"):
            # slack values of the soft state bounds = slack of the rows -sl <= 0 / -su <= 0 (stages 1..N-1)
            if ns == 0 or stage == 0 or stage == N:
                return np.zeros(0)
            v = self.engine.get("t", stage, 1)[0].cpu().numpy()
            o = 2 * (nu + nbx + self.engine.ng) + (0 if field == "sl" else ns)
            return v[o:o + ns].copy()
        if field in ("lam", "t"):
            # acados order within a stage: [lbu, lbx, ubu, ubx, lsbx, usbx] (rlmpc/common/utils.py:4-25); the
            # engine stores [lbu(nu), lbx(nbx), ubu(nu), ubx(nbx), lsbx(ns), usbx(ns)] for every stage 0..N
            if stage == N and len(self.acados_ocp.constraints.idxbx_e) == 0:
                return np.zeros(0)
            v = self.engine.get(field, stage, 1)[0].cpu().numpy()
            ng = self.engine.ng
            nv = nu + nbx + ng  # engine order per side: [u, x[bx], h]
            lo_u, lo_x, lo_h = v[:nu], v[nu:nu + nbx], v[nu + nbx:nv]
            up_u, up_x, up_h = v[nv:nv + nu], v[nv + nu:nv + nu + nbx], v[nv + nu + nbx:2 * nv]
            soft = v[2 * nv:]
            if stage == N:
                if len(self.acados_ocp.constraints.idxbx_e) == 0:
                    return np.zeros(0)
                return np.concatenate([lo_x, up_x])
            if stage > 0:
                return np.concatenate([lo_u, lo_x, lo_h, up_u, up_x, up_h, soft])
            # stage 0 carries the x_0 (and, in Q-mode, u_0) equalities as two opposing bounds on all nx / nu
            rho_x = self.engine.get("rho_x0", 0, 1)[0].cpu().numpy()
            if self.qmode:
                rho_u = self.engine.get("rho_u0", 0, 1)[0].cpu().numpy()
                lo_u, up_u = (np.maximum(rho_u, 0.0), np.maximum(-rho_u, 0.0)) if field == "lam" else (np.zeros(nu), np.zeros(nu))
            lo_x, up_x = (np.maximum(rho_x, 0.0), np.maximum(-rho_x, 0.0)) if field == "lam" else (np.zeros(nx), np.zeros(nx))
            return np.concatenate([lo_u, lo_x, lo_h, up_u, up_x, up_h])
        raise NotImplementedError(field)

    # ---- iterate hand-off (examples/chain_mass.py:119-120) ----
    def _acados_to_engine_rows(self, stage: int, v: np.ndarray) -> np.ndarray:
        """Inverse of the row map of ``get(stage, "lam" / "t")``: acados stage layout -> the engine's row order.  The
        stage-0 rows of the x_0 (and Q-mode u_0) equalities have no engine counterpart and are dropped."""
        nx, nu, N = self.spec.nx, self.spec.nu, self.N
        nbx, ns, ng = self.engine.nbx, len(self.spec.idxsbx), self.engine.ng
        nv = nu + nbx + ng
        out = np.zeros(self.engine.nrows)
        if stage == N:
            nbe = len(self.acados_ocp.constraints.idxbx_e)
            if len(v) != 2 * nbe:
                raise ValueError(f"stage {stage}: expected {2 * nbe} rows, got {len(v)}")
            if nbe:
                out[nu:nu + nbx], out[nv + nu:nv + nu + nbx] = v[:nbe], v[nbe:]
            return out
        if stage == 0:
            if len(v) != 2 * (nu + nx + ng):
                raise ValueError(f"stage 0: expected {2 * (nu + nx + ng)} rows (lbu, lbx_0, lh, ubu, ubx_0, uh), got {len(v)}")
            h = nu + nx + ng
            out[:nu] = v[:nu]; out[nu + nbx:nv] = v[nu + nx:h]
            out[nv:nv + nu] = v[h:h + nu]; out[nv + nu + nbx:2 * nv] = v[h + nu + nx:2 * h]
            return out
        if len(v) != 2 * nv + 2 * ns:
            raise ValueError(f"stage {stage}: expected {2 * nv + 2 * ns} rows, got {len(v)}")
        return np.asarray(v, dtype=np.float64).copy()  # same order: [lbu lbx lh ubu ubx uh lsbx usbx]

    def store_iterate(self, filename: str = "iterate.json", overwrite: bool = False, verbose: bool = True) -> None:
        """acados' iterate JSON: x_k, u_k, pi_k and lam_k / t_k in acados' per-stage layout (what get(k, "lam") returns,
        stage 0 with the rows of the x_0 equalities), so that a real acados install can load_iterate() the file."""
        if os.path.exists(filename) and not overwrite:
            raise FileExistsError(filename)
        d = {}
        for k in range(self.N + 1):
            d[f"x_{k}"] = self.get(k, "x").tolist()
            if k < self.N:
                d[f"u_{k}"] = self.get(k, "u").tolist()
                d[f"pi_{k}"] = self.get(k, "pi").tolist()
            for f in ("lam", "t"):
                d[f"{f}_{k}"] = self.get(k, f).tolist()
        with open(filename, "w") as fh:
            json.dump(d, fh, indent=1)

    def load_iterate(self, filename: str, verbose: bool = True) -> None:
        with open(filename) as fh:
            d = json.load(fh)
        dims = {"x": self.spec.nx, "u": self.spec.nu, "pi": self.spec.nx}
        for key, val in d.items():
            f, k = key.rsplit("_", 1)
            k = int(k)
            v = np.asarray(val, dtype=np.float64).reshape(-1)
            if f in dims:
                if len(v) != dims[f]:
                    raise ValueError(f"{key}: expected {dims[f]} values, got {len(v)}")
            elif f in ("lam", "t"):
                if self.engine.nrows == 0 or (k == self.N and self.spec.model == 4):
                    continue
                if k == self.N and len(v) == 0 and len(self.acados_ocp.constraints.idxbx_e) == 0:
                    continue
                v = self._acados_to_engine_rows(k, v)
            else:
                continue  # sl / su / z entries of acados files: slack values are rows of t here
            self.engine.put(f, k, torch.tensor(v.reshape(1, -1), dtype=torch.float64, device=self._dev))
        self._iterate_written = True
        one = torch.ones(1, 1, dtype=torch.float64, device=self._dev)
        try:
            self.engine.put("meta", 0, one)  # the loaded multipliers are a valid interior-point warm start
        except RuntimeError:
            pass
