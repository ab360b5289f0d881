"""Value holder mirroring the reference's ``NLP`` object (rlmpc/mpc/nlp.py:213-430) and the
``update_nlp`` entry point (nlp.py:1341-1563).

The reference builds a CasADi mirror of the whole OCP and, per sample, evaluates ~15 CasADi
functions and factorises a dense (nz x nz) KKT Jacobian with SuperLU.  Here ``update_nlp`` asks the
engine for the same quantities (one fused CUDA kernel: stage sweep for dL/dp, exact-Hessian Riccati
factorisation + adjoint solves for dpi/dp) and stores them under the same attribute names, so
``mpc.nlp.dL_dp.val.full()``, ``mpc.nlp.dpi_dp.val``, ``mpc.nlp.p.val.cat.full()`` etc. keep working.
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np


class DM(np.ndarray):
    """numpy array that also answers the little bit of CasADi DM/struct API the drivers use."""

    def __new__(cls, a):
        return np.asarray(a, dtype=np.float64).view(cls)

    def full(self):
        return np.asarray(self)

    @property
    def cat(self):
        return self


class NLPEntry:
    def __init__(self):
        self.sym = None
        self.val = None
        self.fun = None


class NLP:
    """Same attribute names as the reference's NLP (only ``.val`` is populated)."""

    def __init__(self, spec):
        self.spec = spec
        self.dims = SimpleNamespace(N=spec.N, nx=spec.nx, nu=spec.nu, np=spec.np_model)
        for name in ("cost", "vars", "w", "x", "u", "z", "p", "g", "pi", "h", "t", "lam", "L", "dL_dw", "dL_dp",
                     "dL_du", "dL_dx", "R", "dR_dp", "dR_dz", "dpi_dp"):
            setattr(self, name, NLPEntry())
        self.p.val = DM(spec.p_nominal)
        self.dL_dp.val = DM(np.zeros((1, spec.ntheta)))
        self.dpi_dp.val = np.zeros((spec.nu, spec.ntheta))
        self.constants = {"gamma": spec.gamma}
        self.residuals = np.full(4, np.nan)  # [stat, eq, ineq, comp]

    # parameter struct access (nlp.py:311-318)
    def set_parameter(self, field_: str, value_) -> None:
        sl, shape = self.spec.p_slices()[field_]
        v = np.asarray(value_, dtype=np.float64)
        p = np.array(self.p.val)
        p[sl] = v.T.reshape(-1) if v.ndim == 2 else v.reshape(-1)
        self.p.val = DM(p)

    def get_parameter(self, field_: str) -> DM:
        sl, shape = self.spec.p_slices()[field_]
        v = np.asarray(self.p.val)[sl]
        return DM(v.reshape(shape[::-1]).T if len(shape) == 2 else v)

    def set_constant(self, field_: str, value_) -> None:
        self.constants[field_] = value_

    # residual helpers (nlp.py:386-415)
    def get_cost(self):
        return self.cost.val

    def get_residuals(self):
        return list(self.residuals)

    def assert_kkt_residual(self, tol: float = 1e-6) -> bool:
        """nlp.py:282-283 -> 1295-1299: ||R||_inf <= tol."""
        assert np.all(self.residuals <= tol), f"KKT residual [stat, eq, ineq, comp] = {self.residuals} exceeds {tol}"
        return True


def update_nlp(nlp: NLP, ocp_solver, print_level: int = 0):
    """Refresh the NLP values from the solver's current primal-dual point (nlp.py:1341-1563)."""
    t0 = time.time()
    dL, dpi, status = ocp_solver.evaluate()
    dt = time.time() - t0
    nlp.dL_dp.val = DM(dL.reshape(1, -1))
    nlp.dpi_dp.val = np.array(dpi)
    nlp.cost.val = ocp_solver.get_cost()
    nlp.L.val = nlp.cost.val  # at a KKT point lam.h = 0 and g = 0, so L == cost (to the residual level)
    nlp.residuals = np.array(ocp_solver._res)
    stat, eq, ineq, comp = nlp.residuals
    # the reference's consistency assertions (nlp.py:1513-1537), same thresholds
    assert eq <= 1e-4, f"Equality constraints are not satisfied. g_inf_norm = {eq} >= 1e-4"
    assert ineq < 1e-6, "Inequality constraints are not satisfied."
    assert comp <= 1e-5, "Complementary slackness not satisfied."
    assert stat <= 1e-3, f"Stationarity not satisfied. dL_dw_inf_norm = {stat} >= 1e-3"
    # one fused kernel produces what the reference times in three phases (nlp.py:1397-1422)
    timing = {"dL_dp": dt, "lin_params": 0.0, "solve_params": 0.0, "fused_kernel": dt}
    if print_level > 0:
        print("Cost", nlp.cost.val, "residuals [stat, eq, ineq, comp]", nlp.residuals)
    return nlp, timing


def get_state_labels(ocp) -> list:
    return list(ocp.model.x_labels)


def get_input_labels(ocp) -> list:
    return list(ocp.model.u_labels)


def get_parameter_labels(ocp) -> list:
    return list(ocp.model.p_labels)
