"""``MPC`` base class with the method surface of the reference (rlmpc/mpc/common/mpc.py:8-414),
re-implemented over the B200 engine.  Scalar calls run as a batch of one sample on the GPU; for
throughput use ``mpc.batched(max_batch)`` (``mpc4rl_b200.BatchedMPC``), which shares the problem
definition and the current parameters.
"""
from __future__ import annotations

from abc import ABC

import numpy as np

from ..nlp import NLP, get_input_labels, get_parameter_labels, get_state_labels, update_nlp
from ..ocp_solver import OcpSolverShim


class MPC(ABC):
    ocp_solver: OcpSolverShim
    nlp: NLP

    def __init__(self, gamma: float = 1.0):
        super().__init__()
        self.discount_factor = gamma
        self.nlp_timing = {}
        self.status = 0

    # ---- policy / value evaluation (mpc.py:27-96, 177-202) ----
    def get_parameters(self) -> np.ndarray:
        return self.get_p()

    def get_action(self, x0: np.ndarray) -> np.ndarray:
        """pi(s): fix x_0, solve, return u_0.  Like the reference, the status is stored, not checked."""
        self.ocp_solver.set(0, "lbx", x0)
        self.ocp_solver.set(0, "ubx", x0)
        self.status = self.ocp_solver.solve()
        return self.ocp_solver.get(0, "u")

    def q_update(self, x0: np.ndarray, u0: np.ndarray) -> int:
        """Q(s,a): additionally clamp u_0 = a, solve, refresh the sensitivities, restore the bounds."""
        self.ocp_solver.set(0, "lbx", x0)
        self.ocp_solver.set(0, "ubx", x0)
        self.ocp_solver.set(0, "u", u0)
        self.ocp_solver.constraints_set(0, "lbu", u0)
        self.ocp_solver.constraints_set(0, "ubu", u0)
        try:
            status = self.ocp_solver.solve()
            if status != 0:
                raise RuntimeError(f"Solver failed q_update with status {status}. Exiting.")
            self.nlp, _ = update_nlp(self.nlp, self.ocp_solver)
        finally:
            self.ocp_solver.constraints_set(0, "lbu", self.ocp_solver.acados_ocp.constraints.lbu)
            self.ocp_solver.constraints_set(0, "ubu", self.ocp_solver.acados_ocp.constraints.ubu)
        return status

    def update(self, x0: np.ndarray) -> int:
        """V(s): fix x_0 and solve.  Does NOT refresh the sensitivities (reference quirk Q6)."""
        self.ocp_solver.set(0, "lbx", x0)
        self.ocp_solver.set(0, "ubx", x0)
        status = self.ocp_solver.solve()
        if status != 0:
            raise RuntimeError(f"Solver failed update with status {status}. Exiting.")
        return status

    def update_nlp(self) -> None:
        self.nlp, self.nlp_timing = update_nlp(self.nlp, self.ocp_solver)

    # ---- getters (mpc.py:104-135, 316-351, 402-414) ----
    def get_dL_dp(self) -> np.ndarray:
        return self.nlp.dL_dp.val.full()

    def get_dV_dp(self) -> np.ndarray:
        return self.get_dL_dp()

    def get_dQ_dp(self) -> np.ndarray:
        return self.get_dL_dp()

    def get_L(self) -> float:
        return float(self.nlp.L.val)

    def get_V(self) -> float:
        return self.ocp_solver.get_cost()

    def get_Q(self) -> float:
        return self.ocp_solver.get_cost()

    def get_pi(self) -> np.ndarray:
        return self.ocp_solver.get(0, "u")

    def get_dpi_dp(self, finite_differences: bool = False, idx: int = 0) -> np.ndarray:
        if not finite_differences:
            return self.nlp.dpi_dp.val
        return self.compute_dpi_dp_finite_differences(self.get_p(), idx=idx)

    def compute_dpi_dp_finite_differences(self, p: np.ndarray, idx: int = None, delta: float = 1e-4) -> np.ndarray:
        """Forward differences of pi wrt p at x0 = the current lbx_0 (mpc.py:353-400)."""
        pi0 = self.get_pi().copy()
        p0 = self.get_p()
        x0 = self.ocp_solver.acados_ocp.constraints.lbx_0.copy()
        nu, nparam = self.ocp_solver.acados_ocp.dims.nu, p0.shape[0]
        dpi_dp = np.zeros((nu, nparam))
        for i in (range(nparam) if idx is None else [idx]):
            pplus = p0.copy()
            pplus[i] += delta
            self.set_p(pplus)
            self.update(x0)
            dpi_dp[:, i] = (self.get_pi() - pi0) / delta
        self.set_p(p0)  # (the reference leaves the last perturbed parameter in the solver)
        return dpi_dp

    # ---- parameters (mpc.py:137-160, 212-257) ----
    def set_p(self, p: np.ndarray, finite_differences: bool = False) -> None:
        """Whole parameter vector p (layout nlp.py:970-989) on every stage."""
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        if p.shape[0] != self.nlp.spec.ntheta:
            raise ValueError(f"p must have {self.nlp.spec.ntheta} entries (the full NLP parameter vector)")
        self.ocp_solver.set_theta(p)
        if not finite_differences:
            from ..nlp import DM

            self.nlp.p.val = DM(p)

    def set(self, stage, field, value, finite_differences: bool = False):
        if field == "p":
            self.set_p(value, finite_differences=finite_differences)  # one theta shared by all stages
        else:
            self.ocp_solver.set(stage, field, value)

    def set_parameter(self, value_, api="new") -> None:
        """mpc.py:233-257: pushes model parameters and W_0/W/yref_0/yref; here theta is one vector."""
        self.set_p(value_)

    def get_p(self) -> np.ndarray:
        return np.asarray(self.nlp.p.val).reshape(-1).copy()

    def get_parameter_values(self) -> np.ndarray:
        return self.get_p()

    def get_parameter_labels(self) -> list:
        return get_parameter_labels(self.ocp_solver.acados_ocp)

    def get_state_labels(self) -> list:
        return get_state_labels(self.ocp_solver.acados_ocp)

    def get_input_labels(self) -> list:
        return get_input_labels(self.ocp_solver.acados_ocp)

    # ---- iterate / discount (mpc.py:204-210, 259-285) ----
    def reset(self, x0: np.ndarray):
        self.ocp_solver.reset()
        self.set_discount_factor(self.discount_factor)
        for stage in range(self.ocp_solver.acados_ocp.dims.N + 1):
            self.ocp_solver.set(stage, "x", x0)

    def set_discount_factor(self, discount_factor_: float) -> None:
        """gamma enters the per-stage cost scaling exactly as in the reference's NLP cost
        (nlp.py:1038-1134).  (acados-side quirk Q3 -- the reference overwrites acados' dT scaling with
        gamma^k -- is not replicated: the engine has a single copy of the scaling.)"""
        self.discount_factor = discount_factor_
        self.nlp.set_constant("gamma", discount_factor_)
        self.ocp_solver.engine.set_cost_scaling(self.nlp.spec.cost_scaling(discount_factor_))

    def get(self, stage, field):
        return self.ocp_solver.get(stage, field)

    # ---- action scaling (mpc.py:290-314) ----
    def scale_action(self, action: np.ndarray) -> np.ndarray:
        low, high = self.ocp_solver.acados_ocp.constraints.lbu, self.ocp_solver.acados_ocp.constraints.ubu
        return 2.0 * ((action - low) / (high - low)) - 1.0

    def unscale_action(self, action: np.ndarray) -> np.ndarray:
        low, high = self.ocp_solver.acados_ocp.constraints.lbu, self.ocp_solver.acados_ocp.constraints.ubu
        return 0.5 * (high - low) * (action + 1.0) + low

    # ---- new: the batched engine for this problem ----
    def batched(self, max_batch: int, device: int = 0):
        from ...batched import BatchedMPC

        eng = BatchedMPC(self.nlp.spec, max_batch=max_batch, device=device)
        eng.set_theta(self.get_p())
        eng.set_cost_scaling(self.nlp.spec.cost_scaling(self.discount_factor))
        return eng
