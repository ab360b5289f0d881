"""Evaporation-process economic NMPC with the constructor of the reference
(rlmpc/mpc/evaporation_process/acados.py:89-139):
``AcadosMPC(model_param, cost_param, x0, u0, gamma, H)``; ``reset()`` puts every stage back on the
steady state (the model needs a non-zero initial guess), ``get_action`` raises on a failed solve."""
from __future__ import annotations

import numpy as np

from ...problems import evaporation_spec
from ..common.mpc import MPC
from ..nlp import NLP
from ..ocp_solver import OcpSolverShim


class AcadosMPC(MPC):
    def __init__(self, model_param: dict = None, cost_param: dict = None, x0: np.ndarray = np.array([25, 49.743]),
                 u0: np.ndarray = np.array([191.713, 215.888, 0.0]), gamma: float = 1.0, H: np.ndarray = None,
                 device: int = 0):
        super().__init__()
        self.gamma = gamma
        self.discount_factor = gamma
        spec = evaporation_spec(model_param=model_param, cost_param=cost_param, gamma=gamma, H=H)
        self.spec = spec
        self.ocp_solver = OcpSolverShim(spec, device=device, max_iter=100, tol=1e-6)
        self.ocp = self.ocp_solver.acados_ocp
        self.nlp = NLP(spec)
        self.u0 = np.asarray(u0, dtype=float)
        self._x_guess = np.asarray(x0, dtype=float)
        self._set_guess(self._x_guess, self.u0)

    def _set_guess(self, x, u):
        for stage in range(self.spec.N + 1):
            self.ocp_solver.set(stage, "x", x)
        for stage in range(self.spec.N):
            self.ocp_solver.set(stage, "u", u)

    def get_action(self, x0: np.ndarray) -> np.ndarray:
        action = super().get_action(x0)
        if self.ocp_solver.status != 0:
            raise RuntimeError(f"Solver failed with status {self.ocp_solver.status}. Exiting.")
        return action

    def reset(self, x0: np.ndarray = None):
        """acados.py:130-139: all stages back to the steady state."""
        self.ocp_solver.reset()
        self._set_guess(np.array([25.0, 49.743]), np.array([191.713, 215.888, 0.0]))
