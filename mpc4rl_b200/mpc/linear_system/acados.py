"""Linear-system MPC with the constructor of the reference
(rlmpc/mpc/linear_system/acados.py:12-24): ``AcadosMPC(param, discount_factor=0.99)``.
``param`` is the dict of tests/test_linear_example.py:9-17 (A, B, Q, R, b, f, V_0)."""
from __future__ import annotations

from ...problems import linear_system_spec
from ..common.mpc import MPC
from ..nlp import NLP
from ..ocp_solver import OcpSolverShim


class AcadosMPC(MPC):
    def __init__(self, param: dict, discount_factor: float = 0.99, device: int = 0):
        super().__init__()
        spec = linear_system_spec(param, gamma=discount_factor)
        self.spec = spec
        self.ocp_solver = OcpSolverShim(spec, device=device, max_iter=100, tol=1e-6)
        self.ocp = self.ocp_solver.acados_ocp
        self.nlp = NLP(spec)
        self.set_discount_factor(discount_factor)
