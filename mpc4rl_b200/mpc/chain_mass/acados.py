"""Chain-of-masses MPC with the constructor of the reference (rlmpc/mpc/chain_mass/acados.py:19-31):
``AcadosMPC(param, discount_factor=0.99)`` with ``param = get_chain_params()`` (ocp_utils.py:319-341).

The reference builds export_parametric_ocp(param, integrator_type="DISCRETE") and hands it to acados; here the same
problem (ERK4 with two sub-steps, EXTERNAL cost with Q, R inside p, |u| <= 1, tol = nlp_tol, nlp_iter SQP iterations)
runs on the warp-cooperative CUDA engine (csrc/chain/)."""
from __future__ import annotations

from ...problems import chain_mass_spec
from ..common.mpc import MPC
from ..nlp import NLP
from ..ocp_solver import OcpSolverShim


class AcadosMPC(MPC):
    def __init__(self, param: dict, discount_factor: float = 0.99, device: int = 0):
        super().__init__()
        spec = chain_mass_spec(param, gamma=discount_factor)
        self.spec = spec
        self.ocp_solver = OcpSolverShim(spec, device=device, max_iter=int(param.get("nlp_iter", 50)),
                                        tol=float(param.get("nlp_tol", 1e-5)))
        self.ocp = self.ocp_solver.acados_ocp
        self.ocp.constraints.x0 = spec.x_ss.copy()  # ocp.constraints.x0 = x_ss (ocp_utils.py:289)
        self.nlp = NLP(spec)
        self.set_discount_factor(discount_factor)
