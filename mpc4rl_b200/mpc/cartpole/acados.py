"""Cart-pole swing-up MPC with the constructor of the reference
(rlmpc/mpc/cartpole/acados.py:162-249): ``AcadosMPC(config["mpc"], build=True)``.
No code generation / gcc step happens here: the device model is part of librlmpc_b200.so."""
from __future__ import annotations

import numpy as np

from ...problems import cartpole_spec
from ..common.mpc import MPC
from ..nlp import NLP
from ..ocp_solver import OcpSolverShim


class AcadosMPC(MPC):
    def __init__(self, config: dict, build: bool = True, device: int = 0):
        super().__init__()
        spec = cartpole_spec(config)
        self.spec = spec
        self.nlp = NLP(spec)
        opts = config.get("ocp_options", {})
        self.ocp_solver = OcpSolverShim(spec, device=device, max_iter=int(opts.get("nlp_solver_max_iter", 100)),
                                        tol=float(opts.get("nlp_solver_tol", 1e-6)))
        self.ocp = self.ocp_solver.acados_ocp
        if "x0" in config.get("constraints", {}):
            x0 = np.asarray(config["constraints"]["x0"], dtype=float)
            self.ocp_solver.set(0, "lbx", x0)
            self.ocp_solver.set(0, "ubx", x0)

    def get_predicted_state_trajectory(self) -> np.ndarray:
        return np.stack([self.ocp_solver.get(i, "x") for i in range(self.spec.N + 1)])

    def get_predicted_control_trajectory(self) -> np.ndarray:
        return np.stack([self.ocp_solver.get(i, "u") for i in range(self.spec.N)])

    def get_action(self, x0: np.ndarray) -> np.ndarray:
        """Action rescaled to [-1, 1] for gym (cartpole/acados.py:239-249)."""
        return self.scale_action(super().get_action(x0))
