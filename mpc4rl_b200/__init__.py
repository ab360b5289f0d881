"""mpc4rl_b200 -- B200-native batched MPC-as-function-approximator engine (hot path of rlmpc).

Layout
  csrc/       CUDA kernels + C ABI (include/rlmpc_b200.h), built in-tree to librlmpc_b200.so
  _cabi.py    ctypes binding
  batched.py  BatchedMPC: torch-tensor API for whole minibatches
  problems.py host-side problem descriptions (the AcadosOcp role)
  mpc/        mirror of the reference's rlmpc.mpc package (MPC, AcadosMPC classes)
"""
__all__ = ["BatchedMPC", "cartpole_spec", "cartpole_original_config", "cartpole_config", "linear_system_spec",
           "linear_system_param_nominal", "evaporation_spec", "chain_mass_spec", "get_chain_params"]


def __getattr__(name):
    if name == "BatchedMPC":
        from .batched import BatchedMPC
        return BatchedMPC
    if name in ("cartpole_spec", "cartpole_original_config", "cartpole_config", "ProblemSpec", "linear_system_spec",
                "linear_system_param_nominal", "evaporation_spec", "chain_mass_spec", "get_chain_params", "chain_define_x0"):
        from . import problems
        return getattr(problems, name)
    raise AttributeError(name)
