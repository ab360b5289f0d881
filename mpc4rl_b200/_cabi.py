"""ctypes binding of librlmpc_b200.so (the C ABI in include/rlmpc_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises on import of the
symbols, and ``rlmpc_create`` fails with RLMPC_ENODEV on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

MAXN, MAXD = 128, 8
MODEL_CARTPOLE = 1
MODEL_LINEAR_SYSTEM = 2
MODEL_EVAPORATION = 3
MODEL_CHAIN_MASS = 4
MODE_V, MODE_Q = 0, 1

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RLMPC_B200_LIB", os.path.join(_PKG, "librlmpc_b200.so"))  # env override: tuning variants

# every symbol include/rlmpc_b200.h declares
SYMBOLS = [
    "rlmpc_create", "rlmpc_destroy", "rlmpc_last_error", "rlmpc_dims", "rlmpc_nrows", "rlmpc_set_theta",
    "rlmpc_set_cost_scaling", "rlmpc_set_bounds", "rlmpc_set_option", "rlmpc_reset", "rlmpc_reset_masked",
    "rlmpc_get_iterate", "rlmpc_store_bytes", "rlmpc_store_copy",
    "rlmpc_put_iterate", "rlmpc_solve", "rlmpc_sens", "rlmpc_solve_sens", "rlmpc_solve_sens_host",
    "rlmpc_td_grad", "rlmpc_launch_count", "rlmpc_get_timings", "rlmpc_cartpole_env_step",
    "rlmpc_set_theta_dev", "rlmpc_set_model_vector", "rlmpc_fp64_peak", "rlmpc_fp64_tensor_peak",
]


class ProblemDesc(C.Structure):
    """struct rlmpc_problem_desc"""
    _fields_ = [
        ("model", C.c_int), ("N", C.c_int),
        ("scale", C.c_double * (MAXN + 1)),
        ("lbu", C.c_double * MAXD), ("ubu", C.c_double * MAXD),
        ("lbx", C.c_double * MAXD), ("ubx", C.c_double * MAXD),
        ("lbx_e", C.c_double * MAXD), ("ubx_e", C.c_double * MAXD),
        ("model_const", C.c_double * 24),
        ("zl", C.c_double * MAXD), ("zu", C.c_double * MAXD),
        ("lg", C.c_double * MAXD), ("ug", C.c_double * MAXD),
    ]


_lib = None


def load():
    """Load the shared library; raises RuntimeError (never falls back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). mpc4rl_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    dp, ip, vp, cp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_char_p
    H = C.c_void_p
    lib.rlmpc_create.argtypes = [C.POINTER(ProblemDesc), C.c_int, C.c_int, C.POINTER(H)]
    lib.rlmpc_destroy.argtypes = [H]; lib.rlmpc_destroy.restype = None
    lib.rlmpc_last_error.restype = C.c_char_p
    lib.rlmpc_dims.argtypes = [H, ip, ip, ip, ip, ip]
    lib.rlmpc_nrows.argtypes = [H]
    lib.rlmpc_set_theta.argtypes = [H, vp, C.c_int, C.c_int]
    lib.rlmpc_set_theta_dev.argtypes = [H, vp, C.c_int, C.c_int, vp]
    lib.rlmpc_set_model_vector.argtypes = [H, cp, vp, C.c_int]
    lib.rlmpc_fp64_peak.argtypes = [C.c_int, dp]
    lib.rlmpc_fp64_tensor_peak.argtypes = [C.c_int, dp]
    lib.rlmpc_set_cost_scaling.argtypes = [H, vp, C.c_int]
    lib.rlmpc_set_bounds.argtypes = [H, cp, vp, C.c_int]
    lib.rlmpc_set_option.argtypes = [H, cp, C.c_double]
    lib.rlmpc_reset.argtypes = [H, C.c_int, vp, vp]
    lib.rlmpc_reset_masked.argtypes = [H, C.c_int, vp, vp, vp]
    lib.rlmpc_store_bytes.argtypes = [H, C.c_int]; lib.rlmpc_store_bytes.restype = C.c_size_t
    lib.rlmpc_store_copy.argtypes = [H, C.c_int, vp, vp, C.c_int, C.c_int, vp]
    lib.rlmpc_get_iterate.argtypes = [H, cp, C.c_int, C.c_int, vp, vp]
    lib.rlmpc_put_iterate.argtypes = [H, cp, C.c_int, C.c_int, vp, vp]
    lib.rlmpc_solve.argtypes = [H, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.rlmpc_sens.argtypes = [H, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.rlmpc_solve_sens.argtypes = [H, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rlmpc_solve_sens_host.argtypes = [H, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rlmpc_td_grad.argtypes = [H, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.rlmpc_get_timings.argtypes = [H, vp, C.c_int]
    lib.rlmpc_cartpole_env_step.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.rlmpc_launch_count.argtypes = [H]; lib.rlmpc_launch_count.restype = C.c_longlong
    for name in SYMBOLS:
        f = getattr(lib, name)
        if name not in ("rlmpc_destroy", "rlmpc_last_error", "rlmpc_launch_count", "rlmpc_store_bytes"):
            f.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().rlmpc_last_error()
        raise RuntimeError(f"rlmpc_b200 error {rc}: {msg.decode() if msg else ''}")
