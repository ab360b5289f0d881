"""ctypes loader of the host build of the engine (TEST INFRASTRUCTURE, see oracle/__init__.py).

Not an independent checker (shares engine.cuh with the CUDA product): used to debug the
templates on a box without a GPU and as bench.py's ``cpu_baseline`` of kind "port".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
MAXN, MAXD = 128, 8


class ProblemData(C.Structure):
    """Mirror of rlmpc::ProblemData (mpc4rl_b200/csrc/common.cuh)."""
    _fields_ = [
        ("N", C.c_int), ("mode", C.c_int), ("max_sqp", C.c_int), ("max_ipm", C.c_int),
        ("warm_ipm", C.c_int), ("param_cost", C.c_int), ("fix0", C.c_int),
        ("tol", C.c_double), ("tau", C.c_double), ("mu0", C.c_double),
        ("sigma_min", C.c_double), ("sigma0", C.c_double), ("as_steps", C.c_double), ("condense", C.c_double), ("comp_accept", C.c_double), ("step_length", C.c_double),
        ("scale", C.c_double * (MAXN + 1)),
        ("lbu", C.c_double * MAXD), ("ubu", C.c_double * MAXD),
        ("lbx", C.c_double * MAXD), ("ubx", C.c_double * MAXD),
        ("lbx_e", C.c_double * MAXD), ("ubx_e", C.c_double * MAXD),
        ("zl", C.c_double * MAXD), ("zu", C.c_double * MAXD),
        ("lg", C.c_double * MAXD), ("ug", C.c_double * MAXD),
        ("mc", C.c_double * 24),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_DIR, "libcpu_port.so")
    # make tracks staleness against engine.cuh / the model headers; on a box without the sources'
    # toolchain the prebuilt .so is used as is
    r = subprocess.run(["make", "-C", _DIR, "-s"] + (["-B"] if force else []), capture_output=True, text=True)
    if r.returncode != 0 and not os.path.exists(so):
        raise RuntimeError("building oracle/cpu_port failed:\n" + r.stdout + r.stderr)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        try:
            _lib = C.CDLL(so)
        except OSError:
            _lib = C.CDLL(build(force=True))
        assert _lib.cpu_port_sizeof_problem_data() == C.sizeof(ProblemData), "ProblemData layout mismatch"
    return _lib


def make_pd(N, scale, lbu, ubu, mc, tol=1e-6, tau=1e-8, mu0=1.0, max_ipm=50, warm_ipm=0, param_cost=0,
            lbx=(), ubx=(), lbx_e=(), ubx_e=(), sigma_min=0.05, sigma0=0.3, zl=(), zu=(), as_steps=20, lg=(), ug=(), condense=0, comp_accept=0.2, step_length=1.0) -> ProblemData:
    pd = ProblemData()
    pd.sigma_min = sigma_min; pd.sigma0 = sigma0; pd.as_steps = as_steps; pd.condense = condense; pd.comp_accept = comp_accept; pd.step_length = step_length
    pd.N = N; pd.max_ipm = max_ipm; pd.warm_ipm = warm_ipm; pd.param_cost = param_cost
    pd.tol = tol; pd.tau = tau; pd.mu0 = mu0
    for i, v in enumerate(scale):
        pd.scale[i] = v
    for i in range(MAXD):
        pd.lbx[i] = pd.lbx_e[i] = pd.lg[i] = -1e30
        pd.ubx[i] = pd.ubx_e[i] = pd.ug[i] = 1e30
    for i, v in enumerate(lbu):
        pd.lbu[i] = v
    for i, v in enumerate(ubu):
        pd.ubu[i] = v
    for name, vals in (("lbx", lbx), ("ubx", ubx), ("lbx_e", lbx_e), ("ubx_e", ubx_e), ("zl", zl), ("zu", zu),
                       ("lg", lg), ("ug", ug)):
        for i, v in enumerate(vals):
            getattr(pd, name)[i] = v
    for i, v in enumerate(mc):
        pd.mc[i] = v
    return pd


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def unit(model: int, pd: ProblemData, mode: int, max_sqp: int, theta, x0, u0=None, iterate=None, nx=4, nu=1,
         do_solve=True, do_sens=True, threads=0):
    """Run solve(+sens) for a batch on the host.  Returns a dict of row-major outputs and the iterate."""
    L = lib()
    L.cpu_port_set_threads(int(threads))
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    B = x0.shape[0]
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    per_sample = int(theta.ndim == 2)
    nth = theta.shape[-1]
    itsz = L.cpu_port_iterate_size(model, pd.N)
    if iterate is None:  # MPC.reset: all stages = x0
        iterate = np.zeros((itsz, B))
        for k in range(pd.N + 1):
            iterate[k * nx:(k + 1) * nx, :] = x0.T
    iterate = np.ascontiguousarray(iterate)
    u0a = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64).reshape(B, nu)
    ng = L.cpu_port_grad_width(C.c_int(model), C.byref(pd))  # model parameters only unless pd.param_cost
    out = dict(u0=np.zeros((B, nu)), cost=np.zeros(B), status=np.zeros(B, dtype=np.int32), dL=np.zeros((B, ng)),
               dpi=np.zeros((B, nu, ng)), res=np.zeros((B, 4)), iters=np.zeros((B, 3), dtype=np.int32))
    r = L.cpu_port_unit(C.c_int(model), C.byref(pd), C.c_int(mode), C.c_int(max_sqp), C.c_int(B), _p(theta),
                        C.c_int(per_sample), _p(x0), _p(u0a), _p(iterate), C.c_int(int(do_solve)), C.c_int(int(do_sens)),
                        _p(out["u0"]), _p(out["cost"]), _p(out["status"], C.c_int), _p(out["dL"]), _p(out["dpi"]),
                        _p(out["res"]), _p(out["iters"], C.c_int))
    assert r == 0
    out["iterate"] = iterate
    return out


# ---- chain of masses: host run of the warp-cooperative engine under the fiber SIMT emulation (chain_port.cpp) ----
_chain = None


def chain_lib():
    global _chain
    if _chain is None:
        so = os.path.join(_DIR, "libchain_port.so")
        r = subprocess.run(["make", "-C", _DIR, "-s", "libchain_port.so"], capture_output=True, text=True)
        if r.returncode != 0 and not os.path.exists(so):
            raise RuntimeError("building oracle/cpu_port/libchain_port.so failed:\n" + r.stdout + r.stderr)
        _chain = C.CDLL(so)
        assert _chain.chain_port_sizeof_problem_data() == C.sizeof(ProblemData), "ProblemData layout mismatch"
    return _chain


def chain_unit(n_mass: int, pd: ProblemData, mode: int, max_sqp: int, theta, xss, x0, u0=None, iterate=None,
               do_solve=True, do_sens=True, threads=0):
    """solve(+sens) of a batch of chain-mass samples on the host.  iterate: [B, it_size] (contiguous per sample)."""
    L = chain_lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    B = x0.shape[0]
    nx, nth, its = C.c_int(), C.c_int(), C.c_int()
    assert L.chain_port_dims(n_mass, pd.N, C.byref(nx), C.byref(nth), C.byref(its)) == 0
    nx, nth, its = nx.value, nth.value, its.value
    if iterate is None:  # MPC.reset: all stages = x0
        iterate = np.zeros((B, its))
        for k in range(pd.N + 1):
            iterate[:, k * nx:(k + 1) * nx] = x0
    iterate = np.ascontiguousarray(iterate)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    xss = np.ascontiguousarray(xss, dtype=np.float64)
    u0a = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64).reshape(B, 3)
    out = dict(u0=np.zeros((B, 3)), cost=np.zeros(B), status=np.zeros(B, dtype=np.int32), dL=np.zeros((B, nth)),
               dpi=np.zeros((B, 3, nth)), res=np.zeros((B, 4)), iters=np.zeros(B, dtype=np.int32))
    r = L.chain_port_run(C.c_int(n_mass), C.byref(pd), C.c_int(mode), C.c_int(max_sqp), C.c_int(B), _p(theta), _p(xss), _p(x0),
                         _p(u0a), _p(iterate), C.c_int(int(do_solve)), C.c_int(int(do_sens)), _p(out["u0"]), _p(out["cost"]),
                         _p(out["status"], C.c_int), _p(out["dL"]), _p(out["dpi"]), _p(out["res"]), _p(out["iters"], C.c_int),
                         C.c_int(int(threads)))
    assert r == 0
    out["iterate"] = iterate
    return out
