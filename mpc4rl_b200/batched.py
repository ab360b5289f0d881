"""Batched MPC engine: torch CUDA tensors in, torch CUDA tensors out, stream-ordered, FP64.

This is the *new* API of SURVEY.md 8(b): one call evaluates a whole replay-buffer minibatch
(what the reference does with a Python loop over samples, e.g.
rlmpc/examples/linear_system_mpc_qlearning.py:178-190, rlmpc/td3/policies.py:197).
torch is used for device memory and streams only; all arithmetic happens in the CUDA library
behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _cabi
from .problems import ProblemSpec


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedMPC:
    def __init__(self, spec: ProblemSpec, max_batch: int, device: int | torch.device = 0):
        self.spec = spec
        self.lib = _cabi.load()
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        if dev.type != "cuda":
            raise RuntimeError("BatchedMPC runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: mpc4rl_b200 has no CPU fallback")
        self.device = dev
        self.max_batch = int(max_batch)
        self._h = C.c_void_p()
        desc = spec.to_desc()
        _cabi.check(self.lib.rlmpc_create(C.byref(desc), self.max_batch, dev.index or 0, C.byref(self._h)))
        self.nx, self.nu, self.ntheta = spec.nx, spec.nu, spec.ntheta
        for name, vec in getattr(spec, "model_vectors", {}).items():
            v = np.ascontiguousarray(np.asarray(vec, dtype=np.float64))
            _cabi.check(self.lib.rlmpc_set_model_vector(self._h, name.encode(), v.ctypes.data_as(C.c_void_p), len(v)))
        # inequality rows per stage: 2*(nu+nbx) + 2*ns, acados order [lbu lbx ubu ubx lsbx usbx]
        self.nrows = int(self.lib.rlmpc_nrows(self._h))
        self.ns = len(spec.idxsbx)
        self.ng = len(spec.lh)  # affine general constraint rows (lh/uh)
        self.nbx = (self.nrows - 2 * self.ns) // 2 - self.nu - self.ng
        self._theta_host, self._theta_dev = None, None
        self.set_theta(np.array(spec.p_nominal, dtype=np.float64))
        # cartpole only: gradient rows over the whole p (incl. W, yref) instead of the model parameters.  The
        # other models have a fixed gradient width (all their parameters).
        self.param_cost = bool(spec.parameterize_tracking_cost) and spec.model == _cabi.MODEL_CARTPOLE
        self.set_option("param_cost", 1.0 if self.param_cost else 0.0)

    @property
    def ngrad(self) -> int:
        """Width of the gradient rows: the model parameters (the only entries of the reference's p
        with non-zero gradient when parameterize_tracking_cost=False, quirk Q8) or all of p."""
        ng = C.c_int()
        _cabi.check(self.lib.rlmpc_dims(self._h, None, None, None, C.byref(ng), None))
        return int(ng.value)

    def full_grad(self, g: torch.Tensor) -> torch.Tensor:
        """Pad [..., ngrad] gradient rows to the reference's [..., ntheta] layout."""
        if g.shape[-1] == self.ntheta:
            return g
        return torch.nn.functional.pad(g, (0, self.ntheta - g.shape[-1]))

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.rlmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk_in(self, t: torch.Tensor, cols: int, name: str) -> torch.Tensor:
        if not (t.is_cuda and t.dtype == torch.float64):
            raise TypeError(f"{name} must be a float64 CUDA tensor")
        if t.dim() != 2 or t.shape[1] != cols:
            raise ValueError(f"{name} must have shape [B, {cols}], got {tuple(t.shape)}")
        return t.contiguous()

    # ---- parameters ----
    @property
    def theta(self) -> np.ndarray:
        """The current parameter vector(s) as a host array (read back once after a device-side set_theta)."""
        if self._theta_host is None:
            self._theta_host = self._theta_dev.detach().cpu().numpy()
        return self._theta_host

    def set_theta(self, theta) -> None:
        """theta: [ntheta] shared by the batch, or [B, ntheta] per sample.  A CUDA tensor is handed over on the
        current stream without any host synchronisation (rlmpc_set_theta_dev)."""
        if isinstance(theta, torch.Tensor) and theta.is_cuda:
            th = theta.detach().to(torch.float64).contiguous()
            if th.shape[-1] != self.ntheta:
                raise ValueError(f"theta must have {self.ntheta} entries per sample")
            per = int(th.dim() == 2)
            _cabi.check(self.lib.rlmpc_set_theta_dev(self._h, _ptr(th), per, th.shape[0] if per else 0, self._stream()))
            self._theta_dev = th  # keep the tensor alive until the copy kernel has run; `theta` reads it back lazily
            self._theta_host = None
            return
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        if th.shape[-1] != self.ntheta:
            raise ValueError(f"theta must have {self.ntheta} entries per sample")
        per = int(th.ndim == 2)
        _cabi.check(self.lib.rlmpc_set_theta(self._h, th.ctypes.data_as(C.c_void_p), per, th.shape[0] if per else 0))
        self._theta_host, self._theta_dev = th, None
        self._theta_full_dev = None  # (cache of the autograd bridge)

    def set_option(self, name: str, value: float) -> None:
        _cabi.check(self.lib.rlmpc_set_option(self._h, name.encode(), float(value)))
        if name == "param_cost":
            self.param_cost = bool(value)

    def set_cost_scaling(self, scale) -> None:
        s = np.ascontiguousarray(np.asarray(scale, dtype=np.float64))
        _cabi.check(self.lib.rlmpc_set_cost_scaling(self._h, s.ctypes.data_as(C.c_void_p), len(s)))

    def set_bounds(self, field: str, v) -> None:
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        _cabi.check(self.lib.rlmpc_set_bounds(self._h, field.encode(), a.ctypes.data_as(C.c_void_p), len(a)))

    # ---- iterate ----
    def reset(self, x0: Optional[torch.Tensor] = None, B: Optional[int] = None, mask: Optional[torch.Tensor] = None) -> None:
        """MPC.reset for the batch (mpc.py:204-210); ``mask`` [B] restricts it to some samples."""
        if x0 is not None:
            x0 = self._chk_in(x0, self.nx, "x0")
            B = x0.shape[0]
        if mask is None:
            _cabi.check(self.lib.rlmpc_reset(self._h, int(B), _ptr(x0), self._stream()))
        else:
            m = mask.to(self.device, torch.int32).contiguous()
            _cabi.check(self.lib.rlmpc_reset_masked(self._h, int(B), _ptr(x0), _ptr(m), self._stream()))

    def get(self, field: str, stage: int, B: int) -> torch.Tensor:
        dim = {"x": self.nx, "u": self.nu, "pi": self.nx, "lam": self.nrows, "t": self.nrows,
               "rho_x0": self.nx, "rho_u0": self.nu, "meta": 1}[field]
        out = torch.empty(B, dim, dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.rlmpc_get_iterate(self._h, field.encode(), int(stage), int(B), _ptr(out), self._stream()))
        return out

    def put(self, field: str, stage: int, value: torch.Tensor) -> None:
        dim = {"x": self.nx, "u": self.nu, "pi": self.nx, "lam": self.nrows, "t": self.nrows,
               "rho_x0": self.nx, "rho_u0": self.nu, "meta": 1}.get(field)
        if dim is None:
            raise ValueError(f"unknown iterate field {field!r}")
        if not (value.is_cuda and value.dtype == torch.float64 and value.dim() == 2 and value.shape[1] == dim):
            raise ValueError(f"{field} rows must be float64 CUDA tensors of shape [B, {dim}], got {tuple(value.shape)} {value.dtype}")
        if value.shape[0] > self.max_batch:
            raise ValueError("more rows than max_batch")
        value = value.contiguous()
        _cabi.check(self.lib.rlmpc_put_iterate(self._h, field.encode(), int(stage), value.shape[0], _ptr(value), self._stream()))

    def iterate_store(self, capacity: int) -> "IterateStore":
        """Per-replay-buffer-entry warm starts."""
        return IterateStore(self, capacity)

    # ---- hot path ----
    def solve(self, x0: torch.Tensor, u0: Optional[torch.Tensor] = None, max_sqp: int = 1):
        """V-mode if u0 is None else Q-mode.  Returns (u0 [B,nu], cost [B], status [B] int32)."""
        x0 = self._chk_in(x0, self.nx, "x0")
        B = x0.shape[0]
        mode = _cabi.MODE_V if u0 is None else _cabi.MODE_Q
        if u0 is not None:
            u0 = self._chk_in(u0, self.nu, "u0")
        uo = torch.empty(B, self.nu, dtype=torch.float64, device=self.device)
        cost = torch.empty(B, dtype=torch.float64, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        _cabi.check(self.lib.rlmpc_solve(self._h, mode, int(max_sqp), B, _ptr(x0), _ptr(u0), _ptr(uo), _ptr(cost),
                                         _ptr(status), self._stream()))
        return uo, cost, status

    def sens(self, B: int, qmode: bool = False):
        """update_nlp at the current iterate: (dL_dtheta [B,ngrad], dpi_dtheta [B,nu,ngrad], cost, res [B,4], status)."""
        dL = torch.zeros(B, self.ngrad, dtype=torch.float64, device=self.device)
        dpi = torch.zeros(B, self.nu, self.ngrad, dtype=torch.float64, device=self.device)
        cost = torch.empty(B, dtype=torch.float64, device=self.device)
        res = torch.empty(B, 4, dtype=torch.float64, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        _cabi.check(self.lib.rlmpc_sens(self._h, _cabi.MODE_Q if qmode else _cabi.MODE_V, int(B), _ptr(dL), _ptr(dpi),
                                        _ptr(cost), _ptr(res), _ptr(status), self._stream()))
        return dL, dpi, cost, res, status

    def solve_sens(self, x0: torch.Tensor, u0: Optional[torch.Tensor] = None, max_sqp: int = 1, out: dict | None = None):
        """One fused launch = one unit of the headline metric per sample."""
        x0 = self._chk_in(x0, self.nx, "x0")
        B = x0.shape[0]
        mode = _cabi.MODE_V if u0 is None else _cabi.MODE_Q
        if u0 is not None:
            u0 = self._chk_in(u0, self.nu, "u0")
        if out is None:
            out = self.alloc_outputs(B)
        elif self.param_cost:
            out["dL"].zero_(); out["dpi"].zero_()
        _cabi.check(self.lib.rlmpc_solve_sens(self._h, mode, int(max_sqp), B, _ptr(x0), _ptr(u0), _ptr(out["u0"]),
                                              _ptr(out["cost"]), _ptr(out["status"]), _ptr(out["dL"]), _ptr(out["dpi"]),
                                              _ptr(out["res"]), self._stream()))
        return out

    def alloc_outputs(self, B: int) -> dict:
        f64 = dict(dtype=torch.float64, device=self.device)
        return dict(u0=torch.empty(B, self.nu, **f64), cost=torch.empty(B, **f64),
                    status=torch.empty(B, dtype=torch.int32, device=self.device),
                    dL=torch.zeros(B, self.ngrad, **f64), dpi=torch.zeros(B, self.nu, self.ngrad, **f64),
                    res=torch.empty(B, 4, **f64))

    def alloc_host_outputs(self, B: int, pinned: bool = True) -> dict:
        """Host result buffers for ``solve_sens_host``; page-locked ones are DMA targets (no staging copy)."""
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pinned).numpy()
        return dict(u0=mk((B, self.nu), torch.float64), cost=mk((B,), torch.float64), status=mk((B,), torch.int32),
                    dL=mk((B, self.ngrad), torch.float64), dpi=mk((B, self.nu, self.ngrad), torch.float64),
                    res=mk((B, 4), torch.float64))

    def solve_sens_host(self, x0: np.ndarray, u0: Optional[np.ndarray] = None, max_sqp: int = 1, out: dict | None = None) -> dict:
        """Host buffers in/out through the C ABI (H2D + kernels + D2H + sync inside the call)."""
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        mode = _cabi.MODE_V if u0 is None else _cabi.MODE_Q
        u0a = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64).reshape(B, self.nu)
        if out is None:
            out = dict(u0=np.empty((B, self.nu)), cost=np.empty(B), status=np.empty(B, dtype=np.int32),
                       dL=np.empty((B, self.ngrad)), dpi=np.empty((B, self.nu, self.ngrad)), res=np.empty((B, 4)))
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        _cabi.check(self.lib.rlmpc_solve_sens_host(self._h, mode, int(max_sqp), B, p(x0), p(u0a), p(out["u0"]),
                                                   p(out["cost"]), p(out["status"]), p(out["dL"]), p(out["dpi"]),
                                                   p(out["res"])))
        return out

    def td_grad(self, td: torch.Tensor, dQ: torch.Tensor, status: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[sum_i td_i dQ_i/dtheta (ncols), sum_i td_i, n_valid] over samples with status 0."""
        B, ncols = td.shape[0], dQ.shape[1]
        acc = torch.empty(ncols + 2, dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.rlmpc_td_grad(self._h, B, ncols, _ptr(td.contiguous()), _ptr(dQ.contiguous()),
                                           _ptr(status), _ptr(acc), self._stream()))
        return acc

    PHASES = ("linearize", "qp_fast", "qp_full", "sens_stage", "sens_sweep", "sens_tail")
    # the chain-mass engine has one QP kernel (no fast path / queue split) and a separate parameter-contraction kernel
    CHAIN_PHASES = ("linearize", "qp", "unused", "sens_stage", "sens_sweep", "param_contraction")

    def timings(self) -> dict:
        """Device milliseconds of the phases of the last call (needs set_option("timing", 1))."""
        ms = np.zeros(8)
        _cabi.check(self.lib.rlmpc_get_timings(self._h, ms.ctypes.data_as(C.c_void_p), 8))
        names = self.CHAIN_PHASES if self.spec.model == _cabi.MODEL_CHAIN_MASS else self.PHASES
        d = dict(zip(names, ms[:6].tolist()))
        d.pop("unused", None)
        d["queue_len"], d["queue_ipm_iters"] = int(ms[6]), int(ms[7])
        return d

    def fp64_peak_tflops(self) -> float:
        """Measured FP64 FMA throughput of this device (rlmpc_fp64_peak)."""
        v = C.c_double()
        _cabi.check(self.lib.rlmpc_fp64_peak(self.device.index or 0, C.byref(v)))
        return float(v.value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.rlmpc_launch_count(self._h))


class IterateStore:
    """``capacity`` primal-dual iterates on the device, addressed by replay-buffer index.

    ``load(idx)`` puts the stored iterates of entries ``idx[b]`` into batch positions b of the engine (before
    an RTI step on a sampled minibatch), ``save(idx)`` writes the engine's current iterates back."""

    def __init__(self, engine: BatchedMPC, capacity: int):
        self.engine, self.capacity = engine, int(capacity)
        nbytes = int(engine.lib.rlmpc_store_bytes(engine._h, self.capacity))
        self.buf = torch.zeros(nbytes // 8, dtype=torch.float64, device=engine.device)

    def _copy(self, idx: torch.Tensor, to_store: int):
        i = idx.to(self.engine.device, torch.int32).contiguous()
        _cabi.check(self.engine.lib.rlmpc_store_copy(self.engine._h, i.numel(), _ptr(i), _ptr(self.buf), self.capacity,
                                                     to_store, self.engine._stream()))

    def save(self, idx: torch.Tensor) -> None:
        self._copy(idx, 1)

    def load_into(self, other: "BatchedMPC", idx: torch.Tensor) -> None:
        """Stored iterates -> batch positions of ANOTHER engine of the same problem (e.g. the actor's iterates as the
        critic's warm start)."""
        if other.spec.model != self.engine.spec.model or other.spec.N != self.engine.spec.N or other.nrows != self.engine.nrows:
            raise ValueError("load_into needs an engine of the same problem")
        i = idx.to(other.device, torch.int32).contiguous()
        _cabi.check(other.lib.rlmpc_store_copy(other._h, i.numel(), _ptr(i), _ptr(self.buf), self.capacity, 0, other._stream()))

    def load(self, idx: torch.Tensor) -> None:
        self._copy(idx, 0)
