"""Restated ``build_nlp`` / ``update_nlp`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows rlmpc/mpc/nlp.py of the reference:
  * variables  w = (u_0..u_{N-1}, x_0..x_N)                       nlp.py:1190-1201
  * g_k = f(x_k,u_k,p_model) - x_{k+1},  multipliers pi_k        nlp.py:822-831, 967
  * h <= 0 rows per stage in acados multiplier order               nlp.py:644-819,
    (lbu, lbx, ubu, ubx, lsbx, usbx; stage 0 carries lbu/ubu on     common/utils.py:4-25
    all nu and lbx/ubx on all nx)
  * L = cost + lam.h + pi.g                                        nlp.py:1180
  * R = [dL/dw ; g ; h+t ; lam*t - tau],  tau = 1e-8               nlp.py:1199,1214
  * z = (u, x, pi, lam, t)                                         nlp.py:1220
  * dz/dp = spsolve(csc(dR/dz), csc(-dR/dp)); dpi_dp = first nu rows  nlp.py:1413-1424
  * dV/dp = dQ/dp = dL/dp at the solution                          nlp.py:1211-1212,1401
Slack values enter h and the cost as constants (quirk Q4).
All derivatives come from torch.func (float64).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import scipy.sparse.linalg as splinalg
import torch
from scipy.sparse import csc_matrix
from torch.func import grad, jacfwd, jacrev

from .problems import F64, Problem, stage_cost_unscaled

TAU = 1e-8  # nlp.py:1199


def _T(a):
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, dtype=float), dtype=F64)


@dataclass
class Row:
    stage: int
    kind: str  # lbu lbx ubu ubx lsbx usbx
    idx: int  # component of u / x, or soft-row number for lsbx/usbx


def build_rows(pb: Problem) -> List[Row]:
    rows: List[Row] = []
    N = pb.N
    for k in range(N + 1):
        if k == 0:
            iu, ix, isx = np.arange(pb.nu), np.arange(pb.nx), np.zeros(0, dtype=int)
        elif k < N:
            iu, ix, isx = pb.idxbu, pb.idxbx, pb.idxsbx
        else:
            iu, ix, isx = np.zeros(0, dtype=int), pb.idxbx_e, pb.idxsbx_e
        ih = range(len(pb.lh)) if k < N else range(0)  # nh rows on stages 0..N-1 (nh_e = 0)
        rows += [Row(k, "lbu", int(i)) for i in iu]
        rows += [Row(k, "lbx", int(i)) for i in ix]
        rows += [Row(k, "lh", int(i)) for i in ih]
        rows += [Row(k, "ubu", int(i)) for i in iu]
        rows += [Row(k, "ubx", int(i)) for i in ix]
        rows += [Row(k, "uh", int(i)) for i in ih]
        rows += [Row(k, "lsbx", int(j)) for j in range(len(isx))]
        rows += [Row(k, "usbx", int(j)) for j in range(len(isx))]
    return rows


@dataclass
class Bounds:
    """Numeric bound/slack data of one sample (what nlp.vars.val holds for lbu_k, lbx_k, slbx_k ...)."""
    lbu: np.ndarray  # (N, nu)  stage 0 row = lbu_0 (clamped to u0 in Q-mode)
    ubu: np.ndarray
    lbx0: np.ndarray  # (nx,)
    ubx0: np.ndarray
    slbx: np.ndarray  # (N+1, nsbx)   slack values (constants, Q4)
    subx: np.ndarray


class RestatedNLP:
    def __init__(self, pb: Problem):
        self.pb = pb
        self.rows = build_rows(pb)
        self.nlam = len(self.rows)
        N, nx, nu = pb.N, pb.nx, pb.nu
        self.nw = N * nu + (N + 1) * nx
        self.npi = N * nx
        self.nz = self.nw + self.npi + 2 * self.nlam
        # per-stage slices of lam (acados `get(stage,"lam")` order)
        self.stage_rows: Dict[int, List[int]] = {k: [] for k in range(N + 1)}
        for i, r in enumerate(self.rows):
            self.stage_rows[r.stage].append(i)

    # ---- packing ----
    def split_w(self, w):
        pb = self.pb
        U = w[: pb.N * pb.nu].reshape(pb.N, pb.nu)
        X = w[pb.N * pb.nu:].reshape(pb.N + 1, pb.nx)
        return U, X

    def pack_w(self, U, X):
        return torch.cat([torch.as_tensor(U, dtype=F64).reshape(-1), torch.as_tensor(X, dtype=F64).reshape(-1)])

    # ---- functions of the reference's NLP ----
    def g(self, w, p):
        pb = self.pb
        U, X = self.split_w(w)
        pm = pb.p_model(p)
        Xn = torch.vmap(pb.f_disc, in_dims=(0, 0, None))(X[:-1], U, pm)
        return (Xn - X[1:]).reshape(-1)

    def cost(self, w, p, b: Bounds):
        pb = self.pb
        N = pb.N
        U, X = self.split_w(w)
        c = pb.stage_scale(0) * stage_cost_unscaled(pb, 0, X[0], U[0], p)
        if N > 1:
            sc = torch.as_tensor([pb.stage_scale(k) for k in range(1, N)], dtype=F64)
            lm = torch.vmap(lambda x, u: stage_cost_unscaled(pb, 1, x, u, p))(X[1:N], U[1:N])
            c = c + sc @ lm
        c = c + pb.stage_scale(N) * stage_cost_unscaled(pb, N, X[N], None, p)
        # linear slack penalties, EXTERNAL branch only (nlp.py:1099-1134)
        if pb.cost_type == "EXTERNAL":
            if len(pb.idxsbx) > 0:
                zl, zu = torch.as_tensor(pb.zl, dtype=F64), torch.as_tensor(pb.zu, dtype=F64)
                sc = torch.as_tensor([pb.dT * pb.gamma**k for k in range(1, N)], dtype=F64)
                ns = len(pb.idxsbx)
                c = c + sc @ (_T(b.slbx)[1:N, :ns] @ zl) + sc @ (_T(b.subx)[1:N, :ns] @ zu)
            if len(pb.idxsbx_e) > 0:
                ne = len(pb.idxsbx_e)
                c = c + _T(b.slbx)[N, :ne] @ torch.as_tensor(pb.zl_e, dtype=F64)
                c = c + _T(b.subx)[N, :ne] @ torch.as_tensor(pb.zu_e, dtype=F64)
        return c

    def _h_structure(self):
        """All rows the reference uses are affine: h = Jw w + Jsl slbx + Jsu subx + c(bounds)
        (nlp.py:644-798).  Built once; c depends on the per-sample bound values."""
        if hasattr(self, "_Jw"):
            return
        pb = self.pb
        N, nx, nu = pb.N, pb.nx, pb.nu
        ns = max(len(pb.idxsbx), len(pb.idxsbx_e))
        Jw = np.zeros((self.nlam, self.nw)); Jsl = np.zeros((self.nlam, (N + 1) * ns)); Jsu = np.zeros_like(Jsl)
        spec = []  # (kind, stage, position-in-bound-vector) to fill c
        for i, r in enumerate(self.rows):
            k = r.stage
            iu = k * nu + r.idx
            ix = N * nu + k * nx + r.idx
            if k == 0:
                idxs, soft = list(range(nx)), {}
            elif k < N:
                idxs, soft = pb.idxbx.tolist(), {int(j): n for n, j in enumerate(pb.idxsbx.tolist())}
            else:
                idxs, soft = pb.idxbx_e.tolist(), {int(j): n for n, j in enumerate(pb.idxsbx_e.tolist())}
            if r.kind == "lbu":
                Jw[i, iu] = -1.0; spec.append(("lbu", k, r.idx if k == 0 else pb.idxbu.tolist().index(r.idx)))
            elif r.kind == "ubu":
                Jw[i, iu] = 1.0; spec.append(("ubu", k, r.idx if k == 0 else pb.idxbu.tolist().index(r.idx)))
            elif r.kind == "lbx":
                Jw[i, ix] = -1.0
                j = idxs.index(r.idx)
                if j in soft:  # soft row j (position within idxbx) carries slack number soft[j] (Q5: Jsbx)
                    Jsl[i, k * ns + soft[j]] = -1.0
                spec.append(("lbx", k, j))
            elif r.kind == "ubx":
                Jw[i, ix] = 1.0
                j = idxs.index(r.idx)
                if j in soft:
                    Jsu[i, k * ns + soft[j]] = -1.0
                spec.append(("ubx", k, j))
            elif r.kind in ("lh", "uh"):
                # lh - h <= 0 / h - uh <= 0 with h = h0 + Ch [x_k; u_k]  (nlp.py:696-716, 745-762)
                sgn = -1.0 if r.kind == "lh" else 1.0
                Jw[i, N * nu + k * nx: N * nu + (k + 1) * nx] = sgn * pb.Ch[r.idx, :nx]
                Jw[i, k * nu: (k + 1) * nu] = sgn * pb.Ch[r.idx, nx:]
                spec.append((r.kind, k, r.idx))
            elif r.kind == "lsbx":
                Jsl[i, k * ns + r.idx] = -1.0; spec.append(("zero", k, 0))
            elif r.kind == "usbx":
                Jsu[i, k * ns + r.idx] = -1.0; spec.append(("zero", k, 0))
        self._Jw, self._Jsl, self._Jsu = (torch.as_tensor(a, dtype=F64) for a in (Jw, Jsl, Jsu))
        self._hspec = spec

    def _h_const(self, b: Bounds):
        pb, N = self.pb, self.pb.N
        c = np.zeros(self.nlam)
        for i, (kind, k, j) in enumerate(self._hspec):
            if kind == "lbu":
                c[i] = b.lbu[k][j]
            elif kind == "ubu":
                c[i] = -b.ubu[k][j]
            elif kind == "lbx":
                c[i] = b.lbx0[j] if k == 0 else (pb.lbx[j] if k < N else pb.lbx_e[j])
            elif kind == "ubx":
                c[i] = -(b.ubx0[j] if k == 0 else (pb.ubx[j] if k < N else pb.ubx_e[j]))
            elif kind == "lh":
                c[i] = pb.lh[j] - pb.h0[j]
            elif kind == "uh":
                c[i] = pb.h0[j] - pb.uh[j]
        return torch.as_tensor(c, dtype=F64)

    def h(self, w, p, b: Bounds):
        self._h_structure()
        out = self._Jw @ w + self._h_const(b)
        if self._Jsl.shape[1] > 0:
            out = out + self._Jsl @ _T(b.slbx).reshape(-1) + self._Jsu @ _T(b.subx).reshape(-1)
        return out

    def lagrangian(self, w, p, pi, lam, b: Bounds):
        return self.cost(w, p, b) + lam @ self.h(w, p, b) + pi @ self.g(w, p)

    def R(self, z, p, b: Bounds):
        nw, npi, nl = self.nw, self.npi, self.nlam
        w, pi, lam, t = z[:nw], z[nw:nw + npi], z[nw + npi:nw + npi + nl], z[nw + npi + nl:]
        dL_dw = grad(lambda w_: self.lagrangian(w_, p, pi, lam, b))(w)
        return torch.cat([dL_dw, self.g(w, p), self.h(w, p, b) + t, lam * t - TAU])

    # ---- restated update_nlp (nlp.py:1341-1424) ----
    def update(self, U, X, pi, lam, t, p, b: Bounds) -> dict:
        w = self.pack_w(U, X)
        T = lambda a: torch.as_tensor(np.asarray(a, dtype=float).reshape(-1), dtype=F64)
        p, pi, lam, t = T(p), T(pi), T(lam), T(t)
        z = torch.cat([w, pi, lam, t])
        out = {}
        out["cost"] = float(self.cost(w, p, b))
        out["g"] = self.g(w, p).numpy()
        out["h"] = self.h(w, p, b).numpy()
        out["L"] = float(self.lagrangian(w, p, pi, lam, b))
        out["dL_dw"] = grad(lambda w_: self.lagrangian(w_, p, pi, lam, b))(w).numpy()
        out["dL_dp"] = grad(lambda p_: self.lagrangian(w, p_, pi, lam, b))(p).numpy().reshape(1, -1)
        out["R"] = self.R(z, p, b).numpy()
        dR_dp = jacrev(lambda p_: self.R(z, p_, b))(p).numpy()
        dR_dz = jacfwd(lambda z_: self.R(z_, p, b))(z).numpy()
        out["dR_dp"], out["dR_dz"] = dR_dp, dR_dz
        dz_dp = splinalg.spsolve(csc_matrix(dR_dz), csc_matrix(-dR_dp))
        dz_dp = dz_dp.toarray() if hasattr(dz_dp, "toarray") else np.asarray(dz_dp).reshape(self.nz, -1)
        out["dz_dp"] = dz_dp
        out["dpi_dp"] = dz_dp[: self.pb.nu, :]
        return out

    def lam_of_stage(self, lam, k):
        return np.asarray(lam)[self.stage_rows[k]]
