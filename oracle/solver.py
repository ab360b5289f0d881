"""Dense SQP + primal-dual interior-point stand-in for ``AcadosOcpSolver.solve()``
(TEST INFRASTRUCTURE, see oracle/__init__.py).

acados (SQP, full steps) + HPIPM (Riccati IPM) are third-party and absent
(SURVEY.md 8(c)); this restates their published algorithm *without* exploiting any
stage structure: dense Jacobians from torch.func, dense KKT solves with numpy.  The
fixed point it converges to is the KKT point the reference's ``R(z,p)`` describes
(rlmpc/mpc/nlp.py:1214, tau = 1e-8 on the complementarity rows), with the equal
bounds of stage 0 (x_0 = s; u_0 = a in Q-mode, rlmpc/mpc/common/mpc.py:63-76)
imposed by fixing the variables -- the tau -> 0 limit of what HPIPM does with
lb == ub (quirk Q7).  Soft state bounds are genuine variables here, as in acados.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch.func import grad, hessian, jacrev

from .nlp import TAU, Bounds, RestatedNLP
from .problems import F64, Problem

T_EQ = 1e-10  # slack reported for the eliminated equal-bound rows (HPIPM returns something tiny)


@dataclass
class Solution:
    U: np.ndarray
    X: np.ndarray
    pi: np.ndarray  # (N*nx,)
    lam: np.ndarray  # all rows, acados order
    t: np.ndarray
    slbx: np.ndarray
    subx: np.ndarray
    cost: float
    status: int
    sqp_iter: int
    kkt: float
    bounds: Bounds


def _qp_ipm(H, c, G, g, J, hbar, free, tau, max_iter=200):
    """min 1/2 d'Hd + c'd  s.t. Gd+g=0, Jd+hbar<=0, d[~free]=0 -- primal-dual interior point, converging to the tau-central
    point lam*t = tau.  First Mehrotra predictor-corrector; on hard QPs (an RTI step far from the solution: many rows
    change sides) that heuristic can cycle, so a conservative long-step method (fixed centring, no corrector) takes over.
    Raises if neither reaches the central point: a fixture must never hold an unconverged QP."""
    n, m, me = len(c), len(hbar), len(g)
    fi = np.where(free)[0]
    nf = len(fi)
    Hf, cf, Gf, Jf = H[np.ix_(fi, fi)], c[fi], G[:, fi], J[:, fi]
    if m == 0:
        K = np.block([[Hf, Gf.T], [Gf, np.zeros((me, me))]])
        sol = np.linalg.solve(K, -np.concatenate([cf, g]))
        out = np.zeros(n); out[fi] = sol[:nf]
        return out, sol[nf:], np.zeros(0), np.zeros(0), 0

    def run(mehrotra, iters):
        d = np.zeros(nf)
        t = np.maximum(-hbar, 1.0)
        lam = np.ones(m)
        pi = np.zeros(me)
        a_prev = 1.0
        settled = 0  # consecutive full Newton steps that ended on the central path
        for it in range(iters):
            mu = float(lam @ t) / m
            C = lam / t
            Kmat = np.block([[Hf + Jf.T @ (C[:, None] * Jf), Gf.T], [Gf, np.zeros((me, me))]])

            def solve(target):  # target = the vector "sigma*mu - corr" of the complementarity rows
                rhs = -np.concatenate([cf + Jf.T @ (target / t + C * (t + hbar)), g])
                sol = np.linalg.solve(Kmat, rhs)
                dh, pih = sol[:nf], sol[nf:]
                th = -(Jf @ dh + hbar)
                lamh = target / t + C * (t - th)
                return dh, pih, th, lamh

            def steplen(dt, dl):
                a = 1.0
                neg = dt < 0
                if neg.any():
                    a = min(a, float(np.min(-t[neg] / dt[neg])))
                neg = dl < 0
                if neg.any():
                    a = min(a, float(np.min(-lam[neg] / dl[neg])))
                return a

            if mehrotra:
                dh, pih, th, lamh = solve(np.full(m, min(tau, mu)))  # predictor (affine scaling towards tau)
                dta, dla = th - t, lamh - lam
                a_aff = steplen(dta, dla)
                mu_aff = float((lam + a_aff * dla) @ (t + a_aff * dta)) / m
                sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
                target = np.maximum(sigma * mu, tau) - dta * dla
                frac = 0.995
            else:
                sigma = 0.1 if a_prev > 0.9 else 0.5
                target = np.full(m, max(sigma * mu, tau))
                frac = 0.9
            dh, pih, th, lamh = solve(target)
            dt_, dl_ = th - t, lamh - lam
            amax = steplen(dt_, dl_)
            a = min(1.0, frac * amax) if amax < 1.0 else 1.0
            step = np.linalg.norm(a * (dh - d), np.inf)
            d = d + a * (dh - d)
            pi = pi + a * (pih - pi)
            t = t + a * dt_
            lam = lam + a * dl_
            a_prev = a
            comp = np.max(np.abs(lam * t - tau))
            settled = settled + 1 if (a == 1.0 and comp < 1e-3 * tau) else 0
            # a fixed point of the Newton iteration: the step has vanished, or (rounding of lam/t ~ 1e16 keeps it from
            # vanishing exactly) three full steps in a row stayed on the central path
            if settled >= 1 and (step < 1e-13 * (1 + np.linalg.norm(d, np.inf)) or settled >= 3):
                return d, pi, lam, t, it + 1, True
        return d, pi, lam, t, iters, False

    d, pi, lam, t, it, ok = run(True, min(max_iter, 60))
    if not ok:
        d, pi, lam, t, it2, ok = run(False, 400)
        it += it2
    if not ok:
        raise RuntimeError("oracle QP: the interior-point iteration did not reach the tau-central point")
    out = np.zeros(n); out[fi] = d
    return out, pi, lam, t, it


class DenseSolver:
    def __init__(self, pb: Problem):
        self.pb = pb
        self.nlp = RestatedNLP(pb)
        pb_ns = max(len(pb.idxsbx), len(pb.idxsbx_e))
        self.ns = pb_ns
        N, nx, nu = pb.N, pb.nx, pb.nu
        self.nw = self.nlp.nw
        self.nv = self.nw + 2 * (N + 1) * pb_ns

    # ---- helpers on v = [w ; slbx ; subx] ----
    def _bounds(self, v, x0, u0):
        pb, ns, N = self.pb, self.ns, self.pb.N
        lbu = np.tile(pb.lbu, (N, 1)).astype(float)
        ubu = np.tile(pb.ubu, (N, 1)).astype(float)
        if u0 is not None:
            lbu[0] = u0; ubu[0] = u0
        sl = v[self.nw:self.nw + (N + 1) * ns].reshape(N + 1, ns)
        su = v[self.nw + (N + 1) * ns:].reshape(N + 1, ns)
        return Bounds(lbu=lbu, ubu=ubu, lbx0=np.asarray(x0, float), ubx0=np.asarray(x0, float), slbx=sl, subx=su)

    def _masks(self, qmode):
        pb, nlp, N, ns = self.pb, self.nlp, self.pb.N, self.ns
        free = np.ones(self.nv, dtype=bool)
        free[N * pb.nu: N * pb.nu + pb.nx] = False  # x_0
        if qmode:
            free[: pb.nu] = False  # u_0
        sl_free = np.zeros((N + 1, ns), dtype=bool)
        if len(pb.idxsbx):
            sl_free[1:N, : len(pb.idxsbx)] = True
        if len(pb.idxsbx_e):
            sl_free[N, : len(pb.idxsbx_e)] = True
        free[self.nw:] = np.concatenate([sl_free.ravel(), sl_free.ravel()])
        genuine = np.array([not (r.stage == 0 and (r.kind in ("lbx", "ubx") or (qmode and r.kind in ("lbu", "ubu"))))
                            for r in nlp.rows])
        return free, genuine

    def solve(self, x0, u0=None, p=None, init=None, tol=1e-9, max_iter=100, hessian_approx=None,
              polish=True, verbose=False) -> Solution:
        pb, nlp = self.pb, self.nlp
        N, nx, nu, ns = pb.N, pb.nx, pb.nu, self.ns
        p_t = torch.as_tensor(pb.p_nominal if p is None else np.asarray(p, float), dtype=F64)
        hess_mode = hessian_approx or pb.hessian_approx
        qmode = u0 is not None
        free, genuine = self._masks(qmode)
        gi = np.where(genuine)[0]
        # initial guess like MPC.reset (mpc.py:204-210): all stages = x0, u = 0
        if init is None and pb.x_init is not None:
            U = np.tile(np.asarray(pb.u_init, float), (N, 1)); X = np.tile(np.asarray(pb.x_init, float), (N + 1, 1))
        elif init is None:
            U = np.zeros((N, nu)); X = np.tile(np.asarray(x0, float), (N + 1, 1))
        else:
            U, X = np.array(init[0], float), np.array(init[1], float)
        X[0] = x0
        if qmode:
            U[0] = u0
        v = torch.cat([nlp.pack_w(U, X), torch.zeros(2 * (N + 1) * ns, dtype=F64)])
        x0n = np.asarray(x0, float)

        def cost_v(v_):
            return nlp.cost(v_[: self.nw], p_t, self._bounds(v_, x0n, u0))

        def h_v(v_):
            return nlp.h(v_[: self.nw], p_t, self._bounds(v_, x0n, u0))[gi]

        def g_v(v_):
            return nlp.g(v_[: self.nw], p_t)

        def lag_v(v_, pi_, lam_):
            return cost_v(v_) + pi_ @ g_v(v_) + lam_ @ h_v(v_)

        m = len(gi)
        pi = np.zeros(N * nx); lam = np.full(m, 0.0); t = np.full(m, 1.0)
        if init is not None and len(init) > 2:  # stored multipliers (pi, lam over ALL rows): the exact Hessian of the first step
            pi = np.array(init[2], float).ravel().copy(); lam = np.array(init[3], float)[gi].copy()
        status, kkt, it = 2, np.inf, 0
        for it in range(max_iter + 1):
            c = grad(cost_v)(v).numpy()
            gv = g_v(v).numpy()
            G = jacrev(g_v)(v).numpy()
            hv = h_v(v).numpy() if m else np.zeros(0)
            J = jacrev(h_v)(v).numpy() if m else np.zeros((0, self.nv))
            r_stat = (c + G.T @ pi + (J.T @ lam if m else 0.0))[free]
            kkt = max(np.max(np.abs(r_stat)), np.max(np.abs(gv)),
                      np.max(np.abs(hv + t)) if m and it > 0 else 0.0,
                      np.max(np.abs(lam * t - TAU)) if m and it > 0 else 0.0)
            if verbose:
                print(f"  oracle sqp it {it}: kkt={kkt:.3e}")
            if it > 0 and kkt < tol:
                status = 0
                break
            if it == max_iter:
                break
            if hess_mode == "GAUSS_NEWTON":
                H = hessian(cost_v)(v).numpy()  # y linear in (x,u): GN == cost Hessian
            else:
                H = hessian(lambda v_: lag_v(v_, torch.as_tensor(pi, dtype=F64), torch.as_tensor(lam, dtype=F64)))(v).numpy()
            d, pi, lam, t, _ = _qp_ipm(H, c, G, gv, J, hv, free, TAU)
            v = v + torch.as_tensor(d, dtype=F64)

        if polish and status == 0:
            v, pi, lam, t, kkt = self._polish(v, pi, lam, t, lag_v, g_v, h_v, free, m)

        # assemble the reference's full multiplier vectors (all rows)
        vb = self._bounds(v, x0n, u0)
        U_, X_ = nlp.split_w(v[: self.nw])
        lam_full = np.zeros(nlp.nlam); t_full = np.zeros(nlp.nlam)
        lam_full[gi] = lam; t_full[gi] = t
        rho = grad(lambda v_: lag_v(v_, torch.as_tensor(pi, dtype=F64), torch.as_tensor(lam, dtype=F64)))(v).numpy()
        for i, r in enumerate(nlp.rows):
            if genuine[i]:
                continue
            rv = rho[N * nu + r.idx] if r.kind in ("lbx", "ubx") else rho[r.idx]
            lower = r.kind in ("lbx", "lbu")
            lam_full[i] = max(rv, 0.0) if lower else max(-rv, 0.0)
            lam_full[i] += TAU / T_EQ
            t_full[i] = T_EQ
        sl = np.array(vb.slbx.detach().numpy() if isinstance(vb.slbx, torch.Tensor) else vb.slbx)
        su = np.array(vb.subx.detach().numpy() if isinstance(vb.subx, torch.Tensor) else vb.subx)
        bnum = Bounds(lbu=vb.lbu, ubu=vb.ubu, lbx0=vb.lbx0, ubx0=vb.ubx0, slbx=sl, subx=su)
        cost = float(cost_v(v))
        return Solution(U=U_.detach().numpy().copy(), X=X_.detach().numpy().copy(), pi=np.array(pi), lam=lam_full,
                        t=t_full, slbx=sl, subx=su, cost=cost, status=status, sqp_iter=it, kkt=float(kkt), bounds=bnum)

    def _polish(self, v, pi, lam, t, lag_v, g_v, h_v, free, m):
        """Newton on the tau-perturbed KKT system with the exact Lagrangian Hessian."""
        fi = np.where(free)[0]
        nf, me = len(fi), len(pi)
        kkt = np.inf
        for _ in range(12):
            pit, lamt = torch.as_tensor(pi, dtype=F64), torch.as_tensor(lam, dtype=F64)
            gradL = grad(lambda v_: lag_v(v_, pit, lamt))(v).numpy()[fi]
            Hx = hessian(lambda v_: lag_v(v_, pit, lamt))(v).numpy()[np.ix_(fi, fi)]
            G = jacrev(g_v)(v).numpy()[:, fi]
            gv = g_v(v).numpy()
            if m:
                J = jacrev(h_v)(v).numpy()[:, fi]
                hv = h_v(v).numpy()
            else:
                J = np.zeros((0, nf)); hv = np.zeros(0)
            res = np.concatenate([gradL, gv, hv + t, lam * t - TAU])
            kkt = np.max(np.abs(res))
            if kkt < 1e-12:
                break
            Z = np.zeros
            Kj = np.block([
                [Hx, G.T, J.T, Z((nf, m))],
                [G, Z((me, me)), Z((me, m)), Z((me, m))],
                [J, Z((m, me)), Z((m, m)), np.eye(m)],
                [Z((m, nf)), Z((m, me)), np.diag(t), np.diag(lam)],
            ])
            dz = np.linalg.solve(Kj, -res)
            dv, dpi, dlam, dt = dz[:nf], dz[nf:nf + me], dz[nf + me:nf + me + m], dz[nf + me + m:]
            a = 1.0
            for cur, dd in ((t, dt), (lam, dlam)):
                neg = dd < 0
                if neg.any():
                    a = min(a, 0.995 * float(np.min(-cur[neg] / dd[neg])))
            step = np.zeros(len(v)); step[fi] = a * dv
            v = v + torch.as_tensor(step, dtype=F64)
            pi = pi + a * dpi; lam = lam + a * dlam; t = t + a * dt
        return v, pi, lam, t, kkt

    # ---- the whole reference unit: solve + update_nlp ----
    def unit(self, x0, u0=None, p=None, **kw):
        sol = self.solve(x0, u0=u0, p=p, **kw)
        upd = self.nlp.update(sol.U, sol.X, sol.pi, sol.lam, sol.t,
                              self.pb.p_nominal if p is None else p, sol.bounds)
        return sol, upd
