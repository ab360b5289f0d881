// Chain of masses (rlmpc/mpc/chain_mass/ocp_utils.py:59-147): the continuous dynamics, their first and second
// derivatives per LINK, and the explicit RK4 integrator with two sub-steps (ocp_utils.py:42-56).
//
//   n_mass masses on a line of springs; mass 0 is fixed at the wall (eliminated), masses 1..M are free
//   (M = n_mass - 2), the last mass is moved by the input (its velocity is u).
//   x = [pos_1 .. pos_{M+1} (3 each) ; vel_1 .. vel_M (3 each)],  nx = 3 (2M + 1),  nu = 3
//   xdot = [vel ; u ; acc],   acc_m = [0 0 -9.81] - T_m + T_{m+1} + w_m,
//   link i = 0..M between mass i and mass i+1 carries the force (ocp_utils.py:80-124)
//       T_i[j] = D_ij / m_i (1 - L_ij / |dist_i|) dist_i[j]  +  C_ij vl_i[j]
//       dist_i = pos_i - pos_{i-1} (pos_{-1} = wall = 0),  vl_0 = vel_0, vl_M = u - vel_{M-1}, else vel_i - vel_{i-1}
//   (the damping term is NOT divided by the mass, as in the reference).
//   theta = [m (n_link) | D (3 n_link) | L (3 n_link) | C (3 n_link) | Q (nx^2, column-major) | R (9) | w (3 M)]
//   (define_param_struct_symSX, ocp_utils.py:353-371); the "dynamic" parameters are m, D, L, C, w.
//
// Everything here is scalar host/device code on small arrays: the kernels decide who calls it (one lane per
// tangent direction, one thread per (sample, stage, right-hand side), ...).
#pragma once
#include "../common.cuh"

namespace rlmpc {

template <int NMASS>
struct ChainModel {
  static constexpr int NM = NMASS, M = NMASS - 2, NL = NMASS - 1;
  static constexpr int NPOS = 3 * (M + 1), NVEL = 3 * M;
  static constexpr int NX = NPOS + NVEL, NU = 3, NW = NX + NU, NC = NW + 1;
  static constexpr int TH_M = 0, TH_D = NL, TH_L = 4 * NL, TH_C = 7 * NL, TH_Q = 10 * NL, TH_R = TH_Q + NX * NX;
  static constexpr int TH_W = TH_R + NU * NU, NTH = TH_W + 3 * M;
  // dynamic parameters in compact numbering: [m | D | L | C | w]
  static constexpr int PD_M = 0, PD_D = NL, PD_L = 4 * NL, PD_C = 7 * NL, PD_W = 10 * NL, NPD = 10 * NL + 3 * M;
  MPC_HD static constexpr int pd_to_theta(int p) { return p < PD_W ? p : TH_W + (p - PD_W); }
  static constexpr int NSUB = 2;  // RK4 sub-steps per stage (export_discrete_erk4_integrator_step, n_stages = 2)
  static constexpr int NSP = 4 * NSUB;  // stage points of one shooting interval

  // geometry of link i at state x (u only for the last link's velocity)
  struct Link {
    double d[3], vl[3], r, ir, im;  // dist, link velocity, |dist|, 1/|dist|, 1/mass of the link (one division per link
                                    // evaluation instead of one per derivative routine)
  };
  MPC_HD static double inv_sqrt(double v) {
#ifdef __CUDA_ARCH__
    return rsqrt(v);  // one MUFU.RSQ64H + refinement instead of sqrt followed by a division
#else
    return 1.0 / sqrt(v);
#endif
  }
  MPC_HD static void link_at(const double* x, const double* u, const double* th, int i, Link& k) {
    MPC_UNROLL for (int j = 0; j < 3; ++j) {
      k.d[j] = x[3 * i + j] - (i > 0 ? x[3 * (i - 1) + j] : 0.0);
      if (i == 0) k.vl[j] = x[NPOS + j];
      else if (i == M) k.vl[j] = u[j] - x[NPOS + 3 * (M - 1) + j];
      else k.vl[j] = x[NPOS + 3 * i + j] - x[NPOS + 3 * (i - 1) + j];
    }
    const double dd = k.d[0] * k.d[0] + k.d[1] * k.d[1] + k.d[2] * k.d[2];
    k.ir = inv_sqrt(dd);
    k.r = dd * k.ir;
    k.im = 1.0 / th[TH_M + i];
  }
  // total link force T_i
  MPC_HD static void link_force(const double* th, int i, const Link& k, double* T) {
    const double im = k.im;
    MPC_UNROLL for (int j = 0; j < 3; ++j)
      T[j] = th[TH_D + 3 * i + j] * im * (1.0 - th[TH_L + 3 * i + j] * k.ir) * k.d[j] + th[TH_C + 3 * i + j] * k.vl[j];
  }
  // S = dT/d(dist) (3 x 3, row-major): S[j][l] = c_j ((1 - L_j / r) delta_jl + L_j d_j d_l / r^3)
  MPC_HD static void link_S(const double* th, int i, const Link& k, double* S) {
    const double im = k.im, ir3 = k.ir * k.ir * k.ir;
    MPC_UNROLL for (int j = 0; j < 3; ++j) {
      const double c = th[TH_D + 3 * i + j] * im, Lj = th[TH_L + 3 * i + j];
      MPC_UNROLL for (int l = 0; l < 3; ++l) S[3 * j + l] = c * ((j == l ? 1.0 - Lj * k.ir : 0.0) + Lj * k.d[j] * k.d[l] * ir3);
    }
  }
  // G = Hessian of nu'T wrt dist, packed [00 01 02 11 12 22]:  a_j = nu_j c_j L_j, s = a'd,
  //   G = (a d' + d a' + s I) / r^3 - 3 s d d' / r^5
  MPC_HD static void link_G(const double* th, int i, const Link& k, const double* nu, double* G) {
    const double im = k.im, ir2 = k.ir * k.ir, ir3 = ir2 * k.ir;
    double a[3], s = 0.0;
    MPC_UNROLL for (int j = 0; j < 3; ++j) {
      a[j] = nu[j] * th[TH_D + 3 * i + j] * im * th[TH_L + 3 * i + j];
      s += a[j] * k.d[j];
    }
    const double q = 3.0 * s * ir2;
    int e = 0;
    MPC_UNROLL for (int l = 0; l < 3; ++l)
      MPC_UNROLL for (int m = l; m < 3; ++m)
        G[e++] = (a[l] * k.d[m] + a[m] * k.d[l] + (l == m ? s : 0.0) - q * k.d[l] * k.d[m]) * ir3;
  }
  MPC_HD static void sym3_mul(const double* G, const double* v, double* out) {
    out[0] = G[0] * v[0] + G[1] * v[1] + G[2] * v[2];
    out[1] = G[1] * v[0] + G[3] * v[1] + G[4] * v[2];
    out[2] = G[2] * v[0] + G[4] * v[1] + G[5] * v[2];
  }
  // adjoint weight of link i's force for an adjoint mu on xdot:  nu_i = -mu_acc[i] (i < M) + mu_acc[i-1] (i > 0)
  MPC_HD static void link_nu(const double* mu, int i, double* nu) {
    MPC_UNROLL for (int j = 0; j < 3; ++j)
      nu[j] = (i < M ? -mu[NPOS + 3 * i + j] : 0.0) + (i > 0 ? mu[NPOS + 3 * (i - 1) + j] : 0.0);
  }
  // gradient of nu'T_i wrt the link's own parameters, added to g (compact numbering).  With tangents (dd, dvl of the
  // geometry, dnu of the weight; all may be null = zero) it adds the directional derivative of that gradient instead.
  MPC_HD static void link_theta_grad(const double* th, int i, const Link& k, const double* nu, double* g) {
    const double im = k.im;
    double gm = 0.0;
    MPC_UNROLL for (int j = 0; j < 3; ++j) {
      const double Dj = th[TH_D + 3 * i + j], Lj = th[TH_L + 3 * i + j];
      const double e = (1.0 - Lj * k.ir) * k.d[j] * im;  // spring force per unit D
      g[PD_D + 3 * i + j] += nu[j] * e;
      gm -= nu[j] * Dj * e * im;
      g[PD_L + 3 * i + j] -= nu[j] * Dj * im * k.d[j] * k.ir;
      g[PD_C + 3 * i + j] += nu[j] * k.vl[j];
    }
    g[PD_M + i] += gm;
  }
  MPC_HD static void link_theta_grad_tan(const double* th, int i, const Link& k, const double* nu, const double* dnu,
                                         const double* dd, const double* dvl, double* g) {
    const double im = k.im, ir3 = k.ir * k.ir * k.ir;
    const double ddot = k.d[0] * dd[0] + k.d[1] * dd[1] + k.d[2] * dd[2];
    double gm = 0.0;
    MPC_UNROLL for (int j = 0; j < 3; ++j) {
      const double Dj = th[TH_D + 3 * i + j], Lj = th[TH_L + 3 * i + j];
      const double e = (1.0 - Lj * k.ir) * k.d[j] * im;
      const double de = ((1.0 - Lj * k.ir) * dd[j] + Lj * k.d[j] * ddot * ir3) * im;
      const double t1 = dnu[j] * e + nu[j] * de;
      g[PD_D + 3 * i + j] += t1;
      gm -= Dj * t1 * im;
      g[PD_L + 3 * i + j] -= Dj * im * (dnu[j] * k.d[j] * k.ir + nu[j] * (dd[j] * k.ir - k.d[j] * ddot * ir3));
      g[PD_C + 3 * i + j] += dnu[j] * k.vl[j] + nu[j] * dvl[j];
    }
    g[PD_M + i] += gm;
  }

  // ---------------- whole-vector versions (scalar code, one caller does everything) ----------------
  MPC_HD static void ode(const double* x, const double* u, const double* th, double* f) {
    MPC_UNROLL for (int c = 0; c < NVEL; ++c) f[c] = x[NPOS + c];
    MPC_UNROLL for (int j = 0; j < 3; ++j) f[NVEL + j] = u[j];
    MPC_UNROLL for (int m = 0; m < M; ++m)
      MPC_UNROLL for (int j = 0; j < 3; ++j) f[NPOS + 3 * m + j] = (j == 2 ? -9.81 : 0.0) + th[TH_W + 3 * m + j];
    for (int i = 0; i <= M; ++i) {
      Link k;
      double T[3];
      link_at(x, u, th, i, k);
      link_force(th, i, k, T);
      MPC_UNROLL for (int j = 0; j < 3; ++j) {
        if (i < M) f[NPOS + 3 * i + j] -= T[j];
        if (i > 0) f[NPOS + 3 * (i - 1) + j] += T[j];
      }
    }
  }
  // tangent of the link geometry for a direction (dx, du)
  MPC_HD static void link_tan(const double* dx, const double* du, int i, double* dd, double* dvl) {
    MPC_UNROLL for (int j = 0; j < 3; ++j) {
      dd[j] = dx[3 * i + j] - (i > 0 ? dx[3 * (i - 1) + j] : 0.0);
      if (i == 0) dvl[j] = dx[NPOS + j];
      else if (i == M) dvl[j] = du[j] - dx[NPOS + 3 * (M - 1) + j];
      else dvl[j] = dx[NPOS + 3 * i + j] - dx[NPOS + 3 * (i - 1) + j];
    }
  }
  MPC_HD static void ode_jvp(const double* x, const double* u, const double* th, const double* dx, const double* du, double* df) {
    MPC_UNROLL for (int c = 0; c < NVEL; ++c) df[c] = dx[NPOS + c];
    MPC_UNROLL for (int j = 0; j < 3; ++j) df[NVEL + j] = du[j];
    MPC_UNROLL for (int c = 0; c < NVEL; ++c) df[NPOS + c] = 0.0;
    for (int i = 0; i <= M; ++i) {
      Link k;
      double S[9], dd[3], dvl[3];
      link_at(x, u, th, i, k);
      link_S(th, i, k, S);
      link_tan(dx, du, i, dd, dvl);
      MPC_UNROLL for (int j = 0; j < 3; ++j) {
        const double dT = S[3 * j] * dd[0] + S[3 * j + 1] * dd[1] + S[3 * j + 2] * dd[2] + th[TH_C + 3 * i + j] * dvl[j];
        if (i < M) df[NPOS + 3 * i + j] -= dT;
        if (i > 0) df[NPOS + 3 * (i - 1) + j] += dT;
      }
    }
  }
  // scatter of link i's adjoints (dbar on dist, vbar on the link velocity) into xbar, ubar
  MPC_HD static void link_scatter(int i, const double* dbar, const double* vbar, double* xbar, double* ubar) {
    MPC_UNROLL for (int j = 0; j < 3; ++j) {
      xbar[3 * i + j] += dbar[j];
      if (i > 0) xbar[3 * (i - 1) + j] -= dbar[j];
      if (i == 0) xbar[NPOS + j] += vbar[j];
      else if (i == M) { ubar[j] += vbar[j]; xbar[NPOS + 3 * (M - 1) + j] -= vbar[j]; }
      else { xbar[NPOS + 3 * i + j] += vbar[j]; xbar[NPOS + 3 * (i - 1) + j] -= vbar[j]; }
    }
  }
  // xbar += J_x' mu, ubar += J_u' mu, gth += d(mu'f)/d theta_dyn (gth may be null)
  MPC_HD static void ode_vjp(const double* x, const double* u, const double* th, const double* mu, double* xbar, double* ubar,
                             double* gth) {
    MPC_UNROLL for (int c = 0; c < NVEL; ++c) xbar[NPOS + c] += mu[c];
    MPC_UNROLL for (int j = 0; j < 3; ++j) ubar[j] += mu[NVEL + j];
    if (gth) MPC_UNROLL for (int c = 0; c < NVEL; ++c) gth[PD_W + c] += mu[NPOS + c];
    for (int i = 0; i <= M; ++i) {
      Link k;
      double S[9], nu[3], dbar[3], vbar[3];
      link_at(x, u, th, i, k);
      link_S(th, i, k, S);
      link_nu(mu, i, nu);
      MPC_UNROLL for (int l = 0; l < 3; ++l) {
        dbar[l] = nu[0] * S[l] + nu[1] * S[3 + l] + nu[2] * S[6 + l];
        vbar[l] = nu[l] * th[TH_C + 3 * i + l];
      }
      link_scatter(i, dbar, vbar, xbar, ubar);
      if (gth) link_theta_grad(th, i, k, nu, gth);
    }
  }
  // directional derivative of ode_vjp along (dx, du) with adjoint tangent dmu:
  //   dxbar += J_x' dmu + [Hessian of mu'f] (dx, du),  likewise dubar, dgth
  MPC_HD static void ode_vjp_tan(const double* x, const double* u, const double* th, const double* dx, const double* du,
                                 const double* mu, const double* dmu, double* dxbar, double* dubar, double* dgth) {
    MPC_UNROLL for (int c = 0; c < NVEL; ++c) dxbar[NPOS + c] += dmu[c];
    MPC_UNROLL for (int j = 0; j < 3; ++j) dubar[j] += dmu[NVEL + j];
    if (dgth) MPC_UNROLL for (int c = 0; c < NVEL; ++c) dgth[PD_W + c] += dmu[NPOS + c];
    for (int i = 0; i <= M; ++i) {
      Link k;
      double S[9], G[6], nu[3], dnu[3], dd[3], dvl[3], gd[3], dbar[3], vbar[3];
      link_at(x, u, th, i, k);
      link_S(th, i, k, S);
      link_nu(mu, i, nu);
      link_nu(dmu, i, dnu);
      link_tan(dx, du, i, dd, dvl);
      link_G(th, i, k, nu, G);
      sym3_mul(G, dd, gd);
      MPC_UNROLL for (int l = 0; l < 3; ++l) {
        dbar[l] = dnu[0] * S[l] + dnu[1] * S[3 + l] + dnu[2] * S[6 + l] + gd[l];
        vbar[l] = dnu[l] * th[TH_C + 3 * i + l];
      }
      link_scatter(i, dbar, vbar, dxbar, dubar);
      if (dgth) link_theta_grad_tan(th, i, k, nu, dnu, dd, dvl, dgth);
    }
  }

  // RK4 tableau as used below: stage point st = x + RA[st] * k_{st-1}, x+ = x + sum RB[st] k_st
  MPC_HD static double rk_a(int st, double h) { return st == 0 ? 0.0 : (st == 3 ? h : 0.5 * h); }
  MPC_HD static double rk_b(int st, double h) { return (st == 0 || st == 3) ? h / 6.0 : h / 3.0; }

  // x+ = F(x, u; theta): NSUB RK4 steps of h each.  xs (optional): the NSP stage points, [NSP][NX]
  MPC_HD static void step(const double* x, const double* u, const double* th, double h, double* xn, double* xs) {
    double xc[NX], k[NX], acc[NX], p[NX];
    MPC_UNROLL for (int c = 0; c < NX; ++c) xc[c] = x[c];
    for (int sub = 0; sub < NSUB; ++sub) {
      MPC_UNROLL for (int c = 0; c < NX; ++c) { acc[c] = 0.0; k[c] = 0.0; }
      for (int st = 0; st < 4; ++st) {
        const double a = rk_a(st, h), b = rk_b(st, h);
        MPC_UNROLL for (int c = 0; c < NX; ++c) p[c] = xc[c] + a * k[c];
        if (xs) MPC_UNROLL for (int c = 0; c < NX; ++c) xs[(sub * 4 + st) * NX + c] = p[c];
        ode(p, u, th, k);
        MPC_UNROLL for (int c = 0; c < NX; ++c) acc[c] += b * k[c];
      }
      MPC_UNROLL for (int c = 0; c < NX; ++c) xc[c] += acc[c];
    }
    MPC_UNROLL for (int c = 0; c < NX; ++c) xn[c] = xc[c];
  }

  // Forward-over-reverse through one shooting interval, for ONE direction and ONE adjoint seed (scalar code):
  //   dirx, diru : direction in (x, u)
  //   seed       : adjoint on x+ for the second-order part   (pi_k)
  //   seed1      : first-order adjoint seed, or null         (y_pi)
  // returns  gth[p] = d/dtheta_p ( seed1' F )  +  d/deps d/dtheta_p ( seed' F )(w + eps dir)      (compact numbering)
  // i.e. the parameter gradient of  seed1'F + seed' (dF/dw) dir  -- the two contractions the policy gradient needs
  // per stage and adjoint right-hand side, at the cost of a few integrator sweeps instead of dense dF/dtheta and
  // d2(pi'F)/dw dtheta blocks.
  MPC_HD static void param_contraction(const double* x, const double* u, const double* th, double h, const double* dirx,
                                       const double* diru, const double* seed, const double* seed1, double* gth) {
    double xs[NSP * NX], dxs[NSP * NX];
    {  // nominal and tangent forward sweeps, stage points kept
      double xc[NX], dxc[NX], k[NX], dk[NX], acc[NX], dacc[NX];
      MPC_UNROLL for (int c = 0; c < NX; ++c) { xc[c] = x[c]; dxc[c] = dirx[c]; }
      for (int sub = 0; sub < NSUB; ++sub) {
        MPC_UNROLL for (int c = 0; c < NX; ++c) { acc[c] = 0.0; dacc[c] = 0.0; k[c] = 0.0; dk[c] = 0.0; }
        for (int st = 0; st < 4; ++st) {
          const double a = rk_a(st, h), b = rk_b(st, h);
          double* p = xs + (sub * 4 + st) * NX;
          double* dp = dxs + (sub * 4 + st) * NX;
          MPC_UNROLL for (int c = 0; c < NX; ++c) { p[c] = xc[c] + a * k[c]; dp[c] = dxc[c] + a * dk[c]; }
          ode(p, u, th, k);
          ode_jvp(p, u, th, dp, diru, dk);
          MPC_UNROLL for (int c = 0; c < NX; ++c) { acc[c] += b * k[c]; dacc[c] += b * dk[c]; }
        }
        MPC_UNROLL for (int c = 0; c < NX; ++c) { xc[c] += acc[c]; dxc[c] += dacc[c]; }
      }
    }
    MPC_UNROLL for (int p = 0; p < NPD; ++p) gth[p] = 0.0;
    // reverse sweeps: lam = adjoint of the sub-step's result, dlam its tangent (seeded with seed1: the first-order
    // part rides along because the tangent recursion is linear in dlam)
    double lam[NX], dlam[NX];
    MPC_UNROLL for (int c = 0; c < NX; ++c) { lam[c] = seed[c]; dlam[c] = seed1 ? seed1[c] : 0.0; }
    for (int sub = NSUB - 1; sub >= 0; --sub) {
      double xb[NX], dxb[NX], lacc[NX], dlacc[NX], ub[3], dub[3];
      MPC_UNROLL for (int c = 0; c < NX; ++c) { xb[c] = 0.0; dxb[c] = 0.0; lacc[c] = 0.0; dlacc[c] = 0.0; }
      MPC_UNROLL for (int j = 0; j < 3; ++j) { ub[j] = 0.0; dub[j] = 0.0; }
      for (int st = 3; st >= 0; --st) {
        const double b = rk_b(st, h), an = st < 3 ? rk_a(st + 1, h) : 0.0;
        double kb[NX], dkb[NX];
        MPC_UNROLL for (int c = 0; c < NX; ++c) { kb[c] = b * lam[c] + an * xb[c]; dkb[c] = b * dlam[c] + an * dxb[c]; }
        MPC_UNROLL for (int c = 0; c < NX; ++c) { xb[c] = 0.0; dxb[c] = 0.0; }
        const double* p = xs + (sub * 4 + st) * NX;
        const double* dp = dxs + (sub * 4 + st) * NX;
        ode_vjp(p, u, th, kb, xb, ub, nullptr);
        ode_vjp_tan(p, u, th, dp, diru, kb, dkb, dxb, dub, gth);
        MPC_UNROLL for (int c = 0; c < NX; ++c) { lacc[c] += xb[c]; dlacc[c] += dxb[c]; }
      }
      MPC_UNROLL for (int c = 0; c < NX; ++c) { lam[c] += lacc[c]; dlam[c] += dlacc[c]; }
    }
  }
};

}  // namespace rlmpc
