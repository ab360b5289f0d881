// Cart-pole swing-up model of the reference (rlmpc/mpc/cartpole/acados.py:28-108):
//   x = [s, s_dot, theta, theta_dot], u = [F], model parameters (M, m, l) with g fixed (the YAML default), or
//   (M, m, l, g) when a driver un-fixes g (scripts/cartpole_mpc_qlearning.py:184-187: 84 parameters).
//   x+ = ONE explicit RK4 step of length h = tf/N/sim_method_num_stages (quirk Q1,
//   cartpole/acados.py:86-92, rlmpc/common/integrator.py:6-33).
//   cost y = [x;u], y_e = x (acados.py:104-106), NONLINEAR_LS with W_0/W/W_e, yref_*.
// theta layout (rlmpc/mpc/nlp.py:970-989, CasADi column-major):
//   [M, m, l (, g) | W_0(25) | W(25) | W_e(16) | yref_0(5) | yref(5) | yref_e(4)]  = 83 (84)
//
// First/second derivatives of the RK4 map are built by hand-written forward/adjoint
// propagation around sympy-generated leaf code (cartpole_gen.cuh) -- this replaces the
// CasADi-generated C the reference compiles at construction time.
#pragma once
#include "../common.cuh"

namespace rlmpc {

#include "cartpole_gen.cuh"

// NBX_ = number of box-constrained states on stages 1..N (0: config/cartpole_original.yaml,
// 4: config/cartpole.yaml with idxbx = [0,1,2,3]).  NPM_ = 3: (M, m, l) learnable, g = mc[1]; 4: g is theta[3].
template <int NBX_, int NPM_ = 3>
struct CartpoleModelT {
  static_assert(NPM_ == 3 || NPM_ == 4, "model parameters: (M, m, l) or (M, m, l, g)");
  static constexpr int NX = 4, NU = 1, NPM = NPM_, NZ = 5 + NPM_;  // NZ = NX+NU+NPM (derivative columns)
  static constexpr int NV = 3 + NPM_;                              // leaf variables (theta, theta_dot, F, parameters)
  static constexpr int NW = NX + NU;
  static constexpr int NBX = NBX_;
  static constexpr int NTH = 80 + NPM_;
  static constexpr int TH_W0 = NPM_, TH_W = TH_W0 + 25, TH_WE = TH_W + 25, TH_YREF0 = TH_WE + 16, TH_YREF = TH_YREF0 + 5,
                       TH_YREFE = TH_YREF + 5;
  MPC_HD static double grav(const double* th, size_t ths, const double* mc) { return NPM_ == 4 ? th[3 * ths] : mc[1]; }
  static constexpr int NSX = 0;              // no soft bounds
  static constexpr int NG = 0;               // no general linear rows
  MPC_HD static double gC(int, int) { return 0.0; }
  MPC_HD static double g0(int) { return 0.0; }
  static constexpr bool STAGE_HESS = false;
  static constexpr bool PARAMS_COST_ONLY = false;  // parameters enter the dynamics: dense per-stage parameter derivatives  // stage Hessians come from the cost table
  MPC_HD static int bx(int j) { return j; }  // idxbx (all states, in order)
  MPC_HD static int sx(int) { return 0; }

  // ---- quadratic tracking cost  l = 1/2 (y-yref)' W (y-yref),  y=[x;u] -------------------
  // kind: 0 initial stage, 1 intermediate, 2 terminal (ny = NX).
  MPC_HD static int ny(int kind) { return kind == 2 ? NX : NX + NU; }
  MPC_HD static int w_off(int kind) { return kind == 0 ? TH_W0 : (kind == 1 ? TH_W : TH_WE); }
  MPC_HD static int yref_off(int kind) { return kind == 0 ? TH_YREF0 : (kind == 1 ? TH_YREF : TH_YREFE); }
  // symmetrised weight entry (i,j): the Hessian of 1/2 e'We is (W+W')/2
  MPC_HD static double W(int kind, int i, int j, const double* th, size_t ths) {
    const int n = ny(kind), o = w_off(kind);
    return 0.5 * (th[(size_t)(o + j * n + i) * ths] + th[(size_t)(o + i * n + j) * ths]);
  }
  MPC_HD static double yref(int kind, int i, const double* th, size_t ths) {
    return th[(size_t)(yref_off(kind) + i) * ths];
  }
  // theta -> the engine's quadratic cost table (Engine::CT_*): per kind
  //   [W packed upper (15) | yref (5) | flin (5) | c0]
  MPC_HD static void cost_table(const double* th, size_t ths, double* ct, size_t cts, const double* /*mc*/) {
    constexpr int NWS = NW * (NW + 1) / 2, REC = NWS + 2 * NW + 1;
    for (int kind = 0; kind < 3; ++kind) {
      double* c = ct + (size_t)(kind * REC) * cts;
      const int n = ny(kind);
      int q = 0;
      for (int i = 0; i < NW; ++i)
        for (int j = i; j < NW; ++j) c[(size_t)(q++) * cts] = (i < n && j < n) ? W(kind, i, j, th, ths) : 0.0;
      for (int i = 0; i < NW; ++i) c[(size_t)(NWS + i) * cts] = (i < n) ? yref(kind, i, th, ths) : 0.0;
      for (int i = 0; i < NW; ++i) c[(size_t)(NWS + NW + i) * cts] = 0.0;  // no linear term
      c[(size_t)(NWS + 2 * NW) * cts] = 0.0;                              // no constant term
    }
  }
  MPC_HD static void cost_sens(int, double, const double*, const double*, size_t, double*, double*) {}  // no model parameter in the cost
  // d(s * l)/d(W, yref) accumulated into the [NTH] row (parameterize_tracking_cost=True semantics,
  // nlp.py:1057-1074): dl/dW_ij = 1/2 e_i e_j, dl/dyref = -W_sym e.
  MPC_HD static void cost_param_grad(int kind, double s, const double* th, size_t ths, const double* x, const double* u,
                                     double* dLdth) {
    const int n = ny(kind);
    double e[NW];
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      e[i] = 0.0;
      if (i < n) e[i] = ((i < NX) ? x[i] : u[i - NX]) - yref(kind, i, th, ths);
    }
    const int wo = w_off(kind), yo = yref_off(kind);
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      if (i >= n) continue;
      double a = 0.0;
      MPC_UNROLL for (int j = 0; j < NW; ++j) {
        if (j >= n) continue;
        a += W(kind, i, j, th, ths) * e[j];
        dLdth[wo + j * n + i] += 0.5 * s * e[i] * e[j];
      }
      dLdth[yo + i] -= s * a;
    }
  }

  // Cost-parameter columns of dpi/dtheta for one adjoint vector yw = [yx ; yu] of this stage (terminal: yu ignored):
  //   row[theta] -= yw' d(grad_w s l)/d theta,   d(grad l)/dW_ij = 1/2 (delta_i e_j + delta_j e_i),  d(grad l)/dyref = -W_sym
  MPC_HD static void cost_param_adj(int kind, double s, const double* th, size_t ths, const double* x, const double* u,
                                    const double* yw, double* row) {
    const int n = ny(kind);
    double e[NW];
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      e[i] = 0.0;
      if (i < n) e[i] = ((i < NX) ? x[i] : u[i - NX]) - yref(kind, i, th, ths);
    }
    const int wo = w_off(kind), yo = yref_off(kind);
    MPC_UNROLL for (int j = 0; j < NW; ++j) {
      if (j >= n) continue;
      double a = 0.0;
      MPC_UNROLL for (int i = 0; i < NW; ++i) {
        if (i >= n) continue;
        a += yw[i] * W(kind, i, j, th, ths);
        row[wo + j * n + i] -= 0.5 * s * (yw[i] * e[j] + yw[j] * e[i]);
      }
      row[yo + j] += s * a;
    }
  }

  // sin / cos of the pole angle at the later RK stage points: theta_0 + d with |d| = O(h theta_dot) small, so the angle
  // addition formulas with degree-13 / degree-12 Taylor polynomials of sin d / cos d (error < 3e-18 for |d| <= 0.25)
  // replace three of the four FP64 sincos calls of a step (ncu: sincos was 24 % of the instructions of k_lin).
  MPC_HD static void stage_sincos(double d, double angle, double sn0, double cs0, double* sn, double* cs) {
    if (!(dabs(d) <= 0.25)) {
      sincos(angle, sn, cs);
      return;
    }
    const double q = d * d;  // Horner, coefficients (-1)^k / (2k+1)!  and  (-1)^k / (2k)!
    const double ps = -1.0 / 6.0 + q * (1.0 / 120.0 + q * (-1.0 / 5040.0 + q * (1.0 / 362880.0 + q * (-1.0 / 39916800.0 + q * (1.0 / 6227020800.0)))));
    const double pc = -0.5 + q * (1.0 / 24.0 + q * (-1.0 / 720.0 + q * (1.0 / 40320.0 + q * (-1.0 / 3628800.0 + q * (1.0 / 479001600.0)))));
    const double sd = d + (d * q) * ps;
    const double cd = 1.0 + q * pc;
    *sn = sn0 * cd + cs0 * sd;
    *cs = cs0 * cd - sn0 * sd;
  }

  // ---- dynamics -------------------------------------------------------------------------
  // One RK4 step with forward propagation of d(.)/d zeta, zeta = (x0..x3, F, M, m, l (, g));
  // NC = 5 -> columns (x,u) only (SQP linearisation), NC = NZ -> also the parameter columns.
  // The accelerations depend on v = (theta, theta_dot, F, M, m, l (, g)) = zeta[2..] only -- neither on the cart position
  // nor on its velocity -- so the columns of s and s_dot are known in closed form (dF/ds = e_0, dF/ds_dot = (h, 1, 0, 0)')
  // and carry no curvature: only the NE = NC - 2 columns of v are propagated (column e <-> zeta[e + 2]).
  template <int NC>
  MPC_HD static void rk4_fwd(const double* x, double F, double M, double m, double l, double g, double h,
                             double* xn, double* DF /* 4 x NC row-major */,
                             double* keep /* optional per-stage data for the adjoint pass, or nullptr */) {
    constexpr int NE = NC - 2;
    double S[4][NE], acc[4][NE], s[4], xa[4], sn0 = 0.0, cs0 = 1.0;
    MPC_UNROLL for (int i = 0; i < 4; ++i) {
      s[i] = x[i];
      xa[i] = 0.0;
      MPC_UNROLL for (int e = 0; e < NE; ++e) {
        S[i][e] = (i == e + 2) ? 1.0 : 0.0;
        acc[i][e] = 0.0;
      }
    }
    MPC_UNROLL for (int st = 0; st < 4; ++st) {
      double sn, cs, xdd, thdd, jx[NV], jt[NV];
      if (st == 0) {
        sincos(s[2], &sn, &cs);
        sn0 = sn; cs0 = cs;
      } else {
        stage_sincos(s[2] - x[2], s[2], sn0, cs0, &sn, &cs);
      }
      if constexpr (NPM_ == 4 && NC > 5) cartpole_f_jac_g(sn, cs, s[3], F, M, m, l, g, &xdd, &thdd, jx, jt);
      else cartpole_f_jac(sn, cs, s[3], F, M, m, l, g, &xdd, &thdd, jx, jt);
      if (keep) {  // stage point + the S rows the adjoint pass needs (first stage: unit vectors, not stored)
        double* kp = keep + st * (3 + 2 * NE);
        kp[0] = sn; kp[1] = cs; kp[2] = s[3];
        if (st > 0) {
          MPC_UNROLL for (int e = 0; e < NE; ++e) { kp[3 + e] = S[2][e]; kp[3 + NE + e] = S[3][e]; }
        }
      }
      const double k[4] = {s[1], xdd, s[3], thdd};
      double Dk[4][NE];
      MPC_UNROLL for (int e = 0; e < NE; ++e) {
        if (st == 0) {  // S = [0 ; 0 ; e_0 ; e_1] at the first stage: written out, a product with an exact zero is not folded away in IEEE arithmetic
          Dk[0][e] = 0.0;
          Dk[2][e] = (e == 1) ? 1.0 : 0.0;
          Dk[1][e] = jx[e];  // v = zeta[2..]: the Jacobian row itself
          Dk[3][e] = jt[e];
          continue;
        }
        Dk[0][e] = S[1][e];
        Dk[2][e] = S[3][e];
        double a = jx[0] * S[2][e] + jx[1] * S[3][e];
        double b = jt[0] * S[2][e] + jt[1] * S[3][e];
        if (e >= 2) { a += jx[e]; b += jt[e]; }
        Dk[1][e] = a;
        Dk[3][e] = b;
      }
      const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
      MPC_UNROLL for (int i = 0; i < 4; ++i) {
        xa[i] += wgt * k[i];
        MPC_UNROLL for (int e = 0; e < NE; ++e) acc[i][e] += wgt * Dk[i][e];
      }
      if (st < 3) {
        const double a = (st < 2) ? 0.5 * h : h;
        MPC_UNROLL for (int i = 0; i < 4; ++i) {
          s[i] = x[i] + a * k[i];
          MPC_UNROLL for (int e = 0; e < NE; ++e) S[i][e] = ((i == e + 2) ? 1.0 : 0.0) + a * Dk[i][e];
        }
      }
    }
    const double h6 = h / 6.0;
    MPC_UNROLL for (int i = 0; i < 4; ++i) {
      xn[i] = x[i] + h6 * xa[i];
      DF[i * NC + 0] = (i == 0) ? 1.0 : 0.0;
      DF[i * NC + 1] = (i == 0) ? h6 * 6.0 : (i == 1) ? 1.0 : 0.0;  // sum of the RK weights of ds_dot/ds_dot = 1
      MPC_UNROLL for (int e = 0; e < NE; ++e) DF[i * NC + 2 + e] = ((i == e + 2) ? 1.0 : 0.0) + h6 * acc[i][e];
    }
  }

  // x+ and its Jacobians wrt x (A, NX x NX row-major) and u (B, NX x NU)
  MPC_HD static void dyn_lin(const double* x, const double* u, const double* th, size_t ths, const double* mc,
                             double* xn, double* A, double* B) {
    double DF[4 * 5];
    rk4_fwd<5>(x, u[0], th[0], th[ths], th[2 * ths], grav(th, ths, mc), mc[0], xn, DF, nullptr);
    MPC_UNROLL for (int i = 0; i < 4; ++i) {
      MPC_UNROLL for (int j = 0; j < 4; ++j) A[i * 4 + j] = DF[i * 5 + j];
      B[i] = DF[i * 5 + 4];
    }
  }

  // Full second-order information at one stage:
  //   xn, A, B, Fp = dF/dp_model (NX x NPM), and the Hessian of pi' F wrt zeta split into
  //   Hww ((NX+NU) x (NX+NU), full symmetric storage) and Hwp ((NX+NU) x NPM).
  MPC_HD static void dyn_sens(const double* x, const double* u, const double* th, size_t ths, const double* mc,
                              const double* pi, double* xn, double* A, double* B, double* Fp, double* Hww,
                              double* Hwp) {
    constexpr int NC = NZ, NE = NC - 2, KS = 3 + 2 * NE;
    static_assert(NE == NV, "the propagated columns are the leaf variables v = zeta[2..]");
    const double F = u[0], M = th[0], m = th[ths], l = th[2 * ths], g = grav(th, ths, mc), h = mc[0];
    double DF[4 * NC], keep[4 * KS];
    rk4_fwd<NC>(x, F, M, m, l, g, h, xn, DF, keep);
    MPC_UNROLL for (int i = 0; i < 4; ++i) {
      MPC_UNROLL for (int j = 0; j < 4; ++j) A[i * 4 + j] = DF[i * NC + j];
      B[i] = DF[i * NC + 4];
      MPC_UNROLL for (int j = 0; j < NPM; ++j) Fp[i * NPM + j] = DF[i * NC + 5 + j];
    }
    // adjoint sweep over the RK stages: mu_i = d(pi'F)/dk_i.  Curvature lives in the v-columns only (see rk4_fwd):
    // Hacc is the NE x NE block of the Hessian over zeta[2..], its rows / columns of s and s_dot are zero.
    double Hacc[NE][NE];
    MPC_UNROLL for (int a = 0; a < NE; ++a) MPC_UNROLL for (int b = 0; b < NE; ++b) Hacc[a][b] = 0.0;
    double mu[4];
    MPC_UNROLL for (int i = 0; i < 4; ++i) mu[i] = (h / 6.0) * pi[i];
    MPC_UNROLL for (int st = 3; st >= 0; --st) {
      const double* kp = keep + st * KS;
      const double sn = kp[0], cs = kp[1], thd = kp[2];
      double xdd, thdd, jx[NV], jt[NV], hs[NV * NV];
      if constexpr (NPM_ == 4) cartpole_f_hess_g(sn, cs, thd, F, M, m, l, g, mu[1], mu[3], &xdd, &thdd, jx, jt, hs);
      else cartpole_f_hess(sn, cs, thd, F, M, m, l, g, mu[1], mu[3], &xdd, &thdd, jx, jt, hs);
      // symmetric NV x NV Hessian of mu1*xdd + mu3*thdd wrt v = (theta, theta_dot, F, M, m, l (, g))
      double Hf[NV][NV];
      MPC_UNROLL for (int a = 0; a < NV; ++a) MPC_UNROLL for (int b = a; b < NV; ++b) {
        Hf[a][b] = hs[a * NV + b];
        Hf[b][a] = hs[a * NV + b];
      }
      const double* Sr0 = kp + 3;       // d s[2] / d v
      const double* Sr1 = kp + 3 + NE;  // d s[3] / d v
      if (st == 0) {  // d (theta, theta_dot) / d v = [e_0 ; e_1] at the first stage: D = I
        MPC_UNROLL for (int a = 0; a < NE; ++a) MPC_UNROLL for (int b = a; b < NE; ++b) Hacc[a][b] += Hf[a][b];
      } else {
        double T[NV][NE];
        MPC_UNROLL for (int p = 0; p < NV; ++p) MPC_UNROLL for (int b = 0; b < NE; ++b) {
          double v = Hf[p][0] * Sr0[b] + Hf[p][1] * Sr1[b];
          if (b >= 2) v += Hf[p][b];
          T[p][b] = v;
        }
        MPC_UNROLL for (int a = 0; a < NE; ++a) MPC_UNROLL for (int b = a; b < NE; ++b) {  // symmetric: upper triangle
          double v = Sr0[a] * T[0][b] + Sr1[a] * T[1][b];
          if (a >= 2) v += T[a][b];
          Hacc[a][b] += v;
        }
      }
      if (st > 0) {
        // adjoint of the stage point s_st = x + a*k_{st-1}:  mu_{st-1} = w*h/6*pi + a * (df/ds)' mu_st
        const double a = (st == 3) ? h : 0.5 * h;
        const double wgt = (st - 1 == 0) ? 1.0 : 2.0;
        const double jx0 = jx[0], jx1 = jx[1], jt0 = jt[0], jt1 = jt[1];  // Jacobian at this stage point (just re-evaluated)
        const double a0 = 0.0;
        const double a1 = mu[0];
        const double a2 = jx0 * mu[1] + jt0 * mu[3];
        const double a3 = mu[2] + jx1 * mu[1] + jt1 * mu[3];
        const double c6 = wgt * h / 6.0;
        mu[0] = c6 * pi[0] + a * a0;
        mu[1] = c6 * pi[1] + a * a1;
        mu[2] = c6 * pi[2] + a * a2;
        mu[3] = c6 * pi[3] + a * a3;
      }
    }
    MPC_UNROLL for (int a = 0; a < 5; ++a) {
      MPC_UNROLL for (int b = 0; b < 5; ++b)
        Hww[a * 5 + b] = (a < 2 || b < 2) ? 0.0 : ((a <= b) ? Hacc[a - 2][b - 2] : Hacc[b - 2][a - 2]);
      MPC_UNROLL for (int b = 0; b < NPM; ++b) Hwp[a * NPM + b] = (a < 2) ? 0.0 : Hacc[a - 2][3 + b];
    }
  }
};

using CartpoleModel = CartpoleModelT<0>;      // input bounds only (config/cartpole_original.yaml)
using CartpoleModelBX = CartpoleModelT<4>;    // + box bounds on all states (config/cartpole.yaml)
using CartpoleModelG = CartpoleModelT<0, 4>;  // input bounds only, g un-fixed: theta = [M, m, l, g | W ...] (84 entries)
using CartpoleModelBXG = CartpoleModelT<4, 4>;

}  // namespace rlmpc
