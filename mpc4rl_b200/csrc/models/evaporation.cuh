// Evaporation-process economic NMPC of the reference (rlmpc/mpc/evaporation_process/acados.py:142-228,
// plant equations rlmpc/gym/evaporation_process/environment.py:50-106):
//   x = [X_2, P_2], u = [P_100, F_200, s]  (s: slack input of the soft constraint), N = 100, dT = 1
//   x+ = sim_method_num_stages (= 4) explicit RK4 steps of length dT/4              (acados.py:74-86)
//   cost NONLINEAR_LS with y = [x;u], yref = (x_ss,u_ss), W = H (5x5), stage k weighted gamma^k; no
//   terminal cost                                                                    (acados.py:160-176, 219-226)
//   u in [100,400]^2 x [0,10];  h = 25 - x - s in [-1e3, 0] (two affine rows)        (acados.py:202-210)
// theta = the tracking-cost parameters of build_nlp(parameterize_tracking_cost=True) (nlp.py:970-989,
//   1057-1074): [W_0 (25, col-major) | W (25) | yref_0 (5) | yref (5)] = 60.  model.p is empty.
//   (W_0/yref_0 present: acados fills the stage-0 cost from the path cost when cost_type_0 is unset,
//   SURVEY.md a12; the 30-parameter variant would leave stage 0 without a cost in the NLP.)
// Model constants mc[0..18] = environment.PARAM in dict order, mc[19] = RK4 step, mc[20] = number of steps.
#pragma once
#include "../common.cuh"

namespace rlmpc {

#include "evaporation_gen.cuh"

struct EvaporationModel {
  static constexpr int NX = 2, NU = 3, NW = 5, NPM = 60, NTH = 60;
  static constexpr int NBX = 0, NSX = 0, NG = 2;
  static constexpr int NZ = 4;  // quantities the dynamics depend on: (X_2, P_2, P_100, F_200)
  static constexpr int TH_W0 = 0, TH_W = 25, TH_Y0 = 50, TH_Y = 55;
  static constexpr int MAXSUB = 8;
  MPC_HD static int bx(int j) { return j; }
  MPC_HD static int sx(int) { return 0; }
  // g_j = 25 - x_j - s,  lh <= g <= uh
  MPC_HD static double gC(int j, int i) { return (i == j || i == NX + 2) ? -1.0 : 0.0; }
  MPC_HD static double g0(int) { return 25.0; }
  static constexpr bool STAGE_HESS = false;

  // ---- cost --------------------------------------------------------------------------------
  MPC_HD static int w_off(int kind) { return kind == 0 ? TH_W0 : TH_W; }
  MPC_HD static int y_off(int kind) { return kind == 0 ? TH_Y0 : TH_Y; }
  MPC_HD static double Wsym(int kind, int i, int j, const double* th, size_t ths) {
    const int o = w_off(kind);
    return 0.5 * (th[(size_t)(o + j * NW + i) * ths] + th[(size_t)(o + i * NW + j) * ths]);
  }
  MPC_HD static void cost_table(const double* th, size_t ths, double* ct, size_t cts, const double* /*mc*/) {
    constexpr int NWS = NW * (NW + 1) / 2, REC = NWS + 2 * NW + 1;
    for (int kind = 0; kind < 3; ++kind) {
      double* c = ct + (size_t)(kind * REC) * cts;
      int q = 0;
      for (int i = 0; i < NW; ++i)
        for (int j = i; j < NW; ++j) c[(size_t)(q++) * cts] = (kind == 2) ? 0.0 : Wsym(kind, i, j, th, ths);
      for (int i = 0; i < NW; ++i) c[(size_t)(NWS + i) * cts] = (kind == 2) ? 0.0 : th[(size_t)(y_off(kind) + i) * ths];
      for (int i = 0; i < NW; ++i) c[(size_t)(NWS + NW + i) * cts] = 0.0;
      c[(size_t)(NWS + 2 * NW) * cts] = 0.0;
    }
  }
  MPC_HD static void cost_param_grad(int, double, const double*, size_t, const double*, const double*, double*) {}
  MPC_HD static void cost_param_adj(int, double, const double*, size_t, const double*, const double*, const double*, double*) {}
  // d(s l)/d theta -> gp,  d(grad_w s l)/d theta -> Hwp  for l = 1/2 e'We, e = y - yref:
  //   dl/dW_ij = 1/2 e_i e_j,  dl/dyref = -W_s e,  d(grad l)_a/dW_ij = 1/2 (d_ai e_j + d_aj e_i),  d(grad l)/dyref = -W_s
  MPC_HD static void cost_sens(int kind, double s, const double* y, const double* th, size_t ths, double* gp, double* Hwp) {
    if (kind == 2) return;
    const int wo = w_off(kind), yo = y_off(kind);
    double e[NW];
    MPC_UNROLL for (int i = 0; i < NW; ++i) e[i] = y[i] - th[(size_t)(yo + i) * ths];
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      double a = 0.0;
      MPC_UNROLL for (int j = 0; j < NW; ++j) {
        const double ws = Wsym(kind, i, j, th, ths);
        a += ws * e[j];
        Hwp[i * NPM + yo + j] -= s * ws;
        const int idx = wo + j * NW + i;  // W_ij, column-major
        gp[idx] += 0.5 * s * e[i] * e[j];
        Hwp[i * NPM + idx] += 0.5 * s * e[j];
        Hwp[j * NPM + idx] += 0.5 * s * e[i];
      }
      gp[yo + i] -= s * a;
    }
  }

  // All 60 parameters sit in the stage cost only (model.p is empty): the engine then does not store the dense
  // d(grad_w L)/d theta (5 x 60) and dF/d theta (2 x 60) of every stage but asks for the two contractions it
  // needs while it sweeps:  gp += d(s l)/d theta   and   acc -= yw' d(grad_w s l)/d theta  (yw: adjoint of [x;u]).
  static constexpr bool PARAMS_COST_ONLY = true;
  MPC_HD static void cost_sens_grad(int kind, double s, const double* y, const double* th, size_t ths, double* gp) {
    if (kind == 2) return;
    const int wo = w_off(kind), yo = y_off(kind);
    double e[NW];
    MPC_UNROLL for (int i = 0; i < NW; ++i) e[i] = y[i] - th[(size_t)(yo + i) * ths];
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      double a = 0.0;
      MPC_UNROLL for (int j = 0; j < NW; ++j) {
        a += Wsym(kind, i, j, th, ths) * e[j];
        gp[wo + j * NW + i] += 0.5 * s * e[i] * e[j];
      }
      gp[yo + i] -= s * a;
    }
  }
  MPC_HD static void cost_sens_adj(int kind, double s, const double* y, const double* th, size_t ths, const double* yw,
                                   double* acc) {
    if (kind == 2) return;
    const int wo = w_off(kind), yo = y_off(kind);
    double e[NW];
    MPC_UNROLL for (int i = 0; i < NW; ++i) e[i] = y[i] - th[(size_t)(yo + i) * ths];
    MPC_UNROLL for (int j = 0; j < NW; ++j) {
      double a = 0.0;
      MPC_UNROLL for (int i = 0; i < NW; ++i) {
        a += yw[i] * Wsym(kind, i, j, th, ths);
        acc[wo + j * NW + i] -= 0.5 * s * (yw[i] * e[j] + yw[j] * e[i]);  // W_ij, column-major
      }
      acc[yo + j] += s * a;  // d(grad l)/d yref = -W_s
    }
  }

  // ---- dynamics ------------------------------------------------------------------------------
  // One RK4 step Phi(x, ud) of length h with forward propagation of d/d(x, ud) (ud = the two inputs
  // that enter f).  D: 2 x 4 row-major.  keep (optional): per RK stage [s(2) | S(2x4) | jac(2x4)].
  static constexpr int KEEP = 2 + 8 + 8;
  MPC_HD static void rk4_fwd(const double* x, const double* ud, const double* pm, double h, double* xn, double* D,
                             double* keep) {
    double s[2] = {x[0], x[1]}, S[8], acc[8], xa[2] = {0.0, 0.0};
    MPC_UNROLL for (int i = 0; i < 2; ++i) MPC_UNROLL for (int c = 0; c < 4; ++c) {
      S[i * 4 + c] = (i == c) ? 1.0 : 0.0;
      acc[i * 4 + c] = 0.0;
    }
    MPC_UNROLL for (int st = 0; st < 4; ++st) {
      double f[2], jac[8], Dk[8];
      evaporation_f_jac(s[0], s[1], ud[0], ud[1], pm, f, jac);
      if (keep) {
        double* kp = keep + st * KEEP;
        kp[0] = s[0]; kp[1] = s[1];
        MPC_UNROLL for (int i = 0; i < 8; ++i) { kp[2 + i] = S[i]; kp[10 + i] = jac[i]; }
      }
      MPC_UNROLL for (int i = 0; i < 2; ++i) MPC_UNROLL for (int c = 0; c < 4; ++c) {
        double a = jac[i * 4 + 0] * S[0 * 4 + c] + jac[i * 4 + 1] * S[1 * 4 + c];
        if (c >= 2) a += jac[i * 4 + c];
        Dk[i * 4 + c] = a;
      }
      const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
      MPC_UNROLL for (int i = 0; i < 2; ++i) {
        xa[i] += wgt * f[i];
        MPC_UNROLL for (int c = 0; c < 4; ++c) acc[i * 4 + c] += wgt * Dk[i * 4 + c];
      }
      if (st < 3) {
        const double a = (st < 2) ? 0.5 * h : h;
        MPC_UNROLL for (int i = 0; i < 2; ++i) {
          s[i] = x[i] + a * f[i];
          MPC_UNROLL for (int c = 0; c < 4; ++c) S[i * 4 + c] = ((i == c) ? 1.0 : 0.0) + a * Dk[i * 4 + c];
        }
      }
    }
    MPC_UNROLL for (int i = 0; i < 2; ++i) {
      xn[i] = x[i] + h / 6.0 * xa[i];
      MPC_UNROLL for (int c = 0; c < 4; ++c) D[i * 4 + c] = ((i == c) ? 1.0 : 0.0) + h / 6.0 * acc[i * 4 + c];
    }
  }
  // Hessian of lam' Phi w.r.t. (x, ud) (4 x 4, added to H) from the kept stage data
  MPC_HD static void rk4_hess(const double* keep, const double* ud, const double* pm, double h, const double* lam, double* H) {
    double mu[2] = {h / 6.0 * lam[0], h / 6.0 * lam[1]};
    MPC_UNROLL for (int st = 3; st >= 0; --st) {
      const double* kp = keep + st * KEEP;
      const double* S = kp + 2;
      const double* jac = kp + 10;
      double f[2], jj[8], hs[16];
      evaporation_f_hess(kp[0], kp[1], ud[0], ud[1], pm, mu[0], mu[1], f, jj, hs);
      double Hf[16];
      MPC_UNROLL for (int a = 0; a < 4; ++a) MPC_UNROLL for (int b = a; b < 4; ++b) {
        Hf[a * 4 + b] = hs[a * 4 + b];
        Hf[b * 4 + a] = hs[a * 4 + b];
      }
      // D = [S ; 0 I] maps (x, ud) to (stage point, ud):  H += D' Hf D
      double T[16];  // T = Hf D
      MPC_UNROLL for (int p = 0; p < 4; ++p) MPC_UNROLL for (int b = 0; b < 4; ++b) {
        double v = Hf[p * 4 + 0] * S[0 * 4 + b] + Hf[p * 4 + 1] * S[1 * 4 + b];
        if (b >= 2) v += Hf[p * 4 + b];
        T[p * 4 + b] = v;
      }
      MPC_UNROLL for (int a = 0; a < 4; ++a) MPC_UNROLL for (int b = 0; b < 4; ++b) {
        double v = S[0 * 4 + a] * T[0 * 4 + b] + S[1 * 4 + a] * T[1 * 4 + b];
        if (a >= 2) v += T[a * 4 + b];
        H[a * 4 + b] += v;
      }
      if (st > 0) {
        const double a = (st == 3) ? h : 0.5 * h;
        const double wgt = (st - 1 == 0) ? 1.0 : 2.0;
        const double m0 = jac[0] * mu[0] + jac[4] * mu[1];  // (df/dx)' mu
        const double m1 = jac[1] * mu[0] + jac[5] * mu[1];
        mu[0] = wgt * h / 6.0 * lam[0] + a * m0;
        mu[1] = wgt * h / 6.0 * lam[1] + a * m1;
      }
    }
  }

  // x+ = Phi^nsub (x, ud); G = d x+ / d(x, ud) (2 x 4)
  MPC_HD static void dyn_lin(const double* x, const double* u, const double*, size_t, const double* mc,
                             double* xn, double* A, double* B) {
    const double h = mc[19];
    const int nsub = (int)mc[20];
    double xc[2] = {x[0], x[1]}, G[8] = {1, 0, 0, 0, 0, 1, 0, 0};
    for (int j = 0; j < nsub; ++j) {
      double xnn[2], D[8], Gn[8];
      rk4_fwd(xc, u, mc, h, xnn, D, nullptr);
      MPC_UNROLL for (int i = 0; i < 2; ++i) MPC_UNROLL for (int c = 0; c < 4; ++c) {
        double a = D[i * 4 + 0] * G[0 * 4 + c] + D[i * 4 + 1] * G[1 * 4 + c];
        if (c >= 2) a += D[i * 4 + c];
        Gn[i * 4 + c] = a;
      }
      MPC_UNROLL for (int i = 0; i < 8; ++i) G[i] = Gn[i];
      xc[0] = xnn[0]; xc[1] = xnn[1];
    }
    xn[0] = xc[0]; xn[1] = xc[1];
    MPC_UNROLL for (int i = 0; i < 2; ++i) {
      A[i * 2 + 0] = G[i * 4 + 0]; A[i * 2 + 1] = G[i * 4 + 1];
      B[i * 3 + 0] = G[i * 4 + 2]; B[i * 3 + 1] = G[i * 4 + 3]; B[i * 3 + 2] = 0.0;  // s does not enter f
    }
  }

  MPC_HD static void dyn_sens(const double* x, const double* u, const double* th, size_t ths, const double* mc,
                              const double* pi, double* xn, double* A, double* B, double* Fp, double* Hww, double* Hwp) {
    const double h = mc[19];
    int nsub = (int)mc[20];
    if (nsub > MAXSUB) nsub = MAXSUB;
    double xs[MAXSUB][2], Gs[MAXSUB][8], Ds[MAXSUB][8];
    double xc[2] = {x[0], x[1]}, G[8] = {1, 0, 0, 0, 0, 1, 0, 0};
    for (int j = 0; j < nsub; ++j) {
      double xnn[2], Gn[8];
      xs[j][0] = xc[0]; xs[j][1] = xc[1];
      MPC_UNROLL for (int i = 0; i < 8; ++i) Gs[j][i] = G[i];
      rk4_fwd(xc, u, mc, h, xnn, Ds[j], nullptr);
      MPC_UNROLL for (int i = 0; i < 2; ++i) MPC_UNROLL for (int c = 0; c < 4; ++c) {
        double a = Ds[j][i * 4 + 0] * G[0 * 4 + c] + Ds[j][i * 4 + 1] * G[1 * 4 + c];
        if (c >= 2) a += Ds[j][i * 4 + c];
        Gn[i * 4 + c] = a;
      }
      MPC_UNROLL for (int i = 0; i < 8; ++i) G[i] = Gn[i];
      xc[0] = xnn[0]; xc[1] = xnn[1];
    }
    xn[0] = xc[0]; xn[1] = xc[1];
    MPC_UNROLL for (int i = 0; i < 2; ++i) {
      A[i * 2 + 0] = G[i * 4 + 0]; A[i * 2 + 1] = G[i * 4 + 1];
      B[i * 3 + 0] = G[i * 4 + 2]; B[i * 3 + 1] = G[i * 4 + 3]; B[i * 3 + 2] = 0.0;
    }
    // second order: lam_nsub = pi; H_total = sum_j Gt_j' Hess(lam_{j+1}' Phi)(x_j, ud) Gt_j, Gt_j = [G_j ; 0 I]
    double Hz[16];
    MPC_UNROLL for (int i = 0; i < 16; ++i) Hz[i] = 0.0;
    double lam[2] = {pi[0], pi[1]};
    for (int j = nsub - 1; j >= 0; --j) {
      double keep[4 * KEEP], xnn[2], D[8], Hj[16];
      rk4_fwd(xs[j], u, mc, h, xnn, D, keep);
      MPC_UNROLL for (int i = 0; i < 16; ++i) Hj[i] = 0.0;
      rk4_hess(keep, u, mc, h, lam, Hj);
      const double* Gj = Gs[j];
      double T[16];
      MPC_UNROLL for (int p = 0; p < 4; ++p) MPC_UNROLL for (int b = 0; b < 4; ++b) {
        double v = Hj[p * 4 + 0] * Gj[0 * 4 + b] + Hj[p * 4 + 1] * Gj[1 * 4 + b];
        if (b >= 2) v += Hj[p * 4 + b];
        T[p * 4 + b] = v;
      }
      MPC_UNROLL for (int a = 0; a < 4; ++a) MPC_UNROLL for (int b = 0; b < 4; ++b) {
        double v = Gj[0 * 4 + a] * T[0 * 4 + b] + Gj[1 * 4 + a] * T[1 * 4 + b];
        if (a >= 2) v += T[a * 4 + b];
        Hz[a * 4 + b] += v;
      }
      const double l0 = D[0] * lam[0] + D[4] * lam[1], l1 = D[1] * lam[0] + D[5] * lam[1];
      lam[0] = l0; lam[1] = l1;
    }
    MPC_UNROLL for (int i = 0; i < NW * NW; ++i) Hww[i] = 0.0;
    MPC_UNROLL for (int a = 0; a < 4; ++a) MPC_UNROLL for (int b = 0; b < 4; ++b) Hww[a * NW + b] = Hz[a * 4 + b];  // z = w[0..4)
    MPC_UNROLL for (int i = 0; i < NX * NPM; ++i) Fp[i] = 0.0;   // no model parameters
    MPC_UNROLL for (int i = 0; i < NW * NPM; ++i) Hwp[i] = 0.0;  // filled by cost_sens
    (void)th; (void)ths;
  }
};

}  // namespace rlmpc
