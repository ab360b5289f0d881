// Linear system MPC of the reference (rlmpc/mpc/linear_system/acados.py:27-131):
//   x+ = A x + B u + b,  nx = 2, nu = 1, N = 40, all 12 parameters learnable:
//   theta = [A (col-major, 4) | B (2) | b (2) | V_0 (1) | f (3)]        (acados.py:60-70, 89-90)
//   stage cost  l = 1/2 y'y + f'y,  y = [x;u]  (+ V_0 at stage 0)        (acados.py:34-49)
//   terminal    l_e = 1/2 x'P x,  P = DARE(A,B,Q,R) of the INITIAL parameters, a constant (acados.py:51-57)
//   u in [-1,1]; x in [0,1] x [-1,1] on stages 1..N-1, the bound on x_0 soft (idxsbx=[0], zl=zu=1e2)
//   EXTERNAL cost, EXACT Hessian (= the constant cost Hessian, the dynamics are linear)
// Model constants: mc[0..2] = P11, P12, P22.
#pragma once
#include "../common.cuh"

namespace rlmpc {

struct LinearSystemModel {
  static constexpr int NX = 2, NU = 1, NW = 3, NPM = 12, NTH = 12;
  static constexpr int NBX = 2, NSX = 1;
  static constexpr int NG = 0;               // no general linear rows
  MPC_HD static double gC(int, int) { return 0.0; }
  MPC_HD static double g0(int) { return 0.0; }
  static constexpr bool STAGE_HESS = false;
  static constexpr bool PARAMS_COST_ONLY = false;  // parameters enter the dynamics: dense per-stage parameter derivatives  // stage Hessians come from the cost table
  static constexpr int TH_A = 0, TH_B = 4, TH_b = 6, TH_V0 = 8, TH_F = 9;
  MPC_HD static int bx(int j) { return j; }  // idxbx = [0, 1]
  MPC_HD static int sx(int) { return 0; }    // idxsbx = [0]: position in idxbx

  MPC_HD static void cost_table(const double* th, size_t ths, double* ct, size_t cts, const double* mc) {
    constexpr int NWS = NW * (NW + 1) / 2, REC = NWS + 2 * NW + 1;
    for (int kind = 0; kind < 3; ++kind) {
      double* c = ct + (size_t)(kind * REC) * cts;
      // packed upper triangle, row-major: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
      const double Wst[NWS] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
      const double Wte[NWS] = {mc[0], mc[1], 0.0, mc[2], 0.0, 0.0};
      for (int q = 0; q < NWS; ++q) c[(size_t)q * cts] = (kind == 2) ? Wte[q] : Wst[q];
      for (int i = 0; i < NW; ++i) c[(size_t)(NWS + i) * cts] = 0.0;  // yref
      for (int i = 0; i < NW; ++i) c[(size_t)(NWS + NW + i) * cts] = (kind == 2) ? 0.0 : th[(size_t)(TH_F + i) * ths];
      c[(size_t)(NWS + 2 * NW) * cts] = (kind == 0) ? th[(size_t)TH_V0 * ths] : 0.0;
    }
  }
  // no W / yref entries in p (EXTERNAL cost)
  MPC_HD static void cost_param_grad(int, double, const double*, size_t, const double*, const double*, double*) {}
  MPC_HD static void cost_param_adj(int, double, const double*, size_t, const double*, const double*, const double*, double*) {}
  // cost terms that depend on model parameters: d(s l)/d theta -> gp, d(grad_w s l)/d theta -> Hwp
  MPC_HD static void cost_sens(int kind, double s, const double* y, const double*, size_t, double* gp, double* Hwp) {
    if (kind == 2) return;
    if (kind == 0) gp[TH_V0] += s;
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      gp[TH_F + i] += s * y[i];
      Hwp[i * NPM + TH_F + i] += s;
    }
  }

  MPC_HD static void dyn_lin(const double* x, const double* u, const double* th, size_t ths, const double*,
                             double* xn, double* A, double* B) {
    MPC_UNROLL for (int i = 0; i < 2; ++i) {
      MPC_UNROLL for (int j = 0; j < 2; ++j) A[i * 2 + j] = th[(size_t)(TH_A + j * 2 + i) * ths];
      B[i] = th[(size_t)(TH_B + i) * ths];
      xn[i] = A[i * 2] * x[0] + A[i * 2 + 1] * x[1] + B[i] * u[0] + th[(size_t)(TH_b + i) * ths];
    }
  }

  MPC_HD static void dyn_sens(const double* x, const double* u, const double* th, size_t ths, const double* mc,
                              const double* pi, double* xn, double* A, double* B, double* Fp, double* Hww,
                              double* Hwp) {
    dyn_lin(x, u, th, ths, mc, xn, A, B);
    MPC_UNROLL for (int i = 0; i < NX * NPM; ++i) Fp[i] = 0.0;
    MPC_UNROLL for (int i = 0; i < NW * NW; ++i) Hww[i] = 0.0;
    MPC_UNROLL for (int i = 0; i < NW * NPM; ++i) Hwp[i] = 0.0;
    MPC_UNROLL for (int i = 0; i < 2; ++i) {
      MPC_UNROLL for (int j = 0; j < 2; ++j) {
        Fp[i * NPM + TH_A + j * 2 + i] = x[j];    // dF_i / dA_ij
        Hwp[j * NPM + TH_A + j * 2 + i] = pi[i];  // d2(pi'F) / dx_j dA_ij
      }
      Fp[i * NPM + TH_B + i] = u[0];
      Hwp[2 * NPM + TH_B + i] = pi[i];
      Fp[i * NPM + TH_b + i] = 1.0;
    }
  }
};

}  // namespace rlmpc
