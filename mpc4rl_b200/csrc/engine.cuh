// rlmpc-b200 engine: per-sample SQP (Gauss-Newton / exact) with a Riccati-structured
// primal-dual interior-point QP solve, followed by the exact-Hessian adjoint KKT solve that
// yields dpi/dtheta and the stage sweep that yields dL/dtheta (= dV/dtheta = dQ/dtheta).
//
// What this replaces in the reference:
//   a4  AcadosOcpSolver.solve()  (acados SQP + HPIPM Riccati IPM)   -> Engine::solve
//   a5  update_nlp()  rlmpc/mpc/nlp.py:1341-1563 (dense dR/dz + SuperLU, dL/dp) -> Engine::sens
// The maths is specified by rlmpc/mpc/nlp.py:884-1275:
//   L = cost + lam'h + pi'g,  g_k = F(x_k,u_k;theta) - x_{k+1},  h <= 0 (bounds),
//   R = [dL/dw ; g ; h + t ; lam*t - tau] = 0,  tau = 1e-8,
//   dpi/dtheta = -(first nu rows of) (dR/dz)^-1 dR/dtheta,  dV/dtheta = dL/dtheta.
// Equal bounds of stage 0 (x_0 = s always; u_0 = a in Q-mode) are imposed by elimination
// (the tau -> 0 limit of the reference, quirk Q7).
//
// One "lane" = one sample.  All per-sample vectors live in batch-minor (SoA) arrays: element
// i of sample b is at base[i*bs + b], so a warp touching element i reads 32 consecutive
// doubles (one 256-byte line pair).  The code is host/device generic: kernels.cu runs one
// lane per CUDA thread; the host build exists only for debugging and the CPU baseline.
#pragma once
#include "common.cuh"

namespace rlmpc {

template <class M>
struct Engine {
  static constexpr int NX = M::NX, NU = M::NU, NW = NX + NU, NPM = M::NPM;
  static constexpr int NPS = NX * (NX + 1) / 2;

  // ---------------- iterate layout (per sample, persistent) ----------------
  //   x[(N+1)NX] | u[N NU] | pi[N NX] | lam_u[N 2NU] | t_u[N 2NU]      (lower rows first)
  MPC_HD static int it_x(int N, int k) { (void)N; return k * NX; }
  MPC_HD static int it_u(int N, int k) { return (N + 1) * NX + k * NU; }
  MPC_HD static int it_pi(int N, int k) { return (N + 1) * NX + N * NU + k * NX; }
  MPC_HD static int it_lu(int N, int k) { return (N + 1) * NX + N * NU + N * NX + k * 2 * NU; }
  MPC_HD static int it_tu(int N, int k) { return (N + 1) * NX + N * NU + N * NX + N * 2 * NU + k * 2 * NU; }
  // multipliers of the eliminated equal bounds of stage 0 (x_0 = s; u_0 = a in Q-mode), i.e. the
  // gradient of the rest of the Lagrangian wrt x_0 / u_0; written by sens()
  MPC_HD static int it_rx0(int N) { return (N + 1) * NX + N * NU + N * NX + 4 * N * NU; }
  MPC_HD static int it_ru0(int N) { return it_rx0(N) + NX; }
  // meta[0] = 1.0 once (lam,t) hold the result of a QP solve (valid IPM warm start)
  MPC_HD static int it_meta(int N) { return it_ru0(N) + NU; }
  MPC_HD static int it_size(int N) { return it_meta(N) + 1; }

  // ---------------- workspace record per stage ----------------
  static constexpr int W_A = 0;
  static constexpr int W_B = W_A + NX * NX;
  static constexpr int W_b = W_B + NX * NU;
  static constexpr int W_q = W_b + NX;
  static constexpr int W_r = W_q + NX;
  static constexpr int W_K = W_r + NU;
  static constexpr int W_k = W_K + NU * NX;
  static constexpr int W_dx = W_k + NU;
  static constexpr int W_du = W_dx + NX;
  static constexpr int W_lh = W_du + NU;       // lam_hat (2NU)
  static constexpr int W_th = W_lh + 2 * NU;   // t_hat   (2NU)
  static constexpr int W_SOLVE_END = W_th + 2 * NU;
  // sensitivity pass re-uses the record: A, B, K stay where they are, the rest is overlaid
  static constexpr int W_P = W_K + NU * NX;            // P_{k+1} packed symmetric (NPS)
  static constexpr int W_Hwp = W_P + NPS;              // d(grad_w L)/d p_model  (NW x NPM)
  static constexpr int W_Fp = W_Hwp + NW * NPM;        // dF/d p_model (NX x NPM)
  static constexpr int W_Gi = W_Fp + NX * NPM;         // stage 0 only: inv(G_0) (NU x NU)
  static constexpr int W_SENS_END = W_Gi + NU * NU;
  static constexpr int W_REC = W_SOLVE_END > W_SENS_END ? W_SOLVE_END : W_SENS_END;
  MPC_HD static int ws_size(int N) { return (N + 1) * W_REC; }

  // ---------------- small helpers ----------------
  template <int n>
  MPC_HD static void ld(const double* p, size_t bs, double* out) {
    MPC_UNROLL for (int i = 0; i < n; ++i) out[i] = p[(size_t)i * bs];
  }
  template <int n>
  MPC_HD static void st(double* p, size_t bs, const double* in) {
    MPC_UNROLL for (int i = 0; i < n; ++i) p[(size_t)i * bs] = in[i];
  }

  // in-place Cholesky-based solve of a small SPD system G X = R (G: n x n, R: n x m); returns false if not PD
  template <int n, int m>
  MPC_HD static bool spd_solve(double* G, double* R) {
    if (n == 1) {
      if (!(G[0] > 0.0)) return false;
      const double inv = 1.0 / G[0];
      MPC_UNROLL for (int j = 0; j < m; ++j) R[j] *= inv;
      return true;
    }
    // LDL^T without pivoting
    bool ok = true;
    MPC_UNROLL for (int j = 0; j < n; ++j) {
      double d = G[j * n + j];
      MPC_UNROLL for (int p = 0; p < j; ++p) d -= G[j * n + p] * G[j * n + p] * G[p * n + p];
      if (!(d > 0.0)) ok = false;
      G[j * n + j] = d;
      const double inv = 1.0 / d;
      MPC_UNROLL for (int i = j + 1; i < n; ++i) {
        double v = G[i * n + j];
        MPC_UNROLL for (int p = 0; p < j; ++p) v -= G[i * n + p] * G[j * n + p] * G[p * n + p];
        G[i * n + j] = v * inv;
      }
    }
    MPC_UNROLL for (int c = 0; c < m; ++c) {
      MPC_UNROLL for (int i = 0; i < n; ++i) {
        double v = R[i * m + c];
        MPC_UNROLL for (int p = 0; p < i; ++p) v -= G[i * n + p] * R[p * m + c];
        R[i * m + c] = v;
      }
      MPC_UNROLL for (int i = 0; i < n; ++i) R[i * m + c] /= G[i * n + i];
      MPC_UNROLL for (int i = n - 1; i >= 0; --i) {
        double v = R[i * m + c];
        MPC_UNROLL for (int p = i + 1; p < n; ++p) v -= G[p * n + i] * R[p * m + c];
        R[i * m + c] = v;
      }
    }
    return ok;
  }

  // Cost weight blocks of stage kind (0 initial / 1 intermediate / 2 terminal), scaled by s
  MPC_HD static void load_W(int kind, double s, const Lane& L, double* Wm /* NW x NW (or NX x NX top-left) */) {
    const int n = (kind == 2) ? NX : NW;
    MPC_UNROLL for (int i = 0; i < NW; ++i) MPC_UNROLL for (int j = 0; j < NW; ++j) {
      Wm[i * NW + j] = (i < n && j < n) ? s * M::W(kind, i, j, L.th, L.ths) : 0.0;
    }
  }

  // One backward Riccati step.  In: P (NX x NX full), p; stage data.  Out: P, p (overwritten), K, kff.
  // Returns false if the reduced Hessian block G is not positive definite.
  MPC_HD static bool riccati_step(double* P, double* p, const double* A, const double* B, const double* b,
                                  const double* Hm /* NW x NW: [Q S'; S R] incl. barrier */,
                                  const double* gq /* NX */, const double* gr /* NU */, double* K, double* kff,
                                  double* Ginv /* optional NU x NU or nullptr */) {
    double v[NX], PA[NX * NX], PB[NX * NU];
    MPC_UNROLL for (int i = 0; i < NX; ++i) {
      double a = p[i];
      MPC_UNROLL for (int j = 0; j < NX; ++j) a += P[i * NX + j] * b[j];
      v[i] = a;
    }
    MPC_UNROLL for (int i = 0; i < NX; ++i) {
      MPC_UNROLL for (int j = 0; j < NX; ++j) {
        double a = 0.0;
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += P[i * NX + l] * A[l * NX + j];
        PA[i * NX + j] = a;
      }
      MPC_UNROLL for (int j = 0; j < NU; ++j) {
        double a = 0.0;
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += P[i * NX + l] * B[l * NU + j];
        PB[i * NU + j] = a;
      }
    }
    double G[NU * NU], H[NU * NX], gv[NU];
    MPC_UNROLL for (int i = 0; i < NU; ++i) {
      MPC_UNROLL for (int j = 0; j < NU; ++j) {
        double a = Hm[(NX + i) * NW + NX + j];
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * PB[l * NU + j];
        G[i * NU + j] = a;
      }
      MPC_UNROLL for (int j = 0; j < NX; ++j) {
        double a = Hm[(NX + i) * NW + j];
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * PA[l * NX + j];
        H[i * NX + j] = a;
      }
      double a = gr[i];
      MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * v[l];
      gv[i] = a;
    }
    // [K | kff | Ginv] = -G^{-1} [H | gv | -I]
    double R[NU * (NX + 1 + NU)];
    MPC_UNROLL for (int i = 0; i < NU; ++i) {
      MPC_UNROLL for (int j = 0; j < NX; ++j) R[i * (NX + 1 + NU) + j] = -H[i * NX + j];
      R[i * (NX + 1 + NU) + NX] = -gv[i];
      MPC_UNROLL for (int j = 0; j < NU; ++j) R[i * (NX + 1 + NU) + NX + 1 + j] = (i == j) ? 1.0 : 0.0;
    }
    const bool ok = spd_solve<NU, NX + 1 + NU>(G, R);
    MPC_UNROLL for (int i = 0; i < NU; ++i) {
      MPC_UNROLL for (int j = 0; j < NX; ++j) K[i * NX + j] = R[i * (NX + 1 + NU) + j];
      kff[i] = R[i * (NX + 1 + NU) + NX];
      if (Ginv) MPC_UNROLL for (int j = 0; j < NU; ++j) Ginv[i * NU + j] = R[i * (NX + 1 + NU) + NX + 1 + j];
    }
    // P <- Q + A'PA + H'K ;  p <- q + A'v + H'kff
    double Pn[NX * NX], pn[NX];
    MPC_UNROLL for (int i = 0; i < NX; ++i) {
      MPC_UNROLL for (int j = i; j < NX; ++j) {
        double a = Hm[i * NW + j];
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[l * NX + i] * PA[l * NX + j];
        MPC_UNROLL for (int l = 0; l < NU; ++l) a += H[l * NX + i] * K[l * NX + j];
        Pn[i * NX + j] = a;
        Pn[j * NX + i] = a;
      }
      double a = gq[i];
      MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[l * NX + i] * v[l];
      MPC_UNROLL for (int l = 0; l < NU; ++l) a += H[l * NX + i] * kff[l];
      pn[i] = a;
    }
    MPC_UNROLL for (int i = 0; i < NX * NX; ++i) P[i] = Pn[i];
    MPC_UNROLL for (int i = 0; i < NX; ++i) p[i] = pn[i];
    return ok;
  }

  struct Residuals {
    double stat, eq, ineq, comp, cost;
  };

  // ---------------------------------------------------------------------------------------
  // Linearise all stages at the current iterate: store A,B,b,q,r per stage, return the cost
  // and the four KKT residual norms (acados' convergence test).
  // ---------------------------------------------------------------------------------------
  MPC_HD static Residuals linearize(const ProblemData& pd, const Lane& L) {
    const int N = pd.N;
    const size_t bs = L.bs;
    Residuals R = {0, 0, 0, 0, 0};
    double pim[NX];  // pi_{k-1}
    MPC_UNROLL for (int i = 0; i < NX; ++i) pim[i] = 0.0;
    for (int k = 0; k < N; ++k) {
      double x[NX], u[NU], xn[NX], xnext[NX], A[NX * NX], B[NX * NU], bb[NX];
      ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      ld<NX>(L.it + (size_t)it_x(N, k + 1) * bs, bs, xnext);
      M::dyn_lin(x, u, L.th, L.ths, pd.mc, xn, A, B);
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        bb[i] = xn[i] - xnext[i];
        R.eq = dmax(R.eq, dabs(bb[i]));
      }
      double* w = L.ws + (size_t)k * W_REC * bs;
      st<NX * NX>(w + (size_t)W_A * bs, bs, A);
      st<NX * NU>(w + (size_t)W_B * bs, bs, B);
      st<NX>(w + (size_t)W_b * bs, bs, bb);
      // cost gradient
      const int kind = (k == 0) ? 0 : 1;
      const double s = pd.scale[k];
      double e[NW], gr[NW];
      MPC_UNROLL for (int i = 0; i < NW; ++i) e[i] = ((i < NX) ? x[i] : u[i - NX]) - M::yref(kind, i, L.th, L.ths);
      double c = M::c0(kind, L.th, L.ths);
      MPC_UNROLL for (int i = 0; i < NW; ++i) {
        double a = 0.0;
        MPC_UNROLL for (int j = 0; j < NW; ++j) a += M::W(kind, i, j, L.th, L.ths) * e[j];
        const double fl = M::flin(kind, i, L.th, L.ths);
        gr[i] = s * (a + fl);
        c += 0.5 * a * e[i] + fl * ((i < NX) ? x[i] : u[i - NX]);
      }
      R.cost += s * c;
      st<NX>(w + (size_t)W_q * bs, bs, gr);
      st<NU>(w + (size_t)W_r * bs, bs, gr + NX);
      // stationarity / complementarity residuals with the current multipliers
      double pik[NX], lu[2 * NU], tu[2 * NU];
      ld<NX>(L.it + (size_t)it_pi(N, k) * bs, bs, pik);
      ld<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
      ld<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
      const bool ufixed = (k == 0 && pd.mode == MODE_Q);
      if (!ufixed) {
        MPC_UNROLL for (int i = 0; i < NU; ++i) {
          double a = gr[NX + i] - lu[i] + lu[NU + i];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * pik[l];
          R.stat = dmax(R.stat, dabs(a));
          R.comp = dmax(R.comp, dmax(dabs(lu[i] * tu[i] - pd.tau), dabs(lu[NU + i] * tu[NU + i] - pd.tau)));
          R.ineq = dmax(R.ineq, dmax(dabs(pd.lbu[i] - u[i] + tu[i]), dabs(u[i] - pd.ubu[i] + tu[NU + i])));
        }
      }
      if (k > 0) {
        MPC_UNROLL for (int i = 0; i < NX; ++i) {
          double a = gr[i] - pim[i];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[l * NX + i] * pik[l];
          R.stat = dmax(R.stat, dabs(a));
        }
      }
      MPC_UNROLL for (int i = 0; i < NX; ++i) pim[i] = pik[i];
    }
    {  // terminal stage
      double x[NX], gr[NX];
      ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, x);
      const double s = pd.scale[N];
      double e[NX];
      MPC_UNROLL for (int i = 0; i < NX; ++i) e[i] = x[i] - M::yref(2, i, L.th, L.ths);
      double c = M::c0(2, L.th, L.ths);
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        double a = 0.0;
        MPC_UNROLL for (int j = 0; j < NX; ++j) a += M::W(2, i, j, L.th, L.ths) * e[j];
        const double fl = M::flin(2, i, L.th, L.ths);
        gr[i] = s * (a + fl);
        c += 0.5 * a * e[i] + fl * x[i];
        R.stat = dmax(R.stat, dabs(gr[i] - pim[i]));
      }
      R.cost += s * c;
      st<NX>(L.ws + ((size_t)N * W_REC + W_q) * bs, bs, gr);
    }
    return R;
  }

  // ---------------------------------------------------------------------------------------
  // Interior-point solve of the stage QP (Riccati factorisation per iteration).
  // State of the method is (lam,t) only (absolute-step form); the primal step dx,du of the last
  // iteration is left in the workspace.  Returns the number of IPM iterations, <0 on failure.
  // ---------------------------------------------------------------------------------------
  static constexpr int WARM_LIMIT = 6;  // IPM iterations granted to a warm start before a cold restart

  // returns sum(lam*t)
  MPC_HD static double ipm_init(const ProblemData& pd, const Lane& L, int k_first, bool warm) {
    const int N = pd.N;
    const size_t bs = L.bs;
    double mu = 0.0;
    for (int k = k_first; k < N; ++k) {
      double u[NU], lu[2 * NU], tu[2 * NU];
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      if (warm) {
        ld<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
        ld<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
      }
      MPC_UNROLL for (int i = 0; i < NU; ++i) {
        const double range = pd.ubu[i] - pd.lbu[i];
        if (warm) {
          tu[i] = dmax(tu[i], 1e-10 * range);
          tu[NU + i] = dmax(tu[NU + i], 1e-10 * range);
          lu[i] = dmax(lu[i], 1e-14);
          lu[NU + i] = dmax(lu[NU + i], 1e-14);
        } else {
          const double tmin = 1e-2 * range;
          tu[i] = dmax(u[i] - pd.lbu[i], tmin);
          tu[NU + i] = dmax(pd.ubu[i] - u[i], tmin);
          lu[i] = pd.mu0 / tu[i];
          lu[NU + i] = pd.mu0 / tu[NU + i];
        }
        mu += lu[i] * tu[i] + lu[NU + i] * tu[NU + i];
      }
      st<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
      st<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
    }
    return mu;
  }

  MPC_HD static int qp_ipm(const ProblemData& pd, const Lane& L, double* alpha_out) {
    const int N = pd.N;
    const size_t bs = L.bs;
    const bool qmode = pd.mode == MODE_Q;
    const int k_first = qmode ? 1 : 0;  // first stage with a free input
    const double m_rows = 2.0 * NU * (N - k_first);
    // ---- initialise (lam,t): warm = keep the multipliers of the previous QP (clipped away from
    // zero), cold = slacks from the current inputs, lam = mu0 / t ----
    bool warm = pd.warm_ipm && L.it[(size_t)it_meta(N) * bs] > 0.5;
    double mu = ipm_init(pd, L, k_first, warm) / m_rows;

    double alpha = 0.0;  // pending step length of the previous iteration (0: nothing pending)
    double sigma = warm ? 0.05 : 0.3;
    int iters = 0, warm_iters = 0;
    bool converged = false, failed = false;
    for (int j = 0; j < pd.max_ipm && !converged; ++j) {
      ++iters;
      if (warm && (warm_iters >= WARM_LIMIT || (warm_iters > 0 && alpha < 0.05))) {
        // the warm start is jammed (active set changed too much): restart from a cold point
        warm = false;
        mu = ipm_init(pd, L, k_first, false) / m_rows;
        alpha = 0.0;
        sigma = 0.3;
      }
      if (warm) ++warm_iters;
      const double target = dmax(sigma * mu, pd.tau);
      // ---------------- backward sweep ----------------
      double P[NX * NX], p[NX];
      {
        double Wm[NW * NW];
        load_W(2, pd.scale[N], L, Wm);
        MPC_UNROLL for (int i = 0; i < NX; ++i) MPC_UNROLL for (int jj = 0; jj < NX; ++jj) P[i * NX + jj] = Wm[i * NW + jj];
        ld<NX>(L.ws + ((size_t)N * W_REC + W_q) * bs, bs, p);
      }
      for (int k = N - 1; k >= 0; --k) {
        double* w = L.ws + (size_t)k * W_REC * bs;
        double A[NX * NX], B[NX * NU], bb[NX], gq[NX], gr[NU], K[NU * NX], kff[NU];
        ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
        ld<NX * NU>(w + (size_t)W_B * bs, bs, B);
        ld<NX>(w + (size_t)W_b * bs, bs, bb);
        ld<NX>(w + (size_t)W_q * bs, bs, gq);
        ld<NU>(w + (size_t)W_r * bs, bs, gr);
        double Hm[NW * NW];
        load_W(k == 0 ? 0 : 1, pd.scale[k], L, Hm);
        const bool ufixed = (k == 0 && qmode);
        if (!ufixed) {
          double u[NU], lu[2 * NU], tu[2 * NU];
          ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
          ld<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
          ld<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
          if (alpha > 0.0) {  // apply the pending damped update lam += alpha (lam_hat - lam)
            double lh[2 * NU], th[2 * NU];
            ld<2 * NU>(w + (size_t)W_lh * bs, bs, lh);
            ld<2 * NU>(w + (size_t)W_th * bs, bs, th);
            MPC_UNROLL for (int i = 0; i < 2 * NU; ++i) {
              lu[i] += alpha * (lh[i] - lu[i]);
              tu[i] += alpha * (th[i] - tu[i]);
            }
            st<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
            st<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
          }
          MPC_UNROLL for (int i = 0; i < NU; ++i) {
            const double itl = 1.0 / tu[i], itu = 1.0 / tu[NU + i];
            const double cl = lu[i] * itl, cu = lu[NU + i] * itu;
            Hm[(NX + i) * NW + NX + i] += cl + cu;
            // J'(target/t + lam + C*hbar), hbar_l = lb - u, hbar_u = u - ub
            gr[i] += -(target * itl + lu[i] - cl * (u[i] - pd.lbu[i])) + (target * itu + lu[NU + i] - cu * (pd.ubu[i] - u[i]));
          }
          if (!riccati_step(P, p, A, B, bb, Hm, gq, gr, K, kff, nullptr)) failed = true;
        } else {
          // u_0 fixed (Q-mode): no feedback; x_0 is fixed as well so P_0, p_0 are not needed
          MPC_UNROLL for (int i = 0; i < NU * NX; ++i) K[i] = 0.0;
          MPC_UNROLL for (int i = 0; i < NU; ++i) kff[i] = 0.0;
        }
        st<NU * NX>(w + (size_t)W_K * bs, bs, K);
        st<NU>(w + (size_t)W_k * bs, bs, kff);
      }
      // ---------------- forward sweep ----------------
      double dx[NX];
      MPC_UNROLL for (int i = 0; i < NX; ++i) dx[i] = 0.0;
      double amax = 1e300, s0 = 0.0, s1 = 0.0, s2 = 0.0, cmax = 0.0;
      for (int k = 0; k < N; ++k) {
        double* w = L.ws + (size_t)k * W_REC * bs;
        double A[NX * NX], B[NX * NU], bb[NX], K[NU * NX], du[NU];
        ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
        ld<NX * NU>(w + (size_t)W_B * bs, bs, B);
        ld<NX>(w + (size_t)W_b * bs, bs, bb);
        ld<NU * NX>(w + (size_t)W_K * bs, bs, K);
        ld<NU>(w + (size_t)W_k * bs, bs, du);
        MPC_UNROLL for (int i = 0; i < NU; ++i) MPC_UNROLL for (int l = 0; l < NX; ++l) du[i] += K[i * NX + l] * dx[l];
        st<NX>(w + (size_t)W_dx * bs, bs, dx);
        st<NU>(w + (size_t)W_du * bs, bs, du);
        if (!(k == 0 && qmode)) {
          double u[NU], lu[2 * NU], tu[2 * NU], lh[2 * NU], th[2 * NU];
          ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
          ld<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
          ld<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
          MPC_UNROLL for (int i = 0; i < NU; ++i) {
            // (u - lb) + du, NOT (u + du) - lb: the slack must use the same rounded distance to the
            // bound as the condensed gradient above, or lam_hat picks up c*ulp(u) ~ 1e-8 of noise
            th[i] = (u[i] - pd.lbu[i]) + du[i];
            th[NU + i] = (pd.ubu[i] - u[i]) - du[i];
          }
          MPC_UNROLL for (int i = 0; i < 2 * NU; ++i) {
            const double it_ = 1.0 / tu[i];
            lh[i] = target * it_ + lu[i] - lu[i] * it_ * th[i];
            const double dt = th[i] - tu[i], dl = lh[i] - lu[i];
            if (dt < 0.0) amax = dmin(amax, -tu[i] / dt);
            if (dl < 0.0) amax = dmin(amax, -lu[i] / dl);
            s0 += lu[i] * tu[i];
            s1 += lu[i] * dt + tu[i] * dl;
            s2 += dl * dt;
            cmax = dmax(cmax, dabs(dl * dt));
          }
          st<2 * NU>(w + (size_t)W_lh * bs, bs, lh);
          st<2 * NU>(w + (size_t)W_th * bs, bs, th);
        }
        double dxn[NX];
        MPC_UNROLL for (int i = 0; i < NX; ++i) {
          double a = bb[i];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[i * NX + l] * dx[l];
          MPC_UNROLL for (int l = 0; l < NU; ++l) a += B[i * NU + l] * du[l];
          dxn[i] = a;
        }
        MPC_UNROLL for (int i = 0; i < NX; ++i) dx[i] = dxn[i];
      }
      st<NX>(L.ws + ((size_t)N * W_REC + W_dx) * bs, bs, dx);
      if (failed || !(amax == amax)) break;
      alpha = (amax >= 1.0 / 0.995) ? 1.0 : 0.995 * amax;
      const double mu_new = (s0 + alpha * s1 + alpha * alpha * s2) / m_rows;
      if (target <= pd.tau && alpha == 1.0 && cmax <= dmin(0.05 * pd.tau, 0.1 * pd.tol)) converged = true;
      // centring heuristic: aggressive after long steps, conservative after short ones
      const double r = 1.0 - alpha;
      sigma = dmin(0.8, dmax(0.05, r * r * 4.0 + 0.05));
      mu = mu_new;
    }
    *alpha_out = alpha;
    if (failed) return -1;
    return converged ? iters : -(iters + 1000);
  }

  // ---------------------------------------------------------------------------------------
  // Apply the QP step: w += dw, (lam,t) <- last IPM update, pi <- QP multipliers (backward
  // recursion of the x-stationarity rows).
  // ---------------------------------------------------------------------------------------
  MPC_HD static void apply_step(const ProblemData& pd, const Lane& L, double alpha) {
    const int N = pd.N;
    const size_t bs = L.bs;
    const bool qmode = pd.mode == MODE_Q;
    double pik[NX];  // pi_k (multiplier of x_{k+1} = F(x_k,u_k))
    {
      double dx[NX], q[NX], x[NX], Wm[NW * NW];
      ld<NX>(L.ws + ((size_t)N * W_REC + W_dx) * bs, bs, dx);
      ld<NX>(L.ws + ((size_t)N * W_REC + W_q) * bs, bs, q);
      load_W(2, pd.scale[N], L, Wm);
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        double a = q[i];
        MPC_UNROLL for (int j = 0; j < NX; ++j) a += Wm[i * NW + j] * dx[j];
        pik[i] = a;
      }
      ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, x);
      MPC_UNROLL for (int i = 0; i < NX; ++i) x[i] += dx[i];
      st<NX>(L.it + (size_t)it_x(N, N) * bs, bs, x);
    }
    for (int k = N - 1; k >= 0; --k) {
      double* w = L.ws + (size_t)k * W_REC * bs;
      st<NX>(L.it + (size_t)it_pi(N, k) * bs, bs, pik);
      double dx[NX], du[NU], x[NX], u[NU];
      ld<NX>(w + (size_t)W_dx * bs, bs, dx);
      ld<NU>(w + (size_t)W_du * bs, bs, du);
      if (!(k == 0 && qmode)) {
        double lu[2 * NU], tu[2 * NU], lh[2 * NU], th[2 * NU];
        ld<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
        ld<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
        ld<2 * NU>(w + (size_t)W_lh * bs, bs, lh);
        ld<2 * NU>(w + (size_t)W_th * bs, bs, th);
        MPC_UNROLL for (int i = 0; i < 2 * NU; ++i) {
          lu[i] += alpha * (lh[i] - lu[i]);
          tu[i] += alpha * (th[i] - tu[i]);
        }
        st<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
        st<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
      }
      if (k > 0) {
        double A[NX * NX], q[NX], Wm[NW * NW], pin[NX];
        ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
        ld<NX>(w + (size_t)W_q * bs, bs, q);
        load_W(1, pd.scale[k], L, Wm);
        MPC_UNROLL for (int i = 0; i < NX; ++i) {
          double a = q[i];
          MPC_UNROLL for (int j = 0; j < NX; ++j) a += Wm[i * NW + j] * dx[j];
          MPC_UNROLL for (int j = 0; j < NU; ++j) a += Wm[i * NW + NX + j] * du[j];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[l * NX + i] * pik[l];
          pin[i] = a;
        }
        MPC_UNROLL for (int i = 0; i < NX; ++i) pik[i] = pin[i];
        ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
        MPC_UNROLL for (int i = 0; i < NX; ++i) x[i] += dx[i];
        st<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
      }
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      MPC_UNROLL for (int i = 0; i < NU; ++i) u[i] += du[i];
      st<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
    }
  }

  // ---------------------------------------------------------------------------------------
  // SQP driver (acados semantics: linearise -> convergence test -> QP -> full step).
  // Outputs the residuals/cost of the last linearisation; `fresh` tells whether they belong to
  // the final iterate (true when converged) or to the iterate before the last step (RTI).
  // ---------------------------------------------------------------------------------------
  struct SolveOut {
    Residuals res;
    int status;
    int sqp_iter;
    int ipm_iter;
    bool fresh;
  };

  MPC_HD static void set_initial(const ProblemData& pd, const Lane& L, const double* x0, size_t x0s, const double* u0,
                                 size_t u0s) {
    const int N = pd.N;
    MPC_UNROLL for (int i = 0; i < NX; ++i) L.it[(size_t)(it_x(N, 0) + i) * L.bs] = x0[(size_t)i * x0s];
    if (pd.mode == MODE_Q) {
      MPC_UNROLL for (int i = 0; i < NU; ++i) L.it[(size_t)(it_u(N, 0) + i) * L.bs] = u0[(size_t)i * u0s];
    }
  }

  MPC_HD static SolveOut solve(const ProblemData& pd, const Lane& L) {
    SolveOut o;
    o.status = ST_MAXITER;
    o.sqp_iter = 0;
    o.ipm_iter = 0;
    o.fresh = false;
    for (int itn = 0;; ++itn) {
      o.res = linearize(pd, L);
      const double rmax = dmax(dmax(o.res.stat, o.res.eq), dmax(o.res.ineq, o.res.comp));
      if (!(rmax == rmax) || !(o.res.cost == o.res.cost)) {
        o.status = ST_NAN;
        o.fresh = true;
        break;
      }
      if (rmax < pd.tol) {
        o.status = ST_OK;
        o.fresh = true;
        break;
      }
      if (itn >= pd.max_sqp) {
        o.fresh = true;
        break;
      }
      double alpha = 0.0;
      const int r = qp_ipm(pd, L, &alpha);
      o.ipm_iter += (r > 0) ? r : ((r == -1) ? 0 : -(r + 1000));
      if (r == -1) {
        o.status = ST_QPFAIL;
        break;
      }
      apply_step(pd, L, alpha);
      L.it[(size_t)it_meta(pd.N) * L.bs] = 1.0;
      o.sqp_iter = itn + 1;
      if (r < 0) {  // IPM hit its iteration limit
        o.status = ST_QPFAIL;
      }
      if (pd.max_sqp == 1) {  // RTI: one QP, no further linearisation here (sens() re-evaluates)
        if (o.status != ST_QPFAIL) o.status = ST_OK;
        break;
      }
    }
    return o;
  }

  // ---------------------------------------------------------------------------------------
  // Evaluation + sensitivities at the current iterate.
  //   * cost and KKT residuals (what update_nlp asserts, nlp.py:1445-1537)
  //   * dL/dtheta (model part; cost part when pd.param_cost)           nlp.py:1211-1212,1401
  //   * dpi/dtheta via ONE exact-Hessian Riccati factorisation and NU adjoint solves, instead of
  //     the reference's dense (nz x nz) sparse LU with ntheta right-hand sides  nlp.py:1413-1424
  // dLdth / dpidth point at this sample's rows of row-major [B, ng] / [B, NU, ng] outputs.
  // ---------------------------------------------------------------------------------------
  // Gradient rows have width ng = grad_width(pd): the model parameters only (the structurally
  // non-zero prefix of the reference's p, quirk Q8) or the whole p when pd.param_cost is set.
  MPC_HD static int grad_width(const ProblemData& pd) { return pd.param_cost ? M::NTH : NPM; }

  MPC_HD static Residuals sens(const ProblemData& pd, const Lane& L, double* dLdth, double* dpidth, int* ok_out) {
    const int N = pd.N;
    const size_t bs = L.bs;
    const bool qmode = pd.mode == MODE_Q;
    Residuals R = {0, 0, 0, 0, 0};
    bool ok = true;
    double gp[NPM];
    MPC_UNROLL for (int i = 0; i < NPM; ++i) gp[i] = 0.0;

    // ---- backward pass: exact Hessian blocks, Riccati factorisation ----
    double P[NX * NX], pdummy[NX];
    double xk1[NX];  // x_{k+1}
    double grn[NX];  // cost gradient at x_{k+1} (for the stationarity residual of stage k+1)
    double An_pik[NX];  // A_{k+1}' pi_{k+1}
    {
      double Wm[NW * NW], e[NX];
      ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, xk1);
      load_W(2, pd.scale[N], L, Wm);
      MPC_UNROLL for (int i = 0; i < NX; ++i) e[i] = xk1[i] - M::yref(2, i, L.th, L.ths);
      double c = pd.scale[N] * M::c0(2, L.th, L.ths);
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        double a = 0.0;
        MPC_UNROLL for (int j = 0; j < NX; ++j) a += Wm[i * NW + j] * e[j];
        const double fl = pd.scale[N] * M::flin(2, i, L.th, L.ths);
        grn[i] = a + fl;
        c += 0.5 * a * e[i] + fl * xk1[i];
        An_pik[i] = 0.0;
        pdummy[i] = 0.0;
        MPC_UNROLL for (int j = 0; j < NX; ++j) P[i * NX + j] = Wm[i * NW + j];
      }
      R.cost += c;
      if (pd.param_cost && dLdth) cost_param_grad(2, pd.scale[N], L, xk1, nullptr, dLdth);
    }
    for (int k = N - 1; k >= 0; --k) {
      double* w = L.ws + (size_t)k * W_REC * bs;
      double x[NX], u[NU], pik[NX], xn[NX], A[NX * NX], B[NX * NU], Fp[NX * NPM], Hww[NW * NW], Hwp[NW * NPM];
      ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      ld<NX>(L.it + (size_t)it_pi(N, k) * bs, bs, pik);
      M::dyn_sens(x, u, L.th, L.ths, pd.mc, pik, xn, A, B, Fp, Hww, Hwp);
      // stationarity residual wrt x_{k+1}:  grad l_{k+1} + A_{k+1}' pi_{k+1} - pi_k
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        R.eq = dmax(R.eq, dabs(xn[i] - xk1[i]));
        R.stat = dmax(R.stat, dabs(grn[i] + An_pik[i] - pik[i]));
      }
      // dL/dp_model += pi_k' dF/dp
      MPC_UNROLL for (int j = 0; j < NPM; ++j) MPC_UNROLL for (int i = 0; i < NX; ++i) gp[j] += pik[i] * Fp[i * NPM + j];
      // cost at stage k
      const int kind = (k == 0) ? 0 : 1;
      const double s = pd.scale[k];
      double Wm[NW * NW], e[NW], gr[NW];
      load_W(kind, s, L, Wm);
      MPC_UNROLL for (int i = 0; i < NW; ++i) e[i] = ((i < NX) ? x[i] : u[i - NX]) - M::yref(kind, i, L.th, L.ths);
      double c = s * M::c0(kind, L.th, L.ths);
      MPC_UNROLL for (int i = 0; i < NW; ++i) {
        double a = 0.0;
        MPC_UNROLL for (int j = 0; j < NW; ++j) a += Wm[i * NW + j] * e[j];
        const double fl = s * M::flin(kind, i, L.th, L.ths);
        gr[i] = a + fl;
        c += 0.5 * a * e[i] + fl * ((i < NX) ? x[i] : u[i - NX]);
      }
      R.cost += c;
      if (pd.param_cost && dLdth) cost_param_grad(kind, s, L, x, u, dLdth);
      // exact Lagrangian Hessian block + barrier terms
      double Hm[NW * NW];
      MPC_UNROLL for (int i = 0; i < NW * NW; ++i) Hm[i] = Wm[i] + Hww[i];
      const bool ufixed = (k == 0 && qmode);
      double lu[2 * NU], tu[2 * NU];
      ld<2 * NU>(L.it + (size_t)it_lu(N, k) * bs, bs, lu);
      ld<2 * NU>(L.it + (size_t)it_tu(N, k) * bs, bs, tu);
      if (!ufixed) {
        MPC_UNROLL for (int i = 0; i < NU; ++i) {
          Hm[(NX + i) * NW + NX + i] += lu[i] / tu[i] + lu[NU + i] / tu[NU + i];
          double a = gr[NX + i] - lu[i] + lu[NU + i];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * pik[l];
          R.stat = dmax(R.stat, dabs(a));
          R.comp = dmax(R.comp, dmax(dabs(lu[i] * tu[i] - pd.tau), dabs(lu[NU + i] * tu[NU + i] - pd.tau)));
          R.ineq = dmax(R.ineq, dmax(dabs(pd.lbu[i] - u[i] + tu[i]), dabs(u[i] - pd.ubu[i] + tu[NU + i])));
        }
      }
      // record for the forward (adjoint) pass
      if (dpidth) {
        st<NX * NX>(w + (size_t)W_A * bs, bs, A);
        st<NX * NU>(w + (size_t)W_B * bs, bs, B);
        double Pp[NPS];
        {
          int c_ = 0;
          MPC_UNROLL for (int i = 0; i < NX; ++i) MPC_UNROLL for (int j = i; j < NX; ++j) Pp[c_++] = P[i * NX + j];
        }
        st<NPS>(w + (size_t)W_P * bs, bs, Pp);
        st<NW * NPM>(w + (size_t)W_Hwp * bs, bs, Hwp);
        st<NX * NPM>(w + (size_t)W_Fp * bs, bs, Fp);
        double K[NU * NX], kff[NU], Ginv[NU * NU], zq[NX], zr[NU], zb[NX];
        MPC_UNROLL for (int i = 0; i < NX; ++i) { zq[i] = 0.0; zb[i] = 0.0; }
        MPC_UNROLL for (int i = 0; i < NU; ++i) zr[i] = 0.0;
        if (!ufixed) {
          if (!riccati_step(P, pdummy, A, B, zb, Hm, zq, zr, K, kff, Ginv)) ok = false;
        } else {
          MPC_UNROLL for (int i = 0; i < NU * NX; ++i) K[i] = 0.0;
          MPC_UNROLL for (int i = 0; i < NU * NU; ++i) Ginv[i] = 0.0;
        }
        st<NU * NX>(w + (size_t)W_K * bs, bs, K);
        if (k == 0) st<NU * NU>(w + (size_t)W_Gi * bs, bs, Ginv);
      }
      // carry to stage k-1
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        double a = 0.0;
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[l * NX + i] * pik[l];
        An_pik[i] = a;
        grn[i] = gr[i];
        xk1[i] = x[i];
      }
      if (k == 0) {  // multipliers of the eliminated stage-0 equalities
        MPC_UNROLL for (int i = 0; i < NX; ++i) L.it[(size_t)(it_rx0(N) + i) * bs] = gr[i] + An_pik[i];
        MPC_UNROLL for (int i = 0; i < NU; ++i) {
          double a = gr[NX + i];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * pik[l];
          L.it[(size_t)(it_ru0(N) + i) * bs] = a;
        }
      }
    }
    if (dLdth) {
      MPC_UNROLL for (int j = 0; j < NPM; ++j) dLdth[j] = gp[j];
    }
    // ---- forward pass: NU adjoint solves  K y_i = e_{u0,i},  dpi_i/dtheta = -y_i' dR/dtheta ----
    if (dpidth) {
      double acc[NU * NPM];
      MPC_UNROLL for (int i = 0; i < NU * NPM; ++i) acc[i] = 0.0;
      if (!qmode) {
        double yx[NU * NX];  // y_x of each right-hand side
        MPC_UNROLL for (int i = 0; i < NU * NX; ++i) yx[i] = 0.0;
        for (int k = 0; k < N; ++k) {
          const double* w = L.ws + (size_t)k * W_REC * bs;
          double A[NX * NX], B[NX * NU], K[NU * NX], Pp[NPS], Hwp[NW * NPM], Fp[NX * NPM];
          ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
          ld<NX * NU>(w + (size_t)W_B * bs, bs, B);
          ld<NU * NX>(w + (size_t)W_K * bs, bs, K);
          ld<NPS>(w + (size_t)W_P * bs, bs, Pp);
          ld<NW * NPM>(w + (size_t)W_Hwp * bs, bs, Hwp);
          ld<NX * NPM>(w + (size_t)W_Fp * bs, bs, Fp);
          double Pf[NX * NX];
          {
            int c_ = 0;
            MPC_UNROLL for (int i = 0; i < NX; ++i) MPC_UNROLL for (int j = i; j < NX; ++j) {
              Pf[i * NX + j] = Pp[c_];
              Pf[j * NX + i] = Pp[c_++];
            }
          }
          double Gi[NU * NU];
          if (k == 0) ld<NU * NU>(w + (size_t)W_Gi * bs, bs, Gi);
          MPC_UNROLL for (int r = 0; r < NU; ++r) {
            double yu[NU], yxn[NX], ypi[NX];
            MPC_UNROLL for (int i = 0; i < NU; ++i) {
              double a = (k == 0) ? Gi[i * NU + r] : 0.0;
              MPC_UNROLL for (int l = 0; l < NX; ++l) a += K[i * NX + l] * yx[r * NX + l];
              yu[i] = a;
            }
            MPC_UNROLL for (int i = 0; i < NX; ++i) {
              double a = 0.0;
              MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[i * NX + l] * yx[r * NX + l];
              MPC_UNROLL for (int l = 0; l < NU; ++l) a += B[i * NU + l] * yu[l];
              yxn[i] = a;
            }
            MPC_UNROLL for (int i = 0; i < NX; ++i) {
              double a = 0.0;
              MPC_UNROLL for (int l = 0; l < NX; ++l) a += Pf[i * NX + l] * yxn[l];
              ypi[i] = a;
            }
            MPC_UNROLL for (int j = 0; j < NPM; ++j) {
              double a = 0.0;
              MPC_UNROLL for (int i = 0; i < NX; ++i) a += yx[r * NX + i] * Hwp[i * NPM + j] + ypi[i] * Fp[i * NPM + j];
              MPC_UNROLL for (int i = 0; i < NU; ++i) a += yu[i] * Hwp[(NX + i) * NPM + j];
              acc[r * NPM + j] -= a;
            }
            MPC_UNROLL for (int i = 0; i < NX; ++i) yx[r * NX + i] = yxn[i];
          }
        }
      }
      MPC_UNROLL for (int r = 0; r < NU; ++r) MPC_UNROLL for (int j = 0; j < NPM; ++j) dpidth[(size_t)r * grad_width(pd) + j] = acc[r * NPM + j];
    }
    *ok_out = ok ? 1 : 0;
    return R;
  }

  // d(s * l)/d(W, yref) accumulated into the [NTH] row (parameterize_tracking_cost=True semantics,
  // nlp.py:1057-1074): dl/dW_ij = 1/2 e_i e_j, dl/dyref = -W_sym e.
  MPC_HD static void cost_param_grad(int kind, double s, const Lane& L, const double* x, const double* u, double* dLdth) {
    const int n = M::ny(kind);
    double e[NW];
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      e[i] = 0.0;
      if (i < n) e[i] = ((i < NX) ? x[i] : u[i - NX]) - M::yref(kind, i, L.th, L.ths);
    }
    const int wo = M::w_off(kind), yo = M::yref_off(kind);
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      if (i >= n) continue;
      double a = 0.0;
      MPC_UNROLL for (int j = 0; j < NW; ++j) {
        if (j >= n) continue;
        a += M::W(kind, i, j, L.th, L.ths) * e[j];
        dLdth[wo + j * n + i] += 0.5 * s * e[i] * e[j];
      }
      dLdth[yo + i] -= s * a;
    }
  }
};

}  // namespace rlmpc
