// rlmpc-b200 engine: per-sample SQP (Gauss-Newton Hessian) with a Riccati-structured primal-dual
// interior-point QP solve, followed by the exact-Hessian adjoint KKT solve that yields
// dpi/dtheta and the stage sweep that yields dL/dtheta (= dV/dtheta = dQ/dtheta).
//
// What this replaces in the reference:
//   a4  AcadosOcpSolver.solve()  (acados SQP + HPIPM Riccati IPM)   -> lin_stage + qp_fast/qp_full
//   a5  update_nlp()  rlmpc/mpc/nlp.py:1341-1563 (dense dR/dz + SuperLU, dL/dp) -> sens_stage + sens_sweep
// The maths is specified by rlmpc/mpc/nlp.py:884-1275:
//   L = cost + lam'h + pi'g,  g_k = F(x_k,u_k;theta) - x_{k+1},  h <= 0 (bounds),
//   R = [dL/dw ; g ; h + t ; lam*t - tau] = 0,  tau = 1e-8,
//   dpi/dtheta = -(first nu rows of) (dR/dz)^-1 dR/dtheta,  dV/dtheta = dL/dtheta.
// Equal bounds of stage 0 (x_0 = s always; u_0 = a in Q-mode) are imposed by elimination
// (the tau -> 0 limit of the reference, quirk Q7).
//
// Work decomposition.  The functions below come in two shapes:
//   *_stage(pd, L, k)  one (sample, stage) pair: function/derivative evaluation, no dependence
//                      between stages -> launched over B x (N+1) threads, compute bound;
//   qp_* / sens_sweep  one sample: the Riccati recursions, sequential in the stage index ->
//                      launched over B threads, streaming the per-stage records from HBM.
// All per-sample vectors live in tiled batch-minor (AoSoA) arrays, see TILE in common.cuh: a warp
// touching element i reads 32 consecutive doubles (256 bytes) at an immediate offset.
// The code is host/device generic: rlmpc_b200.cu runs it in CUDA kernels; a host build of the
// same templates exists only as test infrastructure (debugging, CPU baseline).
#pragma once
#include "common.cuh"

namespace rlmpc {

template <class M>
struct Engine {
  static constexpr int NX = M::NX, NU = M::NU, NW = NX + NU, NPM = M::NPM, NBX = M::NBX;
  static constexpr int NG = M::NG;      // general linear rows lg <= g0 + C [x;u] <= ug (constraints.lh/uh with an affine h)
  static constexpr int NV = NU + NBX + NG;  // constrained quantities of a stage: v = [u ; x[bx] ; g]
  static constexpr int NSX = M::NSX;    // soft state bounds (subset of bx, stages 1..N-1): slack rows
  static constexpr int NSXA = NSX > 0 ? NSX : 1;
  // inequality rows per stage, acados order [lbu lbx lh ubu ubx uh lsbx usbx] (rlmpc/common/utils.py:4-25)
  static constexpr int NR = 2 * NV + 2 * NSX;
  static constexpr int R_LS = 2 * NV, R_US = 2 * NV + NSX;  // first lower / upper slack row
  static constexpr bool NEEDX = NBX > 0 || NG > 0;  // do the rows of a stage depend on x?
  static constexpr int NPS = NX * (NX + 1) / 2;
  static constexpr int NWS = NW * (NW + 1) / 2;

  // ---------------- iterate layout (per sample, persistent) ----------------
  //   x[(N+1)NX] | u[N NU] | pi[N NX] | lam[(N+1) NR] | t[(N+1) NR] | rho_x0 | rho_u0 | meta
  MPC_HD static int it_x(int N, int k) { (void)N; return k * NX; }
  MPC_HD static int it_u(int N, int k) { return (N + 1) * NX + k * NU; }
  MPC_HD static int it_pi(int N, int k) { return (N + 1) * NX + N * NU + k * NX; }
  MPC_HD static int it_lam(int N, int k) { return (N + 1) * NX + N * NU + N * NX + k * NR; }
  MPC_HD static int it_t(int N, int k) { return it_lam(N, 0) + (N + 1) * NR + k * NR; }
  // multipliers of the eliminated equal bounds of stage 0 (x_0 = s; u_0 = a in Q-mode), i.e. the
  // gradient of the rest of the Lagrangian wrt x_0 / u_0; written by sens_sweep()
  MPC_HD static int it_rx0(int N) { return it_t(N, 0) + (N + 1) * NR; }
  MPC_HD static int it_ru0(int N) { return it_rx0(N) + NX; }
  // meta[0] = 1.0 once (lam,t) hold the result of a QP solve (valid IPM warm start)
  MPC_HD static int it_meta(int N) { return it_ru0(N) + NU; }
  MPC_HD static int it_size(int N) { return it_meta(N) + 1; }

  // ---------------- quadratic cost table (derived from theta by M::cost_table) ----------------
  //   l_kind(y) = c0 + flin'y + 1/2 (y - yref)' W (y - yref),  y = [x;u]  (terminal: u part zero)
  static constexpr int CT_W = 0;                 // NWS, packed upper triangle, row-major
  static constexpr int CT_Y = CT_W + NWS;        // yref (NW)
  static constexpr int CT_F = CT_Y + NW;         // flin (NW)
  static constexpr int CT_C = CT_F + NW;         // c0
  static constexpr int CT_REC = CT_C + 1;
  static constexpr int CT_SIZE = 3 * CT_REC;     // kinds: 0 initial, 1 intermediate, 2 terminal
  MPC_HD static constexpr int pidx(int i, int j) { return i * NW - i * (i - 1) / 2 + (j - i); }  // i <= j

  // ---------------- workspace record per stage ----------------
  // solve phase
  static constexpr int W_A = 0;
  static constexpr int W_B = W_A + NX * NX;
  static constexpr int W_b = W_B + NX * NU;
  static constexpr int W_q = W_b + NX;          // scaled cost gradient wrt x
  static constexpr int W_r = W_q + NX;          // ... wrt u
  static constexpr int W_H = W_r + NU;          // only with M::STAGE_HESS: stage Hessian [Q S'; S R], packed upper (NWS)
  static constexpr int W_K = W_H + (M::STAGE_HESS ? NWS : 0);
  static constexpr int W_k = W_K + NU * NX;
  static constexpr int W_lh = W_k + NU;         // lam_hat (NR)
  static constexpr int W_th = W_lh + NR;        // t_hat   (NR)
  static constexpr int W_dx = W_th + NR;        // primal step (written by the forward sweep, read by the step)
  static constexpr int W_du = W_dx + NX;
  static constexpr int W_c = W_du + NU;         // scaled stage cost
  static constexpr int W_e = W_c + 1;           // |F(x_k,u_k) - x_{k+1}|_inf
  static constexpr int W_SOLVE_END = W_e + 1;
  // sensitivity phase (overlays the solve record; A, B stay where they are)
  static constexpr int S_g = W_b;                      // scaled cost gradient [x;u] (NW)
  static constexpr int S_H = S_g + NW;                 // Hessian of pi'F wrt w, packed upper (NWS)
  static constexpr int S_Hwp = S_H + NWS;              // d(grad_w pi'F)/d p_model  (NW x NPM)
  static constexpr int S_Fp = S_Hwp + NW * NPM;        // dF/d p_model (NX x NPM)
  static constexpr int S_gp = S_Fp + NX * NPM;         // pi_k' dF/dp (NPM)
  static constexpr int S_c = S_gp + NPM;               // scaled stage cost
  static constexpr int S_e = S_c + 1;                  // dynamics defect norm
  static constexpr int S_K = S_e + 1;                  // feedback gain of the exact-Hessian factorisation
  static constexpr int S_P = S_K + NU * NX;            // P_{k+1} packed symmetric (NPS)
  static constexpr int S_Gi = S_P + NPS;               // stage 0 only: inv(G_0) (NU x NU)
  static constexpr int W_SENS_END = S_Gi + NU * NU;
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

  // in-place LDL'-based solve of a small symmetric system G X = R (G: n x n, R: n x m).  PD = true: returns false if
  // G is not positive definite (the QP solves: an indefinite reduced Hessian is a failed QP).  PD = false: pivots of
  // either sign are accepted, false only for a zero / NaN pivot (the sensitivities: the reference solves the KKT
  // system of update_nlp with a general sparse LU, nlp.py:1413-1424, which needs it nonsingular, not definite --
  // at an unconverged RTI iterate the exact Hessian of the Lagrangian often is indefinite).
  template <int n, int m, bool PD = true>
  MPC_HD static bool spd_solve(double* G, double* R) {
    if (n == 1) {
      if (PD ? !(G[0] > 0.0) : !(G[0] > 0.0 || G[0] < 0.0)) return false;
      const double inv = 1.0 / G[0];
      MPC_UNROLL for (int j = 0; j < m; ++j) R[j] *= inv;
      return true;
    }
    bool ok = true;
    double dinv[n];  // reciprocal pivots: one division per pivot, the n x m diagonal scalings below multiply
    MPC_UNROLL for (int j = 0; j < n; ++j) {
      double d = G[j * n + j];
      MPC_UNROLL for (int p = 0; p < j; ++p) d -= G[j * n + p] * G[j * n + p] * G[p * n + p];
      if (PD ? !(d > 0.0) : !(d > 0.0 || d < 0.0)) ok = false;
      G[j * n + j] = d;
      const double inv = 1.0 / d;
      dinv[j] = inv;
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
      MPC_UNROLL for (int i = 0; i < n; ++i) R[i * m + c] *= dinv[i];
      MPC_UNROLL for (int i = n - 1; i >= 0; --i) {
        double v = R[i * m + c];
        MPC_UNROLL for (int p = i + 1; p < n; ++p) v -= G[p * n + i] * R[p * m + c];
        R[i * m + c] = v;
      }
    }
    return ok;
  }

  // ---------------- quadratic stage cost ----------------
  struct CostK {
    double W[NW * NW];  // full symmetric, unscaled
    double yref[NW], flin[NW], c0;
  };
  MPC_HD static void load_cost(int kind, const Lane& L, CostK& c) {
    const double* p = L.ct + (size_t)(kind * CT_REC) * (size_t)TILE;
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      MPC_UNROLL for (int j = i; j < NW; ++j) {
        const double v = p[(size_t)(CT_W + pidx(i, j)) * (size_t)TILE];
        c.W[i * NW + j] = v;
        c.W[j * NW + i] = v;
      }
      c.yref[i] = p[(size_t)(CT_Y + i) * (size_t)TILE];
      c.flin[i] = p[(size_t)(CT_F + i) * (size_t)TILE];
    }
    c.c0 = p[(size_t)CT_C * (size_t)TILE];
  }
  // scaled Hessian block only (what the Riccati sweeps need)
  MPC_HD static void load_W(int kind, double s, const Lane& L, double* Wm) {
    const double* p = L.ct + (size_t)(kind * CT_REC + CT_W) * (size_t)TILE;
    MPC_UNROLL for (int i = 0; i < NW; ++i) MPC_UNROLL for (int j = i; j < NW; ++j) {
      const double v = s * p[(size_t)pidx(i, j) * (size_t)TILE];
      Wm[i * NW + j] = v;
      Wm[j * NW + i] = v;
    }
  }
  // Hessian of the stage QP: the scaled cost table, or (M::STAGE_HESS, e.g. condensed blocks) the record
  MPC_HD static void stage_hess(int kind, double s, const Lane& L, const double* wr, double* Hm) {
    if (M::STAGE_HESS) {
      MPC_UNROLL for (int i = 0; i < NW; ++i) MPC_UNROLL for (int j = i; j < NW; ++j) {
        const double v = wr[(size_t)(W_H + pidx(i, j)) * TILE];
        Hm[i * NW + j] = v;
        Hm[j * NW + i] = v;
      }
    } else {
      load_W(kind, s, L, Hm);
    }
  }
  // gradient g = s (W (y - yref) + flin) and value s l(y) of one stage, y = [x;u] (n = NW or NX)
  MPC_HD static double cost_grad(const CostK& c, double s, int n, const double* y, double* g) {
    double e[NW], val = c.c0;
    MPC_UNROLL for (int i = 0; i < NW; ++i) e[i] = (i < n) ? y[i] - c.yref[i] : 0.0;
    MPC_UNROLL for (int i = 0; i < NW; ++i) {
      double a = 0.0;
      MPC_UNROLL for (int j = 0; j < NW; ++j) a += c.W[i * NW + j] * e[j];
      g[i] = (i < n) ? s * (a + c.flin[i]) : 0.0;
      if (i < n) val += 0.5 * a * e[i] + c.flin[i] * y[i];
    }
    return s * val;
  }

  // ---------------- box rows of one stage ----------------
  // Variables v = [u ; x[bx]]; row r < NV is the lower bound of v_r, row NV + r its upper bound
  // (acados order [lbu, lbx, ubu, ubx], rlmpc/common/utils.py:4-25).  A row that does not exist at
  // this stage (u at stage N or clamped in Q-mode, x at stage 0, infinite bound) gets lb = -inf /
  // ub = +inf and is skipped everywhere.
  // Soft state bounds (constraints.idxsbx, linear penalty zl/zu, nlp.py:1099-1134): the bound row
  // reads lb - x - sl <= 0 with the extra row -sl <= 0 (multiplier/slack pair R_LS + j); the slack
  // value is the slack of that extra row, sl = t[R_LS + j].  Stationarity in sl, z = lam_b + lam_s,
  // is eliminated inside the interior-point iteration (soft_coeffs).
  struct Bnd {
    double lb[NV], ub[NV];
    double zl[NSXA], zu[NSXA];  // scaled penalties s_k * z of the soft pairs, < 0: pair not present at this stage
  };
  MPC_HD static int vidx(int r) { return r < NU ? NX + r : M::bx(r - NU); }  // r < NU + NBX: index into w = [x;u]
  // Jacobian row of v_r w.r.t. w = [x;u]: a unit vector for the box rows, the model's constant row for g
  MPC_HD static double jrow(int r, int i) {
    if (r < NU + NBX) return vidx(r) == i ? 1.0 : 0.0;
    return M::gC(r - NU - NBX, i);
  }
  // H += c J_r' J_r ;  gvec += a J_r
  MPC_HD static void jr_rank1(int r, double c, double* Hm) {
    if (r < NU + NBX) {
      const int ix = vidx(r);
      Hm[ix * NW + ix] += c;
    } else {
      MPC_UNROLL for (int a = 0; a < NW; ++a) MPC_UNROLL for (int b = 0; b < NW; ++b)
        Hm[a * NW + b] += c * M::gC(r - NU - NBX, a) * M::gC(r - NU - NBX, b);
    }
  }
  MPC_HD static void jr_axpy(int r, double a, double* gvec) {
    if (r < NU + NBX) {
      gvec[vidx(r)] += a;
    } else {
      MPC_UNROLL for (int i = 0; i < NW; ++i) gvec[i] += a * M::gC(r - NU - NBX, i);
    }
  }
  MPC_HD static double jr_dot(int r, const double* dw) {
    if (r < NU + NBX) return dw[vidx(r)];
    double a = 0.0;
    MPC_UNROLL for (int i = 0; i < NW; ++i) a += M::gC(r - NU - NBX, i) * dw[i];
    return a;
  }
  MPC_HD static int soft_row(int j) { return NU + M::sx(j); }               // variable index (in v) of soft pair j
  MPC_HD static void stage_bounds(const ProblemData& pd, int k, Bnd& bd) {
    const bool uact = (k < pd.N) && !(k == 0 && pd.mode == MODE_Q);
    MPC_UNROLL for (int i = 0; i < NU; ++i) {
      const bool act = uact && !(k == 0 && i < pd.fix0);
      bd.lb[i] = act ? pd.lbu[i] : -1e300;
      bd.ub[i] = act ? pd.ubu[i] : 1e300;
    }
    MPC_UNROLL for (int j = 0; j < NBX; ++j) {
      const int ix = M::bx(j);
      bd.lb[NU + j] = (k == 0) ? -1e300 : (k == pd.N ? pd.lbx_e[ix] : pd.lbx[ix]);
      bd.ub[NU + j] = (k == 0) ? 1e300 : (k == pd.N ? pd.ubx_e[ix] : pd.ubx[ix]);
    }
    MPC_UNROLL for (int j = 0; j < NG; ++j) {  // nh rows live on stages 0..N-1; constant when x_0 and u_0 are both fixed
      bd.lb[NU + NBX + j] = uact ? pd.lg[j] : -1e300;
      bd.ub[NU + NBX + j] = uact ? pd.ug[j] : 1e300;
    }
    MPC_UNROLL for (int j = 0; j < NSXA; ++j) {
      const bool sact = NSX > 0 && k >= 1 && k < pd.N;
      bd.zl[j] = (sact && bd.lb[NSX > 0 ? soft_row(j) : 0] > -BIG) ? pd.scale[k] * pd.zl[j] : -1.0;
      bd.zu[j] = (sact && bd.ub[NSX > 0 ? soft_row(j) : 0] < BIG) ? pd.scale[k] * pd.zu[j] : -1.0;
    }
  }
  // is bound row q (< 2 NV) softened at this stage?  returns the slack row or -1
  MPC_HD static int slack_of(const Bnd& bd, int r, int side) {
    int q = -1;
    MPC_UNROLL for (int j = 0; j < NSX; ++j) {
      if (soft_row(j) == r && (side ? bd.zu[j] : bd.zl[j]) >= 0.0) q = (side ? R_US : R_LS) + j;
    }
    return q;
  }
  MPC_HD static int count_rows(const ProblemData& pd) {
    int m = 0;
    for (int k = 0; k <= pd.N; ++k) {
      Bnd bd;
      stage_bounds(pd, k, bd);
      MPC_UNROLL for (int r = 0; r < NV; ++r) m += (bd.lb[r] > -BIG) + (bd.ub[r] < BIG);
      MPC_UNROLL for (int j = 0; j < NSX; ++j) m += (bd.zl[j] >= 0.0) + (bd.zu[j] >= 0.0);
    }
    return m;
  }
  MPC_HD static void stage_vars(const double* x, const double* u, double* v) {
    MPC_UNROLL for (int i = 0; i < NU; ++i) v[i] = u[i];
    MPC_UNROLL for (int j = 0; j < NBX; ++j) v[NU + j] = x[M::bx(j)];
    MPC_UNROLL for (int j = 0; j < NG; ++j) {
      double a = M::g0(j);
      MPC_UNROLL for (int i = 0; i < NX; ++i) a += M::gC(j, i) * x[i];
      MPC_UNROLL for (int i = 0; i < NU; ++i) a += M::gC(j, NX + i) * u[i];
      v[NU + NBX + j] = a;
    }
  }
  MPC_HD static double row_range(const Bnd& bd, int r) {
    return (bd.lb[r] > -BIG && bd.ub[r] < BIG) ? bd.ub[r] - bd.lb[r] : 1.0;
  }
  // warm-start safeguard: keep (lam,t) strictly inside the cone
  MPC_HD static void clip_rows(const Bnd& bd, double* lam, double* t) {
    MPC_UNROLL for (int r = 0; r < NV; ++r) {
      const double range = row_range(bd, r);
      t[r] = dmax(t[r], 1e-10 * range);
      t[NV + r] = dmax(t[NV + r], 1e-10 * range);
      lam[r] = dmax(lam[r], 1e-14);
      lam[NV + r] = dmax(lam[NV + r], 1e-14);
    }
    MPC_UNROLL for (int j = 0; j < NSX; ++j) {
      const double range = row_range(bd, soft_row(j));
      t[R_LS + j] = dmax(t[R_LS + j], 1e-10 * range);
      t[R_US + j] = dmax(t[R_US + j], 1e-10 * range);
      lam[R_LS + j] = dmax(lam[R_LS + j], 1e-14);
      lam[R_US + j] = dmax(lam[R_US + j], 1e-14);
    }
  }
  // One bound row in "lam_hat = a - c * d" form, d = distance to the bound after the step WITHOUT
  // the slack.  Hard row: c = lam/t, a = target/t + lam.  Soft row (slack row qs, scaled penalty z):
  // eliminating the new slack s_hat from  z = lam_hat_b + lam_hat_s  gives
  //   s_hat = (a_b + a_s - z - c_b d) / (c_b + c_s),  c = c_b c_s/(c_b + c_s),  a = (a_b c_s - c_b (a_s - z))/(c_b + c_s).
  struct RowC {
    double c, a;            // condensed
    double cb, ab, cs, as_, z;  // raw pieces (soft rows)
  };
  MPC_HD static RowC row_coeffs(const double* lam, const double* t, int q, int qs, double z, double target) {
    RowC rc;
    const double itb = 1.0 / t[q];
    rc.cb = lam[q] * itb;
    rc.ab = target * itb + lam[q];
    rc.c = rc.cb;
    rc.a = rc.ab;
    rc.cs = 0.0; rc.as_ = 0.0; rc.z = z;
    if (NSX > 0 && qs >= 0) {
      const double its = 1.0 / t[qs];
      rc.cs = lam[qs] * its;
      rc.as_ = target * its + lam[qs];
      const double den = 1.0 / (rc.cb + rc.cs);
      rc.c = rc.cb * rc.cs * den;
      rc.a = (rc.ab * rc.cs - rc.cb * (rc.as_ - z)) * den;
    }
    return rc;
  }
  // condensed barrier terms: Hm += J' diag(c) J,  g += J'(+-)(a - c d0),  d0 = v - lb / ub - v
  // (g = [gq ; gr] of the stage)
  MPC_HD static void barrier_add(const Bnd& bd, const double* v, const double* lam, const double* t,
                                 double target, double* Hm, double* g) {
    MPC_UNROLL for (int r = 0; r < NV; ++r) {
      if (bd.lb[r] > -BIG) {
        const int qs = slack_of(bd, r, 0);
        const RowC rc = row_coeffs(lam, t, r, qs, qs >= 0 ? bd.zl[qs - R_LS] : 0.0, target);
        jr_rank1(r, rc.c, Hm);
        jr_axpy(r, -(rc.a - rc.c * (v[r] - bd.lb[r])), g);
      }
      if (bd.ub[r] < BIG) {
        const int qs = slack_of(bd, r, 1);
        const RowC rc = row_coeffs(lam, t, NV + r, qs, qs >= 0 ? bd.zu[qs - R_US] : 0.0, target);
        jr_rank1(r, rc.c, Hm);
        jr_axpy(r, rc.a - rc.c * (bd.ub[r] - v[r]), g);
      }
    }
  }
  // Hessian part only (sensitivity factorisation).  Slack values are constants there, like in the
  // reference (quirk Q4: slacks are not part of z), so a softened row counts with its own lam/t.
  MPC_HD static void barrier_hess(const Bnd& bd, const double* lam, const double* t, double* Hm) {
    MPC_UNROLL for (int r = 0; r < NV; ++r) {
      if (bd.lb[r] > -BIG) jr_rank1(r, lam[r] / t[r], Hm);
      if (bd.ub[r] < BIG) jr_rank1(r, lam[NV + r] / t[NV + r], Hm);
    }
  }
  struct StepStats {
    double amax, s0, s1, s2, cmax;
  };
  MPC_HD static void step_stats(double lam, double t, double lh, double th, StepStats& S) {
    const double dt = th - t, dl = lh - lam;
    if (dt < 0.0) S.amax = dmin(S.amax, -t / dt);
    if (dl < 0.0) S.amax = dmin(S.amax, -lam / dl);
    S.s0 += lam * t;
    S.s1 += lam * dt + t * dl;
    S.s2 += dl * dt;
    S.cmax = dmax(S.cmax, dabs(dl * dt));
  }
  // new slacks/multipliers of the rows for the primal step dw = [dx;du]; step-length statistics
  MPC_HD static void rows_forward(const Bnd& bd, const double* v, const double* dw, const double* lam,
                                  const double* t, double target, double* lh, double* th, StepStats& S) {
    MPC_UNROLL for (int q = 2 * NV; q < NR; ++q) {
      lh[q] = 0.0;
      th[q] = 0.0;
    }
    MPC_UNROLL for (int r = 0; r < NV; ++r) {
      const double dv = jr_dot(r, dw);
      MPC_UNROLL for (int side = 0; side < 2; ++side) {
        const int q = side * NV + r;
        const bool act = side ? (bd.ub[r] < BIG) : (bd.lb[r] > -BIG);
        if (!act) {
          lh[q] = 0.0;
          th[q] = 0.0;
          continue;
        }
        // (v - lb) + dv, NOT (v + dv) - lb: the slack must use the same rounded distance to the
        // bound as the condensed gradient, or lam_hat picks up c*ulp(v) ~ 1e-8 of noise
        const double d = side ? (bd.ub[r] - v[r]) - dv : (v[r] - bd.lb[r]) + dv;
        const int qs = slack_of(bd, r, side);
        const RowC rc = row_coeffs(lam, t, q, qs, qs >= 0 ? (side ? bd.zu[qs - R_US] : bd.zl[qs - R_LS]) : 0.0, target);
        if (NSX > 0 && qs >= 0) {
          // Soft pair.  Evaluate in the order that avoids cancellation amplified by a huge lam/t:
          if (rc.cb > rc.cs) {
            // bound row (nearly) active, slack free: lam_hat_b from the condensed form, the rest from it
            lh[q] = rc.a - rc.c * d;
            th[q] = (rc.ab - lh[q]) / rc.cb;
            th[qs] = th[q] - d;
            lh[qs] = rc.z - lh[q];
          } else {
            // slack (nearly) pinned at zero: s_hat first
            th[qs] = (rc.ab + rc.as_ - rc.z - rc.cb * d) / (rc.cb + rc.cs);
            lh[qs] = rc.as_ - rc.cs * th[qs];
            th[q] = d + th[qs];
            lh[q] = rc.ab - rc.cb * th[q];
          }
          step_stats(lam[qs], t[qs], lh[qs], th[qs], S);
        } else {
          th[q] = d;
          lh[q] = rc.ab - rc.cb * d;
        }
        step_stats(lam[q], t[q], lh[q], th[q], S);
      }
    }
  }

  // step-length statistics of given (lam_hat, t_hat), same traversal as rows_forward (used when the Newton
  // step of a stage was computed elsewhere and only its rows are at hand)
  MPC_HD static void rows_stats(const Bnd& bd, const double* lam, const double* t, const double* lh, const double* th,
                                StepStats& S) {
    MPC_UNROLL for (int r = 0; r < NV; ++r) {
      MPC_UNROLL for (int side = 0; side < 2; ++side) {
        const int q = side * NV + r;
        const bool act = side ? (bd.ub[r] < BIG) : (bd.lb[r] > -BIG);
        if (!act) continue;
        const int qs = slack_of(bd, r, side);
        if (NSX > 0 && qs >= 0) step_stats(lam[qs], t[qs], lh[qs], th[qs], S);
        step_stats(lam[q], t[q], lh[q], th[q], S);
      }
    }
  }

  struct Residuals {
    double stat, eq, ineq, comp, cost;
  };
  // comp / ineq residuals of the rows of one stage and their contribution J'lam to stationarity;
  // soft pairs add the stationarity residual in the slack, z - lam_b - lam_s
  MPC_HD static void rows_residual(const ProblemData& pd, const Bnd& bd, const double* v, const double* lam,
                                   const double* t, Residuals& R, double* jl /* NW, += */) {
    MPC_UNROLL for (int r = 0; r < NV; ++r) {
      MPC_UNROLL for (int side = 0; side < 2; ++side) {
        const int q = side * NV + r;
        const bool act = side ? (bd.ub[r] < BIG) : (bd.lb[r] > -BIG);
        if (!act) continue;
        const int qs = slack_of(bd, r, side);
        double sl = 0.0;
        if (NSX > 0 && qs >= 0) {
          sl = t[qs];  // slack value
          const double z = side ? bd.zu[qs - R_US] : bd.zl[qs - R_LS];
          R.stat = dmax(R.stat, dabs(z - lam[q] - lam[qs]));
          R.comp = dmax(R.comp, dabs(lam[qs] * t[qs] - pd.tau));
        }
        jr_axpy(r, side ? lam[q] : -lam[q], jl);
        R.comp = dmax(R.comp, dabs(lam[q] * t[q] - pd.tau));
        const double h = side ? v[r] - bd.ub[r] - sl : bd.lb[r] - v[r] - sl;
        R.ineq = dmax(R.ineq, dabs(h + t[q]));
      }
    }
  }
  // slack penalty of the stage, sum_j z_l sl_j + z_u su_j (scaled), part of the cost (nlp.py:1099-1134)
  MPC_HD static double slack_cost(const Bnd& bd, const double* t) {
    double c = 0.0;
    MPC_UNROLL for (int j = 0; j < NSX; ++j) {
      if (bd.zl[j] >= 0.0) c += bd.zl[j] * t[R_LS + j];
      if (bd.zu[j] >= 0.0) c += bd.zu[j] * t[R_US + j];
    }
    return c;
  }

  // One backward Riccati step.  In: P (NX x NX full), p; stage data.  Out: P, p (overwritten), K, kff.
  // Returns false if the reduced Hessian block G is not positive definite (PD = false: not invertible, see spd_solve).
  template <bool PD = true>
  MPC_HD static bool riccati_step(double* P, double* p, const double* A, const double* B, const double* b,
                                  const double* Hm /* NW x NW: [Q S'; S R] incl. barrier */,
                                  const double* g /* NW: [gq ; gr] */, double* K, double* kff,
                                  double* Ginv /* optional NU x NU or nullptr */) {
    const double* gq = g;
    const double* gr = g + NX;
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
    const bool ok = spd_solve<NU, NX + 1 + NU, PD>(G, R);
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

  // =======================================================================================
  // (sample, stage) function: linearise stage k at the current iterate.
  // k < N: A, B, b = F(x_k,u_k) - x_{k+1}, scaled cost gradient q, r, stage cost, defect norm.
  // k = N: terminal cost gradient and value.
  // =======================================================================================
  MPC_HD static double stage_slack_cost(const ProblemData& pd, const Lane& L, int k) {
    Bnd bd;
    double t[NR];
    stage_bounds(pd, k, bd);
    ld<NR>(L.it + (size_t)it_t(pd.N, k) * TILE, TILE, t);
    return slack_cost(bd, t);
  }

  MPC_HD static void lin_stage(const ProblemData& pd, const Lane& L, int k) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    double* w = L.ws + (size_t)k * W_REC * bs;
    CostK ck;
    double y[NW], g[NW];
    if (k < N) {
      double xn[NX], xnext[NX], A[NX * NX], B[NX * NU], bb[NX];
      ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, y);
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, y + NX);
      ld<NX>(L.it + (size_t)it_x(N, k + 1) * bs, bs, xnext);
      M::dyn_lin(y, y + NX, L.th, (size_t)TILE, pd.mc, xn, A, B);
      double eq = 0.0;
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        bb[i] = xn[i] - xnext[i];
        eq = dmax(eq, dabs(bb[i]));
        if (!(bb[i] == bb[i])) eq = bb[i];  // propagate NaN
      }
      st<NX * NX>(w + (size_t)W_A * bs, bs, A);
      st<NX * NU>(w + (size_t)W_B * bs, bs, B);
      st<NX>(w + (size_t)W_b * bs, bs, bb);
      load_cost(k == 0 ? 0 : 1, L, ck);
      double c = cost_grad(ck, pd.scale[k], NW, y, g);
      if (NSX > 0) c += stage_slack_cost(pd, L, k);
      st<NW>(w + (size_t)W_q * bs, bs, g);
      w[(size_t)W_c * bs] = c;
      w[(size_t)W_e * bs] = eq;
    } else {
      ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, y);
      MPC_UNROLL for (int i = 0; i < NU; ++i) y[NX + i] = 0.0;
      load_cost(2, L, ck);
      const double c = cost_grad(ck, pd.scale[N], NX, y, g);
      st<NX>(w + (size_t)W_q * bs, bs, g);
      w[(size_t)W_c * bs] = c;
      w[(size_t)W_e * bs] = 0.0;
    }
  }

  // =======================================================================================
  // Sample function, fast path of one SQP iteration.  One backward sweep evaluates acados'
  // convergence test (the four KKT residual norms) and, fused with it, the Riccati factorisation
  // of ONE interior-point Newton iteration started at the stored (lam,t) with target tau; a
  // forward sweep computes the step.  If that iteration is a full step that lands on the
  // tau-central point (the usual case when the active set did not change) the step is applied.
  // Otherwise nothing is modified and the sample is handed to qp_full().
  // =======================================================================================
  enum Fast : int { FAST_CONVERGED = 0, FAST_STEPPED = 1, FAST_HARD = 2, FAST_NAN = 3 };

  // swept (optional): set (1 or 2) when FAST_HARD is returned AFTER the forward sweep, i.e. the workspace holds the
  // complete first warm iteration (K, k, dx, du, lam_hat, t_hat at target tau) for the queue kernel to reuse
  // polish: "converged" additionally requires the iterate to sit ON the central path (|lam*t - tau| <= 5 % tau,
  // the accuracy update_nlp's R = 0 assumes for the sensitivities), not merely within the acceptance
  // neighbourhood comp_accept of the steps on the way; otherwise one more (cheap, warm) Newton iteration
  // follows.  Off for the final test-only round of an SQP solve, where the plain acados criterion decides.
  // LIN: linearise every stage right before the backward sweep reads it (the sample's own thread does lin_stage();
  // the record still goes to memory for the forward sweep and the queue kernel, but is read back from cache here, and
  // the separate (sample, stage) launch disappears).
  template <bool LIN = false>
  MPC_HD static int qp_fast(const ProblemData& pd, const Lane& L, Residuals& R, int* swept = nullptr, bool polish = true) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    const bool warm = pd.warm_ipm && L.it[(size_t)it_meta(N) * bs] > 0.5;
    const double target = pd.tau;
    R.stat = R.eq = R.ineq = R.comp = R.cost = 0.0;
    bool failed = false;
    double P[NX * NX], p[NX];
    double carry[NX];  // x-stationarity of stage k+1 without the -pi_k term
    {                  // terminal stage
      if (LIN) lin_stage(pd, L, N);
      const double* w = L.ws + (size_t)N * W_REC * bs;
      double g[NW], Hm[NW * NW];
      ld<NX>(w + (size_t)W_q * bs, bs, g);
      MPC_UNROLL for (int i = 0; i < NU; ++i) g[NX + i] = 0.0;
      R.cost += w[(size_t)W_c * bs];
      MPC_UNROLL for (int i = 0; i < NX; ++i) carry[i] = g[i];
      load_W(2, pd.scale[N], L, Hm);
      if (NBX > 0) {
        Bnd bd;
        double x[NX], u0[NU], v[NV], lam[NR], t[NR], jl[NW];
        stage_bounds(pd, N, bd);
        ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, x);
        MPC_UNROLL for (int i = 0; i < NU; ++i) u0[i] = 0.0;
        stage_vars(x, u0, v);
        ld<NR>(L.it + (size_t)it_lam(N, N) * bs, bs, lam);
        ld<NR>(L.it + (size_t)it_t(N, N) * bs, bs, t);
        MPC_UNROLL for (int i = 0; i < NW; ++i) jl[i] = 0.0;
        rows_residual(pd, bd, v, lam, t, R, jl);
        MPC_UNROLL for (int i = 0; i < NX; ++i) carry[i] += jl[i];
        if (warm) {
          clip_rows(bd, lam, t);
          barrier_add(bd, v, lam, t, target, Hm, g);
        }
      }
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        p[i] = g[i];
        MPC_UNROLL for (int j = 0; j < NX; ++j) P[i * NX + j] = Hm[i * NW + j];
      }
    }
    for (int k = N - 1; k >= 0; --k) {
      if (LIN) lin_stage(pd, L, k);
      double* w = L.ws + (size_t)k * W_REC * bs;
      double A[NX * NX], B[NX * NU], bb[NX], g[NW], x[NX], u[NU], pik[NX], lam[NR], t[NR], v[NV];
      ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
      ld<NX * NU>(w + (size_t)W_B * bs, bs, B);
      ld<NX>(w + (size_t)W_b * bs, bs, bb);
      ld<NW>(w + (size_t)W_q * bs, bs, g);
      R.cost += w[(size_t)W_c * bs];
      {
        const double e = w[(size_t)W_e * bs];
        R.eq = (e == e) ? dmax(R.eq, e) : e;
      }
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      if (NEEDX) ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
      ld<NX>(L.it + (size_t)it_pi(N, k) * bs, bs, pik);
      ld<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
      ld<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
      Bnd bd;
      stage_bounds(pd, k, bd);
      stage_vars(x, u, v);
      // ---- residuals at the current iterate (raw multipliers) ----
      MPC_UNROLL for (int i = 0; i < NX; ++i) R.stat = dmax(R.stat, dabs(carry[i] - pik[i]));
      double jl[NW];
      MPC_UNROLL for (int i = 0; i < NW; ++i) jl[i] = 0.0;
      rows_residual(pd, bd, v, lam, t, R, jl);
      const bool ufixed = (k == 0 && pd.mode == MODE_Q);
      if (!ufixed) {
        MPC_UNROLL for (int i = 0; i < NU; ++i) {
          double a = g[NX + i] + jl[NX + i];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * pik[l];
          R.stat = dmax(R.stat, dabs(a));
        }
      }
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        double a = g[i] + jl[i];
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[l * NX + i] * pik[l];
        carry[i] = a;
      }
      // ---- Riccati step of the Newton iteration ----
      if (warm) {
        double Hm[NW * NW], K[NU * NX], kff[NU];
        load_W(k == 0 ? 0 : 1, pd.scale[k], L, Hm);
        clip_rows(bd, lam, t);
        barrier_add(bd, v, lam, t, target, Hm, g);
        if (!ufixed) {
          if (!riccati_step(P, p, A, B, bb, Hm, g, K, kff, nullptr)) failed = true;
        } else {
          MPC_UNROLL for (int i = 0; i < NU * NX; ++i) K[i] = 0.0;
          MPC_UNROLL for (int i = 0; i < NU; ++i) kff[i] = 0.0;
        }
        st<NU * NX>(w + (size_t)W_K * bs, bs, K);
        st<NU>(w + (size_t)W_k * bs, bs, kff);
      }
    }
    const double rmax = dmax(dmax(R.stat, R.eq), dmax(R.ineq, R.comp));
    if (!(rmax == rmax) || !(R.cost == R.cost)) return FAST_NAN;
    if (rmax < pd.tol && (!polish || !warm || R.comp <= 0.05 * pd.tau)) return FAST_CONVERGED;
    if (!warm || failed) return FAST_HARD;
    StepStats S = {1e300, 0.0, 0.0, 0.0, 0.0};
    forward_sweep(pd, L, target, /*clip=*/true, S);
    if (!(S.amax == S.amax)) return FAST_HARD;
    if (S.amax >= 1.0 / 0.995 && S.cmax <= dmin(pd.comp_accept * pd.tau, 0.1 * pd.tol)) {
      apply_step(pd, L, 1.0, /*clip=*/true);
      return FAST_STEPPED;
    }
    if (swept) *swept = (S.amax >= 1.0 / 0.995) ? 2 : 1;  // 2: a full step, only complementarity left to polish (one more iteration)
    return FAST_HARD;
  }

  // ---------------------------------------------------------------------------------------
  // How a sweep READS the data of stage k (writes always go straight to the arrays).  The default
  // reads in place.  The CUDA build also has a reader that prefetches whole stage records into a
  // shared-memory ring with cp.async, two stages ahead of the recursion (rlmpc_b200.cu): the
  // Riccati recursions are a dependent chain per sample, so when few samples are in flight (the
  // queued full interior-point solves) global-memory latency, not bandwidth, is what they wait for.
  // Interface: begin(k_first, dir, count) starts a sweep over `count` stages; ws(k) / rows(k) return
  // stride-TILE views of the workspace record and of [lam(NR) t(NR) u(NU) x(NX)] of stage k, valid
  // until done(k), which releases the stage and lets the reader fetch ahead.
  // ---------------------------------------------------------------------------------------
  static constexpr int RD_LAM = 0, RD_T = NR, RD_U = 2 * NR, RD_X = 2 * NR + NU, RD_ROWS = 2 * NR + NU + NX;
  static constexpr int RD_WS = W_dx;  // staged workspace elements [W_A, W_dx): everything a sweep reads
  struct DirectReader {
    const Lane& L;
    int N;
    MPC_HD DirectReader(const Lane& L_, int N_) : L(L_), N(N_) {}
    MPC_HD void begin(int, int, int) {}
    MPC_HD const double* ws(int k) const { return L.ws + (size_t)k * W_REC * TILE; }
    MPC_HD void rows(int k, double* lam, double* t, double* u, double* x) const {
      ld<NR>(L.it + (size_t)it_lam(N, k) * TILE, TILE, lam);
      ld<NR>(L.it + (size_t)it_t(N, k) * TILE, TILE, t);
      if (k < N) {
        ld<NU>(L.it + (size_t)it_u(N, k) * TILE, TILE, u);
      } else {
        MPC_UNROLL for (int i = 0; i < NU; ++i) u[i] = 0.0;
      }
      if (NEEDX) ld<NX>(L.it + (size_t)it_x(N, k) * TILE, TILE, x);
    }
    MPC_HD void done(int) {}
  };

  // Forward sweep of one interior-point iteration: dx, du from the feedback law, new slacks and
  // multipliers (lam_hat, t_hat) of every row, fraction-to-boundary statistics.
  MPC_HD static void forward_sweep(const ProblemData& pd, const Lane& L, double target, bool clip, StepStats& S) {
    DirectReader rd(L, pd.N);
    forward_sweep(pd, L, target, clip, S, rd);
  }
  template <class RD>
  MPC_HD static void forward_sweep(const ProblemData& pd, const Lane& L, double target, bool clip, StepStats& S, RD& rd) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    double dw[NW];
    MPC_UNROLL for (int i = 0; i < NW; ++i) dw[i] = 0.0;
    rd.begin(0, +1, N + 1);
    for (int k = 0; k <= N; ++k) {
      double* w = L.ws + (size_t)k * W_REC * bs;
      const double* wr = rd.ws(k);
      double A[NX * NX], B[NX * NU], bb[NX];
      if (k < N) {
        double K[NU * NX];
        ld<NX * NX>(wr + (size_t)W_A * bs, bs, A);
        ld<NX * NU>(wr + (size_t)W_B * bs, bs, B);
        ld<NX>(wr + (size_t)W_b * bs, bs, bb);
        ld<NU * NX>(wr + (size_t)W_K * bs, bs, K);
        ld<NU>(wr + (size_t)W_k * bs, bs, dw + NX);
        MPC_UNROLL for (int i = 0; i < NU; ++i) MPC_UNROLL for (int l = 0; l < NX; ++l) dw[NX + i] += K[i * NX + l] * dw[l];
        st<NU>(w + (size_t)W_du * bs, bs, dw + NX);
      } else {
        MPC_UNROLL for (int i = 0; i < NU; ++i) dw[NX + i] = 0.0;
      }
      st<NX>(w + (size_t)W_dx * bs, bs, dw);
      if (k < N || NBX > 0) {
        Bnd bd;
        double x[NX], u[NU], v[NV], lam[NR], t[NR], lh[NR], th[NR];
        stage_bounds(pd, k, bd);
        rd.rows(k, lam, t, u, x);
        stage_vars(x, u, v);
        if (clip) clip_rows(bd, lam, t);
        rows_forward(bd, v, dw, lam, t, target, lh, th, S);
        st<NR>(w + (size_t)W_lh * bs, bs, lh);
        st<NR>(w + (size_t)W_th * bs, bs, th);
      }
      rd.done(k);
      if (k < N) {
        double dxn[NX];
        MPC_UNROLL for (int i = 0; i < NX; ++i) {
          double a = bb[i];
          MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[i * NX + l] * dw[l];
          MPC_UNROLL for (int l = 0; l < NU; ++l) a += B[i * NU + l] * dw[NX + l];
          dxn[i] = a;
        }
        MPC_UNROLL for (int i = 0; i < NX; ++i) dw[i] = dxn[i];
      }
    }
  }

  // ---------------------------------------------------------------------------------------
  // Full interior-point solve of the stage QP (Riccati factorisation per iteration).
  // State of the method is (lam,t) only (absolute-step form); the primal step dx,du of the last
  // iteration is left in the workspace.  Returns the number of IPM iterations, <0 on failure.
  // ---------------------------------------------------------------------------------------
  static constexpr int WARM_LIMIT = 6;  // IPM iterations granted to a warm start before a cold restart

  // (lam,t) of the rows of one stage at the start of an interior-point solve.  warm: the stored values,
  // clipped into the cone; cold: slacks from the current point, lam = mu0 / t.  Returns sum(lam*t).
  MPC_HD static double ipm_init_stage(const ProblemData& pd, const Bnd& bd, const double* v, bool warm, double* lam, double* t) {
    double mu = 0.0;
    if (warm) {
      clip_rows(bd, lam, t);
    } else {
      MPC_UNROLL for (int q = 0; q < NR; ++q) { lam[q] = 0.0; t[q] = 0.0; }
      MPC_UNROLL for (int r = 0; r < NV; ++r) {
        const double tmin = 1e-2 * row_range(bd, r);
        MPC_UNROLL for (int side = 0; side < 2; ++side) {
          const int q = side * NV + r;
          double d = side ? bd.ub[r] - v[r] : v[r] - bd.lb[r];
          const int qs = slack_of(bd, r, side);
          if (NSX > 0 && qs >= 0) {  // slack: cover a violated bound, stay strictly positive
            t[qs] = dmax(-d, 0.0) + tmin;
            lam[qs] = pd.mu0 / t[qs];
            d += t[qs];
          }
          t[q] = dmax(d, tmin);
          lam[q] = pd.mu0 / t[q];
        }
      }
    }
    MPC_UNROLL for (int r = 0; r < NV; ++r) {
      if (bd.lb[r] > -BIG) mu += lam[r] * t[r]; else { lam[r] = 0.0; t[r] = 0.0; }
      if (bd.ub[r] < BIG) mu += lam[NV + r] * t[NV + r]; else { lam[NV + r] = 0.0; t[NV + r] = 0.0; }
    }
    MPC_UNROLL for (int j = 0; j < NSX; ++j) {
      if (bd.zl[j] >= 0.0) mu += lam[R_LS + j] * t[R_LS + j]; else { lam[R_LS + j] = 0.0; t[R_LS + j] = 0.0; }
      if (bd.zu[j] >= 0.0) mu += lam[R_US + j] * t[R_US + j]; else { lam[R_US + j] = 0.0; t[R_US + j] = 0.0; }
    }
    return mu;
  }
  // returns sum(lam*t)
  MPC_HD static double ipm_init(const ProblemData& pd, const Lane& L, bool warm) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    double mu = 0.0;
    for (int k = 0; k <= N; ++k) {
      if (k == N && NBX == 0) break;
      Bnd bd;
      double x[NX], u[NU], v[NV], lam[NR], t[NR];
      stage_bounds(pd, k, bd);
      if (k < N) {
        ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      } else {
        MPC_UNROLL for (int i = 0; i < NU; ++i) u[i] = 0.0;
      }
      if (NEEDX) ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
      stage_vars(x, u, v);
      if (warm) {
        ld<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
        ld<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
      }
      mu += ipm_init_stage(pd, bd, v, warm, lam, t);
      st<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
      st<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
    }
    return mu;
  }

  // Active-set step of a warm start (primal-dual active-set idea, Hintermueller/Ito/Kunisch): the
  // Newton step from the tau-central point of the previous QP is taken IN FULL, and the rows it
  // drives through zero are flipped instead of cutting the step: a slack that collapses makes
  // its row active (t <- eps, lam kept positive), a multiplier that turns negative releases its row
  // (lam <- tau/t).  All other rows take their Newton values.  A handful of such steps identify the
  // new active set when few rows change; the regular iteration then polishes to lam*t = tau.
  // Returns sum(lam*t).
  MPC_HD static bool row_active(const Bnd& bd, int q) {
    if (q < 2 * NV) {
      const int r = q < NV ? q : q - NV;
      return q < NV ? (bd.lb[r] > -BIG) : (bd.ub[r] < BIG);
    }
    return q < R_US ? (bd.zl[q - R_LS] >= 0.0) : (bd.zu[q - R_US] >= 0.0);
  }
  // the rows of one stage; returns their sum(lam*t)
  MPC_HD static double ipm_project_stage(const ProblemData& pd, const Bnd& bd, double* lam, double* t, const double* lh,
                                         const double* th) {
    double mu = 0.0;
    MPC_UNROLL for (int q = 0; q < NR; ++q) {
      if (!row_active(bd, q)) continue;
      const double range = row_range(bd, q < 2 * NV ? (q < NV ? q : q - NV) : soft_row(q < R_US ? q - R_LS : q - R_US));
      const double eps_t = 1e-9 * range;
      if (!(th[q] > eps_t)) {          // slack collapses: the row becomes (stays) active
        // central for the final target: with t = eps_t a row whose product lam * eps_t exceeds tau asks the next
        // Newton step for dt ~ -t, i.e. a step length of 1 + O(tau / (lam eps_t)) < 1 / 0.995, which reads as
        // "infeasible" again -- the evaporation queue spent all of its active-set steps in that loop
        lam[q] = dmax(dmax(lh[q], lam[q]), 1e-3);
        t[q] = dmin(eps_t, pd.tau / lam[q]);
      } else if (!(lh[q] > 0.0)) {     // multiplier changes sign: the row is released
        // th was computed with the row's barrier weight lam/t still in the Hessian, so it underestimates the free
        // slack by orders of magnitude; keeping it leaves a weight tau/th^2 that only fades over several further
        // active-set steps (measured on the closed-loop cart-pole workload: 8 -> 3.6 iterations per queued QP).
        // Release to a slack of AS_RELEASE * range instead: weight ~ 0, and lam_hat = tau/t (2 - d/t) of the next
        // step stays positive for every d the bounds allow (d <= range = 2 t).
        t[q] = dmax(th[q], AS_RELEASE * range);
        lam[q] = pd.tau / t[q];
      } else {
        t[q] = th[q];
        lam[q] = lh[q];
      }
      mu += lam[q] * t[q];
    }
    return mu;
  }
  MPC_HD static double ipm_project(const ProblemData& pd, const Lane& L) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    double mu = 0.0;
    for (int k = 0; k <= N; ++k) {
      if (k == N && NBX == 0) break;
      const double* w = L.ws + (size_t)k * W_REC * bs;
      Bnd bd;
      double lam[NR], t[NR], lh[NR], th[NR];
      stage_bounds(pd, k, bd);
      ld<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
      ld<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
      ld<NR>(w + (size_t)W_lh * bs, bs, lh);
      ld<NR>(w + (size_t)W_th * bs, bs, th);
      mu += ipm_project_stage(pd, bd, lam, t, lh, th);
      st<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
      st<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
    }
    return mu;
  }

  // pending damped update (lam,t) += alpha ((lam_hat,t_hat) - (lam,t)) of the rows of stage k (lam, t
  // in: current values; wr: read view of the stage record holding lam_hat, t_hat)
  MPC_HD static void rows_update(const Lane& L, int N, int k, double alpha, const double* wr, double* lam, double* t) {
    constexpr size_t bs = TILE;
    if (alpha > 0.0) {
      double lh[NR], th[NR];
      ld<NR>(wr + (size_t)W_lh * bs, bs, lh);
      ld<NR>(wr + (size_t)W_th * bs, bs, th);
      MPC_UNROLL for (int i = 0; i < NR; ++i) {
        lam[i] += alpha * (lh[i] - lam[i]);
        t[i] += alpha * (th[i] - t[i]);
      }
      st<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
      st<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
    }
  }

  // The interior-point loop is written as a resumable state machine: ipm_begin() initialises the rows and the
  // scalars of the method, every ipm_trip() is one trip of the loop (one Newton iteration: two Riccati sweeps and the
  // decision what to do with the step).  qp_ipm() runs it to the end inside one thread; the pass kernels of the CUDA
  // build (k_ipm_pass) run ONE trip per launch over all queued samples and carry IpmState through global memory, so
  // that samples with different iteration counts do not hold each other up inside a warp.
  struct IpmState {
    double mu, alpha, sigma;  // barrier estimate, pending step length of the previous trip (0: none), centring parameter
    int iters, warm_iters, as_iters;
    int warm;                 // still on the warm start (active-set steps allowed)
    int code;                 // 0 running, 1 converged, -1 Riccati failure / NaN, -2 step length collapsed (cold start)
  };
  static constexpr int IPM_STATE_WORDS = 8;  // doubles per sample in the pass kernels' state array
  MPC_HD static void ipm_begin(const ProblemData& pd, const Lane& L, IpmState& s) {
    const int N = pd.N;
    // warm = keep the multipliers of the previous QP (clipped away from zero), cold = slacks from the current
    // point, lam = mu0 / t
    s.warm = (pd.warm_ipm && L.it[(size_t)it_meta(N) * TILE] > 0.5) ? 1 : 0;
    s.mu = ipm_init(pd, L, s.warm != 0) / (double)count_rows(pd);
    s.alpha = 0.0;  // pending step length of the previous iteration (0: nothing pending)
    s.sigma = s.warm ? pd.sigma_min : pd.sigma0;
    s.iters = s.warm_iters = s.as_iters = 0;
    s.code = 0;
  }
  // Would the first trip of a warm start repeat exactly the Newton iteration qp_fast() has just done (target tau from
  // the clipped stored rows)?  Then its lam_hat, t_hat are still in the stage records and the trip can start from them.
  MPC_HD static bool ipm_first_trip_done(const ProblemData& pd, const IpmState& s) {
    return s.iters == 0 && s.warm && dmax(s.sigma * s.mu, pd.tau) == pd.tau && pd.max_ipm > 1;
  }
  // One trip.  reuse: lam_hat, t_hat of this trip's Newton iteration are already in the stage records (see above):
  // only the step statistics are recomputed, the two Riccati sweeps are skipped.
  template <class RD>
  MPC_HD static void ipm_trip(const ProblemData& pd, const Lane& L, IpmState& s, RD& rd, bool reuse = false) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    const bool qmode = pd.mode == MODE_Q;
    const double m_rows = (double)count_rows(pd);
    bool failed = false;
    ++s.iters;
    if (s.warm && (s.warm_iters >= WARM_LIMIT + s.as_iters || (s.warm_iters > s.as_iters && s.alpha > 0.0 && s.alpha < 0.05))) {
      // the warm start is jammed (active set changed too much): restart from a cold point
      s.warm = 0;
      s.mu = ipm_init(pd, L, false) / m_rows;
      s.alpha = 0.0;
      s.sigma = pd.sigma0;
    }
    if (s.warm) ++s.warm_iters;
    const double target = dmax(s.sigma * s.mu, pd.tau);
    StepStats S = {1e300, 0.0, 0.0, 0.0, 0.0};
    if (reuse) {
      for (int k = 0; k <= N; ++k) {
        if (k == N && NBX == 0) break;
        const double* w = L.ws + (size_t)k * W_REC * bs;
        Bnd bd;
        double lam[NR], t[NR], lh[NR], th[NR];
        stage_bounds(pd, k, bd);
        ld<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
        ld<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
        ld<NR>(w + (size_t)W_lh * bs, bs, lh);
        ld<NR>(w + (size_t)W_th * bs, bs, th);
        rows_stats(bd, lam, t, lh, th, S);
      }
    } else {
      // ---------------- backward sweep ----------------
      double P[NX * NX], p[NX];
      rd.begin(N, -1, N + 1);
      {
        double Hm[NW * NW], g[NW];
        const double* wr = rd.ws(N);
        stage_hess(2, pd.scale[N], L, wr, Hm);
        ld<NX>(wr + (size_t)W_q * bs, bs, g);
        MPC_UNROLL for (int i = 0; i < NU; ++i) g[NX + i] = 0.0;
        if (NBX > 0) {
          Bnd bd;
          double x[NX], u0[NU], v[NV], lam[NR], t[NR];
          stage_bounds(pd, N, bd);
          rd.rows(N, lam, t, u0, x);
          stage_vars(x, u0, v);
          rows_update(L, N, N, s.alpha, wr, lam, t);
          barrier_add(bd, v, lam, t, target, Hm, g);
        }
        rd.done(N);
        MPC_UNROLL for (int i = 0; i < NX; ++i) {
          p[i] = g[i];
          MPC_UNROLL for (int jj = 0; jj < NX; ++jj) P[i * NX + jj] = Hm[i * NW + jj];
        }
      }
      for (int k = N - 1; k >= 0; --k) {
        double* w = L.ws + (size_t)k * W_REC * bs;
        const double* wr = rd.ws(k);
        double A[NX * NX], B[NX * NU], bb[NX], g[NW], K[NU * NX], kff[NU];
        ld<NX * NX>(wr + (size_t)W_A * bs, bs, A);
        ld<NX * NU>(wr + (size_t)W_B * bs, bs, B);
        ld<NX>(wr + (size_t)W_b * bs, bs, bb);
        ld<NW>(wr + (size_t)W_q * bs, bs, g);
        double Hm[NW * NW];
        stage_hess(k == 0 ? 0 : 1, pd.scale[k], L, wr, Hm);
        {
          Bnd bd;
          double x[NX], u[NU], v[NV], lam[NR], t[NR];
          stage_bounds(pd, k, bd);
          rd.rows(k, lam, t, u, x);
          stage_vars(x, u, v);
          rows_update(L, N, k, s.alpha, wr, lam, t);
          barrier_add(bd, v, lam, t, target, Hm, g);
        }
        rd.done(k);
        const bool ufixed = (k == 0 && qmode);
        if (!ufixed) {
          if (!riccati_step(P, p, A, B, bb, Hm, g, K, kff, nullptr)) failed = true;
        } else {
          // u_0 fixed (Q-mode): no feedback; x_0 is fixed as well so P_0, p_0 are not needed
          MPC_UNROLL for (int i = 0; i < NU * NX; ++i) K[i] = 0.0;
          MPC_UNROLL for (int i = 0; i < NU; ++i) kff[i] = 0.0;
        }
        st<NU * NX>(w + (size_t)W_K * bs, bs, K);
        st<NU>(w + (size_t)W_k * bs, bs, kff);
      }
      // ---------------- forward sweep ----------------
      forward_sweep(pd, L, target, /*clip=*/false, S, rd);
    }
    if ((failed || !(S.amax == S.amax)) && s.warm) {
      // a warm start (or its active-set steps) went wrong numerically: not an error, start over cold
      s.warm = 0;
      s.mu = ipm_init(pd, L, false) / m_rows;
      s.alpha = 0.0;
      s.sigma = pd.sigma0;
      return;
    }
    if (failed || !(S.amax == S.amax)) {
      s.code = -1;
      return;
    }
    if (s.warm && S.amax < 1.0 / 0.995 && s.as_iters < (int)pd.as_steps) {
      // infeasible Newton step of a warm start: full step + projection instead of a short step
      ++s.as_iters;
#ifdef AS_TRACE  // host debugging aid: which row limits the Newton step (before the projection overwrites lam, t)
      {
        double best = 1e300; int bk = -1, bq = -1, kind = 0; double bl = 0, bt = 0, blh = 0, bth = 0;
        for (int k = 0; k < N; ++k)
          for (int q = 0; q < NR; ++q) {
            const double lam = L.it[(size_t)(it_lam(N, k) + q) * TILE], t = L.it[(size_t)(it_t(N, k) + q) * TILE];
            const double lh = L.ws[((size_t)k * W_REC + W_lh + q) * TILE], th = L.ws[((size_t)k * W_REC + W_th + q) * TILE];
            if (th - t < 0 && -t / (th - t) < best) { best = -t / (th - t); bk = k; bq = q; kind = 0; bl = lam; bt = t; blh = lh; bth = th; }
            if (lh - lam < 0 && -lam / (lh - lam) < best) { best = -lam / (lh - lam); bk = k; bq = q; kind = 1; bl = lam; bt = t; blh = lh; bth = th; }
          }
        printf("      limiting row: stage %d row %d %s lam %.3e t %.3e -> lh %.3e th %.3e (target %.2e)\n", bk, bq, kind ? "dl<0" : "dt<0", bl, bt, blh, bth, target);
      }
#endif
      s.mu = ipm_project(pd, L) / m_rows;
#ifdef AS_TRACE  // ... and the active set after the step (NU = 1 problems)
      {
        char buf[MAXN + 1];
        int n = 0;
        for (int k = 0; k < N; ++k) {
          const double tl = L.it[(size_t)it_t(N, k) * TILE], tu = L.it[(size_t)(it_t(N, k) + NV) * TILE];
          buf[n++] = tl < 1e-6 ? 'L' : (tu < 1e-6 ? 'U' : '.');
        }
        buf[n] = 0;
        printf("   as %2d amax %.3g  %s\n", s.as_iters, S.amax, buf);
      }
#endif
      s.alpha = 0.0;
      s.sigma = pd.sigma_min;
      return;
    }
    s.alpha = (S.amax >= 1.0 / 0.995) ? 1.0 : 0.995 * S.amax;
    if (!s.warm && s.alpha < 1e-9) {  // a cold-started iteration has collapsed onto the boundary (HPIPM: MIN_STEP),
      s.code = -2;                    // e.g. infeasible QP.  (A jammed WARM start restarts cold instead, above.)
      return;
    }
    const double mu_new = (S.s0 + s.alpha * S.s1 + s.alpha * s.alpha * S.s2) / m_rows;
#ifdef IPM_TRACE
    printf("  ipm it %2d warm %d sigma %.3f mu %.3e target %.3e amax %.4g alpha %.4g cmax %.3e mu_new %.3e\n", s.iters, s.warm,
           s.sigma, s.mu, target, S.amax, s.alpha, S.cmax, mu_new);
#endif
    if (target <= pd.tau && s.alpha == 1.0 && S.cmax <= dmin(pd.comp_accept * pd.tau, 0.1 * pd.tol)) s.code = 1;
    // centring heuristic: aggressive after long steps, conservative after short ones
    const double r = 1.0 - s.alpha;
    s.sigma = dmin(0.8, dmax(pd.sigma_min, r * r * 4.0 + pd.sigma_min));
    s.mu = mu_new;
  }
  // return value of qp_ipm for a finished (or abandoned: iteration limit) state
  MPC_HD static int ipm_result(const IpmState& s) {
    if (s.code == -1) return -1;
    if (s.code == -2) return -2;
    return s.code == 1 ? s.iters : -(s.iters + 1000);
  }

  MPC_HD static int qp_ipm(const ProblemData& pd, const Lane& L, double* alpha_out) {
    DirectReader rd(L, pd.N);
    return qp_ipm(pd, L, alpha_out, rd);
  }
  template <class RD>
  MPC_HD static int qp_ipm(const ProblemData& pd, const Lane& L, double* alpha_out, RD& rd) {
    IpmState s;
    ipm_begin(pd, L, s);
    for (int j = 0; j < pd.max_ipm && s.code == 0; ++j) ipm_trip(pd, L, s, rd);
    *alpha_out = s.alpha;
    return ipm_result(s);
  }
  // ---------------------------------------------------------------------------------------
  // Apply the QP step: w += dw, (lam,t) <- last IPM update, pi <- QP multipliers (backward
  // recursion of the x-stationarity rows).
  // ---------------------------------------------------------------------------------------
  // damp_primal: the QP was not solved (iteration limit): move the primal variables by alpha * dw
  // only, the interior iterate of the method, instead of the full Newton target.
  MPC_HD static void apply_step(const ProblemData& pd, const Lane& L, double alpha, bool clip, bool damp_primal = false) {
    const double ap = (damp_primal ? alpha : 1.0) * pd.step_length;
    const int N = pd.N;
    constexpr size_t bs = TILE;
    double pik[NX];  // pi_k (multiplier of x_{k+1} = F(x_k,u_k))
    for (int k = N; k >= 0; --k) {
      double* w = L.ws + (size_t)k * W_REC * bs;
      Bnd bd;
      double dw[NW], g[NW], Wm[NW * NW], lam[NR];
      ld<NX>(w + (size_t)W_dx * bs, bs, dw);
      if (k < N) {
        ld<NU>(w + (size_t)W_du * bs, bs, dw + NX);
        ld<NW>(w + (size_t)W_q * bs, bs, g);
      } else {
        MPC_UNROLL for (int i = 0; i < NU; ++i) dw[NX + i] = 0.0;
        ld<NX>(w + (size_t)W_q * bs, bs, g);
        MPC_UNROLL for (int i = 0; i < NU; ++i) g[NX + i] = 0.0;
      }
      stage_bounds(pd, k, bd);
      MPC_UNROLL for (int i = 0; i < NR; ++i) lam[i] = 0.0;
      if (k < N || NBX > 0) {
        double t[NR], lh[NR], th[NR];
        ld<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
        ld<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
        if (clip) clip_rows(bd, lam, t);
        ld<NR>(w + (size_t)W_lh * bs, bs, lh);
        ld<NR>(w + (size_t)W_th * bs, bs, th);
        MPC_UNROLL for (int q = 0; q < NR; ++q) {
          bool act;
          if (q < 2 * NV) {
            const int r = q < NV ? q : q - NV;
            act = q < NV ? (bd.lb[r] > -BIG) : (bd.ub[r] < BIG);
          } else {
            act = q < R_US ? (bd.zl[q - R_LS] >= 0.0) : (bd.zu[q - R_US] >= 0.0);
          }
          lam[q] = act ? lam[q] + alpha * (lh[q] - lam[q]) : 0.0;
          t[q] = act ? t[q] + alpha * (th[q] - t[q]) : 0.0;
        }
        st<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
        st<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
      }
      if (k < N) st<NX>(L.it + (size_t)it_pi(N, k) * bs, bs, pik);  // pi_k was completed at stage k+1
      if (k > 0) {
        // x-stationarity of the QP at stage k gives pi_{k-1}
        double pin[NX];
        load_W(k == N ? 2 : 1, pd.scale[k], L, Wm);
        MPC_UNROLL for (int i = 0; i < NX; ++i) {
          double a = g[i];
          MPC_UNROLL for (int j = 0; j < NW; ++j) a += Wm[i * NW + j] * dw[j];
          pin[i] = a;
        }
        if (k < N) {
          double A[NX * NX];
          ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
          MPC_UNROLL for (int i = 0; i < NX; ++i) MPC_UNROLL for (int l = 0; l < NX; ++l) pin[i] += A[l * NX + i] * pik[l];
        }
        {
          double jl[NW];
          MPC_UNROLL for (int i = 0; i < NW; ++i) jl[i] = 0.0;
          MPC_UNROLL for (int r = NU; r < NV; ++r) jr_axpy(r, -lam[r] + lam[NV + r], jl);  // rows that touch x
          MPC_UNROLL for (int i = 0; i < NX; ++i) pin[i] += jl[i];
        }
        MPC_UNROLL for (int i = 0; i < NX; ++i) pik[i] = pin[i];
        double x[NX];
        ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
        MPC_UNROLL for (int i = 0; i < NX; ++i) x[i] += ap * dw[i];
        st<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
      }
      if (k < N) {
        double u[NU];
        ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
        MPC_UNROLL for (int i = 0; i < NU; ++i) u[i] += ap * dw[NX + i];
        st<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      }
    }
    L.it[(size_t)it_meta(N) * bs] = 1.0;
  }

  // Sample function, slow path of one SQP iteration (after qp_fast returned FAST_HARD): the full
  // interior-point loop and the step.
  enum Full : int {
    FULL_OK = 0,       // QP solved, step applied
    FULL_MAXITER = 1,  // interior-point iteration limit: the (damped) last step is applied and the SQP
                       // goes on, like acados, which tolerates ACADOS_MAXITER from the QP solver
    FULL_FAILED = 2,   // reduced Hessian not positive definite, or step length collapsed (HPIPM MIN_STEP):
                       // iterate left untouched, acados would return ACADOS_QP_FAILURE (4)
  };
  MPC_HD static int qp_full(const ProblemData& pd, const Lane& L, int* ipm_iters) {
    DirectReader rd(L, pd.N);
    return qp_full(pd, L, ipm_iters, rd);
  }
  template <class RD>
  MPC_HD static int qp_full(const ProblemData& pd, const Lane& L, int* ipm_iters, RD& rd) {
    double alpha = 0.0;
    const int r = qp_ipm(pd, L, &alpha, rd);
    return full_result(pd, L, r, alpha, ipm_iters);
  }
  // r: return value of qp_ipm
  MPC_HD static int full_result(const ProblemData& pd, const Lane& L, int r, double alpha, int* ipm_iters) {
    if (ipm_iters) *ipm_iters += (r > 0) ? r : ((r >= -2) ? 0 : -(r + 1000));
    if (r == -1 || r == -2) return FULL_FAILED;
    apply_step(pd, L, alpha, /*clip=*/false, /*damp_primal=*/r < 0);
    return (r < 0) ? FULL_MAXITER : FULL_OK;
  }

  // One launch of the pass kernel for one queued sample: (first pass: start the method; if qp_fast() has already done
  // the first Newton iteration, pick its result up) + ONE full trip.  st: this sample's IPM_STATE_WORDS state words,
  // stride ss.  Returns the Full code when the sample finished in this pass (the step is applied), -1 when it goes on.
  MPC_HD static int ipm_pass(const ProblemData& pd, const Lane& L, double* st, size_t ss, bool first, bool swept, int* ipm_iters) {
    IpmState s;
    DirectReader rd(L, pd.N);
    if (first) {
      ipm_begin(pd, L, s);
      if (swept && ipm_first_trip_done(pd, s)) ipm_trip(pd, L, s, rd, /*reuse=*/true);
    } else {
      s.mu = st[0]; s.alpha = st[ss]; s.sigma = st[2 * ss];
      s.iters = (int)st[3 * ss]; s.warm_iters = (int)st[4 * ss]; s.as_iters = (int)st[5 * ss]; s.warm = (int)st[6 * ss];
      s.code = 0;
    }
    if (s.code == 0 && s.iters < pd.max_ipm) ipm_trip(pd, L, s, rd, false);
    if (s.code != 0 || s.iters >= pd.max_ipm) return full_result(pd, L, ipm_result(s), s.alpha, ipm_iters);
    st[0] = s.mu; st[ss] = s.alpha; st[2 * ss] = s.sigma;
    st[3 * ss] = (double)s.iters; st[4 * ss] = (double)s.warm_iters; st[5 * ss] = (double)s.as_iters; st[6 * ss] = (double)s.warm;
    return -1;
  }

  MPC_HD static void set_initial(const ProblemData& pd, const Lane& L, const double* x0, size_t x0s, const double* u0,
                                 size_t u0s) {
    const int N = pd.N;
    MPC_UNROLL for (int i = 0; i < NX; ++i) L.it[(size_t)(it_x(N, 0) + i) * TILE] = x0[(size_t)i * x0s];
    if (pd.mode == MODE_Q) {
      MPC_UNROLL for (int i = 0; i < NU; ++i) L.it[(size_t)(it_u(N, 0) + i) * TILE] = u0[(size_t)i * u0s];
    }
  }

  // =======================================================================================
  // Evaluation + sensitivities at the current iterate (update_nlp, nlp.py:1341-1563):
  //   * cost and KKT residuals (what update_nlp asserts, nlp.py:1445-1537)
  //   * dL/dtheta (model part; cost part when pd.param_cost)           nlp.py:1211-1212,1401
  //   * dpi/dtheta via ONE exact-Hessian Riccati factorisation and NU adjoint solves, instead of
  //     the reference's dense (nz x nz) sparse LU with ntheta right-hand sides  nlp.py:1413-1424
  // Gradient rows have width ng = grad_width(pd): the model parameters only (the structurally
  // non-zero prefix of the reference's p, quirk Q8) or the whole p when pd.param_cost is set.
  // =======================================================================================
  MPC_HD static int grad_width(const ProblemData& pd) { return pd.param_cost ? M::NTH : NPM; }

  // (sample, stage) function: exact second-order information of stage k at (x_k, u_k, pi_k).
  MPC_HD static void sens_stage(const ProblemData& pd, const Lane& L, int k) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    double* w = L.ws + (size_t)k * W_REC * bs;
    CostK ck;
    double y[NW], g[NW];
    if (k < N) {
      double pik[NX], xk1[NX], xn[NX], A[NX * NX], B[NX * NU], Fp[NX * NPM], Hww[NW * NW], Hwp[NW * NPM];
      ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, y);
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, y + NX);
      ld<NX>(L.it + (size_t)it_pi(N, k) * bs, bs, pik);
      ld<NX>(L.it + (size_t)it_x(N, k + 1) * bs, bs, xk1);
      M::dyn_sens(y, y + NX, L.th, (size_t)TILE, pd.mc, pik, xn, A, B, Fp, Hww, Hwp);
      double eq = 0.0;
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        const double d = xn[i] - xk1[i];
        eq = dmax(eq, dabs(d));
        if (!(d == d)) eq = d;
      }
      st<NX * NX>(w + (size_t)W_A * bs, bs, A);
      st<NX * NU>(w + (size_t)W_B * bs, bs, B);
      double Hp[NWS];
      MPC_UNROLL for (int i = 0; i < NW; ++i) MPC_UNROLL for (int j = i; j < NW; ++j) Hp[pidx(i, j)] = Hww[i * NW + j];
      st<NWS>(w + (size_t)S_H * bs, bs, Hp);
      if constexpr (!M::PARAMS_COST_ONLY) {
        double gp[NPM];
        MPC_UNROLL for (int j = 0; j < NPM; ++j) {
          double a = 0.0;
          MPC_UNROLL for (int i = 0; i < NX; ++i) a += pik[i] * Fp[i * NPM + j];
          gp[j] = a;
        }
        M::cost_sens(k == 0 ? 0 : 1, pd.scale[k], y, L.th, (size_t)TILE, gp, Hwp);  // model parameters that enter the stage cost
        st<NW * NPM>(w + (size_t)S_Hwp * bs, bs, Hwp);
        st<NX * NPM>(w + (size_t)S_Fp * bs, bs, Fp);
        st<NPM>(w + (size_t)S_gp * bs, bs, gp);
      }  // else: the parameters sit in the cost only; sens_sweep evaluates the contractions it needs on the fly
      load_cost(k == 0 ? 0 : 1, L, ck);
      double c = cost_grad(ck, pd.scale[k], NW, y, g);
      if (NSX > 0) c += stage_slack_cost(pd, L, k);
      st<NW>(w + (size_t)S_g * bs, bs, g);
      w[(size_t)S_c * bs] = c;
      w[(size_t)S_e * bs] = eq;
    } else {
      ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, y);
      MPC_UNROLL for (int i = 0; i < NU; ++i) y[NX + i] = 0.0;
      load_cost(2, L, ck);
      const double c = cost_grad(ck, pd.scale[N], NX, y, g);
      st<NW>(w + (size_t)S_g * bs, bs, g);
      w[(size_t)S_c * bs] = c;
      w[(size_t)S_e * bs] = 0.0;
    }
  }

  // Sample function: residuals, dL/dtheta, exact-Hessian factorisation (backward) and the NU
  // adjoint solves (forward).  dLdth / dpidth point at this sample's rows of row-major
  // [B, ng] / [B, NU, ng] outputs (nullptr: skip).
  MPC_HD static Residuals sens_sweep(const ProblemData& pd, const Lane& L, double* dLdth, double* dpidth, int* ok_out) {
    const int N = pd.N;
    constexpr size_t bs = TILE;
    const bool qmode = pd.mode == MODE_Q;
    Residuals R = {0, 0, 0, 0, 0};
    bool ok = true;
    double gp[NPM];
    MPC_UNROLL for (int i = 0; i < NPM; ++i) gp[i] = 0.0;
    double P[NX * NX], pdummy[NX], carry[NX];
    // dpi/dtheta for a handful of model parameters (cart-pole: 3 or 4): DIRECT sensitivities -- the KKT system of
    // update_nlp with right-hand side -dR/dtheta_j is a Riccati solve with gradient Hwp[:, j] and affine term Fp[:, j],
    // i.e. one more vector recursion p_j per parameter riding on the factorisation of the backward sweep, and
    // dpi/dtheta_j = du_0 = -G_0^{-1} gv_0 (x_0 is fixed).  No forward pass, no P / K records, every stage record is
    // read once.  With many parameters (cost parameters, linear system) the adjoint solves below are cheaper.
    constexpr bool FWD_OK = !M::PARAMS_COST_ONLY && NPM > 0 && NPM <= 4;
    constexpr int NPF = FWD_OK ? NPM : 1;
    const bool fwd = FWD_OK && !pd.param_cost && dpidth != nullptr;
    double pv[NPF * NX], dpi_f[NU * NPF];
    MPC_UNROLL for (int i = 0; i < NPF * NX; ++i) pv[i] = 0.0;
    MPC_UNROLL for (int i = 0; i < NU * NPF; ++i) dpi_f[i] = 0.0;
    {  // terminal stage
      const double* w = L.ws + (size_t)N * W_REC * bs;
      double g[NW], Hm[NW * NW];
      ld<NW>(w + (size_t)S_g * bs, bs, g);
      R.cost += w[(size_t)S_c * bs];
      MPC_UNROLL for (int i = 0; i < NX; ++i) carry[i] = g[i];
      load_W(2, pd.scale[N], L, Hm);
      if (NBX > 0 || (pd.param_cost && dLdth)) {
        double x[NX];
        ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, x);
        if (pd.param_cost && dLdth) M::cost_param_grad(2, pd.scale[N], L.th, (size_t)TILE, x, nullptr, dLdth);
        if (NBX > 0) {
          Bnd bd;
        double u0[NU], v[NV], lam[NR], t[NR], jl[NW];
          stage_bounds(pd, N, bd);
          MPC_UNROLL for (int i = 0; i < NU; ++i) u0[i] = 0.0;
          stage_vars(x, u0, v);
          ld<NR>(L.it + (size_t)it_lam(N, N) * bs, bs, lam);
          ld<NR>(L.it + (size_t)it_t(N, N) * bs, bs, t);
          MPC_UNROLL for (int i = 0; i < NW; ++i) jl[i] = 0.0;
          rows_residual(pd, bd, v, lam, t, R, jl);
          MPC_UNROLL for (int i = 0; i < NX; ++i) carry[i] += jl[i];
          barrier_hess(bd, lam, t, Hm);
        }
      }
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        pdummy[i] = 0.0;
        MPC_UNROLL for (int j = 0; j < NX; ++j) P[i * NX + j] = Hm[i * NW + j];
      }
    }
    for (int k = N - 1; k >= 0; --k) {
      double* w = L.ws + (size_t)k * W_REC * bs;
      double A[NX * NX], B[NX * NU], g[NW], Hp[NWS], gpk[NPM], x[NX], u[NU], pik[NX], lam[NR], t[NR], v[NV];
      ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
      ld<NX * NU>(w + (size_t)W_B * bs, bs, B);
      ld<NW>(w + (size_t)S_g * bs, bs, g);
      ld<NWS>(w + (size_t)S_H * bs, bs, Hp);
      if constexpr (!M::PARAMS_COST_ONLY) ld<NPM>(w + (size_t)S_gp * bs, bs, gpk);
      R.cost += w[(size_t)S_c * bs];
      {
        const double e = w[(size_t)S_e * bs];
        R.eq = (e == e) ? dmax(R.eq, e) : e;
      }
      ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, u);
      if (NEEDX || M::PARAMS_COST_ONLY || (pd.param_cost && dLdth)) ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, x);
      if constexpr (M::PARAMS_COST_ONLY) {
        double yk[NW];
        MPC_UNROLL for (int i = 0; i < NX; ++i) yk[i] = x[i];
        MPC_UNROLL for (int i = 0; i < NU; ++i) yk[NX + i] = u[i];
        M::cost_sens_grad(k == 0 ? 0 : 1, pd.scale[k], yk, L.th, (size_t)TILE, gp);
      } else {
        MPC_UNROLL for (int j = 0; j < NPM; ++j) gp[j] += gpk[j];
      }
      ld<NX>(L.it + (size_t)it_pi(N, k) * bs, bs, pik);
      ld<NR>(L.it + (size_t)it_lam(N, k) * bs, bs, lam);
      ld<NR>(L.it + (size_t)it_t(N, k) * bs, bs, t);
      if (pd.param_cost && dLdth) M::cost_param_grad(k == 0 ? 0 : 1, pd.scale[k], L.th, (size_t)TILE, x, u, dLdth);
      Bnd bd;
      stage_bounds(pd, k, bd);
      stage_vars(x, u, v);
      // ---- residuals ----
      MPC_UNROLL for (int i = 0; i < NX; ++i) R.stat = dmax(R.stat, dabs(carry[i] - pik[i]));
      double jl[NW];
      MPC_UNROLL for (int i = 0; i < NW; ++i) jl[i] = 0.0;
      rows_residual(pd, bd, v, lam, t, R, jl);
      const bool ufixed = (k == 0 && qmode);
      double su[NU];  // u-stationarity remainder (multiplier of the clamped u_0 in Q-mode)
      MPC_UNROLL for (int i = 0; i < NU; ++i) {
        double a = g[NX + i] + jl[NX + i];
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += B[l * NU + i] * pik[l];
        su[i] = a;
        if (!ufixed) R.stat = dmax(R.stat, dabs(a));
      }
      MPC_UNROLL for (int i = 0; i < NX; ++i) {
        double a = g[i] + jl[i];
        MPC_UNROLL for (int l = 0; l < NX; ++l) a += A[l * NX + i] * pik[l];
        carry[i] = a;
      }
      if (k == 0) {  // multipliers of the eliminated stage-0 equalities
        MPC_UNROLL for (int i = 0; i < NX; ++i) L.it[(size_t)(it_rx0(N) + i) * bs] = carry[i];
        MPC_UNROLL for (int i = 0; i < NU; ++i) L.it[(size_t)(it_ru0(N) + i) * bs] = su[i];
      }
      // ---- exact-Hessian Riccati factorisation ----
      if (dpidth) {
        double Hm[NW * NW];
        load_W(k == 0 ? 0 : 1, pd.scale[k], L, Hm);
        MPC_UNROLL for (int i = 0; i < NW; ++i) MPC_UNROLL for (int j = i; j < NW; ++j) {
          Hm[i * NW + j] += Hp[pidx(i, j)];
          if (j != i) Hm[j * NW + i] = Hm[i * NW + j];
        }
        barrier_hess(bd, lam, t, Hm);
        double Hwp[NW * NPF], vj[NPF * NX];
        if (fwd) {  // v_j = p_j + P_{k+1} Fp[:, j]  (before P is overwritten)
          if constexpr (FWD_OK) {
            double Fp[NX * NPM];
            ld<NW * NPM>(w + (size_t)S_Hwp * bs, bs, Hwp);
            ld<NX * NPM>(w + (size_t)S_Fp * bs, bs, Fp);
            MPC_UNROLL for (int j = 0; j < NPM; ++j) MPC_UNROLL for (int i = 0; i < NX; ++i) {
              double a = pv[j * NX + i];
              MPC_UNROLL for (int l = 0; l < NX; ++l) a += P[i * NX + l] * Fp[l * NPM + j];
              vj[j * NX + i] = a;
            }
          }
        } else {
          double Pp[NPS];
          int c_ = 0;
          MPC_UNROLL for (int i = 0; i < NX; ++i) MPC_UNROLL for (int j = i; j < NX; ++j) Pp[c_++] = P[i * NX + j];
          st<NPS>(w + (size_t)S_P * bs, bs, Pp);
        }
        double K[NU * NX], kff[NU], Ginv[NU * NU], zg[NW], zb[NX];
        MPC_UNROLL for (int i = 0; i < NW; ++i) zg[i] = 0.0;
        MPC_UNROLL for (int i = 0; i < NX; ++i) zb[i] = 0.0;
        if (!ufixed) {
#ifdef RLMPC_SENS_REQUIRE_PD  // (round-1 behaviour; a test-side build uses it to label the samples of the indefinite-Hessian fixture)
          if (!riccati_step<true>(P, pdummy, A, B, zb, Hm, zg, K, kff, Ginv)) ok = false;
#else
          if (!riccati_step<false>(P, pdummy, A, B, zb, Hm, zg, K, kff, Ginv)) ok = false;
#endif
        } else {
          MPC_UNROLL for (int i = 0; i < NU * NX; ++i) K[i] = 0.0;
          MPC_UNROLL for (int i = 0; i < NU * NU; ++i) Ginv[i] = 0.0;
        }
        if (fwd) {
          if constexpr (FWD_OK) {
            if (!ufixed) {
              MPC_UNROLL for (int j = 0; j < NPM; ++j) {
                double gv[NU];
                MPC_UNROLL for (int a = 0; a < NU; ++a) {
                  double v = Hwp[(NX + a) * NPM + j];
                  MPC_UNROLL for (int l = 0; l < NX; ++l) v += B[l * NU + a] * vj[j * NX + l];
                  gv[a] = v;
                }
                if (k == 0) {
                  MPC_UNROLL for (int a = 0; a < NU; ++a) {
                    double v = 0.0;
                    MPC_UNROLL for (int b = 0; b < NU; ++b) v -= Ginv[a * NU + b] * gv[b];
                    dpi_f[a * NPM + j] = v;
                  }
                }
                MPC_UNROLL for (int i = 0; i < NX; ++i) {  // p_j <- q_j + A'v_j + H'kff_j,  H'kff_j = K'gv_j
                  double v = Hwp[i * NPM + j];
                  MPC_UNROLL for (int l = 0; l < NX; ++l) v += A[l * NX + i] * vj[j * NX + l];
                  MPC_UNROLL for (int a = 0; a < NU; ++a) v += K[a * NX + i] * gv[a];
                  pv[j * NX + i] = v;
                }
              }
            }
          }
        } else {
          st<NU * NX>(w + (size_t)S_K * bs, bs, K);
          if (k == 0) st<NU * NU>(w + (size_t)S_Gi * bs, bs, Ginv);
        }
      }
    }
    MPC_UNROLL for (int i = 0; i < NX; ++i) (void)pdummy[i];
    if (dLdth) {
      MPC_UNROLL for (int j = 0; j < NPM; ++j) dLdth[j] = gp[j];
    }
    if (fwd) {
      if constexpr (FWD_OK) {
        MPC_UNROLL for (int r = 0; r < NU; ++r) MPC_UNROLL for (int j = 0; j < NPM; ++j)
          dpidth[(size_t)r * grad_width(pd) + j] = qmode ? 0.0 : dpi_f[r * NPM + j];
      }
    }
    // ---- forward pass: NU adjoint solves  K y_i = e_{u0,i},  dpi_i/dtheta = -y_i' dR/dtheta ----
    if (dpidth && !fwd) {
      double acc[NU * NPM];
      MPC_UNROLL for (int i = 0; i < NU * NPM; ++i) acc[i] = 0.0;
      if (!qmode) {
        double yx[NU * NX];  // y_x of each right-hand side
        MPC_UNROLL for (int i = 0; i < NU * NX; ++i) yx[i] = 0.0;
        for (int k = 0; k < N; ++k) {
          const double* w = L.ws + (size_t)k * W_REC * bs;
          constexpr int NDP = M::PARAMS_COST_ONLY ? 1 : NPM;  // dense per-stage parameter derivatives: only if stored
          double A[NX * NX], B[NX * NU], K[NU * NX], Pp[NPS], Hwp[NW * NDP], Fp[NX * NDP], yk[NW];
          ld<NX * NX>(w + (size_t)W_A * bs, bs, A);
          ld<NX * NU>(w + (size_t)W_B * bs, bs, B);
          ld<NU * NX>(w + (size_t)S_K * bs, bs, K);
          ld<NPS>(w + (size_t)S_P * bs, bs, Pp);
          if constexpr (M::PARAMS_COST_ONLY) {
            ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, yk);
            ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, yk + NX);
          } else {
            ld<NW * NPM>(w + (size_t)S_Hwp * bs, bs, Hwp);
            ld<NX * NPM>(w + (size_t)S_Fp * bs, bs, Fp);
          }
          double Pf[NX * NX];
          {
            int c_ = 0;
            MPC_UNROLL for (int i = 0; i < NX; ++i) MPC_UNROLL for (int j = i; j < NX; ++j) {
              Pf[i * NX + j] = Pp[c_];
              Pf[j * NX + i] = Pp[c_++];
            }
          }
          double Gi[NU * NU];
          if (k == 0) ld<NU * NU>(w + (size_t)S_Gi * bs, bs, Gi);
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
            if constexpr (M::PARAMS_COST_ONLY) {
              double yw[NW];
              MPC_UNROLL for (int i = 0; i < NX; ++i) yw[i] = yx[r * NX + i];
              MPC_UNROLL for (int i = 0; i < NU; ++i) yw[NX + i] = yu[i];
              M::cost_sens_adj(k == 0 ? 0 : 1, pd.scale[k], yk, L.th, (size_t)TILE, yw, acc + r * NPM);
            } else {
              MPC_UNROLL for (int j = 0; j < NPM; ++j) {
                double a = 0.0;
                MPC_UNROLL for (int i = 0; i < NX; ++i) a += yx[r * NX + i] * Hwp[i * NPM + j] + ypi[i] * Fp[i * NPM + j];
                MPC_UNROLL for (int i = 0; i < NU; ++i) a += yu[i] * Hwp[(NX + i) * NPM + j];
                acc[r * NPM + j] -= a;
              }
            }
            if (!M::PARAMS_COST_ONLY && pd.param_cost) {
              // cost-parameter columns (parameterize_tracking_cost): contracted on the fly into the caller's zero-initialised row
              double xk[NX], uk[NU], yw[NW];
              ld<NX>(L.it + (size_t)it_x(N, k) * bs, bs, xk);
              ld<NU>(L.it + (size_t)it_u(N, k) * bs, bs, uk);
              MPC_UNROLL for (int i = 0; i < NX; ++i) yw[i] = yx[r * NX + i];
              MPC_UNROLL for (int i = 0; i < NU; ++i) yw[NX + i] = yu[i];
              M::cost_param_adj(k == 0 ? 0 : 1, pd.scale[k], L.th, (size_t)TILE, xk, uk, yw, dpidth + (size_t)r * grad_width(pd));
            }
            MPC_UNROLL for (int i = 0; i < NX; ++i) yx[r * NX + i] = yxn[i];
          }
        }
        if (!M::PARAMS_COST_ONLY && pd.param_cost) {  // terminal stage
          double xk[NX], yw[NW];
          ld<NX>(L.it + (size_t)it_x(N, N) * bs, bs, xk);
          MPC_UNROLL for (int r = 0; r < NU; ++r) {
            MPC_UNROLL for (int i = 0; i < NX; ++i) yw[i] = yx[r * NX + i];
            MPC_UNROLL for (int i = 0; i < NU; ++i) yw[NX + i] = 0.0;
            M::cost_param_adj(2, pd.scale[N], L.th, (size_t)TILE, xk, nullptr, yw, dpidth + (size_t)r * grad_width(pd));
          }
        }
      }
      MPC_UNROLL for (int r = 0; r < NU; ++r) MPC_UNROLL for (int j = 0; j < NPM; ++j) dpidth[(size_t)r * grad_width(pd) + j] = acc[r * NPM + j];
    }
    *ok_out = ok ? 1 : 0;
    return R;
  }
};

}  // namespace rlmpc
