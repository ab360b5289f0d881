// Warp-cooperative SQP / Riccati interior-point / sensitivity engine for problems with DENSE stage blocks too large for
// one thread: the chain of masses (nx = 9 / 21 / 27, nu = 3; rlmpc/mpc/chain_mass/ocp_utils.py).
//
// What this replaces in the reference is the same as engine.cuh (acados SQP + HPIPM, update_nlp's dense KKT solve);
// the mapping is different:
//   * stage_task   one WARP per (sample, stage): the nominal RK4 sweep and its adjoint are shared through shared
//                  memory, every lane then carries ONE tangent direction (a column of [A | B]) through the integrator
//                  in registers; the exact Hessian of pi'F is accumulated as sum_s D_s' G_s D_s over the RK stage
//                  points (D_s: the lanes' tangents of the link geometry, G_s: 3 x 3 link blocks weighted with the
//                  adjoint) -- no second-order sweep per direction.
//   * qp_sample    one WARP per sample: the Riccati recursion on nx x nx blocks held in shared memory, products
//                  register-tiled 3 x CPL per lane, the stage records [A | B | b | q | r] streamed from HBM by TMA bulk
//                  copies one stage ahead of the recursion; primal-dual interior point on the input bounds in the
//                  absolute-step form of engine.cuh (warm start, active-set steps, cold restart).
//   * sens_sample  one WARP per sample: exact-Hessian factorisation, the nu adjoint solves K y = e_u0, the
//                  contractions for the cost parameters Q, R (rank-2 updates per stage, never stored).
//   * param_task   one THREAD per (sample, stage, adjoint right-hand side): the contraction with the dynamic
//                  parameters as one forward-over-reverse sweep through the integrator (ChainModel::param_contraction)
//                  -- dF/dtheta and d2(pi'F)/dw dtheta are never formed.
// Per-sample data is CONTIGUOUS here (not AoSoA): a warp owns a sample, and a stage record is one TMA segment.
#pragma once
#include "chain_model.cuh"
#include "simt.cuh"

namespace rlmpc {

#ifndef CHAIN_USE_DMMA
#define CHAIN_USE_DMMA 1  // dense block products on the FP64 tensor cores (0: register-tiled DFMA versions, kept for comparison)
#endif

MPC_HD constexpr int ev2(int n) { return (n + 1) & ~1; }
MPC_HD double nn(double v) { return (v == v) ? v : 1e300; }  // NaN -> huge, so that maxima keep it

// device pointers and sizes of one call
struct ChainArgs {
  double* it;          // [B][IT]
  double* ws;          // [B][N+1][REC]
  const double* th;    // theta (shared by the batch)
  const double* tab;   // derived tables: sym(Q) [NX*NX] | sym(R) [NU*NU] | x_ss [NX]
  int B;
  int* status;         // acados status per sample
  int* work;           // Work state per sample
  double* cost;        // cost of the last linearisation
  int* counters;       // [0] work counter of the per-sample kernel, [1] samples still active, [2] interior-point iterations
  const double* x0;    // [B, NX] / [B, NU] row-major or null
  const double* u0;
  double* u0_out;
  double* cost_out;
  int* status_out;
  double* dL;          // [B, NTH]
  double* dpi;         // [B, NU, NTH]
  double* res_out;     // [B, 4]
  int last_round, have_solve;
};

template <int NMASS>
struct ChainEngine {
  using Mo = ChainModel<NMASS>;
  static constexpr int NX = Mo::NX, NU = Mo::NU, NW = Mo::NW, NC = Mo::NC, NL = Mo::NL, MI = Mo::M, NPD = Mo::NPD, NTH = Mo::NTH;
  static constexpr int NSP = Mo::NSP, NR = 2 * NU, NK1 = NX + 1, NVEL = Mo::NVEL, NPOS = Mo::NPOS;

  // ---------------- iterate (per sample, contiguous) ----------------
  MPC_HD static int it_x(int N, int k) { (void)N; return k * NX; }
  MPC_HD static int it_u(int N, int k) { return (N + 1) * NX + k * NU; }
  MPC_HD static int it_pi(int N, int k) { return (N + 1) * NX + N * NU + k * NX; }
  MPC_HD static int it_lam(int N, int k) { return (N + 1) * NX + N * NU + N * NX + k * NR; }  // k < N: [lbu(3) ubu(3)]
  MPC_HD static int it_t(int N, int k) { return it_lam(N, 0) + N * NR + k * NR; }
  MPC_HD static int it_rx0(int N) { return it_t(N, 0) + N * NR; }
  MPC_HD static int it_ru0(int N) { return it_rx0(N) + NX; }
  MPC_HD static int it_meta(int N) { return it_ru0(N) + NU; }
  MPC_HD static int it_size(int N) { return ev2(it_meta(N) + 1); }

  // ---------------- stage record (per sample and stage, contiguous; TMA segments start at even offsets) -------------
  // solve phase
  static constexpr int S_M = 0;                      // [A | B | b], NX x NC row-major
  static constexpr int S_G = NX * NC;                // scaled cost gradient [q ; r] (NW)
  static constexpr int SEG_LIN = ev2(S_G + NW);      // TMA segment of the sweeps: [S_M, SEG_LIN)
  static constexpr int S_K = SEG_LIN;                // feedback law [K | kff], NU x (NX+1)
  static constexpr int S_DX = S_K + NU * NK1;        // primal step
  static constexpr int S_DU = S_DX + NX;
  static constexpr int S_C = S_DU + NU;              // scaled stage cost
  static constexpr int S_E = S_C + 1;                // |F(x_k,u_k) - x_{k+1}|_inf
  static constexpr int S_S = S_E + 1;                // stationarity residual of the stage's x_k / u_k rows
  static constexpr int S_END = S_S + 1;
  // sensitivity phase (overlay; A, B stay)
  static constexpr int Z_H = ev2(NX * NC);           // exact Hessian of pi'F wrt w, NW x NW
  static constexpr int SEG_SENS = ev2(Z_H + NW * NW);
  static constexpr int Z_K = SEG_SENS;               // feedback gain K (NU x NX)
  static constexpr int Z_P = ev2(Z_K + NU * NX);     // P_{k+1} (NX x NX)
  static constexpr int Z_PEND = ev2(Z_P + NX * NX);
  static constexpr int Z_Y = Z_PEND;                 // adjoint solution: [r][yx (NX) yu (NU)] then [r][ypi (NX)]
  static constexpr int Z_GT = Z_Y + NU * (NW + NX);  // pi_k' dF/dtheta_dyn (NPD)
  static constexpr int Z_C = Z_GT + NPD;
  static constexpr int Z_E = Z_C + 1;
  static constexpr int Z_S = Z_E + 1;
  static constexpr int Z_END = Z_S + 1;
  static constexpr int REC = ev2(S_END > Z_END ? S_END : Z_END);
  MPC_HD static size_t ws_size(int N) { return (size_t)(N + 1) * REC; }

  // derived tables
  static constexpr int TB_Q = 0, TB_R = NX * NX, TB_XSS = TB_R + NU * NU, TB_SIZE = TB_XSS + NX;

  // ================================================================================================================
  // (sample, stage) task: linearisation (HESS = false) or exact second-order information (HESS = true) of stage k.
  // ================================================================================================================
  // per-lane row of the tangent exchange buffers; with DMMA the rows are the K dimension of m8n8k4 tiles (multiple of 4)
  static constexpr int HKT = (3 * NL + 3) / 4, HMT = (NW + 7) / 8;
  static constexpr int DELW = CHAIN_USE_DMMA ? 4 * HKT : ev2(3 * NL);
  // shared memory of one warp (doubles)
  static constexpr int SM_X0 = 0, SM_XN = SM_X0 + NX, SM_PI = SM_XN + NX, SM_PIM = SM_PI + NX, SM_KK = SM_PIM + NX,
                       SM_LB = SM_KK + NX, SM_KB = SM_LB + NX, SM_XB = SM_KB + NX, SM_XC = SM_XB + NX, SM_U = SM_XC + NX,
                       SM_GV = SM_U + 4, SM_XS = SM_GV + NW, SM_FL = SM_XS + NSP * NX, SM_NU = SM_FL + 3 * NL,
                       SM_DB = SM_NU + 3 * NL, SM_VB = SM_DB + 3 * NL, SM_SB = ev2(SM_VB + 3 * NL),
                       SM_GB = SM_SB + NSP * NL * 9, SM_DEL = ev2(SM_GB + NSP * NL * 6), SM_GDL = SM_DEL + 32 * DELW, SM_OUT = SM_GDL + 32 * DELW,
                       SM_STAGE = ev2(SM_OUT + NW * NC);

  // cost of a stage at (x, u): gradient into g (NW, shared memory), returns the scaled value.  Lanes over rows.
  CH_DEV static double stage_cost(const ChainArgs& a, double s, bool terminal, const double* x, const double* u, double* g,
                                  int lane) {
    double val = 0.0;
    if (lane < NX) {
      double acc = 0.0;
      for (int j = 0; j < NX; ++j) acc += a.tab[TB_Q + lane * NX + j] * (x[j] - a.tab[TB_XSS + j]);
      g[lane] = s * acc;
      val = 0.5 * acc * (x[lane] - a.tab[TB_XSS + lane]);
    } else if (lane < NW) {
      double acc = 0.0;
      if (!terminal) {
        for (int j = 0; j < NU; ++j) acc += a.tab[TB_R + (lane - NX) * NU + j] * u[j];
        val = 0.5 * acc * u[lane - NX];
      }
      g[lane] = s * acc;
    }
    return s * wsum(val);
  }

  template <bool HESS>
  CH_DEV static void stage_task(const ProblemData& pd, const ChainArgs& a, int b, int k, double* S, int lane) {
    const int N = pd.N;
    const double h = pd.mc[0];
    const bool qmode = pd.mode == MODE_Q;
    double* it = a.it + (size_t)b * it_size(N);
    double* rec = a.ws + ((size_t)b * (N + 1) + k) * REC;
    const double* th = a.th;
    double* X0 = S + SM_X0; double* XN = S + SM_XN; double* PI = S + SM_PI; double* PIM = S + SM_PIM; double* KK = S + SM_KK;
    double* LB = S + SM_LB; double* KB = S + SM_KB; double* XB = S + SM_XB; double* XC = S + SM_XC; double* U = S + SM_U;
    double* GV = S + SM_GV; double* XS = S + SM_XS; double* FL = S + SM_FL; double* NUv = S + SM_NU; double* DB = S + SM_DB;
    double* VB = S + SM_VB; double* SB = S + SM_SB; double* GB = S + SM_GB; double* DEL = S + SM_DEL; double* GDL = S + SM_GDL;
    double* OUT = S + SM_OUT;
    constexpr int C_ = HESS ? Z_C : S_C, E_ = HESS ? Z_E : S_E, R_ = HESS ? Z_S : S_S;

    if (lane < NX) {
      X0[lane] = it[it_x(N, k) + lane];
      PIM[lane] = (k > 0) ? it[it_pi(N, k - 1) + lane] : 0.0;
      if (k < N) {
        XN[lane] = it[it_x(N, k + 1) + lane];
        PI[lane] = it[it_pi(N, k) + lane];
      }
    }
    if (lane < NU) U[lane] = (k < N) ? it[it_u(N, k) + lane] : 0.0;
    WSYNC();

    if (k == N) {  // terminal stage: cost only
      const double c = stage_cost(a, pd.scale[N], true, X0, U, GV, lane);
      double sres = 0.0;
      if (lane < NX) {
        rec[S_G + lane] = GV[lane];
        sres = dabs(GV[lane] - PIM[lane]);
        if (!(sres == sres)) sres = 1e300;
      }
      sres = wmax(sres);
      if (lane == 0) { rec[C_] = c; rec[E_] = 0.0; rec[R_] = sres; }
      return;
    }

    // ---------------- nominal forward sweep: stage points XS, link Jacobian blocks SB ----------------
    if (lane < NX) XC[lane] = X0[lane];
    WSYNC();
    for (int sub = 0; sub < Mo::NSUB; ++sub) {
      double accn = 0.0;  // lane c: sum_st b_st k_st[c]
      for (int st = 0; st < 4; ++st) {
        const int s = sub * 4 + st;
        double* xs = XS + s * NX;
        if (lane < NX) xs[lane] = XC[lane] + (st == 0 ? 0.0 : Mo::rk_a(st, h) * KK[lane]);
        WSYNC();
        if (lane < NL) {  // one lane per link: force and Jacobian block
          typename Mo::Link lk;
          Mo::link_at(xs, U, th, lane, lk);
          Mo::link_force(th, lane, lk, FL + 3 * lane);
          Mo::link_S(th, lane, lk, SB + (s * NL + lane) * 9);
        }
        WSYNC();
        if (lane < NX) {
          double kv;
          if (lane < NVEL) kv = xs[NPOS + lane];
          else if (lane < NPOS) kv = U[lane - NVEL];
          else {
            const int m = (lane - NPOS) / 3, j = (lane - NPOS) - 3 * m;
            kv = (j == 2 ? -9.81 : 0.0) + th[Mo::TH_W + 3 * m + j] - FL[3 * m + j] + FL[3 * (m + 1) + j];
          }
          KK[lane] = kv;
          accn += Mo::rk_b(st, h) * kv;
        }
        WSYNC();
      }
      if (lane < NX) XC[lane] += accn;
      WSYNC();
    }
    // XC = F(x_k, u_k)

    // ---------------- nominal adjoint sweep with seed pi_k: link Hessian blocks GB, pi_k' dF/dtheta ----------------
    double gth[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // lane 3 i + j < 3 NL: [D_ij, L_ij, C_ij, part of m_i]; lane < NVEL: [4] = w
    if (HESS) {
      if (lane < NX) LB[lane] = PI[lane];
      WSYNC();
      for (int sub = Mo::NSUB - 1; sub >= 0; --sub) {
        double lacc = 0.0;
        if (lane < NX) XB[lane] = 0.0;
        WSYNC();
        for (int st = 3; st >= 0; --st) {
          const int s = sub * 4 + st;
          const double* xs = XS + s * NX;
          if (lane < NX) KB[lane] = Mo::rk_b(st, h) * LB[lane] + (st < 3 ? Mo::rk_a(st + 1, h) * XB[lane] : 0.0);
          WSYNC();
          if (lane < NL) {
            typename Mo::Link lk;
            double nu[3];
            Mo::link_at(xs, U, th, lane, lk);
            Mo::link_nu(KB, lane, nu);
            const double* Sb = SB + (s * NL + lane) * 9;
            for (int l = 0; l < 3; ++l) {
              NUv[3 * lane + l] = nu[l];
              DB[3 * lane + l] = nu[0] * Sb[l] + nu[1] * Sb[3 + l] + nu[2] * Sb[6 + l];
              VB[3 * lane + l] = nu[l] * th[Mo::TH_C + 3 * lane + l];
            }
            Mo::link_G(th, lane, lk, nu, GB + (s * NL + lane) * 6);
          }
          WSYNC();
          if (lane < 3 * NL) {  // parameter gradient of this stage point, lane = (link i, component j)
            const int i = lane / 3, j = lane - 3 * i;
            typename Mo::Link lk;
            Mo::link_at(xs, U, th, i, lk);
            const double im = lk.im, Dj = th[Mo::TH_D + 3 * i + j], Lj = th[Mo::TH_L + 3 * i + j];
            const double nuj = NUv[lane];
            const double e = (1.0 - Lj * lk.ir) * lk.d[j] * im;
            gth[0] += nuj * e;
            gth[1] -= nuj * Dj * im * lk.d[j] * lk.ir;
            gth[2] += nuj * lk.vl[j];
            gth[3] -= nuj * Dj * e * im;
          }
          if (lane < NVEL) gth[4] += KB[NPOS + lane];
          double xb = 0.0;
          if (lane < NPOS) {
            const int i = lane / 3, l = lane - 3 * i;
            xb = DB[3 * i + l] - (i < MI ? DB[3 * (i + 1) + l] : 0.0);
          } else if (lane < NX) {
            const int m = (lane - NPOS) / 3, j = (lane - NPOS) - 3 * m;
            xb = KB[3 * m + j] + VB[3 * m + j] - VB[3 * (m + 1) + j];
          }
          WSYNC();  // KB, DB, VB of this stage point consumed
          if (lane < NX) XB[lane] = xb;
          lacc += xb;
          WSYNC();
        }
        if (lane < NX) LB[lane] += lacc;
        WSYNC();
      }
    }

    // ---------------- tangent sweep: lane j carries direction e_j of w = [x ; u] ----------------
    double dx[NX], dacc[NX], dp[NX], dkp[NVEL], dka[NVEL];
    double du[3];
#if CHAIN_USE_DMMA
    double hacc[HESS ? HMT : 1][HESS ? HMT : 1][2];  // H tiles of this lane: rows 8 mt + lane / 4, columns 8 nt + 2 (lane % 4) + {0, 1}
    if constexpr (HESS) {
      MPC_UNROLL for (int mt = 0; mt < HMT; ++mt) MPC_UNROLL for (int nt = 0; nt < HMT; ++nt) { hacc[mt][nt][0] = 0.0; hacc[mt][nt][1] = 0.0; }
      for (int e = 3 * NL; e < DELW; ++e) { DEL[lane * DELW + e] = 0.0; GDL[lane * DELW + e] = 0.0; }  // zero padding of the K dimension
    }
#else
    double Hc[HESS ? NW : 1];
    if constexpr (HESS) {
      MPC_UNROLL for (int l = 0; l < NW; ++l) Hc[l] = 0.0;
    }
#endif
    MPC_UNROLL for (int c = 0; c < NX; ++c) dx[c] = (c == lane) ? 1.0 : 0.0;
    MPC_UNROLL for (int j = 0; j < 3; ++j) du[j] = (lane == NX + j) ? 1.0 : 0.0;
    for (int sub = 0; sub < Mo::NSUB; ++sub) {
      MPC_UNROLL for (int c = 0; c < NX; ++c) dacc[c] = 0.0;
      MPC_UNROLL for (int c = 0; c < NVEL; ++c) { dkp[c] = 0.0; dka[c] = 0.0; }
      for (int st = 0; st < 4; ++st) {
        const int s = sub * 4 + st;
        const double aa = (st == 0) ? 0.0 : Mo::rk_a(st, h), bb = Mo::rk_b(st, h);
        // tangent of the stage point: dp = dx + aa dk_prev with dk_prev = [dkp ; du ; dka]
        MPC_UNROLL for (int c = 0; c < NVEL; ++c) dp[c] = dx[c] + aa * dkp[c];
        MPC_UNROLL for (int j = 0; j < 3; ++j) dp[NVEL + j] = dx[NVEL + j] + aa * du[j];
        MPC_UNROLL for (int c = 0; c < NVEL; ++c) dp[NPOS + c] = dx[NPOS + c] + aa * dka[c];
        double dT[3 * NL], gd[HESS ? 3 * NL : 1];
        MPC_UNROLL for (int i = 0; i < NL; ++i) {
          double dd[3], dvl[3];
          Mo::link_tan(dp, du, i, dd, dvl);
          const double* Sb = SB + (s * NL + i) * 9;
          MPC_UNROLL for (int j = 0; j < 3; ++j)
            dT[3 * i + j] = Sb[3 * j] * dd[0] + Sb[3 * j + 1] * dd[1] + Sb[3 * j + 2] * dd[2] + th[Mo::TH_C + 3 * i + j] * dvl[j];
          if constexpr (HESS) {
            Mo::sym3_mul(GB + (s * NL + i) * 6, dd, gd + 3 * i);
            MPC_UNROLL for (int j = 0; j < 3; ++j) {
              DEL[lane * DELW + 3 * i + j] = dd[j];
#if CHAIN_USE_DMMA
              GDL[lane * DELW + 3 * i + j] = gd[3 * i + j];
#endif
            }
          }
        }
        // k of this stage: [dp_vel ; du ; -dT_m + dT_{m+1}]
        MPC_UNROLL for (int c = 0; c < NVEL; ++c) {
          dkp[c] = dp[NPOS + c];
          dka[c] = -dT[c] + dT[3 + c];
          dacc[c] += bb * dkp[c];
          dacc[NPOS + c] += bb * dka[c];
        }
        MPC_UNROLL for (int j = 0; j < 3; ++j) dacc[NVEL + j] += bb * du[j];
        if constexpr (HESS) {
          WSYNC();
#if CHAIN_USE_DMMA
          // H += D' (G D) over the link geometry of this stage point, D = [tangents of the lanes] (3 NL x NW): a dense
          // (NW x 3NL) (3NL x NW) contraction on the FP64 tensor cores, HMT^2 HKT mma.sync per stage point
          MPC_UNROLL for (int kt = 0; kt < HKT; ++kt) {
            double af[HMT], bf[HMT];
            MPC_UNROLL for (int mt = 0; mt < HMT; ++mt) {
              af[mt] = DEL[(8 * mt + (lane >> 2)) * DELW + 4 * kt + (lane & 3)];
              bf[mt] = GDL[(8 * mt + (lane >> 2)) * DELW + 4 * kt + (lane & 3)];
            }
            MPC_UNROLL for (int mt = 0; mt < HMT; ++mt) MPC_UNROLL for (int nt = 0; nt < HMT; ++nt) dmma_8x8x4(af[mt], bf[nt], hacc[mt][nt][0], hacc[mt][nt][1]);
          }
#else
          MPC_UNROLL for (int l = 0; l < NW; ++l) {  // H[l][j] += sum_i Delta_i^(l)' G_i Delta_i^(j)
            const double* dl = DEL + l * DELW;
            double acc = 0.0;
            MPC_UNROLL for (int e = 0; e < 3 * NL; ++e) acc += dl[e] * gd[e];
            Hc[l] += acc;
          }
#endif
          WSYNC();
        }
      }
      MPC_UNROLL for (int c = 0; c < NX; ++c) dx[c] += dacc[c];
    }
    // lane j < NW: dx = column j of [A | B]

    // ---------------- outputs ----------------
    const double cst = stage_cost(a, pd.scale[k], false, X0, U, GV, lane);
    double eq = 0.0, sres = 0.0;
    if (lane < NX) {
      const double bb = XC[lane] - XN[lane];
      OUT[lane * NC + NW] = bb;
      eq = dabs(bb);
      if (!(bb == bb)) eq = 1e300;
    }
    if (lane < NW) {
      MPC_UNROLL for (int c = 0; c < NX; ++c) OUT[c * NC + lane] = dx[c];
      // stationarity of the rows of w_k at the current multipliers:  g + [A B]' pi_k - pi_{k-1} (x) / - lam_l + lam_u (u)
      double r = GV[lane];
      MPC_UNROLL for (int c = 0; c < NX; ++c) r += dx[c] * PI[c];
      if (lane < NX) {
        r -= PIM[lane];
        if (k == 0) {  // x_0 is fixed: this is its multiplier, not a residual
          if (HESS) it[it_rx0(N) + lane] = r;
          r = 0.0;
        }
      } else {
        const int q = lane - NX;
        const bool rows = !(k == 0 && qmode);
        if (rows) r += -it[it_lam(N, k) + q] + it[it_lam(N, k) + NU + q];
        if (k == 0 && HESS) it[it_ru0(N) + q] = r;
        if (!rows) r = 0.0;  // u_0 clamped (Q-mode): multiplier, not a residual
      }
      sres = dabs(r);
      if (!(r == r)) sres = 1e300;
    }
    eq = wmax(eq);
    sres = wmax(sres);
    WSYNC();  // OUT is complete
    for (int e = lane; e < NX * NC; e += 32) rec[S_M + e] = OUT[e];
    if (lane < NW) rec[S_G + lane] = GV[lane];
    if (lane == 0) { rec[C_] = cst; rec[E_] = eq; rec[R_] = sres; }
    if constexpr (HESS) {
      WSYNC();
#if CHAIN_USE_DMMA
      MPC_UNROLL for (int mt = 0; mt < HMT; ++mt) MPC_UNROLL for (int nt = 0; nt < HMT; ++nt) MPC_UNROLL for (int h_ = 0; h_ < 2; ++h_) {
        const int m = 8 * mt + (lane >> 2), n = 8 * nt + 2 * (lane & 3) + h_;
        if (m < NW && n < NW) OUT[m * NW + n] = hacc[mt][nt][h_];
      }
#else
      if (lane < NW) {
        MPC_UNROLL for (int l = 0; l < NW; ++l) OUT[l * NW + lane] = Hc[l];
      }
#endif
      if (lane < 3 * NL) DB[lane] = gth[3];
      WSYNC();
      for (int e = lane; e < NW * NW; e += 32) rec[Z_H + e] = OUT[e];
      if (lane < 3 * NL) {
        rec[Z_GT + Mo::PD_D + lane] = gth[0];
        rec[Z_GT + Mo::PD_L + lane] = gth[1];
        rec[Z_GT + Mo::PD_C + lane] = gth[2];
      }
      if (lane < NL) rec[Z_GT + Mo::PD_M + lane] = DB[3 * lane] + DB[3 * lane + 1] + DB[3 * lane + 2];
      if (lane < NVEL) rec[Z_GT + Mo::PD_W + lane] = gth[4];
    }
  }

  // ================================================================================================================
  // Register-tiled products of the Riccati recursion.  Output tile of a lane: rows 3 rg .. 3 rg + 2, columns
  // cg, cg + CG, cg + 2 CG, ...  (ROWS a multiple of 3; RG = ROWS / 3 row groups, CG = 32 / RG column groups).
  // ================================================================================================================
  template <int ROWS>
  struct Tile {
    static constexpr int RG = ROWS / 3, CG = 32 / RG, CPL = (NC + CG - 1) / CG;
    static_assert(ROWS % 3 == 0 && RG <= 32 && CG >= 1 && CPL * CG >= NC, "tile geometry");
  };
  // acc = Pm (NX x NX) * Mk (NX x NC)
  CH_DEV static void gemm_PM(const double* Pm, const double* Mk, int lane, double (&acc)[3][Tile<NX>::CPL]) {
    using T = Tile<NX>;
    const int rg = lane / T::CG, cg = lane - rg * T::CG;
    const int r0 = (rg < T::RG ? rg : 0) * 3;
    MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) MPC_UNROLL for (int q = 0; q < T::CPL; ++q) acc[a_][q] = 0.0;
    for (int l = 0; l < NX; ++l) {
      double pa[3], mb[T::CPL];
      MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) pa[a_] = Pm[(r0 + a_) * NX + l];
      MPC_UNROLL for (int q = 0; q < T::CPL; ++q) {
        const int c = cg + T::CG * q;
        mb[q] = Mk[l * NC + (c < NC ? c : 0)];
      }
      MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) MPC_UNROLL for (int q = 0; q < T::CPL; ++q) acc[a_][q] += pa[a_] * mb[q];
    }
  }
  // acc = Mk[:, 0:NW]' (NW x NX) * Tm (NX x NC)
  CH_DEV static void gemm_MtT(const double* Mk, const double* Tm, int lane, double (&acc)[3][Tile<NW>::CPL]) {
    using T = Tile<NW>;
    const int rg = lane / T::CG, cg = lane - rg * T::CG;
    const int r0 = (rg < T::RG ? rg : 0) * 3;
    MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) MPC_UNROLL for (int q = 0; q < T::CPL; ++q) acc[a_][q] = 0.0;
    for (int l = 0; l < NX; ++l) {
      double ma[3], tb[T::CPL];
      MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) ma[a_] = Mk[l * NC + r0 + a_];
      MPC_UNROLL for (int q = 0; q < T::CPL; ++q) {
        const int c = cg + T::CG * q;
        tb[q] = Tm[l * NC + (c < NC ? c : 0)];
      }
      MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) MPC_UNROLL for (int q = 0; q < T::CPL; ++q) acc[a_][q] += ma[a_] * tb[q];
    }
  }

  // One backward Riccati step in shared memory.
  //   in : Pm (NX x NX, symmetric), PV (p, NX), Mk = [A | B | b] (b ignored when `affine` is false)
  //        hess(i, c): Hessian entry of the stage incl. barrier terms (i, c < NW), grad(i): gradient entry
  //   out: Pm, PV of this stage; KK = [K | kff] (NU x NK1); returns false if the reduced Hessian is not positive definite (need_pd;
  //        the QP solves) resp. not invertible (the sensitivities: update_nlp solves its KKT system with a general sparse LU,
  //        nlp.py:1413-1424, which needs it nonsingular, not definite)
  // Tm (NX x NC), HU (NU x NC) are scratch.  G0I (optional, 9): inverse of the reduced Hessian G.
  template <class HF, class GF>
  CH_DEV static bool riccati_step(double* Pm, double* PV, double* Tm, double* HU, double* KK, const double* Mk, bool affine,
                                  bool ufixed, bool need_pd, HF hess, GF grad, double* G0I, int lane) {
#if CHAIN_USE_DMMA
    // Both products on the FP64 tensor cores (mma.sync m8n8k4): per stage 2 x MT x NT x KT instructions instead of
    // ~700 DFMA + ~330 LDS per lane -- ncu on the FMA version: shared-memory pipe 37-46 % busy, FP64 pipe 17 %, i.e.
    // operand delivery, not arithmetic, was the limiter (profiles/r02_summary.md).  Operands are zero-padded by the
    // guards to multiples of the 8 x 8 x 4 tile.
    constexpr int MT1 = (NX + 7) / 8, MT2 = (NW + 7) / 8, NT = (NC + 7) / 8, KT = (NX + 3) / 4;
    const int fr = lane >> 2, fk = lane & 3;
    {  // T = P [A | B | b] + [0 | 0 | p]
      double acc[MT1][NT][2];
      MPC_UNROLL for (int mt = 0; mt < MT1; ++mt) MPC_UNROLL for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
      MPC_UNROLL for (int kt = 0; kt < KT; ++kt) {
        double af[MT1], bf[NT];
        const int kk = 4 * kt + fk;
        MPC_UNROLL for (int mt = 0; mt < MT1; ++mt) {
          const int i = 8 * mt + fr;
          af[mt] = (i < NX && kk < NX) ? Pm[i * NX + kk] : 0.0;
        }
        MPC_UNROLL for (int nt = 0; nt < NT; ++nt) {
          const int c = 8 * nt + fr;
          bf[nt] = (kk < NX && c < NC) ? Mk[kk * NC + c] : 0.0;
        }
        MPC_UNROLL for (int mt = 0; mt < MT1; ++mt) MPC_UNROLL for (int nt = 0; nt < NT; ++nt) dmma_8x8x4(af[mt], bf[nt], acc[mt][nt][0], acc[mt][nt][1]);
      }
      WSYNC();  // (every lane has read P and the previous T)
      MPC_UNROLL for (int mt = 0; mt < MT1; ++mt) MPC_UNROLL for (int nt = 0; nt < NT; ++nt) MPC_UNROLL for (int h = 0; h < 2; ++h) {
        const int i = 8 * mt + fr, c = 8 * nt + 2 * fk + h;
        if (i < NX && c < NC) Tm[i * NC + c] = (c == NW) ? (affine ? acc[mt][nt][h] + PV[i] : 0.0) : acc[mt][nt][h];
      }
    }
    WSYNC();
    {  // [A B]' T + stage Hessian / gradient  ->  P (x rows, x columns), p (x rows, last column), HU (u rows)
      double acc[MT2][NT][2];
      MPC_UNROLL for (int mt = 0; mt < MT2; ++mt) MPC_UNROLL for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
      MPC_UNROLL for (int kt = 0; kt < KT; ++kt) {
        double af[MT2], bf[NT];
        const int kk = 4 * kt + fk;
        MPC_UNROLL for (int mt = 0; mt < MT2; ++mt) {
          const int i = 8 * mt + fr;
          af[mt] = (kk < NX && i < NW) ? Mk[kk * NC + i] : 0.0;
        }
        MPC_UNROLL for (int nt = 0; nt < NT; ++nt) {
          const int c = 8 * nt + fr;
          bf[nt] = (kk < NX && c < NC) ? Tm[kk * NC + c] : 0.0;
        }
        MPC_UNROLL for (int mt = 0; mt < MT2; ++mt) MPC_UNROLL for (int nt = 0; nt < NT; ++nt) dmma_8x8x4(af[mt], bf[nt], acc[mt][nt][0], acc[mt][nt][1]);
      }
      MPC_UNROLL for (int mt = 0; mt < MT2; ++mt) MPC_UNROLL for (int nt = 0; nt < NT; ++nt) MPC_UNROLL for (int h = 0; h < 2; ++h) {
        const int i = 8 * mt + fr, c = 8 * nt + 2 * fk + h;
        if (i < NW && c < NC) {
          const double v = acc[mt][nt][h] + (c < NW ? hess(i, c) : grad(i));
          if (i < NX) {
            if (c < NX) Pm[i * NX + c] = v;
            else if (c == NW) PV[i] = v;
          } else {
            HU[(i - NX) * NC + c] = v;
          }
        }
      }
    }
    WSYNC();
#else
    {  // T = P [A | B | b] + [0 | 0 | p]
      double acc[3][Tile<NX>::CPL];
      gemm_PM(Pm, Mk, lane, acc);
      using T = Tile<NX>;
      const int rg = lane / T::CG, cg = lane - rg * T::CG;
      if (rg < T::RG) {
        MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) MPC_UNROLL for (int q = 0; q < T::CPL; ++q) {
          const int i = 3 * rg + a_, c = cg + T::CG * q;
          if (c < NC) Tm[i * NC + c] = (c == NW) ? (affine ? acc[a_][q] + PV[i] : 0.0) : acc[a_][q];
        }
      }
    }
    WSYNC();
    {  // [A B]' T + stage Hessian / gradient  ->  P (x rows, x columns), p (x rows, last column), HU (u rows)
      double acc[3][Tile<NW>::CPL];
      gemm_MtT(Mk, Tm, lane, acc);
      using T = Tile<NW>;
      const int rg = lane / T::CG, cg = lane - rg * T::CG;
      if (rg < T::RG) {
        MPC_UNROLL for (int a_ = 0; a_ < 3; ++a_) MPC_UNROLL for (int q = 0; q < T::CPL; ++q) {
          const int i = 3 * rg + a_, c = cg + T::CG * q;
          if (c < NC) {
            const double v = acc[a_][q] + (c < NW ? hess(i, c) : grad(i));
            if (i < NX) {
              if (c < NX) Pm[i * NX + c] = v;
              else if (c == NW) PV[i] = v;
            } else {
              HU[(i - NX) * NC + c] = v;
            }
          }
        }
      }
    }
    WSYNC();
#endif
    bool ok = true;
    if (ufixed) {
      if (lane < NK1) MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) KK[a_ * NK1 + lane] = 0.0;
      if (G0I && lane < NU * NU) G0I[lane] = 0.0;
      // P, p of a stage whose input is fixed are never used (only stage 0 of a Q-mode solve)
      WSYNC();
      return true;
    }
    {  // K = -G^{-1} H, kff = -G^{-1} gv : every lane factorises the NU x NU block, lane c solves column c
      double G[NU * NU], rhs[NU], dinv[NU];
      MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) MPC_UNROLL for (int b_ = 0; b_ < NU; ++b_) G[a_ * NU + b_] = HU[a_ * NC + NX + b_];
      // LDL' (reciprocal pivots: the diagonal scalings of the solves multiply)
      MPC_UNROLL for (int j = 0; j < NU; ++j) {
        double d = G[j * NU + j];
        MPC_UNROLL for (int p = 0; p < j; ++p) d -= G[j * NU + p] * G[j * NU + p] * G[p * NU + p];
        if (need_pd ? !(d > 0.0) : !(d > 0.0 || d < 0.0)) ok = false;
        G[j * NU + j] = d;
        const double inv = 1.0 / d;
        dinv[j] = inv;
        MPC_UNROLL for (int i = j + 1; i < NU; ++i) {
          double v = G[i * NU + j];
          MPC_UNROLL for (int p = 0; p < j; ++p) v -= G[i * NU + p] * G[j * NU + p] * G[p * NU + p];
          G[i * NU + j] = v * inv;
        }
      }
      auto solve = [&](double* r) {
        MPC_UNROLL for (int i = 0; i < NU; ++i) MPC_UNROLL for (int p = 0; p < i; ++p) r[i] -= G[i * NU + p] * r[p];
        MPC_UNROLL for (int i = 0; i < NU; ++i) r[i] *= dinv[i];
        MPC_UNROLL for (int i = NU - 1; i >= 0; --i) MPC_UNROLL for (int p = i + 1; p < NU; ++p) r[i] -= G[p * NU + i] * r[p];
      };
      if (lane < NK1) {
        const int c = (lane < NX) ? lane : NW;  // column of HU: H (x columns) or the gradient
        MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) rhs[a_] = -HU[a_ * NC + c];
        solve(rhs);
        MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) KK[a_ * NK1 + lane] = rhs[a_];
      } else if (G0I && lane < NK1 + NU) {
        const int c = lane - NK1;
        MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) rhs[a_] = (a_ == c) ? 1.0 : 0.0;
        solve(rhs);
        MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) G0I[a_ * NU + c] = rhs[a_];
      }
    }
    WSYNC();
    // P <- Mxx + H' K (upper triangle computed, mirrored: P stays bitwise symmetric);  p <- g_x + H' kff
    for (int e = lane; e < NX * NX; e += 32) {
      const int i = e / NX, j = e - i * NX;
      if (i <= j) {
        double v = Pm[e];
        MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) v += HU[a_ * NC + i] * KK[a_ * NK1 + j];
        Pm[e] = v;
        Pm[j * NX + i] = v;
      }
    }
    if (lane < NX && affine) {
      double v = PV[lane];
      MPC_UNROLL for (int a_ = 0; a_ < NU; ++a_) v += HU[a_ * NC + lane] * KK[a_ * NK1 + NX];
      PV[lane] = v;
    }
    WSYNC();
    return ok;
  }

  // ================================================================================================================
  // Per-sample solve: convergence test, interior-point QP, step.  One warp.
  // ================================================================================================================
  MPC_HD static int qp_smem_doubles(int N) {
    return NX * NX + NX + NX * NC + NU * NC + NU * NK1 + 2 * NX + 2 * N * NR + 4 * N * NU + 2 * SEG_LIN + 2 + 2;
  }
  // the two TMA slots and their mbarriers sit at the end of the warp's shared memory; initialised ONCE per warp
  CH_DEV static void qp_feed_init(StageFeed& feed, double* S, int N, int lane) {
    double* B0 = S + ev2(NX * NX + NX + NX * NC + NU * NC + NU * NK1 + 2 * NX + 2 * N * NR + 4 * N * NU);
    feed.init(B0, B0 + SEG_LIN, reinterpret_cast<uint64_t*>(B0 + 2 * SEG_LIN), lane);
  }
  enum Res : int { R_CONVERGED = 0, R_STEPPED = 1, R_MAXITER = 2, R_FAILED = 3, R_NAN = 4, R_TESTONLY = 5 };

  CH_DEV static int qp_sample(const ProblemData& pd, const ChainArgs& a, int b, double* S, StageFeed& feed, int lane, int* ipm_iters) {
    const int N = pd.N;
    const bool qmode = pd.mode == MODE_Q;
    double* it = a.it + (size_t)b * it_size(N);
    double* ws = a.ws + (size_t)b * ws_size(N);
    const double* tab = a.tab;
    double* Pm = S; double* PV = Pm + NX * NX; double* Tm = PV + NX; double* HU = Tm + NX * NC; double* KK = HU + NU * NC;
    double* DXA = KK + NU * NK1; double* DXB = DXA + NX; double* LAM = DXB + NX; double* TT = LAM + N * NR; double* UU = TT + N * NR;
    double* DU = UU + N * NU; double* RD = DU + N * NU; double* RG = RD + N * NU; double* B0 = RG + N * NU;
    B0 = S + ev2((int)(B0 - S));
    (void)B0;  // the two slots behind it belong to `feed` (qp_feed_init)

    // ---- rows and per-stage residuals of the linearisation ----
    for (int e = lane; e < N * NR; e += 32) { LAM[e] = it[it_lam(N, 0) + e]; TT[e] = it[it_t(N, 0) + e]; }
    for (int e = lane; e < N * NU; e += 32) UU[e] = it[it_u(N, 0) + e];
    WSYNC();
    double stat = 0.0, eq = 0.0, ineq = 0.0, comp = 0.0, cost = 0.0;
    for (int k = lane; k <= N; k += 32) {
      const double* rec = ws + (size_t)k * REC;
      cost += rec[S_C];
      eq = dmax(eq, rec[S_E]);
      stat = dmax(stat, rec[S_S]);
      if (k < N && !(k == 0 && qmode)) {
        MPC_UNROLL for (int q = 0; q < NU; ++q) {
          const double u = UU[k * NU + q];
          const double ll = LAM[k * NR + q], tl = TT[k * NR + q], lu = LAM[k * NR + NU + q], tu = TT[k * NR + NU + q];
          ineq = dmax(ineq, dmax(nn(dabs(pd.lbu[q] - u + tl)), nn(dabs(u - pd.ubu[q] + tu))));
          comp = dmax(comp, dmax(nn(dabs(ll * tl - pd.tau)), nn(dabs(lu * tu - pd.tau))));
        }
      }
    }
    cost = wsum(cost); stat = wmax(stat); eq = wmax(eq); ineq = wmax(ineq); comp = wmax(comp);
    if (lane == 0) a.cost[b] = cost;
    bool warm = pd.warm_ipm && it[it_meta(N)] > 0.5;
    const double rmax = dmax(dmax(stat, eq), dmax(ineq, comp));
    if (!(rmax < 1e299) || !(cost == cost)) return R_NAN;
    if (rmax < pd.tol && (a.last_round || !warm || comp <= 0.05 * pd.tau)) return R_CONVERGED;
    if (a.last_round) return R_TESTONLY;

    // ---- interior point ----
    const double m_rows = (double)(NR * (qmode ? N - 1 : N));
    const int k_first = qmode ? 1 : 0;  // first stage with a free input
    auto init_rows = [&](bool w) -> double {
      double mu = 0.0;
      for (int k = lane; k < N; k += 32) {
        MPC_UNROLL for (int q = 0; q < NU; ++q) {
          const double range = pd.ubu[q] - pd.lbu[q];
          MPC_UNROLL for (int side = 0; side < 2; ++side) {
            const int r = k * NR + side * NU + q;
            if (k < k_first) { LAM[r] = 0.0; TT[r] = 0.0; continue; }
            if (w) {
              TT[r] = dmax(TT[r], 1e-10 * range);
              LAM[r] = dmax(LAM[r], 1e-14);
            } else {
              const double d = side ? pd.ubu[q] - UU[k * NU + q] : UU[k * NU + q] - pd.lbu[q];
              TT[r] = dmax(d, 1e-2 * range);
              LAM[r] = pd.mu0 / TT[r];
            }
            mu += LAM[r] * TT[r];
          }
        }
      }
      WSYNC();
      return wsum(mu);
    };
    double mu = init_rows(warm) / m_rows;
    double alpha = 0.0, sigma = warm ? pd.sigma_min : pd.sigma0;
    int iters = 0, warm_iters = 0, as_iters = 0;
    bool converged = false, failed = false, minstep = false;
    constexpr int WARM_LIMIT = 6;
    for (int j = 0; j < pd.max_ipm && !converged; ++j) {
      ++iters;
      if (warm && (warm_iters >= WARM_LIMIT + as_iters || (warm_iters > as_iters && alpha > 0.0 && alpha < 0.05))) {
        warm = false;
        mu = init_rows(false) / m_rows;
        alpha = 0.0;
        sigma = pd.sigma0;
      }
      if (warm) ++warm_iters;
      const double target = dmax(sigma * mu, pd.tau);
      // ---- rows: barrier terms of the input rows ----
      for (int e = lane; e < N * NU; e += 32) {
        const int k = e / NU, q = e - k * NU;
        double rd = 0.0, rg = 0.0;
        if (k >= k_first) {
          const double u = UU[e];
          const double ll = LAM[k * NR + q], tl = TT[k * NR + q], lu = LAM[k * NR + NU + q], tu = TT[k * NR + NU + q];
          const double cl = ll / tl, al = target / tl + ll, cu = lu / tu, au = target / tu + lu;
          rd = cl + cu;
          rg = -(al - cl * (u - pd.lbu[q])) + (au - cu * (pd.ubu[q] - u));
        }
        RD[e] = rd;
        RG[e] = rg;
      }
      // ---- backward sweep ----
      {
        const double sN = pd.scale[N];
        for (int e = lane; e < NX * NX; e += 32) Pm[e] = sN * tab[TB_Q + e];
        if (lane < NX) PV[lane] = ws[(size_t)N * REC + S_G + lane];
      }
      WSYNC();
      bool ok_all = true;
      feed.issue((N - 1) & 1, lane, ws + (size_t)(N - 1) * REC, 0, SEG_LIN);
      for (int k = N - 1; k >= k_first; --k) {
        if (k > k_first) feed.issue((k - 1) & 1, lane, ws + (size_t)(k - 1) * REC, 0, SEG_LIN);
        const double* Mk = feed.wait(k & 1);
        const double* gk = Mk + S_G;
        const double s = pd.scale[k];
        const double* rd = RD + k * NU;
        const double* rg = RG + k * NU;
        auto hess = [&](int i, int c) -> double {
          if (i < NX) return (c < NX) ? s * tab[TB_Q + i * NX + c] : 0.0;
          if (c < NX) return 0.0;
          return s * tab[TB_R + (i - NX) * NU + (c - NX)] + (i == c ? rd[i - NX] : 0.0);
        };
        auto grad = [&](int i) -> double { return gk[i] + (i >= NX ? rg[i - NX] : 0.0); };
        const bool ok = riccati_step(Pm, PV, Tm, HU, KK, Mk, true, false, /*need_pd=*/true, hess, grad, nullptr, lane);
        ok_all = ok_all && ok;
        for (int e = lane; e < NU * NK1; e += 32) ws[(size_t)k * REC + S_K + e] = KK[e];
        WSYNC();  // slot k&1 is free again
      }
      if (!ok_all) failed = true;
      // ---- forward sweep ----
      if (lane < NX) DXA[lane] = 0.0;
      if (qmode && lane < NU) DU[lane] = 0.0;
      WSYNC();
      double* dxc = DXA;
      double* dxn = DXB;
      if (lane < NX) ws[S_DX + lane] = 0.0;
      feed.issue(0, lane, ws, 0, SEG_LIN);
      double kreg[(NU * NK1 + 31) / 32];
      auto kload = [&](int k) {
        MPC_UNROLL for (int q = 0; q < (NU * NK1 + 31) / 32; ++q) {
          const int e = lane + 32 * q;
          kreg[q] = (e < NU * NK1 && k >= k_first) ? ws[(size_t)k * REC + S_K + e] : 0.0;
        }
      };
      kload(0);
      for (int k = 0; k < N; ++k) {
        if (k + 1 < N) feed.issue((k + 1) & 1, lane, ws + (size_t)(k + 1) * REC, 0, SEG_LIN);
        MPC_UNROLL for (int q = 0; q < (NU * NK1 + 31) / 32; ++q) {
          const int e = lane + 32 * q;
          if (e < NU * NK1) KK[e] = kreg[q];
        }
        const double* Mk = feed.wait(k & 1);  // (barrier on the host emulation; on the device KK needs the explicit one)
        WSYNC();
        if (k + 1 < N) kload(k + 1);
        if (lane < NU) {
          double v = 0.0;
          if (k >= k_first) {
            v = KK[lane * NK1 + NX];
            for (int l = 0; l < NX; ++l) v += KK[lane * NK1 + l] * dxc[l];
          }
          DU[k * NU + lane] = v;
          ws[(size_t)k * REC + S_DU + lane] = v;
        }
        WSYNC();
        if (lane < NX) {
          double v = Mk[lane * NC + NW];
          for (int l = 0; l < NX; ++l) v += Mk[lane * NC + l] * dxc[l];
          MPC_UNROLL for (int q = 0; q < NU; ++q) v += Mk[lane * NC + NX + q] * DU[k * NU + q];
          dxn[lane] = v;
          ws[(size_t)(k + 1) * REC + S_DX + lane] = v;
        }
        WSYNC();
        double* tmp = dxc; dxc = dxn; dxn = tmp;
      }
      // ---- rows: new slacks / multipliers, step-length statistics ----
      double amax = 1e300, s0 = 0.0, s1 = 0.0, s2 = 0.0, cmax = 0.0;
      bool nanstep = false;
      auto row_hat = [&](int k, int q, int side, double& lh, double& th_) {
        const int r = k * NR + side * NU + q;
        const double u = UU[k * NU + q], dq = DU[k * NU + q];
        const double c = LAM[r] / TT[r], aa = target / TT[r] + LAM[r];
        // (u - lb) + du, NOT (u + du) - lb: the same rounded distance as in the condensed gradient
        const double d = side ? (pd.ubu[q] - u) - dq : (u - pd.lbu[q]) + dq;
        th_ = d;
        lh = aa - c * d;
      };
      for (int e = lane; e < N * NU; e += 32) {
        const int k = e / NU, q = e - k * NU;
        if (k < k_first) continue;
        MPC_UNROLL for (int side = 0; side < 2; ++side) {
          double lh, th_;
          row_hat(k, q, side, lh, th_);
          const int r = k * NR + side * NU + q;
          const double dt = th_ - TT[r], dl = lh - LAM[r];
          if (dt < 0.0) amax = dmin(amax, -TT[r] / dt);
          if (dl < 0.0) amax = dmin(amax, -LAM[r] / dl);
          s0 += LAM[r] * TT[r];
          s1 += LAM[r] * dt + TT[r] * dl;
          s2 += dl * dt;
          cmax = dmax(cmax, dabs(dl * dt));
          if (!(lh == lh) || !(th_ == th_)) nanstep = true;
        }
      }
      nanstep = wany(nanstep);
      amax = wmin(amax); s0 = wsum(s0); s1 = wsum(s1); s2 = wsum(s2); cmax = wmax(cmax);
      if ((failed || nanstep) && warm) {  // a warm start went wrong numerically: start over cold
        warm = false; failed = false;
        mu = init_rows(false) / m_rows;
        alpha = 0.0; sigma = pd.sigma0;
        continue;
      }
      if (failed || nanstep) { failed = true; break; }
      if (warm && amax < 1.0 / 0.995 && as_iters < (int)pd.as_steps) {
        // infeasible Newton step of a warm start: full step + projection (active-set step) instead of a short step
        ++as_iters;
        double m2 = 0.0;
        for (int e = lane; e < N * NU; e += 32) {
          const int k = e / NU, q = e - k * NU;
          if (k < k_first) continue;
          const double eps_t = 1e-9 * (pd.ubu[q] - pd.lbu[q]);
          MPC_UNROLL for (int side = 0; side < 2; ++side) {
            double lh, th_;
            row_hat(k, q, side, lh, th_);
            const int r = k * NR + side * NU + q;
            if (!(th_ > eps_t)) { LAM[r] = dmax(dmax(lh, LAM[r]), 1e-3); TT[r] = dmin(eps_t, pd.tau / LAM[r]); }
            else if (!(lh > 0.0)) { TT[r] = dmax(th_, AS_RELEASE * (pd.ubu[q] - pd.lbu[q])); LAM[r] = pd.tau / TT[r]; }
            else { TT[r] = th_; LAM[r] = lh; }
            m2 += LAM[r] * TT[r];
          }
        }
        WSYNC();
        mu = wsum(m2) / m_rows;
        alpha = 0.0; sigma = pd.sigma_min;
        continue;
      }
      alpha = (amax >= 1.0 / 0.995) ? 1.0 : 0.995 * amax;
      if (!warm && alpha < 1e-9) { minstep = true; break; }
      const double mu_new = (s0 + alpha * s1 + alpha * alpha * s2) / m_rows;
      if (target <= pd.tau && alpha == 1.0 && cmax <= dmin(pd.comp_accept * pd.tau, 0.1 * pd.tol)) converged = true;
      // damped update of (lam, t)
      for (int e = lane; e < N * NU; e += 32) {
        const int k = e / NU, q = e - k * NU;
        if (k < k_first) continue;
        MPC_UNROLL for (int side = 0; side < 2; ++side) {
          double lh, th_;
          row_hat(k, q, side, lh, th_);
          const int r = k * NR + side * NU + q;
          LAM[r] += alpha * (lh - LAM[r]);
          TT[r] += alpha * (th_ - TT[r]);
        }
      }
      WSYNC();
      const double rr = 1.0 - alpha;
      sigma = dmin(0.8, dmax(pd.sigma_min, rr * rr * 4.0 + pd.sigma_min));
      mu = mu_new;
    }
    *ipm_iters = iters;
    if (failed || minstep) return R_FAILED;

    // ---- the step: w += ap dw, (lam, t) as updated, pi from the costate recursion of the QP ----
    const double ap = (converged ? 1.0 : alpha) * pd.step_length;
    for (int e = lane; e < N * NR; e += 32) { it[it_lam(N, 0) + e] = LAM[e]; it[it_t(N, 0) + e] = TT[e]; }
    for (int e = lane; e < N * NU; e += 32) it[it_u(N, 0) + e] = UU[e] + ap * DU[e];
    // pi_{k-1} = q_k + s_k Q dx_k + A_k' pi_k  (k = N: terminal cost), x_k += ap dx_k
    double* pic = DXA;
    double* pin = DXB;
    WSYNC();
    if (N >= 2) feed.issue((N - 1) & 1, lane, ws + (size_t)(N - 1) * REC, 0, SEG_LIN);
    for (int k = N; k >= 1; --k) {
      const double* rec = ws + (size_t)k * REC;
      const double* Mk = nullptr;
      if (k < N) {
        if (k > 1) feed.issue((k - 1) & 1, lane, ws + (size_t)(k - 1) * REC, 0, SEG_LIN);
        Mk = feed.wait(k & 1);
      }
      if (lane < NX) {
        const double s = pd.scale[k];
        double v = (k < N) ? Mk[S_G + lane] : rec[S_G + lane];
        double hv = 0.0;
        for (int l = 0; l < NX; ++l) hv += tab[TB_Q + lane * NX + l] * rec[S_DX + l];
        v += s * hv;
        if (k < N)
          for (int l = 0; l < NX; ++l) v += Mk[l * NC + lane] * pic[l];
        pin[lane] = v;
        it[it_pi(N, k - 1) + lane] = v;
        it[it_x(N, k) + lane] += ap * rec[S_DX + lane];
      }
      WSYNC();
      double* tmp = pic; pic = pin; pin = tmp;
    }
    if (lane == 0) it[it_meta(N)] = 1.0;
    return converged ? R_STEPPED : R_MAXITER;
  }

  // ================================================================================================================
  // Per-sample sensitivities (update_nlp, nlp.py:1341-1563): residuals, dL/dtheta, exact-Hessian factorisation, the
  // NU adjoint solves and the cost-parameter columns of dpi/dtheta.  One warp.  Writes y per stage for param_task.
  // ================================================================================================================
  static constexpr int SBUF = SEG_SENS > (ev2(NX * NC) + (Z_PEND - Z_K)) ? SEG_SENS : (ev2(NX * NC) + (Z_PEND - Z_K));
  static constexpr int NQE = (NX * NX + 31) / 32;  // entries of a NX x NX matrix per lane
  MPC_HD static int sens_smem_doubles(int N) {
    return NX * NX + NX + NX * NC + NU * NC + NU * NK1 + (N + 1) * NX + N * NU + 2 * N * NR + 2 * NU * NX + NU * NU + NU * NU + 2 * SBUF + 4;
  }

  CH_DEV static void sens_feed_init(StageFeed& feed, double* S, int N, int lane) {
    double* B0 = S + ev2(NX * NX + NX + NX * NC + NU * NC + NU * NK1 + (N + 1) * NX + N * NU + 2 * N * NR + 2 * NU * NX + 2 * NU * NU);
    feed.init(B0, B0 + SBUF, reinterpret_cast<uint64_t*>(B0 + 2 * SBUF), lane);
  }
  CH_DEV static void sens_sample(const ProblemData& pd, const ChainArgs& a, int b, double* S, StageFeed& feed, int lane) {
    const int N = pd.N;
    const bool qmode = pd.mode == MODE_Q;
    double* it = a.it + (size_t)b * it_size(N);
    double* ws = a.ws + (size_t)b * ws_size(N);
    const double* tab = a.tab;
    double* Pm = S; double* PV = Pm + NX * NX; double* Tm = PV + NX; double* HU = Tm + NX * NC; double* KK = HU + NU * NC;
    double* E = KK + NU * NK1; double* UU = E + (N + 1) * NX; double* LAM = UU + N * NU; double* TT = LAM + N * NR;
    double* YX = TT + N * NR; double* YN = YX + NU * NX; double* YU = YN + NU * NX; double* G0I = YU + NU * NU; double* B0 = G0I + NU * NU;
    B0 = S + ev2((int)(B0 - S));
    (void)B0;  // the two slots behind it belong to `feed` (sens_feed_init)

    for (int e = lane; e < (N + 1) * NX; e += 32) E[e] = it[it_x(N, 0) + e] - tab[TB_XSS + (e % NX)];
    for (int e = lane; e < N * NU; e += 32) UU[e] = it[it_u(N, 0) + e];
    for (int e = lane; e < N * NR; e += 32) { LAM[e] = it[it_lam(N, 0) + e]; TT[e] = it[it_t(N, 0) + e]; }
    WSYNC();
    // ---- residuals and cost ----
    double stat = 0.0, eq = 0.0, ineq = 0.0, comp = 0.0, cost = 0.0;
    for (int k = lane; k <= N; k += 32) {
      const double* rec = ws + (size_t)k * REC;
      cost += rec[Z_C];
      eq = dmax(eq, rec[Z_E]);
      stat = dmax(stat, rec[Z_S]);
      if (k < N && !(k == 0 && qmode)) {
        MPC_UNROLL for (int q = 0; q < NU; ++q) {
          const double u = UU[k * NU + q];
          const double ll = LAM[k * NR + q], tl = TT[k * NR + q], lu = LAM[k * NR + NU + q], tu = TT[k * NR + NU + q];
          ineq = dmax(ineq, dmax(nn(dabs(pd.lbu[q] - u + tl)), nn(dabs(u - pd.ubu[q] + tu))));
          comp = dmax(comp, dmax(nn(dabs(ll * tl - pd.tau)), nn(dabs(lu * tu - pd.tau))));
        }
      }
    }
    cost = wsum(cost); stat = wmax(stat); eq = wmax(eq); ineq = wmax(ineq); comp = wmax(comp);
    // ---- dL/dtheta ----
    if (a.dL) {
      double* dL = a.dL + (size_t)b * NTH;
      for (int p = lane; p < NPD; p += 32) {
        double g = 0.0;
        for (int k = 0; k < N; ++k) g += ws[(size_t)k * REC + Z_GT + p];
        dL[Mo::pd_to_theta(p)] = g;
      }
      for (int e = lane; e < NX * NX; e += 32) {  // dL/dQ_ij = 1/2 sum_k s_k e_i e_j  (theta index i + j NX, column-major)
        const int i = e % NX, j = e / NX;
        double g = 0.0;
        for (int k = 0; k <= N; ++k) g += pd.scale[k] * E[k * NX + i] * E[k * NX + j];
        dL[Mo::TH_Q + e] = 0.5 * g;
      }
      if (lane < NU * NU) {
        const int i = lane % NU, j = lane / NU;
        double g = 0.0;
        for (int k = 0; k < N; ++k) g += pd.scale[k] * UU[k * NU + i] * UU[k * NU + j];
        dL[Mo::TH_R + lane] = 0.5 * g;
      }
    }
    bool ok_all = true;
    if (a.dpi && !qmode) {
      // ---- backward: exact-Hessian Riccati factorisation ----
      {
        const double sN = pd.scale[N];
        for (int e = lane; e < NX * NX; e += 32) Pm[e] = sN * tab[TB_Q + e];
      }
      WSYNC();
      feed.issue((N - 1) & 1, lane, ws + (size_t)(N - 1) * REC, 0, SEG_SENS);
      for (int k = N - 1; k >= 0; --k) {
        if (k > 0) feed.issue((k - 1) & 1, lane, ws + (size_t)(k - 1) * REC, 0, SEG_SENS);
        const double* Mk = feed.wait(k & 1);
        const double* Hk = Mk + Z_H;
        for (int e = lane; e < NX * NX; e += 32) ws[(size_t)k * REC + Z_P + e] = Pm[e];  // P_{k+1}
        const double s = pd.scale[k];
        const double* lam = LAM + k * NR;
        const double* tt = TT + k * NR;
        auto hess = [&](int i, int c) -> double {
          double v = Hk[i * NW + c];
          if (i < NX) return v + ((c < NX) ? s * tab[TB_Q + i * NX + c] : 0.0);
          if (c < NX) return v;
          v += s * tab[TB_R + (i - NX) * NU + (c - NX)];
          if (i == c) v += lam[i - NX] / tt[i - NX] + lam[NU + i - NX] / tt[NU + i - NX];
          return v;
        };
        auto grad = [&](int) -> double { return 0.0; };
        const bool ok = riccati_step(Pm, PV, Tm, HU, KK, Mk, false, false, /*need_pd=*/false, hess, grad, k == 0 ? G0I : nullptr, lane);
        ok_all = ok_all && ok;
        for (int e = lane; e < NU * NX; e += 32) ws[(size_t)k * REC + Z_K + e] = KK[(e / NX) * NK1 + (e % NX)];
        WSYNC();
      }
      feed.publish(lane);  // K, P were written with ordinary stores and are read back by bulk copies
      // ---- forward: adjoint solves, rank-2 updates of the Q / R columns ----
      double accQ[NU][NQE];
      int qi[NQE], qj[NQE];
      MPC_UNROLL for (int q = 0; q < NQE; ++q) {
        const int e = lane + 32 * q;
        qi[q] = (e < NX * NX) ? e % NX : 0;
        qj[q] = (e < NX * NX) ? e / NX : 0;
        MPC_UNROLL for (int r = 0; r < NU; ++r) accQ[r][q] = 0.0;
      }
      double accR[NU] = {0.0, 0.0, 0.0};  // lane < 9: entry (i, j) = (lane % 3, lane / 3) of dpi_r / dR
      for (int e = lane; e < NU * NX; e += 32) YX[e] = 0.0;
      WSYNC();
      constexpr int OK_ = ev2(NX * NC);  // offset of [K | pad | P] behind [A | B | b] in the slot
      feed.issue(0, lane, ws, 0, ev2(NX * NC), ws + Z_K, OK_, Z_PEND - Z_K);
      for (int k = 0; k < N; ++k) {
        if (k + 1 < N)
          feed.issue((k + 1) & 1, lane, ws + (size_t)(k + 1) * REC, 0, ev2(NX * NC), ws + (size_t)(k + 1) * REC + Z_K, OK_, Z_PEND - Z_K);
        const double* Mk = feed.wait(k & 1);
        const double* Kk = Mk + OK_;
        const double* Pn = Mk + OK_ + (Z_P - Z_K);
        const double s = pd.scale[k];
        if (lane < NU * NU) {  // yu[r][a], lane = r * NU + a
          const int r = lane / NU, a_ = lane - r * NU;
          double v = (k == 0) ? G0I[a_ * NU + r] : 0.0;
          for (int l = 0; l < NX; ++l) v += Kk[a_ * NX + l] * YX[r * NX + l];
          YU[lane] = v;
        }
        WSYNC();
        for (int e = lane; e < NU * NX; e += 32) {  // yx_{k+1}[r][i]
          const int r = e / NX, i = e - r * NX;
          double v = 0.0;
          for (int l = 0; l < NX; ++l) v += Mk[i * NC + l] * YX[r * NX + l];
          MPC_UNROLL for (int q = 0; q < NU; ++q) v += Mk[i * NC + NX + q] * YU[r * NU + q];
          YN[e] = v;
        }
        // Q columns of stage k (yx_0 = 0):  dpi_r/dQ_ij -= 1/2 s_k (yx_i e_j + yx_j e_i)
        if (k > 0) {
          MPC_UNROLL for (int q = 0; q < NQE; ++q)
            MPC_UNROLL for (int r = 0; r < NU; ++r)
              accQ[r][q] -= 0.5 * s * (YX[r * NX + qi[q]] * E[k * NX + qj[q]] + YX[r * NX + qj[q]] * E[k * NX + qi[q]]);
        }
        if (lane < NU * NU) {
          const int i = lane % NU, j = lane / NU;
          MPC_UNROLL for (int r = 0; r < NU; ++r)
            accR[r] -= 0.5 * s * (YU[r * NU + i] * UU[k * NU + j] + YU[r * NU + j] * UU[k * NU + i]);
        }
        WSYNC();
        double* yrec = ws + (size_t)k * REC + Z_Y;
        for (int e = lane; e < NU * NX; e += 32) {
          const int r = e / NX, i = e - r * NX;
          double v = 0.0;
          for (int l = 0; l < NX; ++l) v += Pn[i * NX + l] * YN[r * NX + l];
          yrec[NU * NW + e] = v;            // ypi_k
          yrec[r * NW + i] = YX[e];         // yx_k
        }
        if (lane < NU * NU) yrec[(lane / NU) * NW + NX + (lane % NU)] = YU[lane];
        WSYNC();
        for (int e = lane; e < NU * NX; e += 32) YX[e] = YN[e];
        WSYNC();
      }
      {  // terminal stage
        const double s = pd.scale[N];
        MPC_UNROLL for (int q = 0; q < NQE; ++q)
          MPC_UNROLL for (int r = 0; r < NU; ++r)
            accQ[r][q] -= 0.5 * s * (YX[r * NX + qi[q]] * E[N * NX + qj[q]] + YX[r * NX + qj[q]] * E[N * NX + qi[q]]);
      }
      double* dpi = a.dpi + (size_t)b * NU * NTH;
      MPC_UNROLL for (int q = 0; q < NQE; ++q) {
        const int e = lane + 32 * q;
        if (e < NX * NX) MPC_UNROLL for (int r = 0; r < NU; ++r) dpi[(size_t)r * NTH + Mo::TH_Q + e] = accQ[r][q];
      }
      if (lane < NU * NU) MPC_UNROLL for (int r = 0; r < NU; ++r) dpi[(size_t)r * NTH + Mo::TH_R + lane] = accR[r];
    } else if (a.dpi) {  // Q-mode: u_0 is clamped, dpi/dtheta = 0 (quirk Q7)
      double* dpi = a.dpi + (size_t)b * NU * NTH;
      for (int e = lane; e < NU * NTH; e += 32) dpi[e] = 0.0;
    }
    // ---- outputs ----
    const double rmax = dmax(dmax(stat, eq), dmax(ineq, comp));
    if (lane == 0) {
      if (a.res_out) {
        a.res_out[(size_t)b * 4 + 0] = stat; a.res_out[(size_t)b * 4 + 1] = eq;
        a.res_out[(size_t)b * 4 + 2] = ineq; a.res_out[(size_t)b * 4 + 3] = comp;
      }
      int status = a.have_solve ? a.status[b] : ST_OK;
      if (!(rmax < 1e299)) status = ST_NAN;
      if (!a.have_solve) status = (rmax < 1e299) ? (rmax < pd.tol ? ST_OK : ST_MAXITER) : ST_NAN;
      if (!ok_all && status == ST_OK) status = ST_QPFAIL;
      if (a.cost_out) a.cost_out[b] = cost;
      if (a.status_out) a.status_out[b] = status;
    }
    if (a.u0_out && lane < NU) a.u0_out[(size_t)b * NU + lane] = it[it_u(N, 0) + lane];
  }

  // ================================================================================================================
  // One thread per (sample, stage, right-hand side): contribution of stage k to the dynamic-parameter columns of
  // dpi_r / dtheta.  `part` is the thread's slot of NPD doubles (shared memory); the caller sums over the stages.
  // ================================================================================================================
  CH_DEV static void param_task(const ProblemData& pd, const ChainArgs& a, int b, int k, int r, double* part) {
    const int N = pd.N;
    const double* it = a.it + (size_t)b * it_size(N);
    const double* yrec = a.ws + ((size_t)b * (N + 1) + k) * REC + Z_Y;
    double g[NPD];
    Mo::param_contraction(it + it_x(N, k), it + it_u(N, k), a.th, pd.mc[0], yrec + r * NW, yrec + r * NW + NX, it + it_pi(N, k),
                          yrec + NU * NW + r * NX, g);
    MPC_UNROLL for (int p = 0; p < NPD; ++p) part[p] = -g[p];
  }
};

}  // namespace rlmpc
