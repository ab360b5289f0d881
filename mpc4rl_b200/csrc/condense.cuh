// Partial condensing of the stage QP for the queued interior-point solves.
//
// The reference solves its QPs with PARTIAL_CONDENSING_HPIPM (config/cartpole*.yaml:8): consecutive
// stages are merged into blocks before the Riccati recursion.  The same idea shortens the dependent
// chain that bounds the queue kernel (profiles/r01_summary.md): S stages of (nx, nu) become one block
// stage of (nx, S*nu) -- N/S recursion steps per sweep instead of N, each made of small dense
// matrix products with instruction-level parallelism a lone warp can use.
//
//   block i, stages k0 = i*S .. k0+S-1, y = [dx_{k0} ; dU],  dU = (du_{k0}, .., du_{k0+S-1}):
//     dx_{k0+j} = Phi_j dx_{k0} + Gam_j dU + beta_j        (Gam_j has zero columns for inputs >= j)
//     [dx;du]_j = T_j y + c_j,   T_j = [Phi_j Gam_j ; 0 E_j],  c_j = [beta_j ; 0]
//     block cost   1/2 y' Hb y + gb' y,   Hb = sum_j T_j' H_j T_j,   gb = sum_j T_j' (H_j c_j + g_j)
//     block dynamics  dx_{k0+S} = Phi_S dx_{k0} + Gam_S dU + beta_S
//   Bounds are on inputs only, so the block rows are the stage rows re-ordered and the barrier terms
//   stay diagonal: the interior-point iteration on the blocks IS the iteration on the stages.
// Applies to problems with input bounds only (NBX = NSX = NG = 0), N divisible by S.  Q-mode (u_0
// clamped): the inputs of stage 0 stay in block 0 as decoupled dummy variables (unit Hessian, zero
// gradient, zero column in Gam, no rows), so their step is exactly zero.
// condense_block() is a (sample, block) function, solve + expand a sample function.
#pragma once
#include "engine.cuh"

namespace rlmpc {

template <class M, int S>
struct BlockModel {
  static constexpr int NX = M::NX, NU = S * M::NU, NPM = 1, NTH = 1;
  static constexpr int NBX = 0, NSX = 0, NG = 0;
  static constexpr bool STAGE_HESS = true;
  static constexpr bool PARAMS_COST_ONLY = false;
  MPC_HD static int bx(int j) { return j; }
  MPC_HD static int sx(int) { return 0; }
  MPC_HD static double gC(int, int) { return 0.0; }
  MPC_HD static double g0(int) { return 0.0; }
};

template <class M, int S>
struct Condenser {
  using E = Engine<M>;
  using EB = Engine<BlockModel<M, S>>;
  static constexpr int NX = M::NX, NU = M::NU, NW = NX + NU;
  static constexpr int NUB = S * NU, NWB = NX + NUB;
  static_assert(M::NBX == 0 && M::NSX == 0 && M::NG == 0, "partial condensing: input bounds only");

  // block problem data: N/S stages, input bounds replicated, mode V
  static bool applicable(const ProblemData& pd) { return pd.N % S == 0 && NUB <= MAXD; }
  static ProblemData block_pd(const ProblemData& pd) {
    ProblemData b = pd;
    b.N = pd.N / S;
    b.fix0 = (pd.mode == MODE_Q) ? NU : 0;
    b.mode = MODE_V;
    for (int j = 0; j < S; ++j)
      for (int c = 0; c < NU; ++c) {
        b.lbu[j * NU + c] = pd.lbu[c];
        b.ubu[j * NU + c] = pd.ubu[c];
      }
    return b;
  }

  // (sample, block) function.  L: the sample's stage-form lane (workspace [W_A, W_K) filled by lin_stage,
  // iterate), Lb: its block-form lane.  i in [0, N/S]: i = N/S is the terminal stage.
  MPC_HD static void condense_block(const ProblemData& pd, const Lane& L, const Lane& Lb, int i) {
    const int N = pd.N, Nb = N / S;
    constexpr size_t bs = TILE;
    double* wb = Lb.ws + (size_t)i * EB::W_REC * bs;
    if (i == Nb) {  // terminal: gradient q_N and Hessian s_N W_e
      double g[NX], Hm[NW * NW], Hp[EB::NWS];
      EB::template ld<NX>(L.ws + ((size_t)N * E::W_REC + E::W_q) * bs, bs, g);
      E::load_W(2, pd.scale[N], L, Hm);
      MPC_UNROLL for (int a = 0; a < EB::NWS; ++a) Hp[a] = 0.0;
      MPC_UNROLL for (int a = 0; a < NX; ++a) MPC_UNROLL for (int b = a; b < NX; ++b) Hp[EB::pidx(a, b)] = Hm[a * NW + b];
      EB::template st<NX>(wb + (size_t)EB::W_q * bs, bs, g);
      EB::template st<EB::NWS>(wb + (size_t)EB::W_H * bs, bs, Hp);
      if (i == Nb) Lb.it[(size_t)EB::it_meta(Nb) * bs] = L.it[(size_t)E::it_meta(N) * bs];
      return;
    }
    // Z = [Phi | Gam] (NX x NWB), beta (NX); accumulate Hb (NWB x NWB, upper) and gb (NWB)
    double Z[NX * NWB], beta[NX], Hb[NWB * NWB], gb[NWB];
    MPC_UNROLL for (int a = 0; a < NX; ++a) {
      beta[a] = 0.0;
      MPC_UNROLL for (int b = 0; b < NWB; ++b) Z[a * NWB + b] = (a == b) ? 1.0 : 0.0;
    }
    MPC_UNROLL for (int a = 0; a < NWB * NWB; ++a) Hb[a] = 0.0;
    MPC_UNROLL for (int a = 0; a < NWB; ++a) gb[a] = 0.0;
    MPC_UNROLL for (int j = 0; j < S; ++j) {
      const int k = i * S + j;
      const double* w = L.ws + (size_t)k * E::W_REC * bs;
      double A[NX * NX], B[NX * NU], bb[NX], g[NW], Hm[NW * NW];
      E::template ld<NX * NX>(w + (size_t)E::W_A * bs, bs, A);
      E::template ld<NX * NU>(w + (size_t)E::W_B * bs, bs, B);
      E::template ld<NX>(w + (size_t)E::W_b * bs, bs, bb);
      E::template ld<NW>(w + (size_t)E::W_q * bs, bs, g);
      E::load_W(k == 0 ? 0 : 1, pd.scale[k], L, Hm);
      // T_j = [Z ; E_j] (NW x NWB): rows 0..NX-1 = Z, rows NX..NW-1 select the inputs of stage j
      // HT = H_j T_j (NW x NWB);  v = H_j c_j + g_j with c_j = [beta ; 0]
      double HT[NW * NWB], v[NW];
      MPC_UNROLL for (int a = 0; a < NW; ++a) {
        double va = g[a];
        MPC_UNROLL for (int l = 0; l < NX; ++l) va += Hm[a * NW + l] * beta[l];
        v[a] = va;
        MPC_UNROLL for (int b = 0; b < NWB; ++b) {
          double acc = 0.0;
          MPC_UNROLL for (int l = 0; l < NX; ++l) acc += Hm[a * NW + l] * Z[l * NWB + b];
          if (b >= NX + j * NU && b < NX + (j + 1) * NU) acc += Hm[a * NW + NX + (b - NX - j * NU)];
          HT[a * NWB + b] = acc;
        }
      }
      MPC_UNROLL for (int a = 0; a < NWB; ++a) {
        // row a of T_j': column a of T_j = (Z[:,a] ; e) -- e nonzero iff a is an input of stage j
        const bool isu = (a >= NX + j * NU && a < NX + (j + 1) * NU);
        double ga = 0.0;
        MPC_UNROLL for (int l = 0; l < NX; ++l) ga += Z[l * NWB + a] * v[l];
        if (isu) ga += v[NX + (a - NX - j * NU)];
        gb[a] += ga;
        MPC_UNROLL for (int b = a; b < NWB; ++b) {
          double acc = 0.0;
          MPC_UNROLL for (int l = 0; l < NX; ++l) acc += Z[l * NWB + a] * HT[l * NWB + b];
          if (isu) acc += HT[(NX + (a - NX - j * NU)) * NWB + b];
          Hb[a * NWB + b] += acc;
        }
      }
      // advance: Z <- A Z + [0 | B at the columns of stage j],  beta <- A beta + b
      double Zn[NX * NWB], bn[NX];
      MPC_UNROLL for (int a = 0; a < NX; ++a) {
        double ba = bb[a];
        MPC_UNROLL for (int l = 0; l < NX; ++l) ba += A[a * NX + l] * beta[l];
        bn[a] = ba;
        MPC_UNROLL for (int b = 0; b < NWB; ++b) {
          double acc = 0.0;
          MPC_UNROLL for (int l = 0; l < NX; ++l) acc += A[a * NX + l] * Z[l * NWB + b];
          if (b >= NX + j * NU && b < NX + (j + 1) * NU) acc += B[a * NU + (b - NX - j * NU)];
          Zn[a * NWB + b] = acc;
        }
      }
      MPC_UNROLL for (int a = 0; a < NX * NWB; ++a) Z[a] = Zn[a];
      MPC_UNROLL for (int a = 0; a < NX; ++a) beta[a] = bn[a];
    }
    if (i == 0 && pd.mode == MODE_Q) {  // clamped u_0: decouple the inputs of stage 0
      MPC_UNROLL for (int c = 0; c < NU; ++c) {
        const int q = NX + c;
        MPC_UNROLL for (int a = 0; a < NWB; ++a) {
          Hb[(a < q ? a : q) * NWB + (a < q ? q : a)] = (a == q) ? 1.0 : 0.0;
        }
        gb[q] = 0.0;
        MPC_UNROLL for (int a = 0; a < NX; ++a) Z[a * NWB + q] = 0.0;
      }
    }
    // block record: A = Phi_S, B = Gam_S, b = beta_S, g = gb, H = Hb
    double Ab[NX * NX], Bb[NX * NUB], Hp[EB::NWS];
    MPC_UNROLL for (int a = 0; a < NX; ++a) {
      MPC_UNROLL for (int b = 0; b < NX; ++b) Ab[a * NX + b] = Z[a * NWB + b];
      MPC_UNROLL for (int b = 0; b < NUB; ++b) Bb[a * NUB + b] = Z[a * NWB + NX + b];
    }
    MPC_UNROLL for (int a = 0; a < NWB; ++a) MPC_UNROLL for (int b = a; b < NWB; ++b) Hp[EB::pidx(a, b)] = Hb[a * NWB + b];
    EB::template st<NX * NX>(wb + (size_t)EB::W_A * bs, bs, Ab);
    EB::template st<NX * NUB>(wb + (size_t)EB::W_B * bs, bs, Bb);
    EB::template st<NX>(wb + (size_t)EB::W_b * bs, bs, beta);
    EB::template st<NWB>(wb + (size_t)EB::W_q * bs, bs, gb);
    EB::template st<EB::NWS>(wb + (size_t)EB::W_H * bs, bs, Hp);
    // iterate of the block: inputs, multipliers and slacks of its S stages; stage rows [lb(NU) ub(NU)]
    // become block rows [lb(S NU) ub(S NU)]
    double ub_[NUB], lamb[EB::NR], tb[EB::NR];
    MPC_UNROLL for (int j = 0; j < S; ++j) {
      const int k = i * S + j;
      double u[NU], lam[E::NR], t[E::NR];
      E::template ld<NU>(L.it + (size_t)E::it_u(N, k) * bs, bs, u);
      E::template ld<E::NR>(L.it + (size_t)E::it_lam(N, k) * bs, bs, lam);
      E::template ld<E::NR>(L.it + (size_t)E::it_t(N, k) * bs, bs, t);
      MPC_UNROLL for (int c = 0; c < NU; ++c) {
        ub_[j * NU + c] = u[c];
        lamb[j * NU + c] = lam[c];
        lamb[NUB + j * NU + c] = lam[NU + c];
        tb[j * NU + c] = t[c];
        tb[NUB + j * NU + c] = t[NU + c];
      }
    }
    EB::template st<NUB>(Lb.it + (size_t)EB::it_u(Nb, i) * bs, bs, ub_);
    EB::template st<EB::NR>(Lb.it + (size_t)EB::it_lam(Nb, i) * bs, bs, lamb);
    EB::template st<EB::NR>(Lb.it + (size_t)EB::it_t(Nb, i) * bs, bs, tb);
  }

  // Sample function: interior-point solve on the blocks, then back to stage form (dx, du, lam_hat,
  // t_hat per stage, current lam, t) and the ordinary step.  Returns Engine::Full codes.
  template <class RD>
  MPC_HD static int solve_expand(const ProblemData& pd, const ProblemData& pdb, const Lane& L, const Lane& Lb,
                                 int* ipm_iters, RD& rd) {
    const int N = pd.N, Nb = N / S;
    constexpr size_t bs = TILE;
    double alpha = 0.0;
    const int r = EB::qp_ipm(pdb, Lb, &alpha, rd);
    if (ipm_iters) *ipm_iters += (r > 0) ? r : ((r >= -2) ? 0 : -(r + 1000));
    if (r == -1 || r == -2) return E::FULL_FAILED;
    double dx[NX];
    EB::template ld<NX>(Lb.ws + (size_t)EB::W_dx * bs, bs, dx);  // block 0 starts at dx_0 = 0
    for (int i = 0; i < Nb; ++i) {
      const double* wb = Lb.ws + (size_t)i * EB::W_REC * bs;
      double dU[NUB], lh[EB::NR], th[EB::NR], lamb[EB::NR], tb[EB::NR];
      EB::template ld<NUB>(wb + (size_t)EB::W_du * bs, bs, dU);
      EB::template ld<EB::NR>(wb + (size_t)EB::W_lh * bs, bs, lh);
      EB::template ld<EB::NR>(wb + (size_t)EB::W_th * bs, bs, th);
      EB::template ld<EB::NR>(Lb.it + (size_t)EB::it_lam(Nb, i) * bs, bs, lamb);
      EB::template ld<EB::NR>(Lb.it + (size_t)EB::it_t(Nb, i) * bs, bs, tb);
      for (int j = 0; j < S; ++j) {
        const int k = i * S + j;
        double* w = L.ws + (size_t)k * E::W_REC * bs;
        double A[NX * NX], B[NX * NU], bb[NX], du[NU], lhs[E::NR], ths[E::NR], lam[E::NR], t[E::NR];
        E::template ld<NX * NX>(w + (size_t)E::W_A * bs, bs, A);
        E::template ld<NX * NU>(w + (size_t)E::W_B * bs, bs, B);
        E::template ld<NX>(w + (size_t)E::W_b * bs, bs, bb);
        MPC_UNROLL for (int c = 0; c < NU; ++c) {
          du[c] = dU[j * NU + c];
          lhs[c] = lh[j * NU + c];       lhs[NU + c] = lh[NUB + j * NU + c];
          ths[c] = th[j * NU + c];       ths[NU + c] = th[NUB + j * NU + c];
          lam[c] = lamb[j * NU + c];     lam[NU + c] = lamb[NUB + j * NU + c];
          t[c] = tb[j * NU + c];         t[NU + c] = tb[NUB + j * NU + c];
        }
        E::template st<NX>(w + (size_t)E::W_dx * bs, bs, dx);
        E::template st<NU>(w + (size_t)E::W_du * bs, bs, du);
        E::template st<E::NR>(w + (size_t)E::W_lh * bs, bs, lhs);
        E::template st<E::NR>(w + (size_t)E::W_th * bs, bs, ths);
        E::template st<E::NR>(L.it + (size_t)E::it_lam(N, k) * bs, bs, lam);
        E::template st<E::NR>(L.it + (size_t)E::it_t(N, k) * bs, bs, t);
        double dxn[NX];
        MPC_UNROLL for (int a = 0; a < NX; ++a) {
          double acc = bb[a];
          MPC_UNROLL for (int l = 0; l < NX; ++l) acc += A[a * NX + l] * dx[l];
          MPC_UNROLL for (int l = 0; l < NU; ++l) acc += B[a * NU + l] * du[l];
          dxn[a] = acc;
        }
        MPC_UNROLL for (int a = 0; a < NX; ++a) dx[a] = dxn[a];
      }
    }
    E::template st<NX>(L.ws + ((size_t)N * E::W_REC + E::W_dx) * bs, bs, dx);
    E::apply_step(pd, L, alpha, /*clip=*/false, /*damp_primal=*/r < 0);
    return (r < 0) ? E::FULL_MAXITER : E::FULL_OK;
  }
};

}  // namespace rlmpc
