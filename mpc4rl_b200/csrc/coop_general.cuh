// Warp-per-sample interior-point solve of the queued stage QPs, general version (CUDA only).
//
// Same mapping as coop.cuh (one warp owns one sample, the stage QP lives in shared memory, row work with one
// lane per stage, Riccati recursions with one lane per matrix entry) for every model whose blocks fit a warp:
// NX (NX+NU+1) <= 32 and (NX+NU)(NX+NU+1) <= 32.  All row handling goes through the Engine's own per-stage
// functions (stage_bounds, stage_vars, barrier_add, rows_forward, ipm_init_stage, ipm_project_stage), so
// state bounds, general affine rows and softened bounds behave exactly as in the thread-per-sample kernels;
// NU > 1 uses the Engine's LDL' solve of the NU x NU block, redundantly in every lane.  coop.cuh remains the
// leaner specialisation for NU = 1 with input bounds only (the headline cart-pole).
#pragma once

#include "coop.cuh"

namespace rlmpc {

template <class M>
struct CoopGenOK {
  static constexpr int NW = M::NX + M::NU;
  static constexpr bool value = !M::STAGE_HESS && M::NX * (NW + 1) <= 32 && NW * (NW + 1) <= 32;
};

#ifdef __CUDACC__
template <class M>
struct CoopGen {
  using E = Engine<M>;
  using Bnd = typename E::Bnd;
  using StepStats = typename E::StepStats;
  static constexpr int NX = M::NX, NU = M::NU, NW = NX + NU, NC = NW + 1, NWS = E::NWS;
  static constexpr int NR = E::NR, NV = E::NV, NBX = M::NBX;
  static constexpr int NKK = NU * NX + NU;  // feedback law of a stage: K (NU x NX) | kff (NU)
  static_assert(NKK >= NW, "the feedback-law slot is reused for [dx | du]");
  static constexpr int PER_STAGE = NX * NC + 2 * NW + NWS + NKK + NW + 4 * NR;
  static constexpr int SCRATCH = 2 * NX + NX * NC + NW * NC + 3 * NWS + 1;
  __host__ __device__ static constexpr int smem_doubles(int N) { return PER_STAGE * (N + 1) + SCRATCH; }

  __device__ static __forceinline__ double lds(unsigned a) { return CoopQPBase::lds(a); }
  __device__ static __forceinline__ void sts(unsigned a, double v) { CoopQPBase::sts(a, v); }

  __device__ static int solve(const ProblemData& pd, const Lane& L, double* S, const int lane, const bool swept, int* iters_out) {
    const int N = pd.N, NS = N + 1;
    constexpr size_t bs = TILE;
    double* Mk = S;                    // [k][NX][NC] = [A | B | b]
    double* Gk = Mk + NS * NX * NC;    // [k][NW]     = [q ; r]   (later: x-part of the costate recursion)
    double* GKB = Gk + NS * NW;        // [k][NW]     gradient incl. barrier terms (row phase, every iteration)
    double* HKB = GKB + NS * NW;       // [k][NWS]    scaled cost Hessian + barrier terms, packed upper
    double* Kk = HKB + NS * NWS;       // [k][NKK]    [K | kff]; forward sweep: [dx_{k+1} | du_k]; finally pi_k
    double* XU = Kk + NS * NKK;        // [k][NW]     x_k, u_k of the linearisation point
    double* LAM = XU + NS * NW;        // [q][k]
    double* TT = LAM + NR * NS;
    double* LH = TT + NR * NS;
    double* TH = LH + NR * NS;
    double* pv = TH + NR * NS;         // scratch
    double* pv2 = pv + NX;
    double* PM = pv2 + NX;
    double* T = PM + NX * NC;
    double* Wc = T + NW * NC;
    double* ZERO = Wc + 3 * NWS;
    (void)ZERO;

    const unsigned sT = (unsigned)__cvta_generic_to_shared(T), sPM = (unsigned)__cvta_generic_to_shared(PM);
    const unsigned sKk = (unsigned)__cvta_generic_to_shared(Kk), sMk = (unsigned)__cvta_generic_to_shared(Mk);
    const bool qmode = pd.mode == MODE_Q;
    const double m_rows = (double)E::count_rows(pd);

    // ---------------- stage the QP ----------------
    constexpr int NAB = NX * NX + NX * NU + NX;  // record elements [W_A, W_q): A, B, b
#pragma unroll 4  // several of the scattered global loads in flight per lane
    for (int idx = lane; idx < N * NAB; idx += 32) {
      const int k = idx / NAB, e = idx - k * NAB;
      const double v = L.ws[((size_t)k * E::W_REC + e) * bs];
      int row, col;
      if (e < NX * NX) {
        row = e / NX; col = e - row * NX;
      } else if (e < NX * NX + NX * NU) {
        row = (e - NX * NX) / NU; col = NX + (e - NX * NX) - row * NU;
      } else {
        row = e - NX * NX - NX * NU; col = NW;
      }
      Mk[k * NX * NC + row * NC + col] = v;
    }
#pragma unroll 4  // several of the scattered global loads in flight per lane
    for (int idx = lane; idx < NS * NW; idx += 32) {
      const int k = idx / NW, e = idx - k * NW;
      Gk[idx] = (k < N || e < NX) ? L.ws[((size_t)k * E::W_REC + E::W_q + e) * bs] : 0.0;
      XU[idx] = (e < NX) ? L.it[(size_t)(E::it_x(N, k) + e) * bs] : (k < N ? L.it[(size_t)(E::it_u(N, k) + e - NX) * bs] : 0.0);
    }
#pragma unroll 4  // several of the scattered global loads in flight per lane
    for (int idx = lane; idx < NS * NR; idx += 32) {
      const int k = idx / NR, q = idx - k * NR;
      LAM[q * NS + k] = L.it[(size_t)(E::it_lam(N, k) + q) * bs];
      TT[q * NS + k] = L.it[(size_t)(E::it_t(N, k) + q) * bs];
      LH[q * NS + k] = 0.0;
      TH[q * NS + k] = 0.0;
    }
    for (int idx = lane; idx < 3 * NWS; idx += 32) {
      const int kind = idx / NWS, e = idx - kind * NWS;
      Wc[idx] = L.ct[(size_t)(kind * E::CT_REC + E::CT_W + e) * bs];
    }
    bool warm = pd.warm_ipm && L.it[(size_t)E::it_meta(N) * bs] > 0.5;
    __syncwarp();

    // rows of stage k <-> registers
    auto has_rows = [&](int k) { return k < N || NBX > 0; };
    auto stage_point = [&](int k, Bnd& bd, double* v) {
      double x[NX], u[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = XU[k * NW + i];
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = XU[k * NW + NX + i];
      E::stage_bounds(pd, k, bd);
      E::stage_vars(x, u, v);
    };
    auto ld_rows = [&](const double* A, int k, double* out) {
#pragma unroll
      for (int q = 0; q < NR; ++q) out[q] = A[q * NS + k];
    };
    auto st_rows = [&](double* A, int k, const double* in) {
#pragma unroll
      for (int q = 0; q < NR; ++q) A[q * NS + k] = in[q];
    };

    auto init_rows = [&](bool w) -> double {  // [Engine::ipm_init]
      double mu = 0.0;
      for (int k = lane; k < NS; k += 32) {
        if (!has_rows(k)) continue;
        Bnd bd;
        double v[NV], lam[NR], t[NR];
        stage_point(k, bd, v);
        ld_rows(LAM, k, lam);
        ld_rows(TT, k, t);
        mu += E::ipm_init_stage(pd, bd, v, w, lam, t);
        st_rows(LAM, k, lam);
        st_rows(TT, k, t);
      }
      __syncwarp();
      return CoopQPBase::wsum(mu);
    };

    // ---- per-lane roles of the two recursions ----
    const int iA = lane / NC, cA = lane - iA * NC;
    const bool inA = lane < NX * NC, inB = lane < NW * NC;
    const int iAx = iA < NX ? iA : NX - 1;
    const int k_last = qmode ? 1 : 0;
    const unsigned sMA = sMk + 8 * cA, sMB = sMk + 8 * (iA < NW ? iA : 0);
    unsigned oT[NX];
#pragma unroll
    for (int l = 0; l < NX; ++l) oT[l] = 8 * ((iAx < l ? iAx : l) * NC + (iAx < l ? l : iAx));
    unsigned sBase, baseStride;  // phase B: entry (iA, cA) of [H | g] of stage k
    {
      const int r = iA < NW ? iA : 0;
      if (cA < NW) {
        sBase = (unsigned)__cvta_generic_to_shared(HKB + E::pidx(r < cA ? r : cA, r < cA ? cA : r)); baseStride = 8 * NWS;
      } else {
        sBase = (unsigned)__cvta_generic_to_shared(GKB + r); baseStride = 8 * NW;
      }
    }

    double mu = init_rows(warm) / m_rows;
    double alpha = 0.0;
    double sigma = warm ? pd.sigma_min : pd.sigma0;
    int iters = 0, warm_iters = 0, as_iters = 0;
    bool converged = false, failed = false, minstep = false;

    for (int j = 0; j < pd.max_ipm && !converged; ++j) {
      ++iters;
      if (warm && (warm_iters >= E::WARM_LIMIT + as_iters || (warm_iters > as_iters && alpha > 0.0 && alpha < 0.05))) {
        warm = false;
        mu = init_rows(false) / m_rows;
        alpha = 0.0;
        sigma = pd.sigma0;
      }
      if (warm) ++warm_iters;
      const double target = dmax(sigma * mu, pd.tau);
      const bool reuse = swept && j == 0 && warm && target == pd.tau && pd.max_ipm > 1;
      if (!reuse) {
        // ---- rows: pending damped update; Hessian and gradient of the stage incl. barrier terms ----
        for (int k = lane; k < NS; k += 32) {
          double Hm[NW * NW], g[NW];
          {
            const double s = pd.scale[k];
            const double* W = Wc + (k == 0 ? 0 : (k == N ? 2 : 1)) * NWS;
#pragma unroll
            for (int a = 0; a < NW; ++a) {
#pragma unroll
              for (int b = a; b < NW; ++b) {
                const double v = s * W[E::pidx(a, b)];
                Hm[a * NW + b] = v;
                Hm[b * NW + a] = v;
              }
              g[a] = Gk[k * NW + a];
            }
          }
          if (has_rows(k)) {
            Bnd bd;
            double v[NV], lam[NR], t[NR];
            stage_point(k, bd, v);
            ld_rows(LAM, k, lam);
            ld_rows(TT, k, t);
            if (alpha > 0.0) {
              double lh[NR], th[NR];
              ld_rows(LH, k, lh);
              ld_rows(TH, k, th);
#pragma unroll
              for (int q = 0; q < NR; ++q) {
                lam[q] += alpha * (lh[q] - lam[q]);
                t[q] += alpha * (th[q] - t[q]);
              }
              st_rows(LAM, k, lam);
              st_rows(TT, k, t);
            }
            E::barrier_add(bd, v, lam, t, target, Hm, g);
          }
#pragma unroll
          for (int a = 0; a < NW; ++a) {
#pragma unroll
            for (int b = a; b < NW; ++b) HKB[k * NWS + E::pidx(a, b)] = Hm[a * NW + b];
            GKB[k * NW + a] = g[a];
          }
        }
        __syncwarp();
        // ---- backward Riccati sweep (see coop.cuh; here K = -G^{-1} [H | gv] with an NU x NU block G) ----
        if (inB) {  // terminal "T": P_N, p_N; no input: G = I, H = 0
          double v = 0.0;
          if (iA < NX && cA < NX) v = HKB[N * NWS + E::pidx(iA < cA ? iA : cA, iA < cA ? cA : iA)];
          if (iA < NX && cA == NC - 1) v = GKB[N * NW + iA];
          if (iA >= NX && cA == iA) v = 1.0;
          T[lane] = v;
        }
        double mA[NX], mB[NX], baseB;
        auto fetch = [&](int k) {
          const unsigned ma = sMA + (unsigned)k * (8 * NX * NC), mb = sMB + (unsigned)k * (8 * NX * NC);
#pragma unroll
          for (int l = 0; l < NX; ++l) {
            mA[l] = lds(ma + 8 * l * NC);
            mB[l] = lds(mb + 8 * l * NC);
          }
          baseB = lds(sBase + (unsigned)k * baseStride);
        };
        // feedback law of the stage whose T is in shared memory: K (NU x NX), kff (NU); false if G is not PD.
        // Also column iAx of H and of K (this lane's row of P), so that nothing is indexed by a run-time value.
        double Kf[NU * NX], kf[NU], Hf[NU * NX], Hi[NU], Ki[NU];
        auto feedback = [&]() -> bool {
          double G[NU * NU], R[NU * (NX + 2)];
#pragma unroll
          for (int a = 0; a < NU; ++a) {
#pragma unroll
            for (int b = 0; b < NU; ++b) G[a * NU + b] = lds(sT + 8 * ((NX + a) * NC + NX + b));
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) {
              Hf[a * NX + jj] = lds(sT + 8 * ((NX + a) * NC + jj));
              R[a * (NX + 2) + jj] = -Hf[a * NX + jj];
            }
            R[a * (NX + 2) + NX] = -lds(sT + 8 * ((NX + a) * NC + NC - 1));
            Hi[a] = lds(sT + 8 * ((NX + a) * NC) + 8 * iAx);
            R[a * (NX + 2) + NX + 1] = -Hi[a];
          }
          const bool ok = E::template spd_solve<NU, NX + 2>(G, R);
#pragma unroll
          for (int a = 0; a < NU; ++a) {
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) Kf[a * NX + jj] = R[a * (NX + 2) + jj];
            kf[a] = R[a * (NX + 2) + NX];
            Ki[a] = R[a * (NX + 2) + NX + 1];
          }
          return ok;
        };
        auto store_feedback = [&](int k) {
#pragma unroll
          for (int e = 0; e < NU * NX; ++e)
            if (lane == e) Kk[k * NKK + e] = Kf[e];
#pragma unroll
          for (int a = 0; a < NU; ++a)
            if (lane == NU * NX + a) Kk[k * NKK + NU * NX + a] = kf[a];
        };
        fetch(N - 1);
        __syncwarp();
        for (int k = N - 1; k >= k_last; --k) {
          // ---- phase A: finish stage k+1 (P, p), multiply by [A B b] of stage k ----
          if (!feedback()) failed = true;
          {
            double acc = 0.0;
            if (cA == NC - 1) {
              acc = lds(sT + 8 * NC * iAx + 8 * (NC - 1));
#pragma unroll
              for (int a = 0; a < NU; ++a) acc += Hi[a] * kf[a];
            }
#pragma unroll
            for (int l = 0; l < NX; ++l) {
              // P_il = T_ab + sum_a H_a,min K_a,max with (min, max) of (i, l): the same expression on both
              // sides of the diagonal, so P stays bitwise symmetric
              double Pil = lds(sT + oT[l]);
              const bool up = iAx < l;
#pragma unroll
              for (int a = 0; a < NU; ++a) Pil += (up ? Hi[a] : Hf[a * NX + l]) * (up ? Kf[a * NX + l] : Ki[a]);
              acc += Pil * mA[l];
            }
            if (inA) sts(sPM + 8 * lane, acc);
          }
          if (k + 1 < N) store_feedback(k + 1);
          __syncwarp();
          // ---- phase B ----
          {
            double accB = baseB;
#pragma unroll
            for (int l = 0; l < NX; ++l) accB += mB[l] * lds(sPM + 8 * (l * NC) + 8 * cA);
            if (inB) sts(sT + 8 * lane, accB);
          }
          if (k > k_last) fetch(k - 1);
          __syncwarp();
        }
        if (!feedback()) failed = true;
        store_feedback(k_last);
        if (qmode && lane < NKK) Kk[lane] = 0.0;  // u_0 fixed: no feedback at stage 0
        __syncwarp();
        // ---- forward sweep: [dx_{k+1} | du_k] into the feedback-law slot of stage k ----
        {
          double kr[NKK], ar[NC];
          auto fetch_f = [&](int k) {
#pragma unroll
            for (int l = 0; l < NKK; ++l) kr[l] = Kk[k * NKK + l];
#pragma unroll
            for (int l = 0; l < NC; ++l) ar[l] = (lane < NX) ? Mk[k * NX * NC + lane * NC + l] : 0.0;
          };
          fetch_f(0);
        __syncwarp();  // every lane holds stage 0's feedback law before its slot is overwritten
          for (int k = 0; k < N; ++k) {
            double dxl[NX], du[NU];
#pragma unroll
            for (int l = 0; l < NX; ++l) dxl[l] = (k > 0) ? lds(sKk + 8 * ((k - 1) * NKK + l)) : 0.0;
#pragma unroll
            for (int a = 0; a < NU; ++a) {
              du[a] = kr[NU * NX + a];
#pragma unroll
              for (int l = 0; l < NX; ++l) du[a] += kr[a * NX + l] * dxl[l];
            }
            double an = ar[NC - 1];
#pragma unroll
            for (int l = 0; l < NX; ++l) an += ar[l] * dxl[l];
#pragma unroll
            for (int a = 0; a < NU; ++a) an += ar[NX + a] * du[a];
            if (k + 1 < N) fetch_f(k + 1);
            if (lane < NX) sts(sKk + 8 * (k * NKK + lane), an);
#pragma unroll
            for (int a = 0; a < NU; ++a)
              if (lane == NX + a) sts(sKk + 8 * (k * NKK + NX + a), du[a]);
            __syncwarp();
          }
        }
      }  // !reuse
      // ---- rows: new slacks and multipliers, step-length statistics ----
      StepStats St = {1e300, 0.0, 0.0, 0.0, 0.0};
      for (int k = lane; k < NS; k += 32) {
        if (!has_rows(k)) continue;
        Bnd bd;
        double v[NV], lam[NR], t[NR], lh[NR], th[NR];
        stage_point(k, bd, v);
        ld_rows(LAM, k, lam);
        ld_rows(TT, k, t);
        if (reuse) {
#pragma unroll
          for (int q = 0; q < NR; ++q) {
            lh[q] = L.ws[((size_t)k * E::W_REC + E::W_lh + q) * bs];
            th[q] = L.ws[((size_t)k * E::W_REC + E::W_th + q) * bs];
          }
          E::rows_stats(bd, lam, t, lh, th, St);
        } else {
          double dw[NW];
#pragma unroll
          for (int i = 0; i < NX; ++i) dw[i] = (k > 0) ? Kk[(k - 1) * NKK + i] : 0.0;
#pragma unroll
          for (int i = 0; i < NU; ++i) dw[NX + i] = (k < N) ? Kk[k * NKK + NX + i] : 0.0;
          E::rows_forward(bd, v, dw, lam, t, target, lh, th, St);
        }
        st_rows(LH, k, lh);
        st_rows(TH, k, th);
      }
      __syncwarp();
      double amax = St.amax;
      {
        const double probe = reuse ? 0.0 : Kk[NX] + Kk[(N - 1) * NKK];
        if (!(probe == probe)) amax = probe;
      }
      const bool nan_step = __any_sync(0xffffffffu, !(amax == amax));
      amax = CoopQPBase::wmin(amax);
      const double s0 = CoopQPBase::wsum(St.s0), s1 = CoopQPBase::wsum(St.s1), s2 = CoopQPBase::wsum(St.s2);
      const double cmax = CoopQPBase::wmax(St.cmax);
      failed = __any_sync(0xffffffffu, failed);

      if ((failed || nan_step) && warm) {
        warm = false;
        failed = false;
        mu = init_rows(false) / m_rows;
        alpha = 0.0;
        sigma = pd.sigma0;
        continue;
      }
      if (failed || nan_step) { failed = true; break; }
      if (warm && amax < 1.0 / 0.995 && as_iters < (int)pd.as_steps) {  // [Engine::ipm_project]
        ++as_iters;
        double m2 = 0.0;
        for (int k = lane; k < NS; k += 32) {
          if (!has_rows(k)) continue;
          Bnd bd;
          double lam[NR], t[NR], lh[NR], th[NR];
          E::stage_bounds(pd, k, bd);
          ld_rows(LAM, k, lam);
          ld_rows(TT, k, t);
          ld_rows(LH, k, lh);
          ld_rows(TH, k, th);
          m2 += E::ipm_project_stage(pd, bd, lam, t, lh, th);
          st_rows(LAM, k, lam);
          st_rows(TT, k, t);
        }
        __syncwarp();
        mu = CoopQPBase::wsum(m2) / m_rows;
        alpha = 0.0;
        sigma = pd.sigma_min;
        continue;
      }
      alpha = (amax >= 1.0 / 0.995) ? 1.0 : 0.995 * amax;
      if (!warm && alpha < 1e-9) {
        minstep = true;
        break;
      }
      const double mu_new = (s0 + alpha * s1 + alpha * alpha * s2) / m_rows;
      if (target <= pd.tau && alpha == 1.0 && cmax <= dmin(pd.comp_accept * pd.tau, 0.1 * pd.tol)) converged = true;
      const double r = 1.0 - alpha;
      sigma = dmin(0.8, dmax(pd.sigma_min, r * r * 4.0 + pd.sigma_min));
      mu = mu_new;
    }
    *iters_out = iters;
    if (failed || minstep) return E::FULL_FAILED;

    // ---------------- the step                                      [Engine::apply_step] ----------------
    const double ap = (converged ? 1.0 : alpha) * pd.step_length;
    for (int k = lane; k < NS; k += 32) {
      double dw[NW], lam[NR];
#pragma unroll
      for (int i = 0; i < NX; ++i) dw[i] = (k > 0) ? Kk[(k - 1) * NKK + i] : 0.0;
#pragma unroll
      for (int i = 0; i < NU; ++i) dw[NX + i] = (k < N) ? Kk[k * NKK + NX + i] : 0.0;
#pragma unroll
      for (int q = 0; q < NR; ++q) lam[q] = 0.0;
      if (has_rows(k)) {
        Bnd bd;
        double t[NR], lh[NR], th[NR];
        E::stage_bounds(pd, k, bd);
        ld_rows(LAM, k, lam);
        ld_rows(TT, k, t);
        ld_rows(LH, k, lh);
        ld_rows(TH, k, th);
#pragma unroll
        for (int q = 0; q < NR; ++q) {
          const bool act = E::row_active(bd, q);
          lam[q] = act ? lam[q] + alpha * (lh[q] - lam[q]) : 0.0;
          t[q] = act ? t[q] + alpha * (th[q] - t[q]) : 0.0;
          L.it[(size_t)(E::it_lam(N, k) + q) * bs] = lam[q];
          L.it[(size_t)(E::it_t(N, k) + q) * bs] = t[q];
        }
      }
      if (k > 0) {  // x-part of the costate recursion, c_k = q_k + (W dw)_x + (J' lam)_x, in place of q_k;  x_k += ap dx_k
        const double s = pd.scale[k];
        const double* W = Wc + (k == N ? 2 : 1) * NWS;
        double jl[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) jl[i] = 0.0;
#pragma unroll
        for (int r = NU; r < NV; ++r) E::jr_axpy(r, -lam[r] + lam[NV + r], jl);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          double a = Gk[k * NW + i];
#pragma unroll
          for (int jj = 0; jj < NW; ++jj) a += (s * W[E::pidx(i < jj ? i : jj, i < jj ? jj : i)]) * dw[jj];
          a += jl[i];
          Gk[k * NW + i] = a;
          L.it[(size_t)(E::it_x(N, k) + i) * bs] = XU[k * NW + i] + ap * dw[i];
        }
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < NU; ++i) L.it[(size_t)(E::it_u(N, k) + i) * bs] = XU[k * NW + NX + i] + ap * dw[NX + i];
      }
    }
    __syncwarp();
    // pi_{k-1} = c_k + A_k' pi_k, k = N .. 1
    double* cur = pv;
    double* nxt = pv2;
    if (lane < NX) cur[lane] = Gk[N * NW + lane];
    __syncwarp();
    for (int k = N - 1; k >= 0; --k) {
      if (lane < NX) {
        Kk[k * NKK + lane] = cur[lane];  // pi_k
        if (k > 0) {
          const double* Mc = Mk + k * NX * NC;
          double a = Gk[k * NW + lane];
#pragma unroll
          for (int l = 0; l < NX; ++l) a += Mc[l * NC + lane] * cur[l];
          nxt[lane] = a;
        }
      }
      __syncwarp();
      double* tmp = cur; cur = nxt; nxt = tmp;
    }
    for (int idx = lane; idx < N * NX; idx += 32) {
      const int k = idx / NX, i = idx - k * NX;
      L.it[(size_t)(E::it_pi(N, k) + i) * bs] = Kk[k * NKK + i];
    }
    if (lane == 0) L.it[(size_t)E::it_meta(N) * bs] = 1.0;
    return converged ? E::FULL_OK : E::FULL_MAXITER;
  }
};
#endif  // __CUDACC__

}  // namespace rlmpc
