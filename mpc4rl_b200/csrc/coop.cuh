// Warp-per-sample interior-point solve of the queued stage QPs (CUDA only).
//
// The queue (samples whose warm Newton iteration of qp_fast() was not enough) is a small part of the
// batch, but each of its solves is a long dependent chain: interior-point iterations x 2 sweeps x N
// stages.  With one thread per sample (Engine::qp_full / Condenser::solve_expand) the kernel lasts as
// long as its slowest sample, ~100 us per iteration.  Here ONE WARP owns one sample:
//   * the whole stage QP (A, B, b, q, r of every stage, the multipliers and slacks of the rows) is
//     staged once in shared memory (50 doubles per stage),
//   * everything that is independent between stages (row updates, barrier terms, new slacks and
//     multipliers, step-length statistics, the final step) runs with one lane per stage,
//   * the two Riccati recursions run with one lane per matrix entry: [P A  P B  P b + p] on NX x (NW+1)
//     lanes, then [A B]'[...] + cost on NW x (NW+1) lanes, exchanged through shared memory,
//   * all control flow (warm start, active-set steps, cold restart, exits) is warp-uniform.
// The iteration is the one of Engine::qp_ipm / apply_step, statement by statement (same operation order
// inside every dot product), so both paths walk through the same iterates up to the rounding of the
// reductions over the rows.  Applicable to input-bounds-only problems with NU = 1 and NX <= 4 (cart-pole,
// config/cartpole_original.yaml); everything else keeps the thread-per-sample queue kernels.
#pragma once

#include "engine.cuh"

namespace rlmpc {

template <class M>
struct CoopOK {
  static constexpr int NW = M::NX + M::NU;
  static constexpr bool value = M::NU == 1 && M::NBX == 0 && M::NSX == 0 && M::NG == 0 && !M::STAGE_HESS &&
                                M::NX * (NW + 1) <= 32 && NW * (NW + 1) <= 32;
};

#ifdef __CUDACC__
// warp reductions and plain shared-window accesses used by the cooperative solvers
struct CoopQPBase {
  __device__ static __forceinline__ double wsum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
  }
  __device__ static __forceinline__ double wmin(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
  }
  __device__ static __forceinline__ double wmax(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = dmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
  }
  __device__ static __forceinline__ double lds(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
  }
  __device__ static __forceinline__ void sts(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
  }
};

template <class M>
struct CoopQP {
  using E = Engine<M>;
  static constexpr int NX = M::NX, NU = M::NU, NW = NX + NU, NC = NW + 1, NWS = E::NWS;
  static constexpr int PER_STAGE = NX * NC + NW + NW + 7;  // Mk, g, K|kff (later dx, du), 7 row/scalar fields
  static constexpr int SCRATCH = 2 * NX + NX * NC + NW * NC + 3 * NWS + 1;
  __host__ __device__ static constexpr int smem_doubles(int N) { return PER_STAGE * (N + 1) + SCRATCH; }

  __device__ static __forceinline__ double wsum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
  }
  __device__ static __forceinline__ double wmin(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
  }
  __device__ static __forceinline__ double wmax(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = dmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
  }

  __device__ static __forceinline__ double lds(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
  }
  __device__ static __forceinline__ void sts(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
  }

  // One sample, executed by the 32 lanes of a warp.  S: this warp's shared memory (smem_doubles(N)).
  // Returns Engine::FULL_OK / FULL_MAXITER / FULL_FAILED; the step is applied to the iterate in global
  // memory (L.it) unless FULL_FAILED.
  __device__ static int solve(const ProblemData& pd, const Lane& L, double* S, const int lane, const bool swept, int* iters_out) {
    const int N = pd.N, NS = N + 1;
    constexpr size_t bs = TILE;
    double* Mk = S;                  // [k][NX][NC]  = [A | B | b]
    double* Gk = Mk + NS * NX * NC;  // [k][NW]      = [q ; r]      (later: x-part of the costate recursion)
    // [k][NW] = [K_k | kff_k] after the backward sweep; the forward sweep overwrites stage k's slot with
    // [dx_{k+1} | du_k] once the feedback law of stage k is in registers (dx_0 = 0); finally pi_k
    double* Kk = Gk + NS * NW;
    double* U = Kk + NS * NW;
    double* LL = U + NS;    // lam of the lower / upper input bound
    double* LU = LL + NS;
    double* TL = LU + NS;   // slacks
    double* TU = TL + NS;
    double* LHL = TU + NS;  // lam_hat, t_hat of the last Newton step
    double* LHU = LHL + NS;
    // (t_hat is not stored: it is the distance to the bound after the step, th_l / th_u below, from u_k and du_k)
    // (u,u) Hessian entry and u-gradient of the stage incl. barrier terms: live from the row phase to the end
    // of the backward sweep, when lam_hat of the previous step is dead, so they share its storage
    double* HB = LHL;
    double* GB = LHU;
    double* pv = LHU + NS;  // scratch
    double* pv2 = pv + NX;
    double* PM = pv2 + NX;
    double* T = PM + NX * NC;
    double* Wc = T + NW * NC;
    double* ZERO = Wc + 3 * NWS;  // one 0.0

    // explicit shared-window addresses for the two recursions (plain ld/st.shared with immediate offsets)
    const unsigned sT = (unsigned)__cvta_generic_to_shared(T), sPM = (unsigned)__cvta_generic_to_shared(PM);
    const unsigned sKk = (unsigned)__cvta_generic_to_shared(Kk);

    const bool qmode = pd.mode == MODE_Q;
    const double lb = pd.lbu[0], ub = pd.ubu[0];
    const bool has_l = lb > -BIG, has_u = ub < BIG;
    const double range = (has_l && has_u) ? ub - lb : 1.0;
    const int k_first = qmode ? 1 : 0;  // stages [k_first, N) carry the input rows
    const double m_rows = (double)((N - k_first) * ((has_l ? 1 : 0) + (has_u ? 1 : 0)));
    // t_hat of the last Newton step is not stored: it is the slack of the bound at u_k + du_k, (u_k - lb) + du_k resp.
    // (ub - u_k) - du_k, and du_k sits in the feedback-law slot of stage k from the forward sweep until the next
    // backward sweep (rows that do not exist: 0)

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
        row = e - NX * NX; col = NX;
      } else {
        row = e - NX * NX - NX * NU; col = NX + 1;
      }
      Mk[k * NX * NC + row * NC + col] = v;
    }
#pragma unroll 4  // several of the scattered global loads in flight per lane
    for (int idx = lane; idx < NS * NW; idx += 32) {
      const int k = idx / NW, e = idx - k * NW;
      Gk[idx] = (k < N || e < NX) ? L.ws[((size_t)k * E::W_REC + E::W_q + e) * bs] : 0.0;
    }
    for (int k = lane; k < NS; k += 32) {
      U[k] = (k < N) ? L.it[(size_t)E::it_u(N, k) * bs] : 0.0;
      LL[k] = L.it[(size_t)E::it_lam(N, k) * bs];
      LU[k] = L.it[(size_t)(E::it_lam(N, k) + 1) * bs];
      TL[k] = L.it[(size_t)E::it_t(N, k) * bs];
      TU[k] = L.it[(size_t)(E::it_t(N, k) + 1) * bs];
      LHL[k] = LHU[k] = 0.0;
    }
    for (int idx = lane; idx < 3 * NWS; idx += 32) {
      const int kind = idx / NWS, e = idx - kind * NWS;
      Wc[idx] = L.ct[(size_t)(kind * E::CT_REC + E::CT_W + e) * bs];
    }
    if (lane == 0) ZERO[0] = 0.0;
    bool warm = pd.warm_ipm && L.it[(size_t)E::it_meta(N) * bs] > 0.5;
    __syncwarp();

    // (lam,t) of a warm / cold start; returns sum(lam*t)           [Engine::ipm_init]
    auto init_rows = [&](bool w) -> double {
      double mu = 0.0;
      for (int k = lane; k < NS; k += 32) {
        const bool act = k >= k_first && k < N;
        double ll, lu, tl, tu;
        if (w) {
          tl = dmax(TL[k], 1e-10 * range); tu = dmax(TU[k], 1e-10 * range);
          ll = dmax(LL[k], 1e-14); lu = dmax(LU[k], 1e-14);
        } else {
          const double tmin = 1e-2 * range, u = U[k];
          tl = dmax(u - lb, tmin); ll = pd.mu0 / tl;
          tu = dmax(ub - u, tmin); lu = pd.mu0 / tu;
        }
        if (act && has_l) mu += ll * tl; else { ll = 0.0; tl = 0.0; }
        if (act && has_u) mu += lu * tu; else { lu = 0.0; tu = 0.0; }
        LL[k] = ll; LU[k] = lu; TL[k] = tl; TU[k] = tu;
      }
      __syncwarp();
      return wsum(mu);
    };

    // ---- per-lane roles of the two recursions (constant over the solve) ----
    const int iA = lane / NC, cA = lane - iA * NC;  // phase A: entry (iA, cA) of [P A | P B | v]; phase B: entry (iA, cA) of T
    const bool inA = lane < NX * NC, inB = lane < NW * NC;
    const int iAx = iA < NX ? iA : NX - 1;          // lanes beyond the roles compute on valid addresses, never store
    const int k_last = qmode ? 1 : 0;               // last stage with a feedback law
    const unsigned sMk = (unsigned)__cvta_generic_to_shared(Mk);
    const unsigned sMA = sMk + 8 * cA, sMB = sMk + 8 * (iA < NW ? iA : 0);
    unsigned oT[NX];
#pragma unroll
    for (int l = 0; l < NX; ++l) oT[l] = 8 * ((iAx < l ? iAx : l) * NC + (iAx < l ? l : iAx));
    double wB0 = 0.0, wB1 = 0.0;  // cost Hessian entry (stage 0 / stages 1..N-1), unscaled
    unsigned sBase = (unsigned)__cvta_generic_to_shared(ZERO), baseStride = 0;
    if (inB) {
      if (cA < NW && !(iA == NX && cA == NX)) {
        wB0 = Wc[0 * NWS + E::pidx(iA < cA ? iA : cA, iA < cA ? cA : iA)];
        wB1 = Wc[1 * NWS + E::pidx(iA < cA ? iA : cA, iA < cA ? cA : iA)];
      } else if (cA < NW) {  // (u,u): scaled cost + barrier, written by the row phase
        sBase = (unsigned)__cvta_generic_to_shared(HB); baseStride = 8;
      } else if (iA < NX) {  // gradient wrt x
        sBase = (unsigned)__cvta_generic_to_shared(Gk + iA); baseStride = 8 * NW;
      } else {               // gradient wrt u + barrier
        sBase = (unsigned)__cvta_generic_to_shared(GB); baseStride = 8;
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
        warm = false;  // jammed warm start: restart cold
        mu = init_rows(false) / m_rows;
        alpha = 0.0;
        sigma = pd.sigma0;
      }
      if (warm) ++warm_iters;
      const double target = dmax(sigma * mu, pd.tau);
      // first iteration of a warm start at target tau: k_qp1 (Engine::qp_fast) has just done exactly this
      // Newton iteration and left lam_hat, t_hat in the stage records; pick them up instead of repeating it
      const bool reuse = swept && j == 0 && warm && target == pd.tau && pd.max_ipm > 1;
      if (!reuse) {
      // ---- rows: pending damped update, barrier terms (one lane per stage) ----
      for (int k = lane; k < NS; k += 32) {
        double ll = LL[k], lu = LU[k], tl = TL[k], tu = TU[k];
        if (alpha > 0.0) {
          const bool row = k >= k_first && k < N;
          const double duk = row ? Kk[k * NW + NX] : 0.0, uk = U[k];
          const double thl = (row && has_l) ? (uk - lb) + duk : 0.0, thu = (row && has_u) ? (ub - uk) - duk : 0.0;
          ll += alpha * (LHL[k] - ll); tl += alpha * (thl - tl);
          lu += alpha * (LHU[k] - lu); tu += alpha * (thu - tu);
          LL[k] = ll; LU[k] = lu; TL[k] = tl; TU[k] = tu;
        }
        const bool act = k >= k_first && k < N;
        // (u,u) Hessian entry and u-gradient of the stage with the barrier terms   [Engine::barrier_add]
        double hb = pd.scale[k] * Wc[(k == 0 ? 0 : 1) * NWS + E::pidx(NX, NX)], gb = Gk[k * NW + NX];
        const double u = U[k];
        if (act && has_l) {
          const double itb = 1.0 / tl, cb = ll * itb, ab = target * itb + ll;
          hb += cb;
          gb += -(ab - cb * (u - lb));
        }
        if (act && has_u) {
          const double itb = 1.0 / tu, cb = lu * itb, ab = target * itb + lu;
          hb += cb;
          gb += ab - cb * (ub - u);
        }
        HB[k] = hb; GB[k] = gb;
      }
      __syncwarp();
      // ---- backward Riccati sweep ----
      // Two exchanges per stage through shared memory.  T holds stage k+1's [A B]'[P A | P B | v] + cost
      // (rows 0..NX-1: the x-block and gradient, row NX: H, G, gv); phase A of stage k finishes it on the
      // fly, P = T_xx + H'K, p = T_x5 + H'kff with K = -H/G, and multiplies by [A B b] of stage k; phase B
      // forms stage k's T.  The stage data of the next stage is fetched into registers ahead of the chain.
      if (lane < NW * NC) {  // terminal "T": P_N = scaled W_e, p_N = q_N, no input
        double v = 0.0;
        if (iA < NX && cA < NX) v = pd.scale[N] * Wc[2 * NWS + E::pidx(iA < cA ? iA : cA, iA < cA ? cA : iA)];
        if (iA < NX && cA == NC - 1) v = Gk[N * NW + iA];
        if (iA == NX && cA == NX) v = 1.0;
        T[lane] = v;
      }
      // stage data of stage k for this lane's two roles: column cA of [A B b] (phase A), column iA (phase B),
      // and the cost / gradient term of entry (iA, cA) of T:  scale_k * w + (per-lane array)[k]
      double mA[NX], mB[NX], baseB;
      auto fetch = [&](int k) {
        const unsigned ma = sMA + (unsigned)k * (8 * NX * NC), mb = sMB + (unsigned)k * (8 * NX * NC);
#pragma unroll
        for (int l = 0; l < NX; ++l) {
          mA[l] = lds(ma + 8 * l * NC);
          mB[l] = lds(mb + 8 * l * NC);
        }
        baseB = pd.scale[k] * (k == 0 ? wB0 : wB1) + lds(sBase + (unsigned)k * baseStride);
      };
      fetch(N - 1);
      __syncwarp();
      for (int k = N - 1; k >= k_last; --k) {
        // ---- phase A (all lanes compute, lanes (i, c) with i < NX store) ----
        const double G = lds(sT + 8 * (NX * NC + NX));
        if (!(G > 0.0)) failed = true;
        const double inv = 1.0 / G;
        const double hi = lds(sT + 8 * NX * NC + 8 * iAx);
        double acc = 0.0;
        if (cA == NC - 1) {
          const double kff = -lds(sT + 8 * (NX * NC + NC - 1)) * inv;
          acc = lds(sT + 8 * NC * iAx + 8 * (NC - 1)) + hi * kff;
        }
#pragma unroll
        for (int l = 0; l < NX; ++l) {
          // P_il = T_ab - (H_i H_l) / G, (a, b) = (min, max) of (i, l): bitwise symmetric
          const double hl = lds(sT + 8 * (NX * NC + l));
          const double Pil = lds(sT + oT[l]) - (hi * hl) * inv;
          acc += Pil * mA[l];
        }
        if (inA) sts(sPM + 8 * lane, acc);
        if (lane >= NX * NC && lane < NX * NC + NW && k + 1 < N) {  // feedback law of stage k+1
          const int c = lane - NX * NC;
          Kk[(k + 1) * NW + c] = -lds(sT + 8 * (NX * NC + (c < NX ? c : NC - 1))) * inv;
        }
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
      {  // feedback law of the last stage; u_0 fixed (Q-mode): none
        const double G = lds(sT + 8 * (NX * NC + NX));
        if (!(G > 0.0)) failed = true;
        const double inv = 1.0 / G;
        if (lane < NW) {
          Kk[k_last * NW + lane] = -lds(sT + 8 * (NX * NC + (lane < NX ? lane : NC - 1))) * inv;
          if (qmode) Kk[lane] = 0.0;
        }
      }
      // ---- forward sweep: dx, du ----
      __syncwarp();
      {
        double kr[NW], ar[NC];
        // (shared-window loads with immediate offsets, like the backward sweep; lanes >= NX read row 0 and never store)
        const unsigned sMrow = sMk + 8 * ((lane < NX ? lane : 0) * NC);
        auto fetch_f = [&](int k) {
          const unsigned kk = sKk + (unsigned)k * (8 * NW), mm = sMrow + (unsigned)k * (8 * NX * NC);
#pragma unroll
          for (int l = 0; l < NW; ++l) kr[l] = lds(kk + 8 * l);
#pragma unroll
          for (int l = 0; l < NC; ++l) ar[l] = lds(mm + 8 * l);
        };
        fetch_f(0);
        __syncwarp();  // every lane holds stage 0's feedback law before its slot is overwritten
        for (int k = 0; k < N; ++k) {
          double dxl[NX];
#pragma unroll
          for (int l = 0; l < NX; ++l) dxl[l] = (k > 0) ? lds(sKk + 8 * ((k - 1) * NW + l)) : 0.0;
          double du = kr[NX];
#pragma unroll
          for (int l = 0; l < NX; ++l) du += kr[l] * dxl[l];
          double a = ar[NC - 1];
#pragma unroll
          for (int l = 0; l < NX; ++l) a += ar[l] * dxl[l];
          a += ar[NX] * du;
          if (k + 1 < N) fetch_f(k + 1);
          if (lane < NX) sts(sKk + 8 * (k * NW + lane), a);  // dx_{k+1}
          if (lane == NX) sts(sKk + 8 * (k * NW + NX), du);   // du_k
          __syncwarp();
        }
      }
      }  // !reuse
      // ---- rows: new slacks and multipliers, step-length statistics (one lane per stage) ----
      double amax = 1e300, s0 = 0.0, s1 = 0.0, s2 = 0.0, cmax = 0.0;
      for (int k = lane; k < NS; k += 32) {
        const bool act = k >= k_first && k < N;
        if (reuse && k < N) Kk[k * NW + NX] = L.ws[((size_t)k * E::W_REC + E::W_du) * bs];  // du_k of k_qp1's Newton iteration
        const double dv = (k < N) ? Kk[k * NW + NX] : 0.0, u = U[k];
        double lhl = 0.0, thl = 0.0, lhu = 0.0, thu = 0.0;
        if (act && has_l) {
          const double ll = LL[k], tl = TL[k];
          thl = (u - lb) + dv;
          if (reuse) {
            lhl = L.ws[((size_t)k * E::W_REC + E::W_lh) * bs];
          } else {
            const double itb = 1.0 / tl, cb = ll * itb, ab = target * itb + ll;
            lhl = ab - cb * thl;
          }
          const double dt = thl - tl, dl = lhl - ll;
          if (dt < 0.0) amax = dmin(amax, -tl / dt);
          if (dl < 0.0) amax = dmin(amax, -ll / dl);
          s0 += ll * tl; s1 += ll * dt + tl * dl; s2 += dl * dt;
          cmax = dmax(cmax, dabs(dl * dt));
        }
        if (act && has_u) {
          const double lu = LU[k], tu = TU[k];
          thu = (ub - u) - dv;
          if (reuse) {
            lhu = L.ws[((size_t)k * E::W_REC + E::W_lh + 1) * bs];
          } else {
            const double itb = 1.0 / tu, cb = lu * itb, ab = target * itb + lu;
            lhu = ab - cb * thu;
          }
          const double dt = thu - tu, dl = lhu - lu;
          if (dt < 0.0) amax = dmin(amax, -tu / dt);
          if (dl < 0.0) amax = dmin(amax, -lu / dl);
          s0 += lu * tu; s1 += lu * dt + tu * dl; s2 += dl * dt;
          cmax = dmax(cmax, dabs(dl * dt));
        }
        LHL[k] = lhl; LHU[k] = lhu;
      }
      __syncwarp();
      {  // NaN anywhere in the step must reach amax like in the scalar code (a NaN slack fails "dt < 0")
        const double probe = reuse ? 0.0 : Kk[NX] + Kk[(N - 1) * NW];
        if (!(probe == probe)) amax = probe;
      }
      const bool nan_step = __any_sync(0xffffffffu, !(amax == amax));
      amax = wmin(amax); s0 = wsum(s0); s1 = wsum(s1); s2 = wsum(s2); cmax = wmax(cmax);
      failed = __any_sync(0xffffffffu, failed);

      if ((failed || nan_step) && warm) {  // numerically broken warm start: not an error, start over cold
        warm = false;
        failed = false;
        mu = init_rows(false) / m_rows;
        alpha = 0.0;
        sigma = pd.sigma0;
        continue;
      }
      if (failed || nan_step) { failed = true; break; }
      if (warm && amax < 1.0 / 0.995 && as_iters < (int)pd.as_steps) {
        // active-set step of a warm start                         [Engine::ipm_project]
        ++as_iters;
        double m2 = 0.0;
        const double eps_t = 1e-9 * range;
        for (int k = lane; k < NS; k += 32) {
          const bool act = k >= k_first && k < N;
          if (act && has_l) {
            double ll = LL[k], tl;
            const double lh = LHL[k], th = (U[k] - lb) + Kk[k * NW + NX];
            if (!(th > eps_t)) { ll = dmax(dmax(lh, ll), 1e-3); tl = dmin(eps_t, pd.tau / ll); }
            else if (!(lh > 0.0)) { tl = dmax(th, AS_RELEASE * range); ll = pd.tau / tl; }
            else { tl = th; ll = lh; }
            LL[k] = ll; TL[k] = tl;
            m2 += ll * tl;
          }
          if (act && has_u) {
            double lu = LU[k], tu;
            const double lh = LHU[k], th = (ub - U[k]) - Kk[k * NW + NX];
            if (!(th > eps_t)) { lu = dmax(dmax(lh, lu), 1e-3); tu = dmin(eps_t, pd.tau / lu); }
            else if (!(lh > 0.0)) { tu = dmax(th, AS_RELEASE * range); lu = pd.tau / tu; }
            else { tu = th; lu = lh; }
            LU[k] = lu; TU[k] = tu;
            m2 += lu * tu;
          }
        }
        __syncwarp();
        mu = wsum(m2) / m_rows;
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
      // (not on the picked-up iteration: k_qp1 has declined exactly this step, and its dx is not in shared memory)
      if (!reuse && target <= pd.tau && alpha == 1.0 && cmax <= dmin(pd.comp_accept * pd.tau, 0.1 * pd.tol)) converged = true;
      const double r = 1.0 - alpha;
      sigma = dmin(0.8, dmax(pd.sigma_min, r * r * 4.0 + pd.sigma_min));
      mu = mu_new;
    }
    *iters_out = iters;
    if (failed || minstep) return E::FULL_FAILED;

    // ---------------- the step                                      [Engine::apply_step] ----------------
    const double ap = (converged ? 1.0 : alpha) * pd.step_length;  // iteration limit: damped primal step
    for (int k = lane; k < NS; k += 32) {
      const bool act = k >= k_first && k < N;
      const double ll = (act && has_l) ? LL[k] + alpha * (LHL[k] - LL[k]) : 0.0;
      const double duk = act ? Kk[k * NW + NX] : 0.0;
      const double tl = (act && has_l) ? TL[k] + alpha * (((U[k] - lb) + duk) - TL[k]) : 0.0;
      const double lu = (act && has_u) ? LU[k] + alpha * (LHU[k] - LU[k]) : 0.0;
      const double tu = (act && has_u) ? TU[k] + alpha * (((ub - U[k]) - duk) - TU[k]) : 0.0;
      if (k < N) {
        L.it[(size_t)E::it_lam(N, k) * bs] = ll;
        L.it[(size_t)(E::it_lam(N, k) + 1) * bs] = lu;
        L.it[(size_t)E::it_t(N, k) * bs] = tl;
        L.it[(size_t)(E::it_t(N, k) + 1) * bs] = tu;
        L.it[(size_t)E::it_u(N, k) * bs] = U[k] + ap * Kk[k * NW + NX];
      }
    }
    // x-part of the costate recursion: c_k = q_k + (W dw)_x, in place of q_k;  x_k += ap dx_k  (k >= 1)
    for (int idx = lane; idx < NS * NX; idx += 32) {
      const int k = idx / NX, i = idx - k * NX;
      if (k == 0) continue;
      const double s = pd.scale[k];
      const double* W = Wc + (k == N ? 2 : 1) * NWS;
      double a = Gk[k * NW + i];
#pragma unroll
      for (int jj = 0; jj < NW; ++jj) {
        const double dwj = jj < NX ? Kk[(k - 1) * NW + jj] : (k < N ? Kk[k * NW + NX] : 0.0);
        a += (s * W[E::pidx(i < jj ? i : jj, i < jj ? jj : i)]) * dwj;
      }
      Gk[k * NW + i] = a;
      const size_t o = (size_t)(E::it_x(N, k) + i) * bs;
      L.it[o] = L.it[o] + ap * Kk[(k - 1) * NW + i];
    }
    __syncwarp();
    // pi_{k-1} = c_k + A_k' pi_k, k = N .. 1   (pi_k kept in Kk[k][0..NX), ping-pong through pv / pv2)
    double* cur = pv;
    double* nxt = pv2;
    if (lane < NX) cur[lane] = Gk[N * NW + lane];
    __syncwarp();
    for (int k = N - 1; k >= 0; --k) {
      if (lane < NX) {
        Kk[k * NW + lane] = cur[lane];  // pi_k
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
      L.it[(size_t)(E::it_pi(N, k) + i) * bs] = Kk[k * NW + i];
    }
    if (lane == 0) L.it[(size_t)E::it_meta(N) * bs] = 1.0;
    return converged ? E::FULL_OK : E::FULL_MAXITER;
  }
};
#endif  // __CUDACC__

}  // namespace rlmpc
