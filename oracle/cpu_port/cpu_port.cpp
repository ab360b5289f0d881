// Host build of the engine maths (TEST INFRASTRUCTURE -- see oracle/__init__.py).
// Used (a) to debug the templates without a GPU, (b) as bench.py's cpu_baseline of kind "port":
// a structure-exploiting, OpenMP-over-samples CPU implementation, the fair "acados-like" CPU
// number of BASELINE.md section 4 (B2).  It is NOT an independent checker -- it shares
// engine.cuh with the CUDA product -- and nothing in mpc4rl_b200/ links or loads it.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../mpc4rl_b200/csrc/engine.cuh"
#include "../../mpc4rl_b200/csrc/condense.cuh"
#include "../../mpc4rl_b200/csrc/models/cartpole.cuh"
#include "../../mpc4rl_b200/csrc/models/linear_system.cuh"
#include "../../mpc4rl_b200/csrc/models/evaporation.cuh"

using namespace rlmpc;

static int g_threads = 0;  // 0 = all hardware threads
static int g_passes = 0;   // > 0: queued QPs go through that many single-trip passes (Engine::ipm_pass, what k_ipm_pass does)
                           // before the one-thread loop takes over

// queued QP: partially condensed (blocks of 4 stages) where the product does that too
template <class M>
static int qp_full_maybe_condensed(const ProblemData& pd, const Lane& L, int* ipm_iter) {
  using E = Engine<M>;
  if constexpr (M::NBX == 0 && M::NSX == 0 && M::NG == 0 && M::NX * 1 == 4) {
    using Cn = Condenser<M, 4>;
    if (pd.condense > 0 && Cn::applicable(pd)) {
      using EB = typename Cn::EB;
      const ProblemData pdb = Cn::block_pd(pd);
      std::vector<double> itb((size_t)EB::it_size(pdb.N) * TILE, 0.0), wsb((size_t)EB::ws_size(pdb.N) * TILE, 0.0);
      Lane Lb = L;
      Lb.it = itb.data();
      Lb.ws = wsb.data();
      for (int i = 0; i <= pdb.N; ++i) Cn::condense_block(pd, L, Lb, i);
      typename EB::DirectReader rd(Lb, pdb.N);
      return Cn::solve_expand(pd, pdb, L, Lb, ipm_iter, rd);
    }
  }
  return E::qp_full(pd, L, ipm_iter);
}

// Same pipeline as rlmpc_b200.cu, one sample at a time: K rounds of (linearise all stages |
// convergence test + fast QP | full interior point), one test-only round, then the sensitivities.
template <class M>
static void run(const ProblemData& pd0, int mode, int max_sqp, int B, const double* theta, int per_sample,
                const double* x0, const double* u0, double* iterate, int do_solve, int do_sens, double* u0_out,
                double* cost_out, int* status_out, double* dL, double* dpi, double* res_out, int* iters_out) {
  using E = Engine<M>;
  ProblemData pd = pd0;
  pd.mode = mode;
  pd.max_sqp = max_sqp;
  // tiled (AoSoA) storage like the device: element i of sample b at tile_off(b, n) + i*TILE
  const size_t bs = ((size_t)B + TILE - 1) / TILE * TILE;
  const int itn_ = E::it_size(pd.N);
  std::vector<double> ws((size_t)E::ws_size(pd.N) * bs, 0.0), itt((size_t)itn_ * bs, 0.0);
  std::vector<double> th((size_t)M::NTH * (per_sample ? bs : TILE)), ct((size_t)E::CT_SIZE * (per_sample ? bs : TILE), 0.0);
  for (int b = 0; b < (per_sample ? B : TILE); ++b) {
    for (int i = 0; i < M::NTH; ++i)
      th[tile_off(b, M::NTH) + (size_t)i * TILE] = per_sample ? theta[(size_t)b * M::NTH + i] : theta[i];
    M::cost_table(th.data() + tile_off(b, M::NTH), TILE, ct.data() + tile_off(b, E::CT_SIZE), TILE, pd.mc);
  }
  for (int b = 0; b < B; ++b)  // caller's iterate is [it_size][B] batch-minor
    for (int i = 0; i < itn_; ++i) itt[tile_off(b, itn_) + (size_t)i * TILE] = iterate[(size_t)i * B + b];
  auto body = [&](int b) {
    Lane L;
    L.it = itt.data() + tile_off(b, itn_);
    L.ws = ws.data() + tile_off(b, E::ws_size(pd.N));
    L.th = th.data() + (per_sample ? tile_off(b, M::NTH) : (size_t)(b % TILE));
    L.ct = ct.data() + (per_sample ? tile_off(b, E::CT_SIZE) : (size_t)(b % TILE));
    int status = ST_OK;
    double cost = 0.0;
    int sqp_iter = 0, ipm_iter = 0, fast_steps = 0;
    if (do_solve) {
      E::set_initial(pd, L, x0 + (size_t)b * M::NX, 1, u0 ? u0 + (size_t)b * M::NU : nullptr, 1);
      status = ST_MAXITER;
      const int K = pd.max_sqp, rounds = (K == 1) ? 1 : K + 1;
      for (int r = 0; r < rounds; ++r) {
        const bool last = (K > 1 && r == K);
        for (int k = 0; k <= pd.N; ++k) E::lin_stage(pd, L, k);
        typename E::Residuals R;
        int swept = 0;
        const int code = E::qp_fast(pd, L, R, &swept, /*polish=*/!last);
        cost = R.cost;
        if (code == E::FAST_NAN) { status = ST_NAN; break; }
        if (code == E::FAST_CONVERGED) { status = ST_OK; break; }
        if (last) break;
        if (code == E::FAST_STEPPED) {
          ++sqp_iter; ++ipm_iter; ++fast_steps;
          if (K == 1) { status = ST_OK; break; }
          continue;
        }
        int st = -1;
        if (g_passes > 0) {
          double state[E::IPM_STATE_WORDS];
          for (int p = 0; p < g_passes && st < 0; ++p) st = E::ipm_pass(pd, L, state, 1, p == 0, swept != 0, &ipm_iter);
        }
        if (st < 0) st = qp_full_maybe_condensed<M>(pd, L, &ipm_iter);
        ++sqp_iter;
        if (K == 1 || st == E::FULL_FAILED) { status = (st == E::FULL_OK) ? ST_OK : ST_QPFAIL; break; }
      }
      if (iters_out) {
        iters_out[3 * b] = sqp_iter;
        iters_out[3 * b + 1] = ipm_iter;
        iters_out[3 * b + 2] = fast_steps;
      }
    }
    if (do_sens) {
      int ok = 1;
      const int ng = E::grad_width(pd);
      for (int k = 0; k <= pd.N; ++k) E::sens_stage(pd, L, k);
      typename E::Residuals r = E::sens_sweep(pd, L, dL ? dL + (size_t)b * ng : nullptr,
                                              dpi ? dpi + (size_t)b * M::NU * ng : nullptr, &ok);
      cost = r.cost;
      if (res_out) {
        res_out[4 * b] = r.stat; res_out[4 * b + 1] = r.eq; res_out[4 * b + 2] = r.ineq; res_out[4 * b + 3] = r.comp;
      }
      const double rmax = dmax(dmax(r.stat, r.eq), dmax(r.ineq, r.comp));
      if (!(rmax == rmax)) status = ST_NAN;
      if (!do_solve) status = (rmax == rmax) ? (rmax < pd.tol ? ST_OK : ST_MAXITER) : ST_NAN;
      if (!ok && status == ST_OK) status = ST_QPFAIL;
    }
    if (u0_out)
      for (int i = 0; i < M::NU; ++i) u0_out[(size_t)b * M::NU + i] = L.it[(size_t)(E::it_u(pd.N, 0) + i) * TILE];
    if (cost_out) cost_out[b] = cost;
    if (status_out) status_out[b] = status;
  };
  // dynamic chunks of 16 samples over nthreads std::threads (no OpenMP runtime needed)
  int nt = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > (B + 15) / 16) nt = (B + 15) / 16;
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (;;) {
      const int s0 = next.fetch_add(16);
      if (s0 >= B) break;
      const int s1 = s0 + 16 < B ? s0 + 16 : B;
      for (int b = s0; b < s1; ++b) body(b);
    }
  };
  if (nt <= 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int i = 0; i < nt; ++i) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < itn_; ++i) iterate[(size_t)i * B + b] = itt[tile_off(b, itn_) + (size_t)i * TILE];
}

// model ids of the host port: 1 = cartpole (input bounds only), 2 = cartpole with state bounds, 3 = linear system,
// 4 = evaporation, 5 = cartpole with g as fourth model parameter
#define PORT_DISPATCH(model, expr)                         \
  switch (model) {                                         \
    case 1: { using M = CartpoleModel; expr; } break;      \
    case 2: { using M = CartpoleModelBX; expr; } break;    \
    case 3: { using M = LinearSystemModel; expr; } break;  \
    case 4: { using M = EvaporationModel; expr; } break;   \
    case 5: { using M = CartpoleModelG; expr; } break;     \
    default: return -1;                                    \
  }

extern "C" {

void cpu_port_set_threads(int n) { g_threads = n; }
int cpu_port_get_threads() { return g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency(); }

void cpu_port_set_passes(int n) { g_passes = n; }

int cpu_port_sizeof_problem_data() { return (int)sizeof(ProblemData); }

int cpu_port_iterate_size(int model, int N) {
  PORT_DISPATCH(model, return Engine<M>::it_size(N));
  return -1;
}

int cpu_port_nrows(int model) {
  PORT_DISPATCH(model, return Engine<M>::NR);
  return -1;
}

int cpu_port_grad_width(int model, const ProblemData* pd) {
  PORT_DISPATCH(model, return Engine<M>::grad_width(*pd));
  return -1;
}

// iterate: [it_size][B] batch-minor, in/out (zero it + set x rows for a cold start); iters_out: [B,3]
int cpu_port_unit(int model, const ProblemData* pd, int mode, int max_sqp, int B, const double* theta, int per_sample,
                  const double* x0, const double* u0, double* iterate, int do_solve, int do_sens, double* u0_out,
                  double* cost_out, int* status_out, double* dL, double* dpi, double* res_out, int* iters_out) {
  PORT_DISPATCH(model, run<M>(*pd, mode, max_sqp, B, theta, per_sample, x0, u0, iterate, do_solve, do_sens, u0_out,
                              cost_out, status_out, dL, dpi, res_out, iters_out));
  return 0;
}
}
