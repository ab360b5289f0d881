// Host build of the warp-cooperative chain-mass engine (TEST INFRASTRUCTURE -- see oracle/__init__.py).
// The kernel bodies of mpc4rl_b200/csrc/chain/chain_engine.cuh run here under a fiber emulation of a warp
// (simt_host.h): used to debug the maths without a GPU and as bench.py's cpu_baseline of kind "port" for the
// chain-mass workload.  NOT an independent checker (shares the engine with the CUDA product); nothing in
// mpc4rl_b200/ links or loads it.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "simt_host.h"
#include "../../mpc4rl_b200/csrc/chain/chain_engine.cuh"

using namespace rlmpc;

template <int NM>
static void run(const ProblemData& pd0, int mode, int max_sqp, int B, const double* theta, const double* xss, const double* x0,
                const double* u0, double* iterate, int do_solve, int do_sens, double* u0_out, double* cost_out, int* status_out,
                double* dL, double* dpi, double* res_out, int* iters_out, int threads) {
  using E = ChainEngine<NM>;
  using Mo = ChainModel<NM>;
  ProblemData pd = pd0;
  pd.mode = mode;
  pd.max_sqp = max_sqp;
  const int N = pd.N, IT = E::it_size(N);
  std::vector<double> tab(E::TB_SIZE);
  for (int i = 0; i < E::NX; ++i)
    for (int j = 0; j < E::NX; ++j)
      tab[E::TB_Q + i * E::NX + j] = 0.5 * (theta[Mo::TH_Q + i + j * E::NX] + theta[Mo::TH_Q + j + i * E::NX]);
  for (int i = 0; i < E::NU; ++i)
    for (int j = 0; j < E::NU; ++j)
      tab[E::TB_R + i * E::NU + j] = 0.5 * (theta[Mo::TH_R + i + j * E::NU] + theta[Mo::TH_R + j + i * E::NU]);
  for (int i = 0; i < E::NX; ++i) tab[E::TB_XSS + i] = xss[i];
  std::vector<int> status(B, ST_MAXITER), work(B, WK_ACTIVE);
  std::vector<double> cost(B, 0.0);
  std::atomic<int> next{0};
  auto worker = [&]() {
    std::vector<double> ws(E::ws_size(N), 0.0);
    std::vector<double> Sst(E::SM_STAGE), Sqp(E::qp_smem_doubles(N) + 8), Ssn(E::sens_smem_doubles(N) + 8);
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      ChainArgs a;
      memset(&a, 0, sizeof(a));
      // the arrays are indexed by sample: shift the bases so that sample b lands on this worker's private buffers
      a.it = iterate;
      a.ws = ws.data() - (size_t)b * E::ws_size(N);
      a.th = theta; a.tab = tab.data(); a.B = B;
      a.status = status.data(); a.work = work.data(); a.cost = cost.data();
      a.u0_out = u0_out; a.cost_out = cost_out; a.status_out = status_out; a.dL = dL; a.dpi = dpi; a.res_out = res_out;
      double* it = iterate + (size_t)b * IT;
      int ipm_total = 0;
      if (do_solve) {
        for (int i = 0; i < E::NX; ++i) it[E::it_x(N, 0) + i] = x0[(size_t)b * E::NX + i];
        if (mode == MODE_Q)
          for (int i = 0; i < E::NU; ++i) it[E::it_u(N, 0) + i] = u0[(size_t)b * E::NU + i];
        const int rounds = (max_sqp == 1) ? 1 : max_sqp + 1;
        for (int r = 0; r < rounds && work[b] == WK_ACTIVE; ++r) {
          a.last_round = (max_sqp > 1 && r == max_sqp) ? 1 : 0;
          for (int k = 0; k <= N; ++k)
            simt::run(32, [&](int lane) { E::template stage_task<false>(pd, a, b, k, Sst.data(), lane); });
          int res = 0, iters = 0;
          simt::run(32, [&](int lane) {
            StageFeed feed;
            E::qp_feed_init(feed, Sqp.data(), N, lane);
            int it_ = 0;
            const int r_ = E::qp_sample(pd, a, b, Sqp.data(), feed, lane, &it_);
            if (lane == 0) { res = r_; iters = it_; }
          });
          ipm_total += iters;
          if (res == E::R_NAN) { status[b] = ST_NAN; work[b] = WK_DONE; }
          else if (res == E::R_CONVERGED) { status[b] = ST_OK; work[b] = WK_DONE; }
          else if (res == E::R_TESTONLY) { work[b] = WK_DONE; }
          else if (res == E::R_FAILED) { status[b] = ST_QPFAIL; work[b] = WK_DONE; }
          else if (max_sqp == 1) { status[b] = (res == E::R_STEPPED) ? ST_OK : ST_QPFAIL; work[b] = WK_DONE; }
        }
      }
      if (iters_out) iters_out[b] = ipm_total;
      a.have_solve = do_solve;
      if (do_sens) {
        for (int k = 0; k <= N; ++k)
          simt::run(32, [&](int lane) { E::template stage_task<true>(pd, a, b, k, Sst.data(), lane); });
        simt::run(32, [&](int lane) {
          StageFeed feed;
          E::sens_feed_init(feed, Ssn.data(), N, lane);
          E::sens_sample(pd, a, b, Ssn.data(), feed, lane);
        });
        if (dpi && mode == MODE_V) {
          std::vector<double> part(E::NPD);
          for (int r = 0; r < E::NU; ++r) {
            std::vector<double> acc(E::NPD, 0.0);
            for (int k = 0; k < N; ++k) {
              E::param_task(pd, a, b, k, r, part.data());
              for (int p = 0; p < E::NPD; ++p) acc[p] += part[p];
            }
            for (int p = 0; p < E::NPD; ++p) dpi[((size_t)b * E::NU + r) * E::NTH + Mo::pd_to_theta(p)] = acc[p];
          }
        }
      } else {
        if (u0_out) for (int i = 0; i < E::NU; ++i) u0_out[(size_t)b * E::NU + i] = it[E::it_u(N, 0) + i];
        if (cost_out) cost_out[b] = cost[b];
        if (status_out) status_out[b] = status[b];
      }
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt > B) nt = B;
  if (nt < 1) nt = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
}

extern "C" {
int chain_port_sizeof_problem_data() { return (int)sizeof(ProblemData); }
int chain_port_dims(int n_mass, int N, int* nx, int* nth, int* it_size) {
#define DIMS(NM)                                                                                        \
  case NM: *nx = ChainEngine<NM>::NX; *nth = ChainEngine<NM>::NTH; *it_size = ChainEngine<NM>::it_size(N); return 0;
  switch (n_mass) { DIMS(3) DIMS(5) DIMS(6) }
#undef DIMS
  return -1;
}
// iterate: [B][it_size] in/out (zeros + x_k = x0 for a cold start, like MPC.reset)
int chain_port_run(int n_mass, const ProblemData* pd, int mode, int max_sqp, int B, const double* theta, const double* xss,
                   const double* x0, const double* u0, double* iterate, int do_solve, int do_sens, double* u0_out,
                   double* cost_out, int* status_out, double* dL, double* dpi, double* res_out, int* iters_out, int threads) {
#define RUN(NM)                                                                                                        \
  case NM: run<NM>(*pd, mode, max_sqp, B, theta, xss, x0, u0, iterate, do_solve, do_sens, u0_out, cost_out, status_out, dL, dpi, \
                   res_out, iters_out, threads); return 0;
  switch (n_mass) { RUN(3) RUN(5) RUN(6) }
#undef RUN
  return -1;
}
}
