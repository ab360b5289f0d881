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
#include "../../mpc4rl_b200/csrc/models/cartpole.cuh"

using namespace rlmpc;

static int g_threads = 0;  // 0 = all hardware threads

template <class M>
static void run(const ProblemData& pd0, int mode, int max_sqp, int B, const double* theta, int per_sample,
                const double* x0, const double* u0, double* iterate, int do_solve, int do_sens, double* u0_out,
                double* cost_out, int* status_out, double* dL, double* dpi, double* res_out, int* iters_out) {
  using E = Engine<M>;
  ProblemData pd = pd0;
  pd.mode = mode;
  pd.max_sqp = max_sqp;
  const size_t bs = (size_t)B;
  std::vector<double> ws((size_t)E::ws_size(pd.N) * bs, 0.0);
  std::vector<double> th;
  if (per_sample) {
    th.resize((size_t)M::NTH * bs);
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < M::NTH; ++i) th[(size_t)i * bs + b] = theta[(size_t)b * M::NTH + i];
  }
  auto body = [&](int b) {
    Lane L;
    L.it = iterate + b;
    L.ws = ws.data() + b;
    L.bs = bs;
    L.th = per_sample ? th.data() + b : theta;
    L.ths = per_sample ? bs : 1;
    int status = 0;
    double cost = 0.0;
    if (do_solve) {
      E::set_initial(pd, L, x0 + (size_t)b * M::NX, 1, u0 ? u0 + (size_t)b * M::NU : nullptr, 1);
      typename E::SolveOut o = E::solve(pd, L);
      status = o.status;
      cost = o.res.cost;
      if (iters_out) {
        iters_out[2 * b] = o.sqp_iter;
        iters_out[2 * b + 1] = o.ipm_iter;
      }
    }
    if (do_sens) {
      int ok = 1;
      const int ng = E::grad_width(pd);
      typename E::Residuals r = E::sens(pd, L, dL ? dL + (size_t)b * ng : nullptr,
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
      for (int i = 0; i < M::NU; ++i) u0_out[(size_t)b * M::NU + i] = L.it[(size_t)(E::it_u(pd.N, 0) + i) * bs];
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
}

extern "C" {

void cpu_port_set_threads(int n) { g_threads = n; }
int cpu_port_get_threads() { return g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency(); }

int cpu_port_sizeof_problem_data() { return (int)sizeof(ProblemData); }

int cpu_port_iterate_size(int model, int N) {
  if (model == 1) return Engine<CartpoleModel>::it_size(N);
  return -1;
}

int cpu_port_grad_width(int model, const ProblemData* pd) {
  if (model == 1) return Engine<CartpoleModel>::grad_width(*pd);
  return -1;
}

// iterate: [it_size][B] batch-minor, in/out (zero it + set x rows for a cold start)
int cpu_port_unit(int model, const ProblemData* pd, int mode, int max_sqp, int B, const double* theta, int per_sample,
                  const double* x0, const double* u0, double* iterate, int do_solve, int do_sens, double* u0_out,
                  double* cost_out, int* status_out, double* dL, double* dpi, double* res_out, int* iters_out) {
  if (model == 1) {
    run<CartpoleModel>(*pd, mode, max_sqp, B, theta, per_sample, x0, u0, iterate, do_solve, do_sens, u0_out, cost_out,
                       status_out, dL, dpi, res_out, iters_out);
    return 0;
  }
  return -1;
}
}
