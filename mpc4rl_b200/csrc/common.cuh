// Shared definitions of the rlmpc-b200 engine (host/device generic maths).
// The same templates are instantiated in CUDA kernels (rlmpc_b200.cu, the product) and, for
// timing/debugging only, in a host build under oracle/cpu_port (test infrastructure).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define MPC_HD __host__ __device__ __forceinline__
#define MPC_UNROLL _Pragma("unroll")
#else
#define MPC_HD inline __attribute__((always_inline))
#define MPC_UNROLL
#endif

namespace rlmpc {

constexpr int MAXN = 128;  // max horizon supported by ProblemData
constexpr int MAXD = 8;    // max nx / nu held in ProblemData bound arrays (thread-per-sample engine)
constexpr double BIG = 1e29;  // |bound| >= BIG means "no bound"
constexpr double AS_RELEASE = 0.5;  // active-set step: slack given to a released row, as a fraction of the row's range
// Samples are stored in tiles of TILE (one warp): element i of the sample in lane l of tile T lives
// at base[(T * size + i) * TILE + l].  A warp access to element i is one contiguous 256-byte run
// (fully coalesced) and, because TILE is a compile-time constant, every element offset inside a
// thread is an immediate of the load/store instruction -- no per-access index arithmetic.
constexpr int TILE = 32;

// acados status codes: 0 ok, 1 NaN, 2 max iter, 3 min step, 4 QP failure
enum Status : int { ST_OK = 0, ST_NAN = 1, ST_MAXITER = 2, ST_MINSTEP = 3, ST_QPFAIL = 4 };

enum Mode : int { MODE_V = 0, MODE_Q = 1 };

// per-sample pipeline state between the kernels of one solve call
enum Work : int {
  WK_ACTIVE = 0,   // needs (another) SQP iteration
  WK_HARD = 1,     // linearised, fast QP path declined: queued for the full interior-point solve
  WK_DONE = 2,     // finished (status holds the result)
};

// Everything that is shared by all samples of a batch. Passed to kernels by value
// (fits the 4 KB kernel-parameter space), so it sits in constant memory.
struct ProblemData {
  int N;
  int mode;             // MODE_V / MODE_Q
  int max_sqp;          // max SQP iterations (QP solves); 1 == RTI
  int max_ipm;          // max interior-point iterations per QP
  int warm_ipm;         // 1: start the IPM from the stored (lam,t)
  int param_cost;       // 1: gradient wrt cost parameters requested (parameterize_tracking_cost)
  int fix0;             // leading inputs of stage 0 that carry no rows (the clamped u_0 of Q-mode in block form)
  double tol;           // SQP convergence tolerance on the 4 KKT residual norms
  double tau;           // complementarity target lam*t = tau (nlp.py:1199)
  double mu0;           // initial barrier parameter of a cold-started IPM
  double sigma_min;     // smallest centring parameter (barrier reduction per full step)
  double sigma0;        // centring parameter of the first iteration of a cold start
  double as_steps;      // warm start: number of active-set (full step + projection) iterations tried
                        // when the Newton step is not feasible, before the cold restart
  double condense;      // > 0: queued QPs are solved in partially condensed form (condense.cuh) where applicable
  double comp_accept;   // a Newton step / interior-point solve is accepted when it is a full step and every row ends
                        // within lam*t = tau (1 +- comp_accept), i.e. max |dlam dt| <= comp_accept * tau
  double step_length;   // fixed SQP step length on the primal variables (acados: nlp_solver_step_length, default 1):
                        // w += step_length * dw; the multipliers are the QP's.  < 1 cures the 2-cycles of full-step
                        // Gauss-Newton SQP at the price of a linear rate
  double scale[MAXN + 1];  // per-stage cost scaling s_k (dT, gamma^k dT, ...)
  double lbu[MAXD], ubu[MAXD];
  double lbx[MAXD], ubx[MAXD];      // stages 1..N-1, indexed by state component
  double lbx_e[MAXD], ubx_e[MAXD];  // stage N
  double zl[MAXD], zu[MAXD];        // linear penalties of the soft state bounds (cost.zl / cost.zu), per soft row
  double lg[MAXD], ug[MAXD];        // bounds of the general linear rows (constraints.lh / uh)
  double mc[24];        // model constants (integrator step, gravity, plant parameters, ...)
};

// One sample's strided view: element i of a logical per-sample vector lives at p[i*TILE].
struct Lane {
  double* it;        // iterate (persistent primal-dual state, warm start)
  double* ws;        // per-stage scratch
  const double* th;  // parameter vector theta (a shared theta is stored as one replicated tile)
  const double* ct;  // quadratic cost table derived from theta (Engine::CT_*)
};
// offset of element i of sample b in a tiled array whose per-sample size is n
MPC_HD size_t tile_off(size_t b, size_t n) { return (b / TILE) * n * TILE + (b % TILE); }

MPC_HD double dmax(double a, double b) { return a > b ? a : b; }
MPC_HD double dmin(double a, double b) { return a < b ? a : b; }
MPC_HD double dabs(double a) { return a < 0 ? -a : a; }

}  // namespace rlmpc
