/* rlmpc_b200.h -- C ABI of the B200 batched MPC engine (librlmpc_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path.  What the reference reaches
 * through ctypes into the acados-generated shared library (acados_template.AcadosOcpSolver:
 * set / get / solve / get_cost / constraints_set / cost_set / reset, called from
 * rlmpc/mpc/common/mpc.py:27-96,177-285) plus the CasADi/SuperLU work of update_nlp
 * (rlmpc/mpc/nlp.py:1341-1424) is exposed here as batched entry points with plain pointers
 * and sizes.  No torch / C++ types cross the boundary.
 *
 * Conventions
 *   - all functions return 0 on success and a negative RLMPC_E* code on failure; they never
 *     throw.  rlmpc_last_error() gives a message for the calling thread.
 *   - "_dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors);
 *     "_host" pointers are ordinary host memory.  Matrices are row-major [B, dim].
 *   - one handle = one problem on one device with room for max_batch samples.  The handle
 *     owns the persistent primal-dual iterate of every sample (the warm start, what acados
 *     keeps inside its solver object) and the per-stage scratch.  Calls on one handle must be
 *     serialised by the caller (stream order); the handle is not re-entrant, like an
 *     AcadosOcpSolver.
 *   - per-sample status uses acados' codes: 0 ok, 1 NaN, 2 max iter, 3 min step, 4 QP failure
 *     (what mpc.py:81-82,197-198 turns into RuntimeError).
 */
#ifndef RLMPC_B200_H
#define RLMPC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLMPC_MAXN 128
#define RLMPC_MAXD 8

/* models (device code emitted per problem, the role of acados' c_generated_code) */
#define RLMPC_MODEL_CARTPOLE 1      /* rlmpc/mpc/cartpole/acados.py:28-108 (config/cartpole*.yaml) */
#define RLMPC_MODEL_LINEAR_SYSTEM 2 /* rlmpc/mpc/linear_system/acados.py:27-131 */
#define RLMPC_MODEL_EVAPORATION 3   /* rlmpc/mpc/evaporation_process/acados.py:142-228 */
#define RLMPC_MODEL_CHAIN_MASS 4    /* rlmpc/mpc/chain_mass/ocp_utils.py:59-147,195-316 (n_mass = 3, 5, 6: nx = 9, 21, 27) */

#define RLMPC_MODE_V 0 /* x_0 fixed            -> V(s), pi(s)   mpc.py:177-202 (update, get_action) */
#define RLMPC_MODE_Q 1 /* x_0 and u_0 fixed    -> Q(s,a)        mpc.py:52-96   (q_update) */

#define RLMPC_EINVAL (-1)
#define RLMPC_ECUDA (-2)
#define RLMPC_ENOMEM (-3)
#define RLMPC_ENODEV (-4)

typedef struct rlmpc_handle rlmpc_handle;

/* What AcadosOcp carries for one problem (dims, cost scaling, bounds, integrator constants).
 * Replaces: AcadosOcp + AcadosOcpSolver(ocp, json_file=...) construction
 * (rlmpc/mpc/cartpole/acados.py:169-203). */
typedef struct rlmpc_problem_desc {
  int model;                        /* RLMPC_MODEL_* */
  int N;                            /* horizon, <= RLMPC_MAXN */
  double scale[RLMPC_MAXN + 1];     /* cost scaling per stage: dT, gamma^k dT, ... (nlp.py:1038-1134) */
  double lbu[RLMPC_MAXD], ubu[RLMPC_MAXD]; /* input bounds, all stages (constraints.lbu/ubu) */
  double lbx[RLMPC_MAXD], ubx[RLMPC_MAXD]; /* state bounds stages 1..N-1 (+-1e30 = none) */
  double lbx_e[RLMPC_MAXD], ubx_e[RLMPC_MAXD];
  double model_const[24];           /* cartpole: [0]=RK4 step h, [1]=g, [2] != 0: g is a learnable parameter, theta =
                                       [M, m, l, g | W_0 ...] with 84 entries (scripts/cartpole_mpc_qlearning.py:184-187); linear system: [0..2] = P11,P12,P22 of
                                       the constant terminal cost (linear_system/acados.py:51-57); evaporation:
                                       [0..18] = environment.PARAM in dict order, [19] = RK4 step, [20] = #steps;
                                       chain mass: [0] = RK4 step (Ts / 2: two steps per stage, ocp_utils.py:42-56),
                                       [1] = n_mass; the steady state x_ss of the cost goes through rlmpc_set_model_vector */
  double zl[RLMPC_MAXD], zu[RLMPC_MAXD]; /* linear penalties of the soft state bounds (cost.zl/zu), per idxsbx row */
  double lg[RLMPC_MAXD], ug[RLMPC_MAXD]; /* bounds of the affine general constraints (constraints.lh/uh) */
} rlmpc_problem_desc;

/* ---- lifetime ------------------------------------------------------------------------- */
/* Replaces AcadosOcpSolver.__init__ (code generation + gcc + dlopen). */
int rlmpc_create(const rlmpc_problem_desc* desc, int max_batch, int device, rlmpc_handle** out);
void rlmpc_destroy(rlmpc_handle* h);
const char* rlmpc_last_error(void);
/* dims: nx, nu, ntheta (length of the reference's p vector, nlp.py:970-989), ngrad (width of the
 * gradient rows, see rlmpc_sens), iterate doubles per sample */
int rlmpc_dims(const rlmpc_handle* h, int* nx, int* nu, int* ntheta, int* ngrad, int* iterate_size);
/* inequality rows per stage held in "lam"/"t": 2*(nu + nbx), acados order [lbu, lbx, ubu, ubx]
 * (rlmpc/common/utils.py:4-25); nbx = 0 unless the problem was created with finite state bounds */
int rlmpc_nrows(const rlmpc_handle* h);

/* ---- parameters / options --------------------------------------------------------------- */
/* Replaces ocp_solver.set(stage,"p",..) for all stages + cost_set(stage,"W"/"yref",..)
 * (mpc.py:137-149, 212-257).  theta is the reference's full p vector.  per_sample=0: one theta
 * shared by the batch; per_sample=1: theta_host is [B, ntheta] (parameter sweeps,
 * scripts/cartpole_mpc_sensitivities.py:79-98). */
int rlmpc_set_theta(rlmpc_handle* h, const double* theta_host, int per_sample, int B);
/* Stream-ordered variant: theta_dev is a device pointer ([ntheta], or [B, ntheta] with per_sample = 1), copied and
 * turned into the cost table by kernels on `stream`; no host synchronisation (the learning loop keeps theta on the
 * device).  The caller orders it against solves by using the same stream. */
int rlmpc_set_theta_dev(rlmpc_handle* h, const double* theta_dev, int per_sample, int B, void* stream);
/* Model data too large for model_const.  chain mass: name = "x_ss", the steady state the tracking cost refers to
 * (compute_parametric_steady_state, ocp_utils.py:150-192: a constant computed once when the OCP is built), n = nx. */
int rlmpc_set_model_vector(rlmpc_handle* h, const char* name, const double* v_host, int n);
/* Replaces ocp_nlp_cost_model_set(..., "scaling", ...) (mpc.py:259-285). n = N+1. */
int rlmpc_set_cost_scaling(rlmpc_handle* h, const double* scale, int n);
/* Replaces constraints_set(stage, "lbu"/"ubu"/..) for the nominal bounds. field in
 * {"lbu","ubu","lbx","ubx","lbx_e","ubx_e","zl","zu"}. */
int rlmpc_set_bounds(rlmpc_handle* h, const char* field, const double* v, int n);
/* options: "tol" (1e-6), "tau" (1e-8), "mu0" (1.0), "max_ipm" (50), "warm_ipm" (1: start every QP's
 * interior-point iteration from the multipliers of the previous QP, with a cold restart if jammed),
 * "sync_every" (4: SQP rounds between host checks "has every sample converged" when max_sqp > 1),
 * "timing" (0; 1 = record per-phase CUDA events, see rlmpc_get_timings), "overlap" (0),
 * "sigma_min" (0.05), "sigma0" (0.3): centring parameters of the interior-point iteration,
 * "coop" (1): queued QPs are solved by the warp-per-sample kernel, the whole stage QP of a sample resident in
 * shared memory (all shipped models; needs the horizon to fit: 43..99 doubles per stage); 0 = use the
 * thread-per-sample queue kernels,
 * "condense" (1): (thread-per-sample path) queued QPs of input-bounds-only problems are solved in partially condensed form
 * (blocks of 4 stages, like the reference's PARTIAL_CONDENSING_HPIPM) in V-mode when N % 4 == 0,
 * "ring" (1) / "ring_b" (0): cp.async shared-memory ring reader of the stage-form / condensed queue kernel,
 * "comp_accept" (0.2): an interior-point solve (and the single warm Newton iteration of the fast path) ends on
 * a full step whose rows all land within lam*t = tau (1 +- comp_accept) -- a neighbourhood of the tau-central
 * point, far inside HPIPM's own complementarity tolerance.  Measured against the oracle's exact tau-central RTI step
 * (tests/golden/cartpole_original_rti.npz, 256 states one environment step away from their iterate; worst sample, warm /
 * cold start of the interior point): 0.5 -> |du0| 1.5e-6 / 3.3e-5, 0.2 -> 2.3e-7 / 1.1e-6, 0.05 -> 4e-8 / 1.1e-6; the
 * headline step costs 5.49 / 5.6 / 5.73 ms at 0.5 / 0.2 / 0.05.  0.2 is the largest setting that keeps the north-star
 * |du0| < 1e-5 with a margin on both paths (tools/comp_accept_study.py),
 * "split" (2): an RTI solve(+sens) call on >= 4096 samples runs the two halves of the batch as two independent
 * chains of kernels on two streams (the caller's and an internal one, joined before the call returns to the
 * stream): kernels bound by different units overlap; 1 = one chain; rlmpc_get_timings then describes the first half,
 * "step_length" (1.0): fixed SQP step length on the primal variables, acados' nlp_solver_step_length.  Full steps
 * (the reference's setting) 2-cycle on ~8 % of random cart-pole swing-up states; 0.7 converges those at a linear rate,
 * "graph" (0): 1 = an RTI solve+sens call is captured into a CUDA graph on first use and replayed while its arguments
 * (mode, batch, every pointer, the options) stay the same; for small batches, where the chain of ~14 launches on two
 * streams is launch-bound (strong scaling over 8 GPUs: 8 192 samples per rank).  Not combined with "timing",
 * "as_steps" (20): active-set (full Newton step + projection) iterations a warm start may take when
 * its Newton step is infeasible, before it falls back to a cold start,
 * "fuse_lin" (0): 1 = the convergence-test / fast-path kernel linearises every stage itself instead of reading the
 * records of a separate (sample, stage) launch (measured equal at best, see profiles/r02_summary.md),
 * "ipm_passes" (0): n > 0 = n thread-per-sample pass kernels, each ONE interior-point iteration of every queued
 * sample in place, ahead of the queue kernel (measured slower on every workload, see profiles/r02_summary.md),
 * "param_cost" (0: dL/dtheta only for model parameters = parameterize_tracking_cost False) */
int rlmpc_set_option(rlmpc_handle* h, const char* name, double value);

/* ---- iterate (warm start) --------------------------------------------------------------- */
/* Replaces ocp_solver.reset() + set(stage,"x",x0) for all stages (mpc.py:204-210):
 * x_k = x0, u = 0, multipliers = 0 for samples [0,B). */
int rlmpc_reset(rlmpc_handle* h, int B, const double* x0_dev, void* stream);
/* Same for the samples with mask_dev[b] != 0 only (environments that were reset inside a vectorised
 * closed loop); mask_dev NULL = all. */
int rlmpc_reset_masked(rlmpc_handle* h, int B, const double* x0_dev, const int* mask_dev, void* stream);
/* Replaces ocp_solver.get/set(stage, field). field in {"x","u","pi","lam","t"}; buf_dev is
 * [B, dim(field)] row-major; lam/t are in acados order [lbu, ubu] for this problem class. */
int rlmpc_get_iterate(rlmpc_handle* h, const char* field, int stage, int B, double* buf_dev, void* stream);
int rlmpc_put_iterate(rlmpc_handle* h, const char* field, int stage, int B, const double* buf_dev, void* stream);

/* Warm-start store next to a replay buffer: the reference keeps ONE iterate inside its solver object
 * (mpc.py:204-210), so replayed samples always start cold; here every replay-buffer entry can keep its
 * own primal-dual iterate, which is what keeps an SQP-RTI step (max_sqp = 1) accurate.  The store is a
 * caller-owned device buffer of rlmpc_store_bytes(h, capacity) bytes; idx_dev[b] is the slot of batch
 * sample b (entries outside [0, capacity) are skipped).  to_store = 1: handle iterate -> store,
 * 0: store -> handle iterate. */
size_t rlmpc_store_bytes(const rlmpc_handle* h, int capacity);
int rlmpc_store_copy(rlmpc_handle* h, int B, const int* idx_dev, double* store_dev, int capacity, int to_store,
                     void* stream);

/* ---- the hot path ----------------------------------------------------------------------- */
/* Replaces ocp_solver.set(0,"lbx"/"ubx",x0) [+ constraints_set(0,"lbu"/"ubu",u0)] + solve()
 * + get(0,"u") + get_cost() for B samples (mpc.py:27-50, 52-82, 177-202).
 * max_sqp = 1 is one SQP-RTI step; larger = SQP to convergence (tol).
 * u0_dev may be NULL in V-mode.  Outputs may be NULL.  cost_out is the cost of the last
 * linearisation point (for converged solves: of the solution). */
int rlmpc_solve(rlmpc_handle* h, int mode, int max_sqp, int B, const double* x0_dev, const double* u0_dev,
                double* u0_out_dev, double* cost_out_dev, int* status_out_dev, void* stream);
/* Replaces update_nlp() (nlp.py:1341-1563) at the current iterate: cost, KKT residual norms
 * [stat, eq, ineq, comp], dL/dtheta [B, ngrad] (= dV/dtheta = dQ/dtheta) and
 * dpi/dtheta [B, nu, ngrad].  ngrad = number of model parameters when option "param_cost" is 0
 * (build_nlp(parameterize_tracking_cost=False): every other entry of p has zero gradient, quirk
 * Q8 -- the rows are the non-zero prefix of the reference's [1, ntheta] arrays), = ntheta when it
 * is 1 (rows must then be zero-initialised by the caller). */
int rlmpc_sens(rlmpc_handle* h, int mode, int B, double* dL_dtheta_dev, double* dpi_dtheta_dev,
               double* cost_out_dev, double* res_out_dev, int* status_out_dev, void* stream);
/* One fused launch: solve + sens (one "unit" of BASELINE.json's metric). */
int rlmpc_solve_sens(rlmpc_handle* h, int mode, int max_sqp, int B, const double* x0_dev, const double* u0_dev,
                     double* u0_out_dev, double* cost_out_dev, int* status_out_dev, double* dL_dtheta_dev,
                     double* dpi_dtheta_dev, double* res_out_dev, void* stream);
/* Host-buffer variant of rlmpc_solve_sens: copies x0/u0 in, runs, copies all outputs back and
 * synchronises.  This is the end-to-end call the reference-side binding would make.  It runs on streams of the
 * handle and first waits for everything queued on the device (so it is ordered after earlier stream-ordered
 * calls on the handle).  With page-locked buffers an RTI call is pipelined per part of the batch (option
 * "split"): copy in, kernel chain and copy out of one part overlap with those of the other. */
int rlmpc_solve_sens_host(rlmpc_handle* h, int mode, int max_sqp, int B, const double* x0_host,
                          const double* u0_host, double* u0_out_host, double* cost_out_host, int* status_out_host,
                          double* dL_dtheta_host, double* dpi_dtheta_host, double* res_out_host);

/* ---- TD / policy-gradient accumulator ------------------------------------------------------ */
/* Replaces  dp = mean_i(LR * td_i * dQ_dp_i)  (examples/linear_system_mpc_qlearning.py:193,203):
 * dQ_dtheta_dev is [B, ncols]; acc_out_dev[0..ncols) = sum_i td_i * dQ_dtheta[i,:], acc[ncols] =
 * sum_i td_i, acc[ncols+1] = number of valid samples (status 0; status_dev may be NULL).  The
 * caller all-reduces acc over ranks (NCCL) and divides. */
int rlmpc_td_grad(rlmpc_handle* h, int B, int ncols, const double* td_dev, const double* dQ_dtheta_dev,
                  const int* status_dev, double* acc_out_dev, void* stream);

/* ---- vectorised environment (closed-loop training without host round trips) ------------------- */
/* Replaces ContinuousCartPoleSwingUpVectorEnv.step (rlmpc/gym/continuous_cartpole/environment.py:372-426)
 * for B environments on the current device.  par_dev[13] = [gravity, masscart, masspole, length,
 * force_mag, tau, x_threshold, theta_threshold, max_episode_steps, reset_state(4)]; state_dev [B,4]
 * in/out, action_dev [B] in [-1,1], steps_dev [B] in/out; finished environments are reset in place. */
int rlmpc_cartpole_env_step(const double* par_dev, int B, double* state_dev, const double* action_dev,
                            double* reward_dev, int* terminated_dev, int* truncated_dev, int* steps_dev, void* stream);

/* FP64 FMA throughput of the device, measured with a register-resident probe kernel (8 independent chains per thread,
 * 2048 threads per SM): the compute roofline of this FP64 path.  tflops_out: best of 5 timed launches, counting
 * FMA = 2 flop.  ("measure first"; MEASURED_PEAKS.json has no FP64 figure.) */
int rlmpc_fp64_peak(int device, double* tflops_out);
/* Same for the FP64 tensor-core path (mma.sync m8n8k4, DMMA), 512 flop per instruction and warp. */
int rlmpc_fp64_tensor_peak(int device, double* tflops_out);

/* number of kernels launched through this handle so far (bench.py's gpu_launches) */
long long rlmpc_launch_count(const rlmpc_handle* h);
/* With option "timing" = 1 the library records CUDA events on the caller's stream between the phases
 * of a solve / sens call.  ms_out[0..6) = device time of the last call's
 * [linearise | convergence test + fast QP | full interior point of the queued samples | sens stage
 * evaluation | sens sweeps | tail = sens kernels of the queued samples] (of the last SQP round when
 * max_sqp > 1); waits for the call to finish.  n >= 6; with n >= 8 also ms_out[6] = number of queued samples and
 * ms_out[7] = their interior-point iterations (warp-per-sample queue kernel) in that round.  With option "overlap" = 1 (default 0) an RTI
 * solve_sens call runs the third phase on a side stream concurrently with the fourth and fifth.  The reference's
 * analogue is the nlp_timing dict of update_nlp (nlp.py:1397-1422). */
int rlmpc_get_timings(rlmpc_handle* h, double* ms_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* RLMPC_B200_H */
