// Chain-mass kernels of librlmpc_b200.so (sm_100a): launch wrappers around chain/chain_engine.cuh and the device
// buffers of a chain-mass handle.  The C ABI lives in rlmpc_b200.cu and dispatches here (chain/chain_backend.h).
//
// Kernels per RTI unit (solve + sensitivities), all on the caller's stream:
//   k_chain_stage<false>  warp per (sample, stage)   linearisation                      FP64 issue
//   k_chain_qp            warp per sample            convergence test, Riccati IPM, step FP64 issue + TMA-fed HBM stream
//   k_chain_stage<true>   warp per (sample, stage)   exact Hessian of pi'F, pi'dF/dtheta FP64 issue
//   k_chain_sens          warp per sample            factorisation, adjoint solves, Q/R columns
//   k_chain_param         thread per (sample, stage, rhs)  dynamic-parameter columns of dpi/dtheta
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/rlmpc_b200.h"
#include "chain/chain_backend.h"
#include "chain/chain_engine.cuh"

namespace rlmpc {

namespace {

#define CK(expr)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (expr);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      err = std::string(#expr) + ": " + cudaGetErrorString(e_);                      \
      return RLMPC_ECUDA;                                                            \
    }                                                                                \
  } while (0)

#ifndef CHAIN_STAGE_WARPS
#define CHAIN_STAGE_WARPS 4
#endif
#ifndef CHAIN_QP_WARPS
#define CHAIN_QP_WARPS 2
#endif
#ifndef CHAIN_SENS_WARPS
#define CHAIN_SENS_WARPS 1
#endif
constexpr int STAGE_WARPS = CHAIN_STAGE_WARPS;  // warps per block of the (sample, stage) kernel
constexpr int QP_WARPS = CHAIN_QP_WARPS;        // samples in flight per block of the per-sample kernels
constexpr int SENS_WARPS = CHAIN_SENS_WARPS;

template <int NM>
__global__ void k_chain_begin(const __grid_constant__ ProblemData pd, const ChainArgs a) {
  using E = ChainEngine<NM>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  double* it = a.it + (size_t)b * E::it_size(pd.N);
  for (int i = 0; i < E::NX; ++i) it[E::it_x(pd.N, 0) + i] = a.x0[(size_t)b * E::NX + i];
  if (pd.mode == MODE_Q)
    for (int i = 0; i < E::NU; ++i) it[E::it_u(pd.N, 0) + i] = a.u0[(size_t)b * E::NU + i];
  a.work[b] = WK_ACTIVE;
  a.status[b] = ST_MAXITER;
}

template <int NM, bool HESS>
__global__ void __launch_bounds__(STAGE_WARPS * 32) k_chain_stage(const __grid_constant__ ProblemData pd, const ChainArgs a) {
  using E = ChainEngine<NM>;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long task = (long long)blockIdx.x * STAGE_WARPS + wib;
  const int b = (int)(task / (pd.N + 1)), k = (int)(task - (long long)b * (pd.N + 1));
  if (b >= a.B) return;
  if (!HESS && a.work[b] != WK_ACTIVE) return;
  E::template stage_task<HESS>(pd, a, b, k, smem + (size_t)wib * E::SM_STAGE, lane);
}

template <int NM>
__global__ void __launch_bounds__(QP_WARPS * 32) k_chain_qp(const __grid_constant__ ProblemData pd, const ChainArgs a) {
  using E = ChainEngine<NM>;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* S = smem + (size_t)wib * ev2(E::qp_smem_doubles(pd.N));
  StageFeed feed;
  E::qp_feed_init(feed, S, pd.N, lane);
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(&a.counters[0], 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= a.B) break;
    if (a.work[b] != WK_ACTIVE) continue;
    int iters = 0;
    const int res = E::qp_sample(pd, a, b, S, feed, lane, &iters);
    if (lane == 0) {
      if (iters) atomicAdd(&a.counters[2], iters);
      if (res == E::R_NAN) { a.status[b] = ST_NAN; a.work[b] = WK_DONE; }
      else if (res == E::R_CONVERGED) { a.status[b] = ST_OK; a.work[b] = WK_DONE; }
      else if (res == E::R_TESTONLY) { a.work[b] = WK_DONE; }
      else if (res == E::R_FAILED) { a.status[b] = ST_QPFAIL; a.work[b] = WK_DONE; }
      else if (pd.max_sqp == 1) { a.status[b] = (res == E::R_STEPPED) ? ST_OK : ST_QPFAIL; a.work[b] = WK_DONE; }
    }
    __syncwarp();
  }
}

template <int NM>
__global__ void __launch_bounds__(SENS_WARPS * 32) k_chain_sens(const __grid_constant__ ProblemData pd, const ChainArgs a) {
  using E = ChainEngine<NM>;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* S = smem + (size_t)wib * ev2(E::sens_smem_doubles(pd.N));
  StageFeed feed;
  E::sens_feed_init(feed, S, pd.N, lane);
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(&a.counters[0], 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= a.B) break;
    E::sens_sample(pd, a, b, S, feed, lane);
    __syncwarp();
  }
}

// block per sample; thread t < 3 N owns (stage t / 3, right-hand side t % 3)
template <int NM>
__global__ void k_chain_param(const __grid_constant__ ProblemData pd, const ChainArgs a) {
  using E = ChainEngine<NM>;
  using Mo = ChainModel<NM>;
  extern __shared__ double smem[];
  const int b = blockIdx.x, t = threadIdx.x, nt = E::NU * pd.N;
  if (t < nt) E::param_task(pd, a, b, t / E::NU, t % E::NU, smem + (size_t)t * E::NPD);
  __syncthreads();
  for (int e = t; e < E::NU * E::NPD; e += blockDim.x) {
    const int r = e / E::NPD, p = e - r * E::NPD;
    double acc = 0.0;
    for (int k = 0; k < pd.N; ++k) acc += smem[(size_t)(k * E::NU + r) * E::NPD + p];
    a.dpi[((size_t)b * E::NU + r) * E::NTH + Mo::pd_to_theta(p)] = acc;
  }
}

template <int NM>
__global__ void k_chain_out(const __grid_constant__ ProblemData pd, const ChainArgs a) {
  using E = ChainEngine<NM>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  if (a.u0_out)
    for (int i = 0; i < E::NU; ++i) a.u0_out[(size_t)b * E::NU + i] = a.it[(size_t)b * E::it_size(pd.N) + E::it_u(pd.N, 0) + i];
  if (a.cost_out) a.cost_out[b] = a.cost[b];
  if (a.status_out) a.status_out[b] = a.status[b];
}

__global__ void k_chain_count_active(const int* work, int B, int* counters) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = b < B && work[b] != WK_DONE;
  const unsigned m = __ballot_sync(0xffffffffu, act);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[1], __popc(m));
}

// theta -> sym(Q), sym(R); x_ss appended
template <int NM>
__global__ void k_chain_tables(const double* th, const double* xss, double* tab) {
  using E = ChainEngine<NM>;
  using Mo = ChainModel<NM>;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E::TB_SIZE; e += gridDim.x * blockDim.x) {
    double v;
    if (e < E::TB_R) {
      const int i = e / E::NX, j = e - i * E::NX;
      v = 0.5 * (th[Mo::TH_Q + i + j * E::NX] + th[Mo::TH_Q + j + i * E::NX]);
    } else if (e < E::TB_XSS) {
      const int i = (e - E::TB_R) / E::NU, j = (e - E::TB_R) - i * E::NU;
      v = 0.5 * (th[Mo::TH_R + i + j * E::NU] + th[Mo::TH_R + j + i * E::NU]);
    } else {
      v = xss[e - E::TB_XSS];
    }
    tab[e] = v;
  }
}

// MPC.reset: x_k = x0 for all stages, everything else zero
__global__ void k_chain_reset(double* it, int it_size, int nxtot, int nx, int B, const double* x0, const int* mask) {
  const int b = blockIdx.x;
  if (b >= B || (mask && !mask[b])) return;
  double* p = it + (size_t)b * it_size;
  for (int i = threadIdx.x; i < it_size; i += blockDim.x) p[i] = (x0 && i < nxtot) ? x0[(size_t)b * nx + (i % nx)] : 0.0;
}

__global__ void k_chain_copy_field(double* it, int it_size, int B, int off, int dim, double* buf, int to_iterate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * dim) return;
  const int b = e / dim, i = e - b * dim;
  double* p = it + (size_t)b * it_size + off + i;
  if (to_iterate) *p = buf[e];
  else buf[e] = *p;
}

__global__ void k_chain_store_copy(double* it, double* store, int it_size, int B, const int* idx, int capacity, int to_store) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const int slot = idx[b];
  if (slot < 0 || slot >= capacity) return;
  double* pi_ = it + (size_t)b * it_size;
  double* ps = store + (size_t)slot * it_size;
  for (int i = threadIdx.x; i < it_size; i += blockDim.x) {
    if (to_store) ps[i] = pi_[i];
    else pi_[i] = ps[i];
  }
}

}  // namespace

struct ChainBackend {
  int n_mass, N, max_batch;
  int nx, nu, nth, it_size, rec, tb_size;
  double *it = nullptr, *ws = nullptr, *th = nullptr, *tab = nullptr, *xss = nullptr, *cost = nullptr;
  int *status = nullptr, *work = nullptr, *counters = nullptr, *h_counters = nullptr;
  int qp_grid = 0, sens_grid = 0;
  size_t stage_smem = 0, qp_smem = 0, sens_smem = 0, param_smem = 0;
  long long launches = 0;
  int timing = 0;
  cudaEvent_t ev[6] = {};
  bool ev_set[6] = {};
};

namespace {

#define CHAIN_DISPATCH(cb, ...)                        \
  switch ((cb)->n_mass) {                              \
    case 3: { constexpr int NM = 3; __VA_ARGS__; } break; \
    case 5: { constexpr int NM = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int NM = 6; __VA_ARGS__; } break; \
  }

template <int NM>
int setup(ChainBackend* cb, std::string& err) {
  using E = ChainEngine<NM>;
  const int N = cb->N;
  cb->nx = E::NX; cb->nu = E::NU; cb->nth = E::NTH; cb->it_size = E::it_size(N); cb->rec = E::REC; cb->tb_size = E::TB_SIZE;
  cb->stage_smem = sizeof(double) * STAGE_WARPS * E::SM_STAGE;
  cb->qp_smem = sizeof(double) * QP_WARPS * ev2(E::qp_smem_doubles(N));
  cb->sens_smem = sizeof(double) * SENS_WARPS * ev2(E::sens_smem_doubles(N));
  cb->param_smem = sizeof(double) * (size_t)E::NU * N * E::NPD;
  int dev = 0, sms = 0, max_optin = 0, per_sm = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (cb->stage_smem > (size_t)max_optin || cb->qp_smem > (size_t)max_optin || cb->sens_smem > (size_t)max_optin ||
      cb->param_smem > (size_t)max_optin) {
    err = "horizon too long for the shared memory of the chain-mass kernels";
    return RLMPC_EINVAL;
  }
  CK(cudaFuncSetAttribute(k_chain_stage<NM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
  CK(cudaFuncSetAttribute(k_chain_stage<NM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
  CK(cudaFuncSetAttribute(k_chain_qp<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
  CK(cudaFuncSetAttribute(k_chain_sens<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
  CK(cudaFuncSetAttribute(k_chain_param<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_qp<NM>, QP_WARPS * 32, cb->qp_smem));
  cb->qp_grid = sms * (per_sm > 0 ? per_sm : 1);
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_sens<NM>, SENS_WARPS * 32, cb->sens_smem));
  cb->sens_grid = sms * (per_sm > 0 ? per_sm : 1);
  return 0;
}

void mark(ChainBackend* cb, int i, cudaStream_t s) {
  if (!cb->timing) return;
  cudaEventRecord(cb->ev[i], s);
  cb->ev_set[i] = true;
}

template <int NM>
int run_t(ChainBackend* cb, const ProblemData& pd0, const ChainCall& c, int sync_every, cudaStream_t s, std::string& err) {
  using E = ChainEngine<NM>;
  ProblemData pd = pd0;
  pd.mode = c.mode;
  pd.max_sqp = c.max_sqp;
  const int B = c.B, N = pd.N;
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.it = cb->it; a.ws = cb->ws; a.th = cb->th; a.tab = cb->tab; a.B = B;
  a.status = cb->status; a.work = cb->work; a.cost = cb->cost; a.counters = cb->counters;
  a.x0 = c.x0; a.u0 = (c.mode == MODE_Q) ? c.u0 : nullptr;
  a.u0_out = c.u0_out; a.cost_out = c.cost_out; a.status_out = c.status_out; a.dL = c.dL; a.dpi = c.dpi; a.res_out = c.res_out;
  const long long tasks = (long long)B * (N + 1);
  const int gstage = (int)((tasks + STAGE_WARPS - 1) / STAGE_WARPS);
  for (int i = 0; i < 6; ++i) cb->ev_set[i] = false;
  if (c.do_solve) {
    k_chain_begin<NM><<<(B + 127) / 128, 128, 0, s>>>(pd, a);
    cb->launches++;
    const int K = c.max_sqp, rounds = (K == 1) ? 1 : K + 1;
    for (int r = 0; r < rounds; ++r) {
      a.last_round = (K > 1 && r == K) ? 1 : 0;
      CK(cudaMemsetAsync(cb->counters, 0, 4 * sizeof(int), s));
      mark(cb, 0, s);
      k_chain_stage<NM, false><<<gstage, STAGE_WARPS * 32, cb->stage_smem, s>>>(pd, a);
      mark(cb, 1, s);
      k_chain_qp<NM><<<cb->qp_grid, QP_WARPS * 32, cb->qp_smem, s>>>(pd, a);
      mark(cb, 2, s);
      cb->launches += 2;
      if (K > 1 && !a.last_round && (r % sync_every) == sync_every - 1) {
        k_chain_count_active<<<(B + 127) / 128, 128, 0, s>>>(cb->work, B, cb->counters);
        cb->launches++;
        CK(cudaMemcpyAsync(cb->h_counters, cb->counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (cb->h_counters[1] == 0) break;
      }
    }
  }
  a.have_solve = c.do_solve;
  a.last_round = 0;
  if (c.do_sens) {
    CK(cudaMemsetAsync(cb->counters, 0, sizeof(int), s));
    mark(cb, 2, s);
    k_chain_stage<NM, true><<<gstage, STAGE_WARPS * 32, cb->stage_smem, s>>>(pd, a);
    mark(cb, 3, s);
    k_chain_sens<NM><<<cb->sens_grid, SENS_WARPS * 32, cb->sens_smem, s>>>(pd, a);
    mark(cb, 4, s);
    cb->launches += 2;
    if (c.dpi && c.mode == MODE_V) {
      const int nt = (E::NU * N + 31) / 32 * 32;
      k_chain_param<NM><<<B, nt, cb->param_smem, s>>>(pd, a);
      cb->launches++;
    }
    mark(cb, 5, s);
  } else {
    k_chain_out<NM><<<(B + 127) / 128, 128, 0, s>>>(pd, a);
    cb->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

int chain_create(int n_mass, int N, int max_batch, ChainBackend** out, std::string& err) {
  if (n_mass != 3 && n_mass != 5 && n_mass != 6) {
    err = "chain mass: n_mass must be 3, 5 or 6 (device code is emitted per size)";
    return RLMPC_EINVAL;
  }
  ChainBackend* cb = new (std::nothrow) ChainBackend();
  if (!cb) { err = "out of host memory"; return RLMPC_ENOMEM; }
  cb->n_mass = n_mass; cb->N = N; cb->max_batch = max_batch;
  int rc = RLMPC_EINVAL;
  CHAIN_DISPATCH(cb, rc = setup<NM>(cb, err));
  if (rc) { delete cb; return rc; }
  const size_t nB = (size_t)max_batch;
  cudaError_t e = cudaMalloc(&cb->it, sizeof(double) * cb->it_size * nB);
  if (e == cudaSuccess) e = cudaMalloc(&cb->ws, sizeof(double) * (size_t)(N + 1) * cb->rec * nB);
  if (e == cudaSuccess) e = cudaMalloc(&cb->th, sizeof(double) * cb->nth);
  if (e == cudaSuccess) e = cudaMalloc(&cb->tab, sizeof(double) * cb->tb_size);
  if (e == cudaSuccess) e = cudaMalloc(&cb->xss, sizeof(double) * cb->nx);
  if (e == cudaSuccess) e = cudaMalloc(&cb->cost, sizeof(double) * nB);
  if (e == cudaSuccess) e = cudaMalloc(&cb->status, sizeof(int) * nB);
  if (e == cudaSuccess) e = cudaMalloc(&cb->work, sizeof(int) * nB);
  if (e == cudaSuccess) e = cudaMalloc(&cb->counters, sizeof(int) * 8);
  if (e == cudaSuccess) e = cudaMallocHost(&cb->h_counters, sizeof(int) * 8);
  if (e == cudaSuccess) e = cudaMemset(cb->it, 0, sizeof(double) * cb->it_size * nB);
  if (e == cudaSuccess) e = cudaMemset(cb->ws, 0, sizeof(double) * (size_t)(N + 1) * cb->rec * nB);
  if (e == cudaSuccess) e = cudaMemset(cb->th, 0, sizeof(double) * cb->nth);
  if (e == cudaSuccess) e = cudaMemset(cb->tab, 0, sizeof(double) * cb->tb_size);
  if (e == cudaSuccess) e = cudaMemset(cb->xss, 0, sizeof(double) * cb->nx);
  if (e == cudaSuccess) e = cudaMemset(cb->status, 0, sizeof(int) * nB);
  if (e == cudaSuccess) e = cudaMemset(cb->cost, 0, sizeof(double) * nB);
  if (e == cudaSuccess) e = cudaMemset(cb->counters, 0, sizeof(int) * 8);
  for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&cb->ev[i]);
  if (e != cudaSuccess) {
    err = std::string("allocation failed: ") + cudaGetErrorString(e);
    chain_destroy(cb);
    return e == cudaErrorMemoryAllocation ? RLMPC_ENOMEM : RLMPC_ECUDA;
  }
  *out = cb;
  return 0;
}

void chain_destroy(ChainBackend* cb) {
  if (!cb) return;
  cudaFree(cb->it); cudaFree(cb->ws); cudaFree(cb->th); cudaFree(cb->tab); cudaFree(cb->xss); cudaFree(cb->cost);
  cudaFree(cb->status); cudaFree(cb->work); cudaFree(cb->counters);
  cudaFreeHost(cb->h_counters);
  for (int i = 0; i < 6; ++i)
    if (cb->ev[i]) cudaEventDestroy(cb->ev[i]);
  delete cb;
}

void chain_dims(const ChainBackend* cb, int* nx, int* nu, int* ntheta, int* it_size) {
  if (nx) *nx = cb->nx;
  if (nu) *nu = cb->nu;
  if (ntheta) *ntheta = cb->nth;
  if (it_size) *it_size = cb->it_size;
}

static int refresh_tables(ChainBackend* cb, cudaStream_t s, std::string& err) {
  CHAIN_DISPATCH(cb, (k_chain_tables<NM><<<4, 256, 0, s>>>(cb->th, cb->xss, cb->tab)));
  cb->launches++;
  CK(cudaGetLastError());
  return 0;
}

int chain_set_theta(ChainBackend* cb, const ProblemData&, const double* theta, bool on_device, cudaStream_t s, std::string& err) {
  CK(cudaMemcpyAsync(cb->th, theta, sizeof(double) * cb->nth, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (!on_device) CK(cudaStreamSynchronize(s));  // the host buffer may be reused by the caller
  return refresh_tables(cb, s, err);
}

int chain_set_xss(ChainBackend* cb, const ProblemData&, const double* xss_host, int n, std::string& err) {
  if (n != cb->nx) { err = "x_ss must have nx entries"; return RLMPC_EINVAL; }
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(cb->xss, xss_host, sizeof(double) * n, cudaMemcpyHostToDevice));
  return refresh_tables(cb, nullptr, err);
}

int chain_reset(ChainBackend* cb, const ProblemData& pd, int B, const double* x0_dev, const int* mask_dev, cudaStream_t s, std::string& err) {
  k_chain_reset<<<B, 128, 0, s>>>(cb->it, cb->it_size, (pd.N + 1) * cb->nx, cb->nx, B, x0_dev, mask_dev);
  cb->launches++;
  CK(cudaGetLastError());
  return 0;
}

template <int NM>
static int field_off(const ProblemData& pd, const char* field, int stage, int* off, int* dim, std::string& err) {
  using E = ChainEngine<NM>;
  const int N = pd.N;
  const bool le = stage >= 0 && stage <= N, lt = stage >= 0 && stage < N;
  if (!strcmp(field, "x")) { if (!le) goto bad; *off = E::it_x(N, stage); *dim = E::NX; }
  else if (!strcmp(field, "u")) { if (!lt) goto bad; *off = E::it_u(N, stage); *dim = E::NU; }
  else if (!strcmp(field, "pi")) { if (!lt) goto bad; *off = E::it_pi(N, stage); *dim = E::NX; }
  else if (!strcmp(field, "lam")) { if (!lt) goto bad; *off = E::it_lam(N, stage); *dim = E::NR; }
  else if (!strcmp(field, "t")) { if (!lt) goto bad; *off = E::it_t(N, stage); *dim = E::NR; }
  else if (!strcmp(field, "rho_x0")) { *off = E::it_rx0(N); *dim = E::NX; }
  else if (!strcmp(field, "rho_u0")) { *off = E::it_ru0(N); *dim = E::NU; }
  else if (!strcmp(field, "meta")) { *off = E::it_meta(N); *dim = 1; }
  else { err = std::string("unknown field ") + field; return RLMPC_EINVAL; }
  return 0;
bad:
  err = "stage out of range";
  return RLMPC_EINVAL;
}

int chain_field(ChainBackend* cb, const ProblemData& pd, const char* field, int stage, int B, double* buf_dev, int to_iterate,
                cudaStream_t s, int* dim_out, std::string& err) {
  int off = 0, dim = 0, rc = RLMPC_EINVAL;
  CHAIN_DISPATCH(cb, rc = field_off<NM>(pd, field, stage, &off, &dim, err));
  if (rc) return rc;
  if (dim_out) *dim_out = dim;
  if (B == 0 || !buf_dev) return 0;
  k_chain_copy_field<<<(B * dim + 127) / 128, 128, 0, s>>>(cb->it, cb->it_size, B, off, dim, buf_dev, to_iterate);
  cb->launches++;
  CK(cudaGetLastError());
  return 0;
}

int chain_run(ChainBackend* cb, const ProblemData& pd, const ChainCall& c, int sync_every, cudaStream_t s, std::string& err) {
  int rc = RLMPC_EINVAL;
  CHAIN_DISPATCH(cb, rc = run_t<NM>(cb, pd, c, sync_every < 1 ? 1 : sync_every, s, err));
  return rc;
}

size_t chain_store_bytes(const ChainBackend* cb, int capacity) { return sizeof(double) * (size_t)cb->it_size * (size_t)capacity; }

int chain_store_copy(ChainBackend* cb, int B, const int* idx_dev, double* store_dev, int capacity, int to_store, cudaStream_t s,
                     std::string& err) {
  k_chain_store_copy<<<B, 128, 0, s>>>(cb->it, store_dev, cb->it_size, B, idx_dev, capacity, to_store);
  cb->launches++;
  CK(cudaGetLastError());
  return 0;
}

long long chain_launches(const ChainBackend* cb) { return cb->launches; }
void chain_set_timing(ChainBackend* cb, int on) { cb->timing = on; }

int chain_timings(ChainBackend* cb, double* ms_out, int n, std::string& err) {
  // [linearise | qp | (unused) | sens stage | sens sweep | param contraction], like rlmpc_get_timings' six phases
  const int from[6] = {0, 1, 2, 2, 3, 4}, to[6] = {1, 2, 2, 3, 4, 5};
  for (int i = 0; i < 6 && i < n; ++i) {
    ms_out[i] = 0.0;
    if (from[i] != to[i] && cb->ev_set[from[i]] && cb->ev_set[to[i]]) {
      CK(cudaEventSynchronize(cb->ev[to[i]]));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, cb->ev[from[i]], cb->ev[to[i]]));
      ms_out[i] = ms;
    }
  }
  if (n >= 8) {
    CK(cudaDeviceSynchronize());
    int c[4] = {};
    CK(cudaMemcpy(c, cb->counters, sizeof(c), cudaMemcpyDeviceToHost));
    ms_out[6] = 0.0;
    ms_out[7] = c[2];
  }
  return 0;
}

}  // namespace rlmpc
