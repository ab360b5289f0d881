// librlmpc_b200.so -- CUDA kernels (sm_100a) + the C ABI declared in include/rlmpc_b200.h.
//
// Mapping (nx <= 4 problems): function/derivative evaluation runs one CUDA thread per (sample,
// stage) pair, the Riccati recursions one thread per sample.  The primal-dual iterate and the
// per-stage records live in HBM in tiles of 32 samples (AoSoA), so every load/store a warp issues
// is a fully coalesced 256-byte access at an immediate offset.  See DESIGN.md.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/rlmpc_b200.h"
#include "engine.cuh"
#include "condense.cuh"
#include "coop_general.cuh"
#include "models/cartpole.cuh"
#include "models/linear_system.cuh"
#include "models/evaporation.cuh"
#include "chain/chain_backend.h"

using namespace rlmpc;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(RLMPC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// launch shapes (tuned with tools/bench_variants.sh; -D overrides build the variants)
#ifndef RLMPC_TPB
#define RLMPC_TPB 64  // threads per block of the sample-parallel kernels
#endif
#ifndef RLMPC_QP1_MINB
#define RLMPC_QP1_MINB 8  // 128 registers: 0.58 -> 0.50 ms per 65 536 samples (profiles/r01_summary.md)
#endif
#ifndef RLMPC_SW_MINB
#define RLMPC_SW_MINB 1
#endif
#ifndef RLMPC_STAGE_TPB
#define RLMPC_STAGE_TPB 128  // threads per block of the (sample, stage) kernels
#endif
#ifndef RLMPC_LIN_MINB
#define RLMPC_LIN_MINB 4  // cart-pole: 136 -> 128 registers, 0.201 -> 0.169 ms per 65 536 x 41 (profiles/r02_summary.md)
#endif
#ifndef RLMPC_SS_MINB
#define RLMPC_SS_MINB 1
#endif
constexpr int TPB = RLMPC_TPB;
constexpr int STPB = RLMPC_STAGE_TPB;

// Per-call state shared by the kernels of one pipeline (device pointers into the handle).
struct KArgs {
  double* it;
  double* ws;
  int it_size, ws_size, th_size, ct_size;  // doubles per sample of each tiled array
  double* it2;  // compact copies used by the full interior-point pass (one slot per queued sample)
  double* ws2;
  double* itb;  // block-form iterate / workspace of the partially condensed queue path
  double* wsb;
  int itb_size, wsb_size;
  const double* th;
  const double* ct;
  int th_per_sample;
  int B;
  int* work;    // Work state per sample
  int* status;  // acados status per sample
  double* cost; // cost of the last linearisation
  int* hard;    // queue of samples for the full interior-point pass
  int* ishard;  // != 0 if the sample was queued in this call (written by k_qp1 only); 2: its workspace holds the first warm iteration
  int subset;   // sens kernels: 0 all samples, 1 samples not queued, 2 the queued samples (via the queue)
  int* counters;  // [0] queue length (front), [1] samples still active, [2] work counter of k_qp3, [3] its interior-point
                  // iterations, [4] samples queued from the end of the array (two_ended)
  const double* x0;  // [B, NX] row-major or null
  const double* u0;  // [B, NU] row-major or null
  double* u0_out;    // [B, NU]
  double* cost_out;  // [B]
  int* status_out;   // [B]
  double* dL;        // [B, ng]
  double* dpi;       // [B, NU, ng]
  double* res_out;   // [B, 4]
  int inplace;       // queue kernels work in place on all samples marked WK_HARD (dense queue: no compact copies)
  int last_round;    // this SQP round only evaluates the convergence test
  int have_solve;    // sens: a solve preceded in this call (keep its status)
  int two_ended;     // k_qp1 files samples that probably need one more iteration at the END of the queue array (see k_qp3)
  double* ipm_state;  // [Engine::IPM_STATE_WORDS][state_stride]: interior-point state of the queued samples between pass kernels
  int state_stride;
  int b0;            // first sample of the range [b0, b0 + B) this launch works on (0 unless the batch is split over two streams)
};

template <class M>
__device__ __forceinline__ Lane make_lane(const KArgs& a, int b) {
  Lane L;
  L.it = a.it + tile_off(b, a.it_size);
  L.ws = a.ws + tile_off(b, a.ws_size);
  // a shared theta is one tile with 32 identical lanes (tile 0)
  L.th = a.th + (a.th_per_sample ? tile_off(b, a.th_size) : (size_t)(b % TILE));
  L.ct = a.ct + (a.th_per_sample ? tile_off(b, a.ct_size) : (size_t)(b % TILE));
  return L;
}

// Which sample does thread j of a queue kernel own, and where are its iterate / workspace?  Sparse queue:
// the j-th queued sample, compact copies in slot j.  Dense queue (inplace): sample j itself if it is queued.
template <class M>
__device__ __forceinline__ bool queue_lane(const KArgs& a, int j, int& b, int& slot, Lane& L) {
  if (a.inplace) {
    if (j >= a.B || a.work[j] != WK_HARD) return false;
    b = slot = j;
    L = make_lane<M>(a, b);
    return true;
  }
  if (j >= a.counters[0]) return false;
  b = a.hard[j];
  if (a.work[b] != WK_HARD) return false;  // finished in a pass kernel
  slot = j;
  L = make_lane<M>(a, b);
  L.it = a.it2 + tile_off(j, a.it_size);
  L.ws = a.ws2 + tile_off(j, a.ws_size);
  return true;
}

// start of a solve call: x_0 (and u_0 in Q-mode) into the iterate, pipeline state reset
template <class M>
__global__ void k_begin(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= a.B) return;
  const int b = a.b0 + bi;
  const Lane L = make_lane<M>(a, b);
  E::set_initial(pd, L, a.x0 + (size_t)b * M::NX, 1, a.u0 ? a.u0 + (size_t)b * M::NU : nullptr, 1);
  a.work[b] = WK_ACTIVE;
  a.status[b] = ST_MAXITER;
}

// (sample, stage) kernel: linearisation.  grid = (ceil(B / blockDim), N + 1)
template <class M>
__global__ void __launch_bounds__(STPB, RLMPC_LIN_MINB) k_lin(const __grid_constant__ ProblemData pd, const KArgs a) {
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = a.b0 + bi;
  if (bi >= a.B || a.work[b] != WK_ACTIVE) return;
  Engine<M>::lin_stage(pd, make_lane<M>(a, b), blockIdx.y);
}

// sample kernel: convergence test + one warm interior-point Newton iteration (fast path)
#ifndef RLMPC_QP1L_MINB
#define RLMPC_QP1L_MINB 8
#endif
// LIN: the linearisation of every stage is done by the sample's own thread inside the sweep (option "fuse_lin")
template <class M, bool LIN>
__global__ void __launch_bounds__(TPB, LIN ? RLMPC_QP1L_MINB : RLMPC_QP1_MINB) k_qp1(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = a.b0 + bi;
  if (bi >= a.B || a.work[b] != WK_ACTIVE) return;
  const Lane L = make_lane<M>(a, b);
  typename E::Residuals R;
  int swept = 0;
  const int code = E::template qp_fast<LIN>(pd, L, R, &swept, /*polish=*/!a.last_round);
  a.cost[b] = R.cost;
  a.ishard[b] = (code == E::FAST_HARD && !a.last_round) ? (swept ? 2 : 1) : 0;
  if (code == E::FAST_NAN) {
    a.status[b] = ST_NAN;
    a.work[b] = WK_DONE;
  } else if (code == E::FAST_CONVERGED) {
    a.status[b] = ST_OK;
    a.work[b] = WK_DONE;
  } else if (a.last_round) {
    a.work[b] = WK_DONE;  // status stays ST_MAXITER
  } else if (code == E::FAST_STEPPED) {
    if (pd.max_sqp == 1) {  // RTI: one QP, no further linearisation here
      a.status[b] = ST_OK;
      a.work[b] = WK_DONE;
    }
  } else {
    a.work[b] = WK_HARD;
    // Longest jobs first: samples whose step was cut short (active-set change: 3..20 iterations) go to the front of
    // the queue array, full steps that only miss the complementarity bound (one more iteration) to its end; the
    // warp-per-sample kernel hands out positions front to back, so the short jobs fill the tail of the launch.
    if (a.two_ended && swept == 2)
      a.hard[a.B - 1 - atomicAdd(&a.counters[4], 1)] = b;
    else
      a.hard[atomicAdd(&a.counters[0], 1)] = b;
  }
}

// sample kernel, one interior-point iteration ("trip") of every queued sample, in place: thread b owns sample b like
// in k_qp1, so every access of a warp is a contiguous run of the tiled arrays, whichever of its samples are queued.
// A queued QP needs 3-4 iterations on the closed-loop workload (1-8 over the batch): running the loop to the end
// inside one thread (k_qp2) makes a warp as slow as its slowest sample, the warp-per-sample kernel (k_qp3) spends ~150
// instructions per sample and stage where a thread spends ~11.  Here the loop is cut into launches instead: pass p
// does iteration p + 1 of every sample that still needs one (Engine::ipm_pass), carries the scalars of the method in
// ipm_state, applies the step of the samples that finish, and the few samples left after the last pass go to k_qp3 /
// k_qp2 as before.
#ifndef RLMPC_PASS_MINB
#define RLMPC_PASS_MINB 4
#endif
template <class M>
__global__ void __launch_bounds__(TPB, RLMPC_PASS_MINB) k_ipm_pass(const __grid_constant__ ProblemData pd, const KArgs a, const int first) {
  using E = Engine<M>;
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = a.b0 + bi;
  if (bi >= a.B || a.work[b] != WK_HARD) return;
  const Lane L = make_lane<M>(a, b);
  int iters = 0;
  const int st = E::ipm_pass(pd, L, a.ipm_state + b, (size_t)a.state_stride, first != 0, a.ishard[b] == 2, &iters);
  if (first) a.ishard[b] = 1;  // the stage records no longer hold the Newton iteration of k_qp1
  if (st < 0) return;
  atomicAdd(&a.counters[3], iters);
  if (pd.max_sqp == 1 || st == E::FULL_FAILED) {
    a.status[b] = (st == E::FULL_OK) ? ST_OK : ST_QPFAIL;
    a.work[b] = WK_DONE;
  } else {
    a.work[b] = WK_ACTIVE;
  }
}

// Reader of the queued (latency-bound) interior-point solves: every lane copies the records of the
// next RING_DEPTH stages of its own sample into a shared-memory ring with cp.async (LDGSTS, no
// registers held across the copy) while the recursion works on the current stage.  ncu on the
// direct-load version: 73 % of the warp stalls were long-scoreboard (profiles/r01_summary.md).
// A lane only ever reads what it copied itself, so no barrier is needed, just wait_group.
constexpr int RING_DEPTH = 3;   // stage-form records (13 KB per stage and warp)
constexpr int RING_DEPTH_B = 2; // condensed block records (38 KB per block and warp)

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}

template <class E, int DEPTH = RING_DEPTH>
struct RingReader {
  static constexpr int SLOT = E::RD_WS + E::RD_ROWS;  // doubles per stage and lane
  const Lane& L;
  int N;
  double* ring;  // this lane's column: element e of slot s at ring[(s * SLOT + e) * TILE]
  int k0, dir, count, consumed;
  __device__ RingReader(const Lane& L_, int N_, double* ring_) : L(L_), N(N_), ring(ring_), k0(0), dir(1), count(0), consumed(0) {}
  __device__ __forceinline__ void issue(int i) {
    if (i < count) {
      const int k = k0 + i * dir;
      double* dst = ring + (size_t)((i % DEPTH) * SLOT) * TILE;
      const double* src = L.ws + (size_t)k * E::W_REC * TILE;
#pragma unroll
      for (int e = 0; e < E::RD_WS; ++e) cp_async8(dst + (size_t)e * TILE, src + (size_t)e * TILE);
      double* dr = dst + (size_t)E::RD_WS * TILE;
      const double* sl = L.it + (size_t)E::it_lam(N, k) * TILE;
      const double* stt = L.it + (size_t)E::it_t(N, k) * TILE;
#pragma unroll
      for (int e = 0; e < E::NR; ++e) {
        cp_async8(dr + (size_t)(E::RD_LAM + e) * TILE, sl + (size_t)e * TILE);
        cp_async8(dr + (size_t)(E::RD_T + e) * TILE, stt + (size_t)e * TILE);
      }
      if (k < N) {
        const double* su = L.it + (size_t)E::it_u(N, k) * TILE;
#pragma unroll
        for (int e = 0; e < E::NU; ++e) cp_async8(dr + (size_t)(E::RD_U + e) * TILE, su + (size_t)e * TILE);
      }
      if (E::NEEDX) {
        const double* sx = L.it + (size_t)E::it_x(N, k) * TILE;
#pragma unroll
        for (int e = 0; e < E::NX; ++e) cp_async8(dr + (size_t)(E::RD_X + e) * TILE, sx + (size_t)e * TILE);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per call, empty or not
  }
  __device__ __forceinline__ void begin(int k_first, int dir_, int count_) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // the previous sweep's stores (K, k, lam, t, lam_hat, t_hat) are re-read through the async copies
    __threadfence();
    k0 = k_first; dir = dir_; count = count_; consumed = 0;
#pragma unroll
    for (int i = 0; i < DEPTH; ++i) issue(i);
  }
  __device__ __forceinline__ const double* ws(int) const {
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    return ring + (size_t)((consumed % DEPTH) * SLOT) * TILE;
  }
  __device__ __forceinline__ void rows(int k, double* lam, double* t, double* u, double* x) const {
    const double* r = ws(k) + (size_t)E::RD_WS * TILE;
    E::template ld<E::NR>(r + (size_t)E::RD_LAM * TILE, TILE, lam);
    E::template ld<E::NR>(r + (size_t)E::RD_T * TILE, TILE, t);
    if (k < N) {
      E::template ld<E::NU>(r + (size_t)E::RD_U * TILE, TILE, u);
    } else {
#pragma unroll
      for (int i = 0; i < E::NU; ++i) u[i] = 0.0;
    }
    if (E::NEEDX) E::template ld<E::NX>(r + (size_t)E::RD_X * TILE, TILE, x);
  }
  __device__ __forceinline__ void done(int) {
    issue(consumed + DEPTH);
    ++consumed;
  }
};

// Queue <-> compact copies.  The queued samples are scattered over the batch; these two kernels
// move their iterate and the linearisation part of their workspace into / out of contiguous tiles
// so that the long interior-point solve streams coalesced data.  One warp per (queue tile, chunk of
// GATHER_CHUNK elements): reads are one sector per lane (scattered), writes are coalesced.
constexpr int GATHER_CHUNK = 64;

template <class M>
__global__ void k_gather(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  const int j = blockIdx.x * TILE + (threadIdx.x & 31);
  if (blockIdx.x * TILE >= a.counters[0]) return;
  const bool live = j < a.counters[0];
  const int b = live ? a.hard[j] : 0;
  const int nit = E::it_size(pd.N), nws = (pd.N + 1) * E::W_K;  // iterate, then [W_A, W_K) of every stage
  const int chunk = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int e0 = chunk * GATHER_CHUNK;
  if (!live || e0 >= nit + nws) return;
  const double* si = a.it + tile_off(b, a.it_size);
  const double* sw = a.ws + tile_off(b, a.ws_size);
  double* di = a.it2 + tile_off(j, a.it_size);
  double* dw = a.ws2 + tile_off(j, a.ws_size);
#pragma unroll 8
  for (int e = e0; e < e0 + GATHER_CHUNK && e < nit + nws; ++e) {
    if (e < nit) {
      di[(size_t)e * TILE] = si[(size_t)e * TILE];
    } else {
      const int q = e - nit, k = q / E::W_K, i = q - k * E::W_K;
      const size_t o = ((size_t)k * E::W_REC + i) * TILE;
      dw[o] = sw[o];
    }
  }
}

template <class M>
__global__ void k_scatter(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  const int j = blockIdx.x * TILE + (threadIdx.x & 31);
  if (j >= a.counters[0]) return;
  const int b = a.hard[j];
  const int nit = E::it_size(pd.N);
  const int chunk = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int e0 = chunk * GATHER_CHUNK;
  const double* si = a.it2 + tile_off(j, a.it_size);
  double* di = a.it + tile_off(b, a.it_size);
#pragma unroll 8
  for (int e = e0; e < e0 + GATHER_CHUNK && e < nit; ++e) di[(size_t)e * TILE] = si[(size_t)e * TILE];
}

// sample kernel over the queue: full interior-point solve on the compact copies.
template <class M, bool RING>
__global__ void __launch_bounds__(32) k_qp2(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int b, slot;
  Lane L;  // theta / cost table of the sample; iterate and workspace: the compact copies (or in place)
  if (!queue_lane<M>(a, j, b, slot, L)) return;
  const int N = pd.N;
  int st;
  if (RING) {
    extern __shared__ double ring_smem[];
    RingReader<E> rd(L, N, ring_smem + threadIdx.x);
    st = E::qp_full(pd, L, nullptr, rd);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    st = E::qp_full(pd, L, nullptr);
  }
  if (pd.max_sqp == 1 || st == E::FULL_FAILED) {
    // RTI: done after one QP.  SQP: an indefinite reduced Hessian ends the solve; an interior-point
    // iteration limit does not (the next linearisation may well be solvable)
    a.status[b] = (st == E::FULL_OK) ? ST_OK : ST_QPFAIL;
    a.work[b] = WK_DONE;
  } else {
    a.work[b] = WK_ACTIVE;
  }
}

// ---- partially condensed queue path (condense.cuh): input-bounds-only problems, V-mode ----
#ifndef RLMPC_CBLK
#define RLMPC_CBLK 4
#endif
constexpr int CBLK = RLMPC_CBLK;  // stages per block
template <class M>
struct Condensable {
  static constexpr bool value = M::NBX == 0 && M::NSX == 0 && M::NG == 0 && CBLK * M::NU <= MAXD;
};

// (queue sample, block) kernel: build the block records and the block-form iterate.  grid = (tiles, N/S + 1)
template <class M>
__global__ void __launch_bounds__(64) k_condense(const __grid_constant__ ProblemData pd, const KArgs a) {
  using Cn = Condenser<M, CBLK>;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int b, slot;
  Lane L;
  if (!queue_lane<M>(a, j, b, slot, L)) return;
  Lane Lb = L;
  Lb.it = a.itb + tile_off(slot, a.itb_size);
  Lb.ws = a.wsb + tile_off(slot, a.wsb_size);
  Cn::condense_block(pd, L, Lb, blockIdx.y);
}

// queue sample kernel: interior-point loop on the blocks, expansion, step
template <class M, bool RING>
__global__ void __launch_bounds__(32) k_qp2c(const __grid_constant__ ProblemData pd, const __grid_constant__ ProblemData pdb,
                                             const KArgs a) {
  using E = Engine<M>;
  using Cn = Condenser<M, CBLK>;
  using EB = typename Cn::EB;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int b, slot;
  Lane L;
  if (!queue_lane<M>(a, j, b, slot, L)) return;
  Lane Lb = L;
  Lb.it = a.itb + tile_off(slot, a.itb_size);
  Lb.ws = a.wsb + tile_off(slot, a.wsb_size);
  int st;
  if (RING) {
    extern __shared__ double ring_smem[];
    RingReader<EB, RING_DEPTH_B> rd(Lb, pdb.N, ring_smem + threadIdx.x);
    st = Cn::solve_expand(pd, pdb, L, Lb, nullptr, rd);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    typename EB::DirectReader rd(Lb, pdb.N);
    st = Cn::solve_expand(pd, pdb, L, Lb, nullptr, rd);
  }
  if (pd.max_sqp == 1 || st == E::FULL_FAILED) {
    a.status[b] = (st == E::FULL_OK) ? ST_OK : ST_QPFAIL;
    a.work[b] = WK_DONE;
  } else {
    a.work[b] = WK_ACTIVE;
  }
}

// ---- warp-per-sample queue path (coop.cuh): one warp solves one queued sample in shared memory ----
#ifndef RLMPC_COOP_WARPS
#define RLMPC_COOP_WARPS 4
#endif
#ifndef RLMPC_COOP_MINB
#define RLMPC_COOP_MINB 4
#endif
constexpr int COOP_WARPS = RLMPC_COOP_WARPS;  // samples in flight per block

// Which cooperative solver, and how many samples per block: the lean one (coop.cuh) where it applies, else the
// general one (coop_general.cuh; more shared memory per sample, one warp per block), else none.
template <class M, bool LEAN = CoopOK<M>::value, bool GEN = CoopGenOK<M>::value>
struct CoopSel {
  static constexpr bool value = false;
  static constexpr int WARPS = 1;
  static constexpr int MINB = 1;
};
template <class M, bool GEN>
struct CoopSel<M, true, GEN> {
  static constexpr bool value = true;
  static constexpr int WARPS = COOP_WARPS;
  static constexpr int MINB = RLMPC_COOP_MINB;  // 14.0 KB of shared memory per sample: 4 blocks of 4 samples per SM (16 warps), 128 registers each
  using Solver = CoopQP<M>;
};
template <class M>
struct CoopSel<M, false, true> {
  static constexpr bool value = true;
  static constexpr int WARPS = 1;
  static constexpr int MINB = 1;
  using Solver = CoopGen<M>;
};

// persistent grid; warps fetch queue positions from counters[2].  Works on the samples' own iterate and
// stage records (no compact copies: the QP lives in shared memory for the whole solve).
template <class M>
__global__ void __launch_bounds__(CoopSel<M>::WARPS * 32, CoopSel<M>::MINB) k_qp3(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  using Cq = typename CoopSel<M>::Solver;
  extern __shared__ double coop_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* S = coop_smem + (size_t)wib * Cq::smem_doubles(pd.N);
  const int nf = a.counters[0], n = nf + a.counters[4];  // queue: [0, nf) from the front, the rest from the end of the array
  for (;;) {
    int j = 0;
    if (lane == 0) j = atomicAdd(&a.counters[2], 1);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= n) break;
    const int b = a.hard[j < nf ? j : a.B - 1 - (j - nf)];
    if (a.work[b] != WK_HARD) continue;  // finished in a pass kernel (warp-uniform: one sample per warp)
    const Lane L = make_lane<M>(a, b);
    int iters = 0;
    const int st = Cq::solve(pd, L, S, lane, a.ishard[b] == 2, &iters);
    if (lane == 0) {
      atomicAdd(&a.counters[3], iters);
      if (pd.max_sqp == 1 || st == E::FULL_FAILED) {
        a.status[b] = (st == E::FULL_OK) ? ST_OK : ST_QPFAIL;
        a.work[b] = WK_DONE;
      } else {
        a.work[b] = WK_ACTIVE;
      }
    }
    __syncwarp();
  }
}

template <class M>
__global__ void k_count_active(const KArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = b < a.B && a.work[b] != WK_DONE;
  const unsigned m = __ballot_sync(0xffffffffu, act);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&a.counters[1], __popc(m));
}

// outputs of a solve-only call
template <class M>
__global__ void k_out(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= a.B) return;
  const int b = a.b0 + bi;
  if (a.u0_out) {
#pragma unroll
    for (int i = 0; i < M::NU; ++i)
      a.u0_out[(size_t)b * M::NU + i] = a.it[tile_off(b, a.it_size) + (size_t)(E::it_u(pd.N, 0) + i) * TILE];
  }
  if (a.cost_out) a.cost_out[b] = a.cost[b];
  if (a.status_out) a.status_out[b] = a.status[b];
}

// (sample, stage) kernel: exact second-order stage information.  grid = (ceil(B / blockDim), N + 1)
__device__ __forceinline__ int subset_sample(const KArgs& a) {
  // which sample does this thread own?  -1: none
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (a.subset == 2) return j < a.counters[0] ? a.hard[j] : -1;
  if (j >= a.B) return -1;
  if (a.subset == 1 && a.ishard[a.b0 + j]) return -1;
  return a.b0 + j;
}

template <class M>
__global__ void __launch_bounds__(STPB, RLMPC_SS_MINB) k_sens_stage(const __grid_constant__ ProblemData pd, const KArgs a) {
  const int b = subset_sample(a);
  if (b < 0) return;
  Engine<M>::sens_stage(pd, make_lane<M>(a, b), blockIdx.y);
}

// sample kernel: residuals, dL/dtheta, exact-Hessian factorisation, adjoint solves, outputs
template <class M>
__global__ void __launch_bounds__(TPB, RLMPC_SW_MINB) k_sens_sweep(const __grid_constant__ ProblemData pd, const KArgs a) {
  using E = Engine<M>;
  const int b = subset_sample(a);
  if (b < 0) return;
  const Lane L = make_lane<M>(a, b);
  int ok = 1;
  const int ng = E::grad_width(pd);
  const typename E::Residuals r = E::sens_sweep(pd, L, a.dL ? a.dL + (size_t)b * ng : nullptr,
                                                a.dpi ? a.dpi + (size_t)b * M::NU * ng : nullptr, &ok);
  if (a.res_out) {
    a.res_out[(size_t)b * 4 + 0] = r.stat;
    a.res_out[(size_t)b * 4 + 1] = r.eq;
    a.res_out[(size_t)b * 4 + 2] = r.ineq;
    a.res_out[(size_t)b * 4 + 3] = r.comp;
  }
  const double rmax = dmax(dmax(r.stat, r.eq), dmax(r.ineq, r.comp));
  int status = a.have_solve ? a.status[b] : ST_OK;
  if (!(rmax == rmax)) status = ST_NAN;
  if (!a.have_solve) status = (rmax == rmax) ? (rmax < pd.tol ? ST_OK : ST_MAXITER) : ST_NAN;
  if (!ok && status == ST_OK) status = ST_QPFAIL;  // reduced Hessian not PD: sensitivities invalid
  if (a.u0_out) {
#pragma unroll
    for (int i = 0; i < M::NU; ++i) a.u0_out[(size_t)b * M::NU + i] = L.it[(size_t)(E::it_u(pd.N, 0) + i) * TILE];
  }
  if (a.cost_out) a.cost_out[b] = r.cost;
  if (a.status_out) a.status_out[b] = status;
}

// theta -> quadratic cost table (shared: one thread; per sample: one thread per sample)
template <class M>
__global__ void k_cost_table(const __grid_constant__ ProblemData pd, const double* th, double* ct, int per_sample, int B) {
  using E = Engine<M>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (per_sample ? B : TILE)) return;  // shared theta: the 32 lanes of tile 0
  M::cost_table(th + tile_off(b, M::NTH), TILE, ct + tile_off(b, E::CT_SIZE), TILE, pd.mc);
}

// Warm-start store coupled to a replay buffer: slot idx[b] of a caller-owned store
// (same tiled layout, `capacity` samples) <-> sample b of the handle's iterate.  One warp per
// (batch tile, 64-element chunk); the side addressed through idx is sector-granular, the other coalesced.
__global__ void k_store_copy(double* it, double* store, int it_size, int B, const int* idx, int capacity, int to_store) {
  const int b = blockIdx.x * TILE + (threadIdx.x & 31);
  if (b >= B) return;
  const int slot = idx[b];
  if (slot < 0 || slot >= capacity) return;
  const int chunk = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int e0 = chunk * 64;
  double* pi_ = it + tile_off(b, it_size);
  double* ps = store + tile_off(slot, it_size);
#pragma unroll 8
  for (int e = e0; e < e0 + 64 && e < it_size; ++e) {
    if (to_store)
      ps[(size_t)e * TILE] = pi_[(size_t)e * TILE];
    else
      pi_[(size_t)e * TILE] = ps[(size_t)e * TILE];
  }
}

// Vectorised continuous cart-pole swing-up environment, one thread per environment
// (rlmpc/gym/continuous_cartpole/environment.py:372-426: explicit Euler, tau = 0.02; note the env uses
// polemass_length = m*l in `temp` where the MPC model uses m, quirk Q9).
// par: [gravity, masscart, masspole, length, force_mag, tau, x_threshold, theta_threshold,
//       max_episode_steps, reset_state(4)]
__global__ void k_cartpole_env_step(const double* __restrict__ par, int B, double* state, const double* action,
                                    double* reward, int* terminated, int* truncated, int* steps) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double g = par[0], mc = par[1], mp = par[2], len = par[3], fmag = par[4], tau = par[5];
  const double xth = par[6], thth = par[7];
  const int max_steps = (int)par[8];
  double x = state[4 * b], xd = state[4 * b + 1], th = state[4 * b + 2], thd = state[4 * b + 3];
  const double a = action[b];
  const double force = a * fmag, total = mp + mc, pml = mp * len;
  double sn, cs;
  sincos(th, &sn, &cs);
  const double temp = (force + pml * thd * thd * sn) / total;
  const double thacc = (g * sn - cs * temp) / (len * (4.0 / 3.0 - mp * cs * cs / total));
  const double xacc = temp - pml * thacc * cs / total;
  x += tau * xd;
  xd += tau * xacc;
  th += tau * thd;
  thd += tau * thacc;
  const int term = (x < -xth) || (x > xth) || (th < -thth) || (th > thth);
  const int st = steps[b] + 1;
  const int trunc = st >= max_steps;
  // reward of the state reached (environment.py:448-456), angle wrapped to [-pi, pi)
  const double pi = 3.14159265358979323846;
  double an = fmod(th + pi, 2.0 * pi);
  if (an < 0.0) an += 2.0 * pi;
  an -= pi;
  reward[b] = 2.0 * x * x + 0.01 * xd * xd + 2.0 * an * an + 0.01 * thd * thd + 0.001 * a * a;
  terminated[b] = term;
  truncated[b] = trunc;
  if (term || trunc) {  // auto-reset (environment.py:416-419) to the reset state
    x = par[9]; xd = par[10]; th = par[11]; thd = par[12];
    steps[b] = 0;
  } else {
    steps[b] = st;
  }
  state[4 * b] = x; state[4 * b + 1] = xd; state[4 * b + 2] = th; state[4 * b + 3] = thd;
}

// FP64 FMA throughput probe: 8 independent dependent-chains per thread, enough warps to fill every scheduler.
// The roofline the path is measured against ("FP64 CUDA-core peaks are not in MEASURED_PEAKS.json --
// measure them first").
__global__ void k_fp64_peak(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-3 * (threadIdx.x + j);
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == 12345.678) out[0] = s;  // keeps the loop alive
}

// FP64 tensor-core probe: 8 independent m8n8k4 accumulators per warp (DMMA), operands in registers
__global__ void k_dmma_peak(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
  const double a = 1.0 + 1e-3 * (threadIdx.x & 31), b = 1.0 - 1e-3 * (threadIdx.x & 7);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[j][0]), "+d"(c[j][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  if (s == 12345.678) out[0] = s;
}

// MPC.reset: x_k = x0 for all stages, everything else zero
template <class M>
__global__ void k_reset(int N, double* it, int B, const double* x0, const int* mask) {
  using E = Engine<M>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || (mask && !mask[b])) return;
  const int n = E::it_size(N);
  double* p = it + tile_off(b, n);
  for (int i = 0; i < n; ++i) {
    double v = 0.0;
    if (x0 && i < (N + 1) * M::NX) v = x0[(size_t)b * M::NX + (i % M::NX)];
    p[(size_t)i * TILE] = v;
  }
}

// gather/scatter one field of one stage between the SoA iterate and a row-major [B, dim] buffer
__global__ void k_copy_field(double* it, int it_size, int B, int off, int dim, double* buf, int to_iterate) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double* p = it + tile_off(b, it_size);
  for (int i = 0; i < dim; ++i) {
    if (to_iterate)
      p[(size_t)(off + i) * TILE] = buf[(size_t)b * dim + i];
    else
      buf[(size_t)b * dim + i] = p[(size_t)(off + i) * TILE];
  }
}

// theta [B, nth] row-major -> tiled; shared = 1: the single theta `in` replicated into the lanes of tile 0
__global__ void k_theta_transpose(const double* in, double* out, int B, int nth, int shared) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (shared ? TILE : B)) return;
  double* p = out + tile_off(b, nth);
  for (int i = 0; i < nth; ++i) p[(size_t)i * TILE] = in[shared ? (size_t)i : (size_t)b * nth + i];
}

// acc[j] += sum_b td_b * dQ[b, j] over valid samples; acc[nth] += sum td; acc[nth+1] += count
__global__ void k_td_grad(int B, int nth, const double* td, const double* dQ, const int* status, double* acc) {
  extern __shared__ double sm[];  // [nth + 2]
  for (int j = threadIdx.x; j < nth + 2; j += blockDim.x) sm[j] = 0.0;
  __syncthreads();
  // each warp walks rows; lanes walk theta entries (dQ is row-major [B, nth]: coalesced along j)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int b = blockIdx.x * nwarp + warp; b < B; b += gridDim.x * nwarp) {
    if (status && status[b] != 0) continue;
    const double t = td[b];
    for (int j = lane; j < nth; j += 32) atomicAdd(&sm[j], t * dQ[(size_t)b * nth + j]);
    if (lane == 0) {
      atomicAdd(&sm[nth], t);
      atomicAdd(&sm[nth + 1], 1.0);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nth + 2; j += blockDim.x)
    if (sm[j] != 0.0) atomicAdd(&acc[j], sm[j]);
}

}  // namespace

enum Variant : int { VAR_CARTPOLE = 0, VAR_CARTPOLE_BX = 1, VAR_LINEAR = 2, VAR_EVAPORATION = 3, VAR_CARTPOLE_G = 4, VAR_CARTPOLE_BX_G = 5 };

constexpr int MAX_SPLIT = 4;
constexpr int NCNT = 8;  // ints per set of queue counters

struct rlmpc_handle {
  ChainBackend* chain = nullptr;  // chain-mass problems run on the warp-cooperative engine (rlmpc_chain.cu)
  int model, variant, device, max_batch;
  size_t bs;
  int nx, nu, nth, npm, nr, it_size, ws_size, ct_size;
  int ng() const { return pd.param_cost ? nth : npm; }
  ProblemData pd;
  double *it = nullptr, *ws = nullptr, *it2 = nullptr, *ws2 = nullptr, *th = nullptr, *ct = nullptr, *th_stage = nullptr;
  double *itb = nullptr, *wsb = nullptr;  // partially condensed queue path (only for condensable models)
  int itb_size = 0, wsb_size = 0;
  int condense = 1;    // 1: queued QPs in partially condensed form where applicable (input bounds only, V-mode, N % 4 == 0)
  int split = 2;       // n > 1: an RTI call runs n parts of its batch on n streams (see pipeline())
  cudaStream_t part_stream[MAX_SPLIT - 1] = {};
  cudaEvent_t ev_part[MAX_SPLIT - 1] = {};
  bool marks_off = false;
  int coop = 1;        // warp-per-sample queue kernel (coop.cuh / coop_general.cuh) where the model's blocks fit a warp
  int coop_grid = 0;   // its persistent grid (blocks), sized at create time from the occupancy
  int ring_b = 0;      // ring reader for the condensed kernel: measured slower (3.3 vs 2.3 ms), blocks carry enough work per load batch
                       // (an L1 prefetch of the next block record, CCTL.E.PF1, measured the same 3.3 ms: profiles/r01g_variants_prefetch_overlap.log)
  double* cost = nullptr;
  int *work = nullptr, *status = nullptr, *hard = nullptr, *ishard = nullptr, *counters = nullptr;
  int ring = 1;     // queued interior-point pass reads through the cp.async shared-memory ring (0: direct loads)
  int inplace_queue = 0;
  int fuse_lin = 0;    // 1: k_qp1 linearises the stages itself (no k_lin launch).  Measured equal at best: the fused kernel
                       // takes 0.66 ms (128 registers) against 0.17 + 0.50 ms of the two launches, the step 2.42 vs 2.40 ms --
                       // the FP64 work of the linearisation does not hide under the sweep's memory time at 65 536 threads
  int ipm_passes = 0;  // pass kernels (one interior-point iteration of every queued sample each) before the queue kernel.
                       // Measured slower on every workload (headline: queue 2.48 -> 3.76 ms with 5 passes, evaporation
                       // 4.05 -> 6.56 ms): a pass costs its full ~0.45 ms as long as most WARPS still hold one queued
                       // sample, whatever the fraction of active lanes (profiles/r02_summary.md), so it is off by default
  double* ipm_state = nullptr;
  int overlap = 0;  // 1: RTI + sens runs the full interior-point pass of the queued samples on a side stream,
                    // concurrently with the sensitivity kernels of all other samples.  Measured slower
                    // (the few latency-bound warps of the queue lose issue slots to the bulk kernels).
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int* h_counters = nullptr;  // pinned
  int th_per_sample = 0;
  int sync_every = 4;  // SQP rounds between host checks of the active-sample counter (max_sqp > 1)
  int timing = 0;      // 1: record CUDA events between the phases of a call (rlmpc_get_timings)
  static constexpr int NEV = 7;
  cudaEvent_t ev[NEV] = {};
  bool ev_set[NEV] = {};
  long long launches = 0;
  // staging for the host-buffer entry point
  double *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  int *h_status = nullptr, *d_status = nullptr;
  cudaStream_t own_stream = nullptr;
  // CUDA graph of the RTI solve+sens chain (option "graph"): captured on the first call with a given argument set,
  // replayed while the arguments (mode, batch, every pointer, options) stay the same -- small batches are launch-bound
  int use_graph = 0;
  cudaGraphExec_t graph_exec = nullptr;
  unsigned long long graph_key = 0;
  long long graph_launches = 0;
  cudaEvent_t ev_last = nullptr;  // recorded after every stream-ordered call: the host entry point waits for it
  bool ev_last_set = false;
};

namespace {

// run `expr` with M bound to the model type of the handle
#define DISPATCH_MODEL(h, ...)                                                \
  switch ((h)->variant) {                                                     \
    case VAR_CARTPOLE: { using M = CartpoleModel; __VA_ARGS__; } break;       \
    case VAR_CARTPOLE_BX: { using M = CartpoleModelBX; __VA_ARGS__; } break;  \
    case VAR_CARTPOLE_G: { using M = CartpoleModelG; __VA_ARGS__; } break;    \
    case VAR_CARTPOLE_BX_G: { using M = CartpoleModelBXG; __VA_ARGS__; } break; \
    case VAR_LINEAR: { using M = LinearSystemModel; __VA_ARGS__; } break;     \
    case VAR_EVAPORATION: { using M = EvaporationModel; __VA_ARGS__; } break; \
  }

// stream-ordered entry points leave a marker that rlmpc_solve_sens_host (which runs on the handle's own streams) waits for
void note_stream(rlmpc_handle* h, cudaStream_t s) {
  if (!h->ev_last) cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming);
  if (h->ev_last && cudaEventRecord(h->ev_last, s) == cudaSuccess) h->ev_last_set = true;
}

template <class M>
constexpr bool coop_model() { return CoopSel<M>::value; }

int check_batch(rlmpc_handle* h, int B) {
  if (!h) return fail(RLMPC_EINVAL, "null handle");
  if (B < 0 || B > h->max_batch) return fail(RLMPC_EINVAL, "batch exceeds max_batch");
  return 0;
}

// phase boundaries: 0 start | 1 after k_lin | 2 after k_qp1 | 3 after k_qp2 | 4 after k_sens_stage | 5 after
// k_sens_sweep | 6 end of the call (sens kernels of the queued samples when the side stream is used)
void mark(rlmpc_handle* h, int i, cudaStream_t s) {
  if (!h->timing || h->marks_off) return;
  cudaEventRecord(h->ev[i], s);
  h->ev_set[i] = true;
}

KArgs base_args(rlmpc_handle* h, int B) {
  KArgs a;
  memset(&a, 0, sizeof(a));
  a.it = h->it; a.ws = h->ws; a.it2 = h->it2; a.ws2 = h->ws2;
  a.itb = h->itb; a.wsb = h->wsb; a.itb_size = h->itb_size; a.wsb_size = h->wsb_size;
  a.it_size = h->it_size; a.ws_size = h->ws_size; a.th_size = h->nth; a.ct_size = h->ct_size;
  a.th = h->th; a.ct = h->ct; a.th_per_sample = h->th_per_sample; a.B = B;
  a.work = h->work; a.status = h->status; a.cost = h->cost; a.hard = h->hard; a.counters = h->counters;
  a.ishard = h->ishard;
  a.ipm_state = h->ipm_state; a.state_stride = (int)h->bs;
  return a;
}

template <class M>
constexpr size_t qp2_smem() {
  return sizeof(double) * RING_DEPTH * RingReader<Engine<M>>::SLOT * TILE;
}
template <class M>
constexpr size_t qp2c_smem() {
  return sizeof(double) * RING_DEPTH_B * RingReader<typename Condenser<M, CBLK>::EB, RING_DEPTH_B>::SLOT * TILE;
}

template <class M>
cudaError_t alloc_condensed(rlmpc_handle* h, int N, cudaError_t e) {
  if constexpr (Condensable<M>::value) {
    using EB = typename Condenser<M, CBLK>::EB;
    if (N % CBLK == 0) {
      h->itb_size = EB::it_size(N / CBLK);
      h->wsb_size = EB::ws_size(N / CBLK);
      if (e == cudaSuccess) e = cudaMalloc(&h->itb, sizeof(double) * h->itb_size * h->bs);
      if (e == cudaSuccess) e = cudaMalloc(&h->wsb, sizeof(double) * h->wsb_size * h->bs);
      if (e == cudaSuccess) e = cudaMemset(h->itb, 0, sizeof(double) * h->itb_size * h->bs);
      if (e == cudaSuccess) e = cudaMemset(h->wsb, 0, sizeof(double) * h->wsb_size * h->bs);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_qp2c<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qp2c_smem<M>());
    }
  }
  return e;
}

template <class M>
size_t coop_smem(int N) {
  if constexpr (CoopSel<M>::value) return sizeof(double) * CoopSel<M>::WARPS * CoopSel<M>::Solver::smem_doubles(N);
  return 0;
}
// persistent grid of the warp-per-sample queue kernel: as many blocks as fit on the device
template <class M>
cudaError_t setup_coop(rlmpc_handle* h, int N, cudaError_t e) {
  if constexpr (CoopSel<M>::value) {
    const size_t smem = coop_smem<M>(N);
    int dev = 0, sms = 0, per_sm = 0, max_optin = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess || smem > (size_t)max_optin) return e;  // horizon too long for shared memory: coop_grid stays 0
    // the attribute belongs to the kernel, not to the handle: opt in to the device maximum once, so that handles
    // with different horizons can coexist (the launch passes the size this handle needs)
    e = cudaFuncSetAttribute(k_qp3<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_qp3<M>, CoopSel<M>::WARPS * 32, smem);
    if (const char* cap = getenv("RLMPC_COOP_BLOCKS_PER_SM")) {  // tuning / debugging aid
      const int c = atoi(cap);
      if (c > 0 && c < per_sm) per_sm = c;
    }
    if (e == cudaSuccess) h->coop_grid = sms * per_sm;
  }
  return e;
}

// the queue pass between gather and scatter: condensed where possible, else stage form
template <class M>
void launch_queue_solve(rlmpc_handle* h, const KArgs& a, int B, cudaStream_t sq) {
  if constexpr (Condensable<M>::value) {
    using Cn = Condenser<M, CBLK>;
    if (h->condense && h->itb && Cn::applicable(h->pd)) {
      const ProblemData pdb = Cn::block_pd(h->pd);
      k_condense<M><<<dim3((B + 63) / 64, pdb.N + 1), 64, 0, sq>>>(h->pd, a);
      if (h->ring_b)
        k_qp2c<M, true><<<(B + 31) / 32, 32, qp2c_smem<M>(), sq>>>(h->pd, pdb, a);
      else
        k_qp2c<M, false><<<(B + 31) / 32, 32, 0, sq>>>(h->pd, pdb, a);
      h->launches += 2;
      return;
    }
  }
  if (h->ring)
    k_qp2<M, true><<<(B + 31) / 32, 32, qp2_smem<M>(), sq>>>(h->pd, a);
  else
    k_qp2<M, false><<<(B + 31) / 32, 32, 0, sq>>>(h->pd, a);
  h->launches++;
}

// SQP: K rounds of (linearise | convergence test + fast QP | full interior point on the queue),
// then one test-only round.  RTI (K = 1) is a single round without the final test.
template <class M>
int pipeline_solve(rlmpc_handle* h, KArgs a, cudaStream_t s, bool fork_qp2) {
  const int B = a.B, N = h->pd.N, K = h->pd.max_sqp;
  const int gs = (B + TPB - 1) / TPB;
  const dim3 gstage((B + STPB - 1) / STPB, N + 1);
  k_begin<M><<<(B + 127) / 128, 128, 0, s>>>(h->pd, a);
  h->launches++;
  if (!h->marks_off)
    for (int i = 0; i < rlmpc_handle::NEV; ++i) h->ev_set[i] = false;
  const int rounds = (K == 1) ? 1 : K + 1;
  // Option "inplace_queue": the queue kernels work in place on the samples marked WK_HARD and the gather /
  // scatter copies are skipped.  Only pays when (nearly) every sample is queued; measured slower for a
  // cold SQP-to-convergence solve as a whole (90 vs 86 ms per 65 536: after the first round the queue is
  // sparse and in-place warps carry few active lanes), so it is off by default.
  bool dense_queue = (K > 1) && h->inplace_queue;
  for (int r = 0; r < rounds; ++r) {
    a.last_round = (K > 1 && r == K) ? 1 : 0;
    a.inplace = dense_queue ? 1 : 0;
    CUDA_OK(cudaMemsetAsync(a.counters, 0, NCNT * sizeof(int), s));
    mark(h, 0, s);
    if constexpr (CoopSel<M>::value) a.two_ended = (h->coop && h->coop_grid > 0 && !a.inplace && !fork_qp2) ? 1 : 0;
    if (h->fuse_lin) {
      mark(h, 1, s);
      k_qp1<M, true><<<gs, TPB, 0, s>>>(h->pd, a);
      h->launches += 1;
    } else {
      k_lin<M><<<gstage, STPB, 0, s>>>(h->pd, a);
      mark(h, 1, s);
      k_qp1<M, false><<<gs, TPB, 0, s>>>(h->pd, a);
      h->launches += 2;
    }
    mark(h, 2, s);
    if (!a.last_round) {
      cudaStream_t sq = s;
      if (fork_qp2) {  // queued samples continue on the side stream; the caller joins on ev_join
        CUDA_OK(cudaEventRecord(h->ev_fork, s));
        CUDA_OK(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        sq = h->side_stream;
      }
      for (int p = 0; p < h->ipm_passes; ++p) {
        k_ipm_pass<M><<<gs, TPB, 0, sq>>>(h->pd, a, p == 0 ? 1 : 0);
        h->launches++;
      }
      bool coop_done = false;
      if constexpr (CoopSel<M>::value) {
        if (h->coop && h->coop_grid > 0 && !a.inplace) {
          k_qp3<M><<<h->coop_grid, CoopSel<M>::WARPS * 32, coop_smem<M>(N), sq>>>(h->pd, a);
          h->launches++;
          coop_done = true;
        }
      }
      if (!coop_done) {
        using E = Engine<M>;
        const int wpb = 8;  // warps per block of the copy kernels
        const int n_g = (E::it_size(N) + (N + 1) * E::W_K + GATHER_CHUNK - 1) / GATHER_CHUNK;
        const int n_s = (E::it_size(N) + GATHER_CHUNK - 1) / GATHER_CHUNK;
        if (!a.inplace) k_gather<M><<<dim3((B + 31) / 32, (n_g + wpb - 1) / wpb), 32 * wpb, 0, sq>>>(h->pd, a);
        launch_queue_solve<M>(h, a, B, sq);
        if (!a.inplace) k_scatter<M><<<dim3((B + 31) / 32, (n_s + wpb - 1) / wpb), 32 * wpb, 0, sq>>>(h->pd, a);
        h->launches += 2;
      }
      mark(h, 3, sq);
      if (fork_qp2) CUDA_OK(cudaEventRecord(h->ev_join, sq));
    }
    if (K > 1 && !a.last_round && (r % h->sync_every) == h->sync_every - 1) {
      // all samples converged?  (the only host synchronisation of the library; RTI never gets here)
      k_count_active<M><<<(B + 127) / 128, 128, 0, s>>>(a);
      h->launches++;
      CUDA_OK(cudaMemcpyAsync(h->h_counters, h->counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaStreamSynchronize(s));
      if (h->h_counters[1] == 0) break;
      dense_queue = dense_queue && h->h_counters[0] > B / 2;
    }
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <class M>
int pipeline_sens(rlmpc_handle* h, KArgs a, cudaStream_t s, bool forked) {
  const int B = a.B, N = h->pd.N;
  const dim3 gstage((B + STPB - 1) / STPB, N + 1);
  if (!a.have_solve) {
    for (int i = 0; i < rlmpc_handle::NEV; ++i) h->ev_set[i] = false;
    mark(h, 3, s);
  }
  a.subset = forked ? 1 : 0;
  k_sens_stage<M><<<gstage, STPB, 0, s>>>(h->pd, a);
  mark(h, 4, s);
  k_sens_sweep<M><<<(B + TPB - 1) / TPB, TPB, 0, s>>>(h->pd, a);
  mark(h, 5, s);
  h->launches += 2;
  if (forked) {  // join: the queued samples now have their step; same two kernels over the queue
    CUDA_OK(cudaStreamWaitEvent(s, h->ev_join, 0));
    a.subset = 2;
    k_sens_stage<M><<<gstage, STPB, 0, s>>>(h->pd, a);
    k_sens_sweep<M><<<(B + 31) / 32, 32, 0, s>>>(h->pd, a);
    h->launches += 2;
  }
  mark(h, 6, s);
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <class M>
int pipeline_range(rlmpc_handle* h, KArgs a, int do_solve, int do_sens, cudaStream_t s) {
  const bool fork = do_solve && do_sens && h->pd.max_sqp == 1 && h->overlap;
  if (do_solve) {
    if (int r = pipeline_solve<M>(h, a, s, fork)) return r;
  }
  a.have_solve = do_solve;
  if (do_sens) return pipeline_sens<M>(h, a, s, fork);
  k_out<M><<<(a.B + 127) / 128, 128, 0, s>>>(h->pd, a);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// One RTI call = a chain of dependent kernels that are bound by different units (k_lin, k_sens_stage: FP64
// issue; k_qp1, k_sens_sweep: HBM; k_qp3: shared memory and latency).  Samples are independent, so with option
// "split" = 2 the batch is cut in two halves that run the same chain on two streams: the kernels of one half
// overlap with kernels of the other half that wait for a different unit.
template <class M>
int pipeline(rlmpc_handle* h, KArgs a, int do_solve, int do_sens, cudaStream_t s) {
  bool split = h->split > 1 && do_solve && h->pd.max_sqp == 1 && !h->overlap && a.B >= 4096;
  if constexpr (CoopSel<M>::value) {
    split = split && h->coop && h->coop_grid > 0;
  } else {
    split = false;
  }
  if (!split) return pipeline_range<M>(h, a, do_solve, do_sens, s);
  const int parts = h->split < MAX_SPLIT ? h->split : MAX_SPLIT;
  const int per = ((a.B + parts - 1) / parts + TILE - 1) / TILE * TILE;
  CUDA_OK(cudaEventRecord(h->ev_fork, s));
  int rc = 0;
  for (int p = 0; p < parts && rc == 0; ++p) {
    KArgs ap = a;
    ap.b0 = p * per;
    ap.B = a.B - ap.b0 < per ? a.B - ap.b0 : per;
    if (ap.B <= 0) break;
    ap.hard = a.hard + ap.b0;
    ap.counters = a.counters + NCNT * p;
    cudaStream_t sp = (p == 0) ? s : h->part_stream[p - 1];
    if (p > 0) CUDA_OK(cudaStreamWaitEvent(sp, h->ev_fork, 0));
    h->marks_off = p > 0;  // phase events (option "timing") describe the first part
    rc = pipeline_range<M>(h, ap, do_solve, do_sens, sp);
    h->marks_off = false;
    if (p > 0 && rc == 0) {
      CUDA_OK(cudaEventRecord(h->ev_part[p - 1], sp));
      CUDA_OK(cudaStreamWaitEvent(s, h->ev_part[p - 1], 0));
    }
  }
  return rc;
}

// Works on the samples [b0, b0 + B); array arguments are indexed by the absolute sample number.  part > 0: a
// further range of the same logical call, issued on another stream by the caller (own queue counters, no
// phase events, no internal split).
int run_unit(rlmpc_handle* h, int mode, int max_sqp, int B, const double* x0, const double* u0, double* u0_out,
             double* cost_out, int* status_out, double* dL, double* dpi, double* res_out, int do_solve, int do_sens,
             cudaStream_t s, int b0 = 0, int part = -1) {
  if (int r = check_batch(h, b0 + B)) return r;
  if (B == 0) return 0;
  if (mode != RLMPC_MODE_V && mode != RLMPC_MODE_Q) return fail(RLMPC_EINVAL, "bad mode");
  if (do_solve && !x0) return fail(RLMPC_EINVAL, "x0 is required");
  if (do_solve && mode == RLMPC_MODE_Q && !u0) return fail(RLMPC_EINVAL, "u0 is required in Q-mode");
  if (do_solve && max_sqp < 1) return fail(RLMPC_EINVAL, "max_sqp must be >= 1");
  CUDA_OK(cudaSetDevice(h->device));
  h->pd.mode = mode;
  h->pd.max_sqp = max_sqp;
  if (h->chain) {
    ChainCall c;
    c.mode = mode; c.max_sqp = max_sqp; c.B = B; c.do_solve = do_solve; c.do_sens = do_sens;
    const size_t o = (size_t)b0;
    c.x0 = x0 ? x0 + o * h->nx : nullptr; c.u0 = u0 ? u0 + o * h->nu : nullptr;
    c.u0_out = u0_out ? u0_out + o * h->nu : nullptr; c.cost_out = cost_out ? cost_out + o : nullptr;
    c.status_out = status_out ? status_out + o : nullptr; c.dL = dL ? dL + o * h->nth : nullptr;
    c.dpi = dpi ? dpi + o * h->nth * h->nu : nullptr; c.res_out = res_out ? res_out + o * 4 : nullptr;
    if (b0 != 0) return fail(RLMPC_EINVAL, "chain mass: split batches are not supported");
    std::string err;
    const int rc = chain_run(h->chain, h->pd, c, h->sync_every, s, err);
    return rc ? fail(rc, err) : 0;
  }
  KArgs a = base_args(h, B);
  a.x0 = x0; a.u0 = (mode == RLMPC_MODE_Q) ? u0 : nullptr;
  a.u0_out = u0_out; a.cost_out = cost_out; a.status_out = status_out;
  a.dL = dL; a.dpi = dpi; a.res_out = res_out;
  int rc = fail(RLMPC_EINVAL, "unknown model");
  if (part < 0) {
    const bool graphable = h->use_graph && do_solve && do_sens && max_sqp == 1 && !h->timing && !h->overlap;
    if (graphable) {
      // FNV-1a over everything the captured kernels bake in
      unsigned long long key = 1469598103934665603ull;
      auto mix = [&](const void* p, size_t n) {
        const unsigned char* c = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) { key ^= c[i]; key *= 1099511628211ull; }
      };
      mix(&a, sizeof(a)); mix(&h->pd, sizeof(h->pd)); mix(&h->split, sizeof(int)); mix(&h->coop, sizeof(int)); mix(&h->ipm_passes, sizeof(int)); mix(&h->fuse_lin, sizeof(int));
      mix(&h->th_per_sample, sizeof(int)); mix(&s, sizeof(s));
      if (h->graph_exec && key == h->graph_key) {
        CUDA_OK(cudaGraphLaunch(h->graph_exec, s));
        h->launches += h->graph_launches;
        return 0;
      }
      if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
      const long long l0 = h->launches;
      cudaGraph_t g = nullptr;
      // (the legacy default stream cannot be captured: capture on the handle's own stream, the graph is launched on s)
      cudaStream_t cs = (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread) ? h->own_stream : s;
      CUDA_OK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      DISPATCH_MODEL(h, rc = pipeline<M>(h, a, do_solve, do_sens, cs));
      cudaError_t ec = cudaStreamEndCapture(cs, &g);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (ec != cudaSuccess) return fail(RLMPC_ECUDA, std::string("graph capture: ") + cudaGetErrorString(ec));
      ec = cudaGraphInstantiate(&h->graph_exec, g, 0);
      cudaGraphDestroy(g);
      if (ec != cudaSuccess) { h->graph_exec = nullptr; return fail(RLMPC_ECUDA, std::string("graph instantiate: ") + cudaGetErrorString(ec)); }
      h->graph_key = key;
      h->graph_launches = h->launches - l0;
      CUDA_OK(cudaGraphLaunch(h->graph_exec, s));
      return 0;
    }
    DISPATCH_MODEL(h, rc = pipeline<M>(h, a, do_solve, do_sens, s));
    return rc;
  }
  a.b0 = b0;
  a.hard = h->hard + b0;
  a.counters = h->counters + NCNT * part;
  h->marks_off = part > 0;
  DISPATCH_MODEL(h, rc = pipeline_range<M>(h, a, do_solve, do_sens, s));
  h->marks_off = false;
  return rc;
}

struct Field { int off, dim; };

template <class M>
int field_offset_t(rlmpc_handle* h, const char* field, int stage, int* off, int* dim) {
  using E = Engine<M>;
  const int N = h->pd.N;
  if (!strcmp(field, "x")) {
    if (stage < 0 || stage > N) return fail(RLMPC_EINVAL, "stage out of range");
    *off = E::it_x(N, stage); *dim = E::NX;
  } else if (!strcmp(field, "u")) {
    if (stage < 0 || stage >= N) return fail(RLMPC_EINVAL, "stage out of range");
    *off = E::it_u(N, stage); *dim = E::NU;
  } else if (!strcmp(field, "pi")) {
    if (stage < 0 || stage >= N) return fail(RLMPC_EINVAL, "stage out of range");
    *off = E::it_pi(N, stage); *dim = E::NX;
  } else if (!strcmp(field, "lam")) {
    if (stage < 0 || stage > N) return fail(RLMPC_EINVAL, "stage out of range");
    *off = E::it_lam(N, stage); *dim = E::NR;
  } else if (!strcmp(field, "t")) {
    if (stage < 0 || stage > N) return fail(RLMPC_EINVAL, "stage out of range");
    *off = E::it_t(N, stage); *dim = E::NR;
  } else if (!strcmp(field, "rho_x0")) {
    *off = E::it_rx0(N); *dim = E::NX;
  } else if (!strcmp(field, "rho_u0")) {
    *off = E::it_ru0(N); *dim = E::NU;
  } else if (!strcmp(field, "meta")) {  // [0] = 1: (lam, t) hold the result of a QP solve (valid interior-point warm start)
    *off = E::it_meta(N); *dim = 1;
  } else {
    return fail(RLMPC_EINVAL, std::string("unknown field ") + field);
  }
  return 0;
}

int field_offset(rlmpc_handle* h, const char* field, int stage, int* off, int* dim) {
  int rc = fail(RLMPC_EINVAL, "unknown model");
  DISPATCH_MODEL(h, rc = field_offset_t<M>(h, field, stage, off, dim));
  return rc;
}

int refresh_cost_table(rlmpc_handle* h, int B) {
  DISPATCH_MODEL(h, (k_cost_table<M><<<h->th_per_sample ? (B + 127) / 128 : 1, h->th_per_sample ? 128 : 32>>>(
                        h->pd, h->th, h->ct, h->th_per_sample, B)));
  h->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

}  // namespace

extern "C" {

const char* rlmpc_last_error(void) { return g_err.c_str(); }

int rlmpc_create(const rlmpc_problem_desc* d, int max_batch, int device, rlmpc_handle** out) {
  if (!d || !out || max_batch <= 0) return fail(RLMPC_EINVAL, "bad arguments");
  if (d->N < 1 || d->N > RLMPC_MAXN) return fail(RLMPC_EINVAL, "N out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(RLMPC_ENODEV, "no CUDA device: rlmpc_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(RLMPC_EINVAL, "bad device index");
  rlmpc_handle* h = new (std::nothrow) rlmpc_handle();
  if (!h) return fail(RLMPC_ENOMEM, "out of host memory");
  h->model = d->model;
  h->device = device;
  h->max_batch = max_batch;
  h->bs = ((size_t)max_batch + 127) / 128 * 128;
  if (d->model == RLMPC_MODEL_CHAIN_MASS) {
    // chain of masses: dense nx = 9 / 21 / 27 stage blocks on the warp-cooperative engine; n_mass = model_const[1]
    ProblemData& pd = h->pd;
    memset(&pd, 0, sizeof(pd));
    pd.N = d->N;
    pd.mode = MODE_V; pd.max_sqp = 1; pd.max_ipm = 50; pd.warm_ipm = 1; pd.param_cost = 0;
    pd.tol = 1e-6; pd.tau = 1e-8; pd.mu0 = 1.0; pd.sigma_min = 0.05; pd.sigma0 = 0.3; pd.as_steps = 20; pd.comp_accept = 0.2; pd.step_length = 1.0;
    memcpy(pd.scale, d->scale, sizeof(double) * (d->N + 1));
    memcpy(pd.lbu, d->lbu, sizeof(pd.lbu)); memcpy(pd.ubu, d->ubu, sizeof(pd.ubu));
    memcpy(pd.mc, d->model_const, sizeof(pd.mc));
    h->variant = -1;
    cudaError_t e = cudaSetDevice(device);
    std::string err;
    int rc = 0;
    if (e != cudaSuccess) { rc = RLMPC_ECUDA; err = cudaGetErrorString(e); }
    if (!rc) rc = chain_create((int)(d->model_const[1] + 0.5), d->N, max_batch, &h->chain, err);
    if (!rc) {
      chain_dims(h->chain, &h->nx, &h->nu, &h->nth, &h->it_size);
      h->npm = h->nth; h->nr = 2 * h->nu; h->ws_size = 0; h->ct_size = 0;
      const size_t nio_in = (size_t)max_batch * (h->nx + h->nu);
      const size_t nio_out = (size_t)max_batch * (h->nu + 1 + 4 + (size_t)h->nth * (1 + h->nu));
      if (e == cudaSuccess) e = cudaMalloc(&h->d_in, sizeof(double) * nio_in);
      if (e == cudaSuccess) e = cudaMalloc(&h->d_out, sizeof(double) * nio_out);
      if (e == cudaSuccess) e = cudaMalloc(&h->d_status, sizeof(int) * max_batch);
      if (e == cudaSuccess) e = cudaMallocHost(&h->h_in, sizeof(double) * nio_in);
      if (e == cudaSuccess) e = cudaMallocHost(&h->h_out, sizeof(double) * nio_out);
      if (e == cudaSuccess) e = cudaMallocHost(&h->h_status, sizeof(int) * max_batch);
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
      if (e != cudaSuccess) { rc = (e == cudaErrorMemoryAllocation) ? RLMPC_ENOMEM : RLMPC_ECUDA; err = cudaGetErrorString(e); }
    }
    if (rc) {
      rlmpc_destroy(h);
      return fail(rc, err);
    }
    h->split = 1;
    *out = h;
    return 0;
  }
  switch (d->model) {
    case RLMPC_MODEL_CARTPOLE: {
      // state bounds present?  -> the instantiation that carries the 2*nx extra rows per stage
      bool bx = false;
      for (int i = 0; i < 4; ++i)
        bx = bx || d->lbx[i] > -BIG || d->ubx[i] < BIG || d->lbx_e[i] > -BIG || d->ubx_e[i] < BIG;
      // model_const[2] != 0: g is the fourth model parameter (theta = [M, m, l, g | W ...], 84 entries)
      const bool gfree = d->model_const[2] > 0.5;
      h->variant = gfree ? (bx ? VAR_CARTPOLE_BX_G : VAR_CARTPOLE_G) : (bx ? VAR_CARTPOLE_BX : VAR_CARTPOLE);
      break;
    }
    case RLMPC_MODEL_LINEAR_SYSTEM:
      h->variant = VAR_LINEAR;
      break;
    case RLMPC_MODEL_EVAPORATION:
      h->variant = VAR_EVAPORATION;
      break;
    default:
      delete h;
      return fail(RLMPC_EINVAL, "unknown model");
  }
  DISPATCH_MODEL(h, {
    using E = Engine<M>;
    h->nx = E::NX; h->nu = E::NU; h->nth = M::NTH; h->npm = E::NPM; h->nr = E::NR;
    h->it_size = E::it_size(d->N); h->ws_size = E::ws_size(d->N); h->ct_size = E::CT_SIZE;
  });
  {
    cudaError_t ea = cudaSetDevice(device);
    DISPATCH_MODEL(h, {
      if (ea == cudaSuccess)
        ea = cudaFuncSetAttribute(k_qp2<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qp2_smem<M>());
    });
    if (ea != cudaSuccess) {
      const std::string msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ea);
      delete h;
      return fail(RLMPC_ECUDA, msg);
    }
  }
  ProblemData& pd = h->pd;
  memset(&pd, 0, sizeof(pd));
  pd.N = d->N;
  pd.mode = MODE_V; pd.max_sqp = 1; pd.max_ipm = 50; pd.warm_ipm = 1; pd.param_cost = 0;
  pd.tol = 1e-6; pd.tau = 1e-8; pd.mu0 = 1.0; pd.sigma_min = 0.05; pd.sigma0 = 0.3; pd.as_steps = 20; pd.condense = 0; pd.comp_accept = 0.2; pd.step_length = 1.0;
  memcpy(pd.scale, d->scale, sizeof(double) * (d->N + 1));
  memcpy(pd.lbu, d->lbu, sizeof(pd.lbu)); memcpy(pd.ubu, d->ubu, sizeof(pd.ubu));
  memcpy(pd.lbx, d->lbx, sizeof(pd.lbx)); memcpy(pd.ubx, d->ubx, sizeof(pd.ubx));
  memcpy(pd.lbx_e, d->lbx_e, sizeof(pd.lbx_e)); memcpy(pd.ubx_e, d->ubx_e, sizeof(pd.ubx_e));
  memcpy(pd.mc, d->model_const, sizeof(pd.mc));
  memcpy(pd.zl, d->zl, sizeof(pd.zl)); memcpy(pd.zu, d->zu, sizeof(pd.zu));
  memcpy(pd.lg, d->lg, sizeof(pd.lg)); memcpy(pd.ug, d->ug, sizeof(pd.ug));
  cudaError_t e = cudaSetDevice(device);
  const size_t nio_in = (size_t)max_batch * (h->nx + h->nu);
  const size_t nio_out = (size_t)max_batch * (h->nu + 1 + 4 + (size_t)h->nth * (1 + h->nu));
  const size_t n_it = sizeof(double) * h->it_size * h->bs, n_ws = sizeof(double) * h->ws_size * h->bs;
  if (e == cudaSuccess) e = cudaMalloc(&h->it, n_it);
  if (e == cudaSuccess) e = cudaMalloc(&h->ws, n_ws);
  if (e == cudaSuccess) e = cudaMalloc(&h->it2, n_it);
  if (e == cudaSuccess) e = cudaMalloc(&h->ws2, n_ws);
  DISPATCH_MODEL(h, e = alloc_condensed<M>(h, d->N, e));
  DISPATCH_MODEL(h, e = setup_coop<M>(h, d->N, e));
  if (e == cudaSuccess) e = cudaMalloc(&h->th, sizeof(double) * h->nth * h->bs);
  if (e == cudaSuccess) e = cudaMalloc(&h->ct, sizeof(double) * h->ct_size * h->bs);
  if (e == cudaSuccess) e = cudaMalloc(&h->th_stage, sizeof(double) * h->nth * (size_t)max_batch);
  if (e == cudaSuccess) e = cudaMalloc(&h->cost, sizeof(double) * h->bs);
  if (e == cudaSuccess) e = cudaMalloc(&h->work, sizeof(int) * h->bs);
  if (e == cudaSuccess) e = cudaMalloc(&h->status, sizeof(int) * h->bs);
  if (e == cudaSuccess) e = cudaMalloc(&h->hard, sizeof(int) * h->bs);
  if (e == cudaSuccess) e = cudaMalloc(&h->ishard, sizeof(int) * h->bs);
  if (e == cudaSuccess) e = cudaMemset(h->ishard, 0, sizeof(int) * h->bs);
  if (e == cudaSuccess) e = cudaMalloc(&h->ipm_state, sizeof(double) * 8 * h->bs);
  if (e == cudaSuccess) e = cudaMemset(h->ipm_state, 0, sizeof(double) * 8 * h->bs);
  if (e == cudaSuccess) {  // side stream of the "overlap" option: highest priority, so that the few queue blocks are placed first
    int prio_lo = 0, prio_hi = 0;
    e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, prio_hi);
  }
  for (int i = 0; i < MAX_SPLIT - 1 && e == cudaSuccess; ++i) {
    int prio_lo = 0, prio_hi = 0;  // later parts first: their blocks fill in wherever an earlier part leaves room
    e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->part_stream[i], cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_part[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&h->counters, sizeof(int) * NCNT * MAX_SPLIT);  // one set per part of a split batch
  if (e == cudaSuccess) e = cudaMemset(h->counters, 0, sizeof(int) * NCNT * MAX_SPLIT);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_in, sizeof(double) * nio_in);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_out, sizeof(double) * nio_out);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_status, sizeof(int) * max_batch);
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_in, sizeof(double) * nio_in);
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_out, sizeof(double) * nio_out);
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_status, sizeof(int) * max_batch);
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_counters, sizeof(int) * 4);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  for (int i = 0; i < rlmpc_handle::NEV && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
  if (e == cudaSuccess) e = cudaMemset(h->it, 0, n_it);
  if (e == cudaSuccess) e = cudaMemset(h->ws, 0, n_ws);
  if (e == cudaSuccess) e = cudaMemset(h->it2, 0, n_it);
  if (e == cudaSuccess) e = cudaMemset(h->ws2, 0, n_ws);
  if (e == cudaSuccess) e = cudaMemset(h->th, 0, sizeof(double) * h->nth * h->bs);
  if (e == cudaSuccess) e = cudaMemset(h->ct, 0, sizeof(double) * h->ct_size * h->bs);
  if (e == cudaSuccess) e = cudaMemset(h->status, 0, sizeof(int) * h->bs);
  if (e == cudaSuccess) e = cudaMemset(h->cost, 0, sizeof(double) * h->bs);
  if (e != cudaSuccess) {
    std::string msg = std::string("allocation failed: ") + cudaGetErrorString(e);
    rlmpc_destroy(h);
    return fail(e == cudaErrorMemoryAllocation ? RLMPC_ENOMEM : RLMPC_ECUDA, msg);
  }
  *out = h;
  return 0;
}

void rlmpc_destroy(rlmpc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->chain) chain_destroy(h->chain);
  cudaFree(h->it); cudaFree(h->ws); cudaFree(h->it2); cudaFree(h->ws2); cudaFree(h->itb); cudaFree(h->wsb); cudaFree(h->th); cudaFree(h->ct);
  cudaFree(h->th_stage); cudaFree(h->cost); cudaFree(h->work); cudaFree(h->status); cudaFree(h->hard);
  cudaFree(h->counters); cudaFree(h->ishard); cudaFree(h->ipm_state);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  for (int i = 0; i < MAX_SPLIT - 1; ++i) {
    if (h->part_stream[i]) cudaStreamDestroy(h->part_stream[i]);
    if (h->ev_part[i]) cudaEventDestroy(h->ev_part[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaFree(h->d_in); cudaFree(h->d_out); cudaFree(h->d_status);
  cudaFreeHost(h->h_in); cudaFreeHost(h->h_out); cudaFreeHost(h->h_status); cudaFreeHost(h->h_counters);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->ev_last) cudaEventDestroy(h->ev_last);
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  for (int i = 0; i < rlmpc_handle::NEV; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  delete h;
}

int rlmpc_dims(const rlmpc_handle* h, int* nx, int* nu, int* ntheta, int* ngrad, int* iterate_size) {
  if (!h) return fail(RLMPC_EINVAL, "null handle");
  if (nx) *nx = h->nx;
  if (nu) *nu = h->nu;
  if (ntheta) *ntheta = h->nth;
  if (ngrad) *ngrad = h->ng();
  if (iterate_size) *iterate_size = h->it_size;
  return 0;
}

int rlmpc_nrows(const rlmpc_handle* h) { return h ? h->nr : RLMPC_EINVAL; }

int rlmpc_set_theta(rlmpc_handle* h, const double* theta_host, int per_sample, int B) {
  if (!h || !theta_host) return fail(RLMPC_EINVAL, "bad arguments");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->chain) {
    if (per_sample) return fail(RLMPC_EINVAL, "chain mass: one theta is shared by the batch (per-sample theta is not supported)");
    std::string err;
    CUDA_OK(cudaDeviceSynchronize());  // theta is read by kernels that may still be in flight
    const int rc = chain_set_theta(h->chain, h->pd, theta_host, false, nullptr, err);
    return rc ? fail(rc, err) : 0;
  }
  if (!per_sample) {
    CUDA_OK(cudaMemcpy(h->th_stage, theta_host, sizeof(double) * h->nth, cudaMemcpyHostToDevice));
    k_theta_transpose<<<1, TILE>>>(h->th_stage, h->th, TILE, h->nth, 1);
    h->launches++;
    CUDA_OK(cudaGetLastError());
    h->th_per_sample = 0;
    return refresh_cost_table(h, 0);
  }
  if (int r = check_batch(h, B)) return r;
  CUDA_OK(cudaMemcpy(h->th_stage, theta_host, sizeof(double) * h->nth * (size_t)B, cudaMemcpyHostToDevice));
  k_theta_transpose<<<(B + 127) / 128, 128>>>(h->th_stage, h->th, B, h->nth, 0);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  h->th_per_sample = 1;
  return refresh_cost_table(h, B);
}

int rlmpc_set_theta_dev(rlmpc_handle* h, const double* theta_dev, int per_sample, int B, void* stream) {
  if (!h || !theta_dev) return fail(RLMPC_EINVAL, "bad arguments");
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (h->chain) {
    if (per_sample) return fail(RLMPC_EINVAL, "chain mass: one theta is shared by the batch (per-sample theta is not supported)");
    std::string err;
    const int rc = chain_set_theta(h->chain, h->pd, theta_dev, true, s, err);
    if (!rc) note_stream(h, s);
    return rc ? fail(rc, err) : 0;
  }
  if (per_sample) {
    if (int r = check_batch(h, B)) return r;
    if (B == 0) return 0;
  }
  const int n = per_sample ? B : TILE;
  k_theta_transpose<<<(n + 127) / 128, 128, 0, s>>>(theta_dev, h->th, n, h->nth, per_sample ? 0 : 1);
  h->th_per_sample = per_sample ? 1 : 0;
  DISPATCH_MODEL(h, (k_cost_table<M><<<(n + 127) / 128, 128, 0, s>>>(h->pd, h->th, h->ct, h->th_per_sample, B)));
  h->launches += 2;
  CUDA_OK(cudaGetLastError());
  note_stream(h, (cudaStream_t)stream);
  return 0;
}

int rlmpc_set_model_vector(rlmpc_handle* h, const char* name, const double* v_host, int n) {
  if (!h || !name || !v_host || n < 0) return fail(RLMPC_EINVAL, "bad arguments");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->chain && !strcmp(name, "x_ss")) {
    std::string err;
    const int rc = chain_set_xss(h->chain, h->pd, v_host, n, err);
    return rc ? fail(rc, err) : 0;
  }
  return fail(RLMPC_EINVAL, std::string("this model has no vector ") + name);
}

int rlmpc_set_cost_scaling(rlmpc_handle* h, const double* scale, int n) {
  if (!h || !scale || n != h->pd.N + 1) return fail(RLMPC_EINVAL, "scale must have N+1 entries");
  memcpy(h->pd.scale, scale, sizeof(double) * n);
  return 0;
}

int rlmpc_set_bounds(rlmpc_handle* h, const char* field, const double* v, int n) {
  if (!h || !field || !v || n < 0 || n > RLMPC_MAXD) return fail(RLMPC_EINVAL, "bad arguments");
  double* dst = nullptr;
  const bool is_x = field[0] && field[1] == 'b' && field[2] == 'x';
  if (!strcmp(field, "lbu")) dst = h->pd.lbu;
  else if (!strcmp(field, "ubu")) dst = h->pd.ubu;
  else if (!strcmp(field, "lbx")) dst = h->pd.lbx;
  else if (!strcmp(field, "ubx")) dst = h->pd.ubx;
  else if (!strcmp(field, "lbx_e")) dst = h->pd.lbx_e;
  else if (!strcmp(field, "ubx_e")) dst = h->pd.ubx_e;
  else if (!strcmp(field, "zl") || !strcmp(field, "zu")) dst = nullptr;
  else return fail(RLMPC_EINVAL, std::string("unknown bound field ") + field);
  if (!strcmp(field, "zl")) dst = h->pd.zl;
  if (!strcmp(field, "zu")) dst = h->pd.zu;
  if (h->chain && strcmp(field, "lbu") && strcmp(field, "ubu")) return fail(RLMPC_EINVAL, "chain mass has input bounds only");
  if (is_x && (h->variant == VAR_CARTPOLE || h->variant == VAR_CARTPOLE_G)) {
    for (int i = 0; i < n; ++i)
      if (v[i] > -BIG && v[i] < BIG)
        return fail(RLMPC_EINVAL, "this handle was created without state bounds; create it with finite lbx/ubx");
  }
  memcpy(dst, v, sizeof(double) * n);
  return 0;
}

int rlmpc_set_option(rlmpc_handle* h, const char* name, double value) {
  if (!h || !name) return fail(RLMPC_EINVAL, "bad arguments");
  if (!strcmp(name, "tol")) h->pd.tol = value;
  else if (!strcmp(name, "tau")) h->pd.tau = value;
  else if (!strcmp(name, "mu0")) h->pd.mu0 = value;
  else if (!strcmp(name, "sigma_min")) h->pd.sigma_min = value;
  else if (!strcmp(name, "sigma0")) h->pd.sigma0 = value;
  else if (!strcmp(name, "as_steps")) h->pd.as_steps = value;
  else if (!strcmp(name, "max_ipm")) h->pd.max_ipm = (int)value;
  else if (!strcmp(name, "warm_ipm")) h->pd.warm_ipm = (int)value;
  else if (!strcmp(name, "param_cost")) h->pd.param_cost = (int)value;
  else if (!strcmp(name, "sync_every")) h->sync_every = value < 1 ? 1 : (int)value;
  else if (!strcmp(name, "timing")) { h->timing = (int)value; if (h->chain) chain_set_timing(h->chain, h->timing); }
  else if (!strcmp(name, "overlap")) h->overlap = (int)value;
  else if (!strcmp(name, "ring")) h->ring = (int)value;
  else if (!strcmp(name, "ring_b")) h->ring_b = (int)value;
  else if (!strcmp(name, "coop")) h->coop = (int)value;
  else if (!strcmp(name, "comp_accept")) h->pd.comp_accept = value;
  else if (!strcmp(name, "step_length")) { if (!(value > 0.0 && value <= 1.0)) return fail(RLMPC_EINVAL, "step_length must be in (0, 1]"); h->pd.step_length = value; }
  else if (!strcmp(name, "split")) h->split = (int)value;
  else if (!strcmp(name, "graph")) h->use_graph = (int)value;
  else if (!strcmp(name, "condense")) h->condense = (int)value;
  else if (!strcmp(name, "inplace_queue")) h->inplace_queue = (int)value;
  else if (!strcmp(name, "ipm_passes")) h->ipm_passes = value < 0 ? 0 : (int)value;
  else if (!strcmp(name, "fuse_lin")) h->fuse_lin = (int)value;
  else return fail(RLMPC_EINVAL, std::string("unknown option ") + name);
  return 0;
}

int rlmpc_reset(rlmpc_handle* h, int B, const double* x0_dev, void* stream) {
  return rlmpc_reset_masked(h, B, x0_dev, nullptr, stream);
}

int rlmpc_reset_masked(rlmpc_handle* h, int B, const double* x0_dev, const int* mask_dev, void* stream) {
  if (int r = check_batch(h, B)) return r;
  if (B == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  if (h->chain) {
    std::string err;
    const int rc = chain_reset(h->chain, h->pd, B, x0_dev, mask_dev, (cudaStream_t)stream, err);
    if (!rc) note_stream(h, (cudaStream_t)stream);
    return rc ? fail(rc, err) : 0;
  }
  DISPATCH_MODEL(h, (k_reset<M><<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->pd.N, h->it, B, x0_dev, mask_dev)));
  h->launches++;
  CUDA_OK(cudaGetLastError());
  note_stream(h, (cudaStream_t)stream);
  return 0;
}

int rlmpc_get_iterate(rlmpc_handle* h, const char* field, int stage, int B, double* buf_dev, void* stream) {
  if (int r = check_batch(h, B)) return r;
  if (!field || !buf_dev) return fail(RLMPC_EINVAL, "bad arguments");
  if (h->chain) {
    std::string err;
    CUDA_OK(cudaSetDevice(h->device));
    const int rc = chain_field(h->chain, h->pd, field, stage, B, buf_dev, 0, (cudaStream_t)stream, nullptr, err);
    return rc ? fail(rc, err) : 0;
  }
  int off, dim;
  if (int r = field_offset(h, field, stage, &off, &dim)) return r;
  if (B == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  k_copy_field<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->it, h->it_size, B, off, dim, buf_dev, 0);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int rlmpc_put_iterate(rlmpc_handle* h, const char* field, int stage, int B, const double* buf_dev, void* stream) {
  if (int r = check_batch(h, B)) return r;
  if (!field || !buf_dev) return fail(RLMPC_EINVAL, "bad arguments");
  if (h->chain) {
    std::string err;
    CUDA_OK(cudaSetDevice(h->device));
    const int rc = chain_field(h->chain, h->pd, field, stage, B, const_cast<double*>(buf_dev), 1, (cudaStream_t)stream, nullptr, err);
    if (!rc) note_stream(h, (cudaStream_t)stream);
    return rc ? fail(rc, err) : 0;
  }
  int off, dim;
  if (int r = field_offset(h, field, stage, &off, &dim)) return r;
  if (B == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  k_copy_field<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->it, h->it_size, B, off, dim, const_cast<double*>(buf_dev), 1);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  note_stream(h, (cudaStream_t)stream);
  return 0;
}

int rlmpc_solve(rlmpc_handle* h, int mode, int max_sqp, int B, const double* x0_dev, const double* u0_dev,
                double* u0_out_dev, double* cost_out_dev, int* status_out_dev, void* stream) {
  const int rc = run_unit(h, mode, max_sqp, B, x0_dev, u0_dev, u0_out_dev, cost_out_dev, status_out_dev, nullptr, nullptr,
                          nullptr, 1, 0, (cudaStream_t)stream);
  if (!rc && B > 0) note_stream(h, (cudaStream_t)stream);
  return rc;
}

int rlmpc_sens(rlmpc_handle* h, int mode, int B, double* dL_dtheta_dev, double* dpi_dtheta_dev, double* cost_out_dev,
               double* res_out_dev, int* status_out_dev, void* stream) {
  const int rc = run_unit(h, mode, h ? h->pd.max_sqp : 1, B, nullptr, nullptr, nullptr, cost_out_dev, status_out_dev,
                          dL_dtheta_dev, dpi_dtheta_dev, res_out_dev, 0, 1, (cudaStream_t)stream);
  if (!rc && B > 0) note_stream(h, (cudaStream_t)stream);
  return rc;
}

int rlmpc_solve_sens(rlmpc_handle* h, int mode, int max_sqp, int B, const double* x0_dev, const double* u0_dev,
                     double* u0_out_dev, double* cost_out_dev, int* status_out_dev, double* dL_dtheta_dev,
                     double* dpi_dtheta_dev, double* res_out_dev, void* stream) {
  const int rc = run_unit(h, mode, max_sqp, B, x0_dev, u0_dev, u0_out_dev, cost_out_dev, status_out_dev, dL_dtheta_dev,
                          dpi_dtheta_dev, res_out_dev, 1, 1, (cudaStream_t)stream);
  if (!rc && B > 0) note_stream(h, (cudaStream_t)stream);
  return rc;
}

int rlmpc_solve_sens_host(rlmpc_handle* h, int mode, int max_sqp, int B, const double* x0_host, const double* u0_host,
                          double* u0_out_host, double* cost_out_host, int* status_out_host, double* dL_dtheta_host,
                          double* dpi_dtheta_host, double* res_out_host) {
  if (int r = check_batch(h, B)) return r;
  if (B == 0) return 0;
  if (!x0_host) return fail(RLMPC_EINVAL, "x0 is required");
  if (mode == RLMPC_MODE_Q && !u0_host) return fail(RLMPC_EINVAL, "u0 is required in Q-mode");
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = h->own_stream;
  // This entry point runs on the handle's own streams and is synchronous; order it after whatever the caller
  // has queued for this handle through the stream-ordered entry points (rlmpc_reset, rlmpc_solve, ...): those
  // leave an event behind (note_stream), so no device-wide synchronisation is needed.
  if (h->ev_last_set) CUDA_OK(cudaStreamWaitEvent(s, h->ev_last, 0));
  const size_t nB = (size_t)B, nx = h->nx, nu = h->nu, nth = h->ng();
  // Page-locked caller buffers (cudaHostAlloc / cudaHostRegister, e.g. torch pinned tensors) are used
  // directly as DMA source / destination; pageable ones go through the handle's pinned staging area.
  auto pinned = [](const void* p) {
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return at.type == cudaMemoryTypeHost;
  };
  const bool in_pinned = pinned(x0_host) && pinned(u0_host);
  const bool all_pinned = in_pinned && pinned(u0_out_host) && pinned(cost_out_host) && pinned(res_out_host) &&
                          pinned(dL_dtheta_host) && pinned(dpi_dtheta_host) && pinned(status_out_host);
  // the per-part pipelines address the queue through per-part counters; only the warp-per-sample queue kernel works in
  // place on the samples' own records -- the thread-per-sample fallback shares compact buffers by queue slot, so parts
  // must not run concurrently there (option coop = 0, or a horizon too long for shared memory)
  bool coop_ok = false;
  if (!h->chain) DISPATCH_MODEL(h, coop_ok = coop_model<M>() && h->coop && h->coop_grid > 0);
  if (all_pinned && max_sqp == 1 && h->split > 1 && B >= 4096 && !h->overlap && coop_ok) {
    // Page-locked buffers, RTI: the parts of the batch (option "split") are complete pipelines of their own --
    // copy in, kernel chain, copy out -- on separate streams, so that the copies of one part overlap with the
    // kernels of another.
    const int parts = h->split < MAX_SPLIT ? h->split : MAX_SPLIT;
    const int per = ((B + parts - 1) / parts + TILE - 1) / TILE * TILE;
    double* d_x0 = h->d_in;
    double* d_u0 = u0_host ? h->d_in + nB * nx : nullptr;
    double* d_u0o = h->d_out;
    double* d_cost = d_u0o + nB * nu;
    double* d_res = d_cost + nB;
    double* d_dL = d_res + nB * 4;
    double* d_dpi = d_dL + nB * nth;
    int rc = 0, used = 0;
    for (int p = 0; p < parts && rc == 0; ++p) {
      const size_t b0 = (size_t)p * per;
      if (b0 >= nB) break;
      const size_t nb = nB - b0 < (size_t)per ? nB - b0 : (size_t)per;
      cudaStream_t sp = (p == 0) ? s : h->part_stream[p - 1];
      if (p > 0 && h->ev_last_set) CUDA_OK(cudaStreamWaitEvent(sp, h->ev_last, 0));
      CUDA_OK(cudaMemcpyAsync(d_x0 + b0 * nx, x0_host + b0 * nx, sizeof(double) * nb * nx, cudaMemcpyHostToDevice, sp));
      if (u0_host) CUDA_OK(cudaMemcpyAsync(d_u0 + b0 * nu, u0_host + b0 * nu, sizeof(double) * nb * nu, cudaMemcpyHostToDevice, sp));
      CUDA_OK(cudaMemsetAsync(d_dL + b0 * nth, 0, sizeof(double) * nb * nth, sp));
      CUDA_OK(cudaMemsetAsync(d_dpi + b0 * nth * nu, 0, sizeof(double) * nb * nth * nu, sp));
      rc = run_unit(h, mode, max_sqp, (int)nb, d_x0, d_u0, d_u0o, d_cost, h->d_status, d_dL, d_dpi, d_res, 1, 1, sp, (int)b0, p);
      if (rc) break;
      if (u0_out_host) CUDA_OK(cudaMemcpyAsync(u0_out_host + b0 * nu, d_u0o + b0 * nu, sizeof(double) * nb * nu, cudaMemcpyDeviceToHost, sp));
      if (cost_out_host) CUDA_OK(cudaMemcpyAsync(cost_out_host + b0, d_cost + b0, sizeof(double) * nb, cudaMemcpyDeviceToHost, sp));
      if (res_out_host) CUDA_OK(cudaMemcpyAsync(res_out_host + b0 * 4, d_res + b0 * 4, sizeof(double) * nb * 4, cudaMemcpyDeviceToHost, sp));
      if (dL_dtheta_host) CUDA_OK(cudaMemcpyAsync(dL_dtheta_host + b0 * nth, d_dL + b0 * nth, sizeof(double) * nb * nth, cudaMemcpyDeviceToHost, sp));
      if (dpi_dtheta_host)
        CUDA_OK(cudaMemcpyAsync(dpi_dtheta_host + b0 * nth * nu, d_dpi + b0 * nth * nu, sizeof(double) * nb * nth * nu, cudaMemcpyDeviceToHost, sp));
      if (status_out_host) CUDA_OK(cudaMemcpyAsync(status_out_host + b0, h->d_status + b0, sizeof(int) * nb, cudaMemcpyDeviceToHost, sp));
      used = p + 1;
    }
    for (int p = 0; p < used; ++p) cudaStreamSynchronize(p == 0 ? s : h->part_stream[p - 1]);
    if (rc) return rc;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  double* d_x0 = h->d_in;
  double* d_u0 = u0_host ? h->d_in + nB * nx : nullptr;
  if (in_pinned) {
    CUDA_OK(cudaMemcpyAsync(d_x0, x0_host, sizeof(double) * nB * nx, cudaMemcpyHostToDevice, s));
    if (u0_host) CUDA_OK(cudaMemcpyAsync(d_u0, u0_host, sizeof(double) * nB * nu, cudaMemcpyHostToDevice, s));
  } else {
    memcpy(h->h_in, x0_host, sizeof(double) * nB * nx);
    if (u0_host) memcpy(h->h_in + nB * nx, u0_host, sizeof(double) * nB * nu);
    CUDA_OK(cudaMemcpyAsync(h->d_in, h->h_in, sizeof(double) * nB * (nx + (u0_host ? nu : 0)), cudaMemcpyHostToDevice, s));
  }
  double* d_u0o = h->d_out;
  double* d_cost = d_u0o + nB * nu;
  double* d_res = d_cost + nB;
  double* d_dL = d_res + nB * 4;
  double* d_dpi = d_dL + nB * nth;
  const size_t n_out = nB * (nu + 1 + 4 + nth * (1 + nu));
  CUDA_OK(cudaMemsetAsync(d_dL, 0, sizeof(double) * nB * nth * (1 + nu), s));
  if (int r = run_unit(h, mode, max_sqp, B, d_x0, d_u0, d_u0o, d_cost, h->d_status, d_dL, d_dpi, d_res, 1, 1, s)) return r;
  const bool out_pinned = pinned(u0_out_host) && pinned(cost_out_host) && pinned(res_out_host) && pinned(dL_dtheta_host) &&
                          pinned(dpi_dtheta_host) && pinned(status_out_host);
  if (out_pinned) {
    if (u0_out_host) CUDA_OK(cudaMemcpyAsync(u0_out_host, d_u0o, sizeof(double) * nB * nu, cudaMemcpyDeviceToHost, s));
    if (cost_out_host) CUDA_OK(cudaMemcpyAsync(cost_out_host, d_cost, sizeof(double) * nB, cudaMemcpyDeviceToHost, s));
    if (res_out_host) CUDA_OK(cudaMemcpyAsync(res_out_host, d_res, sizeof(double) * nB * 4, cudaMemcpyDeviceToHost, s));
    if (dL_dtheta_host) CUDA_OK(cudaMemcpyAsync(dL_dtheta_host, d_dL, sizeof(double) * nB * nth, cudaMemcpyDeviceToHost, s));
    if (dpi_dtheta_host)
      CUDA_OK(cudaMemcpyAsync(dpi_dtheta_host, d_dpi, sizeof(double) * nB * nth * nu, cudaMemcpyDeviceToHost, s));
    if (status_out_host) CUDA_OK(cudaMemcpyAsync(status_out_host, h->d_status, sizeof(int) * nB, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
  }
  CUDA_OK(cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int) * nB, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  const double* o = h->h_out;
  if (u0_out_host) memcpy(u0_out_host, o, sizeof(double) * nB * nu);
  if (cost_out_host) memcpy(cost_out_host, o + nB * nu, sizeof(double) * nB);
  if (res_out_host) memcpy(res_out_host, o + nB * (nu + 1), sizeof(double) * nB * 4);
  if (dL_dtheta_host) memcpy(dL_dtheta_host, o + nB * (nu + 5), sizeof(double) * nB * nth);
  if (dpi_dtheta_host) memcpy(dpi_dtheta_host, o + nB * (nu + 5 + nth), sizeof(double) * nB * nth * nu);
  if (status_out_host) memcpy(status_out_host, h->h_status, sizeof(int) * nB);
  return 0;
}

int rlmpc_td_grad(rlmpc_handle* h, int B, int ncols, const double* td_dev, const double* dQ_dtheta_dev,
                  const int* status_dev, double* acc_out_dev, void* stream) {
  if (int r = check_batch(h, B)) return r;
  if (!td_dev || !dQ_dtheta_dev || !acc_out_dev || ncols < 1 || ncols > 4096) return fail(RLMPC_EINVAL, "bad arguments");
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_OK(cudaMemsetAsync(acc_out_dev, 0, sizeof(double) * (ncols + 2), s));
  if (B == 0) return 0;
  const int threads = 256, nwarp = threads / 32;
  int grid = (B + nwarp - 1) / nwarp;
  if (grid > 148 * 4) grid = 148 * 4;
  k_td_grad<<<grid, threads, sizeof(double) * (ncols + 2), s>>>(B, ncols, td_dev, dQ_dtheta_dev, status_dev, acc_out_dev);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

size_t rlmpc_store_bytes(const rlmpc_handle* h, int capacity) {
  if (!h || capacity <= 0) return 0;
  if (h->chain) return chain_store_bytes(h->chain, capacity);
  return sizeof(double) * (size_t)h->it_size * (((size_t)capacity + TILE - 1) / TILE * TILE);
}

int rlmpc_store_copy(rlmpc_handle* h, int B, const int* idx_dev, double* store_dev, int capacity, int to_store,
                     void* stream) {
  if (int r = check_batch(h, B)) return r;
  if (!idx_dev || !store_dev || capacity <= 0) return fail(RLMPC_EINVAL, "bad arguments");
  if (B == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  if (h->chain) {
    std::string err;
    const int rc = chain_store_copy(h->chain, B, idx_dev, store_dev, capacity, to_store, (cudaStream_t)stream, err);
    if (!rc) note_stream(h, (cudaStream_t)stream);
    return rc ? fail(rc, err) : 0;
  }
  const int wpb = 8, chunks = (h->it_size + 63) / 64;
  k_store_copy<<<dim3((B + 31) / 32, (chunks + wpb - 1) / wpb), 32 * wpb, 0, (cudaStream_t)stream>>>(
      h->it, store_dev, h->it_size, B, idx_dev, capacity, to_store);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  note_stream(h, (cudaStream_t)stream);
  return 0;
}

int rlmpc_cartpole_env_step(const double* par_dev, int B, double* state_dev, const double* action_dev,
                            double* reward_dev, int* terminated_dev, int* truncated_dev, int* steps_dev, void* stream) {
  if (!par_dev || !state_dev || !action_dev || !reward_dev || !terminated_dev || !truncated_dev || !steps_dev || B < 0)
    return fail(RLMPC_EINVAL, "bad arguments");
  if (B == 0) return 0;
  k_cartpole_env_step<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(par_dev, B, state_dev, action_dev, reward_dev,
                                                                       terminated_dev, truncated_dev, steps_dev);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int rlmpc_fp64_tensor_peak(int device, double* tflops_out) {
  if (!tflops_out) return fail(RLMPC_EINVAL, "bad arguments");
  CUDA_OK(cudaSetDevice(device));
  int sms = 0;
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  double* d = nullptr;
  CUDA_OK(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  const int blocks = sms * 4, threads = 256, iters = 1 << 13;
  double best = 0.0;
  for (int r = 0; r < 6; ++r) {
    CUDA_OK(cudaEventRecord(e0));
    k_dmma_peak<<<blocks, threads>>>(d, iters);
    CUDA_OK(cudaEventRecord(e1));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 512.0 * 8.0 * (double)iters * blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  CUDA_OK(cudaGetLastError());
  *tflops_out = best;
  return 0;
}

int rlmpc_fp64_peak(int device, double* tflops_out) {
  if (!tflops_out) return fail(RLMPC_EINVAL, "bad arguments");
  CUDA_OK(cudaSetDevice(device));
  int sms = 0;
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  double* d = nullptr;
  CUDA_OK(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256, iters = 1 << 15;
  double best = 0.0;
  for (int r = 0; r < 6; ++r) {
    CUDA_OK(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(d, iters);
    CUDA_OK(cudaEventRecord(e1));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  CUDA_OK(cudaGetLastError());
  *tflops_out = best;
  return 0;
}

long long rlmpc_launch_count(const rlmpc_handle* h) { return h ? h->launches + (h->chain ? chain_launches(h->chain) : 0) : 0; }

int rlmpc_get_timings(rlmpc_handle* h, double* ms_out, int n) {
  if (!h || !ms_out || n < 6) return fail(RLMPC_EINVAL, "bad arguments");
  if (!h->timing) return fail(RLMPC_EINVAL, "option \"timing\" is off");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->chain) {
    std::string err;
    const int rc = chain_timings(h->chain, ms_out, n, err);
    return rc ? fail(rc, err) : 0;
  }
  // [lin | qp1 | qp2 (from the end of qp1, possibly on the side stream) | sens_stage | sens_sweep | tail]
  const int from[6] = {0, 1, 2, 2, 4, 5}, to[6] = {1, 2, 3, 4, 5, 6};
  for (int i = 0; i < 6; ++i) {
    ms_out[i] = 0.0;
    int f = from[i];
    if (i == 3 && !h->overlap && h->ev_set[3]) f = 3;  // serial: sens_stage starts when qp2 ends
    if (i == 3 && !h->ev_set[2]) f = 3;                // sens-only call
    if (h->ev_set[f] && h->ev_set[to[i]]) {
      CUDA_OK(cudaEventSynchronize(h->ev[to[i]]));
      CUDA_OK(cudaEventSynchronize(h->ev[f]));
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, h->ev[f], h->ev[to[i]]));
      ms_out[i] = ms;
    }
  }
  if (n >= 8) {  // queue statistics of the last SQP round: length, interior-point iterations (warp-per-sample kernel only)
    CUDA_OK(cudaDeviceSynchronize());
    int c[NCNT * MAX_SPLIT] = {};
    CUDA_OK(cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
    const int parts = h->split > 1 ? (h->split < MAX_SPLIT ? h->split : MAX_SPLIT) : 1;  // the parts of a split batch count separately
    ms_out[6] = ms_out[7] = 0.0;
    for (int p = 0; p < parts; ++p) {
      ms_out[6] += c[NCNT * p] + c[NCNT * p + 4];
      ms_out[7] += c[NCNT * p + 3];
    }
  }
  return 0;
}

}  // extern "C"
