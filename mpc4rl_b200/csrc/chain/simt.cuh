// Target abstraction of the warp-cooperative chain-mass kernels.
//
// The kernel bodies in chain_engine.cuh are written once, as functions of (lane, shared-memory pointer, ...), in
// terms of the few SIMT primitives below.  Compiled by nvcc they are the product (rlmpc_chain.cu).  The same bodies
// can be compiled by a host compiler against an emulation of these primitives (cooperative fibers, one per lane;
// oracle/cpu_port (test infrastructure)) to debug the maths on a box without a GPU.  Nothing of the
// emulation lives in this tree: a host translation unit must define the CH_* / W* macros before including this.
#pragma once
#include "../common.cuh"

#if defined(__CUDACC__)

#include <cuda_runtime.h>
#include <cstdint>

#define CH_DEV __device__ __forceinline__
#define WSYNC() __syncwarp()
#define BSYNC() __syncthreads()

namespace rlmpc {
CH_DEV double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
CH_DEV double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
CH_DEV double wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
CH_DEV bool wany(bool p) { return __any_sync(0xffffffffu, p); }
CH_DEV int wbcast_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
CH_DEV int atomic_next(int* counter) { return atomicAdd(counter, 1); }
// FP64 tensor-core tile product (DMMA): C (8 x 8) += A (8 x 4) B (4 x 8), fragments in registers.
//   lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and C[l / 4][2 (l % 4)], C[l / 4][2 (l % 4) + 1]
CH_DEV void dmma_8x8x4(double a, double b, double& c0, double& c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- TMA (bulk asynchronous copy) feed of per-stage records: global -> shared, completion on an mbarrier ----
// One elected lane arms the barrier with the byte count and issues cp.async.bulk; every lane waits on the
// barrier's phase parity.  Two slots = the stage being worked on and the next one in flight.
struct StageFeed {
  double* buf[2];
  uint64_t* bar;  // two mbarriers in shared memory
  unsigned phase;  // bit s = parity the next wait on slot s expects
  CH_DEV void init(double* b0, double* b1, uint64_t* bars, int lane) {
    buf[0] = b0; buf[1] = b1; bar = bars; phase = 0;
    if (lane == 0) {
      const unsigned a0 = (unsigned)__cvta_generic_to_shared(bar), a1 = (unsigned)__cvta_generic_to_shared(bar + 1);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a0));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  // make generic-proxy global writes of this warp visible to later bulk copies (call by all lanes)
  CH_DEV void publish(int lane) {
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncwarp();
  }
  // up to three segments (n in doubles, even; 16-byte aligned sources; destinations offset inside the slot)
  CH_DEV void issue(int slot, int lane, const double* s0, int o0, int n0, const double* s1 = nullptr, int o1 = 0, int n1 = 0,
                    const double* s2 = nullptr, int o2 = 0, int n2 = 0) {
    if (lane == 0) {
      const unsigned b = (unsigned)__cvta_generic_to_shared(bar + slot);
      const unsigned bytes = 8u * (unsigned)(n0 + n1 + n2);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
      const unsigned d0 = (unsigned)__cvta_generic_to_shared(buf[slot] + o0);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d0), "l"(s0),
                   "r"(8u * (unsigned)n0), "r"(b)
                   : "memory");
      if (n1 > 0) {
        const unsigned d1 = (unsigned)__cvta_generic_to_shared(buf[slot] + o1);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d1), "l"(s1),
                     "r"(8u * (unsigned)n1), "r"(b)
                     : "memory");
      }
      if (n2 > 0) {
        const unsigned d2 = (unsigned)__cvta_generic_to_shared(buf[slot] + o2);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d2), "l"(s2),
                     "r"(8u * (unsigned)n2), "r"(b)
                     : "memory");
      }
    }
  }
  CH_DEV const double* wait(int slot) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar + slot);
    const unsigned par = (phase >> slot) & 1u;
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(done)
          : "r"(b), "r"(par)
          : "memory");
    }
    phase ^= (1u << slot);
    return buf[slot];
  }
};
}  // namespace rlmpc

#else  // host emulation: the includer provides CH_DEV, WSYNC, BSYNC, wsum, wmax, wmin, wany, atomic_next, StageFeed
#ifndef CH_DEV
#error "host builds must include a SIMT emulation header (simt_host.h of the test infrastructure) before chain/simt.cuh"
#endif
#endif
