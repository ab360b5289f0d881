// Internal interface between the C ABI (rlmpc_b200.cu) and the chain-mass kernels (rlmpc_chain.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../common.cuh"

namespace rlmpc {

struct ChainBackend;  // device buffers of one handle

struct ChainCall {
  int mode, max_sqp, B, do_solve, do_sens;
  const double* x0; const double* u0;
  double* u0_out; double* cost_out; int* status_out; double* dL; double* dpi; double* res_out;
};

// n_mass in {3, 5, 6}.  All functions return 0 or a negative RLMPC_E* code and put a message into err.
int chain_create(int n_mass, int N, int max_batch, ChainBackend** out, std::string& err);
void chain_destroy(ChainBackend* cb);
void chain_dims(const ChainBackend* cb, int* nx, int* nu, int* ntheta, int* it_size);
int chain_set_theta(ChainBackend* cb, const ProblemData& pd, const double* theta, bool on_device, cudaStream_t s, std::string& err);
int chain_set_xss(ChainBackend* cb, const ProblemData& pd, const double* xss_host, int n, std::string& err);
int chain_reset(ChainBackend* cb, const ProblemData& pd, int B, const double* x0_dev, const int* mask_dev, cudaStream_t s, std::string& err);
// field in {x,u,pi,lam,t,rho_x0,rho_u0,meta}; dim_out (optional) receives the row width
int chain_field(ChainBackend* cb, const ProblemData& pd, const char* field, int stage, int B, double* buf_dev, int to_iterate,
                cudaStream_t s, int* dim_out, std::string& err);
int chain_run(ChainBackend* cb, const ProblemData& pd, const ChainCall& c, int sync_every, cudaStream_t s, std::string& err);
size_t chain_store_bytes(const ChainBackend* cb, int capacity);
int chain_store_copy(ChainBackend* cb, int B, const int* idx_dev, double* store_dev, int capacity, int to_store, cudaStream_t s,
                     std::string& err);
long long chain_launches(const ChainBackend* cb);
// device milliseconds of the phases of the last call: [lin, qp, sens_stage, sens, param]; queue statistics: ipm iterations
int chain_timings(ChainBackend* cb, double* ms_out, int n, std::string& err);
void chain_set_timing(ChainBackend* cb, int on);

}  // namespace rlmpc
