// mcd_large.cuh — global-memory pipeline for slabs that do not fit in shared memory.
#pragma once
#include "mcd_common.cuh"
#include "mcd_slab.cuh"
#include <string>

namespace mcd {

struct LargeEnv {
  cudaStream_t stream;
  int sm_count;
  int smem_optin;
  unsigned* flags;
  long long workspace_bytes;
  long long* launches;
  void** work;
  size_t* work_cap;
  const int* d_chain_inds;
};

template <typename T>
static int run_large(LargeEnv& env, const T* dx, long long params, const SplitGeom& g, int nsteps,
                     const Step* steps, int combine, int method, int maxlag, int relative, int ess_nan,
                     double mcse_p, int cps, int nsuper, T* d_ess, T* d_rhat, void* d_arr, std::string& msg) {
  msg = "large-slab pipeline not built yet";
  return -4;
}

}  // namespace mcd
