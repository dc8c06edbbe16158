// mcd_big_api.cuh — what the call driver needs from the big-slab estimator kernel's translation unit (mcd_big.cu).
#pragma once
#include "mcd_common.cuh"

namespace mcd {

template <typename T> struct BigArgs {
  const T* x;
  long long params;
  SplitGeom g;
  int proxy;            // 0 = x (mean), 1 = (x - mean)^2 (std), 2 = x <= median(x)
  int method;           // MCD_AUTOCOV_DIRECT (0) or MCD_AUTOCOV_BDA (2)
  int maxlag, relative, ess_nan, want_ess;
  T rel_ess_max;
  T* ess_out;
  T* rhat_out;
  int off_aux, off_part, off_small;   // shared-memory byte offsets (filled by big_launch)
};

// Shared memory the kernel needs for this geometry (0 = the slab does not fit / is not eligible).
template <typename T> size_t big_smem_bytes(const SplitGeom& g, int maxlag, int* off_aux, int* off_part, int* off_small);

// Persistent grid of `grid` CTAs (one per SM).
template <typename T> cudaError_t big_launch(BigArgs<T> a, unsigned grid, cudaStream_t stream);

}  // namespace mcd
