// mcd_rk2.cu — translation unit of the headline kernel (mcd_rk2.cuh) and its launcher.
#include "mcd_rk2.cuh"

namespace mcd {

template <typename T>
cudaError_t rk2_launch(const FastArgs<T>& a, unsigned grid, size_t extra_smem, cudaStream_t stream) {
  const size_t smem = (size_t)RK_SMEM_BYTES + extra_smem;
  const int mode = (a.rank_x ? 1 : 0) | (a.do_bulk ? 2 : 0) | (a.do_tail ? 4 : 0);
  const bool lng = a.niter > 32 * (RK_EPT - 1);
  void (*kern)(const FastArgs<T>) = nullptr;
  switch (mode) {
    case 7: kern = lng ? rk2_kernel<T, true, 7> : rk2_kernel<T, false, 7>; break;
    case 3: kern = lng ? rk2_kernel<T, true, 3> : rk2_kernel<T, false, 3>; break;
    case 2: kern = lng ? rk2_kernel<T, true, 2> : rk2_kernel<T, false, 2>; break;
    case 4: kern = lng ? rk2_kernel<T, true, 4> : rk2_kernel<T, false, 4>; break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, RK_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

template cudaError_t rk2_launch<double>(const FastArgs<double>&, unsigned, size_t, cudaStream_t);
template cudaError_t rk2_launch<float>(const FastArgs<float>&, unsigned, size_t, cudaStream_t);

}  // namespace mcd
