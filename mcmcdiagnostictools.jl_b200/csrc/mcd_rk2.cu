// mcd_rk2.cu — translation unit of the headline kernel (mcd_rk2.cuh) and its launcher.
#include "mcd_rk2.cuh"

namespace mcd {

template <typename T>
cudaError_t rk2_launch(const FastArgs<T>& a, unsigned grid, size_t extra_smem, cudaStream_t stream) {
  const size_t smem = (size_t)RK_SMEM_BYTES + extra_smem;
  const int mode = (a.rank_x ? 1 : 0) | (a.do_bulk ? 2 : 0) | (a.do_tail ? 4 : 0);
  const bool lng = a.niter > 32 * (RK_EPT - 1);
  void (*kern)(const FastArgs<T>) = nullptr;
  const bool full = a.nch == RK_NCH;
#define RK2_PICK(M) kern = full ? (lng ? rk2_kernel<T, true, M, true> : rk2_kernel<T, false, M, true>) \
                                : (lng ? rk2_kernel<T, true, M, false> : rk2_kernel<T, false, M, false>)
  switch (mode) {
    case 7: RK2_PICK(7); break;
    case 3: RK2_PICK(3); break;
    case 2: RK2_PICK(2); break;
    case 4: RK2_PICK(4); break;
    default: return cudaErrorInvalidValue;
  }
#undef RK2_PICK
  if (a.nch < 1 || a.nch > RK_NCH) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, RK_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

template cudaError_t rk2_launch<double>(const FastArgs<double>&, unsigned, size_t, cudaStream_t);
template cudaError_t rk2_launch<float>(const FastArgs<float>&, unsigned, size_t, cudaStream_t);

}  // namespace mcd
