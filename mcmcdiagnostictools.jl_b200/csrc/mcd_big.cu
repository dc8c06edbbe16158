// mcd_big.cu — translation unit of the big-slab estimator kernel (mcd_big.cuh) and its launcher.
#include "mcd_big.cuh"

namespace mcd {

static size_t bg_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T>
size_t big_smem_bytes(const SplitGeom& g, int maxlag, int* off_aux, int* off_part, int* off_small) {
  const size_t slab = bg_align((size_t)g.n * sizeof(T), 128);
  if (((size_t)g.n * sizeof(T)) % 16 != 0 || g.nch > 64 || g.niter < 2) return 0;
  const size_t part = bg_align((size_t)BG_BINS * 4, 128);   // the histogram (and the candidates) come first
  size_t sm = part + (size_t)g.nch * (3 * BG_LAGS + 3) * 8 + (size_t)g.nch * 2 * sizeof(T) + (size_t)(maxlag + 9) * sizeof(T);
  sm = bg_align(sm, 128);
  *off_aux = (int)slab; *off_part = (int)part; *off_small = (int)sm;
  return slab + sm + 640;
}

template <typename T> cudaError_t big_launch(BigArgs<T> a, unsigned grid, cudaStream_t stream) {
  const size_t smem = big_smem_bytes<T>(a.g, a.maxlag, &a.off_aux, &a.off_part, &a.off_small);
  if (!smem) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(big_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  big_kernel<T><<<grid, BG_THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

template size_t big_smem_bytes<double>(const SplitGeom&, int, int*, int*, int*);
template size_t big_smem_bytes<float>(const SplitGeom&, int, int*, int*, int*);
template cudaError_t big_launch<double>(BigArgs<double>, unsigned, cudaStream_t);
template cudaError_t big_launch<float>(BigArgs<float>, unsigned, cudaStream_t);

}  // namespace mcd
