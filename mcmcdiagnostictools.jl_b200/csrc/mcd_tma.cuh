// mcd_tma.cuh — inline PTX for the Blackwell / Hopper async-copy machinery the kernels use: mbarrier objects in
// shared memory and 1-D bulk copies global -> shared executed by the TMA engine (SASS: UBLKCP, SYNCS).
#pragma once
#include <cuda_runtime.h>

namespace mcd {

// ---- PTX: mbarrier + 1-D bulk async copy (TMA engine) ---------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "MCD_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra MCD_DONE;\n"
      "bra MCD_WAIT;\n"
      "MCD_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}

}  // namespace mcd
