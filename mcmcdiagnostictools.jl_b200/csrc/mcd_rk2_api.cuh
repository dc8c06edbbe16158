// mcd_rk2_api.cuh — what the call driver (mcd_api.cu) needs from the headline kernel's translation unit.
#pragma once
#include "mcd_fast.cuh"

namespace mcd {

constexpr int RK_MAXLAG_CAP = 512;   // largest maxlag the kernel's rho[] array holds (checked in mcd_rk2.cuh)

// Host launcher (defined in mcd_rk2.cu, its own translation unit so that the kernel builds in seconds):
// persistent grid of `grid` CTAs, RK_SMEM_BYTES + extra_smem bytes of dynamic shared memory.
template <typename T>
cudaError_t rk2_launch(const FastArgs<T>& a, unsigned grid, size_t extra_smem, cudaStream_t stream);

}  // namespace mcd
