// mcd_rk2.cuh — the headline kernel, second generation: ess_rhat / rhat for kind in {rank, bulk,
// tail-rhat, basic} with the direct autocovariance, for slabs of exactly 8 split chains of <= 512
// draws each (e.g. the canonical 1000 draws x 4 chains, split_chains = 2).
//
// What changed against mcd_fast.cuh (round 1, 37 k warp-instructions per parameter, 7.5 % of roofline):
//   * PERSISTENT CTAs (2 per SM) with a TMA bulk-copy pipeline: one thread issues
//     `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` for the NEXT parameter's slab
//     (32 KB, one contiguous byte range: Julia column-major, utils.jl:203-211) as soon as the staging
//     buffer is dead, and the CTA waits on the mbarrier phase at the top of the next iteration: HBM
//     latency is hidden behind the autocovariance of the current parameter (SASS: UBLKCP / SYNCS).
//   * ONE ranking instead of two.  The fine-bucket counting rank (65 536 buckets, 4-bit packed
//     counters, see mcd_fast.cuh) now scatters EVERY value to its sorted slot, so the staging buffer
//     becomes the sorted slab S.  The folded series |x - median| (utils.jl:148-158) is monotone on each
//     side of the median, i.e. it is the merge of two sorted runs of S: its ranks come from a
//     merge-path (one diagonal search + 16 sequential merge steps per thread) on the COMPUTED folded
//     values, so ties created by the rounding of x - median are seen exactly as tiedrank sees them
//     (adjacent equal values of the merged sequence; resolved by a run scan only when one occurs).
//   * The median is S[n/2-1], S[n/2]: no capture pass.
//   * Geyer's truncation (ess_rhat.jl:553-594) is a thread-0 state machine whose decision is broadcast
//     through the lag batch's existing barrier (it ran redundantly on all 8 warps before).
// Ranks / ties stay bit-exact (StatsBase.tiedrank); slabs with NaN, an infinite range or a fine bucket
// holding >= 15 values go to the redo list and are recomputed by the general slab kernel, as before.
//
// Reference citations (/root/reference): utils.jl:13-41,148-193; ess_rhat.jl:362-409,488-624.
#pragma once
#include "mcd_common.cuh"
#include "mcd_slab.cuh"
#include "mcd_fast.cuh"
#include "mcd_rk2_api.cuh"
#include "mcd_tma.cuh"

// developer switch (A/B builds): 1 = the merge loop keeps one value of lookahead per run in registers (same-box A/B:
// 11.30 ms per 200 k parameters with it, 11.26 without: fewer selects win once every warp of the SM merges at once)
#ifndef RK_MERGE_LOOKAHEAD
#define RK_MERGE_LOOKAHEAD 0
#endif

namespace mcd {

constexpr int RK_THREADS = 256;
constexpr int RK_EPT = 16;                 // elements per thread
constexpr int RK_NCH = 8;                  // split chains = warps
constexpr int RK_MAXITER = 32 * RK_EPT;    // 512 draws per split chain
constexpr int RK_NMAX = RK_NCH * RK_MAXITER;
constexpr int RK_FINE = 65536;             // fine buckets
constexpr int RK_WORDS = RK_FINE / 8;      // counter words (8 nibbles each)
constexpr int RK_ROW = 616;                // padded centred-chain row (doubles)
constexpr int RK_TMAX = 576;               // the row is zero-filled on [niter, RK_TMAX)
constexpr int RK_SEG = 17;                 // merged values per thread in the fold merge (odd: conflict-free strides)
constexpr int RK_LAGS = 16;                // lags per autocovariance batch

// shared-memory layout (bytes)
//   XS   [0, 32 960)        staged slab (TMA destination); then S = XS + 2 elements: the sorted slab with sentinels S[-1] = NaN,
//                           S[n .. n + RK_SEG] = +inf (an exhausted run A loses every compare, an exhausted run B is endless)
//   A    48 KB multi-use:   FC u32[8192] + WP u16[8192]        (count / scan / position)
//                           WL u32[<=4096] + RES u32[4096]     (shared-bucket work list and its results)
//                           ZC rows + FR u16[4096] (last 8 KB)  (FR: folded table index per sorted slot)
//                           ZC f64[8][616] + rho[maxlag+17] + part f64[8][16]   (autocovariance)
//   small
constexpr int RK_OFF_XS = 0;
constexpr int RK_XS_BYTES = (RK_NMAX + 24) * 8;   // 2 elements in front of S, RK_SEG + 1 sentinels behind it
constexpr int RK_OFF_A = RK_XS_BYTES;
constexpr int RK_A_BYTES = RK_WORDS * 4 + RK_WORDS * 2;
constexpr int RK_A_WP = RK_WORDS * 4;
constexpr int RK_A_FR = RK_WORDS * 4 + RK_WORDS * 2 - (RK_NMAX + 32) * 2;   // end of region A (+ RK_SEG + 1 overrun slots): clear of the ZC rows
constexpr int RK_A_RHO = RK_NCH * RK_ROW * 8;
constexpr int RK_A_PART = RK_A_BYTES - RK_NCH * RK_LAGS * 8;
constexpr int RK_OFF_SMALL = RK_OFF_A + RK_A_BYTES;
constexpr int RK_SMALL_BYTES = 640;
constexpr int RK_SMEM_BYTES = RK_OFF_SMALL + RK_SMALL_BYTES;
static_assert(RK_A_FR >= RK_NCH * RK_ROW * 8, "FR must not overlap the centred rows");
static_assert((RK_A_PART - RK_A_RHO) / 8 - (RK_LAGS + 1) >= RK_MAXLAG_CAP && RK_MAXLAG_CAP >= RK_MAXITER, "rho[] must hold every admissible maxlag");
static_assert(RK_SEG * RK_THREADS >= RK_NMAX, "every merged value needs a thread");

template <typename T> __device__ __forceinline__ T rk_inf();
template <> __device__ __forceinline__ double rk_inf<double>() { return CUDART_INF; }
template <> __device__ __forceinline__ float rk_inf<float>() { return CUDART_INF_F; }


// Order-preserving integer key of a value's leading 32 bits (for double: the high word, i.e. sign, exponent and
// 20 mantissa bits; for float: the whole value).  key(a) < key(b) implies a < b, so the minimum / maximum KEY of
// a slab bounds its range from below / above by less than one unit of those 20 bits: all the bucket map needs.
// Integer min / max are one VIMNMX3 per two values and one REDUX per warp (FP64 min / max: DSETP + 2 FSEL per
// value and ten shuffles per warp).  NaN and +-Inf have keys beyond the finite range (they go to the redo list).
template <typename T> struct RkKey;
template <> struct RkKey<double> {
  static constexpr bool exact = false;
  static constexpr int pos_inf = 0x7ff00000, neg_inf = (int)0x800fffffu;
  static __device__ __forceinline__ int key(double v) { const int h = __double2hiint(v); return h ^ ((h >> 31) & 0x7fffffff); }
  static __device__ __forceinline__ double lower(int k) { const int h = k ^ ((k >> 31) & 0x7fffffff); return __hiloint2double(h, h < 0 ? -1 : 0); }
  static __device__ __forceinline__ double upper(int k) { const int h = k ^ ((k >> 31) & 0x7fffffff); return __hiloint2double(h, h < 0 ? 0 : -1); }
};
template <> struct RkKey<float> {
  static constexpr bool exact = true;
  static constexpr int pos_inf = 0x7f800000, neg_inf = (int)0x807fffffu;
  static __device__ __forceinline__ int key(float v) { const int h = __float_as_int(v); return h ^ ((h >> 31) & 0x7fffffff); }
  static __device__ __forceinline__ float lower(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
  static __device__ __forceinline__ float upper(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
};

// LONG = every split chain has more than 480 draws: only the last of a thread's 16 slots can be empty.
// sums of the eight 4-bit fields of each of the four words of c, as prefix pieces:
// a = s0 << 16, b = (s0 + s1) | (s0 + s1 + s2) << 16, returns s0 + s1 + s2 + s3
__device__ __forceinline__ unsigned rk_group_sums(const uint4 c, unsigned& a, unsigned& b) {
  const unsigned s0 = nibsum(c.x), s1 = nibsum(c.y), s2 = nibsum(c.z), s3 = nibsum(c.w);
  const unsigned s01 = s0 + s1, s012 = s01 + s2;
  a = s0 << 16;
  b = s01 | (s012 << 16);
  return s012 + s3;
}

// MODE: bit 0 = the bulk series is rank-normalised (else x itself: kind basic), bit 1 = bulk / basic step,
// bit 2 = tail step (fold + rank-normalise + R-hat).  rank = 7, bulk = 3, basic = 2, tail R-hat = 4.
// FULL = exactly 8 split chains (the canonical 4 chains x 2 halves).  Otherwise a.nch in 1..7 split chains: the warps
// without a chain hold no values (their slots are never live, their rows are zero) and still take part in the
// ranking's block-wide phases; every count that depends on the number of chains uses nch.
template <typename T, bool LONG, int MODE, bool FULL>
__global__ void __launch_bounds__(RK_THREADS, 2) rk2_kernel(const FastArgs<T> a) {
  constexpr bool RANKX = (MODE & 1) != 0, BULK = (MODE & 2) != 0, TAIL = (MODE & 4) != 0;
  extern __shared__ __align__(128) unsigned char smem[];
  const int niter = a.niter;
  const int nch = FULL ? RK_NCH : a.nch;
  const int n = nch * niter;
  T* XS = reinterpret_cast<T*>(smem + RK_OFF_XS);
  T* S = XS + 2;
  unsigned char* A = smem + RK_OFF_A;
  unsigned* FC = reinterpret_cast<unsigned*>(A);
  unsigned short* WP = reinterpret_cast<unsigned short*>(A + RK_A_WP);
  unsigned* WL = FC;
  unsigned* RES = reinterpret_cast<unsigned*>(A + RK_A_WP);
  T* M = XS;   // merged folded values (block-wide tie pass only): over the dead sorted slab
  unsigned short* FR = reinterpret_cast<unsigned short*>(A + RK_A_FR);
  double* ZC = reinterpret_cast<double*>(A);
  T* rhoa = reinterpret_cast<T*>(A + RK_A_RHO);            // [maxlag + 17]
  double* part = reinterpret_cast<double*>(A + RK_A_PART);  // [8][16]
  unsigned char* small = smem + RK_OFF_SMALL;
  T* cmean = reinterpret_cast<T*>(small);                   // [8]  bulk
  T* cvar = reinterpret_cast<T*>(small + 64);               // [8]
  T* cmean2 = reinterpret_cast<T*>(small + 128);            // [8]  tail
  T* cvar2 = reinterpret_cast<T*>(small + 192);             // [8]
  double* wred = reinterpret_cast<double*>(small + 256);    // [2][8]
  int* iflag = reinterpret_cast<int*>(small + 384);         // [8]
  unsigned* listlen = reinterpret_cast<unsigned*>(small + 416);
  int* decision = reinterpret_cast<int*>(small + 420);
  unsigned long long* mbar_ptr = reinterpret_cast<unsigned long long*>(small + 448);
  int* ikey = reinterpret_cast<int*>(small + 512);          // [2][8]  per-warp min / max keys

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool last_live = lane + 32 * (RK_EPT - 1) < niter;
  const bool has_chain = FULL || w < nch;
  auto live = [&](int k) -> bool { return has_chain && (LONG ? (k < RK_EPT - 1 || last_live) : (lane + 32 * k < niter)); };

  const unsigned xs_addr = smem_u32(XS), mbar = smem_u32(mbar_ptr);
  const unsigned slab_bytes = (unsigned)n * (unsigned)sizeof(T);
  if (tid == 0) {
    *listlen = 0;
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  long long param = blockIdx.x;
  if (tid == 0 && param < a.params) {
    mbar_arrive_expect_tx(mbar, slab_bytes);
    bulk_copy_g2s(xs_addr, a.x + param * (long long)n, slab_bytes, mbar);
  }
  unsigned phase = 0;
  constexpr bool need_rank = RANKX || TAIL;
  const int mA = (n & 1) ? n / 2 : n / 2 - 1, mB = n / 2;

  for (; param < a.params; param += gridDim.x) {
    // issue the bulk copy of the next parameter's slab; every generic-proxy access to XS of this iteration must
    // have been ordered before the calling point by a __syncthreads()
    bool prefetched = false;
    auto prefetch_next = [&]() {
      if (tid == 0) {
        const long long nx = param + gridDim.x;
        if (nx < a.params) {
          fence_proxy_async_smem();
          mbar_arrive_expect_tx(mbar, slab_bytes);
          bulk_copy_g2s(xs_addr, a.x + nx * (long long)n, slab_bytes, mbar);
        }
      }
      prefetched = true;
    };

    mbar_wait_parity(mbar, phase);
    phase ^= 1u;
    T x[RK_EPT];
    {
      const T* xs = XS + w * niter;
#pragma unroll
      for (int k = 0; k < RK_EPT; ++k) x[k] = live(k) ? xs[lane + 32 * k] : (T)0;
    }
    double ess = (double)Traits<T>::nan(), rhat_bulk = 0.0, rhat_tail = 0.0;
    bool redo = false, flat = false;   // flat: every value equal (all ranks tie)
    unsigned pi[RK_EPT];               // z-table index | sorted slot << 16
    T zflat = (T)0;

    do {   // single-trip block: `break` leaves for the redo list
      if (need_rank) {
        // clear the packed counters: the barrier of the min / max exchange covers it
        for (int i = tid; i < RK_WORDS / 4; i += RK_THREADS) reinterpret_cast<uint4*>(FC)[i] = make_uint4(0, 0, 0, 0);
        T vmin, vmax;
        bool exact_minmax = false;
        {
          int kmin = 0x7fffffff, kmax = (int)0x80000000u;
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) {
            const int key = RkKey<T>::key(x[k]);
            kmin = min(kmin, live(k) ? key : 0x7fffffff);
            kmax = max(kmax, live(k) ? key : (int)0x80000000u);
          }
          kmin = __reduce_min_sync(0xffffffffu, kmin);
          kmax = __reduce_max_sync(0xffffffffu, kmax);
          if (lane == 0) { ikey[w] = kmin; ikey[8 + w] = kmax; }
          __syncthreads();   // also: every thread holds its x in registers, XS is free to become S
          kmin = ikey[0]; kmax = ikey[8];
#pragma unroll
          for (int i = 1; i < RK_NCH; ++i) { kmin = min(kmin, ikey[i]); kmax = max(kmax, ikey[8 + i]); }
          if (kmax >= RkKey<T>::pos_inf || kmin <= RkKey<T>::neg_inf) { redo = true; break; }   // NaN or +-Inf
          // fewer than 64 key steps between the extremes: the key bounds would widen the range noticeably (or the
          // slab is constant): take the exact minimum and maximum instead
          exact_minmax = !RkKey<T>::exact && (long long)kmax - (long long)kmin < 64;
          vmin = RkKey<T>::lower(kmin); vmax = RkKey<T>::upper(kmax);
        }
        if (exact_minmax) {
          T lmin = rk_inf<T>(), lmax = -rk_inf<T>();
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) {
            if (live(k)) {
              const T v = x[k];
              lmin = v < lmin ? v : lmin;
              lmax = v > lmax ? v : lmax;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const T p = __shfl_xor_sync(0xffffffffu, lmin, o); lmin = p < lmin ? p : lmin;
            const T q = __shfl_xor_sync(0xffffffffu, lmax, o); lmax = q > lmax ? q : lmax;
          }
          if (lane == 0) { wred[w] = (double)lmin; wred[8 + w] = (double)lmax; }
          __syncthreads();
          vmin = (T)wred[0]; vmax = (T)wred[8];
#pragma unroll
          for (int i = 1; i < RK_NCH; ++i) {
            const T p = (T)wred[i], q = (T)wred[8 + i];
            vmin = p < vmin ? p : vmin; vmax = q > vmax ? q : vmax;
          }
        }
        flat = !(vmax > vmin);
        const T range = vmax - vmin;
        // slightly less than FINE / range: the largest value lands inside the last bucket, the smallest in bucket 0
        const T scale = (T)((double)RK_FINE * (1.0 - 1.0 / 1048576.0)) / range;
        if (!flat && (!(range < rk_inf<T>()) || !(scale > (T)0) || !(scale < rk_inf<T>()))) { redo = true; break; }
        if (flat) {
          zflat = __ldg(&a.ztab[((n - 1) >> 1) + (((n - 1) & 1) ? n : 0)]);   // every value ties: rank (n+1)/2
          __syncthreads();   // wred reads done before anything below reuses the small arrays
          prefetch_next();
        } else {
          // ---- count: 4-bit packed populations, ONE shared-memory atomic per element; it returns the arrival offset ----
          unsigned maxoff = 0, shared_mask = 0;
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) {
            if (live(k)) {
              const unsigned fb = (unsigned)(int)((x[k] - vmin) * scale);
              const unsigned sh = (fb & 7u) * 4u;
              const unsigned off = (atomicAdd(&FC[fb >> 3], 1u << sh) >> sh) & 15u;
              maxoff = off > maxoff ? off : maxoff;
              pi[k] = fb | (off << 16);
            } else pi[k] = 0;
          }
          // a counter that reaches 16 spills into its neighbour: the value that did it saw 15
          if (__syncthreads_or(maxoff >= 15u)) { redo = true; break; }
          // ---- scan: WP[word] = #values in earlier words.  Warp w owns words [1024 w, 1024 w + 1024); a lane reads
          // groups of 4 words g = lane + 32 j (conflict-free 128-bit loads); prefix over g = per-j warp scans, two j
          // per shuffle as 16-bit fields ----
          {
            const uint4* fc4 = reinterpret_cast<const uint4*>(FC + w * (RK_WORDS / RK_NCH));
            unsigned ga[8], gb[8], gt[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) gt[j] = rk_group_sums(fc4[lane + 32 * j], ga[j], gb[j]);
            unsigned incl[4], all[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) incl[q] = gt[2 * q] | (gt[2 * q + 1] << 16);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl[q], o);
                if (lane >= o) incl[q] += t;
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) all[q] = __shfl_sync(0xffffffffu, incl[q], 31);
            unsigned tot = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) tot += (all[q] & 0xffffu) + (all[q] >> 16);
            if (lane == 0) iflag[w] = (int)tot;
            __syncthreads();
            unsigned run = 0;   // values in the ranges of warps 0 .. w-1, then in the earlier j of this warp
#pragma unroll
            for (int i = 0; i < RK_NCH - 1; ++i) run += i < w ? (unsigned)iflag[i] : 0u;
            uint2* wp2 = reinterpret_cast<uint2*>(WP + w * (RK_WORDS / RK_NCH));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const unsigned field = (j & 1) ? (incl[j >> 1] >> 16) : (incl[j >> 1] & 0xffffu);
              const unsigned B = (run + field - gt[j]) * 0x00010001u;   // (base, base)
              wp2[lane + 32 * j] = make_uint2(B + ga[j], B + gb[j]);    // base, base + s0 | base + s01, base + s012
              run += (j & 1) ? (all[j >> 1] >> 16) : (all[j >> 1] & 0xffffu);
            }
            __syncthreads();
          }
          // ---- position: start of the fine bucket, population, own slot; EVERY value goes to its slot of S
          // (members of a shared bucket in arrival order for now) ----
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) {
            const bool valid = live(k);
            const unsigned fb = pi[k] & 0xffffu, off = pi[k] >> 16;
            const unsigned word = fb >> 3, sh = (fb & 7u) * 4u;
            const unsigned fw = FC[word];
            const unsigned st = (unsigned)WP[word] + nibsum(fw & ((1u << sh) - 1u));
            const unsigned c = valid ? ((fw >> sh) & 15u) : 0u;
            if (valid) S[st + off] = x[k];
            pi[k] = st | (c << 12) | (off << 16);   // st <= 4095
            shared_mask |= (c >= 2u ? 1u : 0u) << k;
          }
          __syncthreads();
          // ---- shared buckets (~10 % of the values): compacted into a work list (aliasing the dead counter words),
          // resolved by all threads: exact (less, equal) counts against the bucket mates on the values themselves ----
          {
            const unsigned mine = __popc(shared_mask);
            unsigned incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            unsigned base = 0;
            if (lane == 31) base = atomicAdd(listlen, incl);
            unsigned q = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
            const unsigned slot0 = (unsigned)tid << 20;
            if (mine) {
#pragma unroll
              for (int k = 0; k < RK_EPT; ++k)
                if (shared_mask & (1u << k)) WL[q++] = pi[k] | (slot0 + ((unsigned)k << 28));   // slot = k * 256 + tid
            }
          }
          __syncthreads();
          {
            const unsigned listn = *listlen;
            for (unsigned q = tid; q < listn; q += RK_THREADS) {
              const unsigned it = WL[q];
              const unsigned st = it & 0xfffu, c = (it >> 12) & 15u, off = (it >> 16) & 15u;
              const T v = S[st + off];
              unsigned less = 0, eq = 0, before = 0;   // before: equal mates that arrived earlier (distinct final slots)
              for (unsigned j = 0; j < c; ++j) {
                const T y = S[st + j];
                less += y < v;
                eq += y == v;        // == ties -0.0 with 0.0, as tiedrank's run detection does
                before += (y == v) & (j < off);
              }
              const unsigned lo = st + less;
              const unsigned zi = 2u * lo + eq - 1u;   // doubled average rank - 2
              RES[it >> 20] = ((zi >> 1) + ((zi & 1u) ? (unsigned)n : 0u)) | ((lo + before) << 16);   // split z table | final slot
            }
          }
          __syncthreads();   // all bucket-local reads of S are done
          if (tid == 0) { *listlen = 0; S[-1] = Traits<T>::nan(); }
          if (TAIL && tid >= 32 && tid < 32 + RK_SEG + 1) S[n + tid - 32] = rk_inf<T>();
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) {
            const unsigned st = pi[k] & 0xfffu;
            unsigned r = st * 0x00010001u;   // alone in its bucket: integer rank st + 1, table index st, final slot st
            if (shared_mask & (1u << k)) r = RES[k * RK_THREADS + tid];
            pi[k] = r;
          }
          if (TAIL) {
            if (shared_mask) {   // members of shared buckets move to their final slots: S is sorted
#pragma unroll
              for (int k = 0; k < RK_EPT; ++k)
                if (shared_mask & (1u << k)) S[pi[k] >> 16] = x[k];
            }
          } else { __syncthreads(); prefetch_next(); }
        }
      } else {
        __syncthreads();   // every thread holds its x in registers
        prefetch_next();
      }

      // ---- bulk / basic series: z of the ranks (utils.jl:175-193) or x itself ----
      T z[RK_EPT];
      if (BULK) {
        if (RANKX) {
          if (flat) {
#pragma unroll
            for (int k = 0; k < RK_EPT; ++k) z[k] = live(k) ? zflat : (T)0;
          } else {
#pragma unroll
            for (int k = 0; k < RK_EPT; ++k) z[k] = live(k) ? __ldg(&a.ztab[pi[k] & 0xffffu]) : (T)0;
          }
        } else {
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) z[k] = x[k];
        }
        // split-chain moments (ess_rhat.jl:529-545): warp w owns split chain w
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < RK_EPT; ++k) if (live(k)) s += (double)z[k];
        s = warp_sum(s);
        const T m = (T)(s / (double)niter);
        double q = 0.0;
#pragma unroll
        for (int k = 0; k < RK_EPT; ++k) {
          z[k] = live(k) ? (T)(z[k] - m) : (T)0;   // samples .-= chain_mean (ess_rhat.jl:548)
          q = fma((double)z[k], (double)z[k], q);
        }
        q = warp_sum(q);
        if (lane == 0) { cmean[w] = m; cvar[w] = (T)(q / (double)(niter - 1)); }
      }

      // the centred chain as a padded row (index t + (t >> 4)) of region A: written before the fold merge so that z
      // does not stay in registers across it
      const bool do_ess = BULK && a.want_ess && !a.ess_nan;
      bool rows_written = false;
      auto write_rows = [&]() {
        // t = lane + 32 k sits at t + (t >> 4) = (lane + (lane >> 4)) + 34 k: one base, compile-time offsets
        double* rp = ZC + w * RK_ROW + lane + (lane >> 4);
#pragma unroll
        for (int k = 0; k < RK_EPT; ++k) rp[34 * k] = (double)z[k];
#pragma unroll
        for (int k = RK_EPT; k < RK_TMAX / 32; ++k) rp[34 * k] = 0.0;   // zero fill of [RK_MAXITER, RK_TMAX)
        rows_written = true;
      };

      // ---- tail series: rank-normalised |x - median| (utils.jl:148-158), ranks by merging the two sorted runs of S ----
      if (TAIL) {
        T zt[RK_EPT];
        if (flat) {
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) zt[k] = live(k) ? zflat : (T)0;
        } else {
          __syncthreads();   // S is final; WL / RES are dead
          if (do_ess) write_rows();
          const T sA = S[mA], sB = S[mB];
          // Statistics.median: middle of the two central order statistics
          const T med = (n & 1) ? sA : (T)(sA / (T)2 + sB / (T)2);
          const int L = n >> 1, nA = L, nB = n - L;   // run A = slots [0, L) read downwards, run B = slots [L, n)
          const int m0 = tid * RK_SEG;
          // The folded values are |d| for the signed differences d = S[slot] - med: never materialised, the compares take
          // absolute values of their operands.  tiebits: bit i = merged value i of this segment equals its predecessor
          // (bit 0: the last value of the previous segment)
          unsigned tiebits = 0;
          int pa0 = 0, pb0 = 0;
          bool cross = false;
          if (m0 < n) {
            // merge path: the first m0 merged values take `lo` from run A and m0 - lo from run B (A first among equals)
            int lo = m0 - nB > 0 ? m0 - nB : 0, hi = m0 < nA ? m0 : nA;
            const T* SA = S + (L - 1);          // run A element i is SA[-i]
            const T* SB = S + (L + m0 - 1);     // its opponent on the diagonal is SB[-i]
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              const T da = SA[-mid] - med, db = SB[-mid] - med;
              if (fabs(da) <= fabs(db)) lo = mid + 1; else hi = mid;
            }
            int pa = L - 1 - lo, pb = L + (m0 - lo);   // next slot of each run (sentinels: S[-1] = NaN, S[n ..] = +inf)
            pa0 = pa; pb0 = pb;
            T prev = (T)-1;   // the merged value just before this segment (no folded value is negative... as |.|: use 1 vs -1 trick below)
            bool has_prev = false;
            if (lo > 0) { prev = S[pa + 1] - med; has_prev = true; }
            if (m0 - lo > 0) { const T t = S[pb - 1] - med; if (!has_prev || fabs(t) > fabs(prev)) prev = t; has_prev = true; }
            T da = S[pa] - med, db = S[pb] - med;
#if RK_MERGE_LOOKAHEAD
            T da2 = S[pa - 1] - med, db2 = S[pb + 1] - med;   // one value of lookahead per run: the load of a step is not on its critical path
#endif
            if (!has_prev) prev = rk_inf<T>();   // |prev| = inf never equals a merged value of this segment (data are finite)
#pragma unroll
            for (int i = 0; i < RK_SEG; ++i) {
              // (steps past the end of the slab walk the +inf sentinels behind S; their stores land in the overrun slots
              // behind FR and their tie bits are masked off below)
              const bool takeA = fabs(da) <= fabs(db);
              const int slot = takeA ? pa : pb;
              FR[slot] = (unsigned short)(m0 + i);   // untied: integer rank m0 + i + 1 -> table index m0 + i
              const T f = takeA ? da : db;
              if (fabs(f) == fabs(prev)) tiebits |= 1u << i;
              prev = f;
#if RK_MERGE_LOOKAHEAD
              if (takeA) { --pa; da = da2; } else { ++pb; db = db2; }
              const T d = S[takeA ? pa - 1 : pb + 1] - med;
              if (takeA) da2 = d; else db2 = d;
#else
              if (takeA) --pa; else ++pb;
              const T d = S[takeA ? pa : pb] - med;
              if (takeA) da = d; else db = d;
#endif
            }
            const int cnt = n - m0 < RK_SEG ? n - m0 : RK_SEG;
            tiebits &= (cnt >= 32 ? 0xffffffffu : (1u << cnt) - 1u);
            // a run of equal values that crosses a segment boundary (the head of the remaining runs is the next
            // segment's first value) is left to the block-wide pass below
            const T nxt = fabs(da) <= fabs(db) ? da : db;
            cross = (tiebits & 1u) || (cnt == RK_SEG && m0 + RK_SEG < n && fabs(nxt) == fabs(prev));
          }
          if (__syncthreads_or(cross)) {
            // ---- ties across segments (duplicated draws): average ranks over runs of equal folded values, block-wide.
            // Slot p sits at merged index FR[p]; M = the merged sequence ----
            T fv[RK_EPT];
#pragma unroll
            for (int k = 0; k < RK_EPT; ++k) { const int p = tid + RK_THREADS * k; fv[k] = p < n ? fabs(S[p] - med) : (T)0; }
            __syncthreads();   // every read of S is done: M overwrites it
#pragma unroll
            for (int k = 0; k < RK_EPT; ++k) { const int p = tid + RK_THREADS * k; if (p < n) M[FR[p]] = fv[k]; }
            __syncthreads();
            for (int p = tid; p < n; p += RK_THREADS) {
              const int mi = FR[p];
              const T f = M[mi];
              int lo = mi, hi = mi + 1;
              while (lo > 0 && M[lo - 1] == f) --lo;
              while (hi < n && M[hi] == f) ++hi;
              const unsigned zi = (unsigned)(lo + hi - 1);
              FR[p] = (unsigned short)((zi >> 1) + ((zi & 1u) ? (unsigned)n : 0u));
            }
            __syncthreads();   // M (over the staging buffer) is dead
            prefetch_next();
          } else {
            // runs inside this segment (the two central values always tie): walk the segment's merge again and give
            // the members of each run their average rank.  t: bit j = element j equals element j - 1
            const unsigned t = tiebits;
            if (t) {
              int pa = pa0, pb = pb0;
              const unsigned inrun = t | (t >> 1);
              for (int i = 0; i < RK_SEG && (inrun >> i); ++i) {
                const bool takeA = fabs(S[pa] - med) <= fabs(S[pb] - med);
                const int slot = takeA ? pa : pb;
                if ((inrun >> i) & 1u) {
                  const int back = __clz(~(t << (31 - i)));         // ones of t ending at bit i: the run starts at i - back
                  const int fwd = __ffs(~(t >> (i + 1))) - 1;       // ones of t starting at bit i + 1: it ends at i + fwd
                  const unsigned zi = (unsigned)(2 * m0 + 2 * i - back + fwd);   // lo + hi - 1 with lo = m0 + s, hi = m0 + e + 1
                  FR[slot] = (unsigned short)((zi >> 1) + ((zi & 1u) ? (unsigned)n : 0u));
                }
                if (takeA) --pa; else ++pb;
              }
            }
            __syncthreads();   // every read of S is done
            prefetch_next();
          }
          // (the barrier that made S dead also made FR final)
#pragma unroll
          for (int k = 0; k < RK_EPT; ++k) zt[k] = live(k) ? __ldg(&a.ztab[FR[pi[k] >> 16]]) : (T)0;
        }
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < RK_EPT; ++k) if (live(k)) s += (double)zt[k];
        s = warp_sum(s);
        const T m = (T)(s / (double)niter);
        double q = 0.0;
#pragma unroll
        for (int k = 0; k < RK_EPT; ++k) if (live(k)) { const T d = zt[k] - m; q = fma((double)d, (double)d, q); }
        q = warp_sum(q);
        if (lane == 0) { cmean2[w] = m; cvar2[w] = (T)(q / (double)(niter - 1)); }
      }
      __syncthreads();   // chain moments visible; region A (M / FR / RES) is dead

      // ---- R-hat (ess_rhat.jl:387-408): warp 0 forms W and var_plus, thread 0 the ratios ----
      SplitGeom g8;
      g8.niter = niter; g8.nch = nch;
      T W = (T)0, var_plus = (T)1;
      if (w == 0) {
        if (TAIL) {
          T W2, vp2;
          within_between<T>(cmean2, cvar2, g8, W2, vp2);
          if (tid == 0) rhat_tail = (double)sqrt(vp2 / W2);
        }
        if (BULK) {
          within_between<T>(cmean, cvar, g8, W, var_plus);
          if (tid == 0) rhat_bulk = (double)sqrt(var_plus / W);
        }
      }
      if (!do_ess) break;

      // ---- direct autocovariance, lazily in batches of 16 lags; Geyer's truncation on thread 0 ----
      if (!rows_written) { write_rows(); __syncwarp(); }
      const int maxlag = a.maxlag;
      const T inv_var_plus = (T)1 / var_plus;
      // thread 0: Geyer state (ess_rhat.jl:553-594)
      T g_p = (T)0, g_sum = (T)0, g_even = (T)1;
      int g_k = 2, g_stage = 0;   // 0 = needs rho_1, 1 = pair loop, 2 = needs the final rho_k, 3 = done
      int have = 0;
      for (;;) {
        // lags k0 .. k0 + 15, k0 = 16 b + 1: lane l owns the draws [16 l, 16 l + 16) of its chain (padded index
        // 17 l + i) and slides a window of 31 draws over them; draw 16 l + k0 + j sits at 17 (l + b) + j + (j < 15 ? 1 : 2)
        const int k0 = have + 1, b = have >> 4;
        const double* row = ZC + w * RK_ROW;
        const double* ownp = row + 17 * lane;
        double acc[RK_LAGS];
#pragma unroll
        for (int kk = 0; kk < RK_LAGS; ++kk) acc[kk] = 0.0;
        double win[2 * RK_LAGS - 1];
        if (b <= 3) {   // the whole window lies inside the zero-filled row: no bound checks
          const double* base = row + 17 * (lane + b);
#pragma unroll
          for (int j = 0; j < 2 * RK_LAGS - 1; ++j) win[j] = base[j + (j < 15 ? 1 : 2)];
        } else {
#pragma unroll
          for (int j = 0; j < 2 * RK_LAGS - 1; ++j) { const int t = 16 * lane + k0 + j; win[j] = t < RK_TMAX ? row[t + (t >> 4)] : 0.0; }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const double o = ownp[i];
#pragma unroll
          for (int kk = 0; kk < RK_LAGS; ++kk) acc[kk] = fma(o, win[i + kk], acc[kk]);
        }
        {
          double lo8[8], hi8[8];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) { lo8[kk] = acc[kk]; hi8[kk] = acc[8 + kk]; }
          const double t0 = warp_reduce8(lo8), t1 = warp_reduce8(hi8);
          if ((lane & 3) == 0) { part[w * RK_LAGS + (lane >> 2)] = t0; part[w * RK_LAGS + 8 + (lane >> 2)] = t1; }
        }
        have += RK_LAGS;
        __syncthreads();
        if (w == 0) {
          if (tid < RK_LAGS) {
            const int k = k0 + tid;
            if (k <= maxlag && k < niter) {
              double sum = 0.0;
#pragma unroll
              for (int i = 0; i < RK_NCH; ++i) sum += part[i * RK_LAGS + tid];
              const T gk = (T)(sum / (double)nch) / (T)niter;
              rhoa[k] = (T)1 - inv_var_plus * (W - gk);   // rho_k (ess_rhat.jl:556,566-567)
            }
          }
          __syncwarp();
          if (tid == 0) {
            int done = 0;
            if (g_stage == 0) { const T r1 = rhoa[1]; g_p = (T)1 + r1; g_sum = g_p; g_stage = 1; }
            if (g_stage == 1) {
              for (;;) {
                if (!(g_k < maxlag - 1)) { g_stage = 2; break; }
                if (g_k + 1 > have) break;   // next batch
                g_even = rhoa[g_k];
                const T delta = g_even + rhoa[g_k + 1];
                if (!(delta > (T)0)) { g_stage = 3; break; }   // rho_even = rho_k is already at hand
                g_p = jl_min<T>(delta, g_p);
                g_sum += g_p;
                g_k += 2;
              }
            }
            if (g_stage == 2) {
              if (maxlag > 1) { if (g_k <= have) { g_even = rhoa[g_k]; g_stage = 3; } }
              else { g_even = (T)0; g_stage = 3; }
            }
            if (g_stage == 3) {
              done = 1;
              const T tau = jl_max<T>((T)0, (T)2 * g_sum + jl_max<T>((T)0, g_even) - (T)1);
              T e = jl_min<T>((T)1 / tau, a.rel_ess_max);
              if (!a.relative) e *= (T)(niter * nch);
              ess = (double)e;
            }
            *decision = done;
          }
        }
        __syncthreads();
        if (*decision) break;
      }
    } while (false);

    if (redo) {
      if (tid == 0) { const int idx = atomicAdd(a.redo_count, 1); a.redo_list[idx] = (int)param; }
    } else if (tid == 0) {
      double rhat = rhat_bulk;
      if (TAIL) rhat = BULK ? (double)jl_max<T>((T)rhat_tail, (T)rhat_bulk) : rhat_tail;
      if (a.ess_out) a.ess_out[param] = (T)ess;
      if (a.rhat_out) a.rhat_out[param] = (T)rhat;
    }
    __syncthreads();   // region A / small arrays / XS are free for the next parameter
    if (!prefetched) prefetch_next();
  }
}
}  // namespace mcd
