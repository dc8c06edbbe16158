// mcd_big.cuh — estimator ESS (+ R-hat) for slabs that fill most of an SM's shared memory, e.g. BASELINE config 5:
// 4000 draws x 8 chains Float32 (128 KB, 16 split chains of 2000), kinds mean / std / median, AutocovMethod or
// BDAAutocovMethod.  Reference: _ess(estimator) -> _expectand_proxy (src/ess_rhat.jl:291-297, 628-646) ->
// _ess_rhat_basic! (:488-603) with mean_autocov of :161-179 (direct) or :197-213 (BDA variogram).
//
// One persistent CTA of 512 threads per SM; per parameter:
//   1. the slab arrives in shared memory by TMA bulk copies (cp.async.bulk + mbarrier) and never leaves it;
//   2. the proxy overwrites it in place: x (mean), (x - mean(x))^2 (std), x <= median(x) (median).  The median is
//      two order statistics of the slab: a linear 8192-bucket histogram (shared-memory atomics) finds their bucket(s),
//      the few candidates of those buckets are ranked exactly by counting; slabs whose range is not finite (or whose
//      candidate set is large: heavy ties) use a bisection on the order-preserving integer keys instead;
//   3. split-chain moments (warp per chain, two passes, Float64 accumulation), R-hat, centring in place;
//   4. lagged products, lazily in batches of 8 lags: lane l of the chain's warp owns a contiguous block of draws
//      (odd length: conflict-free) and slides a register window over it (8 own values x 15 window values -> 64 FMAs per
//      16 loads); products are summed in T over 8 terms, then in Float64.  BOTH estimators come from these sums: the
//      BDA variogram sum_t (y_t - y_{t+k})^2 = Q - tail_k + Q - head_k - 2 sum_t y_t y_{t+k} (Q = sum of squares of
//      the chain, head_k / tail_k = of its first / last k values), whose cancellation costs (1 + rho) / (1 - rho) ulps;
//   5. Geyer's initial-positive / monotone truncation (:553-594) as a thread-0 state machine after each batch.
#pragma once
#include "mcd_common.cuh"
#include "mcd_slab.cuh"
#include "mcd_fast.cuh"
#include "mcd_tma.cuh"
#include "mcd_big_api.cuh"

namespace mcd {

constexpr int BG_THREADS = 512;
constexpr int BG_NW = BG_THREADS / 32;
constexpr int BG_BINS = 8192;
constexpr int BG_CAND = 1024;        // candidates of the median buckets kept for exact ranking
constexpr int BG_LAGS = 8;

template <typename T> __device__ __forceinline__ T bg_inf();
template <> __device__ __forceinline__ double bg_inf<double>() { return CUDART_INF; }
template <> __device__ __forceinline__ float bg_inf<float>() { return CUDART_INF_F; }

// k-th smallest (0-based) order-preserving key of Y[0..n) by bisection on the key space: the smallest key K with
// #{key <= K} >= k + 1.  Block-wide; `red` = BG_NW + 1 unsigned words of scratch.  NaNs (largest key) rank last.
template <typename T>
__device__ typename Traits<T>::Key bg_select_key(const T* Y, int n, int k, unsigned* red) {
  using Key = typename Traits<T>::Key;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Key lo = 0, hi = ~(Key)0;
  while (lo < hi) {
    const Key mid = lo + ((hi - lo) >> 1);
    unsigned c = 0;
    for (int i = tid; i < n; i += BG_THREADS) c += order_key(Y[i]) <= mid;
    c = __reduce_add_sync(0xffffffffu, c);
    __syncthreads();
    if (lane == 0) red[w] = c;
    __syncthreads();
    unsigned tot = 0;
#pragma unroll
    for (int i = 0; i < BG_NW; ++i) tot += red[i];
    if (tot >= (unsigned)k + 1u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

template <typename T>
__global__ void __launch_bounds__(BG_THREADS, 1) big_kernel(const BigArgs<T> a) {
  using Key = typename Traits<T>::Key;
  extern __shared__ __align__(128) unsigned char smem[];
  const SplitGeom g = a.g;
  const int n = g.n, niter = g.niter, nch = g.nch;
  T* Y = reinterpret_cast<T*>(smem);
  unsigned char* aux = smem + a.off_aux;
  unsigned* HIST = reinterpret_cast<unsigned*>(aux);                       // [BG_BINS] (median)
  T* CAND = reinterpret_cast<T*>(aux);                                       // [BG_CAND] (after the histogram is dead)
  double* part = reinterpret_cast<double*>(aux + a.off_part);               // [nch][BG_LAGS]
  double* headp = part + (size_t)nch * BG_LAGS;                             // [nch][BG_LAGS] sum of squares of the first k values
  double* tailp = headp + (size_t)nch * BG_LAGS;                            // [nch][BG_LAGS] ... of the last k values
  double* qsum = tailp + (size_t)nch * BG_LAGS;                             // [nch] sum of squares of the centred chain
  double* hsum = qsum + nch;                                                // [nch] running head sums (first `have` values)
  double* tsum = hsum + nch;                                                // [nch] running tail sums (last `have` values)
  T* cmean = reinterpret_cast<T*>(tsum + nch);                               // [nch]
  T* cvar = cmean + nch;                                                     // [nch]
  T* rhoa = cvar + nch;                                                      // [maxlag + 9]
  double* red = reinterpret_cast<double*>(aux + a.off_small);               // [40] block reductions
  unsigned* ured = reinterpret_cast<unsigned*>(red + 40);                   // [BG_NW + 8]
  int* decision = reinterpret_cast<int*>(ured + BG_NW + 8);
  unsigned* ncand = reinterpret_cast<unsigned*>(decision + 1);
  unsigned long long* mbar_ptr = reinterpret_cast<unsigned long long*>(aux + a.off_small + 512);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned y_addr = smem_u32(Y), mbar = smem_u32(mbar_ptr);
  const unsigned slab_bytes = (unsigned)n * (unsigned)sizeof(T);
  if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
  __syncthreads();
  unsigned phase = 0;
  // lane blocks of the lag loop: B draws per lane, odd (bank-conflict-free stride), 32 B >= niter
  int B = (niter + 31) / 32;
  B |= 1;

  for (long long param = blockIdx.x; param < a.params; param += gridDim.x) {
    if (tid == 0) {
      fence_proxy_async_smem();
      mbar_arrive_expect_tx(mbar, slab_bytes);
      const char* src = reinterpret_cast<const char*>(a.x + param * (long long)n);
      for (unsigned off = 0; off < slab_bytes; off += 65536u) {
        const unsigned len = slab_bytes - off < 65536u ? slab_bytes - off : 65536u;
        bulk_copy_g2s(y_addr + off, src + off, len, mbar);
      }
    }
    mbar_wait_parity(mbar, phase);
    phase ^= 1u;

    // ---- expectand proxy, in place (src/ess_rhat.jl:628-646) ----
    if (a.proxy == 1) {
      // (x .- mean(x; dims=(1,2))).^2
      double s = 0.0;
      for (int i = tid; i < n; i += BG_THREADS) s += (double)Y[i];
      s = block_sum<BG_THREADS>(s, red);
      const T m = (T)(s / (double)n);
      for (int i = tid; i < n; i += BG_THREADS) { const T d = Y[i] - m; Y[i] = d * d; }
    } else if (a.proxy == 2) {
      // x .<= median(x): Statistics.median = middle of the two central order statistics; NaN if any NaN
      const int kA = (n - 1) >> 1, kB = n >> 1;
      T lmin = bg_inf<T>(), lmax = -bg_inf<T>();
      int bad = 0;
      for (int i = tid; i < n; i += BG_THREADS) {
        const T v = Y[i];
        bad |= (v != v);
        lmin = v < lmin ? v : lmin;
        lmax = v > lmax ? v : lmax;
      }
      for (int i = tid; i < BG_BINS; i += BG_THREADS) HIST[i] = 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T p = __shfl_xor_sync(0xffffffffu, lmin, o); lmin = p < lmin ? p : lmin;
        const T q = __shfl_xor_sync(0xffffffffu, lmax, o); lmax = q > lmax ? q : lmax;
      }
      bad = __any_sync(0xffffffffu, bad);
      if (lane == 0) { red[w] = (double)lmin; red[BG_NW + w] = (double)lmax; ured[w] = (unsigned)bad; }
      if (tid == 0) *ncand = 0u;
      __syncthreads();
      T vmin = (T)red[0], vmax = (T)red[BG_NW];
      unsigned anybad = ured[0];
#pragma unroll
      for (int i = 1; i < BG_NW; ++i) {
        const T p = (T)red[i], q = (T)red[BG_NW + i];
        vmin = p < vmin ? p : vmin; vmax = q > vmax ? q : vmax;
        anybad |= ured[i];
      }
      __syncthreads();   // red / ured are free again
      T med;
      if (anybad) {
        med = Traits<T>::nan();
      } else {
        const T range = vmax - vmin;
        const T scale = (T)((double)BG_BINS * (1.0 - 1.0 / 1048576.0)) / range;
        bool fast = vmax > vmin && range < bg_inf<T>() && scale > (T)0 && scale < bg_inf<T>();
        T oA = vmin, oB = vmin;   // the two order statistics (all values equal: vmin)
        if (fast) {
          for (int i = tid; i < n; i += BG_THREADS) atomicAdd(&HIST[(unsigned)(int)((Y[i] - vmin) * scale)], 1u);
          __syncthreads();
          // exclusive scan of the 8192 counts: 16 per thread, warp scan, warp totals
          unsigned c[16], run = 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) { c[i] = HIST[tid * 16 + i]; run += c[i]; }
          unsigned incl = run;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          if (lane == 31) ured[w] = incl;
          __syncthreads();
          unsigned before = incl - run;
#pragma unroll
          for (int i = 0; i < BG_NW; ++i) before += i < w ? ured[i] : 0u;
          // the thread whose buckets hold rank kA / kB publishes (bucket, ranks below it)
          unsigned acc = before;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (c[i] && acc <= (unsigned)kA && (unsigned)kA < acc + c[i]) { ured[BG_NW] = tid * 16 + i; ured[BG_NW + 1] = acc; }
            if (c[i] && acc <= (unsigned)kB && (unsigned)kB < acc + c[i]) { ured[BG_NW + 2] = tid * 16 + i; ured[BG_NW + 3] = acc; }
            acc += c[i];
          }
          __syncthreads();   // the histogram is dead: CAND aliases it
          const unsigned bA = ured[BG_NW], baseA = ured[BG_NW + 1], bB = ured[BG_NW + 2], baseB = ured[BG_NW + 3];
          for (int i = tid; i < n; i += BG_THREADS) {
            const T v = Y[i];
            const unsigned b = (unsigned)(int)((v - vmin) * scale);
            if (b == bA || b == bB) {
              const unsigned q = atomicAdd(ncand, 1u);
              if (q < (unsigned)BG_CAND) CAND[q] = v;
            }
          }
          __syncthreads();
          const unsigned nc = *ncand;
          if (nc > (unsigned)BG_CAND) {
            fast = false;   // heavy ties in the middle: bisection below
          } else {
            // exact ranks among the candidates: candidate q is the order statistic of rank base(bucket) + #{less} (+ ties)
            for (unsigned q = tid; q < nc; q += BG_THREADS) {
              const T v = CAND[q];
              const unsigned b = (unsigned)(int)((v - vmin) * scale);
              unsigned less = 0, eq = 0;
              for (unsigned j = 0; j < nc; ++j) {
                const T y = CAND[j];
                const bool same_bucket = (unsigned)(int)((y - vmin) * scale) == b;
                less += same_bucket && y < v;
                eq += same_bucket && y == v;
              }
              const unsigned lo = (b == bA ? baseA : baseB) + less;
              if (lo <= (unsigned)kA && (unsigned)kA < lo + eq) red[0] = (double)v;
              if (lo <= (unsigned)kB && (unsigned)kB < lo + eq) red[1] = (double)v;
            }
            __syncthreads();
            oA = (T)red[0]; oB = (T)red[1];
            __syncthreads();
          }
        }
        if (!fast && vmax > vmin) {
          oA = key_value(bg_select_key<T>(Y, n, kA, ured));
          oB = kB == kA ? oA : key_value(bg_select_key<T>(Y, n, kB, ured));
        }
        med = (n & 1) ? oA : (T)(oA / (T)2 + oB / (T)2);
      }
      __syncthreads();
      for (int i = tid; i < n; i += BG_THREADS) Y[i] = Y[i] <= med ? (T)1 : (T)0;
    }
    __syncthreads();

    // ---- split-chain moments (ess_rhat.jl:529-545), centring (:548), sums of squares of the centred chains ----
    for (int j = w; j < nch; j += BG_NW) {
      T* p = Y + g.chain_start(j);
      double s = 0.0;
      for (int t = lane; t < niter; t += 32) s += (double)p[t];
      s = warp_sum(s);
      const T m = (T)(s / (double)niter);
      double q = 0.0;
      for (int t = lane; t < niter; t += 32) { const T d = p[t] - m; p[t] = d; q = fma((double)d, (double)d, q); }
      q = warp_sum(q);
      if (lane == 0) { cmean[j] = m; cvar[j] = (T)(q / (double)(niter - 1)); qsum[j] = q; hsum[j] = 0.0; tsum[j] = 0.0; }
    }
    __syncthreads();
    T W = (T)0, var_plus = (T)1;
    double ess = (double)Traits<T>::nan(), rhat = 0.0;
    if (w == 0) {
      within_between<T>(cmean, cvar, g, W, var_plus);
      if (tid == 0) rhat = (double)sqrt(var_plus / W);
    }
    if (a.want_ess && !a.ess_nan) {
      const int maxlag = a.maxlag;
      const T inv_var_plus = (T)1 / var_plus;
      T g_p = (T)0, g_sum = (T)0, g_even = (T)1;
      int g_k = 2, g_stage = 0;   // Geyer state of thread 0: 0 = needs rho_1, 1 = pair loop, 2 = needs the final rho_k, 3 = done
      int have = 0;
      for (;;) {
        const int k0 = have + 1;
        for (int j = w; j < nch; j += BG_NW) {
          const T* p = Y + g.chain_start(j);
          auto at = [&](int t) -> T { return t < niter ? p[t] : (T)0; };
          const int t0 = lane * B;
          double acc[BG_LAGS];
#pragma unroll
          for (int kk = 0; kk < BG_LAGS; ++kk) acc[kk] = 0.0;
          T win[BG_LAGS + 7];
#pragma unroll
          for (int i = 0; i < 7; ++i) win[i] = at(t0 + k0 + i);
          for (int i0 = 0; i0 < B && t0 + i0 < niter; i0 += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) win[7 + i] = at(t0 + i0 + k0 + 7 + i);
            T own[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) own[i] = (i0 + i < B) ? at(t0 + i0 + i) : (T)0;   // the block ends at B: the next lane owns the rest
            T blk[BG_LAGS];
#pragma unroll
            for (int kk = 0; kk < BG_LAGS; ++kk) {
              T sacc = (T)0;
#pragma unroll
              for (int i = 0; i < 8; ++i) sacc = fma(own[i], win[i + kk], sacc);
              blk[kk] = sacc;
            }
#pragma unroll
            for (int kk = 0; kk < BG_LAGS; ++kk) acc[kk] += (double)blk[kk];
#pragma unroll
            for (int i = 0; i < 7; ++i) win[i] = win[8 + i];
          }
          const double tot = warp_reduce8(acc);
          if ((lane & 3) == 0) part[j * BG_LAGS + (lane >> 2)] = tot;
          if (a.method == 2) {
            // sums of squares of the first / last k values of the chain for the lags of this batch, continued from the
            // running sums over the first / last `have` values
            double head = hsum[j], tail = tsum[j];
            if (lane < BG_LAGS) {
              for (int t = have; t <= have + lane && t < niter; ++t) {
                const double u = (double)p[t], v = (double)p[niter - 1 - t];
                head = fma(u, u, head); tail = fma(v, v, tail);
              }
              headp[j * BG_LAGS + lane] = head; tailp[j * BG_LAGS + lane] = tail;
            }
            __syncwarp();
            if (lane == BG_LAGS - 1) { hsum[j] = head; tsum[j] = tail; }
          }
        }
        have += BG_LAGS;
        __syncthreads();
        if (w == 0) {
          if (tid < BG_LAGS) {
            const int k = k0 + tid;
            if (k <= maxlag && k < niter) {
              T gk;
              if (a.method == 2) {
                // BDA (ess_rhat.jl:197-213): mean(chain_var) - mean_j sum_{t<=n-k}(y_t - y_{t+k})^2 / (2 (n - k))
                double sv = 0.0;
                for (int j = 0; j < nch; ++j)
                  sv += (double)(T)(2.0 * qsum[j] - headp[j * BG_LAGS + tid] - tailp[j * BG_LAGS + tid] - 2.0 * part[j * BG_LAGS + tid]);
                const T s = (T)(sv / (double)nch);
                gk = W - s / (T)(2 * (niter - k));
              } else {
                double sum = 0.0;
                for (int j = 0; j < nch; ++j) sum += part[j * BG_LAGS + tid];
                gk = (T)(sum / (double)nch) / (T)niter;
              }
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
                if (!(delta > (T)0)) { g_stage = 3; break; }
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
    }
    if (tid == 0) {
      if (a.ess_out) a.ess_out[param] = (T)ess;
      if (a.rhat_out) a.rhat_out[param] = (T)rhat;
    }
    __syncthreads();   // shared memory is free for the next parameter
  }
}

}  // namespace mcd
