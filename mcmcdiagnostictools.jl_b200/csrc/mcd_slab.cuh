// mcd_slab.cuh — the shared-memory "slab" kernel: one CTA owns one parameter's slab
// (draws*chains values), reads it from HBM exactly once, and runs the whole
// transform -> split-chain moments -> autocovariance -> Geyer -> combine pipeline out of
// shared memory.  Used whenever a slab (plus scratch) fits in the 227 KB of an SM.
//
// Reference algorithm restated per function; citations are /root/reference file:line.
#pragma once
#include "mcd_common.cuh"

namespace mcd {

// ---- program description -----------------------------------------------------------------
enum Transform : int {
  TR_NONE = 0,          // Y = X                                  (_expectand_proxy(mean), ess_rhat.jl:629)
  TR_RANKNORM,          // Y = rank-normalise(X)                  (utils.jl:169-193)
  TR_FOLD_RANKNORM,     // Y = rank-normalise(|X - median(X)|)    (utils.jl:148-158 then :169-193)
  TR_IND_MEDIAN,        // Y = X <= median(X)                     (ess_rhat.jl:630-639)
  TR_IND_QUANTILE,      // Y = X <= quantile(X, p)                (ess_rhat.jl:647-659)
  TR_STDPROXY,          // Y = (X - mean(X))^2                    (ess_rhat.jl:640-642)
  TR_FOLD_IND_MEDIAN,   // Y = F <= median(F), F = |X - median X| (ess_rhat.jl:643-646)
  TR_FOLD,              // Y = |X - median(X)|                    (utils.jl:148-158)
  TR_TIEDRANK           // ranks as Float64 straight to arr_out   (StatsBase.tiedrank)
};
enum Reduce : int {
  RD_ESS_RHAT = 0,      // _ess_rhat_basic!  (ess_rhat.jl:488-603)
  RD_RHAT,              // _rhat_basic!      (ess_rhat.jl:362-409)
  RD_NESTED,            // _rhat_nested_basic! (rhat_nested.jl:127-188)
  RD_STORE,             // write Y to arr_out (parity checks of the transforms)
  RD_NOTHING
};
enum Combine : int {
  CB_PLAIN = 0,         // ess = r0.ess, rhat = r0.rhat
  CB_RANK,              // ess = r0.ess, rhat = max(r1.rhat, r0.rhat)       (ess_rhat.jl:617-624, 416-420)
  CB_TAIL,              // ess = min(r0.ess, r1.ess), rhat = r2.rhat       (ess_rhat.jl:301-311, 607-616)
  CB_TAIL_ESS,          // ess = min(r0.ess, r1.ess)
  CB_MAX_RHAT,          // rhat = max(r0.rhat, r1.rhat)                   (rhat_nested.jl:114-125)
  CB_MCSE_MEAN,         // std(x)/sqrt(ess)                                (mcse.jl:45-51)
  CB_MCSE_STD,          // sqrt((m4/m2 - m2)/S)/2                          (mcse.jl:52-65)
  CB_MCSE_QUANTILE      // order-statistic rule                            (mcse.jl:66-118)
};
enum ViewMode : int { VM_NONE = 0, VM_BUCKET, VM_SORTED, VM_CONST };

struct Step {
  int transform;
  int reduce;
  double p;     // quantile probability (TR_IND_QUANTILE)
  int p_f32;    // quantile arithmetic in Float32 (p has the array's Float32 type)
};

constexpr int MAX_STEPS = 3;
constexpr int LAG_BATCH = 8;

template <typename T> struct SlabArgs {
  const T* x;
  long long params;
  SplitGeom g;
  int nsteps;
  Step steps[MAX_STEPS];
  int combine;
  int method;        // MCD_AUTOCOV_*
  int maxlag;        // already clamped to niter - 4
  int relative;
  int ess_nan;       // niter <= 4: ESS is NaN (ess_rhat.jl:472-479)
  T rel_ess_max;     // log10(T(ntotal))
  T* ess_out;
  T* rhat_out;
  void* arr_out;
  const T* ztab;     // 2n-1 entries: z for doubled rank r2 at [r2 - 2]
  int nbuckets;      // power of two, >= THREADS
  int bucket_limit;
  int fft_n;
  const void* twiddle;  // fft_n entries exp(-2 pi i k / N), complex<T>
  const int* chain_inds;
  int cps, nsuper;
  double mcse_p;
  unsigned* flags;
  const int* redo_list;   // when set: recompute only these parameters (fast-kernel fallbacks)
  const int* redo_count;
  // shared-memory byte offsets
  int offX, offY, offK, offCNT, offCM, offCV, offGAM, offGSUM, offPART, offFFT, offMISC;
};

struct ViewState {
  int mode;
  int nnan;
  int maxcnt;
  double vmin;
  double scale;
};

struct Misc {
  double red[40];       // block_sum scratch
  double thr[4];        // thresholds / order statistics
  ViewState vs;
  double wred_min[32];
  double wred_max[32];
  int wred_i[32];
  int ibuf[8];
};

template <typename T> struct Cx { T x, y; };

// ---- view (bucketised / sorted copy of an array) -------------------------------------------
template <typename T> __device__ __forceinline__ int bucket_of(T v, double vmin, double scale, int B) {
  // monotone non-decreasing in v: subtraction and scaling by a positive constant are
  // monotone under rounding, truncation is monotone, clamping is monotone.
  T u = (v - (T)vmin) * (T)scale;
  int b = (int)u;  // cvt.rzi: NaN -> 0
  b = b < 0 ? 0 : b;
  return b >= B ? B - 1 : b;
}

// Bucket counters are stored padded, one extra word per scan chunk (B / THREADS buckets), so
// that the per-thread sequential scan strides by an odd number of words (conflict-free).
__device__ __forceinline__ int cidx(int b, int per_shift) { return b + (b >> per_shift); }

// Ascending bitonic network with virtual +inf padding: every comparator puts the smaller
// key at the lower index, so comparators whose upper index is >= n are no-ops.
template <typename Key, int THREADS>
__device__ void bitonic_sort_keys(Key* K, int n) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int k = 2; k <= np2; k <<= 1) {
    // first stage of each block: compare i with its mirror inside the block
    for (int i = threadIdx.x; i < np2 / 2; i += THREADS) {
      int blk = i / (k / 2), off = i % (k / 2);
      int lo = blk * k + off, hi = blk * k + (k - 1 - off);
      if (hi < n) {
        Key a = K[lo], b = K[hi];
        if (a > b) { K[lo] = b; K[hi] = a; }
      }
    }
    __syncthreads();
    for (int j = k / 4; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2 / 2; i += THREADS) {
        int lo = 2 * j * (i / j) + (i % j), hi = lo + j;
        if (hi < n) {
          Key a = K[lo], b = K[hi];
          if (a > b) { K[lo] = b; K[hi] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// Build the view of V[0..n) into K / CNT.  Block-wide; ends with a barrier.
template <typename T, int THREADS>
__device__ void build_view(const T* V, int n, T* K, unsigned* CNT, int B, int limit, Misc* ms,
                           unsigned* flags) {
  using Key = typename Traits<T>::Key;
  constexpr int NW = THREADS / WARP;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // 1. min / max / NaN count
  T lmin = (T)CUDART_INF, lmax = -(T)CUDART_INF;
  int lnan = 0;
  for (int i = tid; i < n; i += THREADS) {
    T v = V[i];
    if (v != v) ++lnan;
    else { lmin = v < lmin ? v : lmin; lmax = v > lmax ? v : lmax; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T a = __shfl_xor_sync(0xffffffffu, lmin, o); lmin = a < lmin ? a : lmin;
    T b = __shfl_xor_sync(0xffffffffu, lmax, o); lmax = b > lmax ? b : lmax;
    lnan += __shfl_xor_sync(0xffffffffu, lnan, o);
  }
  __syncthreads();
  if (lane == 0) { ms->wred_min[w] = (double)lmin; ms->wred_max[w] = (double)lmax; ms->wred_i[w] = lnan; }
  __syncthreads();
  if (tid == 0) {
    double mn = ms->wred_min[0], mx = ms->wred_max[0];
    int nn = ms->wred_i[0];
    for (int i = 1; i < NW; ++i) {
      mn = ms->wred_min[i] < mn ? ms->wred_min[i] : mn;
      mx = ms->wred_max[i] > mx ? ms->wred_max[i] : mx;
      nn += ms->wred_i[i];
    }
    ViewState vs;
    vs.nnan = nn; vs.vmin = mn; vs.maxcnt = 0; vs.scale = 0.0;
    if (nn > 0) vs.mode = VM_SORTED;
    else if (!(mx > mn)) vs.mode = VM_CONST;
    else {
      T range = (T)mx - (T)mn;
      T sc = (T)B / range;
      // the scaled top value must stay finite and the scale positive
      if (!(range < (T)CUDART_INF) || !(sc > (T)0) || !(sc < (T)CUDART_INF)) vs.mode = VM_SORTED;
      else { vs.mode = VM_BUCKET; vs.scale = (double)sc; }
    }
    ms->vs = vs;
  }
  __syncthreads();
  int mode = ms->vs.mode;
  if (mode == VM_BUCKET) {
    const double vmin = ms->vs.vmin, scale = ms->vs.scale;
    const int psh = 31 - __clz(B / THREADS);
    for (int i = tid; i < B + THREADS; i += THREADS) CNT[i] = 0;
    __syncthreads();
    int lmaxc = 0;
    for (int i = tid; i < n; i += THREADS) {
      int b = bucket_of<T>(V[i], vmin, scale, B);
      int old = (int)atomicAdd(&CNT[cidx(b, psh)], 1u);
      lmaxc = old + 1 > lmaxc ? old + 1 : lmaxc;
    }
    lmaxc = warp_max(lmaxc);
    if (lane == 0) ms->wred_i[w] = lmaxc;
    __syncthreads();
    if (tid == 0) {
      int m = 0;
      for (int i = 0; i < NW; ++i) m = ms->wred_i[i] > m ? ms->wred_i[i] : m;
      ms->vs.maxcnt = m;
      if (m > limit) ms->vs.mode = VM_SORTED;
    }
    __syncthreads();
    mode = ms->vs.mode;
    if (mode == VM_BUCKET) {
      // exclusive scan of CNT[0..B): thread t owns a contiguous chunk
      const int per = B / THREADS;
      unsigned* mine = CNT + tid * (per + 1);   // = CNT + cidx(tid * per)
      unsigned s = 0;
      for (int i = 0; i < per; ++i) s += mine[i];
      unsigned incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) ms->wred_i[w] = (int)incl;
      __syncthreads();
      unsigned woff = 0;
      for (int i = 0; i < w; ++i) woff += (unsigned)ms->wred_i[i];
      unsigned run = woff + incl - s;
      for (int i = 0; i < per; ++i) {
        unsigned c = mine[i];
        mine[i] = run;
        run += c;
      }
      __syncthreads();
      // scatter; afterwards CNT[b] = end of bucket b = start of bucket b+1
      for (int i = tid; i < n; i += THREADS) {
        T v = V[i];
        int b = bucket_of<T>(v, vmin, scale, B);
        unsigned pos = atomicAdd(&CNT[cidx(b, psh)], 1u);
        K[pos] = v;
      }
      __syncthreads();
      return;
    }
  }
  if (mode == VM_SORTED) {
    if (tid == 0 && flags) atomicOr(flags, FLAG_SORT_FALLBACK);
    Key* KK = reinterpret_cast<Key*>(K);
    for (int i = tid; i < n; i += THREADS) KK[i] = order_key(V[i]);
    __syncthreads();
    bitonic_sort_keys<Key, THREADS>(KK, n);
  }
  __syncthreads();
}

// doubled average rank (2*rank) of a non-NaN value v in the view
template <typename T>
__device__ __forceinline__ int view_rank2(T v, int n, const T* K, const unsigned* CNT, int B,
                                          const ViewState& vs) {
  using Key = typename Traits<T>::Key;
  if (vs.mode == VM_BUCKET) {
    const int psh = 31 - __clz(B / (int)blockDim.x);
    int b = bucket_of<T>(v, vs.vmin, vs.scale, B);
    int s = b ? (int)CNT[cidx(b - 1, psh)] : 0, e = (int)CNT[cidx(b, psh)];
    int less = 0, eq = 0;
    for (int j = s; j < e; ++j) {
      T y = K[j];
      less += (y < v);
      eq += (y == v);
    }
    return 2 * (s + less) + eq + 1;
  }
  if (vs.mode == VM_CONST) return n + 1;
  const Key* KK = reinterpret_cast<const Key*>(K);
  const Key key = order_key(v);
  const int m = n - vs.nnan;
  int lo = 0, hi = m;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (KK[mid] < key) lo = mid + 1; else hi = mid; }
  const int lb = lo;
  hi = m;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (KK[mid] <= key) lo = mid + 1; else hi = mid; }
  return lb + lo + 1;
}

// k-th (0-based) order statistic of the viewed array (no NaNs present)
template <typename T>
__device__ T view_select(int k, int n, const T* K, const unsigned* CNT, int B, const ViewState& vs) {
  using Key = typename Traits<T>::Key;
  if (vs.mode == VM_CONST) return (T)vs.vmin;
  if (vs.mode == VM_SORTED) return key_value(reinterpret_cast<const Key*>(K)[k]);
  const int psh = 31 - __clz(B / (int)blockDim.x);
  int lo = 0, hi = B - 1;  // smallest b with CNT[b] > k
  while (lo < hi) { int mid = (lo + hi) >> 1; if ((int)CNT[cidx(mid, psh)] > k) hi = mid; else lo = mid + 1; }
  const int s = lo ? (int)CNT[cidx(lo - 1, psh)] : 0, e = (int)CNT[cidx(lo, psh)];
  const int target = k - s;
  for (int j = s; j < e; ++j) {
    T y = K[j];
    int less = 0, eq = 0;
    for (int i = s; i < e; ++i) { T u = K[i]; less += (u < y); eq += (u == y); }
    if (less <= target && target < less + eq) return y;
  }
  return Traits<T>::nan();
}

// Statistics.median of the viewed array: middle(a, b) = a/2 + b/2  (thread-serial)
template <typename T>
__device__ T view_median(int n, const T* K, const unsigned* CNT, int B, const ViewState& vs) {
  if (vs.nnan > 0 || n == 0) return Traits<T>::nan();
  if (n & 1) return view_select<T>(n / 2, n, K, CNT, B, vs);
  T a = view_select<T>(n / 2 - 1, n, K, CNT, B, vs), b = view_select<T>(n / 2, n, K, CNT, B, vs);
  return a / (T)2 + b / (T)2;
}

// Statistics.quantile type 7 (alpha = beta = 1), call site ess_rhat.jl:655.  Returned as
// double; when p_f32 the arithmetic is Float32 as in Julia for a Float32 p.
template <typename T>
__device__ double view_quantile(double p, int p_f32, int n, const T* K, const unsigned* CNT, int B,
                                const ViewState& vs) {
  if (n == 1) return (double)view_select<T>(0, n, K, CNT, B, vs);
  if (p_f32) {
    float pf = (float)p;
    float m = (float)(1.0 - (double)pf);
    float aleph = fmaf((float)n, pf, m);
    int j = (int)truncf(aleph);
    j = j < 1 ? 1 : (j > n - 1 ? n - 1 : j);
    float g = aleph - (float)j;
    g = g < 0.f ? 0.f : (g > 1.f ? 1.f : g);
    float a = (float)view_select<T>(j - 1, n, K, CNT, B, vs), b = (float)view_select<T>(j, n, K, CNT, B, vs);
    if (isfinite(a) && isfinite(b)) return (double)__fadd_rn(a, __fmul_rn(g, __fsub_rn(b, a)));
    return (double)__fadd_rn(__fmul_rn(__fsub_rn(1.f, g), a), __fmul_rn(g, b));
  }
  double m = 1.0 - p;
  double aleph = fma((double)n, p, m);
  int j = (int)trunc(aleph);
  j = j < 1 ? 1 : (j > n - 1 ? n - 1 : j);
  double g = aleph - (double)j;
  g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
  double a = (double)view_select<T>(j - 1, n, K, CNT, B, vs), b = (double)view_select<T>(j, n, K, CNT, B, vs);
  if (isfinite(a) && isfinite(b)) return __dadd_rn(a, __dmul_rn(g, __dsub_rn(b, a)));
  return __dadd_rn(__dmul_rn(__dsub_rn(1.0, g), a), __dmul_rn(g, b));
}

// Y[i] = z(rank(V[i])) using the view of V.  V == Y allowed.  If ranks_out != nullptr the
// average ranks are written there as Float64 instead.  Block-wide; ends with a barrier.
template <typename T, int THREADS>
__device__ void rank_transform(const T* V, T* Y, int n, T* K, const unsigned* CNT, int B,
                               const Misc* ms, const T* __restrict__ ztab, double* ranks_out) {
  using Key = typename Traits<T>::Key;
  const ViewState vs = ms->vs;
  const int tid = threadIdx.x;
  const int m = n - vs.nnan;
  if (vs.nnan > 0) {
    // NaNs rank last, each distinct, in index order.  Record the index of the q-th NaN in
    // the (otherwise unused) tail of the sorted key array before Y may overwrite V.
    Key* KK = reinterpret_cast<Key*>(K);
    for (int i = tid; i < n; i += THREADS) {
      if (V[i] != V[i]) {
        int q = 0;
        for (int t = 0; t < i; ++t) q += (V[t] != V[t]);
        KK[m + q] = (Key)i;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += THREADS) {
    T v = V[i];
    if (v != v) continue;
    int r2 = view_rank2<T>(v, n, K, CNT, B, vs);
    if (ranks_out) ranks_out[i] = 0.5 * (double)r2;
    else Y[i] = __ldg(&ztab[r2 - 2]);
  }
  if (vs.nnan > 0) {
    // V may be Y (in place): every thread must have read its V[i] before a NaN slot is overwritten
    __syncthreads();
    const Key* KK = reinterpret_cast<const Key*>(K);
    for (int q = tid; q < vs.nnan; q += THREADS) {
      int i = (int)KK[m + q];
      int r2 = 2 * (m + q + 1);
      if (ranks_out) ranks_out[i] = 0.5 * (double)r2;
      else Y[i] = __ldg(&ztab[r2 - 2]);
    }
  }
  __syncthreads();
}

// ---- split-chain moments (ess_rhat.jl:387-406 / 525-545) ------------------------------------
template <typename T, int THREADS>
__device__ void chain_stats(const T* Y, const SplitGeom& g, T* cmean, T* cvar) {
  constexpr int NW = THREADS / WARP;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int j = w; j < g.nch; j += NW) {
    const T* p = Y + g.chain_start(j);
    double s = 0.0;
    for (int t = lane; t < g.niter; t += WARP) s += (double)p[t];
    s = warp_sum(s);
    const T m = (T)(s / (double)g.niter);
    double q = 0.0;
    for (int t = lane; t < g.niter; t += WARP) { T d = p[t] - m; q = fma((double)d, (double)d, q); }
    q = warp_sum(q);
    if (lane == 0) { cmean[j] = m; cvar[j] = (T)(q / (double)(g.niter - 1)); }
  }
  __syncthreads();
}

// W = mean(chain_var); var_plus = (niter-1)/niter * W + var(chain_mean; corrected = nch > 1)
template <typename T>
__device__ __forceinline__ void within_between(const T* cmean, const T* cvar, const SplitGeom& g, T& W,
                                               T& var_plus) {
  double sw = 0.0, sm = 0.0;
  for (int j = 0; j < g.nch; ++j) { sw += (double)cvar[j]; sm += (double)cmean[j]; }
  W = (T)(sw / (double)g.nch);
  const T mm = (T)(sm / (double)g.nch);
  double sb = 0.0;
  for (int j = 0; j < g.nch; ++j) { T d = cmean[j] - mm; sb = fma((double)d, (double)d, sb); }
  const T bvar = (T)(sb / (double)(g.nch - (g.nch > 1 ? 1 : 0)));
  const T cf = (T)(g.niter - 1) / (T)g.niter;
  var_plus = cf * W + bvar;
}

// gamma[k0 .. k0+LAG_BATCH) for the direct (ess_rhat.jl:161-179) or BDA variogram
// (ess_rhat.jl:197-213) estimator on the centred samples.  Block-wide.
template <typename T, int THREADS, bool BDA>
__device__ void lag_batch(const T* Y, const SplitGeom& g, int k0, int kmax, double* part, T* gamma,
                          T mean_chain_var, bool blocked16 = false) {
  constexpr int NW = THREADS / WARP;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double acc[LAG_BATCH];
#pragma unroll
  for (int kk = 0; kk < LAG_BATCH; ++kk) acc[kk] = 0.0;
  const int niter = g.niter;
  if (blocked16 && !BDA) {
    // Same summation order as mcd_fast.cuh (lane l owns the draws [16 l, 16 l + 16), ascending), so
    // that a slab the fast kernel hands back is computed to the same bits the fast kernel would give.
    for (int j = w; j < g.nch; j += NW) {
      const T* p = Y + g.chain_start(j);
      for (int t = 16 * lane; t < 16 * lane + 16 && t + k0 < niter; ++t) {
        const T a = p[t];
#pragma unroll
        for (int kk = 0; kk < LAG_BATCH; ++kk) {
          const int tk = t + k0 + kk;
          if (tk < niter) acc[kk] = fma((double)a, (double)p[tk], acc[kk]);
        }
      }
    }
  } else
  for (int j = w; j < g.nch; j += NW) {
    const T* p = Y + g.chain_start(j);
    for (int t = lane; t + k0 < niter; t += WARP) {
      const T a = p[t];
#pragma unroll
      for (int kk = 0; kk < LAG_BATCH; ++kk) {
        const int tk = t + k0 + kk;
        if (tk < niter) {
          const T b = p[tk];
          if (BDA) { T d = a - b; acc[kk] = fma((double)d, (double)d, acc[kk]); }
          else acc[kk] = fma((double)a, (double)b, acc[kk]);
        }
      }
    }
  }
#pragma unroll
  for (int kk = 0; kk < LAG_BATCH; ++kk) acc[kk] = warp_sum(acc[kk]);
  if (lane == 0) {
#pragma unroll
    for (int kk = 0; kk < LAG_BATCH; ++kk) part[w * LAG_BATCH + kk] = acc[kk];
  }
  __syncthreads();
  if (threadIdx.x < LAG_BATCH) {
    const int k = k0 + threadIdx.x;
    if (k <= kmax && k < niter) {
      double s = 0.0;
      for (int i = 0; i < NW; ++i) s += part[i * LAG_BATCH + threadIdx.x];
      const T mean_s = (T)(s / (double)g.nch);
      if (BDA) gamma[k] = mean_chain_var - mean_s / (T)(2 * (niter - k));
      else gamma[k] = mean_s / (T)niter;
    }
  }
  __syncthreads();
}

// Block-wide Stockham autosort FFT of length N = 2^a 3^b in shared memory.
// tw[k] = exp(-2 pi i k / N).  Returns the buffer holding the (unnormalised) result.
template <typename T, int THREADS>
__device__ Cx<T>* fft_block(Cx<T>* a, Cx<T>* b, int N, const Cx<T>* __restrict__ tw, bool inverse) {
  int Ns = 1, rem = N;
  const T sgn = inverse ? (T)-1 : (T)1;
  while (rem > 1) {
    const int R = (rem % 4 == 0) ? 4 : ((rem % 2 == 0) ? 2 : 3);
    const int M = N / R;
    const int tmul = N / (Ns * R);
    for (int j = threadIdx.x; j < M; j += THREADS) {
      const int k = j % Ns;
      const int j0 = (j - k) * R + k;
      Cx<T> v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r < R) {
          Cx<T> u = a[j + r * M];
          if (r > 0) {
            const int ti = (int)(((long long)r * k * tmul) % N);
            Cx<T> wv = tw[ti];
            wv.y *= sgn;
            Cx<T> t2;
            t2.x = u.x * wv.x - u.y * wv.y;
            t2.y = u.x * wv.y + u.y * wv.x;
            u = t2;
          }
          v[r] = u;
        }
      }
      if (R == 2) {
        Cx<T> o0 = {v[0].x + v[1].x, v[0].y + v[1].y}, o1 = {v[0].x - v[1].x, v[0].y - v[1].y};
        b[j0] = o0; b[j0 + Ns] = o1;
      } else if (R == 4) {
        // forward: multiply by -i is (x,y) -> (y,-x); inverse uses +i
        Cx<T> s02 = {v[0].x + v[2].x, v[0].y + v[2].y}, d02 = {v[0].x - v[2].x, v[0].y - v[2].y};
        Cx<T> s13 = {v[1].x + v[3].x, v[1].y + v[3].y}, d13 = {v[1].x - v[3].x, v[1].y - v[3].y};
        Cx<T> jd = {sgn * d13.y, -sgn * d13.x};  // (-i*sgn) * d13
        Cx<T> o0 = {s02.x + s13.x, s02.y + s13.y}, o2 = {s02.x - s13.x, s02.y - s13.y};
        Cx<T> o1 = {d02.x + jd.x, d02.y + jd.y}, o3 = {d02.x - jd.x, d02.y - jd.y};
        b[j0] = o0; b[j0 + Ns] = o1; b[j0 + 2 * Ns] = o2; b[j0 + 3 * Ns] = o3;
      } else {
        const T c3 = (T)-0.5, s3 = sgn * (T)-0.86602540378443864676;  // exp(-+2 pi i/3)
        Cx<T> s12 = {v[1].x + v[2].x, v[1].y + v[2].y}, d12 = {v[1].x - v[2].x, v[1].y - v[2].y};
        Cx<T> o0 = {v[0].x + s12.x, v[0].y + s12.y};
        Cx<T> m = {v[0].x + c3 * s12.x, v[0].y + c3 * s12.y};
        Cx<T> rot = {-s3 * d12.y, s3 * d12.x};  // i*s3*d12
        Cx<T> o1 = {m.x + rot.x, m.y + rot.y}, o2 = {m.x - rot.x, m.y - rot.y};
        b[j0] = o0; b[j0 + Ns] = o1; b[j0 + 2 * Ns] = o2;
      }
    }
    __syncthreads();
    Cx<T>* t = a; a = b; b = t;
    Ns *= R; rem /= R;
  }
  return a;
}

// gamma[0..maxlag] with the FFT estimator (ess_rhat.jl:130-152, 181-195) on centred samples.
// Summed raw autocorrelation of the chains of one parameter through ONE inverse transform (block-wide, in shared
// memory).  Chains are real: chains 2q and 2q+1 ride one complex transform z = a + i b, |A(k)|^2 + |B(k)|^2 =
// (|Z(k)|^2 + |Z(N-k)|^2) / 2; the power spectra of all chains are summed in P and inverted once (the transform is
// linear, and only the chain average is ever used: mean_autocov, ess_rhat.jl:181-195).  nch / 2 + 1 transforms
// instead of 2 nch.  `cm` = chain means to subtract on load, or null when Y is already centred.  fa, fb: N complex
// each; P: N reals.  Returns the buffer whose real parts are N * sum_j c_j[k] (the scale cancels in the callers'
// ratio c[k] / c[0]).
template <typename T, int THREADS>
__device__ Cx<T>* fft_autocov_summed(const T* Y, const SplitGeom& g, const T* cm, Cx<T>* fa, Cx<T>* fb, T* P, int N,
                                     const Cx<T>* tw) {
  const int tid = threadIdx.x;
  for (int t = tid; t < N; t += THREADS) P[t] = (T)0;
  for (int ja = 0; ja < g.nch; ja += 2) {
    const bool hasb = ja + 1 < g.nch;
    const T* pa = Y + g.chain_start(ja);
    const T* pb = hasb ? Y + g.chain_start(ja + 1) : pa;
    const T ma = cm ? cm[ja] : (T)0, mb = (cm && hasb) ? cm[ja + 1] : (T)0;
    for (int t = tid; t < N; t += THREADS) {
      Cx<T> c;
      c.x = t < g.niter ? (T)(pa[t] - ma) : (T)0;
      c.y = (hasb && t < g.niter) ? (T)(pb[t] - mb) : (T)0;
      fa[t] = c;
    }
    __syncthreads();
    Cx<T>* f = fft_block<T, THREADS>(fa, fb, N, tw, false);
    for (int t = tid; t < N; t += THREADS) {
      const Cx<T> u = f[t], v = f[t == 0 ? 0 : N - t];
      P[t] += (T)0.5 * ((u.x * u.x + u.y * u.y) + (v.x * v.x + v.y * v.y));
    }
    __syncthreads();
  }
  for (int t = tid; t < N; t += THREADS) { Cx<T> c; c.x = P[t]; c.y = (T)0; fa[t] = c; }
  __syncthreads();
  return fft_block<T, THREADS>(fa, fb, N, tw, true);
}

// mean_autocov of FFTAutocovMethod (ess_rhat.jl:181-195) for the centred slab Y: gamma[k] =
// mean_i(c[k,i] / c[0,i] var_i) (niter - 1) / niter.  With c[0,i] = (niter - 1) var_i (the chain's own lag-0 term)
// this is (sum_i c[k,i] / sum_i c[0,i]) W (niter - 1) / niter, W = mean_i(var_i): formed from the summed series.
// A chain whose centred values are all exactly zero (a constant chain, e.g. a tail indicator that never fires) has
// c[0,i] = 0 and makes the reference's term 0 / 0 = NaN: the summed form keeps that by checking the chain variances.
template <typename T, int THREADS>
__device__ void fft_autocov(const T* Y, const SplitGeom& g, int maxlag, T W, const T* cvar, Cx<T>* fa,
                            Cx<T>* fb, int N, const Cx<T>* tw, T* gamma) {
  const int tid = threadIdx.x;
  int deg = 0;
  for (int j = tid; j < g.nch; j += THREADS) deg |= (cvar[j] == (T)0);
  deg = __syncthreads_or(deg);
  Cx<T>* r = fft_autocov_summed<T, THREADS>(Y, g, nullptr, fa, fb, reinterpret_cast<T*>(fb + N), N, tw);
  const T unc = (T)(g.niter - 1) / (T)g.niter;
  const T c0 = deg ? (T)0 : r[0].x;
  for (int k = tid; k <= maxlag; k += THREADS) gamma[k] = (deg ? Traits<T>::nan() : r[k].x / c0) * W * unc;
  __syncthreads();
}

struct Result { double ess, rhat; };

// _rhat_basic! / _ess_rhat_basic! on the slab Y (which is centred in place when ESS is wanted).
template <typename T, int THREADS>
__device__ Result reduce_ess_rhat(T* Y, const SlabArgs<T>& a, unsigned char* smem, bool want_ess) {
  const SplitGeom& g = a.g;
  T* cmean = reinterpret_cast<T*>(smem + a.offCM);
  T* cvar = reinterpret_cast<T*>(smem + a.offCV);
  T* gamma = reinterpret_cast<T*>(smem + a.offGAM);
  double* part = reinterpret_cast<double*>(smem + a.offPART);
  constexpr int NW = THREADS / WARP;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  Result res;
  chain_stats<T, THREADS>(Y, g, cmean, cvar);
  T W, var_plus;
  within_between<T>(cmean, cvar, g, W, var_plus);
  res.rhat = (double)sqrt(var_plus / W);
  res.ess = (double)Traits<T>::nan();
  if (!want_ess || a.ess_nan) return res;

  // samples .-= chain_mean  (ess_rhat.jl:548)
  for (int j = w; j < g.nch; j += NW) {
    T* p = Y + g.chain_start(j);
    const T m = cmean[j];
    for (int t = lane; t < g.niter; t += WARP) p[t] -= m;
  }
  __syncthreads();

  const int maxlag = a.maxlag;
  int have = 0;  // lags 1..have are in gamma[]
  if (a.method == 1) {
    fft_autocov<T, THREADS>(Y, g, maxlag, W, cvar, reinterpret_cast<Cx<T>*>(smem + a.offFFT),
                            reinterpret_cast<Cx<T>*>(smem + a.offFFT) + a.fft_n, a.fft_n,
                            reinterpret_cast<const Cx<T>*>(a.twiddle), gamma);
    have = maxlag;
  }
  const T inv_var_plus = (T)1 / var_plus;
  auto ensure = [&](int k) {
    while (have < k) {
      if (a.method == 2) lag_batch<T, THREADS, true>(Y, g, have + 1, maxlag, part, gamma, W);
      else lag_batch<T, THREADS, false>(Y, g, have + 1, maxlag, part, gamma, W, a.redo_list != nullptr && g.niter <= 512);
      have += LAG_BATCH;
    }
  };
  auto rho = [&](int k) -> T { return (T)1 - inv_var_plus * (W - gamma[k]); };

  // Geyer's initial positive / monotone sequence (ess_rhat.jl:553-594)
  ensure(1);
  T rho_odd = rho(1);
  T rho_even = (T)1;
  T p_t = rho_even + rho_odd;
  T sum_p = p_t;
  int k = 2;
  while (k < maxlag - 1) {
    ensure(k + 1);
    rho_even = rho(k);
    rho_odd = rho(k + 1);
    const T delta = rho_even + rho_odd;
    if (!(delta > (T)0)) break;
    p_t = jl_min<T>(delta, p_t);
    sum_p += p_t;
    k += 2;
  }
  if (maxlag > 1) { ensure(k); rho_even = rho(k); }
  else rho_even = (T)0;
  const T tau = jl_max<T>((T)0, (T)2 * sum_p + jl_max<T>((T)0, rho_even) - (T)1);
  T e = jl_min<T>((T)1 / tau, a.rel_ess_max);
  if (!a.relative) e *= (T)(g.niter * g.nch);
  res.ess = (double)e;
  return res;
}

// _rhat_nested_basic!  (rhat_nested.jl:127-188)
template <typename T, int THREADS>
__device__ Result reduce_nested(const T* Y, const SlabArgs<T>& a, unsigned char* smem) {
  const SplitGeom& g = a.g;
  T* cmean = reinterpret_cast<T*>(smem + a.offCM);
  T* cvar = reinterpret_cast<T*>(smem + a.offCV);
  double* part = reinterpret_cast<double*>(smem + a.offPART);  // [2*nsuper] : sc_mean, Wk+Bk
  chain_stats<T, THREADS>(Y, g, cmean, cvar);
  const int m = a.cps * g.split;
  for (int k = threadIdx.x; k < a.nsuper; k += THREADS) {
    double sm = 0.0, sv = 0.0;
    for (int i = 0; i < a.cps; ++i) {
      const int c = a.chain_inds[k * a.cps + i];
      for (int s = 0; s < g.split; ++s) { sm += (double)cmean[c * g.split + s]; sv += (double)cvar[c * g.split + s]; }
    }
    const T scm = (T)(sm / (double)m);
    const T Wk = (T)(sv / (double)m);
    double sb = 0.0;
    for (int i = 0; i < a.cps; ++i) {
      const int c = a.chain_inds[k * a.cps + i];
      for (int s = 0; s < g.split; ++s) { T d = cmean[c * g.split + s] - scm; sb = fma((double)d, (double)d, sb); }
    }
    const T Bk = (T)(sb / (double)(m - (m > 1 ? 1 : 0)));
    part[2 * k] = (double)scm;
    part[2 * k + 1] = (double)(Wk + Bk);
  }
  __syncthreads();
  double vw = 0.0, sm = 0.0;
  for (int k = 0; k < a.nsuper; ++k) { vw += part[2 * k + 1]; sm += part[2 * k]; }
  const T var_within = (T)(vw / (double)a.nsuper);
  const T mm = (T)(sm / (double)a.nsuper);
  double sb = 0.0;
  for (int k = 0; k < a.nsuper; ++k) { T d = (T)part[2 * k] - mm; sb = fma((double)d, (double)d, sb); }
  const T var_between = (T)(sb / (double)(a.nsuper - 1));
  Result res;
  res.ess = (double)Traits<T>::nan();
  res.rhat = (double)sqrt((T)1 + var_between / var_within);
  __syncthreads();
  return res;
}

// ---- the kernel ---------------------------------------------------------------------------
template <typename T, int THREADS>
__device__ __forceinline__ void slab_body(const SlabArgs<T>& a) {
  extern __shared__ __align__(16) unsigned char smem[];
  T* X = reinterpret_cast<T*>(smem + a.offX);
  T* Y = reinterpret_cast<T*>(smem + a.offY);
  T* K = reinterpret_cast<T*>(smem + a.offK);
  unsigned* CNT = reinterpret_cast<unsigned*>(smem + a.offCNT);
  Misc* ms = reinterpret_cast<Misc*>(smem + a.offMISC);
  const SplitGeom& g = a.g;
  const int n = g.n, tid = threadIdx.x, B = a.nbuckets;

  const long long work = a.redo_list ? (long long)*a.redo_count : a.params;
  for (long long item = blockIdx.x; item < work; item += gridDim.x) {
    const long long param = a.redo_list ? (long long)a.redo_list[item] : item;
    const T* __restrict__ src = a.x + param * (long long)n;
    for (int i = tid; i < n; i += THREADS) X[i] = __ldg(&src[i]);
    __syncthreads();

    bool viewX = false;  // K/CNT currently hold the view of X
    Result r[MAX_STEPS];
    double extra0 = 0.0, extra1 = 0.0;  // mcse side statistics

    for (int s = 0; s < a.nsteps; ++s) {
      const Step st = a.steps[s];
      bool want_rank_store = false;
      switch (st.transform) {
        case TR_NONE:
          if (Y != X) {
            for (int i = tid; i < n; i += THREADS) Y[i] = X[i];
            __syncthreads();
          }
          break;
        case TR_RANKNORM:
        case TR_TIEDRANK:
          if (!viewX) { build_view<T, THREADS>(X, n, K, CNT, B, a.bucket_limit, ms, a.flags); viewX = true; }
          want_rank_store = (st.transform == TR_TIEDRANK);
          rank_transform<T, THREADS>(X, Y, n, K, CNT, B, ms, a.ztab,
                                     want_rank_store ? reinterpret_cast<double*>(a.arr_out) + param * (long long)n : nullptr);
          break;
        case TR_FOLD:
        case TR_FOLD_RANKNORM:
        case TR_FOLD_IND_MEDIAN: {
          if (!viewX) { build_view<T, THREADS>(X, n, K, CNT, B, a.bucket_limit, ms, a.flags); viewX = true; }
          if (tid == 0) ms->thr[0] = (double)view_median<T>(n, K, CNT, B, ms->vs);
          __syncthreads();
          const T med = (T)ms->thr[0];
          for (int i = tid; i < n; i += THREADS) Y[i] = fabs(X[i] - med);
          __syncthreads();
          if (st.transform == TR_FOLD) break;
          build_view<T, THREADS>(Y, n, K, CNT, B, a.bucket_limit, ms, a.flags);
          viewX = false;
          if (st.transform == TR_FOLD_RANKNORM) {
            rank_transform<T, THREADS>(Y, Y, n, K, CNT, B, ms, a.ztab, nullptr);
          } else {
            if (tid == 0) ms->thr[1] = (double)view_median<T>(n, K, CNT, B, ms->vs);
            __syncthreads();
            const T thr = (T)ms->thr[1];
            for (int i = tid; i < n; i += THREADS) Y[i] = (Y[i] <= thr) ? (T)1 : (T)0;
            __syncthreads();
          }
          break;
        }
        case TR_IND_MEDIAN:
        case TR_IND_QUANTILE: {
          if (!viewX) { build_view<T, THREADS>(X, n, K, CNT, B, a.bucket_limit, ms, a.flags); viewX = true; }
          if (tid == 0) {
            if (st.transform == TR_IND_MEDIAN) ms->thr[0] = (double)view_median<T>(n, K, CNT, B, ms->vs);
            else if (ms->vs.nnan > 0) { ms->thr[0] = CUDART_NAN; if (a.flags) atomicOr(a.flags, FLAG_NAN_QUANTILE); }
            else ms->thr[0] = view_quantile<T>(st.p, st.p_f32, n, K, CNT, B, ms->vs);
          }
          __syncthreads();
          const double thr = ms->thr[0];
          for (int i = tid; i < n; i += THREADS) Y[i] = ((double)X[i] <= thr) ? (T)1 : (T)0;
          __syncthreads();
          break;
        }
        case TR_STDPROXY: {
          double sacc = 0.0;
          for (int i = tid; i < n; i += THREADS) sacc += (double)X[i];
          const T mean = (T)(block_sum<THREADS>(sacc, ms->red) / (double)n);
          double s2 = 0.0, s4 = 0.0;
          for (int i = tid; i < n; i += THREADS) {
            T d = X[i] - mean;
            T pz = d * d;
            Y[i] = pz;
            s2 += (double)pz;
            s4 = fma((double)pz, (double)pz, s4);
          }
          if (a.combine == CB_MCSE_STD) {
            extra0 = block_sum<THREADS>(s2, ms->red) / (double)n;  // mean(proxy)
            extra1 = block_sum<THREADS>(s4, ms->red) / (double)n;  // mean(proxy^2)
          }
          __syncthreads();
          break;
        }
        default: break;
      }

      switch (st.reduce) {
        case RD_ESS_RHAT:
          r[s] = reduce_ess_rhat<T, THREADS>(Y, a, smem, true);
          if (a.method == 1 && a.offFFT == a.offK) viewX = false;
          break;
        case RD_RHAT:
          r[s] = reduce_ess_rhat<T, THREADS>(Y, a, smem, false);
          break;
        case RD_NESTED:
          r[s] = reduce_nested<T, THREADS>(Y, a, smem);
          break;
        case RD_STORE: {
          if (!want_rank_store) {
            T* dst = reinterpret_cast<T*>(a.arr_out) + param * (long long)n;
            for (int i = tid; i < n; i += THREADS) dst[i] = Y[i];
          }
          r[s].ess = r[s].rhat = 0.0;
          break;
        }
        default: r[s].ess = r[s].rhat = 0.0; break;
      }
    }

    // ---- combine -------------------------------------------------------------------------
    double ess = r[0].ess, rhat = r[0].rhat;
    switch (a.combine) {
      case CB_RANK: rhat = (double)jl_max<T>((T)r[1].rhat, (T)r[0].rhat); break;
      case CB_TAIL: ess = (double)jl_min<T>((T)r[0].ess, (T)r[1].ess); rhat = r[2].rhat; break;
      case CB_TAIL_ESS: ess = (double)jl_min<T>((T)r[0].ess, (T)r[1].ess); break;
      case CB_MAX_RHAT: rhat = (double)jl_max<T>((T)r[0].rhat, (T)r[1].rhat); break;
      case CB_MCSE_MEAN: {
        // std(samples; dims=(1,2)) ./ sqrt.(S)  (mcse.jl:45-51)
        double sacc = 0.0;
        for (int i = tid; i < n; i += THREADS) sacc += (double)X[i];
        const T mean = (T)(block_sum<THREADS>(sacc, ms->red) / (double)n);
        double q = 0.0;
        for (int i = tid; i < n; i += THREADS) { T d = X[i] - mean; q = fma((double)d, (double)d, q); }
        const T sd = sqrt((T)(block_sum<THREADS>(q, ms->red) / (double)(n - 1)));
        ess = (double)(sd / sqrt((T)r[0].ess));
        break;
      }
      case CB_MCSE_STD: {
        const T mv = (T)extra0, m4 = (T)extra1, S = (T)r[0].ess;
        ess = (double)(sqrt((m4 / mv - mv) / S) / (T)2);
        break;
      }
      case CB_MCSE_QUANTILE: {
        // _mcse_quantile (mcse.jl:96-118)
        const double S = r[0].ess;
        if (S != S) { ess = (double)Traits<T>::nan(); break; }
        if (!viewX) { build_view<T, THREADS>(X, n, K, CNT, B, a.bucket_limit, ms, a.flags); viewX = true; }
        if (ms->vs.nnan > 0) { ess = (double)Traits<T>::nan(); break; }
        {
          const double al = S * a.mcse_p + 1.0, be = S * (1.0 - a.mcse_p) + 1.0;
          // ms->red[32..33] <- betainvcdf(al, be, Phi(+1)), betainvcdf(al, be, Phi(-1))   (mcse.jl:105-109)
          betainc_inv_pair_block<THREADS>(al, be, 0.8413447460685429, 0.15865525393145705, ms->red, ms->red + 32);
          if (tid == 0) {
            long long u = (long long)ceil(ms->red[32] * (double)n);
            u = u > n ? n : (u < 1 ? 1 : u);
            ms->thr[2] = (double)view_select<T>((int)u - 1, n, K, CNT, B, ms->vs);
          } else if (tid == 32) {
            long long l = (long long)floor(ms->red[33] * (double)n);
            l = l < 1 ? 1 : (l > n ? n : l);
            ms->thr[3] = (double)view_select<T>((int)l - 1, n, K, CNT, B, ms->vs);
          }
        }
        __syncthreads();
        ess = (double)(((T)ms->thr[2] - (T)ms->thr[3]) / (T)2);
        break;
      }
      default: break;
    }
    if (tid == 0) {
      if (a.ess_out) a.ess_out[param] = (T)ess;
      if (a.rhat_out) a.rhat_out[param] = (T)rhat;
    }
    __syncthreads();
  }
}

// Three entry points over the same body, chosen by how many CTAs the slab lets an SM hold: <= 80 registers
// so that three or more fit (small slabs), 128 registers without spills for two, and 512 threads for slabs
// that fit one CTA per SM only (16 warps hide more latency).
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 3) slab_kernel(const SlabArgs<T> a) { slab_body<T, THREADS>(a); }
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) slab_kernel_two(const SlabArgs<T> a) { slab_body<T, THREADS>(a); }
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) slab_kernel_wide(const SlabArgs<T> a) { slab_body<T, THREADS>(a); }

}  // namespace mcd
