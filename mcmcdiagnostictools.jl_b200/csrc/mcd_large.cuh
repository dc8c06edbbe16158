// mcd_large.cuh — global-memory pipeline for slabs that do not fit in shared memory
// (long chains, many chains).  Same programs as the slab kernel (mcd_slab.cuh), executed as
// a sequence of kernels per chunk of parameters:
//
//   segmented sort  = shared-memory tile sort (bitonic, 4096 keys) + merge-path passes
//   ranks           = per-element binary search in the sorted segment (average ties exact)
//   order stats     = direct lookups in the sorted segment (median, quantile, MCSE l/u)
//   moments         = one warp or one CTA per split chain
//   autocovariance  = one CTA per parameter, lazy lag batches (direct / BDA), or one CTA per
//                     chain running the shared-memory Stockham FFT
//
// Reference citations are /root/reference file:line, as in mcd_slab.cuh.
#pragma once
#include "mcd_common.cuh"
#include "mcd_slab.cuh"
#include "mcd_crank.cuh"
#include <algorithm>
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
  double rel_ess_max;
  int use_crank = 1;          // counting rank (mcd_crank.cuh) instead of segmented sort + binary searches where it applies
  int crank_factor = 4;       // fine buckets per value (rounded up to a power of two)
  long long crank_chunk = 0;  // cap on the parameters per chunk on the counting-rank path (0 = workspace-bound)
  long long* crank_chunks = nullptr;     // statistics: chunks ranked by counting / sent to the sort path
  long long* crank_fallbacks = nullptr;
  int fft_tc = 0;                        // developer knob: columns per tile of the four-step FFT (0 = default 2)
  int fft_full = 0;                      // developer knob: 1 = transform length nextprod(2 niter - 1) as the reference pads
  int fft_pair = 1;                      // four-step FFT: two real chains per transform + one inverse per parameter
  const void* ztab = nullptr;            // z for the doubled rank r2 at [r2 - 2] (ztab_kernel), or null: evaluate per element
};

constexpr int LG_THREADS = 256;
constexpr int SORT_TILE = 4096;     // keys per shared-memory sort tile
constexpr int MERGE_ITEMS = 8;
constexpr int MERGE_TILE = LG_THREADS * MERGE_ITEMS;  // 2048 outputs per CTA
constexpr int NAN_TILE = 1024;

// ---- segmented sort --------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) tile_sort_kernel(const T* __restrict__ src,
                                                               typename Traits<T>::Key* __restrict__ dst,
                                                               long long n, long long tiles_per_param,
                                                               int* __restrict__ nnan) {
  using Key = typename Traits<T>::Key;
  __shared__ Key sk[SORT_TILE];
  const long long param = blockIdx.x / tiles_per_param, tile = blockIdx.x % tiles_per_param;
  const long long base = param * n + tile * SORT_TILE;
  const int cnt = (int)min((long long)SORT_TILE, n - tile * SORT_TILE);
  int ln = 0;
  for (int i = threadIdx.x; i < cnt; i += LG_THREADS) {
    T v = src[base + i];
    ln += (v != v);
    sk[i] = order_key(v);
  }
  ln = __reduce_add_sync(0xffffffffu, ln);
  if ((threadIdx.x & 31) == 0 && ln) atomicAdd(&nnan[param], ln);
  __syncthreads();
  bitonic_sort_keys<Key, LG_THREADS>(sk, cnt);
  for (int i = threadIdx.x; i < cnt; i += LG_THREADS) dst[base + i] = sk[i];
}

// number of elements taken from A among the first d outputs of merge(A, B)
template <typename Key>
__device__ __forceinline__ int merge_path(const Key* A, int na, const Key* B, int nb, int d) {
  int lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (A[mid] <= B[d - 1 - mid]) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <typename Key>
__global__ void __launch_bounds__(LG_THREADS) merge_pass_kernel(const Key* __restrict__ src, Key* __restrict__ dst,
                                                                long long n, long long run,
                                                                long long tiles_per_param) {
  __shared__ Key sm[MERGE_TILE];
  __shared__ int split[2];
  const long long param = blockIdx.x / tiles_per_param, tile = blockIdx.x % tiles_per_param;
  const Key* seg = src + param * n;
  Key* out = dst + param * n;
  const long long o0 = tile * MERGE_TILE;
  const long long ps = (o0 / (2 * run)) * (2 * run);
  const long long a_end = min(ps + run, n), b_end = min(ps + 2 * run, n);
  const int na = (int)(a_end - ps), nb = (int)(b_end - a_end);
  const Key* A = seg + ps;
  const Key* B = seg + a_end;
  const int d0 = (int)(o0 - ps);
  const int d1 = (int)min((long long)(d0 + MERGE_TILE), (long long)(na + nb));
  if (threadIdx.x < 2) split[threadIdx.x] = merge_path<Key>(A, na, B, nb, threadIdx.x ? d1 : d0);
  __syncthreads();
  const int a0 = split[0], a1 = split[1], b0 = d0 - a0, b1 = d1 - a1;
  const int la = a1 - a0, lb = b1 - b0;
  for (int i = threadIdx.x; i < la; i += LG_THREADS) sm[i] = A[a0 + i];
  for (int i = threadIdx.x; i < lb; i += LG_THREADS) sm[la + i] = B[b0 + i];
  __syncthreads();
  const int total = la + lb;
  const int k0 = min(threadIdx.x * MERGE_ITEMS, total);
  int ia = merge_path<Key>(sm, la, sm + la, lb, k0);
  int ib = k0 - ia;
  Key r[MERGE_ITEMS];
#pragma unroll
  for (int i = 0; i < MERGE_ITEMS; ++i) {
    bool takeA = (ia < la) && (ib >= lb || sm[ia] <= sm[la + ib]);
    r[i] = takeA ? sm[ia] : ((ib < lb) ? sm[la + ib] : (Key)0);
    if (takeA) ++ia; else ++ib;
  }
#pragma unroll
  for (int i = 0; i < MERGE_ITEMS; ++i)
    if (k0 + i < total) out[ps + d0 + k0 + i] = r[i];
}

// ---- ranks --------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) rank_kernel(const T* __restrict__ V,
                                                          const typename Traits<T>::Key* __restrict__ sorted,
                                                          const int* __restrict__ nnan, long long n,
                                                          long long total, T* __restrict__ Yout,
                                                          double* __restrict__ ranks_out) {
  using Key = typename Traits<T>::Key;
  for (long long gid = blockIdx.x * (long long)LG_THREADS + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * LG_THREADS) {
    const long long param = gid / n;
    const T v = V[gid];
    if (v != v) continue;
    const Key* seg = sorted + param * n;
    const long long m = n - nnan[param];
    const Key key = order_key(v);
    long long lo = 0, hi = m;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (seg[mid] < key) lo = mid + 1; else hi = mid; }
    const long long lb = lo;
    long long ub = lb + 1;
    if (ub < m && seg[ub] == key) {
      lo = ub; hi = m;
      while (lo < hi) { long long mid = (lo + hi) >> 1; if (seg[mid] <= key) lo = mid + 1; else hi = mid; }
      ub = lo;
    }
    const long long r2 = lb + ub + 1;
    if (ranks_out) ranks_out[gid] = 0.5 * (double)r2;
    else Yout[gid] = z_from_rank2<T>(r2, n);
  }
}

// counting-rank variant (mcd_crank.cuh): the doubled rank comes from the bucket structure
template <typename T>
__global__ void __launch_bounds__(CR_THREADS) crank_rank_kernel(CrWork<T> w, long long n, T* __restrict__ Yout,
                                                                double* __restrict__ ranks_out,
                                                                const T* __restrict__ ztab) {
  const long long p = blockIdx.y;
  if (w.flag[p]) return;
  const long long t0 = (long long)blockIdx.x * CR_TILE;
  const long long t1 = t0 + CR_TILE < n ? t0 + CR_TILE : n;
  for (long long i0 = t0 + threadIdx.x; i0 < t1; i0 += CR_U * CR_THREADS) {
    long long r2[CR_U];
    cr_rank_body<T>(w, n, p, i0, CR_THREADS, t1, r2);
    T z[CR_U];
#pragma unroll
    for (int u = 0; u < CR_U; ++u)   // ztab[r2 - 2] = z_from_rank2(r2, n) (ztab_kernel)
      z[u] = (ranks_out || r2[u] == 0) ? (T)0 : (ztab ? __ldg(&ztab[r2[u] - 2]) : z_from_rank2<T>(r2[u], n));
#pragma unroll
    for (int u = 0; u < CR_U; ++u) {
      const long long i = i0 + (long long)u * CR_THREADS;
      if (i < t1) {
        if (ranks_out) ranks_out[p * n + i] = 0.5 * (double)r2[u];
        else Yout[p * n + i] = z[u];
      }
    }
  }
}

// NaNs rank last, each distinct, in index order: count NaNs per NAN_TILE, then rank.
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) nan_tile_count_kernel(const T* __restrict__ V, const int* __restrict__ nnan,
                                                                    long long n, long long tiles, int* __restrict__ tilecnt) {
  const long long param = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  if (nnan[param] == 0) return;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int c = 0;
  for (long long i = tile * NAN_TILE + threadIdx.x; i < min(n, (tile + 1) * NAN_TILE); i += LG_THREADS) {
    T v = V[param * n + i];
    c += (v != v);
  }
  if (c) atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tilecnt[param * tiles + tile] = cnt;
}

template <typename T>
__global__ void __launch_bounds__(LG_THREADS) nan_rank_kernel(const T* __restrict__ V, const int* __restrict__ nnan,
                                                              long long n, long long tiles, const int* __restrict__ tilecnt,
                                                              T* __restrict__ Yout, double* __restrict__ ranks_out) {
  const long long param = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int nn = nnan[param];
  if (nn == 0) return;
  __shared__ long long base;
  __shared__ int flags[NAN_TILE];
  if (threadIdx.x == 0) {
    long long b = 0;
    for (long long t = 0; t < tile; ++t) b += tilecnt[param * tiles + t];
    base = b;
  }
  const long long i0 = tile * NAN_TILE;
  const int cnt = (int)min((long long)NAN_TILE, n - i0);
  for (int i = threadIdx.x; i < NAN_TILE; i += LG_THREADS) {
    int f = 0;
    if (i < cnt) { T v = V[param * n + i0 + i]; f = (v != v); }
    flags[i] = f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // serial exclusive scan of 1024 flags (rare path)
    int run = 0;
    for (int i = 0; i < cnt; ++i) { int f = flags[i]; flags[i] = f ? run : -1; run += f; }
  }
  __syncthreads();
  const long long m = n - nn;
  for (int i = threadIdx.x; i < cnt; i += LG_THREADS) {
    if (flags[i] >= 0) {
      const long long r2 = 2 * (m + base + flags[i] + 1);
      if (ranks_out) ranks_out[param * n + i0 + i] = 0.5 * (double)r2;
      else Yout[param * n + i0 + i] = z_from_rank2<T>(r2, n);
    }
  }
}

// ---- order statistics ----------------------------------------------------------------------------
enum { SEL_MEDIAN = 0, SEL_QUANTILE = 1 };
template <typename T>
__global__ void select_kernel(const typename Traits<T>::Key* __restrict__ sorted, const int* __restrict__ nnan,
                              long long n, long long params, int what, double p, int p_f32,
                              double* __restrict__ thr, unsigned* flags) {
  const long long param = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (param >= params) return;
  const auto* seg = sorted + param * n;
  if (nnan[param] > 0) {
    thr[param] = CUDART_NAN;
    if (what == SEL_QUANTILE && flags) atomicOr(flags, FLAG_NAN_QUANTILE);
    return;
  }
  if (what == SEL_MEDIAN) {
    if (n & 1) thr[param] = (double)key_value(seg[n / 2]);
    else {
      T a = key_value(seg[n / 2 - 1]), b = key_value(seg[n / 2]);
      thr[param] = (double)(a / (T)2 + b / (T)2);
    }
    return;
  }
  if (n == 1) { thr[param] = (double)key_value(seg[0]); return; }
  if (p_f32) {
    float pf = (float)p, mm = (float)(1.0 - (double)pf);
    float aleph = fmaf((float)n, pf, mm);
    long long j = (long long)truncf(aleph);
    j = j < 1 ? 1 : (j > n - 1 ? n - 1 : j);
    float g = aleph - (float)j;
    g = g < 0.f ? 0.f : (g > 1.f ? 1.f : g);
    float a = (float)key_value(seg[j - 1]), b = (float)key_value(seg[j]);
    if (isfinite(a) && isfinite(b)) thr[param] = (double)__fadd_rn(a, __fmul_rn(g, __fsub_rn(b, a)));
    else thr[param] = (double)__fadd_rn(__fmul_rn(__fsub_rn(1.f, g), a), __fmul_rn(g, b));
    return;
  }
  double aleph = fma((double)n, p, 1.0 - p);
  long long j = (long long)trunc(aleph);
  j = j < 1 ? 1 : (j > n - 1 ? n - 1 : j);
  double g = aleph - (double)j;
  g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
  double a = (double)key_value(seg[j - 1]), b = (double)key_value(seg[j]);
  if (isfinite(a) && isfinite(b)) thr[param] = __dadd_rn(a, __dmul_rn(g, __dsub_rn(b, a)));
  else thr[param] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, g), a), __dmul_rn(g, b));
}

// ---- elementwise transforms ------------------------------------------------------------------------
enum { EW_FOLD = 0, EW_INDICATOR = 1, EW_SQDEV = 2 };
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) elementwise_kernel(const T* __restrict__ X, T* __restrict__ Y,
                                                                 const double* __restrict__ thr, long long n,
                                                                 long long total, int op) {
  for (long long gid = blockIdx.x * (long long)LG_THREADS + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * LG_THREADS) {
    const double t = thr[gid / n];
    const T x = X[gid];
    T y;
    if (op == EW_FOLD) y = fabs(x - (T)t);
    else if (op == EW_INDICATOR) y = ((double)x <= t) ? (T)1 : (T)0;
    else { T d = x - (T)t; y = d * d; }
    Y[gid] = y;
  }
}

// per-parameter mean over all draws*chains values, optional second moments about it:
// out0 = mean; out1 = sum (x-mean)^2 / (n-1); or (MODE 1) out0 = mean(y), out1 = mean(y^2)
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) slab_moments_kernel(const T* __restrict__ X, long long n, int mode,
                                                                  double* __restrict__ out0, double* __restrict__ out1) {
  __shared__ double red[40];
  const long long param = blockIdx.x;
  const T* p = X + param * n;
  double s = 0.0, s2 = 0.0;
  for (long long i = threadIdx.x; i < n; i += LG_THREADS) { double v = (double)p[i]; s += v; s2 = fma(v, v, s2); }
  const double sum = block_sum<LG_THREADS>(s, red);
  if (mode == 1) {
    const double sum2 = block_sum<LG_THREADS>(s2, red);
    if (threadIdx.x == 0) { out0[param] = (double)(T)(sum / (double)n); out1[param] = (double)(T)(sum2 / (double)n); }
    return;
  }
  const T mean = (T)(sum / (double)n);
  double q = 0.0;
  for (long long i = threadIdx.x; i < n; i += LG_THREADS) { T d = p[i] - mean; q = fma((double)d, (double)d, q); }
  const double qq = block_sum<LG_THREADS>(q, red);
  if (threadIdx.x == 0) { out0[param] = (double)mean; out1[param] = (double)(T)(qq / (double)(n - 1)); }
}

// ---- split-chain moments ---------------------------------------------------------------------------
// One CTA (or one warp when WARP_PER_CHAIN) per (parameter, split chain).
template <typename T, bool WARP_PER_CHAIN>
__global__ void __launch_bounds__(LG_THREADS) chain_stats_kernel(const T* __restrict__ Y, SplitGeom g, long long params,
                                                                 T* __restrict__ cm, T* __restrict__ cv) {
  __shared__ double red[40];
  const long long nwork = params * g.nch;
  if (WARP_PER_CHAIN) {
    const int lane = threadIdx.x & 31;
    const long long wid = (blockIdx.x * (long long)LG_THREADS + threadIdx.x) >> 5;
    if (wid >= nwork) return;
    const long long param = wid / g.nch;
    const int j = (int)(wid % g.nch);
    const T* p = Y + param * (long long)g.n + g.chain_start(j);
    double s = 0.0;
    for (int t = lane; t < g.niter; t += WARP) s += (double)p[t];
    s = warp_sum(s);
    const T m = (T)(s / (double)g.niter);
    double q = 0.0;
    for (int t = lane; t < g.niter; t += WARP) { T d = p[t] - m; q = fma((double)d, (double)d, q); }
    q = warp_sum(q);
    if (lane == 0) { cm[wid] = m; cv[wid] = (T)(q / (double)(g.niter - 1)); }
  } else {
    const long long wid = blockIdx.x;
    const long long param = wid / g.nch;
    const int j = (int)(wid % g.nch);
    const T* p = Y + param * (long long)g.n + g.chain_start(j);
    double s = 0.0;
    for (int t = threadIdx.x; t < g.niter; t += LG_THREADS) s += (double)p[t];
    const T m = (T)(block_sum<LG_THREADS>(s, red) / (double)g.niter);
    double q = 0.0;
    for (int t = threadIdx.x; t < g.niter; t += LG_THREADS) { T d = p[t] - m; q = fma((double)d, (double)d, q); }
    const double qq = block_sum<LG_THREADS>(q, red);
    if (threadIdx.x == 0) { cm[wid] = m; cv[wid] = (T)(qq / (double)(g.niter - 1)); }
  }
}

// Short split chains (many-short-chains shapes: 50 draws per split chain): G lanes per chain, 32 / G chains per
// warp (a warp per chain spends its time in ten 64-bit shuffles for a handful of values per lane).
template <typename T, int G>
__global__ void __launch_bounds__(LG_THREADS) chain_stats_group_kernel(const T* __restrict__ Y, SplitGeom g, long long params,
                                                                       T* __restrict__ cm, T* __restrict__ cv) {
  const long long nwork = params * g.nch;
  const long long gid = (blockIdx.x * (long long)LG_THREADS + threadIdx.x) / G;
  const int l = threadIdx.x % G;
  const bool live = gid < nwork;
  const long long wid = live ? gid : nwork - 1;   // (every lane takes part in the shuffles)
  const long long param = wid / g.nch;
  const int j = (int)(wid % g.nch);
  const T* p = Y + param * (long long)g.n + g.chain_start(j);
  double s = 0.0;
  for (int t = l; t < g.niter; t += G) s += (double)p[t];
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const T m = (T)(s / (double)g.niter);
  double q = 0.0;
  for (int t = l; t < g.niter; t += G) { T d = p[t] - m; q = fma((double)d, (double)d, q); }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if (live && l == 0) { cm[wid] = m; cv[wid] = (T)(q / (double)(g.niter - 1)); }
}

// ---- FFT autocovariance, one CTA per PARAMETER: chains paired, spectra summed, one inverse (fft_autocov_summed) ----
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fft_param_kernel(const T* __restrict__ Y, SplitGeom g, const T* __restrict__ cm,
                                                               int N, const Cx<T>* __restrict__ tw, int maxlag,
                                                               T* __restrict__ ac) {
  extern __shared__ __align__(16) unsigned char smem_fft[];
  Cx<T>* fa = reinterpret_cast<Cx<T>*>(smem_fft);
  Cx<T>* fb = fa + N;
  T* P = reinterpret_cast<T*>(fb + N);
  const long long param = blockIdx.x;
  Cx<T>* r = fft_autocov_summed<T, LG_THREADS>(Y + param * (long long)g.n, g, cm + param * g.nch, fa, fb, P, N, tw);
  T* dst = ac + param * (long long)(maxlag + 1);
  for (int k = threadIdx.x; k <= maxlag; k += LG_THREADS) dst[k] = r[k].x;   // summed raw Re c[k]; the ratio is formed later
}

// ---- FFT autocovariance, one CTA per (parameter, chain), shared-memory Stockham ------------------
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fft_chain_kernel(const T* __restrict__ Y, SplitGeom g,
                                                               const T* __restrict__ cm, const T* __restrict__ cv,
                                                               int N, const Cx<T>* __restrict__ tw, int maxlag,
                                                               T* __restrict__ ac) {
  extern __shared__ __align__(16) unsigned char smem_fft[];
  Cx<T>* fa = reinterpret_cast<Cx<T>*>(smem_fft);
  Cx<T>* fb = fa + N;
  const long long wid = blockIdx.x;
  const long long param = wid / g.nch;
  const int j = (int)(wid % g.nch);
  const T* p = Y + param * (long long)g.n + g.chain_start(j);
  const T m = cm[wid];
  for (int t = threadIdx.x; t < N; t += LG_THREADS) {
    Cx<T> c; c.x = t < g.niter ? (T)(p[t] - m) : (T)0; c.y = (T)0;
    fa[t] = c;
  }
  __syncthreads();
  Cx<T>* f = fft_block<T, LG_THREADS>(fa, fb, N, tw, false);
  Cx<T>* o = (f == fa) ? fb : fa;
  for (int t = threadIdx.x; t < N; t += LG_THREADS) { Cx<T> c = f[t]; c.x = c.x * c.x + c.y * c.y; c.y = (T)0; f[t] = c; }
  __syncthreads();
  Cx<T>* r = fft_block<T, LG_THREADS>(f, o, N, tw, true);
  T* dst = ac + wid * (long long)(maxlag + 1);
  for (int k = threadIdx.x; k <= maxlag; k += LG_THREADS) dst[k] = r[k].x;   // raw Re c[k]; the ratio is formed later
}

// ---- four-step FFT for chains whose transform does not fit shared memory ---------------------------
// N = N1 * N2 (both 2^a 3^b).  With n = N2 n1 + n2 and k = k1 + N1 k2:
//   X[k1 + N1 k2] = sum_n2 W_N^(n2 k1) [ sum_n1 x[N2 n1 + n2] W_N1^(n1 k1) ] W_N2^(n2 k2)
// step 1 (cols_fwd): N1-point FFTs down the columns n2 (tiles of TC columns per CTA), times the
//                    twiddle W_N^(n2 k1), stored row-major B[k1][n2];
// step 2+3 (rows):   per row k1: N2-point FFT, |.|^2, inverse N2-point FFT, times W_N^(-n2 k1);
// step 4 (cols_inv): inverse N1-point FFTs down the columns, only for the columns and outputs with
//                    n = N2 n1 + n2 <= maxlag (the only lags ever read, ess_rhat.jl:181-195).
// W_N^m is the product of two table entries, m = mh * 1024 + ml.
template <typename T>
__device__ __forceinline__ Cx<T> big_twiddle(const Cx<T>* __restrict__ twh, const Cx<T>* __restrict__ twl, long long m, bool conj) {
  const Cx<T> a = twh[m >> 10], b = twl[m & 1023];
  Cx<T> r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  if (conj) r.y = -r.y;
  return r;
}

template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fft4_cols_fwd_kernel(const T* __restrict__ Y, SplitGeom g, const T* __restrict__ cm,
                                                                   int N1, int N2, int TC, const Cx<T>* __restrict__ tw1,
                                                                   const Cx<T>* __restrict__ twh, const Cx<T>* __restrict__ twl,
                                                                   Cx<T>* __restrict__ B) {
  extern __shared__ __align__(16) unsigned char smem_fft[];
  Cx<T>* buf = reinterpret_cast<Cx<T>*>(smem_fft);          // [TC][2][N1]
  const int tiles = N2 / TC;
  const long long wid = blockIdx.x / tiles;                  // chain id within the chunk
  const int col0 = (int)(blockIdx.x % tiles) * TC;
  const long long param = wid / g.nch;
  const int j = (int)(wid % g.nch);
  const T* p = Y + param * (long long)g.n + g.chain_start(j);
  const T m = cm[wid];
  for (int idx = threadIdx.x; idx < N1 * TC; idx += LG_THREADS) {
    const int n1 = idx / TC, c = idx - n1 * TC;
    const long long nn = (long long)N2 * n1 + col0 + c;
    Cx<T> v; v.x = nn < g.niter ? (T)(p[nn] - m) : (T)0; v.y = (T)0;
    buf[(2 * c) * N1 + n1] = v;
  }
  __syncthreads();
  const long long N = (long long)N1 * N2;
  Cx<T>* Bc = B + wid * N;
  for (int c = 0; c < TC; ++c) {
    Cx<T>* r = fft_block<T, LG_THREADS>(buf + (2 * c) * N1, buf + (2 * c + 1) * N1, N1, tw1, false);
    for (int k1 = threadIdx.x; k1 < N1; k1 += LG_THREADS) {
      const Cx<T> w = big_twiddle<T>(twh, twl, (long long)(col0 + c) * k1, false);
      const Cx<T> u = r[k1];
      Cx<T> o; o.x = u.x * w.x - u.y * w.y; o.y = u.x * w.y + u.y * w.x;
      Bc[(long long)k1 * N2 + col0 + c] = o;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fft4_rows_kernel(Cx<T>* __restrict__ B, int N1, int N2, const Cx<T>* __restrict__ tw2,
                                                               const Cx<T>* __restrict__ twh, const Cx<T>* __restrict__ twl) {
  extern __shared__ __align__(16) unsigned char smem_fft[];
  Cx<T>* fa = reinterpret_cast<Cx<T>*>(smem_fft);
  Cx<T>* fb = fa + N2;
  const long long wid = blockIdx.x / N1;
  const int k1 = (int)(blockIdx.x % N1);
  Cx<T>* row = B + wid * (long long)N1 * N2 + (long long)k1 * N2;
  for (int t = threadIdx.x; t < N2; t += LG_THREADS) fa[t] = row[t];
  __syncthreads();
  Cx<T>* f = fft_block<T, LG_THREADS>(fa, fb, N2, tw2, false);
  Cx<T>* o = (f == fa) ? fb : fa;
  for (int t = threadIdx.x; t < N2; t += LG_THREADS) { Cx<T> c = f[t]; c.x = c.x * c.x + c.y * c.y; c.y = (T)0; f[t] = c; }
  __syncthreads();
  Cx<T>* r = fft_block<T, LG_THREADS>(f, o, N2, tw2, true);
  for (int n2 = threadIdx.x; n2 < N2; n2 += LG_THREADS) {
    const Cx<T> w = big_twiddle<T>(twh, twl, (long long)n2 * k1, true);
    const Cx<T> u = r[n2];
    Cx<T> v; v.x = u.x * w.x - u.y * w.y; v.y = u.x * w.y + u.y * w.x;
    row[n2] = v;
  }
}

// ---- paired variant (round 2): two real chains per complex transform, ONE inverse per parameter ----------------
// The chains are real, so chains 2q and 2q+1 ride one complex transform z = a + i b, and
//   |A(k)|^2 + |B(k)|^2 = (|Z(k)|^2 + |Z(N-k)|^2) / 2.
// Only the chain-AVERAGE of the autocovariances is ever used (mean_autocov, ess_rhat.jl:181-195), and the transform
// is linear: the power spectra of all chains of a parameter are summed and inverted once.  The reference weights
// chain j by var_j / c_j[0] with c_j[0] = (niter - 1) var_j (its own lag-0 term), i.e. every chain by 1 / (niter - 1):
// mean_j(c_j[k] / c_j[0] var_j) = (sum_j c_j[k] / sum_j c_j[0]) mean_j(var_j), which ess_kernel forms from the summed
// series (scale-free; the two differ by the rounding of c_j[0] only).  Per parameter: nch / 2 forward + 1 inverse
// transforms instead of nch + nch.  With k = k1 + N1 k2 the mirror N - k of row k1 > 0 is row N1 - k1 read backwards
// (k2 -> N2 - 1 - k2); row 0 mirrors onto itself (k2 -> (N2 - k2) mod N2): a CTA takes a row and its mirror row.
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fft4p_cols_fwd_kernel(const T* __restrict__ Y, SplitGeom g, const T* __restrict__ cm,
                                                                    int npair, int N1, int N2, int TC,
                                                                    const Cx<T>* __restrict__ tw1, const Cx<T>* __restrict__ twh,
                                                                    const Cx<T>* __restrict__ twl, Cx<T>* __restrict__ B) {
  extern __shared__ __align__(16) unsigned char smem_fft[];
  Cx<T>* buf = reinterpret_cast<Cx<T>*>(smem_fft);          // [TC][2][N1]
  const int tiles = N2 / TC;
  const long long pid = blockIdx.x / tiles;                  // (parameter, pair) within the chunk
  const int col0 = (int)(blockIdx.x % tiles) * TC;
  const long long param = pid / npair;
  const int ja = 2 * (int)(pid % npair), jb = ja + 1;
  const T* pa = Y + param * (long long)g.n + g.chain_start(ja);
  const T ma = cm[param * g.nch + ja];
  const bool hasb = jb < g.nch;
  const T* pb = hasb ? Y + param * (long long)g.n + g.chain_start(jb) : pa;
  const T mb = hasb ? cm[param * g.nch + jb] : (T)0;
  for (int idx = threadIdx.x; idx < N1 * TC; idx += LG_THREADS) {
    const int n1 = idx / TC, c = idx - n1 * TC;
    const long long nn = (long long)N2 * n1 + col0 + c;
    Cx<T> v;
    v.x = nn < g.niter ? (T)(pa[nn] - ma) : (T)0;
    v.y = (hasb && nn < g.niter) ? (T)(pb[nn] - mb) : (T)0;
    buf[(2 * c) * N1 + n1] = v;
  }
  __syncthreads();
  const long long N = (long long)N1 * N2;
  Cx<T>* Bc = B + pid * N;
  for (int c = 0; c < TC; ++c) {
    Cx<T>* r = fft_block<T, LG_THREADS>(buf + (2 * c) * N1, buf + (2 * c + 1) * N1, N1, tw1, false);
    for (int k1 = threadIdx.x; k1 < N1; k1 += LG_THREADS) {
      const Cx<T> w = big_twiddle<T>(twh, twl, (long long)(col0 + c) * k1, false);
      const Cx<T> u = r[k1];
      Cx<T> o; o.x = u.x * w.x - u.y * w.y; o.y = u.x * w.y + u.y * w.x;
      Bc[(long long)k1 * N2 + col0 + c] = o;
    }
  }
}

// grid = parameters x (N1 / 2 + 1) row pairs.  Shared memory: three N2-point complex buffers + the N2 summed powers.
// The inverse rows are written over pair 0's buffer of the parameter (this CTA is the only reader of those rows).
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fft4p_rows_kernel(Cx<T>* __restrict__ B, int npair, int N1, int N2,
                                                                const Cx<T>* __restrict__ tw2, const Cx<T>* __restrict__ twh,
                                                                const Cx<T>* __restrict__ twl) {
  extern __shared__ __align__(16) unsigned char smem_fft[];
  Cx<T>* X0 = reinterpret_cast<Cx<T>*>(smem_fft);
  Cx<T>* X1 = X0 + N2;
  Cx<T>* X2 = X1 + N2;
  T* P = reinterpret_cast<T*>(X2 + N2);
  const int nrp = N1 / 2 + 1;
  const long long param = blockIdx.x / nrp;
  const int k1 = (int)(blockIdx.x % nrp);
  const int k1m = (N1 - k1) % N1;                 // the mirror row
  const bool self = k1m == k1;                    // row 0, and row N1 / 2 of an even N1
  const long long N = (long long)N1 * N2;
  for (int t = threadIdx.x; t < N2; t += LG_THREADS) P[t] = (T)0;
  for (int q = 0; q < npair; ++q) {
    const Cx<T>* Bq = B + (param * npair + q) * N;
    for (int t = threadIdx.x; t < N2; t += LG_THREADS) {
      X0[t] = Bq[(long long)k1 * N2 + t];
      if (!self) X1[t] = Bq[(long long)k1m * N2 + t];
    }
    __syncthreads();
    Cx<T>* f0 = fft_block<T, LG_THREADS>(X0, X2, N2, tw2, false);
    Cx<T>* fm = f0;
    if (!self) fm = fft_block<T, LG_THREADS>(X1, f0 == X0 ? X2 : X0, N2, tw2, false);
    for (int t = threadIdx.x; t < N2; t += LG_THREADS) {
      const int m = (k1 == 0) ? (N2 - t) % N2 : N2 - 1 - t;     // k2 of the mirror frequency in row k1m
      const Cx<T> a = f0[t], b = fm[m];
      P[t] += (T)0.5 * ((a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y));
    }
    __syncthreads();
  }
  Cx<T>* D = B + param * npair * N;
  for (int pass = 0; pass < (self ? 1 : 2); ++pass) {
    const int kr = pass == 0 ? k1 : k1m;
    // the power of row k1m at k2 is the power of row k1 at its mirror k2 (k1m > 0 here: N2 - 1 - k2)
    for (int t = threadIdx.x; t < N2; t += LG_THREADS) { Cx<T> c; c.x = pass == 0 ? P[t] : P[N2 - 1 - t]; c.y = (T)0; X0[t] = c; }
    __syncthreads();
    Cx<T>* r = fft_block<T, LG_THREADS>(X0, X1, N2, tw2, true);
    for (int n2 = threadIdx.x; n2 < N2; n2 += LG_THREADS) {
      const Cx<T> w = big_twiddle<T>(twh, twl, (long long)n2 * kr, true);
      const Cx<T> u = r[n2];
      Cx<T> v; v.x = u.x * w.x - u.y * w.y; v.y = u.x * w.y + u.y * w.x;
      D[(long long)kr * N2 + n2] = v;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(LG_THREADS) fft4_cols_inv_kernel(const Cx<T>* __restrict__ D, int N1, int N2, int TC,
                                                                   const Cx<T>* __restrict__ tw1, int maxlag, int tiles,
                                                                   T* __restrict__ ac, long long dstride) {
  extern __shared__ __align__(16) unsigned char smem_fft[];
  Cx<T>* buf = reinterpret_cast<Cx<T>*>(smem_fft);          // [TC][2][N1]
  const long long wid = blockIdx.x / tiles;
  const int col0 = (int)(blockIdx.x % tiles) * TC;
  const Cx<T>* Dc = D + wid * dstride;   // dstride = N1 N2 (a buffer per chain) or npair N1 N2 (one per parameter)
  for (int idx = threadIdx.x; idx < N1 * TC; idx += LG_THREADS) {
    const int k1 = idx / TC, c = idx - k1 * TC;
    buf[(2 * c) * N1 + k1] = Dc[(long long)k1 * N2 + col0 + c];
  }
  __syncthreads();
  T* dst = ac + wid * (long long)(maxlag + 1);
  for (int c = 0; c < TC; ++c) {
    Cx<T>* r = fft_block<T, LG_THREADS>(buf + (2 * c) * N1, buf + (2 * c + 1) * N1, N1, tw1, true);
    for (int n1 = threadIdx.x; n1 < N1; n1 += LG_THREADS) {
      const long long nn = (long long)N2 * n1 + col0 + c;
      if (nn <= maxlag) dst[nn] = r[n1].x;
    }
  }
}

// exp(-2 pi i j stride / N) for j < count
template <typename T> __global__ void twiddle_stride_kernel(Cx<T>* tw, long long N, long long stride, int count) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) {
    double s, c;
    const long long m = ((long long)k * stride) % N;
    sincospi(-2.0 * (double)m / (double)N, &s, &c);
    Cx<T> w; w.x = (T)c; w.y = (T)s;
    tw[k] = w;
  }
}

// ---- R-hat / ESS per parameter -----------------------------------------------------------------------
template <typename T>
struct EssArgs {
  const T* Y; SplitGeom g; long long params;
  const T* cm; const T* cv;
  int want_ess, method, maxlag, relative, ess_nan;
  T rel_ess_max;
  T* gam;          // [params][maxlag + 1 + LAG_BATCH] scratch
  const T* ac;     // FFT: [params][nch][maxlag+1] raw Re c[k] per chain, or (ac_summed) [params][maxlag+1] summed over the chains
  int ac_summed = 0;
  double* r_ess; double* r_rhat;
};

template <typename T>
__global__ void __launch_bounds__(LG_THREADS) ess_kernel(const EssArgs<T> a) {
  __shared__ double red[40];
  __shared__ double part[LAG_BATCH];
  const SplitGeom& g = a.g;
  const long long param = blockIdx.x;
  const T* cm = a.cm + param * g.nch;
  const T* cv = a.cv + param * g.nch;
  // W, var_plus with block reductions over the chains
  double sw = 0.0, sm = 0.0;
  for (int j = threadIdx.x; j < g.nch; j += LG_THREADS) { sw += (double)cv[j]; sm += (double)cm[j]; }
  const T W = (T)(block_sum<LG_THREADS>(sw, red) / (double)g.nch);
  const T mm = (T)(block_sum<LG_THREADS>(sm, red) / (double)g.nch);
  double sb = 0.0;
  for (int j = threadIdx.x; j < g.nch; j += LG_THREADS) { T d = cm[j] - mm; sb = fma((double)d, (double)d, sb); }
  const T bvar = (T)(block_sum<LG_THREADS>(sb, red) / (double)(g.nch - (g.nch > 1 ? 1 : 0)));
  const T cf = (T)(g.niter - 1) / (T)g.niter;
  const T var_plus = cf * W + bvar;
  if (threadIdx.x == 0) a.r_rhat[param] = (double)sqrt(var_plus / W);
  if (!a.want_ess) return;
  if (a.ess_nan) { if (threadIdx.x == 0) a.r_ess[param] = (double)Traits<T>::nan(); return; }

  const int maxlag = a.maxlag, niter = g.niter;
  T* gamma = a.gam + param * (long long)(maxlag + 1 + LAG_BATCH);
  const T* Yp = a.Y + param * (long long)g.n;
  int have = 0;
  if (a.method == 1 && a.ac_summed) {
    // one series per parameter: sum over the chains of the raw Re c[k] (paired four-step FFT).  mean_i(c[k,i] /
    // c[0,i] var_i) with c[0,i] = (niter - 1) var_i is (sum_i c[k,i] / sum_i c[0,i]) mean_i(var_i)
    // (a chain whose centred values are all exactly zero has c[0,i] = 0: the reference's term is 0 / 0 = NaN)
    const T* ac = a.ac + param * (long long)(maxlag + 1);
    const T unc = (T)(niter - 1) / (T)niter;
    int deg = 0;
    for (int j = threadIdx.x; j < g.nch; j += LG_THREADS) deg |= (cv[j] == (T)0);
    deg = __syncthreads_or(deg);
    for (int k = threadIdx.x; k <= maxlag; k += LG_THREADS) gamma[k] = (deg ? Traits<T>::nan() : ac[k] / ac[0]) * W * unc;
    __syncthreads();
    have = maxlag;
  } else if (a.method == 1) {
    const T* ac = a.ac + param * (long long)g.nch * (maxlag + 1);
    const T unc = (T)(niter - 1) / (T)niter;
    for (int k = threadIdx.x; k <= maxlag; k += LG_THREADS) {
      double s = 0.0;
      // mean_i( Re c[k,i] / Re c[0,i] * var_i ) * (niter-1)/niter   (ess_rhat.jl:181-195)
      for (int j = 0; j < g.nch; ++j) {
        const T* aj = ac + (long long)j * (maxlag + 1);
        s += (double)((aj[k] / aj[0]) * cv[j]);
      }
      gamma[k] = (T)(s / (double)g.nch) * unc;
    }
    __syncthreads();
    have = maxlag;
  }
  const bool bda = a.method == 2;
  auto batch = [&](int k0) {
    double acc[LAG_BATCH];
#pragma unroll
    for (int kk = 0; kk < LAG_BATCH; ++kk) acc[kk] = 0.0;
    for (int j = 0; j < g.nch; ++j) {
      const T* p = Yp + g.chain_start(j);
      const T m = cm[j];
      for (int t = threadIdx.x; t + k0 < niter; t += LG_THREADS) {
        const T x0 = p[t] - m;
#pragma unroll
        for (int kk = 0; kk < LAG_BATCH; ++kk) {
          const int tk = t + k0 + kk;
          if (tk < niter) {
            const T x1 = p[tk] - m;
            if (bda) { T d = x0 - x1; acc[kk] = fma((double)d, (double)d, acc[kk]); }
            else acc[kk] = fma((double)x0, (double)x1, acc[kk]);
          }
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < LAG_BATCH; ++kk) {
      const double s = block_sum<LG_THREADS>(acc[kk], red);
      if (threadIdx.x == 0) part[kk] = s;
    }
    __syncthreads();
    if (threadIdx.x < LAG_BATCH) {
      const int k = k0 + threadIdx.x;
      if (k <= maxlag && k < niter) {
        const T mean_s = (T)(part[threadIdx.x] / (double)g.nch);
        gamma[k] = bda ? (T)(W - mean_s / (T)(2 * (niter - k))) : (T)(mean_s / (T)niter);
      }
    }
    __syncthreads();
  };
  auto ensure = [&](int k) { while (have < k) { batch(have + 1); have += LAG_BATCH; } };
  const T inv_var_plus = (T)1 / var_plus;
  auto rho = [&](int k) -> T { return (T)1 - inv_var_plus * (W - gamma[k]); };
  ensure(1);
  T rho_odd = rho(1), rho_even = (T)1;
  T p_t = rho_even + rho_odd, sum_p = p_t;
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
  if (maxlag > 1) { ensure(k); rho_even = rho(k); } else rho_even = (T)0;
  const T tau = jl_max<T>((T)0, (T)2 * sum_p + jl_max<T>((T)0, rho_even) - (T)1);
  T e = jl_min<T>((T)1 / tau, a.rel_ess_max);
  if (!a.relative) e *= (T)((long long)niter * g.nch);
  if (threadIdx.x == 0) a.r_ess[param] = (double)e;
}

// _rhat_nested_basic! from the per-chain moments (rhat_nested.jl:127-188); one CTA per parameter
template <typename T>
__global__ void __launch_bounds__(LG_THREADS) nested_kernel(const T* __restrict__ cmA, const T* __restrict__ cvA, SplitGeom g,
                                                            const int* __restrict__ chain_inds, int cps, int nsuper,
                                                            double* __restrict__ scratch, double* __restrict__ r_rhat) {
  __shared__ double red[40];
  const long long param = blockIdx.x;
  const T* cm = cmA + param * g.nch;
  const T* cv = cvA + param * g.nch;
  double* sc = scratch + param * 2ll * nsuper;
  const int m = cps * g.split;
  for (int k = threadIdx.x; k < nsuper; k += LG_THREADS) {
    double sm = 0.0, sv = 0.0;
    for (int i = 0; i < cps; ++i) {
      const int c = chain_inds[k * cps + i];
      for (int s = 0; s < g.split; ++s) { sm += (double)cm[c * g.split + s]; sv += (double)cv[c * g.split + s]; }
    }
    const T scm = (T)(sm / (double)m), Wk = (T)(sv / (double)m);
    double sb = 0.0;
    for (int i = 0; i < cps; ++i) {
      const int c = chain_inds[k * cps + i];
      for (int s = 0; s < g.split; ++s) { T d = cm[c * g.split + s] - scm; sb = fma((double)d, (double)d, sb); }
    }
    const T Bk = (T)(sb / (double)(m - (m > 1 ? 1 : 0)));
    sc[2 * k] = (double)scm;
    sc[2 * k + 1] = (double)(T)(Wk + Bk);
  }
  __syncthreads();
  double vw = 0.0, sm = 0.0;
  for (int k = threadIdx.x; k < nsuper; k += LG_THREADS) { vw += sc[2 * k + 1]; sm += sc[2 * k]; }
  const T var_within = (T)(block_sum<LG_THREADS>(vw, red) / (double)nsuper);
  const T mm = (T)(block_sum<LG_THREADS>(sm, red) / (double)nsuper);
  double sb = 0.0;
  for (int k = threadIdx.x; k < nsuper; k += LG_THREADS) { T d = (T)sc[2 * k] - mm; sb = fma((double)d, (double)d, sb); }
  const T var_between = (T)(block_sum<LG_THREADS>(sb, red) / (double)(nsuper - 1));
  if (threadIdx.x == 0) r_rhat[param] = (double)sqrt((T)1 + var_between / var_within);
}

// ---- combine ---------------------------------------------------------------------------------------------
template <typename T>
struct CombineArgs {
  long long params, n;
  int combine;
  const double* r_ess[MAX_STEPS];
  const double* r_rhat[MAX_STEPS];
  const double* ex0; const double* ex1;    // mcse side statistics
  const typename Traits<T>::Key* sortedX; const int* nnanX;
  double mcse_p;
  T* ess_out; T* rhat_out;
};

template <typename T>
__global__ void combine_kernel(const CombineArgs<T> a) {
  const long long param = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (param >= a.params) return;
  double ess = a.r_ess[0] ? a.r_ess[0][param] : 0.0, rhat = a.r_rhat[0] ? a.r_rhat[0][param] : 0.0;
  switch (a.combine) {
    case CB_RANK: rhat = (double)jl_max<T>((T)a.r_rhat[1][param], (T)rhat); break;
    case CB_TAIL: ess = (double)jl_min<T>((T)ess, (T)a.r_ess[1][param]); rhat = a.r_rhat[2][param]; break;
    case CB_TAIL_ESS: ess = (double)jl_min<T>((T)ess, (T)a.r_ess[1][param]); break;
    case CB_MAX_RHAT: rhat = (double)jl_max<T>((T)rhat, (T)a.r_rhat[1][param]); break;
    case CB_MCSE_MEAN: { const T sd = sqrt((T)a.ex1[param]); ess = (double)(sd / sqrt((T)ess)); break; }
    case CB_MCSE_STD: {
      const T mv = (T)a.ex0[param], m4 = (T)a.ex1[param], S = (T)ess;
      ess = (double)(sqrt((m4 / mv - mv) / S) / (T)2);
      break;
    }
    case CB_MCSE_QUANTILE: {
      const double S = ess;
      if (S != S || a.nnanX[param] > 0) { ess = (double)Traits<T>::nan(); break; }
      const double al = S * a.mcse_p + 1.0, be = S * (1.0 - a.mcse_p) + 1.0;
      const double pu = betainc_inv(al, be, 0.8413447460685429), pl = betainc_inv(al, be, 0.15865525393145705);
      long long u = (long long)ceil(pu * (double)a.n), l = (long long)floor(pl * (double)a.n);
      u = u > a.n ? a.n : (u < 1 ? 1 : u);
      l = l < 1 ? 1 : (l > a.n ? a.n : l);
      const auto* seg = a.sortedX + param * a.n;
      ess = (double)(((T)key_value(seg[u - 1]) - (T)key_value(seg[l - 1])) / (T)2);
      break;
    }
    default: break;
  }
  if (a.ess_out) a.ess_out[param] = (T)ess;
  if (a.rhat_out) a.rhat_out[param] = (T)rhat;
}

template <typename T> __global__ void twiddle_kernel_l(Cx<T>* tw, int N) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < N) {
    double s, c;
    sincospi(-2.0 * (double)k / (double)N, &s, &c);
    Cx<T> w; w.x = (T)c; w.y = (T)s;
    tw[k] = w;
  }
}

// ---- host driver --------------------------------------------------------------------------------------------
static inline long long nextprod23_l(long long n) {
  long long best = -1;
  for (long long p3 = 1;; p3 *= 3) {
    long long v = p3;
    while (v < n) v *= 2;
    if (best < 0 || v < best) best = v;
    if (p3 >= n) break;
  }
  return best;
}

#define LCU(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) { msg = std::string(#call) + " failed: " + cudaGetErrorString(e_);       \
      return e_ == cudaErrorMemoryAllocation ? -3 : -2; }                                           \
  } while (0)
#define LAUNCHED() do { (*env.launches)++; LCU(cudaGetLastError()); } while (0)

template <typename T>
static int run_large(LargeEnv& env, const T* dx, long long params, const SplitGeom& g, int nsteps,
                     const Step* steps, int combine, int method, int maxlag, int relative, int ess_nan,
                     double mcse_p, int cps, int nsuper, T* d_ess, T* d_rhat, void* d_arr, std::string& msg) {
  using Key = typename Traits<T>::Key;
  const long long n = g.n;
  const size_t ts = sizeof(T);
  bool any_ess = false, any_nested = false, needs_sort = false, needs_y = false;
  for (int s = 0; s < nsteps; ++s) {
    any_ess |= steps[s].reduce == RD_ESS_RHAT;
    any_nested |= steps[s].reduce == RD_NESTED;
    needs_sort |= steps[s].transform != TR_NONE && steps[s].transform != TR_STDPROXY;
    needs_y |= steps[s].transform != TR_NONE && steps[s].transform != TR_TIEDRANK;
  }
  needs_sort |= combine == CB_MCSE_QUANTILE;
  const bool use_fft = any_ess && method == 1 && !ess_nan;
  // Transform length.  The reference pads to nextprod([2, 3], 2 niter - 1) (src/ess_rhat.jl:103-118), which makes the
  // circular correlation linear for EVERY lag; only lags <= maxlag are ever read (:181-195), and those are free of
  // wrap-around as soon as N >= niter + maxlag (a term x_t x_{t+k-N} needs t >= N - k >= niter).  For long chains
  // (C3: niter = 5e5, maxlag = 250) this halves N: 2^19 instead of 2^20.  Same quantity, different rounding (1e-15).
  const long long fft_need = env.fft_full ? 2ll * g.niter - 1 : std::min<long long>(2ll * g.niter - 1, (long long)g.niter + maxlag);
  const long long fftN = use_fft ? nextprod23_l(fft_need) : 0;
  // FFT plan: one CTA per chain while two N-point complex buffers fit shared memory, else four-step
  const bool fft_big = use_fft && (size_t)fftN * 4 * ts > (size_t)env.smem_optin - 1024;
  long long fN1 = 0, fN2 = 0;
  int fTC = 1;
  if (fft_big) {
    const long long cap2 = ((long long)env.smem_optin - 1024) / (long long)(4 * ts);   // row FFT: 2 buffers of N2
    long long best = 0;
    for (long long a3 = 1; a3 <= fftN; a3 *= 3) {
      if (fftN % a3) break;
      for (long long f = a3; f <= fftN; f *= 2) {
        if (fftN % f) break;
        const long long other = fftN / f;
        if (f > other || other > cap2 || (fftN >> 10) + 1 > (1ll << 31)) continue;
        if (f > best) best = f;
      }
    }
    if (best == 0) { msg = "FFTAutocovMethod: FFT length " + std::to_string(fftN) + " is beyond the four-step plan of this build"; return -4; }
    fN1 = best; fN2 = fftN / best;
    const long long capc = ((long long)env.smem_optin - 1024) / (long long)(4 * ts * fN1);  // column tiles: TC * 2 buffers of N1
    // columns per tile: 2 keeps every 32-byte sector of a strided column read fully used (Float64) and leaves room for
    // three CTAs per SM; 4 (one CTA per SM) measured 12 % of the warps active (profiles/r2_fft_*: latency-bound)
    const int tc_max = env.fft_tc > 0 ? env.fft_tc : 2;
    for (int tc : {4, 3, 2, 1}) if (tc <= tc_max && tc <= capc && fN2 % tc == 0) { fTC = tc; break; }
    if (capc < 1) { msg = "FFTAutocovMethod: column FFT does not fit shared memory"; return -4; }
  }
  // paired variant: three N2-point complex buffers + N2 powers per CTA of the row kernel
  const bool fft_pair = fft_big && env.fft_pair && (size_t)fN2 * 7 * ts + 1024 <= (size_t)env.smem_optin;
  const int npair = (g.nch + 1) / 2;
  // chains that fit shared memory: one CTA per parameter with paired chains and one inverse, if N reals more fit
  const bool fft_smem_pair = use_fft && !fft_big && env.fft_pair && (size_t)fftN * 5 * ts + 1024 <= (size_t)env.smem_optin;
  const long long fft_bufs = fft_pair ? npair : g.nch;   // N-point complex buffers per parameter
  const long long nan_tiles = (n + NAN_TILE - 1) / NAN_TILE;
  const long long gam_stride = maxlag + 1 + LAG_BATCH;
  // counting rank: every use of the sorted copy must be one it provides (ranks, median)
  bool crankable = env.use_crank && needs_sort && combine != CB_MCSE_QUANTILE && n >= 1024 && n <= (long long)CR_MAX_N;
  for (int s = 0; s < nsteps; ++s) {
    const int tr = steps[s].transform;
    crankable &= tr == TR_NONE || tr == TR_STDPROXY || tr == TR_RANKNORM || tr == TR_TIEDRANK || tr == TR_FOLD ||
                 tr == TR_FOLD_RANKNORM;
  }
  unsigned cr_buckets = 1u << 16;
  while ((long long)cr_buckets < (long long)std::max(1, env.crank_factor) * n && cr_buckets < (1u << 28)) cr_buckets <<= 1;
  const long long cr_nw = cr_buckets / 8;

  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  // bytes per parameter of workspace
  size_t per = 0;
  if (needs_y) per += al((size_t)n * ts);
  if (needs_sort) per += 2 * al((size_t)n * ts) + al((size_t)nan_tiles * 4);
  per += 2 * al((size_t)g.nch * ts);
  if (any_ess) per += al((size_t)gam_stride * ts);
  if (use_fft) per += al((size_t)g.nch * (maxlag + 1) * ts);
  if (fft_big) per += al((size_t)fft_bufs * fftN * 2 * ts);
  if (any_nested) per += al((size_t)2 * nsuper * 8);
  per += 256 * 2;  // thresholds, nnan, results (per-param scalars; generous)
  if (crankable) per += al((size_t)cr_nw * 8) + al((size_t)CR_NSEG * 4) + 64;
  long long chunk = std::max<long long>(1, env.workspace_bytes / (long long)per);
  chunk = std::min(chunk, params);
  if (crankable) {
    chunk = std::min<long long>(chunk, 65535);   // grid.y of the counting-rank kernels
    if (env.crank_chunk > 0) chunk = std::min(chunk, env.crank_chunk);
  }
  // grid limits: blocks = chunk * tiles must stay below 2^31
  const long long sort_tiles = (n + SORT_TILE - 1) / SORT_TILE, merge_tiles = (n + MERGE_TILE - 1) / MERGE_TILE;
  const long long max_tiles = std::max<long long>(std::max(sort_tiles, merge_tiles), std::max<long long>(nan_tiles, g.nch));
  const long long fft_tiles = fft_big ? (long long)g.nch * std::max(fN1, fN2 / fTC) : 0;
  chunk = std::max<long long>(1, std::min(chunk, ((1ll << 31) - 1) / std::max(max_tiles, fft_tiles)));

  const size_t scal = al((size_t)chunk * 8);
  size_t need = 0;
  auto carve = [&](size_t bytes) { size_t o = need; need += al(bytes); return o; };
  const size_t oY = carve(needs_y ? (size_t)chunk * n * ts : 0);
  const size_t oSA = carve(needs_sort ? (size_t)chunk * n * ts : 0);
  const size_t oSB = carve(needs_sort ? (size_t)chunk * n * ts : 0);
  const size_t oTC = carve(needs_sort ? (size_t)chunk * nan_tiles * 4 : 0);
  const size_t oCM = carve((size_t)chunk * g.nch * ts);
  const size_t oCV = carve((size_t)chunk * g.nch * ts);
  const size_t oGAM = carve(any_ess ? (size_t)chunk * gam_stride * ts : 0);
  const size_t oAC = carve(use_fft ? (size_t)chunk * g.nch * (maxlag + 1) * ts : 0);
  const size_t oNS = carve(any_nested ? (size_t)chunk * 2 * nsuper * 8 : 0);
  const size_t oTW = carve(use_fft ? (size_t)(fft_big ? (fN1 + fN2 + (fftN >> 10) + 1 + 1024) : fftN) * 2 * ts : 0);
  const size_t oFB = carve(fft_big ? (size_t)chunk * fft_bufs * fftN * 2 * ts : 0);
  const size_t oTHR = carve(scal), oTHR2 = carve(scal), oEX0 = carve(scal), oEX1 = carve(scal);
  const size_t oNNX = carve(scal), oNNY = carve(scal);
  const size_t oCRW = carve(crankable ? (size_t)chunk * cr_nw * 8 : 0);
  const size_t oCRP = carve(crankable ? (size_t)chunk * CR_NSEG * 4 : 0);
  const size_t oCRK0 = carve(scal), oCRK1 = carve(scal), oCRM = carve(2 * scal), oCRF = carve(scal + 256);
  size_t oRE[MAX_STEPS], oRR[MAX_STEPS];
  for (int s = 0; s < MAX_STEPS; ++s) { oRE[s] = carve(scal); oRR[s] = carve(scal); }
  (void)scal;
  if (*env.work_cap < need) {
    if (*env.work) { cudaFree(*env.work); *env.work = nullptr; *env.work_cap = 0; }
    LCU(cudaMalloc(env.work, need));
    *env.work_cap = need;
  }
  char* wb = (char*)*env.work;
  T* Y = (T*)(wb + oY);
  Key* SA = (Key*)(wb + oSA);
  Key* SB = (Key*)(wb + oSB);
  int* tilecnt = (int*)(wb + oTC);
  T* cm = (T*)(wb + oCM);
  T* cv = (T*)(wb + oCV);
  T* gam = (T*)(wb + oGAM);
  T* ac = (T*)(wb + oAC);
  double* nscr = (double*)(wb + oNS);
  Cx<T>* tw = (Cx<T>*)(wb + oTW);
  Cx<T>* fB = (Cx<T>*)(wb + oFB);
  Cx<T>* tw1 = tw; Cx<T>* tw2 = tw + fN1; Cx<T>* twh = tw2 + fN2; Cx<T>* twl = twh + (fftN >> 10) + 1;
  double* thr = (double*)(wb + oTHR);
  double* thr2 = (double*)(wb + oTHR2);
  double* ex0 = (double*)(wb + oEX0);
  double* ex1 = (double*)(wb + oEX1);
  int* nnX = (int*)(wb + oNNX);
  int* nnY = (int*)(wb + oNNY);
  cudaStream_t st = env.stream;

  if (use_fft && !fft_big) {
    twiddle_kernel_l<T><<<(unsigned)((fftN + 255) / 256), 256, 0, st>>>(tw, (int)fftN);
    LAUNCHED();
    LCU(cudaFuncSetAttribute(fft_chain_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fftN * 4 * ts)));
    if (fft_smem_pair) LCU(cudaFuncSetAttribute(fft_param_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fftN * 5 * ts)));
  }
  if (fft_big) {
    twiddle_stride_kernel<T><<<(unsigned)((fN1 + 255) / 256), 256, 0, st>>>(tw1, fN1, 1, (int)fN1); LAUNCHED();
    twiddle_stride_kernel<T><<<(unsigned)((fN2 + 255) / 256), 256, 0, st>>>(tw2, fN2, 1, (int)fN2); LAUNCHED();
    const int nh = (int)((fftN >> 10) + 1);
    twiddle_stride_kernel<T><<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(twh, fftN, 1024, nh); LAUNCHED();
    twiddle_stride_kernel<T><<<4, 256, 0, st>>>(twl, fftN, 1, 1024); LAUNCHED();
    LCU(cudaFuncSetAttribute(fft4_cols_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fTC * fN1 * 4 * ts)));
    LCU(cudaFuncSetAttribute(fft4_cols_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fTC * fN1 * 4 * ts)));
    LCU(cudaFuncSetAttribute(fft4_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fN2 * 4 * ts)));
    if (fft_pair) {
      LCU(cudaFuncSetAttribute(fft4p_cols_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fTC * fN1 * 4 * ts)));
      LCU(cudaFuncSetAttribute(fft4p_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fN2 * 7 * ts)));
    }
  }
  const T rel_ess_max = (T)env.rel_ess_max;
  const int ew_blocks_cap = env.sm_count * 16;

  for (long long done = 0; done < params; done += chunk) {
    const long long pc = std::min(chunk, params - done);
    const T* X = dx + done * n;
    const long long total = pc * n;
    const unsigned ew_blocks = (unsigned)std::min<long long>((total + LG_THREADS - 1) / LG_THREADS, ew_blocks_cap);
    const Key* sortedX = nullptr;  // non-null while SA/SB holds the sorted copy of X

    // sort V (pc segments of n) ; returns pointer to the sorted keys, NaN counts in nn
    auto seg_sort = [&](const T* V, int* nn, const Key** out) -> int {
      LCU(cudaMemsetAsync(nn, 0, (size_t)pc * sizeof(int), st));
      tile_sort_kernel<T><<<(unsigned)(pc * sort_tiles), LG_THREADS, 0, st>>>(V, SA, n, sort_tiles, nn);
      LAUNCHED();
      Key* src = SA; Key* dst = SB;
      for (long long run = SORT_TILE; run < n; run *= 2) {
        merge_pass_kernel<Key><<<(unsigned)(pc * merge_tiles), LG_THREADS, 0, st>>>(src, dst, n, run, merge_tiles);
        LAUNCHED();
        std::swap(src, dst);
      }
      *out = src;
      return 0;
    };
    auto rank_of = [&](const T* V, const Key* sorted, const int* nn, T* Yout, double* ranks_out) -> int {
      rank_kernel<T><<<ew_blocks, LG_THREADS, 0, st>>>(V, sorted, nn, n, total, Yout, ranks_out);
      LAUNCHED();
      nan_tile_count_kernel<T><<<(unsigned)(pc * nan_tiles), LG_THREADS, 0, st>>>(V, nn, n, nan_tiles, tilecnt);
      LAUNCHED();
      nan_rank_kernel<T><<<(unsigned)(pc * nan_tiles), LG_THREADS, 0, st>>>(V, nn, n, nan_tiles, tilecnt, Yout, ranks_out);
      LAUNCHED();
      return 0;
    };
    auto ensure_sortedX = [&]() -> int {
      if (sortedX) return 0;
      return seg_sort(X, nnX, &sortedX);
    };
    // counting rank of V (pc segments of n): 1 = done (Yout / ranks_out written when do_rank; the structures stay
    // valid for crank_median_kernel until the next call), 0 = a slab of the chunk needs the sort path, < 0 = error
    CrWork<T> cw;
    cw.kmin = (typename CrKeyOf<T>::type*)(wb + oCRK0); cw.kmax = (typename CrKeyOf<T>::type*)(wb + oCRK1);
    cw.map = (CrMap<T>*)(wb + oCRM); cw.flag = (int*)(wb + oCRF); cw.cw = (uint2*)(wb + oCRW);
    cw.part = (unsigned*)(wb + oCRP); cw.info = (unsigned*)SB; cw.srt = (T*)SA;
    cw.nw = cr_nw; cw.buckets = cr_buckets; cw.seg = (cr_nw + CR_NSEG - 1) / CR_NSEG;
    auto crank = [&](const T* V, T* Yout, double* ranks_out, bool do_rank) -> int {
      using CK = typename CrKeyOf<T>::type;
      LCU(cudaMemsetAsync(cw.kmin, 0xff, (size_t)pc * sizeof(CK), st));
      LCU(cudaMemsetAsync(cw.kmax, 0, (size_t)pc * sizeof(CK), st));
      LCU(cudaMemsetAsync(cw.flag, 0, (size_t)(pc + 1) * sizeof(int), st));
      LCU(cudaMemsetAsync(cw.cw, 0, (size_t)pc * cr_nw * 8, st));
      const dim3 grid((unsigned)((n + CR_TILE - 1) / CR_TILE), (unsigned)pc);
      const unsigned pb = (unsigned)((pc + 127) / 128);
      crank_minmax_kernel<T><<<grid, CR_THREADS, 0, st>>>(cw, V, n); LAUNCHED();
      crank_setup_kernel<T><<<pb, 128, 0, st>>>(cw, pc); LAUNCHED();
      crank_count_kernel<T><<<grid, CR_THREADS, 0, st>>>(cw, V, n); LAUNCHED();
      const unsigned sb = (unsigned)((pc * CR_NSEG * 32 + CR_THREADS - 1) / CR_THREADS);
      crank_scan1_kernel<T><<<sb, CR_THREADS, 0, st>>>(cw, pc); LAUNCHED();
      crank_scan3_kernel<T><<<sb, CR_THREADS, 0, st>>>(cw, pc); LAUNCHED();
      crank_anyflag_kernel<T><<<pb, 128, 0, st>>>(cw, pc); LAUNCHED();
      int any = 0;
      LCU(cudaMemcpyAsync(&any, cw.flag + pc, sizeof(int), cudaMemcpyDeviceToHost, st));
      LCU(cudaStreamSynchronize(st));
      if (any) { if (env.crank_fallbacks) ++*env.crank_fallbacks; return 0; }
      crank_place_kernel<T><<<grid, CR_THREADS, 0, st>>>(cw, V, n); LAUNCHED();
      if (do_rank) { crank_rank_kernel<T><<<grid, CR_THREADS, 0, st>>>(cw, n, Yout, ranks_out, (const T*)env.ztab); LAUNCHED(); }
      if (env.crank_chunks) ++*env.crank_chunks;
      return 1;
    };
    bool crankX = false;            // the counting structures describe X
    bool sortX = !crankable;        // X goes (or went) through the sort path
    const unsigned pblocks = (unsigned)((pc + 127) / 128);

    for (int s = 0; s < nsteps; ++s) {
      const Step stp = steps[s];
      const T* proxy = X;
      int rc = 0;
      switch (stp.transform) {
        case TR_NONE: break;
        case TR_RANKNORM: case TR_TIEDRANK: {
          T* yo = stp.transform == TR_RANKNORM ? Y : nullptr;
          double* ro = stp.transform == TR_RANKNORM ? nullptr : (double*)d_arr + done * n;
          if (stp.transform == TR_RANKNORM) proxy = Y;
          if (!sortX) {
            const int cr = crank(X, yo, ro, true);
            if (cr < 0) return cr;
            if (cr == 1) { crankX = true; break; }
            sortX = true;
          }
          crankX = false;
          if ((rc = ensure_sortedX())) return rc;
          if ((rc = rank_of(X, sortedX, nnX, yo, ro))) return rc;
          break;
        }
        case TR_FOLD: case TR_FOLD_RANKNORM: case TR_FOLD_IND_MEDIAN: {
          if (!sortX && !crankX) {
            const int cr = crank(X, nullptr, nullptr, false);
            if (cr < 0) return cr;
            if (cr == 1) crankX = true; else sortX = true;
          }
          if (crankX) {
            crank_median_kernel<T><<<pblocks, 128, 0, st>>>(cw, n, pc, thr);
            LAUNCHED();
          } else {
            if ((rc = ensure_sortedX())) return rc;
            select_kernel<T><<<pblocks, 128, 0, st>>>(sortedX, nnX, n, pc, SEL_MEDIAN, 0.0, 0, thr, nullptr);
            LAUNCHED();
          }
          elementwise_kernel<T><<<ew_blocks, LG_THREADS, 0, st>>>(X, Y, thr, n, total, EW_FOLD);
          LAUNCHED();
          proxy = Y;
          if (stp.transform == TR_FOLD) break;
          if (crankable && stp.transform == TR_FOLD_RANKNORM) {
            crankX = false;       // the counting structures (and the sort buffers under them) are about to be reused
            sortedX = nullptr;
            const int cr = crank(Y, Y, nullptr, true);
            if (cr < 0) return cr;
            if (cr == 1) break;
          }
          const Key* sortedY = nullptr;
          sortedX = nullptr;  // the sort buffers are about to be reused
          crankX = false;
          if ((rc = seg_sort(Y, nnY, &sortedY))) return rc;
          if (stp.transform == TR_FOLD_RANKNORM) {
            if ((rc = rank_of(Y, sortedY, nnY, Y, nullptr))) return rc;
          } else {
            select_kernel<T><<<pblocks, 128, 0, st>>>(sortedY, nnY, n, pc, SEL_MEDIAN, 0.0, 0, thr2, nullptr);
            LAUNCHED();
            elementwise_kernel<T><<<ew_blocks, LG_THREADS, 0, st>>>(Y, Y, thr2, n, total, EW_INDICATOR);
            LAUNCHED();
          }
          break;
        }
        case TR_IND_MEDIAN: case TR_IND_QUANTILE:
          if ((rc = ensure_sortedX())) return rc;
          select_kernel<T><<<pblocks, 128, 0, st>>>(sortedX, nnX, n, pc,
                                                     stp.transform == TR_IND_MEDIAN ? SEL_MEDIAN : SEL_QUANTILE, stp.p,
                                                     stp.p_f32, thr, env.flags);
          LAUNCHED();
          elementwise_kernel<T><<<ew_blocks, LG_THREADS, 0, st>>>(X, Y, thr, n, total, EW_INDICATOR);
          LAUNCHED();
          proxy = Y;
          break;
        case TR_STDPROXY:
          slab_moments_kernel<T><<<(unsigned)pc, LG_THREADS, 0, st>>>(X, n, 0, thr, thr2);
          LAUNCHED();
          elementwise_kernel<T><<<ew_blocks, LG_THREADS, 0, st>>>(X, Y, thr, n, total, EW_SQDEV);
          LAUNCHED();
          if (combine == CB_MCSE_STD) {
            slab_moments_kernel<T><<<(unsigned)pc, LG_THREADS, 0, st>>>(Y, n, 1, ex0, ex1);
            LAUNCHED();
          }
          proxy = Y;
          break;
        default: msg = "bad transform"; return -1;
      }

      double* r_ess = (double*)(wb + oRE[s]);
      double* r_rhat = (double*)(wb + oRR[s]);
      if (stp.reduce == RD_ESS_RHAT || stp.reduce == RD_RHAT || stp.reduce == RD_NESTED) {
        const long long nwork = pc * g.nch;
        if (g.niter >= 2048) {
          chain_stats_kernel<T, false><<<(unsigned)nwork, LG_THREADS, 0, st>>>(proxy, g, pc, cm, cv);
        } else if (g.niter <= 128) {
          chain_stats_group_kernel<T, 8><<<(unsigned)((nwork * 8 + LG_THREADS - 1) / LG_THREADS), LG_THREADS, 0, st>>>(proxy, g, pc, cm, cv);
        } else {
          chain_stats_kernel<T, true><<<(unsigned)((nwork * 32 + LG_THREADS - 1) / LG_THREADS), LG_THREADS, 0, st>>>(proxy, g, pc, cm, cv);
        }
        LAUNCHED();
        if (stp.reduce == RD_NESTED) {
          nested_kernel<T><<<(unsigned)pc, LG_THREADS, 0, st>>>(cm, cv, g, env.d_chain_inds, cps, nsuper, nscr, r_rhat);
          LAUNCHED();
        } else {
          const bool want_ess = stp.reduce == RD_ESS_RHAT;
          if (want_ess && use_fft && !fft_big) {
            if (fft_smem_pair) fft_param_kernel<T><<<(unsigned)pc, LG_THREADS, (size_t)fftN * 5 * ts, st>>>(proxy, g, cm, (int)fftN, tw, maxlag, ac);
            else fft_chain_kernel<T><<<(unsigned)nwork, LG_THREADS, (size_t)fftN * 4 * ts, st>>>(proxy, g, cm, cv, (int)fftN, tw, maxlag, ac);
            LAUNCHED();
          }
          if (want_ess && fft_big) {
            const int tiles_f = (int)(fN2 / fTC);
            const long long ncols = std::min<long long>(fN2, (long long)maxlag + 1);
            const int tiles_i = (int)((ncols + fTC - 1) / fTC);
            if (fft_pair) {
              fft4p_cols_fwd_kernel<T><<<(unsigned)(pc * npair * tiles_f), LG_THREADS, (size_t)fTC * fN1 * 4 * ts, st>>>(
                  proxy, g, cm, npair, (int)fN1, (int)fN2, fTC, tw1, twh, twl, fB);
              LAUNCHED();
              fft4p_rows_kernel<T><<<(unsigned)(pc * (fN1 / 2 + 1)), LG_THREADS, (size_t)fN2 * 7 * ts, st>>>(
                  fB, npair, (int)fN1, (int)fN2, tw2, twh, twl);
              LAUNCHED();
              fft4_cols_inv_kernel<T><<<(unsigned)(pc * tiles_i), LG_THREADS, (size_t)fTC * fN1 * 4 * ts, st>>>(
                  fB, (int)fN1, (int)fN2, fTC, tw1, maxlag, tiles_i, ac, (long long)npair * fftN);
              LAUNCHED();
            } else {
              fft4_cols_fwd_kernel<T><<<(unsigned)(nwork * tiles_f), LG_THREADS, (size_t)fTC * fN1 * 4 * ts, st>>>(
                  proxy, g, cm, (int)fN1, (int)fN2, fTC, tw1, twh, twl, fB);
              LAUNCHED();
              fft4_rows_kernel<T><<<(unsigned)(nwork * fN1), LG_THREADS, (size_t)fN2 * 4 * ts, st>>>(fB, (int)fN1, (int)fN2, tw2, twh, twl);
              LAUNCHED();
              fft4_cols_inv_kernel<T><<<(unsigned)(nwork * tiles_i), LG_THREADS, (size_t)fTC * fN1 * 4 * ts, st>>>(
                  fB, (int)fN1, (int)fN2, fTC, tw1, maxlag, tiles_i, ac, fftN);
              LAUNCHED();
            }
          }
          EssArgs<T> ea;
          ea.Y = proxy; ea.g = g; ea.params = pc; ea.cm = cm; ea.cv = cv;
          ea.want_ess = want_ess; ea.method = method; ea.maxlag = maxlag; ea.relative = relative; ea.ess_nan = ess_nan;
          ea.rel_ess_max = rel_ess_max; ea.gam = gam; ea.ac = ac; ea.r_ess = r_ess; ea.r_rhat = r_rhat;
          ea.ac_summed = want_ess && ((fft_big && fft_pair) || fft_smem_pair);
          ess_kernel<T><<<(unsigned)pc, LG_THREADS, 0, st>>>(ea);
          LAUNCHED();
        }
      } else if (stp.reduce == RD_STORE) {
        if (stp.transform != TR_TIEDRANK)
          LCU(cudaMemcpyAsync((T*)d_arr + done * n, proxy, (size_t)total * ts, cudaMemcpyDeviceToDevice, st));
      }
    }

    // combine
    if (d_ess || d_rhat) {
      CombineArgs<T> ca;
      ca.params = pc; ca.n = n; ca.combine = combine; ca.mcse_p = mcse_p;
      for (int s = 0; s < MAX_STEPS; ++s) {
        const bool has = s < nsteps && steps[s].reduce != RD_STORE && steps[s].reduce != RD_NOTHING;
        ca.r_ess[s] = has && steps[s].reduce == RD_ESS_RHAT ? (double*)(wb + oRE[s]) : nullptr;
        ca.r_rhat[s] = has ? (double*)(wb + oRR[s]) : nullptr;
      }
      ca.ex0 = ex0; ca.ex1 = ex1;
      if (combine == CB_MCSE_MEAN) {
        slab_moments_kernel<T><<<(unsigned)pc, LG_THREADS, 0, st>>>(X, n, 0, ex0, ex1);
        LAUNCHED();
      }
      if (combine == CB_MCSE_QUANTILE) { int rc = ensure_sortedX(); if (rc) return rc; }
      ca.sortedX = sortedX; ca.nnanX = nnX;
      ca.ess_out = d_ess ? d_ess + done : nullptr;
      ca.rhat_out = d_rhat ? d_rhat + done : nullptr;
      combine_kernel<T><<<(unsigned)((pc + 63) / 64), 64, 0, st>>>(ca);
      LAUNCHED();
    }
  }
  return 0;
}

#undef LCU
#undef LAUNCHED

}  // namespace mcd
