// mcd_crank.cuh — counting rank for slabs that do not fit shared memory (the large-slab pipeline's
// replacement for "segmented sort + per-element binary search" on NaN-free, untied-enough data).
//
// What it computes: for every element of a slab of n values its doubled average rank r2 = lb + ub + 1
// (lb = #values less, ub = #values less or equal; StatsBase.tiedrank as _rank_normalize uses it,
// /root/reference/src/utils.jl:169-193), bit-exactly, without sorting:
//
//   1. min / max of the slab (order-preserving integer keys, atomicMin / atomicMax per parameter)
//   2. count   every value increments the 4-bit counter of its fine bucket (monotone linear map, B >= 4 n buckets,
//              eight counters per 32-bit word, ONE global atomic per value that also returns the arrival offset
//              of the value inside its bucket)
//   3. scan    exclusive prefix of the populations per counter word (two coalesced sweeps, a warp per segment) (interleaved with the counters: one 8-byte
//              entry {counters, prefix} per word, so a lookup is a single sector)
//   4. place   start of the bucket = prefix + populations of the lower counters of the word; members of shared
//              buckets (and the bucket(s) holding the median) copy their value to the bucket's slots of S
//   5. rank    alone in its bucket: lb = start, ub = start + 1; else exact (less, equal) counts against the bucket
//              mates on the values themselves (== ties -0.0 with 0.0 as tiedrank's run detection does)
//   6. select  the central order statistic(s) for Statistics.median from the bucket that holds them
//
// The same idea as the headline kernel's shared-memory ranking (mcd_rk2.cuh), with global memory / L2 in place
// of shared memory.  A parameter whose slab holds NaN or +-Inf, is constant, or puts >= 15 values into one
// bucket raises its flag; the caller then runs the sort-based path for the chunk (results do not depend on the
// path: both produce the exact r2).
//
// Every per-element / per-segment body is a __host__ __device__ function so that tests/crank_emul.cu can run
// the algorithm on the CPU (thread by thread, atomics as plain updates) against numpy's ranks without a GPU;
// the __global__ wrappers only enumerate indices.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

namespace mcd {

#define CR_HD __host__ __device__ __forceinline__

constexpr int CR_THREADS = 256;
constexpr int CR_TILE = 8192;            // elements per CTA of the per-element kernels
constexpr unsigned CR_MAX_N = (1u << 24) - 1u;   // start | count << 24 | offset << 28 in one word
constexpr int CR_NSEG = 256;             // scan segments per parameter

template <typename T> struct CrKeyOf;
template <> struct CrKeyOf<double> { using type = unsigned long long; };
template <> struct CrKeyOf<float> { using type = unsigned int; };

// order-preserving integer key (NaN -> largest key; -0.0 and 0.0 share a key), and its inverse
CR_HD unsigned long long cr_key(double v) {
  if (v != v) return ~0ull;
  v = v + 0.0;
  long long b;
#ifdef __CUDA_ARCH__
  b = __double_as_longlong(v);
#else
  memcpy(&b, &v, 8);
#endif
  const unsigned long long u = (unsigned long long)b;
  return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}
CR_HD unsigned int cr_key(float v) {
  if (v != v) return ~0u;
  v = v + 0.0f;
  int b;
#ifdef __CUDA_ARCH__
  b = __float_as_int(v);
#else
  memcpy(&b, &v, 4);
#endif
  const unsigned int u = (unsigned int)b;
  return (b < 0) ? ~u : (u | 0x80000000u);
}
CR_HD double cr_unkey(unsigned long long k) {
  const unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  double v;
#ifdef __CUDA_ARCH__
  v = __longlong_as_double((long long)u);
#else
  memcpy(&v, &u, 8);
#endif
  return v;
}
CR_HD float cr_unkey(unsigned int k) {
  const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  float v;
#ifdef __CUDA_ARCH__
  v = __int_as_float((int)u);
#else
  memcpy(&v, &u, 4);
#endif
  return v;
}

CR_HD unsigned cr_nibsum(unsigned w) {
  unsigned t = (w & 0x0f0f0f0fu) + ((w >> 4) & 0x0f0f0f0fu);
  t = (t & 0x00ff00ffu) + ((t >> 8) & 0x00ff00ffu);
  return (t & 0xffffu) + (t >> 16);
}

template <typename K> CR_HD void cr_atomic_min(K* p, K v) {
#ifdef __CUDA_ARCH__
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}
template <typename K> CR_HD void cr_atomic_max(K* p, K v) {
#ifdef __CUDA_ARCH__
  atomicMax(p, v);
#else
  if (v > *p) *p = v;
#endif
}
CR_HD unsigned cr_atomic_add(unsigned* p, unsigned v) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, v);
#else
  const unsigned o = *p; *p = o + v; return o;
#endif
}

// per-parameter bucket map and status
template <typename T> struct CrMap {
  T vmin;
  T scale;
};

// workspace of one crank call over `pc` parameters (device pointers)
template <typename T> struct CrWork {
  typename CrKeyOf<T>::type* kmin;   // [pc]
  typename CrKeyOf<T>::type* kmax;   // [pc]
  CrMap<T>* map;                     // [pc]
  int* flag;                         // [pc + 1]: per-parameter "use the sort path"; [pc] = any
  uint2* cw;                         // [pc][nw]  {eight 4-bit counters, #values in earlier words}
  unsigned* part;                    // [pc][CR_NSEG] scan partials
  unsigned* info;                    // [pc][n]
  T* srt;                            // [pc][n]   bucket-ordered copy (only the slots that are needed)
  long long nw;                      // counter words per parameter (buckets / 8)
  unsigned buckets;                  // B
  long long seg;                     // words per scan segment
};

// ---- 1. min / max -----------------------------------------------------------------------------------
// one thread: elements i0, i0 + stride, ... < i1 of parameter p
template <typename T>
CR_HD void cr_minmax_body(const T* x, long long i0, long long i1, long long stride,
                          typename CrKeyOf<T>::type& kmin, typename CrKeyOf<T>::type& kmax) {
  for (long long i = i0; i < i1; i += stride) {
    const typename CrKeyOf<T>::type k = cr_key(x[i]);
    kmin = k < kmin ? k : kmin;
    kmax = k > kmax ? k : kmax;
  }
}

// one thread per parameter: the bucket map, or the flag
template <typename T>
CR_HD void cr_setup_body(const CrWork<T>& w, long long p) {
  const T vmin = cr_unkey(w.kmin[p]), vmax = cr_unkey(w.kmax[p]);
  const T inf = (T)INFINITY;
  const T range = vmax - vmin;
  // slightly less than B / range: the largest value lands inside the last bucket
  const T scale = (T)((double)w.buckets * (1.0 - 1.0 / 1048576.0)) / range;
  const bool ok = (vmin == vmin) && (vmax == vmax) && vmin > -inf && vmax < inf && vmax > vmin &&
                  range < inf && scale > (T)0 && scale < inf;
  CrMap<T> m;
  m.vmin = vmin;
  m.scale = ok ? scale : (T)0;
  w.map[p] = m;
  w.flag[p] = ok ? 0 : 1;
}

// ---- 2. count ---------------------------------------------------------------------------------------
// The per-element bodies below take a BATCH of CR_U elements (i0, i0 + stride, ... below i1): every load / atomic
// of the batch is issued before the first dependent store, so a thread keeps CR_U random accesses in flight (the
// compiler may not hoist the loads of one element over the stores of the previous one: the pointers may alias).
constexpr int CR_U = 4;   // (8 in flight measured no faster: 6.75 vs 6.64 ms on C4, P = 400)

template <typename T>
CR_HD void cr_count_body(const CrWork<T>& w, const T* x, long long n, long long p, long long i0, long long stride,
                         long long i1) {
  const CrMap<T> m = w.map[p];
  if (!(m.scale > (T)0)) return;   // flagged by the setup (NaN, Inf, constant): nothing to count
  T v[CR_U];
#pragma unroll
  for (int u = 0; u < CR_U; ++u) { const long long i = i0 + u * stride; v[u] = i < i1 ? x[p * n + i] : m.vmin; }
  unsigned fb[CR_U], old[CR_U];
#pragma unroll
  for (int u = 0; u < CR_U; ++u) {
    unsigned f = (unsigned)(long long)((v[u] - m.vmin) * m.scale);
    fb[u] = f < w.buckets ? f : w.buckets - 1u;
  }
#pragma unroll
  for (int u = 0; u < CR_U; ++u) {
    old[u] = 0;
    if (i0 + u * stride < i1) old[u] = cr_atomic_add(&w.cw[p * w.nw + (fb[u] >> 3)].x, 1u << ((fb[u] & 7u) * 4u));
  }
#pragma unroll
  for (int u = 0; u < CR_U; ++u) {
    const long long i = i0 + u * stride;
    if (i < i1) {
      const unsigned off = (old[u] >> ((fb[u] & 7u) * 4u)) & 15u;
      if (off >= 15u) w.flag[p] = 1;   // the counter reaches 16 and spills into its neighbour
      w.info[p * n + i] = fb[u] | (off << 28);
    }
  }
}

// ---- 3. scan ----------------------------------------------------------------------------------------
// Two sweeps over the counter words of a parameter, split into CR_NSEG segments.  On the device a WARP owns a
// segment (lanes read consecutive words: coalesced); these bodies state what a segment's sweep computes.
// (a) populations of the segment's words
template <typename T>
CR_HD void cr_scan1_body(const CrWork<T>& w, long long p, int s) {
  const long long w0 = (long long)s * w.seg, w1 = (w0 + w.seg < w.nw) ? w0 + w.seg : w.nw;
  const uint2* cw = w.cw + p * w.nw;
  unsigned sum = 0;
  for (long long i = w0; i < w1; ++i) sum += cr_nibsum(cw[i].x);
  w.part[p * CR_NSEG + s] = sum;
}
// (b) prefix of every word of the segment: #values in the earlier segments + in the earlier words of this one
template <typename T>
CR_HD void cr_scan3_body(const CrWork<T>& w, long long p, int s) {
  const long long w0 = (long long)s * w.seg, w1 = (w0 + w.seg < w.nw) ? w0 + w.seg : w.nw;
  uint2* cw = w.cw + p * w.nw;
  unsigned run = 0;
  for (int j = 0; j < s; ++j) run += w.part[p * CR_NSEG + j];
  for (long long i = w0; i < w1; ++i) {
    const unsigned c = cw[i].x;
    cw[i].y = run;
    run += cr_nibsum(c);
  }
}

// ---- 4. place ---------------------------------------------------------------------------------------
template <typename T>
CR_HD void cr_place_body(const CrWork<T>& w, const T* x, long long n, long long p, long long i0, long long stride,
                         long long i1) {
  unsigned info[CR_U];
  uint2 e[CR_U];
  T v[CR_U];
#pragma unroll
  for (int u = 0; u < CR_U; ++u) { const long long i = i0 + u * stride; info[u] = i < i1 ? w.info[p * n + i] : 0u; }
#pragma unroll
  for (int u = 0; u < CR_U; ++u) {
    const long long i = i0 + u * stride;
    e[u] = w.cw[p * w.nw + ((info[u] & 0x0fffffffu) >> 3)];
    v[u] = i < i1 ? x[p * n + i] : (T)0;
  }
  // the bucket's slots of S are needed when it is shared, or when it holds a central order statistic
  const unsigned mlo = (unsigned)((n - 1) >> 1), mhi = (unsigned)(n >> 1);
#pragma unroll
  for (int u = 0; u < CR_U; ++u) {
    const long long i = i0 + u * stride;
    if (i < i1) {
      const unsigned fb = info[u] & 0x0fffffffu, off = info[u] >> 28;
      const unsigned sh = (fb & 7u) * 4u;
      const unsigned st = e[u].y + cr_nibsum(e[u].x & ((1u << sh) - 1u));
      const unsigned c = (e[u].x >> sh) & 15u;
      if ((c >= 2u || st == mlo || st == mhi) && (long long)st + off < n) w.srt[p * n + st + off] = v[u];
      w.info[p * n + i] = st | (c << 24) | (off << 28);
    }
  }
}

// ---- 5. rank ----------------------------------------------------------------------------------------
// doubled average ranks r2 = lb + ub + 1 of a batch of elements (0 for the slots beyond i1)
template <typename T>
CR_HD void cr_rank_body(const CrWork<T>& w, long long n, long long p, long long i0, long long stride, long long i1,
                        long long* r2) {
  unsigned info[CR_U];
#pragma unroll
  for (int u = 0; u < CR_U; ++u) { const long long i = i0 + u * stride; info[u] = i < i1 ? w.info[p * n + i] : (1u << 24); }
#pragma unroll
  for (int u = 0; u < CR_U; ++u) {
    const unsigned st = info[u] & 0x00ffffffu, c = (info[u] >> 24) & 15u, off = info[u] >> 28;
    unsigned less = 0, eq = 1;
    if (c >= 2u) {
      const T* s = w.srt + p * n + st;
      const T v = s[off];
      eq = 0;
      for (unsigned j = 0; j < c; ++j) {
        const T y = s[j];
        less += y < v;
        eq += y == v;
      }
    }
    const long long lb = (long long)st + less;
    r2[u] = (i0 + u * stride < i1) ? lb + (lb + eq) + 1 : 0;
  }
}

// ---- 6. select --------------------------------------------------------------------------------------
// the k-th (0-based) smallest value of parameter p: find the counter word, the bucket, then the bucket mate
template <typename T>
CR_HD T cr_order_stat(const CrWork<T>& w, long long n, long long p, long long k) {
  const uint2* cw = w.cw + p * w.nw;
  // first word whose prefix exceeds k, minus one: prefix[wd] <= k < prefix[wd] + population(wd)
  long long lo = 0, hi = w.nw;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)cw[mid].y > k) hi = mid; else lo = mid + 1;
  }
  const uint2 e = cw[lo - 1];
  long long pos = e.y;
  unsigned c = 0;
  for (int b = 0; b < 8; ++b) {
    c = (e.x >> (4 * b)) & 15u;
    if (k < pos + (long long)c) break;
    pos += c;
  }
  const T* s = w.srt + p * n + pos;
  const unsigned want = (unsigned)(k - pos);
  T res = s[0];
  for (unsigned j = 0; j < c; ++j) {
    const T v = s[j];
    unsigned less = 0, eq = 0;
    for (unsigned i = 0; i < c; ++i) { less += s[i] < v; eq += s[i] == v; }
    if (less <= want && want < less + eq) { res = v; break; }
  }
  return res;
}
// Statistics.median of parameter p (no NaN): middle of the two central order statistics
template <typename T>
CR_HD T cr_median_body(const CrWork<T>& w, long long n, long long p) {
  if (n & 1) return cr_order_stat<T>(w, n, p, n / 2);
  const T a = cr_order_stat<T>(w, n, p, n / 2 - 1), b = cr_order_stat<T>(w, n, p, n / 2);
  return a / (T)2 + b / (T)2;
}

#ifdef __CUDACC__
// =====================================================================================================
// kernels: index enumeration only.  grid.y = parameter of the chunk (<= 65535), grid.x = tile.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(CR_THREADS) crank_minmax_kernel(CrWork<T> w, const T* __restrict__ x, long long n) {
  using K = typename CrKeyOf<T>::type;
  const long long p = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * CR_TILE;
  const long long t1 = t0 + CR_TILE < n ? t0 + CR_TILE : n;
  K kmin = ~(K)0, kmax = 0;
  cr_minmax_body<T>(x + p * n, t0 + threadIdx.x, t1, CR_THREADS, kmin, kmax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const K a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  if ((threadIdx.x & 31) == 0) { cr_atomic_min<K>(&w.kmin[p], kmin); cr_atomic_max<K>(&w.kmax[p], kmax); }
}
template <typename T>
__global__ void crank_setup_kernel(CrWork<T> w, long long pc) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < pc) cr_setup_body<T>(w, p);
}
template <typename T>
__global__ void __launch_bounds__(CR_THREADS) crank_count_kernel(CrWork<T> w, const T* __restrict__ x, long long n) {
  const long long p = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * CR_TILE;
  const long long t1 = t0 + CR_TILE < n ? t0 + CR_TILE : n;
  for (long long i = t0 + threadIdx.x; i < t1; i += CR_U * CR_THREADS) cr_count_body<T>(w, x, n, p, i, CR_THREADS, t1);
}
// scan sweeps: one warp per (parameter, segment), global warp index = p * CR_NSEG + s; grid = pc * CR_NSEG / 8 CTAs
template <typename T>
__global__ void __launch_bounds__(CR_THREADS) crank_scan1_kernel(CrWork<T> w, long long pc) {
  const long long gw = (blockIdx.x * (long long)CR_THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= pc * CR_NSEG) return;   // (a whole warp leaves together)
  const long long p = gw / CR_NSEG;
  const int s = (int)(gw % CR_NSEG);
  const long long w0 = (long long)s * w.seg, w1 = (w0 + w.seg < w.nw) ? w0 + w.seg : w.nw;
  const uint2* cw = w.cw + p * w.nw;
  unsigned sum = 0;
  for (long long i = w0 + lane; i < w1; i += 32) sum += cr_nibsum(cw[i].x);
  sum = __reduce_add_sync(0xffffffffu, sum);
  if (lane == 0) w.part[gw] = sum;
}
template <typename T>
__global__ void __launch_bounds__(CR_THREADS) crank_scan3_kernel(CrWork<T> w, long long pc) {
  const long long gw = (blockIdx.x * (long long)CR_THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= pc * CR_NSEG) return;
  const long long p = gw / CR_NSEG;
  const int s = (int)(gw % CR_NSEG);
  const long long w0 = (long long)s * w.seg, w1 = (w0 + w.seg < w.nw) ? w0 + w.seg : w.nw;
  uint2* cw = w.cw + p * w.nw;
  unsigned run = 0;   // #values in the earlier segments of this parameter
  for (int j = lane; j < s; j += 32) run += w.part[p * CR_NSEG + j];
  run = __reduce_add_sync(0xffffffffu, run);
  for (long long i0 = w0; i0 < w1; i0 += 32) {   // (warp-uniform trip count)
    const long long i = i0 + lane;
    uint2 e = make_uint2(0u, 0u);
    if (i < w1) e = cw[i];
    const unsigned c = cr_nibsum(e.x);
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (i < w1) { e.y = run + incl - c; cw[i] = e; }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
}
template <typename T>
__global__ void __launch_bounds__(CR_THREADS) crank_place_kernel(CrWork<T> w, const T* __restrict__ x, long long n) {
  const long long p = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * CR_TILE;
  const long long t1 = t0 + CR_TILE < n ? t0 + CR_TILE : n;
  if (w.flag[p]) return;   // (the counters of a flagged slab may be corrupt; the chunk is redone by the sort path)
  for (long long i = t0 + threadIdx.x; i < t1; i += CR_U * CR_THREADS) cr_place_body<T>(w, x, n, p, i, CR_THREADS, t1);
}
template <typename T>
__global__ void crank_median_kernel(CrWork<T> w, long long n, long long pc, double* __restrict__ thr) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < pc) thr[p] = w.flag[p] ? 0.0 : (double)cr_median_body<T>(w, n, p);
}
// any flag of the chunk -> flag[pc]
template <typename T>
__global__ void crank_anyflag_kernel(CrWork<T> w, long long pc) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < pc && w.flag[p]) w.flag[pc] = 1;
}
#endif  // __CUDACC__

}  // namespace mcd
