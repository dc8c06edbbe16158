// mcd_fast.cuh — the headline kernel: ess_rhat / rhat for kind in {rank, bulk, tail-rhat, basic}
// with the direct autocovariance, for slabs of exactly 8 split chains of <= 512 draws each
// (e.g. the canonical 1000 draws x 4 chains, split_chains = 2).
//
// One CTA (8 warps) per parameter, ONE WARP PER SPLIT CHAIN, the slab held in registers:
//   * global -> registers, coalesced, 16 independent loads in flight per thread (the slab is
//     read from HBM exactly once and never staged);
//   * rank-normalisation by counting: a monotone linear bucket map (8192 buckets), one
//     shared-memory atomic per element, a packed {start,count} scan, a scatter of the
//     order-preserving keys split into 32-bit hi / lo planes (the resolve loop gathers only the
//     hi plane; lo is touched on the rare hi tie), exact average-tie ranks, z from a table
//     indexed by the doubled rank;
//   * the median for the fold is captured from the ranks (no selection pass), the folded
//     values are ranked by the same code (second pass of the loop), so `:rank` costs one
//     HBM read;
//   * split-chain moments are warp-shuffle reductions on registers;
//   * direct autocovariance from a padded (conflict-free) shared-memory copy of the centred
//     chain, register-blocked 8 lags x 8 draws per lane, lazily in batches of 8 lags with
//     Geyer's truncation deciding after each batch.
// Slabs that need the general machinery (NaN, infinite range, a bucket over the limit) are
// appended to a redo list and recomputed by the general slab kernel.
//
// Reference citations (/root/reference): utils.jl:13-41,148-193; ess_rhat.jl:362-409,488-624.
#pragma once
#include "mcd_common.cuh"
#include "mcd_slab.cuh"

namespace mcd {

constexpr int FAST_THREADS = 256;
constexpr int FAST_EPT = 16;              // elements per thread
constexpr int FAST_NCH = 8;               // split chains = warps
constexpr int FAST_MAXITER = 32 * FAST_EPT;  // 512 draws per split chain
constexpr int FAST_ROW = 616;             // padded centred-chain row (doubles): pad(575) = 610
constexpr int FAST_TMAX = 576;            // the row is zero-filled on [niter, FAST_TMAX)

template <typename T> struct FastArgs {
  const T* x;
  long long params;
  int niter;            // draws per split chain; n = 8 * niter
  int rank_x;           // bulk step rank-normalises x (kind bulk / rank); 0 = basic
  int do_bulk;          // compute the bulk / basic step
  int want_ess;         // bulk step computes ESS
  int do_tail;          // fold + rank-normalise + R-hat
  int maxlag, relative, ess_nan;
  T rel_ess_max;
  T* ess_out;
  T* rhat_out;
  const T* ztab;        // [2n-1]
  int nbuckets;         // multiple of 2048
  int bucket_limit;
  int* redo_list;
  int* redo_count;
};

template <typename T> struct FastKeys;
template <> struct FastKeys<double> { static constexpr bool TWO = true; };
template <> struct FastKeys<float> { static constexpr bool TWO = false; };

// sum of 8 per-lane accumulators over the warp with 9 double shuffles; lane (l & 7) ... see below:
// after the call, lanes whose (l >> 2) == q hold the total of acc[q] (q = 0..7).
__device__ __forceinline__ double warp_reduce8(double (&a)[8]) {
  const int lane = threadIdx.x & 31;
  // step 1: partner = lane ^ 16; lower half keeps a[0..3], upper half keeps a[4..7]
  double b[4];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double send = up ? a[i] : a[i + 4];
      const double keep = up ? a[i + 4] : a[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  double c[2];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double send = up ? b[i] : b[i + 2];
      const double keep = up ? b[i + 2] : b[i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  double d;
  {
    const bool up = lane & 4;
    const double send = up ? c[0] : c[1];
    const double keep = up ? c[1] : c[0];
    d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;  // lane l holds total of acc[(l >> 2)]: bit4 -> +4, bit3 -> +2, bit2 -> +1
}

// exact (less, eq) of an element inside its bucket on the full key; returns less | eq << 16.
// Kept out of line: it runs only for elements whose hi word collides with a bucket-mate's.
template <bool TWO>
__device__ __noinline__ unsigned resolve_exact(const unsigned* Khi, const unsigned* Klo, int st, int c, unsigned vhi,
                                               unsigned vlo) {
  unsigned less = 0, eq = 0;
  for (int j = st; j < st + c; ++j) {
    const unsigned yhi = Khi[j];
    if (TWO) {
      const unsigned ylo = Klo[j];
      less += (yhi < vhi) | ((yhi == vhi) & (ylo < vlo));
      eq += (yhi == vhi) & (ylo == vlo);
    } else { less += (yhi < vhi); eq += (yhi == vhi); }
  }
  return less | (eq << 16);
}

template <typename T>
__global__ void __launch_bounds__(FAST_THREADS, 2) fast_kernel(const FastArgs<T> a) {
  using Key = typename Traits<T>::Key;
  extern __shared__ __align__(16) unsigned char smem[];
  const int B = a.nbuckets;
  // layout: [Khi n*4][Klo n*4][CNT (B+4)*4] aliased by ZC[8][FAST_ROW] doubles ; then small arrays
  const int n = FAST_NCH * a.niter;
  unsigned* Khi = reinterpret_cast<unsigned*>(smem);
  unsigned* Klo = Khi + FAST_NCH * FAST_MAXITER;
  unsigned* CNT = Klo + FAST_NCH * FAST_MAXITER;
  double* ZC = reinterpret_cast<double*>(smem);
  constexpr int BIG = FAST_NCH * FAST_ROW * 8;  // bytes of ZC
  const int big_bytes = max(BIG, (2 * FAST_NCH * FAST_MAXITER + B + 4) * 4);
  unsigned char* small = smem + ((big_bytes + 15) & ~15);
  T* cmean = reinterpret_cast<T*>(small);                  // [8]
  T* cvar = cmean + 8;                                     // [8]
  double* part = reinterpret_cast<double*>(small + 128);   // [8][8]
  double* wred = part + 64;                                // [2][8]
  double* thr = wred + 16;                                 // [4]
  int* iflag = reinterpret_cast<int*>(thr + 4);            // [8]
  T* gamma = reinterpret_cast<T*>(iflag + 8);              // [maxlag + 9]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int niter = a.niter;

  for (long long param = blockIdx.x; param < a.params; param += gridDim.x) {
    const T* __restrict__ src = a.x + param * (long long)n + w * niter;
    T x[FAST_EPT], z[FAST_EPT];
#pragma unroll
    for (int k = 0; k < FAST_EPT; ++k) {
      const int t = lane + 32 * k;
      x[k] = t < niter ? __ldg(&src[t]) : (T)0;
    }
    double ess = (double)Traits<T>::nan(), rhat_bulk = 0.0, rhat_tail = 0.0;
    bool redo = false;
    T vmin = (T)0, vmax = (T)0;

    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 0 && !a.do_bulk && !a.do_tail) break;
      if (pass == 1 && !a.do_tail) break;
      const bool need_rank = pass == 1 || a.rank_x || a.do_tail;  // pass 0 ranks x also to find the median
      const bool need_reduce = pass == 1 || a.do_bulk;
      if (pass == 1) {
        __syncthreads();  // thr[] written by the pass-0 resolve is visible
        // _fold_around_median: Statistics.median = middle of the two central order statistics
        const T med = (n & 1) ? (T)thr[0] : (T)((T)thr[0] / (T)2 + (T)thr[1] / (T)2);
        // |x - med| is monotone on each side of med, so its maximum sits at an extreme of x
        const T fa = fabs(vmin - med), fb = fabs(vmax - med);
        vmax = fa > fb ? fa : fb;
        vmin = (T)0;
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) x[k] = fabs(x[k] - med);
      }
      if (need_rank) {
        __syncthreads();  // previous users of the big region (ZC / K / CNT) are done
        if (pass == 0) {
          T lmin = (T)CUDART_INF, lmax = -(T)CUDART_INF;
          int bad = 0;
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            if (lane + 32 * k < niter) {
              const T v = x[k];
              bad |= (v != v);
              lmin = v < lmin ? v : lmin;
              lmax = v > lmax ? v : lmax;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const T p = __shfl_xor_sync(0xffffffffu, lmin, o); lmin = p < lmin ? p : lmin;
            const T q = __shfl_xor_sync(0xffffffffu, lmax, o); lmax = q > lmax ? q : lmax;
          }
          bad = __any_sync(0xffffffffu, bad);
          if (lane == 0) { wred[w] = (double)lmin; wred[8 + w] = (double)lmax; iflag[w] = bad; }
          __syncthreads();
          vmin = (T)wred[0]; vmax = (T)wred[8];
          int anybad = iflag[0];
#pragma unroll
          for (int i = 1; i < FAST_NCH; ++i) {
            const T p = (T)wred[i], q = (T)wred[8 + i];
            vmin = p < vmin ? p : vmin; vmax = q > vmax ? q : vmax;
            anybad |= iflag[i];
          }
          if (anybad) { redo = true; break; }
        }
        const bool is_const = !(vmax > vmin);
        const T range = vmax - vmin;
        const T scale = (T)B / range;
        if (!is_const && (!(range < (T)CUDART_INF) || !(scale > (T)0) || !(scale < (T)CUDART_INF))) { redo = true; break; }
        if (is_const) {
          // every value ties: rank (n+1)/2
          const T zc = __ldg(&a.ztab[n - 1]);
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) z[k] = zc;
          if (pass == 0 && tid == 0) { thr[0] = (double)vmin; thr[1] = (double)vmin; }
        } else {
          // ---- count ---------------------------------------------------------------------------
          for (int i = tid; i < (B + 4) / 4; i += FAST_THREADS) reinterpret_cast<uint4*>(CNT)[i] = make_uint4(0, 0, 0, 0);
          __syncthreads();
          unsigned bo[FAST_EPT];
          unsigned maxoff = 0;
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            if (lane + 32 * k < niter) {
              const int b = bucket_of<T>(x[k], (double)vmin, (double)scale, B);
              const unsigned old = atomicAdd(&CNT[b], 0x10000u) >> 16;
              maxoff = old > maxoff ? old : maxoff;
              bo[k] = (unsigned)b | (old << 16);
            } else bo[k] = 0;
          }
          if (__syncthreads_or(maxoff >= (unsigned)a.bucket_limit)) { redo = true; break; }
          // ---- scan: CNT[b] = start | count << 16 ---------------------------------------------------
          {
            const int per_warp = B / FAST_NCH;       // buckets per warp, multiple of 256
            unsigned carry = 0;
            // each lane owns 4 buckets in each of two consecutive 128-bucket groups; the two
            // lane totals ride one 32-bit shuffle scan packed as 16-bit halves (n <= 4096)
            for (int it = 0; it < per_warp / 256; ++it) {
              uint4* p4 = reinterpret_cast<uint4*>(CNT + w * per_warp + it * 256) + lane;
              uint4 c4 = p4[0], d4 = p4[32];
              const unsigned c0 = c4.x >> 16, c1 = c4.y >> 16, c2 = c4.z >> 16, c3 = c4.w >> 16;
              const unsigned d0 = d4.x >> 16, d1 = d4.y >> 16, d2 = d4.z >> 16, d3 = d4.w >> 16;
              const unsigned tot = (c0 + c1 + c2 + c3) | ((d0 + d1 + d2 + d3) << 16);
              unsigned incl = tot;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
              }
              const unsigned all = __shfl_sync(0xffffffffu, incl, 31);
              const unsigned excl = incl - tot;
              unsigned s0 = carry + (excl & 0xffffu);
              c4.x = s0 | (c0 << 16); s0 += c0;
              c4.y = s0 | (c1 << 16); s0 += c1;
              c4.z = s0 | (c2 << 16); s0 += c2;
              c4.w = s0 | (c3 << 16);
              unsigned s1 = carry + (all & 0xffffu) + (excl >> 16);
              d4.x = s1 | (d0 << 16); s1 += d0;
              d4.y = s1 | (d1 << 16); s1 += d1;
              d4.z = s1 | (d2 << 16); s1 += d2;
              d4.w = s1 | (d3 << 16);
              p4[0] = c4; p4[32] = d4;
              carry += (all & 0xffffu) + (all >> 16);
            }
            if (lane == 0) iflag[w] = (int)carry;
            __syncthreads();
            unsigned woff = 0;
            for (int i = 0; i < w; ++i) woff += (unsigned)iflag[i];
            if (woff) {
              for (int it = 0; it < per_warp / 128; ++it) {  // 128 buckets = 32 lanes x uint4
                uint4* p4 = reinterpret_cast<uint4*>(CNT + w * per_warp + it * 128) + lane;
                uint4 c4 = *p4;
                c4.x += woff; c4.y += woff; c4.z += woff; c4.w += woff;
                *p4 = c4;
              }
            }
            __syncthreads();
          }
          // ---- scatter the keys (hi / lo planes) --------------------------------------------------
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            if (lane + 32 * k < niter) {
              const unsigned pos = (CNT[bo[k] & 0xffffu] & 0xffffu) + (bo[k] >> 16);
              const Key key = order_key_nonan(x[k]);
              if (FastKeys<T>::TWO) { Khi[pos] = (unsigned)((unsigned long long)key >> 32); Klo[pos] = (unsigned)key; }
              else Khi[pos] = (unsigned)key;
            }
          }
          __syncthreads();
          // ---- resolve: exact doubled average rank, z lookup, median capture ------------------------
          // Round r compares every element with the r-th member of its bucket, for all 16 elements
          // of the thread at once (16 independent shared-memory gathers in flight).  The trip count
          // is the warp-wide maximum bucket population; an element past its bucket end re-reads its
          // own slot, which contributes nothing.  Only the 32-bit hi plane is gathered; a hi tie
          // with another element (two values within 2^-20 relative) is settled exactly on (hi, lo)
          // by resolve_exact().
          const int mA = (n & 1) ? n / 2 : n / 2 - 1, mB = n / 2;
          unsigned vhi[FAST_EPT], acc[FAST_EPT];
          int cm = 0;
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            const bool valid = lane + 32 * k < niter;
            const unsigned cw = CNT[bo[k] & 0xffffu];
            const unsigned st = cw & 0xffffu, c = valid ? (cw >> 16) : 0u;
            const Key key = order_key_nonan(x[k]);
            vhi[k] = FastKeys<T>::TWO ? (unsigned)((unsigned long long)key >> 32) : (unsigned)key;
            bo[k] = st | (c << 13) | ((st + (bo[k] >> 16)) << 20);   // start | count | own position
            cm = (int)c > cm ? (int)c : cm;
            acc[k] = 0;
          }
          const int rounds = __reduce_max_sync(0xffffffffu, cm);
          for (int r = 0; r < rounds; ++r) {
#pragma unroll
            for (int k = 0; k < FAST_EPT; ++k) {
              const unsigned st = bo[k] & 0x1fffu, c = (bo[k] >> 13) & 0x7fu, mypos = bo[k] >> 20;
              const unsigned idx = (unsigned)r < c ? st + (unsigned)r : mypos;
              const unsigned yhi = Khi[idx];
              acc[k] += (unsigned)(yhi < vhi[k]) + ((unsigned)((yhi == vhi[k]) & (idx != mypos)) << 16);
            }
          }
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            const bool valid = lane + 32 * k < niter;
            const int st = (int)(bo[k] & 0x1fffu), c = (int)((bo[k] >> 13) & 0x7fu);
            int less = (int)(acc[k] & 0xffffu), eq = 1;
            if (__any_sync(0xffffffffu, acc[k] >> 16)) {
              if (acc[k] >> 16) {
                const Key key = order_key_nonan(x[k]);
                const unsigned le = resolve_exact<FastKeys<T>::TWO>(Khi, Klo, st, c, vhi[k], (unsigned)key);
                less = (int)(le & 0xffffu); eq = (int)(le >> 16);
              }
            }
            const int lo = st + less, hi = lo + eq;
            if (pass == 0 && a.do_tail && valid) {
              if (lo <= mA && mA < hi) thr[0] = (double)x[k];
              if (lo <= mB && mB < hi) thr[1] = (double)x[k];
            }
            bo[k] = (unsigned)(lo + hi - 1);   // r2 - 2 = 2*lo + eq - 1
          }
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) z[k] = (lane + 32 * k < niter) ? __ldg(&a.ztab[bo[k]]) : (T)0;
        }
      } else {
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) z[k] = x[k];
      }
      if (!need_reduce) continue;

      // ---- split-chain moments: warp w owns split chain w -------------------------------------------
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < FAST_EPT; ++k) if (lane + 32 * k < niter) s += (double)z[k];
      s = warp_sum(s);
      const T m = (T)(s / (double)niter);
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < FAST_EPT; ++k) if (lane + 32 * k < niter) { const T d = z[k] - m; q = fma((double)d, (double)d, q); }
      q = warp_sum(q);
      __syncthreads();  // all resolve loops are done with K / CNT; cmean / cvar free
      if (lane == 0) { cmean[w] = m; cvar[w] = (T)(q / (double)(niter - 1)); }
      const bool do_ess = pass == 0 && a.want_ess && !a.ess_nan;
      if (do_ess) {
        // centred chain into the padded row: index t + (t >> 4)
        double* row = ZC + w * FAST_ROW;
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) {
          const int t = lane + 32 * k;
          row[t + (t >> 4)] = t < niter ? (double)(T)(z[k] - m) : 0.0;
        }
        for (int t = FAST_MAXITER + lane; t < FAST_TMAX; t += 32) row[t + (t >> 4)] = 0.0;
      }
      __syncthreads();
      SplitGeom g8;
      g8.niter = niter; g8.nch = FAST_NCH;
      T W, var_plus;
      within_between<T>(cmean, cvar, g8, W, var_plus);
      const double rh = (double)sqrt(var_plus / W);
      if (pass == 0) rhat_bulk = rh; else rhat_tail = rh;
      if (!do_ess) continue;

      // ---- direct autocovariance, lazily, Geyer truncation (ess_rhat.jl:553-594) -----------------
      const int maxlag = a.maxlag;
      int have = 0;
      auto batch = [&](int k0) {
        const double* row = ZC + w * FAST_ROW;
        double acc[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) acc[kk] = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int t0 = 16 * lane + 8 * h;
          double own[8], win[15];
#pragma unroll
          for (int i = 0; i < 8; ++i) own[i] = row[t0 + i + lane];      // (t0+i)>>4 == lane
#pragma unroll
          for (int i = 0; i < 15; ++i) { const int t = t0 + k0 + i; win[i] = t < FAST_TMAX ? row[t + (t >> 4)] : 0.0; }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[kk] = fma(own[i], win[i + kk], acc[kk]);
        }
        const double tot = warp_reduce8(acc);
        if ((lane & 3) == 0) part[w * 8 + (lane >> 2)] = tot;
        __syncthreads();
        if (tid < 8) {
          const int k = k0 + tid;
          if (k <= maxlag && k < niter) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < FAST_NCH; ++i) sum += part[i * 8 + tid];
            gamma[k] = (T)(sum / (double)FAST_NCH) / (T)niter;
          }
        }
        __syncthreads();
      };
      auto ensure = [&](int k) { while (have < k) { batch(have + 1); have += 8; } };
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
      if (!a.relative) e *= (T)(niter * FAST_NCH);
      ess = (double)e;
    }

    if (redo) {
      if (tid == 0) { const int idx = atomicAdd(a.redo_count, 1); a.redo_list[idx] = (int)param; }
    } else if (tid == 0) {
      double rhat = rhat_bulk;
      if (a.do_tail) rhat = a.do_bulk ? (double)jl_max<T>((T)rhat_tail, (T)rhat_bulk) : rhat_tail;
      if (a.ess_out) a.ess_out[param] = (T)ess;
      if (a.rhat_out) a.rhat_out[param] = (T)rhat;
    }
    __syncthreads();
  }
}

}  // namespace mcd
