// mcd_fast.cuh — the headline kernel: ess_rhat / rhat for kind in {rank, bulk, tail-rhat, basic}
// with the direct autocovariance, for slabs of exactly 8 split chains of <= 512 draws each
// (e.g. the canonical 1000 draws x 4 chains, split_chains = 2).
//
// One CTA (8 warps) per parameter, ONE WARP PER SPLIT CHAIN, the slab held in registers:
//   * global -> registers, coalesced, 16 independent loads in flight per thread (the slab is
//     read from HBM exactly once and never staged);
//   * rank-normalisation by counting into 65 536 FINE buckets (monotone linear map, scaled so that
//     no clamping is needed) whose populations live in 4-bit packed counters (8 per word, 32 KB):
//     one shared-memory atomic per element returns its arrival offset, a second one keeps the
//     population of the counter word as a byte, so that one warp-shuffle scan over 4 words per
//     u32 gives a 16-bit prefix per counter word; an element's sorted position is prefix +
//     nibble-sum of the lower counters of its word.  With n <= 4096 values ~90 % of the elements
//     are alone in their bucket and are ranked in O(1) (their z-table index is their position);
//     members of shared buckets store their VALUE at their sorted slot, are compacted into a work
//     list and resolved by all 256 threads: exact (less, equal) counts against the few bucket mates
//     by comparing values.  Exact average-tie ranks; z from a table indexed by the doubled rank
//     (integer ranks first, so untied data keeps half of it hot in L1);
//   * the median for the fold is captured from the ranks (no selection pass); the folded values
//     are ranked by the same code (second pass of the loop), so `:rank` costs one HBM read;
//   * split-chain moments are warp-shuffle reductions on registers;
//   * direct autocovariance from a padded (conflict-free) shared-memory copy of the centred
//     chain, register-blocked 8 lags x 8 draws per lane, lazily in batches of 8 lags with
//     Geyer's truncation deciding after each batch.
// Slabs that need the general machinery (NaN, infinite range, a fine bucket holding >= 15
// values, i.e. heavy ties) are appended to a redo list and recomputed by the general slab kernel.
//
// Compile-time specialisation LONG (split chains longer than 480 draws, e.g. the canonical 500): only
// the last of a thread's 16 slots can be empty, so the validity tests of the other 15 fold away.
// Work that only a few threads consume (R-hat, rho_k, the ESS finalisation) runs on those threads
// only: anything repeated by all 8 warps costs 8 x its instruction count.
// History and measurements: profiles/README.md.
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
constexpr int FAST_FINE = 65536;          // fine buckets
constexpr int FAST_WORDS = FAST_FINE / 8; // counter words (8 nibbles each)
constexpr int FAST_NMAX = FAST_NCH * FAST_MAXITER;  // 4096

// shared-memory layout (bytes):
//   [ FC  : FAST_WORDS u32 ][ WP : FAST_WORDS u16 ]     <- aliased by ZC[8][FAST_ROW] doubles
//   [ KV : FAST_NMAX values of T (the region is sized for two u32 planes: FAST_NMAX + 32 and FAST_NMAX words) ]
//   [ small arrays ]
constexpr int FAST_OFF_WP = FAST_WORDS * 4;
constexpr int FAST_OFF_KHI = FAST_OFF_WP + FAST_WORDS * 2;
constexpr int FAST_OFF_KLO = FAST_OFF_KHI + (FAST_NMAX + 32) * 4;
constexpr int FAST_OFF_SMALL = FAST_OFF_KLO + FAST_NMAX * 4;
static_assert(FAST_NCH * FAST_ROW * 8 <= FAST_OFF_KHI, "ZC must fit in the counter region");
template <typename T> static inline size_t fast_smem_bytes(int maxlag) {
  return (size_t)FAST_OFF_SMALL + 128 + 64 * 8 + 16 * 8 + 4 * 8 + 16 * 4 + (size_t)(maxlag + 9) * sizeof(T) + 64;
}

// sum of the eight 4-bit fields of w (each <= 15)
__device__ __forceinline__ unsigned nibsum(unsigned w) {
  const unsigned t = (w & 0x0f0f0f0fu) + ((w >> 4) & 0x0f0f0f0fu);
  return __dp4a(t, 0x01010101u, 0u);   // sum of the four bytes
}
template <typename T> struct FastArgs {
  const T* x;
  long long params;
  int niter;            // draws per split chain; n = nch * niter
  int nch;              // split chains: 8 for fast_kernel / fastgen_kernel, 1..8 for rk2_kernel
  int rank_x;           // bulk step rank-normalises x (kind bulk / rank); 0 = basic
  int do_bulk;          // compute the bulk / basic step
  int want_ess;         // bulk step computes ESS
  int do_tail;          // fold + rank-normalise + R-hat
  int maxlag, relative, ess_nan;
  T rel_ess_max;
  T* ess_out;
  T* rhat_out;
  const T* ztab;        // split layout: [n integer ranks][n-1 half ranks]
  int nbuckets;         // FAST_FINE
  int bucket_limit;     // unused (a 4-bit counter caps a bucket at 15)
  int* redo_list;
  int* redo_count;
};

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

// 15 consecutive draws of a padded row (index t + (t >> 4)) starting at draw 16 q + 8 E + 1, given
// base = row + 17 q: every offset is a compile-time constant.
template <int E>
__device__ __forceinline__ void fast_window(const double* base, double (&win)[15]) {
#pragma unroll
  for (int i = 0; i < 15; ++i) win[i] = base[(8 * E + 1 + i) + ((8 * E + 1 + i) >> 4)];
}

// LONG = every split chain has more than 480 draws: only the last of a thread's 16 slots can be empty, so
// the validity test of the other 15 folds away at compile time (the canonical 500-draw split chains).
template <typename T, bool LONG>
__global__ void __launch_bounds__(FAST_THREADS, 2) fast_kernel(const FastArgs<T> a) {
  using Key = typename Traits<T>::Key;
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = FAST_NCH * a.niter;
  unsigned* FC = reinterpret_cast<unsigned*>(smem);
  unsigned short* WP = reinterpret_cast<unsigned short*>(smem + FAST_OFF_WP);
  T* KV = reinterpret_cast<T*>(smem + FAST_OFF_KHI);   // values of shared-bucket members at their sorted slots
  unsigned* WC = reinterpret_cast<unsigned*>(smem + FAST_OFF_KHI);   // per-word populations as bytes (count + scan only; KV is written later)
  double* ZC = reinterpret_cast<double*>(smem);
  unsigned char* small = smem + FAST_OFF_SMALL;
  T* cmean = reinterpret_cast<T*>(small);                  // [8]
  T* cvar = cmean + 8;                                     // [8]
  double* part = reinterpret_cast<double*>(small + 128);   // [8][8]
  double* wred = part + 64;                                // [2][8]
  double* thr = wred + 16;                                 // [4]
  int* iflag = reinterpret_cast<int*>(thr + 4);            // [8] warp totals / flags
  int* woffx = iflag + 8;                                  // [8] [0] = length of the work list
  T* gamma = reinterpret_cast<T*>(woffx + 8);              // [maxlag + 9]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int niter = a.niter;
  const bool last_live = lane + 32 * (FAST_EPT - 1) < niter;
  auto live = [&](int k) -> bool { return LONG ? (k < FAST_EPT - 1 || last_live) : (lane + 32 * k < niter); };
  unsigned* listlen = reinterpret_cast<unsigned*>(woffx);
  if (tid == 0) *listlen = 0;

  for (long long param = blockIdx.x; param < a.params; param += gridDim.x) {
    const T* __restrict__ src = a.x + param * (long long)n + w * niter;
    T x[FAST_EPT], z[FAST_EPT];
#pragma unroll
    for (int k = 0; k < FAST_EPT; ++k) {
      const int t = lane + 32 * k;
      x[k] = live(k) ? __ldg(&src[t]) : (T)0;
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
        // clear the packed counters now: the barrier of the min / max exchange (pass 0) covers it
        for (int i = tid; i < FAST_WORDS / 4; i += FAST_THREADS) reinterpret_cast<uint4*>(FC)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < FAST_WORDS / 16; i += FAST_THREADS) reinterpret_cast<uint4*>(WC)[i] = make_uint4(0, 0, 0, 0);
        if (pass == 1) __syncthreads();
        if (pass == 0) {
          T lmin = (T)CUDART_INF, lmax = -(T)CUDART_INF;
          int bad = 0;
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            if (live(k)) {
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
        // slightly less than FINE / range: the largest value lands inside the last bucket, the smallest in bucket 0,
        // so the map needs no clamping (still monotone)
        const T scale = (T)((double)FAST_FINE * (1.0 - 1.0 / 1048576.0)) / range;
        if (!is_const && (!(range < (T)CUDART_INF) || !(scale > (T)0) || !(scale < (T)CUDART_INF))) { redo = true; break; }
        if (is_const) {
          // every value ties: rank (n+1)/2
          const T zc = __ldg(&a.ztab[((n - 1) >> 1) + (((n - 1) & 1) ? n : 0)]);
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) z[k] = zc;
          if (pass == 0 && tid == 0) { thr[0] = (double)vmin; thr[1] = (double)vmin; }
        } else {
          // ---- count: 4-bit packed populations, one atomic per element --------------------------
          unsigned bo[FAST_EPT];   // fine bucket | arrival offset << 16 ; later: packed rank info
          unsigned maxoff = 0, shared_mask = 0;
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            if (live(k)) {
              const unsigned fb = (unsigned)(int)((x[k] - vmin) * scale);
              const unsigned sh = (fb & 7u) * 4u;
              const unsigned off = (atomicAdd(&FC[fb >> 3], 1u << sh) >> sh) & 15u;
              atomicAdd(&WC[fb >> 5], 1u << ((fb >> 3) & 3u) * 8u);   // population of the counter word, one byte per word
              maxoff = off > maxoff ? off : maxoff;
              bo[k] = fb | (off << 16);
            } else bo[k] = 0;
          }
          // a counter that reaches 16 spills into its neighbour: the value that did it saw 15
          if (__syncthreads_or(maxoff >= 15u)) { redo = true; break; }
          // ---- scan: WP[word] = #values in earlier words of this warp's 1024-word range.  The byte
          // populations of 4 words sit in one u32; a lane owns two runs of 16 words (conflict-free 128-bit
          // loads) and one warp scan carries both runs as two 16-bit partial sums -----------------------------
          {
            const uint4* wc4 = reinterpret_cast<const uint4*>(WC + w * (FAST_WORDS / 32));
            const uint4 ca = wc4[lane], cb = wc4[32 + lane];
            const unsigned C[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
            unsigned lo[8], pr[8], tt[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              lo[i] = C[i] & 0x00ff00ffu;                        // (b2, b0) in 16-bit lanes
              pr[i] = lo[i] + ((C[i] >> 8) & 0x00ff00ffu);       // (b2 + b3, b0 + b1)
              tt[i] = (pr[i] & 0xffffu) + (pr[i] >> 16);         // population of the 4 words
            }
            const unsigned totA = tt[0] + tt[1] + tt[2] + tt[3], totB = tt[4] + tt[5] + tt[6] + tt[7];
            const unsigned tot = totA | (totB << 16);
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            const unsigned all = __shfl_sync(0xffffffffu, incl, 31);
            const unsigned excl = incl - tot;
            unsigned out[16];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              unsigned base = hf == 0 ? (excl & 0xffffu) : (all & 0xffffu) + (excl >> 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int j = 4 * hf + i;
                const unsigned B = base * 0x00010001u;                                       // (base, base)
                out[2 * j] = B + (lo[j] << 16);                                              // words 4j, 4j+1: base, base + b0
                out[2 * j + 1] = B + (pr[j] & 0xffffu) * 0x00010001u + (lo[j] & 0xffff0000u);  // base + b0 + b1, ... + b2
                base += tt[j];
              }
            }
            uint4* wp4 = reinterpret_cast<uint4*>(WP + w * (FAST_WORDS / 8)) + 2 * lane;
            wp4[0] = make_uint4(out[0], out[1], out[2], out[3]);
            wp4[1] = make_uint4(out[4], out[5], out[6], out[7]);
            wp4[FAST_WORDS / 128] = make_uint4(out[8], out[9], out[10], out[11]);
            wp4[FAST_WORDS / 128 + 1] = make_uint4(out[12], out[13], out[14], out[15]);
            const unsigned carry = (all & 0xffffu) + (all >> 16);
            if (lane == 0) iflag[w] = (int)carry;
            __syncthreads();
          }
          // lane i < 8 holds the number of values in the ranges of warps 0 .. i-1
          unsigned woff;
          {
            const unsigned tot = lane < FAST_NCH ? (unsigned)iflag[lane] : 0u;
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < FAST_NCH; o <<= 1) {
              const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            woff = incl - tot;
          }
          // ---- position: start of the fine bucket, population, own slot; shared buckets scatter ----
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            const bool valid = live(k);
            const unsigned fb = bo[k] & 0xffffu, off = bo[k] >> 16;
            const unsigned word = fb >> 3, sh = (fb & 7u) * 4u;
            const unsigned fw = FC[word];
            const unsigned st = (unsigned)WP[word] + __shfl_sync(0xffffffffu, woff, (int)(word >> 10)) + nibsum(fw & ((1u << sh) - 1u));
            const unsigned c = valid ? ((fw >> sh) & 15u) : 0u;
            if (c >= 2u) {
              KV[st + off] = x[k];
            }
            bo[k] = st | (c << 12) | (off << 16);   // st <= 4095
            shared_mask |= (c >= 2u ? 1u : 0u) << k;
          }
          __syncthreads();
          // ---- resolve shared buckets: their members (~10 % of the values) are compacted into a work
          // list (aliasing the dead counter words) and compared with their bucket mates by all threads;
          // the z-table index of each comes back through RES (aliasing the dead prefixes).  A value
          // alone in its bucket has the integer rank st + 1, i.e. table index st. -----------------------
          const int mA = (n & 1) ? n / 2 : n / 2 - 1, mB = n / 2;
          const bool capture = pass == 0 && a.do_tail;
          unsigned* WL = FC;
          unsigned short* RES = WP;
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
              for (int k = 0; k < FAST_EPT; ++k)
                if (shared_mask & (1u << k)) WL[q++] = bo[k] | (slot0 + ((unsigned)k << 28));   // slot = k * 256 + tid
            }
          }
          __syncthreads();
          {
            const unsigned listn = *listlen;
            for (unsigned q = tid; q < listn; q += FAST_THREADS) {
              const unsigned it = WL[q];
              const unsigned st = it & 0xfffu, c = (it >> 12) & 15u, off = (it >> 16) & 15u;
              // exact counts against the bucket mates on the values themselves (== ties -0.0 with 0.0, as tiedrank does)
              const T v = KV[st + off];
              unsigned less = 0, eq = 0;
              for (unsigned j = st; j < st + c; ++j) { const T y = KV[j]; less += y < v; eq += y == v; }
              const int lo = (int)(st + less), hi = lo + (int)eq;
              if (capture) {
                if (lo <= mA && mA < hi) thr[0] = (double)v;
                if (lo <= mB && mB < hi) thr[1] = (double)v;
              }
              const unsigned zi = (unsigned)(lo + hi - 1);   // r2 - 2 = 2*lo + eq - 1
              RES[it >> 20] = (unsigned short)((zi >> 1) + ((zi & 1u) ? (unsigned)n : 0u));   // split z table
            }
          }
          __syncthreads();
          if (tid == 0) *listlen = 0;
          if (capture) {
#pragma unroll
            for (int k = 0; k < FAST_EPT; ++k) {
              const int st = (int)(bo[k] & 0xfffu);
              if (live(k) && !(shared_mask & (1u << k))) {
                if (st == mA) thr[0] = (double)x[k];
                if (st == mB) thr[1] = (double)x[k];
              }
            }
          }
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k)
            bo[k] = (shared_mask & (1u << k)) ? (unsigned)RES[k * FAST_THREADS + tid] : (bo[k] & 0xfffu);
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) z[k] = live(k) ? __ldg(&a.ztab[bo[k]]) : (T)0;
        }
      } else {
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) z[k] = x[k];
      }
      if (!need_reduce) continue;

      // ---- split-chain moments: warp w owns split chain w -------------------------------------------
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < FAST_EPT; ++k) if (live(k)) s += (double)z[k];
      s = warp_sum(s);
      const T m = (T)(s / (double)niter);
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < FAST_EPT; ++k) if (live(k)) { const T d = z[k] - m; q = fma((double)d, (double)d, q); }
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
          row[t + (t >> 4)] = live(k) ? (double)(T)(z[k] - m) : 0.0;
        }
        for (int t = FAST_MAXITER + lane; t < FAST_TMAX; t += 32) row[t + (t >> 4)] = 0.0;
      }
      __syncthreads();
      SplitGeom g8;
      g8.niter = niter; g8.nch = FAST_NCH;
      // W and var_plus are used by the threads that form rho_k (tid < 8) and by thread 0 (R-hat): warp 0 only
      T W = (T)0, var_plus = (T)1;
      if (w == 0) within_between<T>(cmean, cvar, g8, W, var_plus);
      if (tid == 0) {
        const double rh = (double)sqrt(var_plus / W);
        if (pass == 0) rhat_bulk = rh; else rhat_tail = rh;
      }
      if (!do_ess) continue;

      // ---- direct autocovariance, lazily, Geyer truncation (ess_rhat.jl:553-594) -----------------
      const int maxlag = a.maxlag;
      int have = 0;
      const T inv_var_plus = (T)1 / var_plus;
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
          // window draws t0 + k0 .. t0 + k0 + 14 with k0 = 8 b + 1: in units of 8 draws the window starts at
          // c = b + h, and with the parity of c known every padded index is the lane's base + a constant
          const int c = (k0 >> 3) + h;
          if (c <= 8) {   // the whole window lies inside the zero-filled row: no bound checks
            const double* base = row + 17 * (lane + (c >> 1));
            if (c & 1) fast_window<1>(base, win); else fast_window<0>(base, win);
          } else {
#pragma unroll
            for (int i = 0; i < 15; ++i) { const int t = t0 + k0 + i; win[i] = t < FAST_TMAX ? row[t + (t >> 4)] : 0.0; }
          }
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
            const T gk = (T)(sum / (double)FAST_NCH) / (T)niter;
            gamma[k] = (T)1 - inv_var_plus * (W - gk);   // rho_k (ess_rhat.jl:556,566-567): stored instead of gamma_k
          }
        }
        __syncthreads();
      };
      auto ensure = [&](int k) { while (have < k) { batch(have + 1); have += 8; } };
      auto rho = [&](int k) -> T { return gamma[k]; };
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
      if (tid == 0) {
        const T tau = jl_max<T>((T)0, (T)2 * sum_p + jl_max<T>((T)0, rho_even) - (T)1);
        T e = jl_min<T>((T)1 / tau, a.rel_ess_max);
        if (!a.relative) e *= (T)(niter * FAST_NCH);
        ess = (double)e;
      }
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
