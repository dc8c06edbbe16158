// mcd_fastgen.cuh — the general form of the fast kernel (mcd_fast.cuh): same shapes (8 split chains
// of <= 512 draws, direct autocovariance), same register-resident ranking, but a small program of
// REDUCTIONS per pass, so that the kinds whose proxies are order-statistic indicators reuse the
// ranks instead of re-sorting:
//   ess_rhat / ess (kind = :tail)   min of the two quantile-indicator ESS  (src/ess_rhat.jl:301-311)
//                                   [+ tail R-hat from the folded pass]
//   ess(kind = median | quantile(p)) indicator x <= threshold              (:630-639, :647-659)
//   ess(kind = std)                  (x - mean)^2                          (:640-642)
//   ess(kind = mad)                  fold, then indicator f <= median(f)   (:643-646)
// The order statistics behind the thresholds (median, type-7 quantile neighbours) are captured
// while the ranks are at hand; thresholds are then formed exactly as the reference does and the
// indicator compares the values themselves against them.
// The headline programs (rank / bulk / basic / tail R-hat) stay on the leaner mcd_fast.cuh kernel.
#pragma once
#include "mcd_fast.cuh"

namespace mcd {

// What a reduction (moments [+ autocovariance + Geyer]) runs on
enum FastSrc : int {
  FS_X = 0,      // the values themselves                                  (kind basic / estimator mean)
  FS_RANKZ = 1,  // z of the ranks just computed (must be the first reduction of its pass)
  FS_IND = 2,    // indicator  value <= threshold                          (median / quantile / tail ESS, mad)
  FS_SQDEV = 3   // (x - mean over draws and chains)^2                     (estimator std)
};
struct FastRed { int src, want_ess, thr; };
// threshold from captured order statistics: median a/2 + b/2, or type-7 quantile a + g (b - a)
struct FastThr { int quantile, capA, capB, f32; double g; };

constexpr int FASTGEN_WIN = 1024;           // window of sorted positions kept for late order statistics (mcse quantile rule)
static_assert(FAST_NCH * FAST_ROW * 8 + FASTGEN_WIN * 8 <= FAST_OFF_KHI, "WIN must fit behind ZC in the counter / prefix region");
static_assert(FAST_NMAX * 4 + FAST_NMAX * 2 + FAST_NMAX <= FAST_WORDS * 4, "work list + RES + WANT must fit in the counter region");
constexpr int FASTGEN_EXTRA_SMEM = 24 * 8;   // cap[8] + thrv[4] + res[12] + side[4] replace the lean kernel's thr[4]

template <typename T> struct FastGenArgs {
  const T* x;
  long long params;
  int niter;            // draws per split chain; n = 8 * niter
  // pass 0 works on x, pass 1 (do_fold) on |x - median(x)|
  int p0_rank;          // pass 0 needs the ranks of x (for z, or for order statistics)
  int p0_nred;          // reductions on pass-0 data -> result slots 0..4
  FastRed p0_red[5];
  int ncap;             // order statistics of x to capture: sorted positions (0-based) -> cap[0..ncap)
  int cap_pos[6];       // when do_fold: cap_pos[0], cap_pos[1] must be the two median positions
  int nthr;             // thresholds computed from the captures -> thrv[0..nthr)
  FastThr thr[3];
  int do_fold;
  FastRed p1_red;       // reduction on the folded data -> result slot 5 (FS_IND uses the median of the folded values)
  int ess_mode;         // 0 none, 1 slot 0, 2 min(slot 0, slot 1), 4 slot 5
  int rhat_mode;        // 0 none, 1 slot 0, 2 slot 5, 3 max(slot 5, slot 0)
  int mcse_mode;        // 0 none, 1 mean: std(x)/sqrt(ess) (mcse.jl:45-51), 2 std: sqrt((m4/m2 - m2)/ess)/2 (mcse.jl:52-65),
                        // 3 median / quantile: order-statistic rule (mcse.jl:96-118) on result slot 0
  double mcse_p;        // mode 3: the quantile's probability
  int win_lo;           // mode 3: first sorted position kept in WIN
  int need_side;        // mean / std / m2 / m4 over all draws and chains (mcse rules, FS_SQDEV, summary)
  // fused summary (mcd_summary): columns mean, std, mcse_mean, mcse_std, ess_bulk, ess_tail, rhat
  // (null = not requested); s_* = result slot of the reduction feeding a column (-1 = absent)
  int sum_mode;
  int s_bulk, s_tlo, s_thi, s_mean, s_std;
  T* col[7];
  int maxlag, relative, ess_nan;
  T rel_ess_max;
  T* ess_out;
  T* rhat_out;
  const T* ztab;        // split layout: [n integer ranks][n-1 half ranks]
  int* redo_list;
  int* redo_count;
};

// QRULE = the program ends with the MCSE order-statistic rule (mcse_mode 3): only that instance carries the
// window of sorted values and the Beta inverse, so the other programs do not pay for them.
// LONG: as in mcd_fast.cuh (split chains longer than 480 draws: 15 of the 16 slots are always occupied).
template <typename T, bool QRULE, bool LONG>
__global__ void __launch_bounds__(FAST_THREADS, 2) fastgen_kernel(const FastGenArgs<T> a) {
  using Key = typename Traits<T>::Key;
  extern __shared__ __align__(16) unsigned char smem[];
  const int n = FAST_NCH * a.niter;
  unsigned* FC = reinterpret_cast<unsigned*>(smem);
  unsigned short* WP = reinterpret_cast<unsigned short*>(smem + FAST_OFF_WP);
  T* KV = reinterpret_cast<T*>(smem + FAST_OFF_KHI);   // values of shared-bucket members at their sorted slots
  double* ZC = reinterpret_cast<double*>(smem);
  double* WIN = reinterpret_cast<double*>(smem + FAST_NCH * FAST_ROW * 8);   // [FASTGEN_WIN] sorted values, behind ZC
  unsigned char* small = smem + FAST_OFF_SMALL;
  T* cmean = reinterpret_cast<T*>(small);                  // [8]
  T* cvar = cmean + 8;                                     // [8]
  double* part = reinterpret_cast<double*>(small + 128);   // [8][8]
  double* wred = part + 64;                                // [2][8]
  double* cap = wred + 16;                                 // [8] captured order statistics (6,7: folded median)
  double* thrv = cap + 8;                                  // [4] thresholds (3: median of the folded values)
  double* res = thrv + 4;                                  // [12] ess[6], rhat[6] per result slot
  double* side = res + 12;                                 // [4] mean, std, mean(proxy), mean(proxy^2) over the slab
  int* iflag = reinterpret_cast<int*>(side + 4);           // [8] warp totals / flags
  int* woffx = iflag + 8;                                  // [8] [0] = length of the work list
  T* gamma = reinterpret_cast<T*>(woffx + 8);              // [maxlag + 9]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int niter = a.niter;
  auto live = [&](int k) -> bool { return LONG ? (k < FAST_EPT - 1 || lane + 32 * (FAST_EPT - 1) < niter) : (lane + 32 * k < niter); };
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
    bool redo = false;
    T vmin = (T)0, vmax = (T)0;
    // statistics over all draws x chains (Statistics.mean / std with dims=(1,2); src/mcse.jl:50,56,62-63)
    if (a.need_side) {
      auto block_total = [&](double v) -> double {
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) wred[w] = v;
        __syncthreads();
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < FAST_NCH; ++i) t += wred[i];
        return t;
      };
      double sx = 0.0;
#pragma unroll
      for (int k = 0; k < FAST_EPT; ++k) if (live(k)) sx += (double)x[k];
      const T mean_all = (T)(block_total(sx) / (double)n);
      double s2 = 0.0, s4 = 0.0;
#pragma unroll
      for (int k = 0; k < FAST_EPT; ++k) if (live(k)) {
        const T d = x[k] - mean_all;
        const T pz = d * d;
        s2 += (double)pz;
        s4 = fma((double)pz, (double)pz, s4);
      }
      const double t2 = block_total(s2), t4 = block_total(s4);
      if (tid == 0) {
        const T sd = sqrt((T)(t2 / (double)(n - 1)));          // std(x; corrected)
        side[0] = (double)mean_all; side[1] = (double)sd;
        side[2] = (double)(T)(t2 / (double)n); side[3] = (double)(T)(t4 / (double)n);   // mean(proxy), mean(proxy^2)
        if (a.sum_mode) {
          if (a.col[0]) a.col[0][param] = mean_all;
          if (a.col[1]) a.col[1][param] = sd;
        }
      }
      __syncthreads();   // side[] is read by every thread (FS_SQDEV) and by thread 0 at the end
    }

    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1 && !a.do_fold) break;
      const bool need_rank = pass == 1 || a.p0_rank;
      const int nred = pass == 0 ? a.p0_nred : 1;
      if (pass == 0 && !need_rank && nred == 0) continue;
      if (pass == 1) {
        __syncthreads();  // cap[] written by the pass-0 resolve is visible
        // _fold_around_median: Statistics.median = middle of the two central order statistics
        const T med = (n & 1) ? (T)cap[0] : (T)((T)cap[0] / (T)2 + (T)cap[1] / (T)2);
        // |x - med| is monotone on each side of med, so its maximum sits at an extreme of x
        const T fa = fabs(vmin - med), fb = fabs(vmax - med);
        vmax = fa > fb ? fa : fb;
        vmin = (T)0;
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) x[k] = fabs(x[k] - med);
      }
      const bool first_is_rankz = (pass == 0 ? (nred > 0 && a.p0_red[0].src == FS_RANKZ) : a.p1_red.src == FS_RANKZ);
      if (need_rank) {
        __syncthreads();  // previous users of the big region (ZC / K / CNT) are done
        // clear the packed counters now: the barrier of the min / max exchange (pass 0) covers it
        for (int i = tid; i < FAST_WORDS / 4; i += FAST_THREADS) reinterpret_cast<uint4*>(FC)[i] = make_uint4(0, 0, 0, 0);
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
        const T scale = (T)FAST_FINE / range;
        if (!is_const && (!(range < (T)CUDART_INF) || !(scale > (T)0) || !(scale < (T)CUDART_INF))) { redo = true; break; }
        if (is_const) {
          // every value ties: rank (n+1)/2
          const T zc = __ldg(&a.ztab[((n - 1) >> 1) + (((n - 1) & 1) ? n : 0)]);
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) z[k] = zc;
          if (tid < 8) cap[tid] = (double)vmin;   // every order statistic equals the common value
        } else {
          // ---- count: 4-bit packed populations, one atomic per element --------------------------
          unsigned bo[FAST_EPT];   // fine bucket | arrival offset << 16 ; later: packed rank info
          unsigned maxoff = 0, shared_mask = 0;
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k) {
            if (live(k)) {
              const unsigned fb = (unsigned)bucket_of<T>(x[k], (double)vmin, (double)scale, FAST_FINE);
              const unsigned sh = (fb & 7u) * 4u;
              const unsigned off = (atomicAdd(&FC[fb >> 3], 1u << sh) >> sh) & 15u;
              maxoff = off > maxoff ? off : maxoff;
              bo[k] = fb | (off << 16);
            } else bo[k] = 0;
          }
          // a counter that reaches 16 spills into its neighbour: the value that did it saw 15
          if (__syncthreads_or(maxoff >= 15u)) { redo = true; break; }
          // ---- scan: WP[word] = #values in earlier words of this warp's 1024-word range ---------
          {
            unsigned carry = 0;
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
              const int wbase = w * 1024 + it * 256 + 4 * lane;
              const uint4 c4 = *reinterpret_cast<const uint4*>(FC + wbase);
              const uint4 d4 = *reinterpret_cast<const uint4*>(FC + wbase + 128);
              const unsigned c0 = nibsum(c4.x), c1 = nibsum(c4.y), c2 = nibsum(c4.z), c3 = nibsum(c4.w);
              const unsigned d0 = nibsum(d4.x), d1 = nibsum(d4.y), d2 = nibsum(d4.z), d3 = nibsum(d4.w);
              const unsigned tot = (c0 + c1 + c2 + c3) | ((d0 + d1 + d2 + d3) << 16);
              unsigned incl = tot;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
              }
              const unsigned all = __shfl_sync(0xffffffffu, incl, 31);
              const unsigned excl = incl - tot;
              const unsigned s0 = carry + (excl & 0xffffu);
              const unsigned s1 = carry + (all & 0xffffu) + (excl >> 16);
              uint2 pc, pd;
              pc.x = s0 | ((s0 + c0) << 16); pc.y = (s0 + c0 + c1) | ((s0 + c0 + c1 + c2) << 16);
              pd.x = s1 | ((s1 + d0) << 16); pd.y = (s1 + d0 + d1) | ((s1 + d0 + d1 + d2) << 16);
              *reinterpret_cast<uint2*>(WP + wbase) = pc;
              *reinterpret_cast<uint2*>(WP + wbase + 128) = pd;
              carry += (all & 0xffffu) + (all >> 16);
            }
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
          // ---- resolve shared buckets through a compacted work list (see mcd_fast.cuh); order
          // statistics are captured while the ranks are at hand ------------------------------------
          const int ncap = pass == 0 ? a.ncap : (a.p1_red.src == FS_IND ? 2 : 0);
          const int cbase = pass == 0 ? 0 : 6;
          const int fmA = (n & 1) ? n / 2 : n / 2 - 1, fmB = n / 2;   // median positions (pass 1)
          unsigned* WL = FC;
          // the work list needs at most FAST_NMAX words = half of the dead counter region; RES and WANT take
          // the rest, which leaves the dead prefix region to WIN (it must outlive the reductions)
          unsigned short* RES = reinterpret_cast<unsigned short*>(FC + FAST_NMAX);
          // WANT[pos] = mask of the captures that ask for sorted position pos
          unsigned char* WANT = reinterpret_cast<unsigned char*>(FC + FAST_NMAX) + 2 * FAST_NMAX;
          const int wlo = a.win_lo;
          const bool fill_win = QRULE && pass == 0;
          if (ncap > 0 && w == 0) {
#pragma unroll
            for (int i = 0; i < FAST_NMAX / (16 * 32); ++i) reinterpret_cast<uint4*>(WANT)[lane + 32 * i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            if (lane == 0) {
#pragma unroll
              for (int ci = 0; ci < 6; ++ci) {
                const int cp = pass == 0 ? a.cap_pos[ci] : (ci == 0 ? fmA : fmB);
                if (ci < ncap) WANT[cp] |= (unsigned char)(1u << ci);
              }
            }
          }
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
              if (ncap > 0) {
                unsigned m = 0;
                for (int pos = lo; pos < hi; ++pos) m |= WANT[pos];
                if (m) for (int ci = 0; ci < 6; ++ci) if (m & (1u << ci)) cap[cbase + ci] = (double)v;
              }
              if (fill_win) {
                for (int pos = lo; pos < hi; ++pos) if ((unsigned)(pos - wlo) < (unsigned)FASTGEN_WIN) WIN[pos - wlo] = (double)v;
              }
              const unsigned zi = (unsigned)(lo + hi - 1);   // r2 - 2 = 2*lo + eq - 1
              RES[it >> 20] = (unsigned short)((zi >> 1) + ((zi & 1u) ? (unsigned)n : 0u));   // split z table
            }
          }
          __syncthreads();
          if (tid == 0) *listlen = 0;
          if (ncap > 0) {
#pragma unroll
            for (int k = 0; k < FAST_EPT; ++k) {
              if (live(k) && !(shared_mask & (1u << k))) {
                const unsigned m = WANT[bo[k] & 0xfffu];
                if (m) for (int ci = 0; ci < 6; ++ci) if (m & (1u << ci)) cap[cbase + ci] = (double)x[k];
              }
            }
          }
          if (fill_win) {
#pragma unroll
            for (int k = 0; k < FAST_EPT; ++k) {
              const unsigned rel = (bo[k] & 0xfffu) - (unsigned)wlo;
              if (live(k) && !(shared_mask & (1u << k)) && rel < (unsigned)FASTGEN_WIN) WIN[rel] = (double)x[k];
            }
          }
#pragma unroll
          for (int k = 0; k < FAST_EPT; ++k)
            bo[k] = (shared_mask & (1u << k)) ? (unsigned)RES[k * FAST_THREADS + tid] : (bo[k] & 0xfffu);
          if (first_is_rankz) {
#pragma unroll
            for (int k = 0; k < FAST_EPT; ++k) z[k] = live(k) ? __ldg(&a.ztab[bo[k]]) : (T)0;
          }
        }
        // thresholds from the captured order statistics
        __syncthreads();
        if (tid == 0) {
          if (pass == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              if (i >= a.nthr) break;
              const FastThr th = a.thr[i];
              const double ca = cap[th.capA], cb = cap[th.capB];
              double v;
              if (!th.quantile) v = (n & 1) ? ca : (double)((T)((T)ca / (T)2 + (T)cb / (T)2));
              else if (th.f32) {
                const float fa = (float)ca, fb = (float)cb, g = (float)th.g;
                v = (isfinite(fa) && isfinite(fb)) ? (double)__fadd_rn(fa, __fmul_rn(g, __fsub_rn(fb, fa)))
                                                   : (double)__fadd_rn(__fmul_rn(__fsub_rn(1.f, g), fa), __fmul_rn(g, fb));
              } else {
                v = (isfinite(ca) && isfinite(cb)) ? __dadd_rn(ca, __dmul_rn(th.g, __dsub_rn(cb, ca)))
                                                   : __dadd_rn(__dmul_rn(__dsub_rn(1.0, th.g), ca), __dmul_rn(th.g, cb));
              }
              thrv[i] = v;
            }
          } else if (a.p1_red.src == FS_IND) {
            thrv[3] = (n & 1) ? cap[6] : (double)((T)((T)cap[6] / (T)2 + (T)cap[7] / (T)2));
          }
        }
        __syncthreads();
      }

      for (int r = 0; r < nred; ++r) {
      const FastRed rd = pass == 1 ? a.p1_red : (r == 0 ? a.p0_red[0] : (r == 1 ? a.p0_red[1] : (r == 2 ? a.p0_red[2] : (r == 3 ? a.p0_red[3] : a.p0_red[4]))));
      const int slot = pass == 0 ? r : 5;
      // ---- the series this reduction runs on -----------------------------------------------------
      if (rd.src == FS_X) {
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) z[k] = x[k];
      } else if (rd.src == FS_IND) {
        const double tv = thrv[pass == 0 ? rd.thr : 3];
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) z[k] = ((double)x[k] <= tv) ? (T)1 : (T)0;
      } else if (rd.src == FS_SQDEV) {
        const T mean_all = (T)side[0];
#pragma unroll
        for (int k = 0; k < FAST_EPT; ++k) { const T d = x[k] - mean_all; z[k] = d * d; }
      }

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
      const bool do_ess = rd.want_ess && !a.ess_nan;
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
      if (tid == 0) { res[6 + slot] = (double)sqrt(var_plus / W); res[slot] = (double)Traits<T>::nan(); }
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
        res[slot] = (double)e;
      }
      }   // reductions
      if (redo) break;
    }

    T mcse_q = Traits<T>::nan();
    if (QRULE && !redo) {
      // _mcse_quantile (mcse.jl:96-118): Beta(S p + 1, S (1 - p) + 1) quantiles at Phi(-1), Phi(+1) -> order statistics
      __syncthreads();   // res[0] (thread 0) is visible; part[] is free
      const double S = res[0];
      if (S == S) {
        const double al = S * a.mcse_p + 1.0, be = S * (1.0 - a.mcse_p) + 1.0;
        betainc_inv_pair_block<FAST_THREADS>(al, be, 0.8413447460685429, 0.15865525393145705, part, part + 34);
        long long u = (long long)ceil(part[34] * (double)n);
        u = u > n ? n : (u < 1 ? 1 : u);
        long long l = (long long)floor(part[35] * (double)n);
        l = l < 1 ? 1 : (l > n ? n : l);
        const long long ru = u - 1 - a.win_lo, rl = l - 1 - a.win_lo;
        if (ru < 0 || ru >= FASTGEN_WIN || rl < 0 || rl >= FASTGEN_WIN) redo = true;   // outside the kept window: general kernel
        else mcse_q = ((T)WIN[ru] - (T)WIN[rl]) / (T)2;
      }
    }
    if (redo) {
      if (tid == 0) { const int idx = atomicAdd(a.redo_count, 1); a.redo_list[idx] = (int)param; }
    } else if (tid == 0) {
      if (QRULE) a.ess_out[param] = mcse_q;
      if (a.sum_mode) {
        // the reference calls each column stands for: src/mcse.jl:45-69, src/ess_rhat.jl:298-311, 410-420, 604-624
        if (a.col[2]) a.col[2][param] = (T)side[1] / sqrt((T)res[a.s_mean]);
        if (a.col[3]) { const T m2 = (T)side[2], m4 = (T)side[3]; a.col[3][param] = sqrt((m4 / m2 - m2) / (T)res[a.s_std]) / (T)2; }
        if (a.col[4]) a.col[4][param] = (T)res[a.s_bulk];
        if (a.col[5]) a.col[5][param] = jl_min<T>((T)res[a.s_tlo], (T)res[a.s_thi]);
        if (a.col[6]) a.col[6][param] = jl_max<T>((T)res[6 + 5], (T)res[6 + a.s_bulk]);
      }
      if (a.ess_out && !QRULE) {
        T e = (T)res[a.ess_mode == 4 ? 5 : 0];
        if (a.ess_mode == 2) e = jl_min<T>(e, (T)res[1]);
        if (a.mcse_mode == 1) e = (T)side[1] / sqrt(e);
        else if (a.mcse_mode == 2) { const T m2 = (T)side[2], m4 = (T)side[3]; e = sqrt((m4 / m2 - m2) / e) / (T)2; }
        a.ess_out[param] = e;
      }
      if (a.rhat_out) {
        T rh = (T)res[6 + (a.rhat_mode == 2 ? 5 : 0)];
        if (a.rhat_mode == 3) rh = jl_max<T>((T)res[6 + 5], rh);
        a.rhat_out[param] = rh;
      }
    }
    __syncthreads();
  }
}

}  // namespace mcd
