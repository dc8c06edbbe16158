// mcd_common.cuh — device helpers shared by the slab (shared-memory) and large-slab
// (global-memory) kernels of libmcmcdiag_b200.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace mcd {

constexpr int WARP = 32;

// ---------------------------------------------------------------------------------------
// status flags raised by kernels (one word per context, read back after the call)
// ---------------------------------------------------------------------------------------
enum : unsigned {
  FLAG_NAN_QUANTILE = 1u,   // quantile requested on a slab containing NaN (Julia throws)
  FLAG_SORT_FALLBACK = 2u,  // at least one slab used the full-sort fallback (statistics only)
};

// ---------------------------------------------------------------------------------------
// type traits
// ---------------------------------------------------------------------------------------
template <typename T> struct Traits;
template <> struct Traits<double> {
  using Key = unsigned long long;
  static constexpr Key KEY_NAN = ~0ull;
  __device__ static double nan() { return __longlong_as_double(0x7ff8000000000000ll); }
};
template <> struct Traits<float> {
  using Key = unsigned int;
  static constexpr Key KEY_NAN = ~0u;
  __device__ static float nan() { return __int_as_float(0x7fc00000); }
};

// Order-preserving map float -> unsigned: isless order, except that -0.0 and 0.0 are
// merged (they tie under `==`, which is what tiedrank's run detection uses) and every NaN
// maps to the single largest key (NaNs rank last; their mutual order is by index and is
// handled by the callers).
__device__ __forceinline__ unsigned long long order_key(double v) {
  if (v != v) return ~0ull;
  long long b = __double_as_longlong(v + 0.0);
  unsigned long long u = (unsigned long long)b;
  return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned int order_key(float v) {
  if (v != v) return ~0u;
  int b = __float_as_int(v + 0.0f);
  unsigned int u = (unsigned int)b;
  return (b < 0) ? ~u : (u | 0x80000000u);
}
// same map for values already known not to be NaN
__device__ __forceinline__ unsigned long long order_key_nonan(double v) {
  const long long b = __double_as_longlong(v + 0.0);
  return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ unsigned int order_key_nonan(float v) {
  const int b = __float_as_int(v + 0.0f);
  return (unsigned int)(b ^ ((b >> 31) | (int)0x80000000u));
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ float key_value(unsigned int k) {
  unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __int_as_float((int)u);
}

// Julia's min/max propagate NaN (src/ess_rhat.jl:574,591,594); fmin/fmax do not.
template <typename T> __device__ __forceinline__ T jl_min(T a, T b) {
  if (a != a || b != b) return Traits<T>::nan();
  return a < b ? a : b;
}
template <typename T> __device__ __forceinline__ T jl_max(T a, T b) {
  if (a != a || b != b) return Traits<T>::nan();
  return a > b ? a : b;
}

// ---------------------------------------------------------------------------------------
// warp / block reductions with a fixed tree (deterministic, batch-invariant)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `scratch` holds >= 33 doubles.  All threads get the result.
// Contains two __syncthreads(); must be called by every thread of the block.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  constexpr int NW = THREADS / WARP;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    double s = (lane < NW) ? scratch[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) scratch[32] = s;
  }
  __syncthreads();
  return scratch[32];
}

// ---------------------------------------------------------------------------------------
// special functions
// ---------------------------------------------------------------------------------------
// z for a doubled rank r2 = 2*rank (rank in 1..n, half-integers allowed), as the reference
// computes it (src/utils.jl:189-193 then :182): q = (r - 3/8)/(n + 1/4) in Float64, stored
// as T, z = norminvcdf(q) in T.
template <typename T> __device__ __forceinline__ T z_from_rank2(long long r2, long long n) {
  double q = (0.5 * (double)r2 - 0.375) / ((double)n + 0.25);
  if (sizeof(T) == 4) {
    float qf = (float)q;
    return (T)(float)normcdfinv((double)qf);
  }
  return (T)normcdfinv(q);
}

// log of the complete beta function
__device__ __forceinline__ double lbeta(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

// continued fraction for the incomplete beta function (modified Lentz)
__device__ inline double betacf(double a, double b, double x) {
  const double FPMIN = 1e-300, EPS = 1e-16;
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < FPMIN) d = FPMIN;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 20000; ++m) {
    double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}

// regularised incomplete beta I_x(a,b)
__device__ inline double betainc_reg(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  double lbt = a * log(x) + b * log1p(-x) - lbeta(a, b);
  if (x < (a + 1.0) / (a + b + 2.0)) return exp(lbt) * betacf(a, b, x) / a;
  return 1.0 - exp(lbt) * betacf(b, a, 1.0 - x) / b;
}

// inverse of I_x(a,b) = p  (StatsFuns.betainvcdf, call site src/mcse.jl:108-109).
// Safeguarded Newton from the normal approximation.  The continued fraction (the expensive
// part: O(sqrt(max(a,b))) terms) is evaluated ONCE, at the starting point; every later value of
// I is carried forward by integrating the density over the (small) Newton step with a 5-point
// Gauss-Legendre rule, I(x + h) = I(x) + int_x^{x+h} pdf.  A final continued-fraction check
// polishes the root if the carried value drifted.
__device__ inline double beta_logpdf(double a, double b, double lb, double x) {
  return (a - 1.0) * log(x) + (b - 1.0) * log1p(-x) - lb;
}
__device__ inline double beta_pdf_integral(double a, double b, double lb, double x0, double x1) {
  // 5-point Gauss-Legendre on [x0, x1]
  const double xm = 0.5 * (x0 + x1), xr = 0.5 * (x1 - x0);
  const double n1 = 0.5384693101056831, n2 = 0.9061798459386640;
  const double w0 = 0.5688888888888889, w1 = 0.4786286704993665, w2 = 0.2369268850561891;
  double s = w0 * exp(beta_logpdf(a, b, lb, xm));
  s += w1 * (exp(beta_logpdf(a, b, lb, xm - n1 * xr)) + exp(beta_logpdf(a, b, lb, xm + n1 * xr)));
  s += w2 * (exp(beta_logpdf(a, b, lb, xm - n2 * xr)) + exp(beta_logpdf(a, b, lb, xm + n2 * xr)));
  return s * xr;
}
__device__ inline double betainc_inv(double a, double b, double p) {
  if (!(a > 0.0) || !(b > 0.0) || !(p == p)) return CUDART_NAN;
  if (p <= 0.0) return 0.0;
  if (p >= 1.0) return 1.0;
  const double mu = a / (a + b);
  const double sd = sqrt(a * b / ((a + b) * (a + b) * (a + b + 1.0)));
  double x = mu + normcdfinv(p) * sd;
  double lo = 0.0, hi = 1.0;
  if (!(x > 0.0 && x < 1.0)) x = mu;
  const double lb = lbeta(a, b);
  const bool smooth = a > 2.0 && b > 2.0;   // density bounded and smooth: the quadrature carry is safe
  double Ix = betainc_reg(a, b, x);
  for (int it = 0; it < 60; ++it) {
    const double f = Ix - p;
    if (f > 0.0) hi = x; else lo = x;
    if (f == 0.0) break;
    const double dx = f / exp(beta_logpdf(a, b, lb, x));
    double xn = x - dx;
    bool bis = false;
    if (!(xn > lo && xn < hi)) { xn = 0.5 * (lo + hi); bis = true; }
    // with the carried value the iteration is only asked for ~1e-9: the exact polish below finishes it
    const bool done = fabs(xn - x) <= (smooth ? 1e-9 : 4e-16) * fabs(xn) || hi - lo <= 1e-17;
    if (smooth && !bis && fabs(xn - x) < 0.5 * sd) Ix += beta_pdf_integral(a, b, lb, x, xn);
    else Ix = betainc_reg(a, b, xn);
    x = xn;
    if (done) break;
  }
  if (smooth) {
    // one exact evaluation to remove any drift of the carried value
    // (Newton from 1e-9 is quadratically convergent: one exact step reaches ~1e-15)
    const double f = betainc_reg(a, b, x) - p;
    const double xn = x - f / exp(beta_logpdf(a, b, lb, x));
    if (xn > 0.0 && xn < 1.0) x = xn;
  }
  return x;
}

// Block-cooperative pair of inverses I_x(a,b) = p0, p1 (the two calls of src/mcse.jl:108-109 share
// a and b).  A single thread needs ~100 us of dependent FP64 work per inverse (continued fraction
// + Newton); here every thread integrates the density over one panel of the window
// [mode - 40 sd, mode + 40 sd] with an 8-point Gauss-Legendre rule, a block scan of the panel
// masses gives the distribution function at every panel edge (normalised by the total, so no
// log-gamma is needed and the density is evaluated relative to the mode, which keeps the
// exponent small), and the panel that brackets each p finishes with <= 8 Newton steps whose
// I(x) is the edge value + the same quadrature over [edge, x].  Agrees with the serial routine
// and with scipy.special.betaincinv to ~4e-15 relative (tests).  Falls back to the serial routine
// for small a or b (density not bell-shaped).  All threads must call; `red` = 34 doubles of
// shared scratch; results in out[0], out[1] (shared) after the closing barrier.
__device__ __forceinline__ double beta_rel_pdf(double am1, double bm1, double x0, double x) {
  const double d = x - x0;
  return exp(am1 * log1p(d / x0) + bm1 * log1p(-d / (1.0 - x0)));
}
__device__ inline double beta_rel_gl8(double am1, double bm1, double x0, double x1, double x2) {
  const double N0 = 0.1834346424956498, N1 = 0.5255324099163290, N2 = 0.7966664774136267, N3 = 0.9602898564975363;
  const double W0 = 0.3626837833783620, W1 = 0.3137066458778873, W2 = 0.2223810344533745, W3 = 0.1012285362903763;
  const double xm = 0.5 * (x1 + x2), xr = 0.5 * (x2 - x1);
  double s = W0 * (beta_rel_pdf(am1, bm1, x0, xm - xr * N0) + beta_rel_pdf(am1, bm1, x0, xm + xr * N0));
  s += W1 * (beta_rel_pdf(am1, bm1, x0, xm - xr * N1) + beta_rel_pdf(am1, bm1, x0, xm + xr * N1));
  s += W2 * (beta_rel_pdf(am1, bm1, x0, xm - xr * N2) + beta_rel_pdf(am1, bm1, x0, xm + xr * N2));
  s += W3 * (beta_rel_pdf(am1, bm1, x0, xm - xr * N3) + beta_rel_pdf(am1, bm1, x0, xm + xr * N3));
  return s * xr;
}
template <int THREADS>
__device__ void betainc_inv_pair_block(double a, double b, double p0, double p1, double* red, double* out) {
  static_assert(THREADS >= 64 && THREADS % 32 == 0 && THREADS <= 1024, "two warps run the Newton steps");
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool ok = a >= 16.0 && b >= 16.0 && p0 > 0.0 && p0 < 1.0 && p1 > 0.0 && p1 < 1.0;   // uniform over the block
  if (!ok) {
    if (tid == 0) out[0] = betainc_inv(a, b, p0);
    if (tid == 32) out[1] = betainc_inv(a, b, p1);
    __syncthreads();
    return;
  }
  const double am1 = a - 1.0, bm1 = b - 1.0;
  const double x0 = am1 / (am1 + bm1);   // mode
  const double sd = sqrt(a * b / ((a + b) * (a + b) * (a + b + 1.0)));
  const double L = fmax(0.0, x0 - 40.0 * sd), R = fmin(1.0, x0 + 40.0 * sd);
  const double h = (R - L) / (double)THREADS;
  const double xl = L + h * (double)tid, xr = tid == THREADS - 1 ? R : L + h * (double)(tid + 1);
  // a panel whose density is < 1e-30 of the mode's at its nearer edge carries no mass that matters
  // (the window is 80 sd wide: whole warps skip the quadrature)
  const double edge = xr < x0 ? xr : (xl > x0 ? xl : x0);
  const double mass = beta_rel_pdf(am1, bm1, x0, edge) < 1e-30 ? 0.0 : beta_rel_gl8(am1, bm1, x0, xl, xr);
  double incl = mass;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();   // red is free
  if (lane == 31) red[w] = incl;
  __syncthreads();
  double before = 0.0, total = 0.0;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) { const double t = red[i]; if (i < w) before += t; total += t; }
  const double cr = (before + incl) / total, cl = (before + incl - mass) / total;
  __syncthreads();   // every thread has read red[0 .. THREADS/32): reuse red[0..7] for the two brackets
  if (tid == 0) { red[0] = red[4] = x0; red[1] = red[5] = x0; red[2] = red[6] = 0.0; red[3] = red[7] = 0.0; }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double p = q == 0 ? p0 : p1;
    const bool mine = (cl <= p && p < cr) || (tid == THREADS - 1 && p >= cr);
    if (mine && cr > cl) { red[4 * q] = xl; red[4 * q + 1] = xr; red[4 * q + 2] = cl; red[4 * q + 3] = cr; }
  }
  __syncthreads();
  if (w < 2) {
    // warp q solves I(x) = p_q inside its bracketing panel: lanes 0..7 hold the Gauss-Legendre nodes of
    // [edge, x], lane 8 the density at x; one density evaluation of latency per Newton step
    const double p = w == 0 ? p0 : p1;
    const double bl = red[4 * w], br = red[4 * w + 1], bcl = red[4 * w + 2], bcr = red[4 * w + 3];
    double x = bl;
    if (bcr > bcl) {
      const double NODE[4] = {0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
      const double WGT[4] = {0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
      const double nd = lane < 8 ? ((lane & 1) ? NODE[(lane >> 1) & 3] : -NODE[(lane >> 1) & 3]) : 0.0;
      const double wt = lane < 8 ? WGT[(lane >> 1) & 3] : 0.0;
      x = bl + (br - bl) * (p - bcl) / (bcr - bcl);
      for (int it = 0; it < 8; ++it) {
        const double xm = 0.5 * (bl + x), xh = 0.5 * (x - bl);
        const double at = lane < 8 ? xm + xh * nd : x;
        double v = lane <= 8 ? beta_rel_pdf(am1, bm1, x0, at) : 0.0;
        const double fx = __shfl_sync(0xffffffffu, v, 8);
        v *= wt;
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const double integral = __shfl_sync(0xffffffffu, v, 0) * xh;
        const double dx = (bcl + integral / total - p) / (fx / total);
        x -= dx;
        x = x < bl ? bl : (x > br ? br : x);
        if (fabs(dx) <= 1e-16 * x) break;
      }
    }
    if (lane == 0) out[w] = x;
  }
  __syncthreads();
}

// Split-chain addressing (copyto_split!, src/utils.jl:13-41).  Split chain j = c*split + k
// of a slab (draws x chains, chain-contiguous) starts at element
//   c*draws + k*niter + min(k, draws % split)
// and holds niter = draws / split consecutive elements.
struct SplitGeom {
  int draws, chains, split, niter, nch, n, rem;
  __host__ __device__ SplitGeom() {}
  __host__ __device__ SplitGeom(int draws_, int chains_, int split_)
      : draws(draws_), chains(chains_), split(split_) {
    niter = draws / split;
    rem = draws % split;
    nch = chains * split;
    n = draws * chains;
  }
  __host__ __device__ __forceinline__ int chain_start(int j) const {
    int c = j / split, k = j - c * split;
    return c * draws + k * niter + (k < rem ? k : rem);
  }
};

}  // namespace mcd
