// ref_port.cpp — C++17/OpenMP restatement of the reference's ess_rhat path.
//
// TEST INFRASTRUCTURE ONLY (second, independent oracle + the timed CPU baseline).  Nothing in
// the product links or calls this; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do.
//
// It follows MCMCDiagnosticTools.jl v0.3.19 pass for pass (citations are /root/reference
// file:line) and deliberately keeps the reference's work profile: a sortperm-style tied
// ranking over the flattened slab for every rank-normalisation (src/utils.jl:175-184), one
// norminvcdf evaluation per element (:182), median / quantile by sorting a copy
// (:155, src/ess_rhat.jl:636,655), full-size temporaries per transform, and the lazy Geyer
// lag loop with one dot product per chain per lag (src/ess_rhat.jl:553-594).  The only thing
// added is `#pragma omp parallel for` over parameters: the reference itself is serial
// (SURVEY.md F2), so this is the "multithreaded CPU path" the north star asks to be timed.
//
// PARITY STATUS: no reference-output pin (no Julia in this image, and the reference's tests hold no
// literal ESS / R-hat vectors); pinned against the NumPy oracle (<= 1e-12), the committed fixtures
// (tests/golden/) and, through the oracle, the reference's known-answer tests and identities.
//
// Scope: ess / rhat / ess_rhat for kind in {basic,bulk,tail,rank} and the estimator ESS kinds
// (mean, median, std, mad, quantile), AutocovMethod and BDAAutocovMethod, Float64.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double NaN = std::numeric_limits<double>::quiet_NaN();

// StatsFuns.norminvcdf(p) = -erfcinv(2p)*sqrt(2).  Acklam's rational start + two Halley steps
// on erfc, which lands within an ulp or two of the exact value (checked against
// scipy.special.ndtri in tests/test_oracle_port.py).
double norminvcdf(double p) {
  if (!(p > 0.0)) return p == 0.0 ? -INFINITY : NaN;
  if (!(p < 1.0)) return p == 1.0 ? INFINITY : NaN;
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                             3.754408661907416e+00};
  const double plow = 0.02425;
  double x;
  if (p < plow) {
    double q = std::sqrt(-2 * std::log(p));
    x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  } else if (p <= 1 - plow) {
    double q = p - 0.5, r = q * q;
    x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
        (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
  } else {
    double q = std::sqrt(-2 * std::log1p(-p));
    x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  for (int it = 0; it < 2; ++it) {
    // e = Phi(x) - p, evaluated on the side that avoids cancellation
    double e = x < 0 ? 0.5 * std::erfc(-x * M_SQRT1_2) - p : (1.0 - p) - 0.5 * std::erfc(x * M_SQRT1_2);
    double u = e * std::sqrt(2 * M_PI) * std::exp(0.5 * x * x);
    x = x - u / (1 + 0.5 * x * u);
  }
  return x;
}

inline double jl_min(double a, double b) { return (a != a || b != b) ? NaN : (a < b ? a : b); }
inline double jl_max(double a, double b) { return (a != a || b != b) ? NaN : (a > b ? a : b); }

struct Geom {
  int draws, chains, split, niter, nch, n, rem;
  Geom(int d, int c, int s) : draws(d), chains(c), split(s) {
    niter = d / s; rem = d % s; nch = c * s; n = d * c;
  }
  // copyto_split! (src/utils.jl:13-41): start of split chain j inside the slab
  int start(int j) const { int c = j / split, k = j % split; return c * draws + k * niter + std::min(k, rem); }
};

// StatsBase.tiedrank (call site src/utils.jl:180)
void tiedrank(const double* x, int n, std::vector<int>& perm, double* rk) {
  perm.resize(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
    double u = x[a], v = x[b];  // isless: NaN last, -0.0 < 0.0
    if (u != u) return false;
    if (v != v) return true;
    if (u == v) return std::signbit(u) && !std::signbit(v);
    return u < v;
  });
  int s = 0;
  for (int e = 1; e <= n; ++e) {
    if (e == n || x[perm[e]] != x[perm[s]]) {
      double ar = (s + 1 + e) / 2.0;
      for (int i = s; i < e; ++i) rk[perm[i]] = ar;
      s = e;
    }
  }
}

// _rank_normalize! (src/utils.jl:175-193)
void rank_normalize(const double* x, int n, std::vector<int>& perm, double* z) {
  tiedrank(x, n, perm, z);
  for (int i = 0; i < n; ++i) z[i] = norminvcdf((z[i] - 0.375) / (n + 0.25));
}

// Statistics.median(vec(x)) (src/utils.jl:155)
double median(const double* x, int n, std::vector<double>& tmp) {
  for (int i = 0; i < n; ++i) if (x[i] != x[i]) return NaN;
  tmp.assign(x, x + n);
  if (n & 1) { std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end()); return tmp[n / 2]; }
  std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
  double b = tmp[n / 2];
  double a = *std::max_element(tmp.begin(), tmp.begin() + n / 2);
  return a / 2 + b / 2;
}

// Statistics.quantile type 7 (src/ess_rhat.jl:655); returns false if NaN present (Julia throws)
bool quantile(const double* x, int n, double p, std::vector<double>& tmp, double& q) {
  for (int i = 0; i < n; ++i) if (x[i] != x[i]) return false;
  tmp.assign(x, x + n);
  std::sort(tmp.begin(), tmp.end());
  if (n == 1) { q = tmp[0]; return true; }
  double aleph = std::fma((double)n, p, 1.0 - p);
  long j = (long)std::trunc(aleph);
  j = std::min<long>(std::max<long>(j, 1), n - 1);
  double g = std::min(std::max(aleph - (double)j, 0.0), 1.0);
  double a = tmp[j - 1], b = tmp[j];
  q = (std::isfinite(a) && std::isfinite(b)) ? a + g * (b - a) : (1 - g) * a + g * b;
  return true;
}

struct Scratch {
  std::vector<double> y, y2, samples, tmp, cmean, cvar;
  std::vector<int> perm;
};

// _rhat_basic! / _ess_rhat_basic! on one slab (src/ess_rhat.jl:362-409, 488-603)
void basic(const double* y, const Geom& g, bool want_ess, int method, int maxlag, bool relative, bool ess_nan,
           Scratch& sc, double& ess, double& rhat) {
  const int niter = g.niter, nch = g.nch;
  sc.samples.resize((size_t)niter * nch);
  sc.cmean.resize(nch); sc.cvar.resize(nch);
  double* s = sc.samples.data();
  for (int j = 0; j < nch; ++j) std::copy(y + g.start(j), y + g.start(j) + niter, s + (size_t)j * niter);
  double W = 0;
  for (int j = 0; j < nch; ++j) {
    const double* c = s + (size_t)j * niter;
    double m = 0;
    for (int t = 0; t < niter; ++t) m += c[t];
    m /= niter;
    double v = 0;
    for (int t = 0; t < niter; ++t) { double d = c[t] - m; v += d * d; }
    sc.cmean[j] = m; sc.cvar[j] = v / (niter - 1);
    W += sc.cvar[j];
  }
  W /= nch;
  double mm = 0;
  for (int j = 0; j < nch; ++j) mm += sc.cmean[j];
  mm /= nch;
  double bv = 0;
  for (int j = 0; j < nch; ++j) { double d = sc.cmean[j] - mm; bv += d * d; }
  bv /= (nch - (nch > 1 ? 1 : 0));
  const double var_plus = ((double)(niter - 1) / (double)niter) * W + bv;
  rhat = std::sqrt(var_plus / W);
  ess = NaN;
  if (!want_ess || ess_nan) return;
  for (int j = 0; j < nch; ++j) { double* c = s + (size_t)j * niter; for (int t = 0; t < niter; ++t) c[t] -= sc.cmean[j]; }
  const double inv = 1.0 / var_plus;
  auto mac = [&](int k) {
    double acc = 0;
    if (method == 2) {  // BDA (src/ess_rhat.jl:197-213)
      int n = niter - k;
      for (int j = 0; j < nch; ++j) {
        const double* c = s + (size_t)j * niter;
        double q = 0;
        for (int t = 0; t < n; ++t) { double d = c[t] - c[t + k]; q += d * d; }
        acc += q;
      }
      return W - (acc / nch) / (2.0 * n);
    }
    for (int j = 0; j < nch; ++j) {  // direct (src/ess_rhat.jl:161-179)
      const double* c = s + (size_t)j * niter;
      double q = 0;
      for (int t = 0; t < niter - k; ++t) q += c[t] * c[t + k];
      acc += q;
    }
    return (acc / nch) / niter;
  };
  auto rho = [&](int k) { return 1.0 - inv * (W - mac(k)); };
  double rho_odd = rho(1), rho_even = 1.0;
  double pt = rho_even + rho_odd, sum = pt;
  int k = 2;
  while (k < maxlag - 1) {
    rho_even = rho(k); rho_odd = rho(k + 1);
    double delta = rho_even + rho_odd;
    if (!(delta > 0)) break;
    pt = jl_min(delta, pt);
    sum += pt;
    k += 2;
  }
  rho_even = maxlag > 1 ? rho(k) : 0.0;
  double tau = jl_max(0.0, 2 * sum + jl_max(0.0, rho_even) - 1);
  const long ntotal = (long)niter * nch;
  ess = jl_min(1.0 / tau, (double)log10l((long double)ntotal));
  if (!relative) ess *= (double)ntotal;
}

void fold(const double* x, int n, Scratch& sc, double* out) {
  double med = median(x, n, sc.tmp);
  for (int i = 0; i < n; ++i) out[i] = std::fabs(x[i] - med);
}

}  // namespace

extern "C" {

// kind: 0 basic, 1 bulk, 2 tail, 3 rank.  method: 0 direct, 2 bda.  ess/rhat may be NULL.
// Returns 0, or -5 if a quantile was requested on data containing NaN, -1 on bad arguments.
int oracle_ess_rhat(const double* x, long draws, long chains, long params, int kind, int method, int split,
                    int maxlag, int relative, double tail_prob, double* ess, double* rhat, int nthreads) {
  if (draws <= 0 || chains <= 0 || split < 1 || (method != 0 && method != 2)) return -1;
  Geom g((int)draws, (int)chains, split);
  const bool want_ess = ess != nullptr;
  bool ess_nan = !(g.niter > 4);
  if (want_ess && !ess_nan) { if (maxlag <= 0) return -1; maxlag = std::min(maxlag, g.niter - 4); }
  int status = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    Scratch sc;
    sc.y.resize(g.n); sc.y2.resize(g.n);
#pragma omp for schedule(dynamic, 16)
    for (long p = 0; p < params; ++p) {
      const double* xp = x + (size_t)p * g.n;
      double e = NaN, r = NaN, e2, r2;
      switch (kind) {
        case 0: basic(xp, g, want_ess, method, maxlag, relative, ess_nan, sc, e, r); break;
        case 1:
          rank_normalize(xp, g.n, sc.perm, sc.y.data());
          basic(sc.y.data(), g, want_ess, method, maxlag, relative, ess_nan, sc, e, r);
          break;
        case 2: {
          if (want_ess) {
            double q;
            double ps[2] = {tail_prob / 2, 1 - tail_prob / 2};
            double es[2];
            for (int i = 0; i < 2; ++i) {
              if (!quantile(xp, g.n, ps[i], sc.tmp, q)) { status = -5; es[i] = NaN; continue; }
              for (int t = 0; t < g.n; ++t) sc.y[t] = xp[t] <= q ? 1.0 : 0.0;
              basic(sc.y.data(), g, true, method, maxlag, relative, ess_nan, sc, es[i], r2);
            }
            e = jl_min(es[0], es[1]);
          }
          if (rhat) {
            fold(xp, g.n, sc, sc.y2.data());
            rank_normalize(sc.y2.data(), g.n, sc.perm, sc.y.data());
            basic(sc.y.data(), g, false, method, maxlag, relative, ess_nan, sc, e2, r);
          }
          break;
        }
        case 3: {
          rank_normalize(xp, g.n, sc.perm, sc.y.data());
          basic(sc.y.data(), g, want_ess, method, maxlag, relative, ess_nan, sc, e, r);
          fold(xp, g.n, sc, sc.y2.data());
          rank_normalize(sc.y2.data(), g.n, sc.perm, sc.y.data());
          basic(sc.y.data(), g, false, method, maxlag, relative, ess_nan, sc, e2, r2);
          r = jl_max(r2, r);
          break;
        }
        default: break;
      }
      if (ess) ess[p] = e;
      if (rhat) rhat[p] = r;
    }
  }
  return status;
}

// estimator: 0 mean, 1 median, 2 std, 3 mad, 4 quantile(p)
int oracle_ess_estimator(const double* x, long draws, long chains, long params, int estimator, double prob,
                         int method, int split, int maxlag, int relative, double* ess, int nthreads) {
  if (draws <= 0 || chains <= 0 || split < 1 || !ess || (method != 0 && method != 2)) return -1;
  Geom g((int)draws, (int)chains, split);
  bool ess_nan = !(g.niter > 4);
  if (!ess_nan) { if (maxlag <= 0) return -1; maxlag = std::min(maxlag, g.niter - 4); }
  int status = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    Scratch sc;
    sc.y.resize(g.n); sc.y2.resize(g.n);
#pragma omp for schedule(dynamic, 16)
    for (long p = 0; p < params; ++p) {
      const double* xp = x + (size_t)p * g.n;
      double e = NaN, r;
      const double* proxy = xp;
      if (estimator == 1) {
        double med = median(xp, g.n, sc.tmp);
        for (int t = 0; t < g.n; ++t) sc.y[t] = xp[t] <= med ? 1.0 : 0.0;
        proxy = sc.y.data();
      } else if (estimator == 2) {
        double m = 0;
        for (int t = 0; t < g.n; ++t) m += xp[t];
        m /= g.n;
        for (int t = 0; t < g.n; ++t) { double d = xp[t] - m; sc.y[t] = d * d; }
        proxy = sc.y.data();
      } else if (estimator == 3) {
        fold(xp, g.n, sc, sc.y2.data());
        double med = median(sc.y2.data(), g.n, sc.tmp);
        for (int t = 0; t < g.n; ++t) sc.y[t] = sc.y2[t] <= med ? 1.0 : 0.0;
        proxy = sc.y.data();
      } else if (estimator == 4) {
        double q;
        if (!quantile(xp, g.n, prob, sc.tmp, q)) { status = -5; ess[p] = NaN; continue; }
        for (int t = 0; t < g.n; ++t) sc.y[t] = xp[t] <= q ? 1.0 : 0.0;
        proxy = sc.y.data();
      }
      basic(proxy, g, true, method, maxlag, relative, ess_nan, sc, e, r);
      ess[p] = e;
    }
  }
  return status;
}

void oracle_tiedrank(const double* x, long n, double* ranks) {
  std::vector<int> perm;
  tiedrank(x, (int)n, perm, ranks);
}

double oracle_norminvcdf(double p) { return norminvcdf(p); }

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
