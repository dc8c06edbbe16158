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

// StatsFuns.norminvcdf(p) = -erfcinv(2p)*sqrt(2): one rational approximation per call, as Julia's erfcinv is
// (round 1 used Acklam's start + two Halley steps, each an erfc and an exp: a pessimistic CPU baseline).  Wichura's
// AS 241 (PPND16, Applied Statistics 37 (1988) 477-484), relative error about 1e-16; checked against
// scipy.special.ndtri in tests/test_oracle_port.py.
inline double horner8(const double* c, double x) {
  double r = c[7];
  for (int k = 6; k >= 0; --k) r = r * x + c[k];
  return r;
}
double norminvcdf(double p) {
  if (!(p > 0.0)) return p == 0.0 ? -INFINITY : NaN;
  if (!(p < 1.0)) return p == 1.0 ? INFINITY : NaN;
  static const double a[] = {3.3871328727963666080e0, 1.3314166789178437745e+2, 1.9715909503065514427e+3, 1.3731693765509461125e+4,
                             4.5921953931549871457e+4, 6.7265770927008700853e+4, 3.3430575583588128105e+4, 2.5090809287301226727e+3};
  static const double b[] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
                             2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4, 5.2264952788528545610e+3};
  static const double c[] = {1.42343711074968357734e0, 4.63033784615654529590e0, 5.76949722146069140550e0, 3.64784832476320460504e0,
                             1.27045825245236838258e0, 2.41780725177450611770e-1, 2.27238449892691845833e-2, 7.74545014278341407640e-4};
  static const double d[] = {1.0, 2.05319162663775882187e0, 1.67638483018380384940e0, 6.89767334985100004550e-1,
                             1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4, 1.05075007164441684324e-9};
  static const double e[] = {6.65790464350110377720e0, 5.46378491116411436990e0, 1.78482653991729133580e0, 2.96560571828504891230e-1,
                             2.65321895265761230930e-2, 1.24266094738807843860e-3, 2.71155556874348757815e-5, 2.01033439929228813265e-7};
  static const double f[] = {1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
                             7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7, 2.04426310338993978564e-15};
  const double q = p - 0.5;
  if (std::fabs(q) <= 0.425) {
    const double r = 0.180625 - q * q;
    return q * horner8(a, r) / horner8(b, r);
  }
  double r = std::sqrt(-std::log(q < 0 ? p : 1.0 - p));
  double v;
  if (r <= 5.0) { r -= 1.6; v = horner8(c, r) / horner8(d, r); }
  else { r -= 5.0; v = horner8(e, r) / horner8(f, r); }
  return q < 0 ? -v : v;
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
