// crank_emul.cu — CPU emulation of the counting-rank kernels (mcd_crank.cuh): every kernel is replayed thread by
// thread in launch order, with the atomics as plain updates, so the algorithm (bucket map, packed counters,
// scan, placement, exact tie handling, median selection) is checked against NumPy / SciPy without a GPU.
// Test infrastructure only; built by tests/test_crank_emul.py with nvcc (host code only is executed).
#include "../../mcmcdiagnostictools.jl_b200/csrc/mcd_crank.cuh"
#include <vector>
#include <algorithm>

using namespace mcd;

template <typename T>
static int emul(const T* x, long long n, long long pc, int bucket_factor, double* ranks, double* med, int* flags) {
  using K = typename CrKeyOf<T>::type;
  unsigned B = 1u << 16;
  while ((long long)B < (long long)bucket_factor * n && B < (1u << 28)) B <<= 1;
  CrWork<T> w;
  w.buckets = B; w.nw = B / 8; w.seg = (w.nw + CR_NSEG - 1) / CR_NSEG;
  std::vector<K> kmin(pc, ~(K)0), kmax(pc, 0);
  std::vector<CrMap<T>> map(pc);
  std::vector<int> flag(pc + 1, 0);
  std::vector<uint2> cw((size_t)pc * w.nw, make_uint2(0, 0));
  std::vector<unsigned> part((size_t)pc * CR_NSEG, 0), info((size_t)pc * n, 0);
  std::vector<T> srt((size_t)pc * n, (T)-12345);
  w.kmin = kmin.data(); w.kmax = kmax.data(); w.map = map.data(); w.flag = flag.data(); w.cw = cw.data();
  w.part = part.data(); w.info = info.data(); w.srt = srt.data();
  const long long tiles = (n + CR_TILE - 1) / CR_TILE;
  // 1. min / max: grid (tiles, pc) x CR_THREADS
  for (long long p = 0; p < pc; ++p)
    for (long long t = 0; t < tiles; ++t)
      for (int th = 0; th < CR_THREADS; ++th) {
        const long long t0 = t * CR_TILE, t1 = std::min<long long>(t0 + CR_TILE, n);
        K a = ~(K)0, b = 0;
        cr_minmax_body<T>(x + p * n, t0 + th, t1, CR_THREADS, a, b);
        cr_atomic_min<K>(&w.kmin[p], a); cr_atomic_max<K>(&w.kmax[p], b);
      }
  for (long long p = 0; p < pc; ++p) cr_setup_body<T>(w, p);
  // 2. count (threads of a CTA interleaved the way the hardware would not: reversed order, to make arrival
  // order differ from index order)
  for (long long p = 0; p < pc; ++p)
    for (long long t = tiles - 1; t >= 0; --t) {
      const long long t0 = t * CR_TILE, t1 = std::min<long long>(t0 + CR_TILE, n);
      for (int th = CR_THREADS - 1; th >= 0; --th)
        for (long long i = t0 + th; i < t1; i += CR_U * CR_THREADS) cr_count_body<T>(w, x, n, p, i, CR_THREADS, t1);
    }
  // 3. scan
  for (long long p = 0; p < pc; ++p) for (int s = 0; s < CR_NSEG; ++s) cr_scan1_body<T>(w, p, s);
  for (long long p = 0; p < pc; ++p) for (int s = 0; s < CR_NSEG; ++s) cr_scan3_body<T>(w, p, s);
  // 4. place
  for (long long p = 0; p < pc; ++p) {
    if (w.flag[p]) continue;
    for (long long t = 0; t < tiles; ++t) {
      const long long t0 = t * CR_TILE, t1 = std::min<long long>(t0 + CR_TILE, n);
      for (int th = 0; th < CR_THREADS; ++th)
        for (long long i = t0 + th; i < t1; i += CR_U * CR_THREADS) cr_place_body<T>(w, x, n, p, i, CR_THREADS, t1);
    }
  }
  // 5. rank, 6. median
  for (long long p = 0; p < pc; ++p) {
    flags[p] = w.flag[p];
    if (w.flag[p]) { med[p] = 0.0; continue; }
    for (long long t = 0; t < tiles; ++t) {
      const long long t0 = t * CR_TILE, t1 = std::min<long long>(t0 + CR_TILE, n);
      for (int th = 0; th < CR_THREADS; ++th)
        for (long long i = t0 + th; i < t1; i += CR_U * CR_THREADS) {
          long long r2[CR_U];
          cr_rank_body<T>(w, n, p, i, CR_THREADS, t1, r2);
          for (int u = 0; u < CR_U; ++u)
            if (i + (long long)u * CR_THREADS < t1) ranks[p * n + i + (long long)u * CR_THREADS] = 0.5 * (double)r2[u];
        }
    }
    med[p] = (double)cr_median_body<T>(w, n, p);
  }
  return 0;
}

extern "C" int crank_emul_f64(const double* x, long long n, long long pc, int bucket_factor, double* ranks, double* med, int* flags) {
  return emul<double>(x, n, pc, bucket_factor, ranks, med, flags);
}
extern "C" int crank_emul_f32(const float* x, long long n, long long pc, int bucket_factor, double* ranks, double* med, int* flags) {
  return emul<float>(x, n, pc, bucket_factor, ranks, med, flags);
}
