// mcd_api.cu — C ABI of libmcmcdiag_b200.so (see include/mcmcdiag_b200.h).
// Context, program construction (which transform / reduce steps a reference call maps
// to), path selection (shared-memory slab kernel vs global-memory large-slab pipeline),
// host staging pipeline, lookup tables and the AR(1) generator.
#include "../../include/mcmcdiag_b200.h"
#include "mcd_common.cuh"
#include "mcd_slab.cuh"
#include "mcd_fast.cuh"
#include "mcd_rk2_api.cuh"
#include "mcd_big_api.cuh"
#include "mcd_fastgen.cuh"
#include "mcd_large.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace mcd;

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
struct mcd_ctx {
  int device = 0;
  std::vector<mcd_ctx*> children;   // non-empty: a multi-GPU group (mcd_create_multi); the group itself owns no device state
  std::vector<unsigned char> skip;  // per-parameter skip mask of the NEXT hot-path call (mcd_set_param_mask); empty = none
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_switch = nullptr;   // orders the per-context scratch across a change of stream
  std::mutex mu;
  std::string err;
  int sm_count = 0;
  int smem_optin = 0;
  // cached tables
  void* ztab = nullptr; long long ztab_n = 0; int ztab_dtype = -1; size_t ztab_cap = 0;
  void* tw = nullptr; int tw_n = 0; int tw_dtype = -1; size_t tw_cap = 0;
  unsigned* d_flags = nullptr;
  int* d_chain_inds = nullptr; size_t chain_inds_cap = 0;
  int* d_redo = nullptr; size_t redo_cap = 0;   // [0] = count, [1..] = parameter indices
  // staging / workspace
  void* stage[2] = {nullptr, nullptr}; size_t stage_cap[2] = {0, 0};
  void* d_out = nullptr; size_t out_cap = 0;
  void* d_arr = nullptr; size_t arr_cap = 0;
  void* work = nullptr; size_t work_cap = 0;
  // options
  int force_path = 0;
  long long h2d_chunk_bytes = 256ll << 20;
  long long workspace_bytes = 6ll << 30;
  int bucket_limit = 64;
  int fast_pad_smem = 0;   // developer knob: extra dynamic shared memory (lowers CTAs/SM)
  int slab_wide = 1;       // developer knob: 512-thread general kernel for slabs that fit one CTA per SM only
  int slab_three = 1;      // developer knob: 0 = never use the 80-register entry (three CTAs per SM) of the general kernel
  int fast_grid_mult = 0;  // developer knob: 0 = one CTA per parameter, k = persistent grid of k * 2 * SMs CTAs
  int use_rk2 = 1;         // developer knob: 0 = round-1 register-resident kernel instead of the TMA-staged one
  int use_big = 1;         // developer knob: 0 = never use the big-slab estimator kernel (mcd_big.cuh)
  int use_crank = 1;       // developer knob: 0 = large slabs are always ranked by segmented sort + binary searches
  int crank_factor = 4;    // counting rank: fine buckets per value
  long long crank_chunk = 0;   // counting rank: cap on the parameters per chunk (0 = workspace-bound)
  int fft_tc = 0;          // developer knob: columns per tile of the four-step FFT (0 = default)
  int fft_full = 0;        // developer knob: 1 = FFT length nextprod(2 niter - 1) (default: niter + maxlag)
  int fft_pair = 1;        // developer knob: 0 = four-step FFT with one transform per chain (round-1 data flow)
  int ztab_max_mb = 32;    // the counting rank looks z up in the ztab table while the table is at most this large
  // stats
  long long crank_chunks = 0, crank_fallbacks = 0;
  long long launches = 0, h2d_bytes = 0, d2h_bytes = 0;
  int last_path = 0;
};

static thread_local std::string g_create_err;

static int fail(mcd_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? MCD_ENOMEM : MCD_ECUDA, "%s failed: %s", \
                  #call, cudaGetErrorString(e_));                                             \
  } while (0)

// The library switches the calling thread's current CUDA device to the context's device for the duration of an entry
// point and restores it on return: a host application (torch, CUDA.jl, ...) keeps whatever device it had selected.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      changed = err == cudaSuccess;
    }
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(c_) DeviceGuard dev_guard_((c_)->device); CU(dev_guard_.err)

static int ensure_cap(mcd_ctx* ctx, void** p, size_t* cap, size_t need) {
  if (*cap >= need && *p) return MCD_OK;
  if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
  if (need == 0) need = 256;
  CU(cudaMalloc(p, need));
  *cap = need;
  return MCD_OK;
}

// ---------------------------------------------------------------------------------------
// small kernels: tables and the generator
// ---------------------------------------------------------------------------------------
// z table for the doubled rank r2 (2 .. 2n): entry r2 - 2 of the interleaved table, and a second,
// split copy starting at 2n whose first n entries are the integer ranks (the only ones untied data
// touches: half the cache footprint) followed by the n - 1 half-integer ranks.
template <typename T> __global__ void ztab_kernel(T* ztab, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < 2 * n - 1) {
    const T z = z_from_rank2<T>(i + 2, n);
    ztab[i] = z;
    ztab[2 * n + (i >> 1) + ((i & 1) ? n : 0)] = z;
  }
}

template <typename T> __global__ void twiddle_kernel(Cx<T>* tw, int N) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < N) {
    double s, c;
    sincospi(-2.0 * (double)k / (double)N, &s, &c);
    Cx<T> w; w.x = (T)c; w.y = (T)s;
    tw[k] = w;
  }
}

// Philox4x32-10 counter-based generator (Salmon et al., SC'11).
__device__ __forceinline__ void philox4x32_10(unsigned c[4], unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// One thread per (chain, param) series: test/helpers.jl:4-12.  Draw pairs (2i, 2i+1) come
// from one Philox block keyed by seed with counter (i, chain, param_lo, param_hi).
template <typename T>
__global__ void ar1_kernel(T* x, long long draws, long long chains, long long params,
                           long long param_offset, double phi, double sigma, unsigned long long seed) {
  long long sid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (sid >= chains * params) return;
  long long param = sid / chains, chain = sid % chains;
  unsigned long long gp = (unsigned long long)(param + param_offset);
  T* dst = x + sid * draws;
  double prev = 0.0;
  for (long long t0 = 0; t0 < draws; t0 += 2) {
    unsigned c[4] = {(unsigned)(t0 >> 1), (unsigned)chain, (unsigned)gp, (unsigned)(gp >> 32)};
    philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
    // two uniforms in (0,1) from 64 bits each
    const double u1 = ((double)c[0] * 4294967296.0 + (double)c[1] + 0.5) * 5.421010862427522e-20;
    const double u2 = ((double)c[2] * 4294967296.0 + (double)c[3] + 0.5) * 5.421010862427522e-20;
    double r = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    double e0 = r * co, e1 = r * s;
    double v0 = (t0 == 0) ? sigma * e0 : fma(phi, prev, sigma * e0);
    dst[t0] = (T)v0;
    prev = (double)(T)v0;
    if (t0 + 1 < draws) {
      double v1 = fma(phi, prev, sigma * e1);
      dst[t0 + 1] = (T)v1;
      prev = (double)(T)v1;
    }
  }
}

// ---------------------------------------------------------------------------------------
// program = what one reference call asks for
// ---------------------------------------------------------------------------------------
struct Program {
  int nsteps = 0;
  Step steps[MAX_STEPS];
  int combine = CB_PLAIN;
  int method = 0, maxlag = 1, relative = 0, ess_nan = 0;
  double mcse_p = 0.5;
  bool want_ess = false, want_rhat = false, want_arr = false;
  int arr_elem_bytes = 0;  // element size of arr_out (0 = none)
  const int32_t* chain_inds = nullptr; long long cps = 0, nsuper = 0;
  void add(int tr, int rd, double p = 0.0, int p_f32 = 0) {
    steps[nsteps].transform = tr; steps[nsteps].reduce = rd; steps[nsteps].p = p; steps[nsteps].p_f32 = p_f32;
    ++nsteps;
  }
};

// One unit of work of a call: a program with its per-parameter outputs, or (moments) the plain
// mean / corrected standard deviation over all draws and chains of each parameter.
enum JobRole { ROLE_NONE = 0, ROLE_MOMENTS, ROLE_MCSE_MEAN, ROLE_MCSE_STD, ROLE_BULK_RHAT, ROLE_TAIL };
struct Job {
  Program pg;
  int role = ROLE_NONE;   // set by mcd_summary: lets the driver fuse the jobs into one kernel
  bool moments = false;
  void* out0 = nullptr;   // ess / mcse / mean
  void* out1 = nullptr;   // rhat / std
  void* arr = nullptr;    // per-element output of the transform calls
};

// NaN for the outputs of skipped parameters (mcd_set_param_mask) when the outputs live in device memory.
template <typename T> __global__ void fill_kernel(T* out, long long n, T v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = v;
}

// Statistics.mean(x; dims=(1,2)) and Statistics.std(x; dims=(1,2)) (the quantities src/mcse.jl:50
// uses), one CTA per parameter, two passes in double; NaN propagates.
template <typename T>
__global__ void __launch_bounds__(256) moments_kernel(const T* __restrict__ x, long long params, long long n,
                                                     T* __restrict__ mean_out, T* __restrict__ std_out) {
  __shared__ double red[8];
  __shared__ double bc;
  for (long long p = blockIdx.x; p < params; p += gridDim.x) {
    const T* src = x + p * n;
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s += (double)__ldg(&src[i]);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < 8; ++i) t += red[i]; bc = t / (double)n; }
    __syncthreads();
    const double m = bc;
    double q = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) { const double d = (double)__ldg(&src[i]) - m; q = fma(d, d, q); }
    q = warp_sum(q);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0; for (int i = 0; i < 8; ++i) t += red[i];
      if (mean_out) mean_out[p] = (T)m;
      if (std_out) std_out[p] = (T)sqrt(t / (double)(n - 1));
    }
    __syncthreads();
  }
}

// Mean and corrected variance of every split chain of every parameter (the per-chain quantities of
// src/ess_rhat.jl:387-399 and of src/gelmandiag.jl:9-17), one warp per (parameter, split chain), two
// passes in double.  Outputs are (nch, params) column-major.
template <typename T>
__global__ void __launch_bounds__(256) chain_moments_kernel(const T* __restrict__ x, long long params, SplitGeom g,
                                                           T* __restrict__ mean_out, T* __restrict__ var_out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long item = warp0; item < params * g.nch; item += nwarps) {
    const long long p = item / g.nch;
    const int j = (int)(item - p * g.nch);
    const T* src = x + p * (long long)g.n + g.chain_start(j);
    double s = 0.0;
    for (int t = lane; t < g.niter; t += 32) s += (double)__ldg(&src[t]);
    s = warp_sum(s);
    const T m = (T)(s / (double)g.niter);
    double q = 0.0;
    for (int t = lane; t < g.niter; t += 32) { const T d = __ldg(&src[t]) - m; q = fma((double)d, (double)d, q); }
    q = warp_sum(q);
    if (lane == 0) {
      if (mean_out) mean_out[item] = m;
      if (var_out) var_out[item] = (T)(q / (double)(g.niter - 1));
    }
  }
}

// bfmi (src/bfmi.jl:36-43): mean(abs2, diff(energy)) / var(energy) per chain, one warp per chain.
template <typename T>
__global__ void __launch_bounds__(256) bfmi_kernel(const T* __restrict__ e, long long draws, long long chains,
                                                  T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = warp0; c < chains; c += nwarps) {
    const T* src = e + c * draws;
    double s = 0.0;
    for (long long t = lane; t < draws; t += 32) s += (double)__ldg(&src[t]);
    const T m = (T)(warp_sum(s) / (double)draws);
    double q = 0.0, d2 = 0.0;
    for (long long t = lane; t < draws; t += 32) {
      const T v = __ldg(&src[t]);
      const T d = v - m;
      q = fma((double)d, (double)d, q);
      if (t + 1 < draws) { const T df = __ldg(&src[t + 1]) - v; d2 = fma((double)df, (double)df, d2); }
    }
    q = warp_sum(q); d2 = warp_sum(d2);
    if (lane == 0) out[c] = (T)(d2 / (double)(draws - 1)) / (T)(q / (double)(draws - 1));
  }
}

// log10(oftype(one(T), ntotal)) (src/ess_rhat.jl:514), correctly rounded: glibc's double
// log10 is not (log10(40.0) is off by one ulp), so evaluate in long double and round once.
template <typename T> static T rel_ess_max_of(long long ntotal) {
  if (sizeof(T) == 8) return (T)(double)log10l((long double)ntotal);
  return (T)(float)log10((double)(float)ntotal);
}

static int next_pow2(long long v) { int p = 1; while (p < v) p <<= 1; return p; }

static long long nextprod23(long long n) {
  long long best = -1;
  for (long long p3 = 1;; p3 *= 3) {
    long long v = p3;
    while (v < n) v *= 2;
    if (best < 0 || v < best) best = v;
    if (p3 >= n) break;
  }
  return best;
}

static int ensure_ztab(mcd_ctx* ctx, int dtype, long long n) {
  if (ctx->ztab && ctx->ztab_n == n && ctx->ztab_dtype == dtype) return MCD_OK;
  size_t ts = dtype == MCD_F64 ? 8 : 4;
  int rc = ensure_cap(ctx, &ctx->ztab, &ctx->ztab_cap, (size_t)(4 * n) * ts);
  if (rc) return rc;
  long long cnt = 2 * n - 1;
  int blocks = (int)((cnt + 255) / 256);
  if (dtype == MCD_F64) ztab_kernel<double><<<blocks, 256, 0, ctx->stream>>>((double*)ctx->ztab, n);
  else ztab_kernel<float><<<blocks, 256, 0, ctx->stream>>>((float*)ctx->ztab, n);
  ctx->launches++;
  CU(cudaGetLastError());
  ctx->ztab_n = n; ctx->ztab_dtype = dtype;
  return MCD_OK;
}

static int ensure_twiddle(mcd_ctx* ctx, int dtype, int N) {
  if (ctx->tw && ctx->tw_n == N && ctx->tw_dtype == dtype) return MCD_OK;
  size_t ts = dtype == MCD_F64 ? 16 : 8;
  int rc = ensure_cap(ctx, &ctx->tw, &ctx->tw_cap, (size_t)N * ts);
  if (rc) return rc;
  int blocks = (N + 255) / 256;
  if (dtype == MCD_F64) twiddle_kernel<double><<<blocks, 256, 0, ctx->stream>>>((Cx<double>*)ctx->tw, N);
  else twiddle_kernel<float><<<blocks, 256, 0, ctx->stream>>>((Cx<float>*)ctx->tw, N);
  ctx->launches++;
  CU(cudaGetLastError());
  ctx->tw_n = N; ctx->tw_dtype = dtype;
  return MCD_OK;
}

static bool program_needs_ranks(const Program& pg) {
  for (int s = 0; s < pg.nsteps; ++s) {
    int t = pg.steps[s].transform;
    if (t == TR_RANKNORM || t == TR_FOLD_RANKNORM) return true;
  }
  return false;
}

constexpr int SLAB_THREADS = 256;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Lay the slab kernel's shared memory out; returns total bytes.
template <typename T>
static size_t slab_layout(SlabArgs<T>& a, const Program& pg, int threads = SLAB_THREADS) {
  const size_t ts = sizeof(T);
  const int n = a.g.n;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (int)o; };
  a.offX = take((size_t)n * ts);
  // single-step programs never look at the raw slab again: transform in place (Y aliases X)
  const bool alias_xy = pg.nsteps == 1 && pg.combine == CB_PLAIN;
  a.offY = alias_xy ? a.offX : take((size_t)n * ts);
  a.offK = take((size_t)n * ts);
  a.offCNT = take((size_t)(a.nbuckets + threads) * 4);
  size_t view_bytes = off - (size_t)a.offK;
  a.offCM = take((size_t)a.g.nch * ts);
  a.offCV = take((size_t)a.g.nch * ts);
  a.offGAM = take((size_t)(a.maxlag + 1 + LAG_BATCH) * ts);
  a.offGSUM = take(pg.method == MCD_AUTOCOV_FFT ? (size_t)(a.maxlag + 1) * 8 : 0);
  size_t part = (size_t)(threads / 32) * LAG_BATCH;
  if ((size_t)(2 * pg.nsuper) > part) part = (size_t)(2 * pg.nsuper);
  a.offPART = take(part * 8);
  a.offFFT = a.offK;
  if (pg.method == MCD_AUTOCOV_FFT && pg.want_ess) {
    size_t need = (size_t)a.fft_n * (2 * 2 + 1) * ts;   // two complex buffers + the summed power spectrum
    if (need > view_bytes) a.offFFT = take(need);
  }
  a.offMISC = take(sizeof(Misc));
  return off;
}

template <typename T>
static int run_slab(mcd_ctx* ctx, const T* dx, long long params, const SplitGeom& g, const Program& pg,
                    T* d_ess, T* d_rhat, void* d_arr, bool* handled, const int* redo_list = nullptr,
                    const int* redo_count = nullptr) {
  *handled = false;
  SlabArgs<T> a;
  memset(&a, 0, sizeof a);
  a.x = dx; a.params = params; a.g = g;
  a.nsteps = pg.nsteps;
  for (int s = 0; s < pg.nsteps; ++s) a.steps[s] = pg.steps[s];
  a.combine = pg.combine; a.method = pg.method; a.maxlag = pg.maxlag; a.relative = pg.relative;
  a.ess_nan = pg.ess_nan;
  const long long ntotal = (long long)g.niter * g.nch;
  a.rel_ess_max = rel_ess_max_of<T>(ntotal);
  a.ess_out = d_ess; a.rhat_out = d_rhat; a.arr_out = d_arr;
  a.nbuckets = std::min(std::max(next_pow2(g.n), SLAB_THREADS), 16384);
  a.bucket_limit = ctx->bucket_limit;
  // transform length: lags <= maxlag are free of wrap-around once N >= niter + maxlag (see run_large in mcd_large.cuh)
  const long long fft_need = ctx->fft_full ? 2ll * g.niter - 1 : std::min<long long>(2ll * g.niter - 1, (long long)g.niter + pg.maxlag);
  a.fft_n = (pg.method == MCD_AUTOCOV_FFT && pg.want_ess && !pg.ess_nan) ? (int)nextprod23(fft_need) : 0;
  a.cps = (int)pg.cps; a.nsuper = (int)pg.nsuper;
  a.mcse_p = pg.mcse_p;
  a.flags = ctx->d_flags;
  a.redo_list = redo_list; a.redo_count = redo_count;
  if (g.nch > 4096 || pg.nsuper > 4096) return MCD_OK;
  size_t smem = slab_layout<T>(a, pg);
  // two CTAs per SM beat finer buckets: halve the bucket count when that is what it takes to fit two
  const size_t two_cta = ((size_t)ctx->smem_optin + 1024 - 2 * 1024) / 2;   // (228 KB - 1 KB per CTA) / 2
  if (smem > two_cta && a.nbuckets > 2 * SLAB_THREADS) {
    const int full = a.nbuckets;
    a.nbuckets = full / 2;
    const size_t s2 = slab_layout<T>(a, pg);
    if (s2 <= two_cta) smem = s2; else { a.nbuckets = full; smem = slab_layout<T>(a, pg); }
  }
  // a slab that leaves room for one CTA per SM only gets 16 warps instead of 8 to hide latency
  int threads = SLAB_THREADS;
  if (smem > two_cta && g.n >= 4096 && !redo_list && ctx->slab_wide) {
    const size_t s512 = slab_layout<T>(a, pg, 512);
    if (s512 <= (size_t)ctx->smem_optin) { threads = 512; smem = s512; } else smem = slab_layout<T>(a, pg);
  }
  if (smem > (size_t)ctx->smem_optin) return MCD_OK;  // not handled: caller uses the large path

  const int dtype = sizeof(T) == 8 ? MCD_F64 : MCD_F32;
  if (program_needs_ranks(pg)) {
    int rc = ensure_ztab(ctx, dtype, g.n);
    if (rc) return rc;
    a.ztab = (const T*)ctx->ztab;
  }
  if (a.fft_n) {
    int rc = ensure_twiddle(ctx, dtype, a.fft_n);
    if (rc) return rc;
    a.twiddle = ctx->tw;
  }
  if (pg.chain_inds) a.chain_inds = ctx->d_chain_inds;

  const bool three = ctx->slab_three && 3 * (smem + 1024) <= (size_t)ctx->smem_optin + 1024;   // three CTAs of this slab fit an SM
  auto kern = threads == 512 ? slab_kernel_wide<T, 512> : (three ? slab_kernel<T, SLAB_THREADS> : slab_kernel_two<T, SLAB_THREADS>);
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = std::min<long long>(params, 1ll << 30);
  if (redo_list) grid = std::min<long long>(grid, 2ll * ctx->sm_count);
  if (grid > 0) {
    kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    ctx->launches++;
    CU(cudaGetLastError());
  }
  if (!redo_list) ctx->last_path = 1;
  *handled = true;
  return MCD_OK;
}

// The register-resident fast kernel (mcd_fast.cuh) for 8 split chains of <= 512 draws.
// type-7 quantile position and weight (Statistics.quantile, call site src/ess_rhat.jl:655): the two
// order statistics to capture (slots slot, slot + 1) and the threshold built from them
template <typename T>
static void plan_quantile(FastGenArgs<T>& a, int n, double p, int f32, int slot, int thr_index) {
  long long j; double gq;
  if (f32) {
    const float pf = (float)p, mm = (float)(1.0 - (double)pf);
    const float aleph = std::fmaf((float)n, pf, mm);
    j = (long long)std::trunc(aleph);
    j = std::min<long long>(std::max<long long>(j, 1), n - 1);
    float gf = aleph - (float)j;
    gf = gf < 0.f ? 0.f : (gf > 1.f ? 1.f : gf);
    gq = (double)gf;
  } else {
    const double aleph = std::fma((double)n, p, 1.0 - p);
    j = (long long)std::trunc(aleph);
    j = std::min<long long>(std::max<long long>(j, 1), n - 1);
    gq = aleph - (double)j;
    gq = gq < 0.0 ? 0.0 : (gq > 1.0 ? 1.0 : gq);
  }
  a.cap_pos[slot] = (int)(j - 1); a.cap_pos[slot + 1] = (int)j;
  a.thr[thr_index].quantile = 1; a.thr[thr_index].capA = slot; a.thr[thr_index].capB = slot + 1;
  a.thr[thr_index].f32 = f32; a.thr[thr_index].g = gq;
}

template <typename T>
static int run_fast(mcd_ctx* ctx, const T* dx, long long params, const SplitGeom& g, const Program& pg,
                    T* d_ess, T* d_rhat, bool* handled) {
  *handled = false;
  if (g.nch < 1 || g.nch > FAST_NCH || g.rem != 0 || g.niter < 2 || g.niter > FAST_MAXITER || params >= (1ll << 31)) return MCD_OK;
  if (pg.want_arr || pg.chain_inds) return MCD_OK;
  const bool ess_live = pg.want_ess && !pg.ess_nan;
  if (ess_live && pg.method != MCD_AUTOCOV_DIRECT) return MCD_OK;
  FastArgs<T> a;
  memset(&a, 0, sizeof a);
  bool lean = true;
  const Step& s0 = pg.steps[0];
  if (pg.nsteps == 1 && pg.combine == CB_PLAIN && (s0.transform == TR_NONE || s0.transform == TR_RANKNORM) &&
      (s0.reduce == RD_ESS_RHAT || s0.reduce == RD_RHAT)) {
    a.do_bulk = 1; a.rank_x = s0.transform == TR_RANKNORM; a.want_ess = s0.reduce == RD_ESS_RHAT;
  } else if (pg.nsteps == 1 && pg.combine == CB_PLAIN && s0.transform == TR_FOLD_RANKNORM && s0.reduce == RD_RHAT) {
    a.do_tail = 1;
  } else if (pg.nsteps == 2 && pg.combine == CB_RANK && s0.transform == TR_RANKNORM &&
             (s0.reduce == RD_ESS_RHAT || s0.reduce == RD_RHAT) && pg.steps[1].transform == TR_FOLD_RANKNORM &&
             pg.steps[1].reduce == RD_RHAT) {
    a.do_bulk = 1; a.rank_x = 1; a.want_ess = s0.reduce == RD_ESS_RHAT; a.do_tail = 1;
  } else lean = false;
  // the TMA-staged kernel takes 1..8 split chains (a slab must be a whole number of 16-byte units for the bulk copy);
  // the register-resident kernels exactly 8
  const bool rk2_ok = lean && ctx->use_rk2 && ((uintptr_t)dx & 15u) == 0 && pg.maxlag <= RK_MAXLAG_CAP &&
                      ((size_t)g.n * sizeof(T)) % 16 == 0;
  if (g.nch != FAST_NCH && !rk2_ok) return MCD_OK;
  a.nch = g.nch;
  FastGenArgs<T> ga;
  memset(&ga, 0, sizeof ga);
  if (!lean) {
  FastGenArgs<T>& a = ga;
  const int n = g.n;
  const int mA = (n & 1) ? n / 2 : n / 2 - 1, mB = n / 2;
  a.cap_pos[0] = mA; a.cap_pos[1] = mB;
  auto quantile_plan = [&](double p, int f32, int slot, int thr_index) { plan_quantile<T>(a, n, p, f32, slot, thr_index); };
  auto ind_red = [&](int i, int thr_index) { a.p0_red[i].src = FS_IND; a.p0_red[i].want_ess = 1; a.p0_red[i].thr = thr_index; };
  const bool plain1 = pg.nsteps == 1 && pg.combine == CB_PLAIN;
  if (pg.nsteps == 1 && pg.combine == CB_MCSE_MEAN && s0.transform == TR_NONE && s0.reduce == RD_ESS_RHAT) {
    a.p0_nred = 1; a.p0_red[0].src = FS_X; a.p0_red[0].want_ess = 1; a.ess_mode = 1; a.mcse_mode = 1; a.need_side = 1;
  } else if (pg.nsteps == 1 && pg.combine == CB_MCSE_STD && s0.transform == TR_STDPROXY && s0.reduce == RD_ESS_RHAT) {
    a.p0_nred = 1; a.p0_red[0].src = FS_SQDEV; a.p0_red[0].want_ess = 1; a.ess_mode = 1; a.mcse_mode = 2; a.need_side = 1;
  } else if (pg.nsteps == 1 && pg.combine == CB_MCSE_QUANTILE && s0.reduce == RD_ESS_RHAT &&
             (s0.transform == TR_IND_MEDIAN || s0.transform == TR_IND_QUANTILE)) {
    // _mcse_quantile (src/mcse.jl:96-118): indicator ESS, then two order statistics around p n from the kept window
    a.p0_rank = 1; a.p0_nred = 1; ind_red(0, 0); a.ess_mode = 1; a.mcse_mode = 3; a.mcse_p = pg.mcse_p;
    if (s0.transform == TR_IND_MEDIAN) { a.ncap = 2; a.nthr = 1; a.thr[0].quantile = 0; a.thr[0].capA = 0; a.thr[0].capB = 1; }
    else { a.ncap = 4; a.nthr = 1; quantile_plan(s0.p, s0.p_f32, 2, 0); }
    const long long centre = std::llround(pg.mcse_p * (double)n);
    a.win_lo = (int)std::min<long long>(std::max<long long>(centre - FASTGEN_WIN / 2, 0), std::max<long long>(n - FASTGEN_WIN, 0));
  } else if ((pg.combine == CB_TAIL && pg.nsteps == 3) || (pg.combine == CB_TAIL_ESS && pg.nsteps == 2)) {
    // _ess(Val(:tail)): min of the two quantile-indicator ESS (src/ess_rhat.jl:301-311) [+ tail R-hat]
    if (s0.transform != TR_IND_QUANTILE || pg.steps[1].transform != TR_IND_QUANTILE) return MCD_OK;
    a.p0_rank = 1; a.ncap = 6; a.nthr = 2; a.p0_nred = 2;
    quantile_plan(s0.p, s0.p_f32, 2, 0); quantile_plan(pg.steps[1].p, pg.steps[1].p_f32, 4, 1);
    ind_red(0, 0); ind_red(1, 1);
    a.ess_mode = 2;
    if (pg.combine == CB_TAIL) {
      if (pg.steps[2].transform != TR_FOLD_RANKNORM || pg.steps[2].reduce != RD_RHAT) return MCD_OK;
      a.do_fold = 1; a.p1_red.src = FS_RANKZ; a.rhat_mode = 2;
    }
  } else if (plain1 && s0.reduce == RD_ESS_RHAT && !pg.want_rhat && s0.transform == TR_IND_MEDIAN) {
    a.p0_rank = 1; a.ncap = 2; a.nthr = 1; a.thr[0].quantile = 0; a.thr[0].capA = 0; a.thr[0].capB = 1;
    a.p0_nred = 1; ind_red(0, 0); a.ess_mode = 1;
  } else if (plain1 && s0.reduce == RD_ESS_RHAT && !pg.want_rhat && s0.transform == TR_IND_QUANTILE) {
    a.p0_rank = 1; a.ncap = 4; a.nthr = 1; quantile_plan(s0.p, s0.p_f32, 2, 0);
    a.p0_nred = 1; ind_red(0, 0); a.ess_mode = 1;
  } else if (plain1 && s0.reduce == RD_ESS_RHAT && !pg.want_rhat && s0.transform == TR_STDPROXY) {
    a.p0_nred = 1; a.p0_red[0].src = FS_SQDEV; a.p0_red[0].want_ess = 1; a.ess_mode = 1; a.need_side = 1;
  } else if (plain1 && s0.reduce == RD_ESS_RHAT && !pg.want_rhat && s0.transform == TR_FOLD_IND_MEDIAN) {
    a.p0_rank = 1; a.ncap = 2; a.do_fold = 1; a.p1_red.src = FS_IND; a.p1_red.want_ess = 1; a.ess_mode = 4;
  } else return MCD_OK;
  }
  a.x = ga.x = dx; a.params = ga.params = params; a.niter = ga.niter = g.niter;
  a.maxlag = ga.maxlag = pg.maxlag; a.relative = ga.relative = pg.relative; a.ess_nan = ga.ess_nan = pg.ess_nan;
  a.rel_ess_max = ga.rel_ess_max = rel_ess_max_of<T>((long long)g.niter * g.nch);
  a.ess_out = ga.ess_out = d_ess; a.rhat_out = ga.rhat_out = d_rhat;
  const int dtype = sizeof(T) == 8 ? MCD_F64 : MCD_F32;
  if (lean ? (a.rank_x || a.do_tail) : (ga.p0_rank || ga.do_fold)) {
    int rc = ensure_ztab(ctx, dtype, g.n);
    if (rc) return rc;
    a.ztab = ga.ztab = (const T*)ctx->ztab + 2 * (size_t)g.n;   // split layout
  }
  int rc = ensure_cap(ctx, (void**)&ctx->d_redo, &ctx->redo_cap, (size_t)(params + 1) * sizeof(int));
  if (rc) return rc;
  a.redo_count = ga.redo_count = ctx->d_redo; a.redo_list = ga.redo_list = ctx->d_redo + 1;
  CU(cudaMemsetAsync(ctx->d_redo, 0, sizeof(int), ctx->stream));
  if (rk2_ok) {
    // persistent CTAs, two per SM, each streaming its parameters through the bulk-copy pipeline
    const int mult = ctx->fast_grid_mult > 0 ? ctx->fast_grid_mult : 1;
    const unsigned grid = (unsigned)std::min<long long>(params, (long long)mult * 2 * ctx->sm_count);
    CU(rk2_launch<T>(a, grid, (size_t)ctx->fast_pad_smem, ctx->stream));
  } else if (lean) {
    const size_t smem = fast_smem_bytes<T>(pg.maxlag) + (size_t)ctx->fast_pad_smem;
    auto kern = g.niter > 32 * (FAST_EPT - 1) ? fast_kernel<T, true> : fast_kernel<T, false>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = ctx->fast_grid_mult ? (unsigned)std::min<long long>(params, (long long)ctx->fast_grid_mult * 2 * ctx->sm_count) : (unsigned)params;
    kern<<<grid, FAST_THREADS, smem, ctx->stream>>>(a);
  } else {
    const size_t smem = fast_smem_bytes<T>(pg.maxlag) + FASTGEN_EXTRA_SMEM + (size_t)ctx->fast_pad_smem;
    const bool lng = g.niter > 32 * (FAST_EPT - 1);
    auto kern = ga.mcse_mode == 3 ? (lng ? fastgen_kernel<T, true, true> : fastgen_kernel<T, true, false>)
                                  : (lng ? fastgen_kernel<T, false, true> : fastgen_kernel<T, false, false>);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)params, FAST_THREADS, smem, ctx->stream>>>(ga);
  }
  ctx->launches++;
  CU(cudaGetLastError());
  // slabs the fast kernel declined (NaN, infinite range, heavy ties): general slab kernel
  bool h2 = false;
  rc = run_slab<T>(ctx, dx, params, g, pg, d_ess, d_rhat, nullptr, &h2, ctx->d_redo + 1, ctx->d_redo);
  if (rc) return rc;
  if (!h2) return fail(ctx, MCD_EUNSUPPORTED, "internal: redo kernel unavailable");
  ctx->last_path = 3;
  *handled = true;
  return MCD_OK;
}

// The jobs of mcd_summary in ONE launch of the register-resident kernel (8 split chains of <= 512
// draws, direct autocovariance): x is read from HBM once and ranked twice (raw, folded) for all
// seven columns.  o[j][0..1] are the device outputs of job j.
template <typename T>
static int run_fast_summary(mcd_ctx* ctx, const T* dx, long long params, const SplitGeom& g, const Job* jobs, int njobs,
                            T* const (*o)[2], bool* handled) {
  *handled = false;
  if (g.nch != FAST_NCH || g.rem != 0 || g.niter < 2 || g.niter > FAST_MAXITER || params >= (1ll << 31) || params == 0) return MCD_OK;
  FastGenArgs<T> a;
  memset(&a, 0, sizeof a);
  a.sum_mode = 1; a.need_side = 1;
  a.s_bulk = a.s_tlo = a.s_thi = a.s_mean = a.s_std = -1;
  const int n = g.n;
  a.cap_pos[0] = (n & 1) ? n / 2 : n / 2 - 1; a.cap_pos[1] = n / 2;
  const Program* ess_pg = nullptr;
  int bulk = -1, tail = -1, mm = -1, ms = -1, mom = -1;
  for (int j = 0; j < njobs; ++j) {
    switch (jobs[j].role) {
      case ROLE_MOMENTS: mom = j; break;
      case ROLE_MCSE_MEAN: mm = j; break;
      case ROLE_MCSE_STD: ms = j; break;
      case ROLE_BULK_RHAT: bulk = j; break;
      case ROLE_TAIL: tail = j; break;
      default: return MCD_OK;
    }
    if (jobs[j].role != ROLE_MOMENTS && jobs[j].pg.want_ess) {
      if (!jobs[j].pg.ess_nan && jobs[j].pg.method != MCD_AUTOCOV_DIRECT) return MCD_OK;
      ess_pg = &jobs[j].pg;
    }
  }
  if (bulk < 0 && tail < 0 && mm < 0 && ms < 0) return MCD_OK;   // moments alone: the plain kernel is HBM-bound already
  int nred = 0;
  if (bulk >= 0) {   // must be the first reduction of its pass (FS_RANKZ)
    a.p0_rank = 1; a.s_bulk = nred;
    a.p0_red[nred].src = FS_RANKZ; a.p0_red[nred].want_ess = jobs[bulk].pg.want_ess ? 1 : 0; ++nred;
    a.col[4] = o[bulk][0];
    if (jobs[bulk].pg.want_rhat) { a.col[6] = o[bulk][1]; a.do_fold = 1; a.p1_red.src = FS_RANKZ; }
  }
  if (tail >= 0) {
    const Program& pg = jobs[tail].pg;
    if (pg.nsteps != 2 || pg.steps[0].transform != TR_IND_QUANTILE || pg.steps[1].transform != TR_IND_QUANTILE) return MCD_OK;
    a.p0_rank = 1; a.ncap = 6; a.nthr = 2;
    plan_quantile<T>(a, n, pg.steps[0].p, pg.steps[0].p_f32, 2, 0);
    plan_quantile<T>(a, n, pg.steps[1].p, pg.steps[1].p_f32, 4, 1);
    a.s_tlo = nred; a.p0_red[nred].src = FS_IND; a.p0_red[nred].want_ess = 1; a.p0_red[nred].thr = 0; ++nred;
    a.s_thi = nred; a.p0_red[nred].src = FS_IND; a.p0_red[nred].want_ess = 1; a.p0_red[nred].thr = 1; ++nred;
    a.col[5] = o[tail][0];
  }
  if (mm >= 0) { a.s_mean = nred; a.p0_red[nred].src = FS_X; a.p0_red[nred].want_ess = 1; ++nred; a.col[2] = o[mm][0]; }
  if (ms >= 0) { a.s_std = nred; a.p0_red[nred].src = FS_SQDEV; a.p0_red[nred].want_ess = 1; ++nred; a.col[3] = o[ms][0]; }
  if (a.do_fold && !a.ncap) a.ncap = 2;   // the fold needs the median of x
  a.p0_nred = nred;
  if (mom >= 0) { a.col[0] = o[mom][0]; a.col[1] = o[mom][1]; }
  a.x = dx; a.params = params; a.niter = g.niter;
  a.maxlag = ess_pg ? ess_pg->maxlag : 1; a.relative = 0; a.ess_nan = ess_pg ? ess_pg->ess_nan : 1;
  a.rel_ess_max = rel_ess_max_of<T>((long long)g.niter * g.nch);
  const int dtype = sizeof(T) == 8 ? MCD_F64 : MCD_F32;
  if (a.p0_rank || a.do_fold) {
    int rc = ensure_ztab(ctx, dtype, g.n);
    if (rc) return rc;
    a.ztab = (const T*)ctx->ztab + 2 * (size_t)g.n;   // split layout
  }
  int rc = ensure_cap(ctx, (void**)&ctx->d_redo, &ctx->redo_cap, (size_t)(params + 1) * sizeof(int));
  if (rc) return rc;
  a.redo_count = ctx->d_redo; a.redo_list = ctx->d_redo + 1;
  CU(cudaMemsetAsync(ctx->d_redo, 0, sizeof(int), ctx->stream));
  const size_t smem = fast_smem_bytes<T>(a.maxlag) + FASTGEN_EXTRA_SMEM + (size_t)ctx->fast_pad_smem;
  auto kern = g.niter > 32 * (FAST_EPT - 1) ? fastgen_kernel<T, false, true> : fastgen_kernel<T, false, false>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)params, FAST_THREADS, smem, ctx->stream>>>(a);
  ctx->launches++;
  CU(cudaGetLastError());
  // slabs the kernel declined (NaN, infinite range, heavy ties): every program again on the general kernel
  for (int j = 0; j < njobs; ++j) {
    if (jobs[j].role == ROLE_MOMENTS) continue;
    bool h2 = false;
    rc = run_slab<T>(ctx, dx, params, g, jobs[j].pg, o[j][0], o[j][1], nullptr, &h2, ctx->d_redo + 1, ctx->d_redo);
    if (rc) return rc;
    if (!h2) return fail(ctx, MCD_EUNSUPPORTED, "internal: redo kernel unavailable");
  }
  ctx->last_path = 3;
  *handled = true;
  return MCD_OK;
}

// The big-slab estimator kernel (mcd_big.cuh): slabs too large for two CTAs of the general kernel per SM but small
// enough to live in one SM's shared memory (e.g. 4000 x 8 Float32), estimator ESS with the mean / std / median proxies,
// direct or BDA autocovariance.
template <typename T>
static int run_big(mcd_ctx* ctx, const T* dx, long long params, const SplitGeom& g, const Program& pg,
                   T* d_ess, T* d_rhat, bool* handled) {
  *handled = false;
  if (!ctx->use_big || pg.nsteps != 1 || pg.combine != CB_PLAIN || pg.want_arr || pg.chain_inds) return MCD_OK;
  const Step& s0 = pg.steps[0];
  if (s0.reduce != RD_ESS_RHAT) return MCD_OK;
  int proxy;
  if (s0.transform == TR_NONE) proxy = 0;
  else if (s0.transform == TR_STDPROXY) proxy = 1;
  else if (s0.transform == TR_IND_MEDIAN) proxy = 2;
  else return MCD_OK;
  const bool ess_live = pg.want_ess && !pg.ess_nan;
  if (ess_live && pg.method != MCD_AUTOCOV_DIRECT && pg.method != MCD_AUTOCOV_BDA) return MCD_OK;
  if ((size_t)g.n * sizeof(T) <= (64u << 10) || ((uintptr_t)dx & 15u) != 0) return MCD_OK;
  BigArgs<T> a;
  memset(&a, 0, sizeof a);
  const size_t smem = big_smem_bytes<T>(g, pg.maxlag, &a.off_aux, &a.off_part, &a.off_small);
  if (!smem || smem > (size_t)ctx->smem_optin) return MCD_OK;
  a.x = dx; a.params = params; a.g = g; a.proxy = proxy; a.method = pg.method; a.maxlag = pg.maxlag;
  a.relative = pg.relative; a.ess_nan = pg.ess_nan; a.want_ess = pg.want_ess ? 1 : 0;
  a.rel_ess_max = rel_ess_max_of<T>((long long)g.niter * g.nch);
  a.ess_out = pg.want_ess ? d_ess : nullptr; a.rhat_out = pg.want_rhat ? d_rhat : nullptr;
  CU(big_launch<T>(a, (unsigned)std::min<long long>(params, ctx->sm_count), ctx->stream));
  ctx->launches++;
  ctx->last_path = 4;
  *handled = true;
  return MCD_OK;
}

// Run a program on device-resident data (params slabs).  Outputs are device pointers.
template <typename T>
static int run_device(mcd_ctx* ctx, const T* dx, long long params, const SplitGeom& g, const Program& pg,
                      T* d_ess, T* d_rhat, void* d_arr) {
  if (params == 0) return MCD_OK;
  bool handled = false;
  if (ctx->force_path == 0 || ctx->force_path == 3) {
    int rc = run_fast<T>(ctx, dx, params, g, pg, d_ess, d_rhat, &handled);
    if (rc) return rc;
    if (handled) return MCD_OK;
    if (ctx->force_path == 3) return fail(ctx, MCD_EUNSUPPORTED, "this call is outside the fast kernel's shapes/programs");
  }
  if (ctx->force_path == 0) {
    int rc = run_big<T>(ctx, dx, params, g, pg, d_ess, d_rhat, &handled);
    if (rc) return rc;
    if (handled) return MCD_OK;
  }
  if (ctx->force_path != 2) {
    int rc = run_slab<T>(ctx, dx, params, g, pg, d_ess, d_rhat, d_arr, &handled);
    if (rc) return rc;
    if (handled) return MCD_OK;
    if (ctx->force_path == 1)
      return fail(ctx, MCD_EUNSUPPORTED, "slab of %d values does not fit the shared-memory kernel", g.n);
  }
  LargeEnv env;
  env.stream = ctx->stream; env.sm_count = ctx->sm_count; env.flags = ctx->d_flags;
  env.workspace_bytes = ctx->workspace_bytes; env.launches = &ctx->launches;
  env.work = &ctx->work; env.work_cap = &ctx->work_cap; env.smem_optin = ctx->smem_optin;
  env.d_chain_inds = ctx->d_chain_inds;
  env.rel_ess_max = (double)rel_ess_max_of<T>((long long)g.niter * g.nch);
  env.use_crank = ctx->use_crank; env.crank_factor = ctx->crank_factor; env.crank_chunk = ctx->crank_chunk;
  env.fft_tc = ctx->fft_tc; env.fft_full = ctx->fft_full; env.fft_pair = ctx->fft_pair;
  env.crank_chunks = &ctx->crank_chunks; env.crank_fallbacks = &ctx->crank_fallbacks;
  // the counting rank turns ranks into z by a table lookup while the table stays cache-sized (<= 32 MB)
  if (ctx->use_crank && g.n >= 1024 && (size_t)g.n * 4 * sizeof(T) <= ((size_t)ctx->ztab_max_mb << 20)) {
    const int zrc = ensure_ztab(ctx, sizeof(T) == 8 ? MCD_F64 : MCD_F32, g.n);
    if (zrc) return zrc;
    env.ztab = ctx->ztab;
  }
  std::string msg;
  int rc = run_large<T>(env, dx, params, g, pg.nsteps, pg.steps, pg.combine, pg.method, pg.maxlag, pg.relative,
                        pg.ess_nan, pg.mcse_p, (int)pg.cps, (int)pg.nsuper, d_ess, d_rhat, d_arr, msg);
  if (rc) return fail(ctx, rc, "%s", msg.c_str());
  ctx->last_path = 2;
  return MCD_OK;
}

static int read_flags(mcd_ctx* ctx, unsigned* out) {
  CU(cudaMemcpyAsync(out, ctx->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MCD_OK;
}

// Common driver: runs the jobs of one call over the same input.  Handles host staging (chunked,
// double-buffered, overlapped: every chunk crosses PCIe once however many jobs read it) or
// device-resident input, then surfaces kernel-raised flags.
template <typename T>
static int run_job(mcd_ctx* ctx, const Job& jb, const T* dx, long long params, const SplitGeom& g, T* o0, T* o1, void* arr) {
  if (jb.moments) {
    if (params == 0) return MCD_OK;
    const int grid = (int)std::min<long long>(params, (long long)ctx->sm_count * 16);
    moments_kernel<T><<<grid, 256, 0, ctx->stream>>>(dx, params, (long long)g.draws * g.chains, o0, o1);
    CU(cudaGetLastError());
    ++ctx->launches;
    return MCD_OK;
  }
  return run_device<T>(ctx, dx, params, g, jb.pg, o0, o1, arr);
}

// All jobs of a call on one device-resident chunk (fused into one kernel when they are the
// columns of mcd_summary on a fast-kernel shape).
template <typename T>
static int run_jobs(mcd_ctx* ctx, const Job* jobs, int njobs, const T* dx, long long params, const SplitGeom& g,
                    T* const (*o)[2], void* const* arrs) {
  if (njobs > 1 && ctx->force_path == 0) {
    bool handled = false;
    int rc = run_fast_summary<T>(ctx, dx, params, g, jobs, njobs, o, &handled);
    if (rc) return rc;
    if (handled) return MCD_OK;
  }
  for (int j = 0; j < njobs; ++j) {
    int rc = run_job<T>(ctx, jobs[j], dx, params, g, o[j][0], o[j][1], arrs[j]);
    if (rc) return rc;
  }
  return MCD_OK;
}

template <typename T>
static int execute_jobs_t(mcd_ctx* ctx, const void* x, int mem, long long draws, long long chains, long long params,
                          int split, Job* jobs, int njobs) {
  if (draws <= 0 || chains <= 0 || params < 0) return fail(ctx, MCD_EINVAL, "draws and chains must be positive");
  if (split < 1) return fail(ctx, MCD_EINVAL, "split_chains must be >= 1");
  if (draws * chains > (1ll << 30)) return fail(ctx, MCD_EUNSUPPORTED, "slab too large (draws*chains > 2^30)");
  if (params > 0 && !x) return fail(ctx, MCD_EINVAL, "x is NULL");
  SplitGeom g((int)draws, (int)chains, split);
  const long long n = g.n;
  ON_DEVICE(ctx);
  CU(cudaMemsetAsync(ctx->d_flags, 0, sizeof(unsigned), ctx->stream));
  bool can_raise = false;   // flags are only consulted for programs that can raise an error
  for (int j = 0; j < njobs; ++j) {
    const Program& pg = jobs[j].pg;
    for (int s = 0; s < pg.nsteps; ++s) can_raise |= (pg.steps[s].transform == TR_IND_QUANTILE);
    if (pg.chain_inds) {
      size_t bytes = (size_t)(pg.cps * pg.nsuper) * sizeof(int);
      int rc = ensure_cap(ctx, (void**)&ctx->d_chain_inds, &ctx->chain_inds_cap, bytes);
      if (rc) return rc;
      CU(cudaMemcpyAsync(ctx->d_chain_inds, pg.chain_inds, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  if (mem == MCD_DEVICE) {
    {
      T* o[8][2]; void* arrs[8];
      for (int j = 0; j < njobs; ++j) { o[j][0] = (T*)jobs[j].out0; o[j][1] = (T*)jobs[j].out1; arrs[j] = jobs[j].arr; }
      int rc = run_jobs<T>(ctx, jobs, njobs, (const T*)x, params, g, o, arrs);
      if (rc) return rc;
    }
    if (can_raise) {
      unsigned fl = 0;
      int rc = read_flags(ctx, &fl);
      if (rc) return rc;
      if (fl & FLAG_NAN_QUANTILE) return fail(ctx, MCD_ENAN, "quantiles are undefined in presence of NaNs");
    }
    return MCD_OK;
  }
  if (mem != MCD_HOST) return fail(ctx, MCD_EINVAL, "bad mem kind %d", mem);

  // ---- host-resident input: chunk the parameter axis, overlap H2D with compute ------------
  const size_t slab_bytes = (size_t)n * sizeof(T);
  long long chunk = std::max<long long>(1, ctx->h2d_chunk_bytes / (long long)slab_bytes);
  chunk = std::min(chunk, std::max<long long>(params, 1));
  int rc = MCD_OK;
  for (int i = 0; i < 2; ++i) {
    rc = ensure_cap(ctx, &ctx->stage[i], &ctx->stage_cap[i], (size_t)chunk * slab_bytes);
    if (rc) return rc;
  }
  const size_t col = (size_t)std::max<long long>(params, 1);   // device outputs: [job][2][params]
  rc = ensure_cap(ctx, &ctx->d_out, &ctx->out_cap, (size_t)njobs * 2 * col * sizeof(T));
  if (rc) return rc;
  size_t arr_elem = 0;
  for (int j = 0; j < njobs; ++j) if (jobs[j].pg.want_arr) arr_elem = std::max(arr_elem, (size_t)jobs[j].pg.arr_elem_bytes);
  if (arr_elem) {
    rc = ensure_cap(ctx, &ctx->d_arr, &ctx->arr_cap, (size_t)chunk * n * arr_elem);
    if (rc) return rc;
  }
  T* dout = (T*)ctx->d_out;
  long long done = 0;
  int it = 0;
  while (done < params) {
    const long long cnt = std::min(chunk, params - done);
    const int b = it & 1;
    if (it >= 2) CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
    CU(cudaMemcpyAsync(ctx->stage[b], (const char*)x + (size_t)done * slab_bytes, (size_t)cnt * slab_bytes,
                       cudaMemcpyHostToDevice, ctx->copy_stream));
    ctx->h2d_bytes += cnt * (long long)slab_bytes;
    CU(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
    T* o[8][2]; void* arrs[8];
    for (int j = 0; j < njobs; ++j) {
      o[j][0] = jobs[j].out0 ? dout + (size_t)(2 * j) * col + done : nullptr;
      o[j][1] = jobs[j].out1 ? dout + (size_t)(2 * j + 1) * col + done : nullptr;
      arrs[j] = jobs[j].pg.want_arr ? ctx->d_arr : nullptr;
    }
    rc = run_jobs<T>(ctx, jobs, njobs, (const T*)ctx->stage[b], cnt, g, o, arrs);
    if (rc) return rc;
    for (int j = 0; j < njobs; ++j) {
      const Job& jb = jobs[j];
      if (jb.pg.want_arr) {
        size_t bytes = (size_t)cnt * n * jb.pg.arr_elem_bytes;
        CU(cudaMemcpyAsync((char*)jb.arr + (size_t)done * n * jb.pg.arr_elem_bytes, ctx->d_arr, bytes,
                           cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h_bytes += (long long)bytes;
      }
    }
    CU(cudaEventRecord(ctx->ev_done[b], ctx->stream));
    done += cnt;
    ++it;
  }
  for (int j = 0; j < njobs; ++j) {
    void* outs[2] = {jobs[j].out0, jobs[j].out1};
    for (int h = 0; h < 2; ++h) {
      if (!outs[h] || params == 0) continue;
      CU(cudaMemcpyAsync(outs[h], dout + (size_t)(2 * j + h) * col, (size_t)params * sizeof(T), cudaMemcpyDeviceToHost,
                         ctx->stream));
      ctx->d2h_bytes += params * (long long)sizeof(T);
    }
  }
  unsigned fl = 0;
  rc = read_flags(ctx, &fl);  // also synchronises the stream
  if (rc) return rc;
  if (fl & FLAG_NAN_QUANTILE) return fail(ctx, MCD_ENAN, "quantiles are undefined in presence of NaNs");
  return MCD_OK;
}

// A multi-GPU group (mcd_create_multi): parameters are independent (the reference's loops at src/ess_rhat.jl:380,517,
// src/rhat_nested.jl:145), so device i of D gets the contiguous parameter range [i P / D, (i + 1) P / D) -- one
// contiguous byte range of the column-major host array -- and writes its results straight into the same range of the
// caller's output buffers.  One host thread per device drives that device's own staging pipeline; no collective is
// needed because the results land in host memory.
static int execute_jobs(mcd_ctx* ctx, const void* x, int mem, int dtype, long long draws, long long chains,
                        long long params, int split, Job* jobs, int njobs);

static int execute_jobs_multi(mcd_ctx* ctx, const void* x, int mem, int dtype, long long draws, long long chains,
                              long long params, int split, Job* jobs, int njobs) {
  if (mem != MCD_HOST)
    return fail(ctx, MCD_EUNSUPPORTED, "a multi-GPU context takes host arrays (MCD_HOST); device-resident input belongs to one device's context");
  if (dtype != MCD_F64 && dtype != MCD_F32) return fail(ctx, MCD_EINVAL, "bad dtype %d", dtype);
  const size_t ts = dtype == MCD_F64 ? 8 : 4;
  const int D = (int)ctx->children.size();
  const size_t slab_bytes = (size_t)draws * (size_t)chains * ts;
  std::vector<int> rcs(D, MCD_OK);
  std::vector<std::thread> threads;
  for (int d = 0; d < D; ++d) {
    const long long lo = (long long)d * params / D, hi = (long long)(d + 1) * params / D;
    if (hi == lo && params > 0) continue;
    threads.emplace_back([=, &rcs]() {
      mcd_ctx* c = ctx->children[d];
      std::lock_guard<std::mutex> lk(c->mu);
      c->err.clear();
      std::vector<Job> sub(jobs, jobs + njobs);
      for (Job& jb : sub) {
        if (jb.out0) jb.out0 = (char*)jb.out0 + (size_t)lo * ts;
        if (jb.out1) jb.out1 = (char*)jb.out1 + (size_t)lo * ts;
        if (jb.arr) jb.arr = (char*)jb.arr + (size_t)lo * (size_t)draws * (size_t)chains * (size_t)jb.pg.arr_elem_bytes;
      }
      rcs[d] = execute_jobs(c, (const char*)x + (size_t)lo * slab_bytes, mem, dtype, draws, chains, hi - lo, split,
                            sub.data(), njobs);
    });
  }
  for (auto& t : threads) t.join();
  for (int d = 0; d < D; ++d)
    if (rcs[d] != MCD_OK) return fail(ctx, rcs[d], "device %d: %s", ctx->children[d]->device, ctx->children[d]->err.c_str());
  return MCD_OK;
}

// Per-parameter skip mask (the reference skips parameters that contain `missing`, src/ess_rhat.jl:382-385,519-523):
// the call runs on the maximal runs of kept parameters -- contiguous byte ranges of the caller's array, so a skipped
// parameter is never copied, staged or read -- and the outputs of skipped parameters are NaN (the shim turns them
// into `missing`).  No compacted copy of the samples is made anywhere.
template <typename T>
static int fill_skipped(mcd_ctx* ctx, int mem, void* out, long long lo, long long cnt) {
  if (!out || cnt <= 0) return MCD_OK;
  T* o = (T*)out + lo;
  const T nanv = std::numeric_limits<T>::quiet_NaN();
  if (mem == MCD_HOST) { for (long long i = 0; i < cnt; ++i) o[i] = nanv; return MCD_OK; }
  mcd_ctx* c = ctx->children.empty() ? ctx : ctx->children[0];
  DeviceGuard dev_guard_c_(c->device); CU(dev_guard_c_.err);
  fill_kernel<T><<<(unsigned)std::min<long long>((cnt + 255) / 256, 4096), 256, 0, c->stream>>>(o, cnt, nanv);
  CU(cudaGetLastError());
  return MCD_OK;
}

static int execute_jobs_masked(mcd_ctx* ctx, const void* x, int mem, int dtype, long long draws, long long chains,
                               long long params, int split, Job* jobs, int njobs) {
  std::vector<unsigned char> skip;
  skip.swap(ctx->skip);   // consumed by this call
  if ((long long)skip.size() != params)
    return fail(ctx, MCD_EINVAL, "the parameter mask has %lld entries but the call has %lld parameters", (long long)skip.size(), params);
  if (dtype != MCD_F64 && dtype != MCD_F32) return fail(ctx, MCD_EINVAL, "bad dtype %d", dtype);
  const size_t ts = dtype == MCD_F64 ? 8 : 4;
  const size_t slab_bytes = (size_t)draws * (size_t)chains * ts;
  long long p = 0;
  while (p < params) {
    long long q = p;
    const bool skipped = skip[p] != 0;
    while (q < params && (skip[q] != 0) == skipped) ++q;
    std::vector<Job> sub(jobs, jobs + njobs);
    for (Job& jb : sub) {
      if (skipped) {
        int rc = dtype == MCD_F64 ? fill_skipped<double>(ctx, mem, jb.out0, p, q - p) : fill_skipped<float>(ctx, mem, jb.out0, p, q - p);
        if (!rc) rc = dtype == MCD_F64 ? fill_skipped<double>(ctx, mem, jb.out1, p, q - p) : fill_skipped<float>(ctx, mem, jb.out1, p, q - p);
        if (rc) return rc;
        if (jb.arr) {   // per-element outputs of the transform calls: arr_elem_bytes is 4 or 8
          const long long e0 = p * draws * chains, ec = (q - p) * draws * chains;
          rc = jb.pg.arr_elem_bytes == 8 ? fill_skipped<double>(ctx, mem, jb.arr, e0, ec) : fill_skipped<float>(ctx, mem, jb.arr, e0, ec);
          if (rc) return rc;
        }
        continue;
      }
      if (jb.out0) jb.out0 = (char*)jb.out0 + (size_t)p * ts;
      if (jb.out1) jb.out1 = (char*)jb.out1 + (size_t)p * ts;
      if (jb.arr) jb.arr = (char*)jb.arr + (size_t)p * (size_t)draws * (size_t)chains * (size_t)jb.pg.arr_elem_bytes;
    }
    if (!skipped) {
      const int rc = execute_jobs(ctx, (const char*)x + (size_t)p * slab_bytes, mem, dtype, draws, chains, q - p, split, sub.data(), njobs);
      if (rc) return rc;
    }
    p = q;
  }
  return MCD_OK;
}

static int execute_jobs(mcd_ctx* ctx, const void* x, int mem, int dtype, long long draws, long long chains,
                        long long params, int split, Job* jobs, int njobs) {
  if (!ctx->skip.empty()) return execute_jobs_masked(ctx, x, mem, dtype, draws, chains, params, split, jobs, njobs);
  if (!ctx->children.empty()) return execute_jobs_multi(ctx, x, mem, dtype, draws, chains, params, split, jobs, njobs);
  if (dtype == MCD_F64) return execute_jobs_t<double>(ctx, x, mem, draws, chains, params, split, jobs, njobs);
  if (dtype == MCD_F32) return execute_jobs_t<float>(ctx, x, mem, draws, chains, params, split, jobs, njobs);
  return fail(ctx, MCD_EINVAL, "bad dtype %d", dtype);
}

static int execute(mcd_ctx* ctx, const void* x, int mem, int dtype, long long draws, long long chains,
                   long long params, int split, Program& pg, void* ess_out, void* rhat_out, void* arr_out) {
  Job jb;
  jb.pg = pg; jb.out0 = pg.want_ess ? ess_out : nullptr; jb.out1 = pg.want_rhat ? rhat_out : nullptr; jb.arr = arr_out;
  return execute_jobs(ctx, x, mem, dtype, draws, chains, params, split, &jb, 1);
}

// maxlag / niter rules of _ess_rhat(Val(:basic)) (src/ess_rhat.jl:469-484)
static int setup_ess(mcd_ctx* ctx, Program& pg, long long draws, int split, int method, int maxlag, int relative) {
  if (method < 0 || method > 2) return fail(ctx, MCD_EINVAL, "unknown autocov_method %d", method);
  if (split < 1) return fail(ctx, MCD_EINVAL, "split_chains must be >= 1");
  const long long niter = draws / split;
  pg.method = method; pg.relative = relative ? 1 : 0;
  if (!(niter > 4)) { pg.ess_nan = 1; pg.maxlag = 1; return MCD_OK; }
  if (!(maxlag > 0)) return fail(ctx, MCD_EINVAL, "maxlag must be >0.");
  pg.maxlag = (int)std::min<long long>(maxlag, niter - 4);
  return MCD_OK;
}

// Secondary entry points: plain moment kernels.  Host input is staged in parameter
// chunks through the first staging buffer (synchronously: these calls are HBM/PCIe-trivial).
template <typename T>
static int chain_moments_t(mcd_ctx* ctx, const void* x, int mem, long long draws, long long chains, long long params,
                           int split, void* mean_out, void* var_out) {
  SplitGeom g((int)draws, (int)chains, split);
  if (g.niter < 1) return fail(ctx, MCD_EINVAL, "fewer draws than split chains");
  ON_DEVICE(ctx);
  auto launch = [&](const T* dx, long long cnt, T* dm, T* dv) -> int {
    const long long warps = cnt * g.nch;
    const unsigned grid = (unsigned)std::min<long long>((warps + 7) / 8, (long long)ctx->sm_count * 32);
    chain_moments_kernel<T><<<grid ? grid : 1, 256, 0, ctx->stream>>>(dx, cnt, g, dm, dv);
    CU(cudaGetLastError());
    ++ctx->launches;
    return MCD_OK;
  };
  if (params == 0) return MCD_OK;
  if (mem == MCD_DEVICE) return launch((const T*)x, params, (T*)mean_out, (T*)var_out);
  if (mem != MCD_HOST) return fail(ctx, MCD_EINVAL, "bad mem kind %d", mem);
  const size_t slab_bytes = (size_t)g.n * sizeof(T), row = (size_t)g.nch * sizeof(T);
  long long chunk = std::max<long long>(1, ctx->h2d_chunk_bytes / (long long)slab_bytes);
  chunk = std::min(chunk, params);
  int rc = ensure_cap(ctx, &ctx->stage[0], &ctx->stage_cap[0], (size_t)chunk * slab_bytes);
  if (rc) return rc;
  rc = ensure_cap(ctx, &ctx->d_out, &ctx->out_cap, 2 * (size_t)chunk * row);
  if (rc) return rc;
  T* dm = (T*)ctx->d_out; T* dv = dm + chunk * g.nch;
  for (long long done = 0; done < params; done += chunk) {
    const long long cnt = std::min(chunk, params - done);
    CU(cudaMemcpyAsync(ctx->stage[0], (const char*)x + (size_t)done * slab_bytes, (size_t)cnt * slab_bytes,
                       cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += cnt * (long long)slab_bytes;
    rc = launch((const T*)ctx->stage[0], cnt, dm, dv);
    if (rc) return rc;
    if (mean_out) CU(cudaMemcpyAsync((char*)mean_out + (size_t)done * row, dm, (size_t)cnt * row, cudaMemcpyDeviceToHost, ctx->stream));
    if (var_out) CU(cudaMemcpyAsync((char*)var_out + (size_t)done * row, dv, (size_t)cnt * row, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += 2 * cnt * (long long)row;
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return MCD_OK;
}

template <typename T>
static int bfmi_t(mcd_ctx* ctx, const void* e, int mem, long long draws, long long chains, void* out) {
  ON_DEVICE(ctx);
  const T* de = (const T*)e;
  T* dout = (T*)out;
  const size_t bytes = (size_t)draws * chains * sizeof(T);
  if (mem == MCD_HOST) {
    int rc = ensure_cap(ctx, &ctx->stage[0], &ctx->stage_cap[0], bytes);
    if (rc) return rc;
    rc = ensure_cap(ctx, &ctx->d_out, &ctx->out_cap, (size_t)chains * sizeof(T));
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->stage[0], e, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (long long)bytes;
    de = (const T*)ctx->stage[0]; dout = (T*)ctx->d_out;
  } else if (mem != MCD_DEVICE) return fail(ctx, MCD_EINVAL, "bad mem kind %d", mem);
  const unsigned grid = (unsigned)std::min<long long>((chains + 7) / 8, (long long)ctx->sm_count * 32);
  bfmi_kernel<T><<<grid ? grid : 1, 256, 0, ctx->stream>>>(de, draws, chains, dout);
  CU(cudaGetLastError());
  ++ctx->launches;
  if (mem == MCD_HOST) {
    CU(cudaMemcpyAsync(out, dout, (size_t)chains * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += chains * (long long)sizeof(T);
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return MCD_OK;
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int mcd_abi_version(void) { return MCD_ABI_VERSION; }
const char* mcd_create_error(void) { return g_create_err.c_str(); }

int mcd_create(mcd_ctx** out, int device) {
  if (!out) return MCD_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, MCD_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, MCD_EINVAL, "device %d out of range [0,%d)", device, ndev);
  mcd_ctx* ctx = new (std::nothrow) mcd_ctx();
  if (!ctx) return MCD_ENOMEM;
  ctx->device = device;
  auto bail = [&](const char* what, cudaError_t er) {
    fail(nullptr, MCD_ECUDA, "%s: %s", what, cudaGetErrorString(er));
    delete ctx;
    return MCD_ECUDA;
  };
  DeviceGuard dev_guard_(device);
  if ((e = dev_guard_.err) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  if (prop.major < 10) {
    fail(nullptr, MCD_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
         prop.major, prop.minor);
    delete ctx;
    return MCD_ECUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  ctx->stream = ctx->own_stream;
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
  }
  if ((e = cudaMalloc(&ctx->d_flags, sizeof(unsigned))) != cudaSuccess) return bail("cudaMalloc", e);
  *out = ctx;
  return MCD_OK;
}

int mcd_create_multi(mcd_ctx** out, const int* devices, int ndev) {
  if (!out || !devices || ndev < 1) return MCD_EINVAL;
  *out = nullptr;
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return fail(nullptr, MCD_EINVAL, "device %d listed twice", devices[i]);
  mcd_ctx* grp = new (std::nothrow) mcd_ctx();
  if (!grp) return MCD_ENOMEM;
  grp->device = devices[0];
  for (int i = 0; i < ndev; ++i) {
    mcd_ctx* c = nullptr;
    const int rc = mcd_create(&c, devices[i]);
    if (rc != MCD_OK) {
      for (mcd_ctx* k : grp->children) mcd_destroy(k);
      delete grp;
      return rc;   // g_create_err holds the message
    }
    grp->children.push_back(c);
  }
  grp->sm_count = grp->children[0]->sm_count;
  grp->smem_optin = grp->children[0]->smem_optin;
  *out = grp;
  return MCD_OK;
}

void mcd_destroy(mcd_ctx* ctx) {
  if (!ctx) return;
  if (!ctx->children.empty()) {
    for (mcd_ctx* c : ctx->children) mcd_destroy(c);
    delete ctx;
    return;
  }
  DeviceGuard dev_guard_(ctx->device);
  cudaDeviceSynchronize();
  void* ptrs[] = {ctx->ztab, ctx->tw, ctx->d_flags, ctx->d_chain_inds, ctx->d_redo, ctx->stage[0], ctx->stage[1],
                  ctx->d_out, ctx->d_arr, ctx->work};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
  }
  if (ctx->ev_switch) cudaEventDestroy(ctx->ev_switch);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

const char* mcd_last_error(const mcd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int mcd_set_stream(mcd_ctx* ctx, void* cuda_stream, int use_own) {
  if (!ctx) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->children.empty()) {
    if (use_own) return MCD_OK;   // every device of a group always runs on its own streams
    return fail(ctx, MCD_EUNSUPPORTED, "a multi-GPU context runs on its devices' own streams");
  }
  cudaStream_t next = use_own ? ctx->own_stream : (cudaStream_t)cuda_stream;
  if (next != ctx->stream) {
    // Every call shares the context's scratch (status flags, redo list, cached z / twiddle tables, workspace):
    // work queued on the new stream must not start before the work already queued on the old one is done.
    ON_DEVICE(ctx);
    if (!ctx->ev_switch) CU(cudaEventCreateWithFlags(&ctx->ev_switch, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_switch, ctx->stream));
    CU(cudaStreamWaitEvent(next, ctx->ev_switch, 0));
    ctx->stream = next;
  }
  return MCD_OK;
}

int mcd_synchronize(mcd_ctx* ctx) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) {
    for (mcd_ctx* c : ctx->children) { const int rc = mcd_synchronize(c); if (rc) return rc; }
    return MCD_OK;
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ON_DEVICE(ctx);
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->copy_stream));
  return MCD_OK;
}

int mcd_set_param_mask(mcd_ctx* ctx, const unsigned char* skip, int64_t params) {
  if (!ctx || params < 0 || (params > 0 && !skip)) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->skip.assign(skip, skip + params);
  bool any = false;
  for (int64_t i = 0; i < params && !any; ++i) any = skip[i] != 0;
  if (!any) ctx->skip.clear();   // nothing to skip: the plain path
  return MCD_OK;
}

int mcd_set_option(mcd_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return MCD_EINVAL;
  if (!ctx->children.empty()) {
    for (mcd_ctx* c : ctx->children) {
      const int rc = mcd_set_option(c, key, value);
      if (rc) return fail(ctx, rc, "%s", c->err.c_str());
    }
    return MCD_OK;
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::string k(key);
  if (k == "force_path") { if (value < 0 || value > 3) return fail(ctx, MCD_EINVAL, "force_path in 0..3"); ctx->force_path = (int)value; }
  else if (k == "h2d_chunk_bytes") { if (value < 1) return fail(ctx, MCD_EINVAL, "h2d_chunk_bytes >= 1"); ctx->h2d_chunk_bytes = value; }
  else if (k == "workspace_bytes") { if (value < (1 << 20)) return fail(ctx, MCD_EINVAL, "workspace_bytes >= 1 MiB"); ctx->workspace_bytes = value; }
  else if (k == "fast_pad_smem") { ctx->fast_pad_smem = (int)value; }
  else if (k == "fast_grid_mult") { ctx->fast_grid_mult = (int)value; }
  else if (k == "use_rk2") { ctx->use_rk2 = (int)value; }
  else if (k == "use_big") { ctx->use_big = (int)value; }
  else if (k == "use_crank") { ctx->use_crank = (int)value; }
  else if (k == "crank_factor") { if (value < 1 || value > 64) return fail(ctx, MCD_EINVAL, "crank_factor in 1..64"); ctx->crank_factor = (int)value; }
  else if (k == "ztab_max_mb") { if (value < 0 || value > 4096) return fail(ctx, MCD_EINVAL, "ztab_max_mb in 0..4096"); ctx->ztab_max_mb = (int)value; }
  else if (k == "fft_pair") { ctx->fft_pair = value ? 1 : 0; }
  else if (k == "fft_full") { ctx->fft_full = value ? 1 : 0; }
  else if (k == "fft_tc") { if (value < 0 || value > 4) return fail(ctx, MCD_EINVAL, "fft_tc in 0..4"); ctx->fft_tc = (int)value; }
  else if (k == "crank_chunk") { if (value < 0) return fail(ctx, MCD_EINVAL, "crank_chunk >= 0"); ctx->crank_chunk = value; }
  else if (k == "slab_wide") ctx->slab_wide = value ? 1 : 0;
  else if (k == "slab_three") ctx->slab_three = value ? 1 : 0;
  else if (k == "sort_bucket_limit") { if (value < 0) return fail(ctx, MCD_EINVAL, "sort_bucket_limit >= 0"); ctx->bucket_limit = (int)value; }
  else return fail(ctx, MCD_EINVAL, "unknown option '%s'", key);
  return MCD_OK;
}

int64_t mcd_get_stat(const mcd_ctx* ctx, const char* key) {
  if (!ctx || !key) return -1;
  std::string k(key);
  if (k == "ndev") return ctx->children.empty() ? 1 : (int64_t)ctx->children.size();
  if (!ctx->children.empty()) {
    if (k == "kernel_launches" || k == "h2d_bytes" || k == "d2h_bytes" || k == "redo_count") {
      int64_t tot = 0;
      for (const mcd_ctx* c : ctx->children) { const int64_t v = mcd_get_stat(c, key); if (v < 0) return v; tot += v; }
      return tot;
    }
    return mcd_get_stat(ctx->children[0], key);
  }
  if (k == "kernel_launches") return ctx->launches;
  if (k == "crank_chunks") return ctx->crank_chunks;        // chunks of large slabs ranked by counting (cumulative)
  if (k == "crank_fallbacks") return ctx->crank_fallbacks;  // counting-rank attempts handed to the sort path (cumulative)
  if (k == "last_path") return ctx->last_path;
  if (k == "h2d_bytes") return ctx->h2d_bytes;
  if (k == "d2h_bytes") return ctx->d2h_bytes;
  if (k == "redo_count") {   // parameters the register-resident kernel handed to the general kernel in the last call (last chunk)
    if (!ctx->d_redo) return 0;
    int v = 0;
    DeviceGuard dev_guard_(ctx->device);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, ctx->d_redo, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
  }
  if (k == "sm_count") return ctx->sm_count;
  if (k == "smem_optin") return ctx->smem_optin;
  return -1;
}

// Program of ess_rhat / ess / rhat for a Symbol kind (pg.want_ess / pg.want_rhat already set).
static int build_ess_rhat(mcd_ctx* ctx, Program& pg, int dtype, int64_t draws, int kind, int autocov_method,
                          int split_chains, int maxlag, int relative, double tail_prob, int tail_prob_f64) {
  if (kind < 0 || kind > 3) return fail(ctx, MCD_EINVAL, "the `kind` %d is not supported", kind);
  if (pg.want_ess) { int rc = setup_ess(ctx, pg, draws, split_chains, autocov_method, maxlag, relative); if (rc) return rc; }
  const int rd = pg.want_ess ? RD_ESS_RHAT : RD_RHAT;
  const int p_f32 = (dtype == MCD_F32 && !tail_prob_f64) ? 1 : 0;
  double pl, pu;
  if (p_f32) { pl = (double)(float)(tail_prob / 2); pu = (double)(float)(1.0 - tail_prob / 2); }
  else { pl = tail_prob / 2; pu = 1.0 - tail_prob / 2; }
  switch (kind) {
    case MCD_KIND_BASIC: pg.add(TR_NONE, rd); break;
    case MCD_KIND_BULK: pg.add(TR_RANKNORM, rd); break;
    case MCD_KIND_TAIL:
      if (pg.want_ess) {
        if (!(pl >= 0.0 && pl <= 1.0 && pu >= 0.0 && pu <= 1.0)) return fail(ctx, MCD_EINVAL, "tail_prob out of range");
        pg.add(TR_IND_QUANTILE, RD_ESS_RHAT, pl, p_f32);
        pg.add(TR_IND_QUANTILE, RD_ESS_RHAT, pu, p_f32);
        if (pg.want_rhat) { pg.add(TR_FOLD_RANKNORM, RD_RHAT); pg.combine = CB_TAIL; }
        else pg.combine = CB_TAIL_ESS;
      } else pg.add(TR_FOLD_RANKNORM, RD_RHAT);
      break;
    case MCD_KIND_RANK:
      if (!pg.want_rhat) return fail(ctx, MCD_EINVAL, "the `kind` `rank` is not supported by `ess`");
      pg.add(TR_RANKNORM, rd);
      pg.add(TR_FOLD_RANKNORM, RD_RHAT);
      pg.combine = CB_RANK;
      break;
  }
  return MCD_OK;
}

int mcd_ess_rhat(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains, int64_t params,
                 int kind, int autocov_method, int split_chains, int maxlag, int relative, double tail_prob,
                 int tail_prob_f64, void* ess_out, void* rhat_out) {
  if (!ctx) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!ess_out && !rhat_out) return fail(ctx, MCD_EINVAL, "both outputs are NULL");
  Program pg;
  pg.want_ess = ess_out != nullptr; pg.want_rhat = rhat_out != nullptr;
  int rc = build_ess_rhat(ctx, pg, dtype, draws, kind, autocov_method, split_chains, maxlag, relative, tail_prob, tail_prob_f64);
  if (rc) return rc;
  return execute(ctx, x, mem, dtype, draws, chains, params, split_chains, pg, ess_out, rhat_out, nullptr);
}

static int estimator_step(mcd_ctx* ctx, Program& pg, int estimator, double p, int p_f32) {
  switch (estimator) {
    case MCD_EST_MEAN: pg.add(TR_NONE, RD_ESS_RHAT); break;
    case MCD_EST_MEDIAN: pg.add(TR_IND_MEDIAN, RD_ESS_RHAT); break;
    case MCD_EST_STD: pg.add(TR_STDPROXY, RD_ESS_RHAT); break;
    case MCD_EST_MAD: pg.add(TR_FOLD_IND_MEDIAN, RD_ESS_RHAT); break;
    case MCD_EST_QUANTILE:
      if (!(p >= 0.0 && p <= 1.0)) return fail(ctx, MCD_EINVAL, "input probability out of [0,1] range");
      pg.add(TR_IND_QUANTILE, RD_ESS_RHAT, p, p_f32);
      break;
    default: return fail(ctx, MCD_EINVAL, "the estimator %d is not yet supported by `ess`", estimator);
  }
  return MCD_OK;
}

int mcd_ess_estimator(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains,
                      int64_t params, int estimator, double p, int p_f64, int autocov_method, int split_chains,
                      int maxlag, int relative, void* ess_out) {
  if (!ctx) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!ess_out) return fail(ctx, MCD_EINVAL, "ess_out is NULL");
  Program pg;
  pg.want_ess = true;
  int rc = setup_ess(ctx, pg, draws, split_chains, autocov_method, maxlag, relative);
  if (rc) return rc;
  const int p_f32 = (dtype == MCD_F32 && !p_f64) ? 1 : 0;
  rc = estimator_step(ctx, pg, estimator, p_f32 ? (double)(float)p : p, p_f32);
  if (rc) return rc;
  return execute(ctx, x, mem, dtype, draws, chains, params, split_chains, pg, ess_out, nullptr, nullptr);
}

// Program of the ESS-based mcse rules (src/mcse.jl:45-118)
static int build_mcse(mcd_ctx* ctx, Program& pg, int dtype, int64_t draws, int estimator, double p, int p_f64,
                      int autocov_method, int split_chains, int maxlag) {
  pg.want_ess = true;
  int rc = setup_ess(ctx, pg, draws, split_chains, autocov_method, maxlag, 0);
  if (rc) return rc;
  const int p_f32 = (dtype == MCD_F32 && !p_f64) ? 1 : 0;
  switch (estimator) {
    case MCD_EST_MEAN: pg.combine = CB_MCSE_MEAN; break;
    case MCD_EST_STD: pg.combine = CB_MCSE_STD; break;
    case MCD_EST_MEDIAN: pg.combine = CB_MCSE_QUANTILE; pg.mcse_p = 0.5; break;
    case MCD_EST_QUANTILE: pg.combine = CB_MCSE_QUANTILE; pg.mcse_p = p; break;
    default:
      return fail(ctx, MCD_EUNSUPPORTED, "mcse for estimator %d uses the subsampling bootstrap, which stays in the host language", estimator);
  }
  return estimator_step(ctx, pg, estimator, p_f32 ? (double)(float)p : p, p_f32);
}

int mcd_mcse(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains, int64_t params,
             int estimator, double p, int p_f64, int autocov_method, int split_chains, int maxlag, void* mcse_out) {
  if (!ctx) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!mcse_out) return fail(ctx, MCD_EINVAL, "mcse_out is NULL");
  Program pg;
  int rc = build_mcse(ctx, pg, dtype, draws, estimator, p, p_f64, autocov_method, split_chains, maxlag);
  if (rc) return rc;
  return execute(ctx, x, mem, dtype, draws, chains, params, split_chains, pg, mcse_out, nullptr, nullptr);
}

// Fused per-parameter summary: what MCMCChains.summarystats / PosteriorStats.summarize ask of this
// package for every parameter (mean, std, mcse(mean), mcse(std), ess(:bulk), ess(:tail),
// rhat(:rank)), one call, the input staged (PCIe) once.  Each column equals the corresponding
// reference call with the same keywords.
int mcd_summary(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains, int64_t params,
                unsigned fields, int autocov_method, int split_chains, int maxlag, double tail_prob,
                int tail_prob_f64, void* out) {
  if (!ctx) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!out) return fail(ctx, MCD_EINVAL, "out is NULL");
  if (fields == 0 || (fields & ~(unsigned)MCD_SUM_ALL)) return fail(ctx, MCD_EINVAL, "bad summary field mask 0x%x", fields);
  if (dtype != MCD_F32 && dtype != MCD_F64) return fail(ctx, MCD_EINVAL, "bad dtype %d", dtype);
  const size_t esz = dtype == MCD_F64 ? 8 : 4;
  void* colp[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int ncol = 0;
  for (int f = 0; f < 7; ++f)
    if (fields & (1u << f)) colp[f] = (char*)out + (size_t)(ncol++) * (size_t)params * esz;
  Job jobs[5];
  int nj = 0, rc;
  if (colp[0] || colp[1]) { Job& j = jobs[nj++]; j.moments = true; j.role = ROLE_MOMENTS; j.out0 = colp[0]; j.out1 = colp[1]; }
  if (colp[2]) {
    Job& j = jobs[nj++];
    if ((rc = build_mcse(ctx, j.pg, dtype, draws, MCD_EST_MEAN, 0.0, 0, autocov_method, split_chains, maxlag))) return rc;
    j.out0 = colp[2]; j.role = ROLE_MCSE_MEAN;
  }
  if (colp[3]) {
    Job& j = jobs[nj++];
    if ((rc = build_mcse(ctx, j.pg, dtype, draws, MCD_EST_STD, 0.0, 0, autocov_method, split_chains, maxlag))) return rc;
    j.out0 = colp[3]; j.role = ROLE_MCSE_STD;
  }
  if (colp[4] || colp[6]) {
    Job& j = jobs[nj++];
    j.pg.want_ess = colp[4] != nullptr; j.pg.want_rhat = colp[6] != nullptr;
    const int kind = j.pg.want_rhat ? MCD_KIND_RANK : MCD_KIND_BULK;   // ess(:rank) = ess(:bulk) (src/ess_rhat.jl:604-624)
    if ((rc = build_ess_rhat(ctx, j.pg, dtype, draws, kind, autocov_method, split_chains, maxlag, 0, tail_prob, tail_prob_f64))) return rc;
    j.out0 = colp[4]; j.out1 = colp[6]; j.role = ROLE_BULK_RHAT;
  }
  if (colp[5]) {
    Job& j = jobs[nj++];
    j.pg.want_ess = true;
    if ((rc = build_ess_rhat(ctx, j.pg, dtype, draws, MCD_KIND_TAIL, autocov_method, split_chains, maxlag, 0, tail_prob, tail_prob_f64))) return rc;
    j.out0 = colp[5]; j.role = ROLE_TAIL;
  }
  return execute_jobs(ctx, x, mem, dtype, draws, chains, params, split_chains, jobs, nj);
}

int mcd_rhat_nested(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains, int64_t params,
                    const int32_t* chain_inds, int64_t chains_per_super, int64_t nsuper, int kind, int split_chains,
                    void* rhat_out) {
  if (!ctx) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!rhat_out || !chain_inds) return fail(ctx, MCD_EINVAL, "NULL argument");
  if (nsuper < 2) return fail(ctx, MCD_EINVAL, "at least 2 superchains are required, got %lld", (long long)nsuper);
  if (chains_per_super < 1 || chains_per_super * nsuper > chains)
    return fail(ctx, MCD_EINVAL, "chain_inds shape (%lld x %lld) does not match %lld chains", (long long)chains_per_super,
                (long long)nsuper, (long long)chains);
  for (int64_t i = 0; i < chains_per_super * nsuper; ++i)
    if (chain_inds[i] < 0 || chain_inds[i] >= chains) return fail(ctx, MCD_EINVAL, "chain index %d out of range", chain_inds[i]);
  if (kind < 0 || kind > 3) return fail(ctx, MCD_EINVAL, "the `kind` %d is not supported by `rhat_nested`", kind);
  Program pg;
  pg.want_rhat = true;
  pg.chain_inds = chain_inds; pg.cps = chains_per_super; pg.nsuper = nsuper;
  switch (kind) {
    case MCD_KIND_BASIC: pg.add(TR_NONE, RD_NESTED); break;
    case MCD_KIND_BULK: pg.add(TR_RANKNORM, RD_NESTED); break;
    case MCD_KIND_TAIL: pg.add(TR_FOLD_RANKNORM, RD_NESTED); break;
    case MCD_KIND_RANK: pg.add(TR_RANKNORM, RD_NESTED); pg.add(TR_FOLD_RANKNORM, RD_NESTED); pg.combine = CB_MAX_RHAT; break;
  }
  return execute(ctx, x, mem, dtype, draws, chains, params, split_chains, pg, nullptr, rhat_out, nullptr);
}

static int transform_call(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains,
                          int64_t params, int tr, int elem_bytes, void* out) {
  if (!ctx) return MCD_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!out) return fail(ctx, MCD_EINVAL, "output is NULL");
  Program pg;
  pg.want_arr = true; pg.arr_elem_bytes = elem_bytes;
  pg.add(tr, RD_STORE);
  return execute(ctx, x, mem, dtype, draws, chains, params, 1, pg, nullptr, nullptr, out);
}

int mcd_tiedrank(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains, int64_t params,
                 double* ranks_out) {
  return transform_call(ctx, x, mem, dtype, draws, chains, params, TR_TIEDRANK, 8, ranks_out);
}
int mcd_rank_normalize(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains,
                       int64_t params, void* out) {
  return transform_call(ctx, x, mem, dtype, draws, chains, params, TR_RANKNORM, dtype == MCD_F64 ? 8 : 4, out);
}
int mcd_fold_around_median(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains,
                           int64_t params, void* out) {
  return transform_call(ctx, x, mem, dtype, draws, chains, params, TR_FOLD, dtype == MCD_F64 ? 8 : 4, out);
}

int mcd_chain_moments(mcd_ctx* ctx, const void* x, int mem, int dtype, int64_t draws, int64_t chains, int64_t params,
                      int split_chains, void* mean_out, void* var_out) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_chain_moments(ctx->children[0], x, mem, dtype, draws, chains, params, split_chains, mean_out, var_out); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!mean_out && !var_out) return fail(ctx, MCD_EINVAL, "both outputs are NULL");
  if (draws <= 0 || chains <= 0 || params < 0) return fail(ctx, MCD_EINVAL, "draws and chains must be positive");
  if (split_chains < 1) return fail(ctx, MCD_EINVAL, "split_chains must be >= 1");
  if (draws * chains > (1ll << 30)) return fail(ctx, MCD_EUNSUPPORTED, "slab too large (draws*chains > 2^30)");
  if (params > 0 && !x) return fail(ctx, MCD_EINVAL, "x is NULL");
  if (dtype == MCD_F64) return chain_moments_t<double>(ctx, x, mem, draws, chains, params, split_chains, mean_out, var_out);
  if (dtype == MCD_F32) return chain_moments_t<float>(ctx, x, mem, draws, chains, params, split_chains, mean_out, var_out);
  return fail(ctx, MCD_EINVAL, "bad dtype %d", dtype);
}

int mcd_bfmi(mcd_ctx* ctx, const void* energy, int mem, int dtype, int64_t draws, int64_t chains, void* out) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_bfmi(ctx->children[0], energy, mem, dtype, draws, chains, out); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!energy || !out) return fail(ctx, MCD_EINVAL, "NULL argument");
  if (draws <= 0 || chains <= 0) return fail(ctx, MCD_EINVAL, "draws and chains must be positive");
  if (dtype == MCD_F64) return bfmi_t<double>(ctx, energy, mem, draws, chains, out);
  if (dtype == MCD_F32) return bfmi_t<float>(ctx, energy, mem, draws, chains, out);
  return fail(ctx, MCD_EINVAL, "bad dtype %d", dtype);
}

int mcd_generate_ar1(mcd_ctx* ctx, int dtype, int64_t draws, int64_t chains, int64_t params, int64_t param_offset,
                     double phi, double sigma, uint64_t seed, void* dev_x) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_generate_ar1(ctx->children[0], dtype, draws, chains, params, param_offset, phi, sigma, seed, dev_x); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->err.clear();
  if (!dev_x || draws <= 0 || chains <= 0 || params < 0) return fail(ctx, MCD_EINVAL, "bad argument");
  ON_DEVICE(ctx);
  long long series = chains * params;
  if (series == 0) return MCD_OK;
  long long blocks = (series + 127) / 128;
  if (dtype == MCD_F64)
    ar1_kernel<double><<<(unsigned)blocks, 128, 0, ctx->stream>>>((double*)dev_x, draws, chains, params, param_offset, phi, sigma, seed);
  else if (dtype == MCD_F32)
    ar1_kernel<float><<<(unsigned)blocks, 128, 0, ctx->stream>>>((float*)dev_x, draws, chains, params, param_offset, phi, sigma, seed);
  else return fail(ctx, MCD_EINVAL, "bad dtype");
  ctx->launches++;
  CU(cudaGetLastError());
  return MCD_OK;
}

int mcd_device_alloc(mcd_ctx* ctx, int64_t bytes, void** dev_ptr) {
  if (!ctx || !dev_ptr || bytes < 0) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_device_alloc(ctx->children[0], bytes, dev_ptr); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ON_DEVICE(ctx);
  CU(cudaMalloc(dev_ptr, (size_t)std::max<int64_t>(bytes, 1)));
  return MCD_OK;
}
int mcd_device_free(mcd_ctx* ctx, void* dev_ptr) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_device_free(ctx->children[0], dev_ptr); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ON_DEVICE(ctx);
  CU(cudaFree(dev_ptr));
  return MCD_OK;
}
int mcd_memcpy_h2d(mcd_ctx* ctx, void* dev_dst, const void* host_src, int64_t bytes) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_memcpy_h2d(ctx->children[0], dev_dst, host_src, bytes); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ON_DEVICE(ctx);
  CU(cudaMemcpyAsync(dev_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MCD_OK;
}
int mcd_memcpy_d2h(mcd_ctx* ctx, void* host_dst, const void* dev_src, int64_t bytes) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_memcpy_d2h(ctx->children[0], host_dst, dev_src, bytes); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ON_DEVICE(ctx);
  CU(cudaMemcpyAsync(host_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MCD_OK;
}
int mcd_host_alloc(mcd_ctx* ctx, int64_t bytes, void** host_ptr) {
  if (!ctx || !host_ptr || bytes < 0) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_host_alloc(ctx->children[0], bytes, host_ptr); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  ON_DEVICE(ctx);
  CU(cudaHostAlloc(host_ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
  return MCD_OK;
}
int mcd_host_free(mcd_ctx* ctx, void* host_ptr) {
  if (!ctx) return MCD_EINVAL;
  if (!ctx->children.empty()) { const int rc_ = mcd_host_free(ctx->children[0], host_ptr); if (rc_) ctx->err = ctx->children[0]->err; return rc_; }
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaFreeHost(host_ptr));
  return MCD_OK;
}

}  // extern "C"
