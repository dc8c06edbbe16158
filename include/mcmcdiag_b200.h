/*
 * mcmcdiag_b200.h — C ABI of libmcmcdiag_b200.so
 *
 * B200-native (sm_100a) implementation of the ESS / R-hat hot path of
 * MCMCDiagnosticTools.jl.  This is the drop-in boundary: plain pointers and sizes,
 * no C++ or torch types.  The Julia host shim (julia/MCMCDiagB200.jl) binds these with
 * `ccall`; the Python host (used by the tests, because no Julia binary exists in this
 * image) binds them with ctypes.  INTEGRATION.md shows both bindings.
 *
 * The seam replaced is the reference's *coarse* internal layer (array in, one value per
 * parameter out), cited per entry point below as /root/reference file:line.
 *
 * Array layout: Julia column-major.  `x` holds `params` contiguous slabs of
 * `draws*chains` elements; inside a slab every chain is `draws` contiguous elements
 * (reference: src/utils.jl:203-211 `_params_array`).
 *
 * Memory kinds: MCD_HOST pointers are ordinary (ideally pinned) host memory, staged to
 * the GPU by the library in overlapped chunks; MCD_DEVICE pointers are device memory on
 * the context's GPU.  Outputs live in the same memory kind as `x` and hold `params`
 * elements of `dtype`.
 *
 * Ownership: the caller owns every buffer it passes; the library never retains a
 * pointer after return and never frees caller memory.  Workspaces, streams and tables
 * belong to the context and are reused across calls.
 *
 * Errors: 0 on success, negative code otherwise; mcd_last_error() gives the message.
 * No exception, abort or exit crosses this boundary.  NaN is a value, not an error.
 * There is no CPU fallback: without a usable CUDA device every call returns MCD_ECUDA.
 *
 * Threading: calls on one context are serialised by an internal mutex; distinct
 * contexts may be used from distinct threads.
 */
#ifndef MCMCDIAG_B200_H
#define MCMCDIAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCD_ABI_VERSION 1

typedef struct mcd_ctx mcd_ctx;

/* status codes */
enum {
  MCD_OK = 0,
  MCD_EINVAL = -1,        /* bad argument (the host shim turns these into ArgumentError/DomainError) */
  MCD_ECUDA = -2,         /* CUDA runtime failure, or no device */
  MCD_ENOMEM = -3,        /* workspace allocation failed */
  MCD_EUNSUPPORTED = -4,  /* shape outside what this build handles */
  MCD_ENAN = -5           /* quantile of data containing NaN (Statistics.quantile throws ArgumentError) */
};

enum { MCD_F32 = 0, MCD_F64 = 1 };                 /* dtype */
enum { MCD_HOST = 0, MCD_DEVICE = 1 };             /* mem kind */

/* `kind::Symbol` of ess / rhat / ess_rhat / rhat_nested  (src/ess_rhat.jl:9-20, 245-254) */
enum { MCD_KIND_BASIC = 0, MCD_KIND_BULK = 1, MCD_KIND_TAIL = 2, MCD_KIND_RANK = 3 };

/* AbstractAutocovMethod subtypes (src/ess_rhat.jl:38,55,73); north-star aliases
 * ESSMethod / FFTESSMethod / BDAESSMethod map to the same three values. */
enum { MCD_AUTOCOV_DIRECT = 0, MCD_AUTOCOV_FFT = 1, MCD_AUTOCOV_BDA = 2 };

/* estimator `kind`s with an expectand proxy (src/ess_rhat.jl:628-659):
 * Statistics.mean, Statistics.median, Statistics.std, StatsBase.mad,
 * Base.Fix2(Statistics.quantile, p) */
enum { MCD_EST_MEAN = 0, MCD_EST_MEDIAN = 1, MCD_EST_STD = 2, MCD_EST_MAD = 3, MCD_EST_QUANTILE = 4 };

/* ---- context --------------------------------------------------------------------- */

/* Create a context bound to CUDA device `device` (one context per GPU; multi-GPU jobs
 * run one process or one context per GPU and shard the parameter axis). */
int mcd_create(mcd_ctx** out, int device);
/* Create ONE context that drives `ndev` GPUs of this process (distinct CUDA device ordinals).  Every hot-path call on
 * it takes a HOST array (MCD_HOST), shards the parameter axis -- the reference's only parallel axis, the loops at
 * src/ess_rhat.jl:380,517 and src/rhat_nested.jl:145 -- into `ndev` contiguous ranges, runs each range through its
 * device's own pinned-staging pipeline on its own host thread, and writes every device's results into the matching
 * range of the caller's single output buffer (results are independent of ndev).  This is how the Julia shim reaches all
 * GPUs of a box with no torch / NCCL in the process (mcd_create(ctx**, devices, ndev)).  Device-resident
 * input (MCD_DEVICE), mcd_set_stream and the device helpers below act on / belong to a single-device context; on a
 * group the helpers are forwarded to its first device. */
int mcd_create_multi(mcd_ctx** out, const int* devices, int ndev);
void mcd_destroy(mcd_ctx* ctx);
/* Last error message of this context ("" if none).  Valid until the next call on ctx. */
const char* mcd_last_error(const mcd_ctx* ctx);
/* Message of a failed mcd_create (no context exists yet). */
const char* mcd_create_error(void);
int mcd_abi_version(void);

/* Run the device work of subsequent calls on the caller's cudaStream_t `cuda_stream`
 * (NULL is CUDA's default stream), or, with use_own != 0, on the context's own
 * non-blocking stream (the initial state).  The caller keeps ownership of its stream. */
int mcd_set_stream(mcd_ctx* ctx, void* cuda_stream, int use_own);
/* Block until all work queued by this context has finished. */
int mcd_synchronize(mcd_ctx* ctx);

/* Per-parameter skip mask for the NEXT hot-path call on this context (consumed by it): parameters with skip[p] != 0
 * are not computed -- their bytes are never copied to the device or read -- and their outputs are NaN.  This is the
 * reference's `missing` handling (a parameter containing `missing` yields `missing`, src/ess_rhat.jl:382-385,519-523)
 * without a compacted copy of the samples: the host shim passes the array as it is (any finite filler in the masked
 * entries) and maps the NaNs of masked parameters back to `missing`.  `params` must equal the call's parameter count. */
int mcd_set_param_mask(mcd_ctx* ctx, const unsigned char* skip, int64_t params);

/* Tuning / test knobs.  Keys: "force_path" (0 auto, 1 general shared-memory slab kernel,
 * 2 global-memory large-slab pipeline, 3 register-resident fast kernel only),
 * "h2d_chunk_bytes", "workspace_bytes", "sort_bucket_limit"; developer switches that select between equivalent
 * implementations (results agree to rounding or exactly; used by the A/B scripts and the tests): "use_rk2",
 * "use_big", "use_crank", "crank_factor", "crank_chunk", "ztab_max_mb", "fft_pair", "fft_full", "fft_tc",
 * "slab_wide", "slab_three", "fast_grid_mult", "fast_pad_smem". */
int mcd_set_option(mcd_ctx* ctx, const char* key, int64_t value);
/* Counters.  Keys: "kernel_launches" (since creation), "last_path" (1 slab, 2 large, 3 fast),
 * "h2d_bytes", "d2h_bytes", "sm_count", "smem_optin", "ndev" (devices of the context), "redo_count" (parameters the register-resident
 * kernel handed to the general kernel in the last launch: NaN, infinite range, heavy ties, ...). */
int64_t mcd_get_stat(const mcd_ctx* ctx, const char* key);

/* ---- the hot path ---------------------------------------------------------------- */

/* ess_rhat / ess / rhat for kind in {:basic,:bulk,:tail,:rank}.
 * Replaces `_ess_rhat(::Val{kind}, x; relative, autocov_method, split_chains, maxlag)`
 * (src/ess_rhat.jl:456-487, 604-624), `_rhat(::Val{kind}, x; split_chains)`
 * (src/ess_rhat.jl:350-361, 410-420) and `_ess(::Val{:tail}, x; tail_prob)`
 * (src/ess_rhat.jl:301-311), including `_rank_normalize` / `_fold_around_median`
 * (src/utils.jl:148-193), `copyto_split!` (src/utils.jl:13-41), `_rhat_basic!`
 * (src/ess_rhat.jl:362-409), `_ess_rhat_basic!` (src/ess_rhat.jl:488-603) and the three
 * autocovariance methods (src/ess_rhat.jl:95-213).
 *
 * ess_out or rhat_out may be NULL (=> `rhat` alone / `ess` alone).  With ess_out set,
 * kind == MCD_KIND_RANK gives ESS_bulk (src/ess_rhat.jl:617-624).
 * niter = draws / split_chains <= 4  => ESS is NaN, R-hat still computed (:472-479; the
 * host shim emits the @warn).  maxlag <= 0 => MCD_EINVAL (DomainError, :481).
 * tail_prob is used by MCD_KIND_TAIL only; tail_prob_f64 != 0 means the caller's
 * tail_prob was a Float64 (so quantile arithmetic is Float64 even for Float32 data),
 * 0 means it promotes to the array's float type (the default Rational 1//10). */
int mcd_ess_rhat(mcd_ctx* ctx, const void* x, int mem, int dtype,
                 int64_t draws, int64_t chains, int64_t params,
                 int kind, int autocov_method, int split_chains, int maxlag, int relative,
                 double tail_prob, int tail_prob_f64,
                 void* ess_out, void* rhat_out);

/* ess(x; kind=estimator): `_ess(estimator, x; ...)` = `_expectand_proxy` +
 * `_ess(Val(:basic))` (src/ess_rhat.jl:291-297, 628-659).  p is used by
 * MCD_EST_QUANTILE; p_f64 as tail_prob_f64 above. */
int mcd_ess_estimator(mcd_ctx* ctx, const void* x, int mem, int dtype,
                      int64_t draws, int64_t chains, int64_t params,
                      int estimator, double p, int p_f64,
                      int autocov_method, int split_chains, int maxlag, int relative,
                      void* ess_out);

/* mcse(x; kind=estimator) for the ESS-based estimators mean / std / median / quantile:
 * `_mcse` + `_mcse_quantile` (src/mcse.jl:45-118).  MCD_EST_MAD and arbitrary callables use
 * the reference's subsampling-bootstrap fallback (src/mcse.jl:120-148), which stays in the
 * host language: MCD_EUNSUPPORTED here. */
int mcd_mcse(mcd_ctx* ctx, const void* x, int mem, int dtype,
             int64_t draws, int64_t chains, int64_t params,
             int estimator, double p, int p_f64,
             int autocov_method, int split_chains, int maxlag,
             void* mcse_out);

/* Fused per-parameter summary: the per-parameter columns that the callers of
 * this package (MCMCChains.summarystats, PosteriorStats.summarize) assemble from separate calls,
 * computed in ONE call with the input crossing PCIe once.  `fields` is a mask of MCD_SUM_*; `out`
 * receives one column of `params` values per selected field, in the order of the bit positions
 * (column-major, params x popcount(fields), element type = dtype).  Column definitions, each
 * identical to the reference call it replaces:
 *   MCD_SUM_MEAN       Statistics.mean(x; dims=(1,2))          (as used at src/mcse.jl:50)
 *   MCD_SUM_STD        Statistics.std(x; dims=(1,2))           (src/mcse.jl:50)
 *   MCD_SUM_MCSE_MEAN  mcse(x; kind=mean, ...)                 (src/mcse.jl:45-52)
 *   MCD_SUM_MCSE_STD   mcse(x; kind=std, ...)                  (src/mcse.jl:53-69)
 *   MCD_SUM_ESS_BULK   ess(x; kind=:bulk, ...)                 (src/ess_rhat.jl:604-624)
 *   MCD_SUM_ESS_TAIL   ess(x; kind=:tail, tail_prob, ...)      (src/ess_rhat.jl:298-311)
 *   MCD_SUM_RHAT       rhat(x; kind=:rank, split_chains)       (src/ess_rhat.jl:410-420)
 * autocov_method / split_chains / maxlag / tail_prob apply to every ESS-based column. */
enum {
  MCD_SUM_MEAN = 1, MCD_SUM_STD = 2, MCD_SUM_MCSE_MEAN = 4, MCD_SUM_MCSE_STD = 8,
  MCD_SUM_ESS_BULK = 16, MCD_SUM_ESS_TAIL = 32, MCD_SUM_RHAT = 64, MCD_SUM_ALL = 127
};
int mcd_summary(mcd_ctx* ctx, const void* x, int mem, int dtype,
                int64_t draws, int64_t chains, int64_t params,
                unsigned fields, int autocov_method, int split_chains, int maxlag,
                double tail_prob, int tail_prob_f64, void* out);

/* Per split-chain mean and corrected variance of every parameter: the per-chain quantities that
 * `_rhat_basic!` (src/ess_rhat.jl:387-399) and `_gelmandiag` (src/gelmandiag.jl:9-17: diag of the
 * per-chain covariance, chain means) start from.  Outputs (chains*split_chains, params) column-major,
 * element type = dtype; either may be NULL.   */
int mcd_chain_moments(mcd_ctx* ctx, const void* x, int mem, int dtype,
                      int64_t draws, int64_t chains, int64_t params, int split_chains,
                      void* mean_out, void* var_out);

/* bfmi(energy::AbstractMatrix; dims=1) (src/bfmi.jl:36-43): mean(abs2, diff(energy)) / var(energy) per
 * chain; energy is (draws, chains) column-major, out has `chains` elements. */
int mcd_bfmi(mcd_ctx* ctx, const void* energy, int mem, int dtype, int64_t draws, int64_t chains, void* out);

/* rhat_nested: `_rhat_nested(::Val{kind}, x, chain_inds; split_chains)` +
 * `_rhat_nested_basic!` (src/rhat_nested.jl:83-188).  chain_inds is a HOST array,
 * column-major (chains_per_super x nsuper), 0-based chain indices, as produced by
 * `_validate_superchain_ids` (src/rhat_nested.jl:68-81) minus one. */
int mcd_rhat_nested(mcd_ctx* ctx, const void* x, int mem, int dtype,
                    int64_t draws, int64_t chains, int64_t params,
                    const int32_t* chain_inds, int64_t chains_per_super, int64_t nsuper,
                    int kind, int split_chains,
                    void* rhat_out);

/* ---- transforms exposed for parity checks ---------------------------------------- */

/* StatsBase.tiedrank of each parameter's flattened slab (call site src/utils.jl:180):
 * average ranks as Float64, `draws*chains*params` values, same layout as x. */
int mcd_tiedrank(mcd_ctx* ctx, const void* x, int mem, int dtype,
                 int64_t draws, int64_t chains, int64_t params, double* ranks_out);
/* `_rank_normalize` (src/utils.jl:169-193): output has x's shape and dtype. */
int mcd_rank_normalize(mcd_ctx* ctx, const void* x, int mem, int dtype,
                       int64_t draws, int64_t chains, int64_t params, void* out);
/* `_fold_around_median` (src/utils.jl:148-158): output has x's shape and dtype. */
int mcd_fold_around_median(mcd_ctx* ctx, const void* x, int mem, int dtype,
                           int64_t draws, int64_t chains, int64_t params, void* out);

/* ---- synthetic input --------------------------------------------------------------- */

/* AR(1) chains as test/helpers.jl:4-12 (`ar1`): eps ~ N(0,1), x_1 = sigma*eps_1,
 * x_t = phi*x_{t-1} + sigma*eps_t, generated on the device with a counter-based RNG keyed
 * (seed, param_offset + param, chain, t), so any sharding of the parameter axis yields the
 * same values.  dev_x is DEVICE memory for draws*chains*params elements. */
int mcd_generate_ar1(mcd_ctx* ctx, int dtype, int64_t draws, int64_t chains, int64_t params,
                     int64_t param_offset, double phi, double sigma, uint64_t seed, void* dev_x);

/* Device memory helpers for hosts without a CUDA binding (Julia shim, ctypes tests). */
int mcd_device_alloc(mcd_ctx* ctx, int64_t bytes, void** dev_ptr);
int mcd_device_free(mcd_ctx* ctx, void* dev_ptr);
int mcd_memcpy_h2d(mcd_ctx* ctx, void* dev_dst, const void* host_src, int64_t bytes);
int mcd_memcpy_d2h(mcd_ctx* ctx, void* host_dst, const void* dev_src, int64_t bytes);
/* Pinned host memory (for full-rate, asynchronous staging of MCD_HOST inputs). */
int mcd_host_alloc(mcd_ctx* ctx, int64_t bytes, void** host_ptr);
int mcd_host_free(mcd_ctx* ctx, void* host_ptr);

#ifdef __cplusplus
}
#endif
#endif /* MCMCDIAG_B200_H */
