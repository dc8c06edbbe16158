# MCMCDiagB200.jl — Julia host shim over libmcmcdiag_b200.so (C ABI: include/mcmcdiag_b200.h).
#
# Drop-in for the ESS / R-hat hot path of MCMCDiagnosticTools.jl: it exports the reference's own
# names with the reference's signatures (src/MCMCDiagnosticTools.jl:19,23) and keeps in Julia
# exactly what SURVEY.md §8(b) assigns to the host: `kind` dispatch and the reference's
# exceptions (raised before `ccall`), the `niter <= 4` @warn, `missing` masking, Int -> Float64
# promotion, N-d / offset parameter axes via `similar`, `_maybescalar`, and the superchain label
# -> index matrix.  Everything numeric is one `ccall` per public call.  No CUDA.jl, no
# KernelAbstractions, no CPU fallback: if the library or a GPU is missing every call throws.
#
# NOTE: no Julia toolchain exists in the build image or on the GPU box, so this file is
# unexecuted there; the same host logic is implemented (and tested, tests/test_gpu_*.py) in
# Python in ../api.py.  Keep the two in sync.
module MCMCDiagB200

using Statistics: Statistics
using StatsBase: StatsBase
using SpecialFunctions: SpecialFunctions

export ess, ess_rhat, rhat, rhat_nested, mcse
export summary_columns
export gewekediag, heideldiag
export bfmi, chain_moments
export AutocovMethod, FFTAutocovMethod, BDAAutocovMethod
export ESSMethod, FFTESSMethod, BDAESSMethod

const LIB = get(ENV, "MCMCDIAG_B200_LIB", joinpath(@__DIR__, "..", "libmcmcdiag_b200.so"))

abstract type AbstractAutocovMethod end
struct AutocovMethod <: AbstractAutocovMethod end      # src/ess_rhat.jl:38
struct FFTAutocovMethod <: AbstractAutocovMethod end   # src/ess_rhat.jl:55
struct BDAAutocovMethod <: AbstractAutocovMethod end   # src/ess_rhat.jl:73
const ESSMethod = AutocovMethod                         # north-star spellings
const FFTESSMethod = FFTAutocovMethod
const BDAESSMethod = BDAAutocovMethod
_code(::AutocovMethod) = Cint(0)
_code(::FFTAutocovMethod) = Cint(1)
_code(::BDAAutocovMethod) = Cint(2)

const MCD_HOST = Cint(0)
_dtype(::Type{Float32}) = Cint(0)
_dtype(::Type{Float64}) = Cint(1)
const KIND = Dict(:basic => Cint(0), :bulk => Cint(1), :tail => Cint(2), :rank => Cint(3))

# ---- context (one per GPU, created lazily) -------------------------------------------------
const _ctx = Ref{Ptr{Cvoid}}(C_NULL)
function context(device::Integer=0)
    if _ctx[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:mcd_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device)
        rc == 0 || error("mcd_create failed ($rc): " * unsafe_string(ccall((:mcd_create_error, LIB), Cstring, ())))
        _ctx[] = h[]
        atexit(() -> ccall((:mcd_destroy, LIB), Cvoid, (Ptr{Cvoid},), _ctx[]))
    end
    return _ctx[]
end

function _check(rc::Cint, maxlag=nothing)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:mcd_last_error, LIB), Cstring, (Ptr{Cvoid},), context()))
    if rc == -1
        occursin("maxlag", msg) && maxlag !== nothing && throw(DomainError(maxlag, "maxlag must be >0."))
        throw(ArgumentError(msg))
    elseif rc == -5
        throw(ArgumentError(msg))          # Statistics.quantile on NaN data
    elseif rc == -3
        throw(OutOfMemoryError())
    end
    return error("libmcmcdiag_b200 error $rc: $msg")
end

# ---- shape helpers (src/utils.jl:197-215) -----------------------------------------------------
_param_dims(x::AbstractArray) = ntuple(i -> i + 2, max(0, ndims(x) - 2))
_param_axes(x::AbstractArray) = map(Base.Fix1(axes, x), _param_dims(x))
_maybescalar(x::AbstractArray{<:Any,0}) = x[]
_maybescalar(x::AbstractArray) = x
_floattype(x) = promote_type(nonmissingtype(eltype(x)), typeof(zero(nonmissingtype(eltype(x))) / 1))

# dense column-major (draws, chains, P) copy of the parameters without `missing`, plus the mask
function _pack(x::AbstractArray{<:Union{Missing,Real}})
    T = _floattype(x)
    T <: Union{Float32,Float64} || (T = Float64)
    draws = size(x, 1)
    chains = ndims(x) > 1 ? size(x, 2) : 1
    x3 = reshape(x, draws, chains, :)
    keep = [!any(ismissing, view(x3, :, :, p)) for p in axes(x3, 3)]
    dense = Array{T}(undef, draws, chains, count(keep))
    j = 0
    for p in axes(x3, 3)
        keep[p] || continue
        j += 1
        dense[:, :, j] .= view(x3, :, :, p)
    end
    return T, dense, keep
end

function _unpack(x, ::Type{T}, vals::Vector, keep) where {T}
    TM = Missing <: eltype(x) ? Union{Missing,T} : T
    out = similar(x, TM, _param_axes(x))
    lin = LinearIndices(out)
    j = 0
    for (p, k) in enumerate(keep)
        if k
            j += 1
            out[lin[p]] = vals[j]
        else
            out[lin[p]] = missing
        end
    end
    return _maybescalar(out)
end

_tailprob(tp::Rational) = (Float64(tp), Cint(0))
_tailprob(tp::Float32) = (Float64(tp), Cint(0))
_tailprob(tp::Real) = (Float64(tp), Cint(1))

function _ess_rhat_call(x, kind::Symbol, want_ess::Bool, want_rhat::Bool; relative::Bool=false,
                        autocov_method::AbstractAutocovMethod=AutocovMethod(), split_chains::Int=2,
                        maxlag::Int=250, tail_prob::Real=1//10)
    T, dense, keep = _pack(x)
    all(ismissing, x) && eltype(x) === Missing && return (similar(x, Missing, _param_axes(x)), similar(x, Missing, _param_axes(x)))
    niter = size(dense, 1) ÷ split_chains
    if want_ess
        if !(niter > 4)
            @warn "number of draws after splitting must be >4 but is $niter. ESS cannot be computed."
        else
            maxlag > 0 || throw(DomainError(maxlag, "maxlag must be >0."))
        end
    end
    P = size(dense, 3)
    S = Vector{T}(undef, want_ess ? P : 0)
    R = Vector{T}(undef, want_rhat ? P : 0)
    tp, tp64 = _tailprob(tail_prob)
    GC.@preserve dense S R begin
        rc = ccall((:mcd_ess_rhat, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Cint, Cint, Cint, Cint, Cdouble, Cint,
                    Ptr{Cvoid}, Ptr{Cvoid}),
                   context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, KIND[kind],
                   _code(autocov_method), split_chains, clamp(maxlag, -1, typemax(Cint)), relative, tp, tp64,
                   want_ess ? pointer(S) : C_NULL, want_rhat ? pointer(R) : C_NULL)
        _check(rc, maxlag)
    end
    return (want_ess ? _unpack(x, T, S, keep) : nothing, want_rhat ? _unpack(x, T, R, keep) : nothing)
end

# estimator -> (code, p, p_is_f64)   (src/ess_rhat.jl:628-659)
_estimator(::typeof(Statistics.mean)) = (Cint(0), 0.0, Cint(0))
_estimator(::typeof(Statistics.median)) = (Cint(1), 0.0, Cint(0))
_estimator(::typeof(Statistics.std)) = (Cint(2), 0.0, Cint(0))
_estimator(::typeof(StatsBase.mad)) = (Cint(3), 0.0, Cint(0))
_estimator(f::Base.Fix2{typeof(Statistics.quantile),<:Real}) = (Cint(4), Float64(f.x), Cint(f.x isa Float32 || f.x isa Rational ? 0 : 1))
_estimator(f) = nothing
# north-star symbol spellings
_estimator(s::Symbol) = s === :mean ? _estimator(Statistics.mean) : s === :median ? _estimator(Statistics.median) :
                        s in (:std, :squared) ? _estimator(Statistics.std) : s in (:mad, :abs, :folded) ? _estimator(StatsBase.mad) : nothing

function _estimator_call(fname::Symbol, x, est; relative::Bool=false, autocov_method::AbstractAutocovMethod=AutocovMethod(),
                         split_chains::Int=2, maxlag::Int=250)
    code, p, p64 = est
    T, dense, keep = _pack(x)
    niter = size(dense, 1) ÷ split_chains
    if !(niter > 4)
        @warn "number of draws after splitting must be >4 but is $niter. ESS cannot be computed."
    else
        maxlag > 0 || throw(DomainError(maxlag, "maxlag must be >0."))
    end
    P = size(dense, 3)
    out = Vector{T}(undef, P)
    GC.@preserve dense out begin
        rc = if fname === :mcd_mcse
            ccall((:mcd_mcse, LIB), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Cdouble, Cint, Cint, Cint, Cint, Ptr{Cvoid}),
                  context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, code, p, p64,
                  _code(autocov_method), split_chains, maxlag, out)
        else
            ccall((:mcd_ess_estimator, LIB), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Cdouble, Cint, Cint, Cint, Cint, Cint, Ptr{Cvoid}),
                  context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, code, p, p64,
                  _code(autocov_method), split_chains, maxlag, relative, out)
        end
        _check(rc, maxlag)
    end
    return _unpack(x, T, out, keep)
end

# ---- public API: same names, keywords and errors as the reference -------------------------------
"""`ess(samples; kind=:bulk, relative=false, autocov_method=AutocovMethod(), split_chains=2, maxlag=250, [tail_prob=1//10])`
(src/ess_rhat.jl:215-311)"""
function ess(samples::AbstractArray{<:Union{Missing,Real}}; kind=:bulk, kwargs...)
    if kind === :bulk || kind === :tail || kind === :basic
        return _ess_rhat_call(samples, kind, true, false; kwargs...)[1]
    elseif kind === :rank
        throw(ArgumentError("the `kind` `$kind` is not supported by `ess`"))
    end
    est = _estimator(kind)
    if est === nothing
        kind isa Symbol && throw(ArgumentError("the `kind` `$kind` is not supported by `ess`"))
        throw(ArgumentError("the estimator $kind is not yet supported by `ess`"))
    end
    return _estimator_call(:mcd_ess_estimator, samples, est; kwargs...)
end

"""`rhat(samples; kind=:rank, split_chains=2)` (src/ess_rhat.jl:313-420)"""
function rhat(samples::AbstractArray{<:Union{Missing,Real}}; kind::Symbol=:rank, split_chains::Int=2)
    haskey(KIND, kind) || throw(ArgumentError("the `kind` `$kind` is not supported by `rhat`"))
    return _ess_rhat_call(samples, kind, false, true; split_chains)[2]
end

"""`ess_rhat(samples; kind=:rank, kwargs...) -> (; ess, rhat)` (src/ess_rhat.jl:422-455)"""
function ess_rhat(samples::AbstractArray{<:Union{Missing,Real}}; kind::Symbol=:rank, kwargs...)
    haskey(KIND, kind) || throw(ArgumentError("the `kind` `$kind` is not supported by `ess_rhat`"))
    S, R = _ess_rhat_call(samples, kind, true, true; kwargs...)
    return (; ess=S, rhat=R)
end

"""`mcse(samples; kind=Statistics.mean, kwargs...)` (src/mcse.jl:5-42).  mean / std / median / quantile run on the
GPU; every other estimator uses the reference's subsampling bootstrap, which needs a Julia closure per window and
therefore stays in the reference package (SURVEY.md §2: out of scope)."""
function mcse(samples::AbstractArray{<:Union{Missing,Real}}; kind=Statistics.mean, kwargs...)
    est = _estimator(kind)
    (est === nothing || est[1] == 3) &&
        throw(ArgumentError("mcse for $kind uses the subsampling bootstrap of MCMCDiagnosticTools (src/mcse.jl:120-148)"))
    return _estimator_call(:mcd_mcse, samples, est; kwargs...)
end

const SUMMARY_FIELDS = (:mean, :std, :mcse_mean, :mcse_std, :ess_bulk, :ess_tail, :rhat)

"""`summary_columns(samples; fields=SUMMARY_FIELDS, autocov_method, split_chains, maxlag, tail_prob)`: the
per-parameter columns MCMCChains.summarystats / PosteriorStats.summarize assemble from separate calls
(`mean`, `std`, `mcse(; kind=mean)`, `mcse(; kind=std)`, `ess(; kind=:bulk)`, `ess(; kind=:tail)`, `rhat(; kind=:rank)`)
from ONE library call (`mcd_summary`): the host array crosses PCIe once.  Returns a NamedTuple of arrays."""
function summary_columns(samples::AbstractArray{<:Union{Missing,Real}}; fields=SUMMARY_FIELDS,
                         autocov_method::AbstractAutocovMethod=AutocovMethod(), split_chains::Int=2,
                         maxlag::Int=250, tail_prob::Real=1//10)
    all(f -> f in SUMMARY_FIELDS, fields) || throw(ArgumentError("unknown summary field in $fields"))
    names = Tuple(f for f in SUMMARY_FIELDS if f in fields)
    isempty(names) && throw(ArgumentError("no summary field requested"))
    mask = UInt32(sum(1 << (findfirst(==(f), SUMMARY_FIELDS) - 1) for f in names))
    T, dense, keep = _pack(samples)
    niter = size(dense, 1) ÷ split_chains
    if mask & 0x3c != 0
        niter > 4 ? (maxlag > 0 || throw(DomainError(maxlag, "maxlag must be >0."))) :
                    @warn "number of draws after splitting must be >4 but is $niter. ESS cannot be computed."
    end
    P = size(dense, 3)
    out = Matrix{T}(undef, P, length(names))
    tp, tp64 = _tailprob(tail_prob)
    GC.@preserve dense out begin
        rc = ccall((:mcd_summary, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cuint, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cvoid}),
                   context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, mask,
                   _code(autocov_method), split_chains, maxlag, tp, tp64, out)
        _check(rc, maxlag)
    end
    return NamedTuple{names}(Tuple(_unpack(samples, T, out[:, i], keep) for i in eachindex(names)))
end

# `_validate_superchain_ids` + `unique_indices` (src/rhat_nested.jl:68-81, src/utils.jl:50-64)
function _validate_superchain_ids(superchain_ids, nchains)
    length(superchain_ids) == nchains || throw(DimensionMismatch(
        "`superchain_ids` has length $(length(superchain_ids)) but `samples` has $nchains chains"))
    groups = Dict{eltype(superchain_ids),Vector{Int32}}()
    for (i, s) in enumerate(superchain_ids)
        push!(get!(() -> Int32[], groups, s), Int32(i - 1))
    end
    ks = sort!(collect(keys(groups)))
    length(ks) >= 2 || throw(ArgumentError("at least 2 superchains are required, got $(length(ks))"))
    allequal(length(groups[k]) for k in ks) || throw(ArgumentError("all superchains must contain the same number of chains"))
    return reduce(hcat, (groups[k] for k in ks))
end

"""`rhat_nested(samples, superchain_ids; kind=:rank, split_chains=2)` (src/rhat_nested.jl:1-66)"""
function rhat_nested(samples::AbstractArray{<:Union{Missing,Real}}, superchain_ids::AbstractVector;
                     kind::Symbol=:rank, split_chains::Int=2)
    ndims(samples) >= 2 || throw(ArgumentError("`samples` must have at least 2 dimensions `(draws, chains[, parameters…])`"))
    inds = _validate_superchain_ids(superchain_ids, size(samples, 2))
    haskey(KIND, kind) || throw(ArgumentError("the `kind` `$kind` is not supported by `rhat_nested`"))
    T, dense, keep = _pack(samples)
    P = size(dense, 3)
    out = Vector{T}(undef, P)
    GC.@preserve dense out inds begin
        rc = ccall((:mcd_rhat_nested, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Ptr{Int32}, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
                   context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, inds, size(inds, 1),
                   size(inds, 2), KIND[kind], split_chains, out)
        _check(rc)
    end
    return _unpack(samples, T, out, keep)
end


# ---- in-package callers of the path (src/gewekediag.jl:19-35, src/heideldiag.jl:16-71): unchanged host logic,
# their two `mcse(...; split_chains=1)` calls land on the device through `mcse` above --------------------------
function gewekediag(x::AbstractVector{<:Real}; first::Real=0.1, last::Real=0.5, kwargs...)
    0 < first < 1 || throw(ArgumentError("`first` is not in (0, 1)"))
    0 < last < 1 || throw(ArgumentError("`last` is not in (0, 1)"))
    first + last <= 1 || throw(ArgumentError("`first` and `last` proportions overlap"))
    n = length(x)
    x1 = x[1:round(Int, first * n)]
    x2 = x[round(Int, n - last * n + 1):n]
    s = hypot(Base.first(mcse(reshape(x1, :, 1, 1); split_chains=1, kwargs...)),
              Base.first(mcse(reshape(x2, :, 1, 1); split_chains=1, kwargs...)))
    z = (Statistics.mean(x1) - Statistics.mean(x2)) / s
    return (zscore=z, pvalue=SpecialFunctions.erfc(abs(z) / sqrt(2)))
end

function heideldiag(x::AbstractVector{<:Real}; alpha::Real=1//20, eps::Real=0.1, start::Int=1, kwargs...)
    n = length(x)
    delta = trunc(Int, 0.10 * n)
    y = x[trunc(Int, n / 2):end]
    T = typeof(zero(eltype(x)) / 1)
    s = Base.first(mcse(reshape(y, :, 1, 1); split_chains=1, kwargs...))
    S0 = length(y) * s^2
    i, pvalue, converged, ybar = 1, one(T), false, T(NaN)
    while i < n / 2
        y = x[i:end]
        m = length(y)
        ybar = Statistics.mean(y)
        B = cumsum(y) - ybar * collect(1:m)
        I = sum((B .* B) ./ (m * S0)) / m
        pvalue = 1 - T(_pcramer(I))
        converged = pvalue > alpha
        converged && break
        i += delta
    end
    s = Base.first(mcse(reshape(y, :, 1, 1); split_chains=1, kwargs...))
    halfwidth = sqrt(2) * SpecialFunctions.erfcinv(T(alpha)) * s
    return (burnin=i + start - 2, stationarity=converged, pvalue=pvalue, mean=ybar, halfwidth=halfwidth,
            test=halfwidth / abs(ybar) <= eps)
end

# Csorgo & Faraway (1996) series for the Cramer-von Mises distribution
function _pcramer(q::Real)
    p = 0.0
    for k in 0:3
        c1 = 4.0 * k + 1.0
        c2 = c1^2 / (16.0 * q)
        p += SpecialFunctions.gamma(k + 0.5) / factorial(k) * sqrt(c1) * exp(-c2) * SpecialFunctions.besselk(0.25, c2)
    end
    return p / (pi^1.5 * sqrt(q))
end


# ---- SURVEY §8(f)4: moment kernels behind a different combine ----------------------------------------------
"""`bfmi(energy; dims=1)` (src/bfmi.jl:36-43) on the device: one value per chain."""
function bfmi(energy::AbstractVector{<:Real})
    return first(bfmi(reshape(energy, :, 1)))
end
function bfmi(energy::AbstractMatrix{<:Real}; dims::Int=1)
    T = float(eltype(energy)) === Float32 ? Float32 : Float64
    e = Matrix{T}(dims == 1 ? energy : permutedims(energy))
    out = Vector{T}(undef, size(e, 2))
    GC.@preserve e out begin
        rc = ccall((:mcd_bfmi, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Cvoid}),
                   context(), e, MCD_HOST, _dtype(T), size(e, 1), size(e, 2), out)
        _check(rc)
    end
    return out
end

"""Per split-chain means and corrected variances, `(chains * split_chains, params)` each: the quantities
`_gelmandiag` (src/gelmandiag.jl:9-17) needs for `psrf` / `psrfci` (only the diagonals of W and B enter)."""
function chain_moments(samples::AbstractArray{<:Real,3}; split_chains::Int=1)
    T = float(eltype(samples)) === Float32 ? Float32 : Float64
    x = Array{T,3}(samples)
    nch, P = size(x, 2) * split_chains, size(x, 3)
    m = Matrix{T}(undef, nch, P); v = Matrix{T}(undef, nch, P)
    GC.@preserve x m v begin
        rc = ccall((:mcd_chain_moments, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                   context(), x, MCD_HOST, _dtype(T), size(x, 1), size(x, 2), P, split_chains, m, v)
        _check(rc)
    end
    return (mean=m, var=v)
end

end # module
