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

export ess, ess_rhat, rhat, rhat_nested, mcse
export summary_columns
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

# ---- context (created lazily) -----------------------------------------------------------------
# One GPU by default (device 0, or ENV["MCMCDIAG_B200_DEVICES"] = "3"); with several ordinals,
# ENV["MCMCDIAG_B200_DEVICES"] = "0,1,2,3,4,5,6,7", ONE multi-GPU context (mcd_create_multi) shards the parameter axis of
# every host array over the listed GPUs inside the library: no CUDA.jl, no NCCL, no extra processes on the Julia side.
const _ctx = Ref{Ptr{Cvoid}}(C_NULL)
function _devices()
    spec = get(ENV, "MCMCDIAG_B200_DEVICES", "0")
    return Cint[parse(Cint, strip(t)) for t in split(spec, ',') if !isempty(strip(t))]
end
function context()
    if _ctx[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        devs = _devices()
        rc = length(devs) == 1 ?
            ccall((:mcd_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, devs[1]) :
            ccall((:mcd_create_multi, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), h, devs, length(devs))
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

# (draws, chains, P) view of the samples as a dense Array{T} plus the per-parameter `missing` mask.  Nothing is compacted:
# parameters that contain `missing` are handed to the library as they are (filler zero(T)) together with a skip mask
# (mcd_set_param_mask), which neither stages nor reads them and NaN-fills their outputs; `_unpack` maps those to `missing`
# (src/ess_rhat.jl:382-385,519-523).  For a plain Array{Float64/Float32} without Missing the "copy" is a reshape.
function _pack(x::AbstractArray{<:Union{Missing,Real}})
    T = _floattype(x)
    T <: Union{Float32,Float64} || (T = Float64)
    draws = size(x, 1)
    chains = ndims(x) > 1 ? size(x, 2) : 1
    x3 = reshape(x, draws, chains, :)
    if !(Missing <: eltype(x))
        dense = x3 isa Array{T,3} ? x3 : Array{T,3}(x3)
        return T, dense, nothing
    end
    skip = UInt8[any(ismissing, view(x3, :, :, p)) for p in axes(x3, 3)]
    dense = Array{T,3}(undef, draws, chains, size(x3, 3))
    @inbounds for i in eachindex(x3)
        v = x3[i]
        dense[i] = v === missing ? zero(T) : T(v)
    end
    return T, dense, skip
end

# hand the skip mask (if any) to the library: it is consumed by the next call on the context
function _set_mask(skip)
    (skip === nothing || !any(!iszero, skip)) && return nothing
    GC.@preserve skip begin
        rc = ccall((:mcd_set_param_mask, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), context(), skip, length(skip))
        _check(rc)
    end
    return nothing
end

function _unpack(x, ::Type{T}, vals::Vector, skip) where {T}
    TM = Missing <: eltype(x) ? Union{Missing,T} : T
    out = similar(x, TM, _param_axes(x))
    lin = LinearIndices(out)
    for p in eachindex(vals)
        out[lin[p]] = (skip !== nothing && skip[p] != 0) ? missing : vals[p]
    end
    return _maybescalar(out)
end

_clamp_maxlag(maxlag::Integer) = Cint(clamp(maxlag, -1, typemax(Cint)))

_tailprob(tp::Rational) = (Float64(tp), Cint(0))
_tailprob(tp::Float32) = (Float64(tp), Cint(0))
_tailprob(tp::Real) = (Float64(tp), Cint(1))

function _ess_rhat_call(x, kind::Symbol, want_ess::Bool, want_rhat::Bool; relative::Bool=false,
                        autocov_method::AbstractAutocovMethod=AutocovMethod(), split_chains::Int=2,
                        maxlag::Int=250, tail_prob::Real=1//10)
    if eltype(x) === Missing   # nothing but `missing`: no float type to compute in (checked before `_pack`)
        return (want_ess ? similar(x, Missing, _param_axes(x)) : nothing, want_rhat ? similar(x, Missing, _param_axes(x)) : nothing)
    end
    T, dense, keep = _pack(x)
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
    _set_mask(keep)
    GC.@preserve dense S R begin
        rc = ccall((:mcd_ess_rhat, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Cint, Cint, Cint, Cint, Cdouble, Cint,
                    Ptr{Cvoid}, Ptr{Cvoid}),
                   context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, KIND[kind],
                   _code(autocov_method), split_chains, _clamp_maxlag(maxlag), relative, tp, tp64,
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
    _set_mask(keep)
    GC.@preserve dense out begin
        rc = if fname === :mcd_mcse
            ccall((:mcd_mcse, LIB), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Cdouble, Cint, Cint, Cint, Cint, Ptr{Cvoid}),
                  context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, code, p, p64,
                  _code(autocov_method), split_chains, _clamp_maxlag(maxlag), out)
        else
            ccall((:mcd_ess_estimator, LIB), Cint,
                  (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cint, Cdouble, Cint, Cint, Cint, Cint, Cint, Ptr{Cvoid}),
                  context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, code, p, p64,
                  _code(autocov_method), split_chains, _clamp_maxlag(maxlag), relative, out)
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

"""`mcse(samples; kind=Statistics.mean, kwargs...)` (src/mcse.jl:5-42).  mean / std / median / quantile run on the GPU
(the ESS-based rules, src/mcse.jl:45-118).  Every other estimator takes the reference's subsampling-bootstrap fallback
`_mcse_sbm(f, x; batch_size)` (src/mcse.jl:120-148), which evaluates a Julia closure per batch: it is host logic in the
reference and stays host logic here, by delegating to the loaded MCMCDiagnosticTools package (the drop-in keeps the
reference's own code for it, SURVEY.md §2)."""
function mcse(samples::AbstractArray{<:Union{Missing,Real}}; kind=Statistics.mean, kwargs...)
    est = _estimator(kind)
    if est === nothing || est[1] == 3
        f = kind isa Symbol ? StatsBase.mad : kind
        return _reference_mcse_sbm(f, samples; kwargs...)
    end
    haskey(kwargs, :relative) && kwargs[:relative] &&
        throw(ArgumentError("mcse(...; relative=true) is not supported by the accelerated path"))
    return _estimator_call(:mcd_mcse, samples, est; (k => v for (k, v) in kwargs if k !== :relative)...)
end

# `_mcse_sbm` of the reference package itself (found among the loaded modules: the shim does not depend on it)
function _reference_mcse_sbm(f, samples; kwargs...)
    id = Base.PkgId(Base.UUID("be115224-59cd-429b-ad48-344e309966f0"), "MCMCDiagnosticTools")
    pkg = get(Base.loaded_modules, id, nothing)
    pkg === nothing && throw(ArgumentError(
        "mcse for $f uses the subsampling bootstrap of MCMCDiagnosticTools (src/mcse.jl:120-148): load MCMCDiagnosticTools next to this shim"))
    return pkg._mcse_sbm(f, samples; kwargs...)
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
    _set_mask(keep)
    GC.@preserve dense out begin
        rc = ccall((:mcd_summary, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Cuint, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cvoid}),
                   context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, mask,
                   _code(autocov_method), split_chains, _clamp_maxlag(maxlag), tp, tp64, out)
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
    _set_mask(keep)
    GC.@preserve dense out inds begin
        rc = ccall((:mcd_rhat_nested, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Ptr{Int32}, Int64, Int64, Cint, Cint, Ptr{Cvoid}),
                   context(), dense, MCD_HOST, _dtype(T), size(dense, 1), size(dense, 2), P, inds, size(inds, 1),
                   size(inds, 2), KIND[kind], split_chains, out)
        _check(rc)
    end
    return _unpack(samples, T, out, keep)
end


# (`gewekediag` and `heideldiag`, src/gewekediag.jl:19-35 / src/heideldiag.jl:16-71, are NOT redefined here: they live outside the
# C boundary, and the package's own definitions reach the device unchanged through the `mcse` above once it is the
# `mcse` in scope -- see INTEGRATION.md.)

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
