# baseline/ref_run.jl — dumps outputs of the REAL MCMCDiagnosticTools.jl for the committed golden inputs, so that parity
# can be pinned against the reference itself (SURVEY.md §8(c), last row).  No Julia exists in the build image or on the
# GPU box, so this script is unexecuted there; anyone with Julia runs
#
#     julia --project=/path/to/MCMCDiagnosticTools.jl baseline/ref_run.jl tests/golden
#
# (needs MCMCDiagnosticTools, NPZ and FFTW in the environment).  It reads every `*_inputs.npz` written by
# tests/golden/make_julia_inputs.py and writes `julia_reference_vectors.npz` next to them;
# tests/test_julia_reference_vectors.py consumes that file when it is present (CPU: the oracle; GPU: the CUDA path)
# with the north star's tolerances (1e-8 Float64, 1e-4 Float32, ranks bit-exact).
#
# Entry points exercised (the seam the C ABI replaces): `ess_rhat` (src/ess_rhat.jl:438-455), `ess` / `rhat` for every
# kind and autocovariance method (:276-349), `mcse` (src/mcse.jl:40-42), `rhat_nested` (src/rhat_nested.jl:43-66), and the
# rank transform internals `_rank_normalize` / `_fold_around_median` (src/utils.jl:148-193).
using MCMCDiagnosticTools
using NPZ
using FFTW          # FFTAutocovMethod needs an AbstractFFTs backend
using Statistics
using StatsBase: StatsBase

dir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
inputs = npzread(joinpath(dir, "julia_inputs.npz"))
out = Dict{String,Any}()

methods = Dict("direct" => AutocovMethod(), "fft" => FFTAutocovMethod(), "bda" => BDAAutocovMethod())
put!(name, v) = (out[name] = v isa Number ? [Float64(v)] : Float64.(collect(v)))

for (key, x) in inputs
    startswith(key, "x_") || continue
    tag = key[3:end]
    for kind in (:rank, :bulk, :tail, :basic)
        for (mname, meth) in methods, split in (1, 2, 3)
            r = ess_rhat(x; kind=kind, autocov_method=meth, split_chains=split)
            put!("ess_rhat.$tag.$kind.$mname.s$split.ess", r.ess)
            put!("ess_rhat.$tag.$kind.$mname.s$split.rhat", r.rhat)
        end
        for maxlag in (1, 10)
            r = ess_rhat(x; kind=kind, maxlag=maxlag)
            put!("ess_rhat.$tag.$kind.maxlag$maxlag.ess", r.ess)
        end
        put!("ess.$tag.$kind.relative", kind === :rank ? ess(x; kind=:bulk, relative=true) : ess(x; kind=kind, relative=true))
    end
    for (ename, est) in (("mean", Statistics.mean), ("median", Statistics.median), ("std", Statistics.std),
                         ("mad", StatsBase.mad), ("q25", Base.Fix2(Statistics.quantile, 0.25)))
        put!("ess.$tag.$ename", ess(x; kind=est))
        ename == "mad" || put!("mcse.$tag.$ename", mcse(x; kind=est))
    end
    put!("ranknorm.$tag", MCMCDiagnosticTools._rank_normalize(x))
    put!("fold.$tag", MCMCDiagnosticTools._fold_around_median(x))
end

if haskey(inputs, "nested_x")
    x, ids = inputs["nested_x"], Int.(inputs["nested_ids"])
    for kind in (:rank, :bulk, :tail, :basic), split in (1, 2)
        put!("rhat_nested.$kind.s$split", rhat_nested(x, ids; kind=kind, split_chains=split))
    end
end

npzwrite(joinpath(dir, "julia_reference_vectors.npz"), out)
println("wrote ", length(out), " arrays to ", joinpath(dir, "julia_reference_vectors.npz"))
