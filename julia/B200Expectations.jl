# B200Expectations.jl -- thin ccall shim that plugs libb200expect.so into SciMLExpectations.jl.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It is a literal mirror of the
# ctypes calls in scimlexpectations.jl_b200/_capi.py (which ARE tested, on a B200, through the same C
# ABI), kept short so a maintainer can check it against include/b200_expect.h line by line.  What can be
# checked without Julia is: tests/test_capi_load.py::test_julia_shim_matches_the_header compares the struct
# definitions below (field order and machine types) and every ccall's symbol and argument count with the header.
#
# What it does: defines `EnsembleB200 <: SciMLBase.EnsembleAlgorithm` for the `ensemblealg` slot of
# `SystemMap(prob, alg, ensemblealg; kwargs...)` (reference src/system_utils.jl:48-53) and specialises the
# two methods that consume that slot,
#     build_integrand(::ExpectationProblem{<:SystemMap{..,EnsembleB200}}, ::Koopman, mid, p, batch::Integer)
#                                                        (reference src/expectation.jl:169-200)
#     solve(::ExpectationProblem{<:SystemMap{..,EnsembleB200}}, ::MonteCarlo)
#                                                        (reference src/expectation.jl:268-294)
#     solve(::ExpectationProblem{<:SystemMap{..,EnsembleB200}}, ::Koopman, ::B200Cubature)   (optional: the whole
#                                                        h-adaptive solve as one C call, src/expectation.jl:336-354)
#     the same two methods for ExpectationProblem{<:ProcessNoiseSystemMap} whose forwarded solve arguments are
#     (EM() | LambaEM(), EnsembleB200(...))             (reference src/expectation.jl:209-249, :295-325)
# so that `solve(exprob, Koopman(); batch=..., quadalg=CubatureJLh())` and `solve(exprob, MonteCarlo(n))`
# keep their signatures.  Julia closures cannot run on the device: f, g, h travel as CUDA source
# (fields of `B200Model`); `h`/`g` of the structured forms (scatter x into u0/p, end-state components)
# can be generated with `hmap_scatter` / `observable_components` below.
module B200Expectations

using SciMLExpectations, SciMLBase, Distributions, Statistics
import KernelDensity                     # the reference module imports it too (src/SciMLExpectations.jl:4-10) but does not re-export it
import SciMLExpectations: build_integrand, ExpectationProblem, SystemMap, ProcessNoiseSystemMap, Koopman, MonteCarlo,
                          ExpectationSolution, GenericDistribution
import Integrals: BatchIntegralFunction, HCubatureJL

const LIB = get(ENV, "B200E_LIB", "libb200expect.so")

struct PdfDim            # b200e_pdf_dim
    kind::Int32; _pad::Int32
    lo::Float64; hi::Float64; mu::Float64; sigma::Float64
    table::Ptr{Float64}; ntable::Int64       # kind 4: density values on the uniform grid lo .. hi (copied by b200e_create)
end
struct CModel            # b200e_model (field order = include/b200_expect.h)
    struct_size::UInt32
    nstate::Int32; nparam::Int32; nx::Int32; nout::Int32
    solver::Int32; n_kkl::Int32; fp_mode::Int32
    source::Cstring
    u0::Ptr{Float64}; p::Ptr{Float64}
    t0::Float64; t1::Float64; abstol::Float64; reltol::Float64; dtmax::Float64; dt_fixed::Float64
    maxiters::Int64
    pdf::Ptr{PdfDim}
    n_devices::Int32; devices::Ptr{Int32}
    block_size::Int32; blocks_per_sm::Int32; refill_threshold::Int32; chunk_points::Int32
    schedule::Int32                     # 0 dynamic (default), 1 static, 2 dynamic + reproducible (B200E_SCHED_*)
    nsave::Int32; save_times::Ptr{Float64}   # time-sampled observables (saveat / sol(t)): nout = nsave * outputs per time
    cache_dir::Cstring
end
struct Counters; nf::UInt64; naccept::UInt64; nreject::UInt64; nfail::UInt64; end

"CUDA sources of the closures the hot path needs (dialect: csrc/kernels/expect_prelude.cuh)."
Base.@kwdef struct B200Model
    rhs::String                      # B200E_FN void rhs(real* du, const real* u, const real* p, real t)
    hmap::String                     # B200E_FN void hmap(const real* x, const real* u0n, const real* pn, real* u0, real* p)
    observable::String               # B200E_FN void observable(const real* u_end, const real* p, real t_end, real* out)
    nout::Int
    diffusion::String = ""           # process noise only: B200E_FN void diffusion(real* g, const real* u, const real* p, real t)
end

# One library handle (compiled kernels, streams, device buffers) per (problem, density) of an EnsembleB200.  The
# finalizer returns it to the library when the cache entry and every integrand closure holding it are gone, so
# repeated solves neither recompile nor grow device memory.
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(ptr::Ptr{Cvoid})
        h = new(ptr)
        finalizer(release!, h)
        return h
    end
end
function release!(h::Handle)
    if h.ptr != C_NULL
        ccall((:b200e_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr)
        h.ptr = C_NULL
    end
    return nothing
end
# ccall(..., (Ptr{Cvoid}, ...), h, ...): cconvert keeps the Handle itself rooted for the duration of the call
Base.cconvert(::Type{Ptr{Cvoid}}, h::Handle) = h
Base.unsafe_convert(::Type{Ptr{Cvoid}}, h::Handle) = h.ptr

Base.@kwdef struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    model::B200Model
    # factors of the product density, one per dimension of x: Uniform / Normal / truncated(Normal) /
    # KernelDensity.UnivariateKDE.  `nothing`: taken from GenericDistribution(dists...) (the product constructor,
    # src/distribution_utils.jl:40-48).  The 4-argument GenericDistribution(pdf_func, rand_func, lb, ub) of the KDE
    # tutorial (docs/src/tutorials/gpu_bayesian.md:163-166) hides its factors in a user closure: pass them here.
    dists::Any = nothing
    devices::Vector{Int32} = Int32[]           # empty: current device; several: sharded + one NCCL all-reduce
    fp32::Bool = false
    schedule::Int32 = 0                        # 0 dynamic work distribution, 1 static, 2 dynamic with bit-reproducible sums
    sampler::Symbol = :device                  # MonteCarlo: :device draws rand(d) in the kernel (counter-based
                                               # stream, reproducible by b200e_sample_host); :host ships rand(d)
    seed::UInt64 = 0x0000000000000000
    cache::Dict{UInt64, Handle} = Dict{UInt64, Handle}()   # handles by problem / density digest (see create_handle)
end

pdfdim(d::Uniform) = PdfDim(1, 0, minimum(d), maximum(d), 0.0, 1.0, C_NULL, 0)
pdfdim(d::Normal) = PdfDim(2, 0, -Inf, Inf, mean(d), std(d), C_NULL, 0)
pdfdim(d::Truncated{<:Normal}) = PdfDim(3, 0, minimum(d), maximum(d), mean(d.untruncated), std(d.untruncated), C_NULL, 0)
# a KernelDensity.UnivariateKDE (gpu_bayesian.md:153-166); `k.density` must stay alive until b200e_create returns
pdfdim(k::KernelDensity.UnivariateKDE) = PdfDim(4, 0, first(k.x), last(k.x), 0.0, 1.0, pointer(k.density), length(k.density))
pdfdim(d) = error("B200Expectations: no device pdf table for $(typeof(d)); supported: Uniform, Normal, truncated(Normal), " *
                  "Exponential, Gamma, Beta, LogNormal, Rayleigh, Chisq, Chi, Erlang, Pareto and their truncations, KernelDensity.UnivariateKDE")

# Log-polynomial factors (B200E_PDF_LOGPOLY): logpdf = c0 + c1 x + c2 x^2 + c3 log x + c4 log(1 - x) + c5 log(x)^2 on the
# support.  `logpoly_shape` gives c1 .. c5; c0 (the normaliser, and log(tp) of a truncation) is read off Distributions'
# own logpdf at an interior point, so no special function is restated here.
logpoly_shape(d) = nothing
logpoly_shape(d::Exponential) = (-1 / scale(d), 0.0, 0.0, 0.0, 0.0)
logpoly_shape(d::Gamma) = (-1 / scale(d), 0.0, shape(d) - 1, 0.0, 0.0)
logpoly_shape(d::Beta) = (0.0, 0.0, params(d)[1] - 1, params(d)[2] - 1, 0.0)
logpoly_shape(d::LogNormal) = (0.0, 0.0, params(d)[1] / params(d)[2]^2 - 1, 0.0, -1 / (2 * params(d)[2]^2))
logpoly_shape(d::Rayleigh) = (0.0, -1 / (2 * scale(d)^2), 1.0, 0.0, 0.0)
logpoly_shape(d::Chisq) = (-0.5, 0.0, dof(d) / 2 - 1, 0.0, 0.0)
logpoly_shape(d::Chi) = (0.0, -0.5, dof(d) - 1, 0.0, 0.0)
logpoly_shape(d::Erlang) = (-1 / scale(d), 0.0, shape(d) - 1.0, 0.0, 0.0)
logpoly_shape(d::Pareto) = (0.0, 0.0, -(shape(d) + 1), 0.0, 0.0)
logpoly_shape(d::Truncated) = d.untruncated isa Normal ? nothing : logpoly_shape(d.untruncated)
function logpoly(d)
    s = logpoly_shape(d)
    s === nothing && return nothing
    x = quantile(d, 0.5)
    c0 = logpdf(d, x) - (s[1] * x + s[2] * x^2 + s[3] * log(x) + (s[4] == 0 ? 0.0 : s[4] * log1p(-x)) + s[5] * log(x)^2)
    return Float64[c0, s...]
end
# `c` must stay alive until b200e_create returns (create_handle preserves it)
pdfdim(d, c::Vector{Float64}) = PdfDim(5, 0, minimum(d), maximum(d), 0.0, 1.0, pointer(c), 6)
pdfdim(d, ::Nothing) = pdfdim(d)
"true when the kernel's sampler can draw `rand(d)` itself (an inverse CDF in closed form)"
device_samplable(d) = d isa Uniform || d isa Normal || (d isa Truncated && d.untruncated isa Normal)

# what identifies a factor in the handle cache (a KDE by its grid ends and a digest of its values, not by identity)
factor_key(d) = d
factor_key(k::KernelDensity.UnivariateKDE) = (first(k.x), last(k.x), length(k.density), hash(k.density))

"Factors of the product density of `d` as the library needs them: explicit on the ensemble algorithm, else from the product constructor."
function density_factors(ens, d)
    ens.dists === nothing || return ens.dists
    f = d.pdf_func
    # GenericDistribution(dists...) builds pdf_func as a closure over `dists` (src/distribution_utils.jl:41-42)
    hasfield(typeof(f), :dists) && return getfield(f, :dists)
    error("B200Expectations: this GenericDistribution was built from a user pdf_func; pass its factors as EnsembleB200(dists = (...))")
end

check(rc, h) = rc == 0 ? nothing :
    error("libb200expect: ", unsafe_string(ccall((:b200e_last_error, LIB), Cstring, (Ptr{Cvoid},), h)), " (status $rc)")

"The library handle of `exprob`: created on first use, then served from the ensemble algorithm's cache."
create_handle(exprob::ExpectationProblem) =
    create_handle(exprob.S.prob, exprob.S.ensemblealg, exprob.S.kwargs, density_factors(exprob.S.ensemblealg, exprob.d))
function create_handle(prob, ens::EnsembleB200, kw, dists; solver = 0, n_kkl = 0, dt_fixed = 0.0,
                       default_tol = (1e-6, 1e-3))
    m = ens.model
    u0 = collect(Float64, prob.u0); p = collect(Float64, prob.p === nothing || prob.p isa SciMLBase.NullParameters ? Float64[] : prob.p)
    key = hash((u0, p, prob.tspan, collect(pairs(kw)), collect(pairs(prob.kwargs)), solver, n_kkl, dt_fixed, default_tol,
                map(factor_key, Tuple(dists)), m.rhs, m.diffusion, m.hmap, m.observable, m.nout, ens.fp32, ens.schedule, ens.devices))
    cached = get(ens.cache, key, nothing)
    cached === nothing || cached.ptr == C_NULL || return cached
    coefs = [logpoly(d) for d in dists]
    pdf = PdfDim[pdfdim(d, c) for (d, c) in zip(dists, coefs)]
    src = m.rhs * "\n" * m.diffusion * "\n" * m.hmap * "\n" * m.observable
    # saveat of the ODEProblem (docs/src/tutorials/introduction.md:207-212): the observable is then an
    # `observable_at(k, t, u, p, out)` source and nout = length(saveat) * outputs per save time
    saves = collect(Float64, get(prob.kwargs, :saveat, Float64[]))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve u0 p pdf src ens saves dists coefs begin
        cm = CModel(sizeof(CModel), length(u0), length(p), length(pdf), m.nout, solver, n_kkl, ens.fp32 ? 1 : 0,
                    Base.unsafe_convert(Cstring, src), pointer(u0), pointer(p),
                    prob.tspan[1], prob.tspan[2], get(kw, :abstol, default_tol[1]), get(kw, :reltol, default_tol[2]),
                    get(kw, :dtmax, 0.0), dt_fixed, get(kw, :maxiters, 100000), pointer(pdf),
                    length(ens.devices), isempty(ens.devices) ? Ptr{Int32}(C_NULL) : pointer(ens.devices),
                    0, 0, 0, 0, ens.schedule, length(saves), isempty(saves) ? Ptr{Float64}(C_NULL) : pointer(saves),
                    Cstring(C_NULL))
        rc = ccall((:b200e_create, LIB), Cint, (Ref{CModel}, Ref{Ptr{Cvoid}}), cm, h)
        rc == 0 || error("b200e_create: ", unsafe_string(ccall((:b200e_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    return ens.cache[key] = Handle(h[])
end

const B200Problem = ExpectationProblem{<:SystemMap{<:Any, <:Any, <:EnsembleB200}}

# ---- Koopman batch integrand (reference src/expectation.jl:169-200) --------------------------------
function build_integrand(prob::B200Problem, ::Koopman, mid, p, batch::Integer)   # the reference's signature: no extra arguments
    h = create_handle(prob)                       # captured by the closure below: alive as long as the integrand is
    nout = prob.S.ensemblealg.model.nout
    function integrand_koopman_systemmap_batch!(dx, x, _)
        npts = size(x)[end]                       # x :: dim x npts, column-major == point-major layout 0
        check(ccall((:b200e_eval_nodes, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Float64}),
                    h, x, npts, 0, 1, dx), h)     # weight_by_pdf = 1: g(sol,p) * pdf(d, x[:, i])
        return nothing
    end
    prototype = nout == 1 ? zeros(1) : zeros(nout, 1)
    integrand_koopman_systemmap_batch!(prototype, reshape(collect(mid), size(mid)..., 1), p)   # :197-198
    return BatchIntegralFunction{true}(integrand_koopman_systemmap_batch!, prototype; max_batch = batch)
end

# ---- Koopman with the cubature heap inside the library (reference src/expectation.jl:336-354) --------
"`quadalg = B200Cubature()`: h-adaptive Genz-Malik with the rule applied on the device (b200e_koopman_solve)."
struct B200Cubature end
struct CubatureOpts; reltol::Float64; abstol::Float64; maxevals::Int64; end
struct CubatureStats
    nevals::Int64; nrounds::Int64; nregions::Int64; max_batch::Int64
    converged::Int32; _pad::Int32
    kernel_ms::Float64; total_ms::Float64
    counters::Counters
end
# `quadalg` is a KEYWORD of the reference's solve (src/expectation.jl:336-343): this method intercepts it for B200 problems
# and hands every other quadrature algorithm to the reference's own method (which calls build_integrand above).
function SciMLBase.solve(exprob::B200Problem, expalg::Koopman, args...; quadalg = HCubatureJL(), kwargs...)
    if quadalg isa B200Cubature || any(a -> a isa B200Cubature, args)
        return koopman_in_library(exprob, create_handle(exprob), exprob.S.ensemblealg.model.nout; kwargs...)
    end
    return invoke(SciMLBase.solve, Tuple{ExpectationProblem, Koopman, Vararg{Any}}, exprob, expalg, args...;
                  quadalg = quadalg, kwargs...)
end
function koopman_in_library(exprob, h::Handle, nout; ireltol = 1e-2, iabstol = 1e-2, maxiters = 1000000, batch = nothing, kwargs...)
    lb, ub = collect(Float64, extrema(exprob.d)[1]), collect(Float64, extrema(exprob.d)[2])
    u = zeros(nout); resid = zeros(nout); st = Ref{CubatureStats}()
    GC.@preserve h begin
        check(ccall((:b200e_koopman_solve, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{CubatureOpts}, Ptr{Float64}, Ptr{Float64}, Ref{CubatureStats}),
                    h, lb, ub, CubatureOpts(ireltol, iabstol, maxiters), u, resid, st), h)
    end
    return ExpectationSolution(nout == 1 ? u[1] : u, nout == 1 ? resid[1] : resid, st[])
end

# ---- MonteCarlo (reference src/expectation.jl:268-294) ---------------------------------------------
function SciMLBase.solve(exprob::B200Problem, expalg::MonteCarlo; block = 1 << 20)
    h = create_handle(exprob)
    nout = exprob.S.ensemblealg.model.nout
    total = zeros(nout); s = zeros(nout); s2 = zeros(nout); c = Ref(Counters(0, 0, 0, 0))
    done = 0; n = expalg.trajectories
    ens = exprob.S.ensemblealg
    if ens.sampler === :device && all(device_samplable, density_factors(ens, exprob.d))
        # rand(d) hoisted into the kernel: samples 0 .. n-1 of the counter-based stream `seed`; the CPU arm gets
        # the identical set from b200e_sample_host(model, seed, 0, n, 0, x)
        check(ccall((:b200e_reduce_generated, LIB), Cint,
                    (Ptr{Cvoid}, UInt64, UInt64, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Counters}),
                    h, ens.seed, 0, n, s, s2, c), h)
        c[].nfail > 0 && @warn "libb200expect: $(c[].nfail) trajectories aborted (maxiters / non-finite / dt underflow)"
        total .+= s; done = n
    end
    while done < n
        m = min(block, n - done)
        x = reduce(hcat, (rand(exprob.d) for _ in 1:m))          # dim x m: host-generated sample set
        check(ccall((:b200e_reduce_samples, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Counters}),
                    h, x, m, 0, 0, s, s2, c), h)                   # weight 1, as expectation.jl:282
        c[].nfail > 0 && @warn "libb200expect: $(c[].nfail) trajectories aborted (maxiters / non-finite / dt underflow)"
        total .+= s; done += m
    end
    u = total ./ n                                                 # mean(sol.u), expectation.jl:293
    return ExpectationSolution(nout == 1 ? u[1] : u, nothing, nothing)
end

# ---- process noise (reference src/expectation.jl:209-249 Koopman, :295-325 MonteCarlo) ----------------
# ProcessNoiseSystemMap(prob, n, args...; kwargs...) forwards `args` to solve (src/system_utils.jl:98-100), so the
# ensemble algorithm is its second positional argument:  ProcessNoiseSystemMap(prob, 8, LambaEM(), EnsembleB200(...);
# abstol = 1e-3, reltol = 1e-3)  (test/processnoise.jl:15 plus the ensemble slot).  Dispatch is on the type of `args`.
const B200NoiseMap = ProcessNoiseSystemMap{<:Any, <:Tuple{Any, EnsembleB200}}
const B200NoiseProblem = ExpectationProblem{<:B200NoiseMap}

function create_noise_handle(exprob::B200NoiseProblem)
    S = exprob.S; alg, ens = S.args
    # the n-fold +-4 sigma truncated normal the reference's constructor builds (src/problem_types.jl:81-85)
    dists = density_factors(ens, exprob.d)
    name = string(nameof(typeof(alg)))
    if name == "LambaEM"                                   # adaptive: SciML's SDE default tolerances are 1e-2 / 1e-2
        return create_handle(S.prob, ens, S.kwargs, dists; solver = 2, n_kkl = S.n, default_tol = (1e-2, 1e-2))
    elseif name == "EM"                                    # fixed step: solve(...; dt = ...)
        return create_handle(S.prob, ens, S.kwargs, dists; solver = 1, n_kkl = S.n, dt_fixed = Float64(S.kwargs[:dt]))
    end
    error("B200Expectations: the process-noise path carries EM() and LambaEM(), not $name")
end

function build_integrand(prob::B200NoiseProblem, ::Koopman, mid, p, batch::Integer)
    h = create_noise_handle(prob)
    nout = prob.S.args[2].model.nout
    function integrand_koopman_processnoisesystemmap_batch!(dx, x, _)
        check(ccall((:b200e_eval_nodes, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Float64}),
                    h, x, size(x)[end], 0, 1, dx), h)     # Z = x[:, i]; g(sol, p) * pdf(d, x[:, i])  (:227)
        return nothing
    end
    prototype = nout == 1 ? zeros(1) : zeros(nout, 1)
    integrand_koopman_processnoisesystemmap_batch!(prototype, reshape(collect(mid), size(mid)..., 1), p)   # :245
    return BatchIntegralFunction{true}(integrand_koopman_processnoisesystemmap_batch!, prototype; max_batch = batch)
end

function SciMLBase.solve(exprob::B200NoiseProblem, expalg::Koopman, args...; quadalg = HCubatureJL(), kwargs...)
    if quadalg isa B200Cubature || any(a -> a isa B200Cubature, args)
        return koopman_in_library(exprob, create_noise_handle(exprob), exprob.S.args[2].model.nout; kwargs...)
    end
    return invoke(SciMLBase.solve, Tuple{ExpectationProblem, Koopman, Vararg{Any}}, exprob, expalg, args...;
                  quadalg = quadalg, kwargs...)
end

function SciMLBase.solve(exprob::B200NoiseProblem, expalg::MonteCarlo)
    h = create_noise_handle(exprob)
    ens = exprob.S.args[2]; nout = ens.model.nout; n = expalg.trajectories
    s = zeros(nout); s2 = zeros(nout); c = Ref(Counters(0, 0, 0, 0))
    check(ccall((:b200e_reduce_generated, LIB), Cint,     # Z = rand(d) drawn in the kernel (:305), weight 1
                (Ptr{Cvoid}, UInt64, UInt64, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Counters}),
                h, ens.seed, 0, n, s, s2, c), h)
    c[].nfail > 0 && @warn "libb200expect: $(c[].nfail) trajectories aborted (maxiters / non-finite / dt underflow)"
    u = s ./ n                                             # mean(sol.u), :324
    return ExpectationSolution(nout == 1 ? u[1] : u, nothing, nothing)
end

hmap_noise(n) =                                                 # cov(x, u, p) = x, p  (test/processnoise.jl:16): Z = x
    "B200E_FN void hmap(const real* x, const real* u0_nom, const real* p_nom, real* z, real* p) {\n" *
    join(["  z[$(k-1)] = x[$(k-1)];\n" for k in 1:n]) * "}\n"

# ---- structured h / g source generators --------------------------------------------------------------
hmap_scatter(; u0 = Pair{Int,Int}[], p = Pair{Int,Int}[]) =     # u0 = [i => j]: u0[i] = x[j]  (1-based in, 0-based out)
    "B200E_FN void hmap(const real* x, const real* u0_nom, const real* p_nom, real* u0, real* p) {\n" *
    join(["  u0[$(i-1)] = x[$(j-1)];\n" for (i, j) in u0]) * join(["  p[$(i-1)] = x[$(j-1)];\n" for (i, j) in p]) * "}\n"
observable_components(idx) =                                    # g(sol,p) = sol[idx, end]
    "B200E_FN void observable(const real* u, const real* p, real t_end, real* out) {\n" *
    join(["  out[$(k-1)] = u[$(i-1)];\n" for (k, i) in enumerate(idx)]) * "}\n"

observable_at_components(idx) =                                 # g(sol,p) = sol[idx, :] with saveat / [sol(t_k)[i] ...]
    "B200E_FN void observable_at(int k, real t, const real* u, const real* p, real* out) {\n" *
    join(["  out[$(j-1)] = u[$(i-1)];\n" for (j, i) in enumerate(idx)]) * "}\n"

export EnsembleB200, B200Model, B200Cubature, release!, hmap_scatter, hmap_noise, observable_components, observable_at_components
end # module
