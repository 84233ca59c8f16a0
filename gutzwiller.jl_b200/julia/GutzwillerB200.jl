# GutzwillerB200.jl -- Julia host shim over libgutzwiller_b200.so (C ABI: include/gutzwiller_b200.h).
#
# Keeps the call forms of mtsch/Gutzwiller.jl for the HubbardReal1D + GutzwillerAnsatz path:
#   kinetic_vqmc / kinetic_vqmc! / KineticVQMC / LocalEnergyEvaluator / val_and_grad / val_err_and_grad /
#   local_energy_estimator / mean / var / amsgrad / gradient_descent
# (reference: src/kinetic-vqmc/qmc.jl:73-146, state.jl:137-164,265-298, wrapper.jl:45-121, localenergy.jl:81-173,
#  amsgrad.jl:105-263).  Every method is a thin `ccall`; no arithmetic on sampled data happens in Julia.
#
# STATUS: UNEXECUTED.  Julia is not installed in the build image (nor on the GPU boxes), so this file has never been
# run; it is kept correct by inspection and by tests/test_capi_and_host.py, which checks every `ccall` (symbol,
# argument count and kinds) and every mirrored struct (field order, types, array lengths) against the C header that
# gcc compiles for tests/c_consumer.c.  The Python mirror (gutzwiller.jl_b200/*.py) binds exactly the same symbols
# and is what the GPU test-suite drives; see INTEGRATION.md for how a maintainer wires this module into Gutzwiller.jl.
module GutzwillerB200

using Rimu
using Rimu.StatsTools: BlockingResult
using StaticArrays
using Statistics

export GPUContext, GPUKineticVQMCResult, GPUKineticVQMC, GPULocalEnergyEvaluator, GPUGradientDescentResult
export GPUAnsatz, GPUGenericVQMC, GPUGenericLocalEnergyEvaluator
export GPUSparseHamiltonian, GPUSparseVQMC, GPUSparseLocalEnergyEvaluator, sampled_vector, recorded_series
export kinetic_vqmc, kinetic_vqmc!, val_and_grad, val_err_and_grad, local_energy_estimator
export amsgrad, gradient_descent

const LIB = get(ENV, "GUTZWILLER_B200_LIB",
                joinpath(@__DIR__, "..", "lib", "libgutzwiller_b200.so"))

const GWZ_KERNEL_ENUMERATE = Cint(0)
const GWZ_KERNEL_BITPARALLEL = Cint(1)
const GWZ_RUN_KEEP_ACC = Cuint(1)
const GWZ_RUN_POOLED = Cuint(2)
const GWZ_RUN_SERIES = Cuint(4)
const GWZ_RUN_BLOCKING = Cuint(8)
const GWZ_NUM_PARAMS = 1
const GWZ_NUM_ACC = 13
const DEFAULT_SEED = 0x9E3779B97F4A7C15

# mirrors `struct gwz_blocking_summary`
struct BlockingSummary
    mean::Cdouble
    err::Cdouble
    err_err::Cdouble
    err_ok::Cdouble
    grad::NTuple{GWZ_NUM_PARAMS,Cdouble}
    walkers::UInt64
    failed::UInt64
    k_min::Int32
    k_max::Int32
    series_steps::UInt64
    resample_len::UInt64
    kernel_ms::Cfloat
    reserved::Cfloat
end
BlockingSummary() = BlockingSummary(0.0, 0.0, 0.0, 0.0, (0.0,), 0, 0, Int32(-1), Int32(-1), 0, 0, 0f0, 0f0)

# mirrors `struct gwz_vqmc_result`
struct VqmcResult
    val::Cdouble
    err::Cdouble
    grad::NTuple{GWZ_NUM_PARAMS,Cdouble}
    grad_err::NTuple{GWZ_NUM_PARAMS,Cdouble}
    pooled_val::Cdouble
    pooled_grad::NTuple{GWZ_NUM_PARAMS,Cdouble}
    pooled_var::Cdouble
    walker_var::Cdouble
    total_time::Cdouble
    walkers::UInt64
    steps::UInt64
    hops::UInt64
    acc::NTuple{GWZ_NUM_ACC,Cdouble}
    kernel_ms::Cfloat
    reserved::Cfloat
    blocking::BlockingSummary
end
VqmcResult() = VqmcResult(0.0, 0.0, (0.0,), (0.0,), 0.0, (0.0,), 0.0, 0.0, 0.0, 0, 0, 0,
                          ntuple(_ -> 0.0, GWZ_NUM_ACC), 0f0, 0f0, BlockingSummary())

function check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:gwz_last_error, LIB), Cstring, ()))
    rc == 1 && throw(ArgumentError(msg))           # GWZ_ERR_INVALID
    rc == 4 && error(msg)                          # "Infinite residence time" (qmc.jl:34-36)
    rc == 5 && throw(OutOfMemoryError())           # GWZ_ERR_NOMEM
    error("libgutzwiller_b200 [status $rc]: $msg")
end

# ---------------------------------------------------------------------------------------------
# handles
# ---------------------------------------------------------------------------------------------
mutable struct GPUContext
    ptr::Ptr{Cvoid}
    function GPUContext(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:gwz_context_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        ctx = new(r[])
        finalizer(c -> ccall((:gwz_context_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), ctx)
        return ctx
    end
end
const DEFAULT_CONTEXT = Ref{Union{Nothing,GPUContext}}(nothing)
function default_context()
    if isnothing(DEFAULT_CONTEXT[])
        DEFAULT_CONTEXT[] = GPUContext(0)
    end
    return DEFAULT_CONTEXT[]::GPUContext
end

"Join the contexts of several devices of this process in one NCCL communicator."
function comm_init_all(ctxs::Vector{GPUContext})
    ptrs = [c.ptr for c in ctxs]
    check(ccall((:gwz_context_comm_init_all, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), ptrs, length(ptrs)))
end

# u and t are type parameters of Rimu's HubbardReal1D{TT,A,U,T}
_ut(::HubbardReal1D{<:Any,<:Any,U,T}) where {U,T} = (Float64(U), Float64(T))

mutable struct GPUModel
    ptr::Ptr{Cvoid}
    N::Int
    M::Int
    W::Int
    ctx::GPUContext
end
function GPUModel(H::HubbardReal1D, ctx::GPUContext=default_context())
    addr = starting_address(H)
    N, M = num_particles(addr), num_modes(addr)
    u, t = _ut(H)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gwz_model_create_hubbard1d_gutzwiller, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble, Ref{Ptr{Cvoid}}), ctx.ptr, N, M, u, t, r))
    m = GPUModel(r[], N, M, cld(N + M - 1, 64), ctx)
    finalizer(x -> ccall((:gwz_model_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), m)
    return m
end

# BoseFS <-> W 64-bit words through occupation numbers (independent of Rimu's chunk order)
function address_words(m::GPUModel, addr::BoseFS)
    words = zeros(UInt64, m.W)
    check(ccall((:gwz_model_from_onr, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{UInt64}),
                m.ptr, Int32.(collect(onr(addr))), words))
    return words
end
function address_from_words(m::GPUModel, ::Type{A}, words::AbstractVector{UInt64}) where {A<:BoseFS}
    o = zeros(Int32, m.M)
    check(ccall((:gwz_model_to_onr, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Int32}), m.ptr, words, o))
    return A(Tuple(Int.(o)))
end

# ---------------------------------------------------------------------------------------------
# kinetic_vqmc / KineticVQMCResult  (qmc.jl:73-146, state.jl:137-164,265-298)
# ---------------------------------------------------------------------------------------------
"""
    GPUKineticVQMCResult

Stands in for `KineticVQMCResult` (state.jl:75-77): the walkers live in HBM behind `walkers`; the parameters are kept
with the result like `KineticVQMCWalkerState.params` (state.jl:13), so `kinetic_vqmc!(res; steps)` needs none.
The (residence_time, local_energy) series of the accumulated steps stay in HBM (`GWZ_RUN_SERIES`) for
`val_err_and_grad(res)` / `local_energy_estimator(res)`.
"""
mutable struct GPUKineticVQMCResult{H,A}
    hamiltonian::H
    ansatz::A
    model::GPUModel
    walkers::Ptr{Cvoid}
    n_walkers::Int
    params::SVector{GWZ_NUM_PARAMS,Float64}
    steps::Int          # per walker, accumulated in the estimators
    burn_in::Int
    keep_series::Bool
    last::VqmcResult
end
Base.length(res::GPUKineticVQMCResult) = res.steps * res.n_walkers      # sum(length, res.states)
Base.keytype(res::GPUKineticVQMCResult) = typeof(starting_address(res.hamiltonian))
Base.valtype(::GPUKineticVQMCResult) = Float64

function _make_walkers(model::GPUModel, H, walkers::Integer; seed::UInt64=DEFAULT_SEED, offset::Integer=0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    start = address_words(model, starting_address(H))      # state.jl:26
    check(ccall((:gwz_walkers_create, LIB), Cint,
                (Ptr{Cvoid}, UInt64, UInt64, Ptr{UInt64}, UInt64, Ref{Ptr{Cvoid}}),
                model.ptr, walkers, offset, start, seed, r))
    return r[]
end

# one gwz_vqmc_run; flags select what is kept (series) and analysed (blocking) on the device
function _run!(res::GPUKineticVQMCResult, steps::Integer; keep::Bool, flags::Cuint=Cuint(0),
               kernel::Cint=GWZ_KERNEL_BITPARALLEL)
    p = Float64[res.params[1]]
    out = Ref(VqmcResult())
    fl = flags | (keep ? GWZ_RUN_KEEP_ACC : Cuint(0)) | (res.keep_series ? GWZ_RUN_SERIES : Cuint(0))
    burn = (keep && res.steps > 0) ? 0 : res.burn_in
    check(ccall((:gwz_vqmc_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, UInt64, UInt64, Cint, Cuint, Ref{VqmcResult}),
                res.walkers, p, steps, burn, kernel, fl, out))
    res.last = out[]
    res.steps = keep ? res.steps + steps : steps
    return res
end

_default_walkers(samples) = min(2^20, prevpow(2, max(1, round(Int, samples) ÷ 1024)))

"""
    kinetic_vqmc(H::HubbardReal1D, ansatz::GutzwillerAnsatz, params; samples=1e6, walkers, steps)

Drop-in for `Gutzwiller.kinetic_vqmc` (qmc.jl:129-146) running on the GPU.  `walkers` defaults to the
largest power of two leaving >= 1024 steps per walker (the reference default is thread-count derived,
qmc.jl:132); an explicit `walkers`/`steps`/`samples` is honoured exactly.  `keep_series=false` drops the per-step
series (16 B of HBM per walker-step); `val_err_and_grad(res)` then reports the cross-walker standard error.
"""
function kinetic_vqmc(H::HubbardReal1D, ansatz, params;
                      samples=1e6,
                      walkers=_default_walkers(samples),
                      steps=round(Int, samples / walkers),
                      burn_in=0, seed=DEFAULT_SEED, keep_series=true, ctx=default_context())
    model = GPUModel(H, ctx)
    w = _make_walkers(model, H, walkers; seed)
    res = GPUKineticVQMCResult(H, ansatz, model, w, Int(walkers), SVector{GWZ_NUM_PARAMS,Float64}(params...), 0,
                               Int(burn_in), keep_series, VqmcResult())
    finalizer(r -> ccall((:gwz_walkers_destroy, LIB), Cint, (Ptr{Cvoid},), r.walkers), res)
    return _run!(res, steps; keep=false)
end

"`kinetic_vqmc!(res; steps=length(res.states[1]))` (qmc.jl:108-111): continue from the walkers' current addresses."
kinetic_vqmc!(res::GPUKineticVQMCResult; steps=res.steps) = _run!(res, steps; keep=true)

# state.jl:137-149
val_and_grad(res::GPUKineticVQMCResult) = (res.last.val, SVector(res.last.grad))

# gwz_walkers_blocking: per-walker blocking_analysis(resample(tau, E; len)) on the device
function _blocking(res::GPUKineticVQMCResult; resample_length::Integer=0, per_walker::Bool=false)
    n = per_walker ? res.n_walkers : 0
    mean_w, err_w, ee_w = zeros(n), zeros(n), zeros(n)
    k_w, blocks_w = zeros(Int32, n), zeros(Int64, n)
    summ = Ref(BlockingSummary())
    # arrays are passed as such (ccall keeps them alive); C_NULL = "not wanted"
    check(ccall((:gwz_walkers_blocking, LIB), Cint,
                (Ptr{Cvoid}, UInt64, Ref{BlockingSummary}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ptr{Int64}),
                res.walkers, resample_length, summ,
                per_walker ? mean_w : C_NULL, per_walker ? err_w : C_NULL, per_walker ? ee_w : C_NULL,
                per_walker ? k_w : C_NULL, per_walker ? blocks_w : C_NULL))
    return summ[], mean_w, err_w, ee_w, k_w, blocks_w
end

"""
    val_err_and_grad(res::GPUKineticVQMCResult)

state.jl:150-164: per walker `blocking_analysis(resample(residence_times, local_energies))`, value = mean of the
blocking means, error = sqrt(sum of err^2) / walkers, gradient with the blocking mean -- computed by the device kernel
over the series kept in HBM.  Without kept series (`keep_series=false`) the error is the cross-walker standard error.
"""
function val_err_and_grad(res::GPUKineticVQMCResult)
    if res.keep_series
        b = res.last.blocking.walkers > 0 ? res.last.blocking : _blocking(res)[1]
        return b.mean, b.err, SVector(b.grad)
    end
    return res.last.val, res.last.err, SVector(res.last.grad)
end

"`CombinedBlockingResult` (state.jl:238-256) of a GPU run; `results` as in the reference for up to 2^16 walkers."
struct GPUCombinedBlockingResult
    mean::Float64
    err::Float64
    err_err::Float64
    results::Vector{BlockingResult{Float64}}
    k_range::Tuple{Int,Int}
    failed::Int
end
function Base.show(io::IO, r::GPUCombinedBlockingResult)
    println(io, "CombinedBlockingResult{Float64}")
    println(io, "  mean = ", r.mean, " ± ", r.err)
    println(io, "  with uncertainty of ± ", r.err_err)
    if r.failed > 0
        print(io, "  Blocking unsuccessful for ", r.failed, " walkers.")
    else
        print(io, "  Combined blocking results. (k ∈ ", r.k_range[1], " … ", r.k_range[2], ")")
    end
end

"`local_energy_estimator(res; resample_length=length(first(res.states)))` (state.jl:265-285)."
function local_energy_estimator(res::GPUKineticVQMCResult; resample_length=res.steps)
    res.keep_series || throw(ArgumentError("this run kept no series (keep_series=false)"))
    per_walker = res.n_walkers <= 2^16
    b, m, e, ee, k, blocks = _blocking(res; resample_length, per_walker)
    results = [BlockingResult(m[i], e[i], ee[i], e[i]^2, Int(k[i]), Int(blocks[i])) for i in 1:length(m)]
    return GPUCombinedBlockingResult(b.mean, b.err, b.err_err, results, (Int(b.k_min), Int(b.k_max)), Int(b.failed))
end

Statistics.mean(res::GPUKineticVQMCResult) = res.last.val            # state.jl:287-292
Statistics.var(res::GPUKineticVQMCResult) = res.last.walker_var      # state.jl:293-298

function Base.show(io::IO, res::GPUKineticVQMCResult)                 # state.jl:78-86
    v, e, _ = val_err_and_grad(res)
    println(io, "KineticVQMCResult (GPU)")
    println(io, "  walkers:      ", res.n_walkers)
    println(io, "  samples:      ", length(res))
    print(io, "  local energy: ", round(v; sigdigits=5), " ± ", round(e; sigdigits=5))
end

"""
    recorded_series(res; steps) -> (residence_times, local_energies, addresses)

A recording run (`gwz_vqmc_run_record`): `steps` more steps whose per-step rows (state.jl:91-132) are shipped to the
host: `residence_times[step, walker]`, `local_energies[step, walker]`, `addresses[step, walker]`.
"""
function recorded_series(res::GPUKineticVQMCResult{H}; steps::Integer=res.steps) where {H}
    n, W = res.n_walkers, res.model.W
    tau = zeros(Cdouble, n, steps)            # C layout [step][walker] = column-major (walker, step)
    eloc = zeros(Cdouble, n, steps)
    addr = zeros(UInt64, W, n, steps)
    out = Ref(VqmcResult())
    p = Float64[res.params[1]]
    check(ccall((:gwz_vqmc_run_record, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, UInt64, Cint, Cuint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt64}, Ref{VqmcResult}),
                res.walkers, p, steps, GWZ_KERNEL_BITPARALLEL, Cuint(0), tau, eloc, addr, out))
    res.last = out[]
    res.steps = steps
    A = typeof(starting_address(res.hamiltonian))
    addresses = [address_from_words(res.model, A, addr[:, w, k]) for k in 1:steps, w in 1:n]
    return permutedims(tau), permutedims(eloc), addresses
end

"""
    sampled_vector(res; steps) -> DVec

`DVec(res)` (state.jl:303-319) for single-word addresses: `steps` more steps whose residence times are deposited on
their addresses in an HBM hash table (`gwz_vqmc_sample_vector`), normalised.
"""
function sampled_vector(res::GPUKineticVQMCResult; steps::Integer=res.steps)
    cap = UInt64(min(big(steps) * res.n_walkers, big(dimension(res.hamiltonian))))
    addr, vals = zeros(UInt64, cap), zeros(Cdouble, cap)
    count, out = Ref{UInt64}(0), Ref(VqmcResult())
    p = Float64[res.params[1]]
    check(ccall((:gwz_vqmc_sample_vector, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, UInt64, Cint, Cuint, UInt64, Ptr{UInt64}, Ptr{Cdouble}, Ref{UInt64}, Ref{VqmcResult}),
                res.walkers, p, steps, GWZ_KERNEL_BITPARALLEL, Cuint(0), cap, addr, vals, count, out))
    res.last = out[]
    res.steps = steps
    A = typeof(starting_address(res.hamiltonian))
    dv = DVec{A,Float64}()
    for i in 1:Int(count[])
        dv[address_from_words(res.model, A, addr[i:i])] = vals[i]
    end
    return dv
end

# ---------------------------------------------------------------------------------------------
# KineticVQMC (wrapper.jl:45-121)
# ---------------------------------------------------------------------------------------------
struct GPUKineticVQMC{R}
    steps::Int
    walkers::Int
    result::R
end
function GPUKineticVQMC(H::HubbardReal1D, ansatz; samples=1e6,
                        walkers=_default_walkers(samples),
                        steps=round(Int, samples / walkers),                               # wrapper.jl:56
                        burn_in=0, seed=DEFAULT_SEED, ctx=default_context())
    model = GPUModel(H, ctx)
    w = _make_walkers(model, H, walkers; seed)              # all walkers at starting_address(H), nothing sampled yet
    res = GPUKineticVQMCResult(H, ansatz, model, w, Int(walkers), SVector{GWZ_NUM_PARAMS,Float64}(0.0), 0,
                               Int(burn_in), false, VqmcResult())
    finalizer(r -> ccall((:gwz_walkers_destroy, LIB), Cint, (Ptr{Cvoid},), r.walkers), res)
    return GPUKineticVQMC(Int(steps), Int(walkers), res)
end
# _reset! (wrapper.jl:74-79): new parameters, estimators cleared, walker positions kept
function _reset_and_run!(ke::GPUKineticVQMC, params, steps; flags::Cuint=Cuint(0))
    ke.result.params = SVector{GWZ_NUM_PARAMS,Float64}(params...)
    return _run!(ke.result, steps; keep=false, flags)
end
function (ke::GPUKineticVQMC)(params; samples=ke.steps * ke.walkers, steps=samples ÷ ke.walkers)
    return _reset_and_run!(ke, params, steps).last.val                                      # wrapper.jl:81-88
end
function val_and_grad(ke::GPUKineticVQMC, params; samples=ke.steps * ke.walkers, steps=samples ÷ ke.walkers)
    return val_and_grad(_reset_and_run!(ke, params, steps))                                 # wrapper.jl:90-97
end
function val_err_and_grad(ke::GPUKineticVQMC, params; samples=ke.steps * ke.walkers, steps=samples ÷ ke.walkers)
    b = _reset_and_run!(ke, params, steps; flags=GWZ_RUN_BLOCKING).last.blocking            # wrapper.jl:98-105
    return b.mean, b.err, SVector(b.grad)
end
function (ke::GPUKineticVQMC)(F, G, params)                                                  # wrapper.jl:107-121
    res = _reset_and_run!(ke, params, ke.steps)
    !isnothing(G) && (G .= res.last.grad)
    return isnothing(F) ? nothing : res.last.val
end

# ---------------------------------------------------------------------------------------------
# LocalEnergyEvaluator (localenergy.jl:81-173)
# ---------------------------------------------------------------------------------------------
mutable struct GPULocalEnergyEvaluator{H,A}
    hamiltonian::H
    ansatz::A
    model::GPUModel
    sweep::Ptr{Cvoid}
end
function GPULocalEnergyEvaluator(H::HubbardReal1D, ansatz; basis=nothing, ctx=default_context())
    model = GPUModel(H, ctx)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    if isnothing(basis)   # enumerate all C(N+M-1,N) states by rank on the GPU
        check(ccall((:gwz_sweep_create, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{UInt64}, UInt64, UInt64, UInt64, Ref{Ptr{Cvoid}}),
                    model.ptr, C_NULL, 0, 0, 0, r))
    else                  # stored basis, e.g. build_basis(H): dim x W words, row-major
        words = reduce(vcat, (address_words(model, b) for b in basis))
        check(ccall((:gwz_sweep_create, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{UInt64}, UInt64, UInt64, UInt64, Ref{Ptr{Cvoid}}),
                    model.ptr, words, length(basis), 0, 0, r))
    end
    le = GPULocalEnergyEvaluator(H, ansatz, model, r[])
    finalizer(x -> ccall((:gwz_sweep_destroy, LIB), Cint, (Ptr{Cvoid},), x.sweep), le)
    return le
end
function val_and_grad(le::GPULocalEnergyEvaluator, params)                                   # localenergy.jl:94-129
    p = Float64[only(params)]
    val = Ref{Cdouble}(0.0)
    grad = zeros(Cdouble, 1)
    check(ccall((:gwz_sweep_val_and_grad, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), le.sweep, p, val, grad, C_NULL))
    return val[], SVector{1,Float64}(grad)
end
val_err_and_grad(le::GPULocalEnergyEvaluator, params) = ((v, g) = val_and_grad(le, params); (v, 0.0, g))  # abstract.jl:59-62
function (le::GPULocalEnergyEvaluator)(params)                                               # localenergy.jl:131-155
    p = Float64[only(params)]
    val = Ref{Cdouble}(0.0)
    check(ccall((:gwz_sweep_val, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}), le.sweep, p, val))
    return val[]
end
(le::GPULocalEnergyEvaluator)(p1::Number, ps::Number...) = le((p1, ps...))                   # localenergy.jl:171-173
function (le::GPULocalEnergyEvaluator)(F, G, params)                                         # localenergy.jl:157-169
    if !isnothing(G)
        val, grad = val_and_grad(le, params)
        G .= grad
    else
        val = le(params)
    end
    return isnothing(F) ? nothing : val
end

# ---------------------------------------------------------------------------------------------
# amsgrad / gradient_descent (amsgrad.jl:105-263).  The loop runs inside the library: natively for the two bit-level
# objectives (no Julia code between iterations), through C callbacks for any other objective that implements
# val_and_grad / val_err_and_grad (any number of parameters).  Gutzwiller.amsgrad itself also works unchanged on all
# GPU objectives, since they implement those two functions.
# ---------------------------------------------------------------------------------------------
struct GdOpts
    maxiter::Int32
    use_amsgrad::Int32
    early_stop::Int32
    reserved::Int32
    alpha::Cdouble
    beta1::Cdouble
    beta2::Cdouble
    grad_tol::Cdouble
    param_tol::Cdouble
    val_tol::Cdouble
    err_tol::Cdouble
    first_moment_init::Ptr{Cdouble}
    second_moment_init::Ptr{Cdouble}
    fix_params::Ptr{UInt8}
end
mutable struct GdTrace
    iterations::Int32
    converged::Int32
    reason::Int32
    np::Int32
    param::Ptr{Cdouble}
    value::Ptr{Cdouble}
    error::Ptr{Cdouble}
    gradient::Ptr{Cdouble}
    first_moment::Ptr{Cdouble}
    second_moment::Ptr{Cdouble}
    param_delta::Ptr{Cdouble}
end

"`GradientDescentResult` (amsgrad.jl:9-25) of a GPU-side descent; same field names."
struct GPUGradientDescentResult{N,F}
    fun::F
    hyperparameters::NamedTuple
    initial_params::SVector{N,Float64}
    iterations::Int
    converged::Bool
    reason::String
    param::Vector{SVector{N,Float64}}
    value::Vector{Float64}
    error::Vector{Float64}
    gradient::Vector{SVector{N,Float64}}
    first_moment::Vector{SVector{N,Float64}}
    second_moment::Vector{SVector{N,Float64}}
    param_delta::Vector{SVector{N,Float64}}
end

# callbacks for gwz_adaptive_gradient_descent: `user` points at an ObjectiveState
mutable struct ObjectiveState
    f::Any
    fkwargs::NamedTuple
    exception::Any
end
function _store!(val, err, grad, np, v, e, g)
    unsafe_store!(val, Float64(v))
    unsafe_store!(err, Float64(e))
    for i in 1:np
        unsafe_store!(grad, Float64(g[i]), i)
    end
    return Cint(0)
end
function _cb_val_and_grad(user::Ptr{Cvoid}, p::Ptr{Cdouble}, np::Cint, val::Ptr{Cdouble}, err::Ptr{Cdouble},
                          grad::Ptr{Cdouble})::Cint
    st = unsafe_pointer_to_objref(user)::ObjectiveState
    try
        v, g = val_and_grad(st.f, copy(unsafe_wrap(Array, p, Int(np))); st.fkwargs...)        # amsgrad.jl:136
        return _store!(val, err, grad, np, v, 0.0, g)
    catch ex
        st.exception = ex
        return Cint(1)
    end
end
function _cb_val_err_and_grad(user::Ptr{Cvoid}, p::Ptr{Cdouble}, np::Cint, val::Ptr{Cdouble}, err::Ptr{Cdouble},
                              grad::Ptr{Cdouble})::Cint
    st = unsafe_pointer_to_objref(user)::ObjectiveState
    try
        v, e, g = val_err_and_grad(st.f, copy(unsafe_wrap(Array, p, Int(np))); st.fkwargs...)  # amsgrad.jl:168
        return _store!(val, err, grad, np, v, e, g)
    catch ex
        st.exception = ex
        return Cint(1)
    end
end

_params_vector(p::GPUGradientDescentResult) = collect(Float64, p.param[end])
_params_vector(p) = collect(Float64, p isa Number ? (p,) : p)

function _descend(f, p_init; use_amsgrad::Bool, maxiter=100, verbose=false, α=0.01, β1=0.1, β2=0.01,
                  first_moment_init=nothing, second_moment_init=nothing, fix_params=nothing, early_stop=true,
                  grad_tol=√eps(Float64), param_tol=√eps(Float64), val_tol=√eps(Float64), err_tol=0.0,
                  fkwargs=(;))
    if p_init isa GPUGradientDescentResult                       # continue a previous run (amsgrad.jl:109-116)
        isnothing(first_moment_init) && (first_moment_init = p_init.first_moment[end])
        isnothing(second_moment_init) && (second_moment_init = p_init.second_moment[end])
    end
    p = _params_vector(p_init)
    N = length(p)
    m1 = isnothing(first_moment_init) ? nothing : collect(Float64, first_moment_init)
    m2 = isnothing(second_moment_init) ? nothing : collect(Float64, second_moment_init)
    fix = isnothing(fix_params) ? nothing : UInt8.(collect(fix_params))
    for v in (m1, m2, fix)
        isnothing(v) || length(v) == N || throw(ArgumentError("moment / fix_params length does not match the number of parameters"))
    end
    rows = maxiter + 1
    vecs = [zeros(Cdouble, N, rows) for _ in 1:5]               # param, gradient, first_moment, second_moment, param_delta
    value, errs = zeros(Cdouble, rows), zeros(Cdouble, rows)
    st = ObjectiveState(f, NamedTuple(fkwargs), nothing)
    GC.@preserve vecs value errs p m1 m2 fix st begin
        opts = GdOpts(maxiter, use_amsgrad, early_stop, 0, α, β1, β2, grad_tol, param_tol, val_tol, err_tol,
                      isnothing(m1) ? C_NULL : pointer(m1), isnothing(m2) ? C_NULL : pointer(m2),
                      isnothing(fix) ? C_NULL : pointer(fix))
        tr = GdTrace(0, 0, 0, N, pointer(vecs[1]), pointer(value), pointer(errs), pointer(vecs[2]),
                     pointer(vecs[3]), pointer(vecs[4]), pointer(vecs[5]))
        if f isa GPUKineticVQMC && N == 1 && isempty(st.fkwargs)
            last = Ref(VqmcResult())
            check(ccall((:gwz_amsgrad_vqmc, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cdouble}, Ref{GdOpts}, UInt64, UInt64, Cint, Cuint, Ref{GdTrace}, Ref{VqmcResult}),
                        f.result.walkers, p, opts, f.steps, f.result.burn_in, GWZ_KERNEL_BITPARALLEL,
                        GWZ_RUN_BLOCKING, tr, last))
            f.result.last = last[]                                # what the reference leaves in qmc.result
            f.result.steps = f.steps
        elseif f isa GPULocalEnergyEvaluator && N == 1 && isempty(st.fkwargs)
            check(ccall((:gwz_amsgrad_sweep, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cdouble}, Ref{GdOpts}, Ref{GdTrace}), f.sweep, p, opts, tr))
        else
            cb1 = @cfunction(_cb_val_and_grad, Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}))
            cb2 = @cfunction(_cb_val_err_and_grad, Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}))
            rc = ccall((:gwz_adaptive_gradient_descent, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ref{GdOpts}, Ref{GdTrace}),
                       cb1, cb2, pointer_from_objref(st), p, N, opts, tr)
            isnothing(st.exception) || throw(st.exception)
            check(rc)
        end
        n = Int(tr.iterations)
        if f isa GPUKineticVQMC && n > 0
            f.result.params = SVector{GWZ_NUM_PARAMS,Float64}(vecs[1][1, n])
        end
        reason = ("iterations", "params", "gradient", "value", "errorbars")[tr.reason + 1]
        sv(m) = [SVector{N,Float64}(m[:, i]) for i in 1:n]
        hyper = use_amsgrad ? (; α, β1, β2) : (; α)
        return GPUGradientDescentResult{N,typeof(f)}(f, hyper, SVector{N,Float64}(p), n, tr.converged != 0, reason,
                                                     sv(vecs[1]), value[1:n], errs[1:n], sv(vecs[2]), sv(vecs[3]),
                                                     sv(vecs[4]), sv(vecs[5]))
    end
end
const GPUObjective = Union{GPUKineticVQMC,GPULocalEnergyEvaluator}
amsgrad(f::GPUObjective, p0; kw...) = _descend(f, p0; use_amsgrad=true, kw...)                # amsgrad.jl:259-263
gradient_descent(f::GPUObjective, p0; kw...) = _descend(f, p0; use_amsgrad=false, kw...)      # amsgrad.jl:240-244

# ------------------------------------------------------------------------------------------
# Ansatz-generic engine (gwz_gen_* / gwz_gwalkers_* / gwz_gsweep_*): the reference's other
# single-component ansaetze and Rimu's ExtendedHubbardReal1D.  P = num_parameters(ansatz).
#   src/ansatz/{extgutzwiller,multinomial,densityprofile,jastrow}.jl
# ------------------------------------------------------------------------------------------
ansatz_kind(::GutzwillerAnsatz) = Cint(0)
ansatz_kind(::ExtendedGutzwillerAnsatz) = Cint(1)
ansatz_kind(::MultinomialAnsatz) = Cint(2)
ansatz_kind(::DensityProfileAnsatz) = Cint(3)
ansatz_kind(::RelativeJastrowAnsatz) = Cint(4)
ansatz_kind(::JastrowAnsatz) = Cint(5)
ansatz_normalize(a::MultinomialAnsatz) = Cint(a.normalization != 1.0)    # multinomial.jl:22-31
ansatz_normalize(_) = Cint(0)

function GPUModel(H::ExtendedHubbardReal1D, ctx::GPUContext=default_context())
    addr = starting_address(H)
    r = Ref{Ptr{Cvoid}}()
    check(ccall((:gwz_model_create_ext_hubbard1d, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble, Cdouble, Ref{Ptr{Cvoid}}),
                ctx.ptr, num_particles(addr), num_modes(addr), H.u, H.v, H.t, r))
    N, M = num_particles(addr), num_modes(addr)
    m = GPUModel(r[], N, M, cld(N + M - 1, 64), ctx)
    finalizer(x -> ccall((:gwz_model_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), m)
    return m
end

mutable struct GPUAnsatz
    ptr::Ptr{Cvoid}
    model::GPUModel
    np::Int
end
function GPUAnsatz(H, ansatz, ctx::GPUContext=default_context())
    model = GPUModel(H, ctx)
    r = Ref{Ptr{Cvoid}}()
    if ansatz isa CombinationAnsatz                                    # a1 + a2 (combination.jl:13-15)
        a1, a2 = ansatz.ansatz1, ansatz.ansatz2
        check(ccall((:gwz_ansatz_create_combination, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                    model.ptr, ansatz_kind(a1), ansatz_normalize(a1), ansatz_kind(a2), ansatz_normalize(a2), r))
    else
        check(ccall((:gwz_ansatz_create, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Ptr{Cvoid}}),
                    model.ptr, ansatz_kind(ansatz), ansatz_normalize(ansatz), r))
    end
    a = GPUAnsatz(r[], model, num_parameters(ansatz))
    finalizer(x -> ccall((:gwz_ansatz_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), a)
    return a
end

struct GenResult                       # gwz_gen_result
    val::Cdouble; err::Cdouble; pooled_val::Cdouble; pooled_var::Cdouble; total_time::Cdouble
    walkers::UInt64; steps::UInt64; hops::UInt64; kernel_ms::Cfloat; np::Int32
end

mutable struct GPUGenericVQMC{H,A}      # KineticVQMC(H, ansatz; samples, walkers, steps) (wrapper.jl:45-63)
    hamiltonian::H
    ansatz::A
    gpu::GPUAnsatz
    walkers_ptr::Ptr{Cvoid}
    walkers::Int
    steps::Int
    last::GenResult
    grad::Vector{Float64}
    grad_err::Vector{Float64}
end
function GPUGenericVQMC(H, ansatz; samples=1e6, walkers=2^16, steps=round(Int, samples / walkers),
                        seed::UInt64=0x9E3779B97F4A7C15, ctx=default_context())
    gpu = GPUAnsatz(H, ansatz, ctx)
    start = address_words(gpu.model, starting_address(H))
    r = Ref{Ptr{Cvoid}}()
    check(ccall((:gwz_gwalkers_create, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, Ptr{UInt64}, UInt64, Ref{Ptr{Cvoid}}),
                gpu.ptr, walkers, 0, start, seed, r))
    np = gpu.np
    q = GPUGenericVQMC(H, ansatz, gpu, r[], Int(walkers), Int(steps),
                       GenResult(0, 0, 0, 0, 0, 0, 0, 0, 0, np), zeros(np), zeros(np))
    finalizer(x -> ccall((:gwz_gwalkers_destroy, LIB), Cint, (Ptr{Cvoid},), x.walkers_ptr), q)
    return q
end
function _run!(q::GPUGenericVQMC, params, steps::Integer; keep::Bool=false, burn_in::Integer=0)
    p = collect(Float64, params)
    length(p) == q.gpu.np || throw(ArgumentError("expected $(q.gpu.np) parameters"))
    res = Ref(q.last)
    check(ccall((:gwz_gen_vqmc_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, UInt64, UInt64, Cuint, Ref{GenResult}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                q.walkers_ptr, p, steps, burn_in, keep ? GWZ_RUN_KEEP_ACC : Cuint(0), res, q.grad, q.grad_err, C_NULL))
    q.last = res[]
    return q
end
(q::GPUGenericVQMC)(params; samples=q.steps * q.walkers, steps=samples ÷ q.walkers) = _run!(q, params, steps).last.val
function val_and_grad(q::GPUGenericVQMC, params; samples=q.steps * q.walkers, steps=samples ÷ q.walkers)
    _run!(q, params, steps)
    return q.last.val, SVector{q.gpu.np,Float64}(q.grad)                                       # wrapper.jl:90-97
end
function val_err_and_grad(q::GPUGenericVQMC, params; samples=q.steps * q.walkers, steps=samples ÷ q.walkers)
    _run!(q, params, steps)
    return q.last.val, q.last.err, SVector{q.gpu.np,Float64}(q.grad)                           # wrapper.jl:98-105
end

mutable struct GPUGenericLocalEnergyEvaluator{H,A}   # LocalEnergyEvaluator(H, ansatz; basis) (localenergy.jl:81-88)
    hamiltonian::H
    ansatz::A
    gpu::GPUAnsatz
    sweep::Ptr{Cvoid}
end
function GPUGenericLocalEnergyEvaluator(H, ansatz; basis=nothing, ctx=default_context())
    gpu = GPUAnsatz(H, ansatz, ctx)
    r = Ref{Ptr{Cvoid}}()
    if basis === nothing
        check(ccall((:gwz_gsweep_create, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, UInt64, UInt64, UInt64, Ref{Ptr{Cvoid}}),
                    gpu.ptr, C_NULL, 0, 0, 0, r))
    else
        words = reduce(vcat, (address_words(gpu.model, a) for a in basis))
        check(ccall((:gwz_gsweep_create, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, UInt64, UInt64, UInt64, Ref{Ptr{Cvoid}}),
                    gpu.ptr, words, length(basis), 0, 0, r))
    end
    le = GPUGenericLocalEnergyEvaluator(H, ansatz, gpu, r[])
    finalizer(x -> ccall((:gwz_gsweep_destroy, LIB), Cint, (Ptr{Cvoid},), x.sweep), le)
    return le
end
function val_and_grad(le::GPUGenericLocalEnergyEvaluator, params)                              # localenergy.jl:94-129
    p = collect(Float64, params); val = Ref(0.0); grad = zeros(le.gpu.np)
    check(ccall((:gwz_gsweep_val_and_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}),
                le.sweep, p, val, grad))
    return val[], SVector{le.gpu.np,Float64}(grad)
end
function (le::GPUGenericLocalEnergyEvaluator)(params)                                          # localenergy.jl:131-155
    p = collect(Float64, params); val = Ref(0.0)
    check(ccall((:gwz_gsweep_val_and_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}),
                le.sweep, p, val, C_NULL))
    return val[]
end
# amsgrad / gradient_descent of the reference (amsgrad.jl:105-263) are generic over objectives with
# val_and_grad / val_err_and_grad, so they drive GPUGenericVQMC and GPUGenericLocalEnergyEvaluator unmodified.

# ------------------------------------------------------------------------------------------
# VectorAnsatz (src/ansatz/vector.jl:1-12): the pairs of the DVec live in an HBM hash table
# ------------------------------------------------------------------------------------------
mutable struct GPUVectorAnsatz
    ptr::Ptr{Cvoid}
    model::GPUModel
end
function GPUVectorAnsatz(H, va::VectorAnsatz, ctx::GPUContext=default_context())
    model = GPUModel(H, ctx)
    ks = collect(keys(va.vector))
    words = reduce(vcat, (address_words(model, k) for k in ks))
    vals = Float64[va.vector[k] for k in ks]
    r = Ref{Ptr{Cvoid}}()
    check(ccall((:gwz_vector_ansatz_create, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Cdouble}, UInt64, Ref{Ptr{Cvoid}}),
                model.ptr, words, vals, length(ks), r))
    a = GPUVectorAnsatz(r[], model)
    finalizer(x -> ccall((:gwz_vector_ansatz_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), a)
    return a
end
# LocalEnergyEvaluator(H, VectorAnsatz(v))() over basis = keys(v): the Rayleigh quotient
function rayleigh(a::GPUVectorAnsatz)
    val = Ref(0.0)
    check(ccall((:gwz_vector_ansatz_rayleigh, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), a.ptr, val))
    return val[]
end
# kinetic_vqmc(H, VectorAnsatz(v), Float64[]; walkers, steps) (test/qmc.jl:22-23): returns the gwz_gen_result
function kinetic_vqmc(H, va::VectorAnsatz, ::AbstractVector; walkers=2^12, steps=1000, seed::UInt64=0x9E3779B97F4A7C15,
                      ctx=default_context())
    a = GPUVectorAnsatz(H, va, ctx)
    start = address_words(a.model, starting_address(H))
    w = Ref{Ptr{Cvoid}}()
    check(ccall((:gwz_vwalkers_create, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, Ptr{UInt64}, UInt64, Ref{Ptr{Cvoid}}),
                a.ptr, walkers, 0, start, seed, w))
    res = Ref(GenResult(0, 0, 0, 0, 0, 0, 0, 0, 0, 0))
    check(ccall((:gwz_vwalkers_run, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, Cuint, Ref{GenResult}), w[], steps, 0, Cuint(0), res))
    ccall((:gwz_vwalkers_destroy, LIB), Cint, (Ptr{Cvoid},), w[])
    return res[]
end

# ------------------------------------------------------------------------------------------
# Stored operators: ANY Rimu Hamiltonian and ANY ansatz (HubbardMom1D, HubbardRealSpace, ... of test/qmc.jl:10-14)
# The operator is materialised over build_basis(H): row k = offdiagonals(H, basis[k]) in Rimu's order, and the
# ansatz is tabulated over the basis per parameter set (GutzwillerAnsatz is filled on the device).
# ------------------------------------------------------------------------------------------
mutable struct GPUSparseHamiltonian{H,A}
    hamiltonian::H
    basis::Vector{A}
    ptr::Ptr{Cvoid}
    start::UInt32
end
function GPUSparseHamiltonian(H; basis=build_basis(H), ctx::GPUContext=default_context())
    index = Dict(a => UInt32(i - 1) for (i, a) in enumerate(basis))
    row_ptr = UInt64[0]; col = UInt32[]; melem = Float64[]; diag = Float64[]
    for k in basis
        push!(diag, diagonal_element(H, k))
        for (l, m) in offdiagonals(H, k)                        # reference order (qmc.jl:15,25)
            push!(col, index[l]); push!(melem, m)
        end
        push!(row_ptr, length(col))
    end
    r = Ref{Ptr{Cvoid}}()
    check(ccall((:gwz_sparse_create, LIB), Cint,
                (Ptr{Cvoid}, UInt64, Ptr{UInt64}, Ptr{UInt32}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                ctx.ptr, length(basis), row_ptr, col, melem, diag, r))
    s = GPUSparseHamiltonian(H, collect(basis), r[], index[starting_address(H)])
    finalizer(x -> ccall((:gwz_sparse_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), s)
    return s
end
# tabulate ansatz(., params) over the basis: psi and grad psi / psi (qmc.jl:18,38)
function load!(s::GPUSparseHamiltonian, ansatz, params)
    if ansatz isa GutzwillerAnsatz
        check(ccall((:gwz_sparse_set_gutzwiller, LIB), Cint, (Ptr{Cvoid}, Cdouble), s.ptr, Float64(first(params))))
        return s
    end
    np = num_parameters(ansatz); dim = length(s.basis)
    psi = Vector{Float64}(undef, dim); gr = Matrix{Float64}(undef, np, dim)   # column k = grad_ratio[k][:]
    p = SVector{np,Float64}(params)
    for (i, k) in enumerate(s.basis)
        v, g = val_and_grad(ansatz, k, p)
        psi[i] = v
        gr[:, i] .= g ./ v
    end
    check(ccall((:gwz_sparse_set_table, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int32),
                s.ptr, psi, np == 0 ? C_NULL : gr, np))
    return s
end
mutable struct GPUSparseVQMC{S,A}         # KineticVQMC(H, ansatz; samples, walkers, steps) (wrapper.jl:45-63)
    op::S
    ansatz::A
    walkers_ptr::Ptr{Cvoid}
    walkers::Int
    steps::Int
    last::GenResult
    grad::Vector{Float64}
    grad_err::Vector{Float64}
end
function GPUSparseVQMC(op::GPUSparseHamiltonian, ansatz; samples=1e6, walkers=2^16,
                       steps=round(Int, samples / walkers), seed::UInt64=0x9E3779B97F4A7C15)
    r = Ref{Ptr{Cvoid}}()
    check(ccall((:gwz_swalkers_create, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, UInt32, UInt64, Ref{Ptr{Cvoid}}),
                op.ptr, walkers, 0, op.start, seed, r))
    np = num_parameters(ansatz)
    q = GPUSparseVQMC(op, ansatz, r[], Int(walkers), Int(steps), GenResult(0, 0, 0, 0, 0, 0, 0, 0, 0, np),
                      zeros(max(np, 1)), zeros(max(np, 1)))
    finalizer(x -> ccall((:gwz_swalkers_destroy, LIB), Cint, (Ptr{Cvoid},), x.walkers_ptr), q)
    return q
end
function _run!(q::GPUSparseVQMC, params, steps::Integer; keep::Bool=false, burn_in::Integer=0)
    load!(q.op, q.ansatz, params)
    res = Ref(q.last)
    check(ccall((:gwz_swalkers_run, LIB), Cint,
                (Ptr{Cvoid}, UInt64, UInt64, Cuint, Ref{GenResult}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                q.walkers_ptr, steps, burn_in, keep ? GWZ_RUN_KEEP_ACC : Cuint(0), res, q.grad, q.grad_err, C_NULL))
    q.last = res[]
    return q
end
(q::GPUSparseVQMC)(params; samples=q.steps * q.walkers, steps=samples ÷ q.walkers) = _run!(q, params, steps).last.val
function val_and_grad(q::GPUSparseVQMC, params; samples=q.steps * q.walkers, steps=samples ÷ q.walkers)
    _run!(q, params, steps)
    np = num_parameters(q.ansatz)
    return q.last.val, SVector{np,Float64}(q.grad[1:np])                                       # wrapper.jl:90-97
end
function val_err_and_grad(q::GPUSparseVQMC, params; samples=q.steps * q.walkers, steps=samples ÷ q.walkers)
    _run!(q, params, steps)
    np = num_parameters(q.ansatz)
    return q.last.val, q.last.err, SVector{np,Float64}(q.grad[1:np])                           # wrapper.jl:98-105
end
# DVec(res) (state.jl:303-319): residence time per address, accumulated on the device
function sampled_vector(q::GPUSparseVQMC, params; steps=q.steps, burn_in=0)
    load!(q.op, q.ansatz, params)
    hist = zeros(length(q.op.basis)); res = Ref(q.last)
    check(ccall((:gwz_swalkers_sample_vector, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, Cuint, Ptr{Cdouble}, Ref{GenResult}),
                q.walkers_ptr, steps, burn_in, Cuint(0), hist, res))
    dv = DVec(zip(q.op.basis, hist)...)
    return normalize!(dv)
end
# LocalEnergyEvaluator(H, ansatz) over the stored rows (localenergy.jl:81-173)
struct GPUSparseLocalEnergyEvaluator{S,A}
    op::S
    ansatz::A
end
function val_and_grad(le::GPUSparseLocalEnergyEvaluator, params)                               # localenergy.jl:94-129
    load!(le.op, le.ansatz, params)
    np = num_parameters(le.ansatz); val = Ref(0.0); grad = zeros(max(np, 1))
    check(ccall((:gwz_sparse_sweep, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, Ref{Cdouble}, Ptr{Cdouble}),
                le.op.ptr, 0, length(le.op.basis), val, np == 0 ? C_NULL : grad))
    return val[], SVector{np,Float64}(grad[1:np])
end
function (le::GPUSparseLocalEnergyEvaluator)(params)                                           # localenergy.jl:131-155
    load!(le.op, le.ansatz, params)
    val = Ref(0.0)
    check(ccall((:gwz_sparse_sweep, LIB), Cint, (Ptr{Cvoid}, UInt64, UInt64, Ref{Cdouble}, Ptr{Cdouble}),
                le.op.ptr, 0, length(le.op.basis), val, C_NULL))
    return val[]
end

end # module
