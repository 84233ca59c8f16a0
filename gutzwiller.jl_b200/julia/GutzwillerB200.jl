# GutzwillerB200.jl -- Julia host shim over libgutzwiller_b200.so (C ABI: include/gutzwiller_b200.h).
#
# Keeps the call forms of mtsch/Gutzwiller.jl for the HubbardReal1D + GutzwillerAnsatz path:
#   kinetic_vqmc / kinetic_vqmc! / KineticVQMC / LocalEnergyEvaluator / val_and_grad /
#   val_err_and_grad / amsgrad / gradient_descent
# (reference: src/kinetic-vqmc/qmc.jl:129-146, wrapper.jl:45-121, localenergy.jl:81-173,
#  amsgrad.jl:105-263).  Every method is a thin `ccall`; no arithmetic happens in Julia.
#
# NOTE: Julia is not installed in the build image, so this file has not been executed there.  The
# Python mirror (gutzwiller.jl_b200/*.py) binds exactly the same symbols and is what the test-suite
# drives; see INTEGRATION.md for how a maintainer wires this module into Gutzwiller.jl.
module GutzwillerB200

using Rimu
using StaticArrays

export GPUContext, GPUKineticVQMCResult, GPUKineticVQMC, GPULocalEnergyEvaluator
export GPUAnsatz, GPUGenericVQMC, GPUGenericLocalEnergyEvaluator
export GPUSparseHamiltonian, GPUSparseVQMC, GPUSparseLocalEnergyEvaluator, sampled_vector
export kinetic_vqmc, kinetic_vqmc!, val_and_grad, val_err_and_grad, amsgrad, gradient_descent

const LIB = get(ENV, "GUTZWILLER_B200_LIB",
                joinpath(@__DIR__, "..", "lib", "libgutzwiller_b200.so"))

const GWZ_KERNEL_ENUMERATE = Cint(0)
const GWZ_KERNEL_BITPARALLEL = Cint(1)
const GWZ_RUN_KEEP_ACC = Cuint(1)
const GWZ_NUM_ACC = 12

# mirrors `struct gwz_vqmc_result` (P = 1)
struct VqmcResult
    val::Cdouble
    err::Cdouble
    grad::NTuple{1,Cdouble}
    grad_err::NTuple{1,Cdouble}
    pooled_val::Cdouble
    pooled_grad::NTuple{1,Cdouble}
    pooled_var::Cdouble
    total_time::Cdouble
    walkers::UInt64
    steps::UInt64
    hops::UInt64
    acc::NTuple{GWZ_NUM_ACC,Cdouble}
    kernel_ms::Cfloat
end

function check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:gwz_last_error, LIB), Cstring, ()))
    rc == 1 && throw(ArgumentError(msg))           # GWZ_ERR_INVALID
    rc == 4 && error(msg)                          # "Infinite residence time" (qmc.jl:34-36)
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
default_context() = something(DEFAULT_CONTEXT[], (DEFAULT_CONTEXT[] = GPUContext(0)))

"Join the contexts of several devices of this process in one NCCL communicator."
function comm_init_all(ctxs::Vector{GPUContext})
    ptrs = [c.ptr for c in ctxs]
    check(ccall((:gwz_context_comm_init_all, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), ptrs, length(ptrs)))
end

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
# kinetic_vqmc / KineticVQMCResult  (qmc.jl:73-146, state.jl:137-164,287-292)
# ---------------------------------------------------------------------------------------------
mutable struct GPUKineticVQMCResult{H,A}
    hamiltonian::H
    ansatz::A
    model::GPUModel
    walkers::Ptr{Cvoid}
    n_walkers::Int
    steps::Int          # per walker, accumulated in the estimators
    last::VqmcResult
end
Base.length(res::GPUKineticVQMCResult) = res.steps * res.n_walkers

function _make_walkers(model::GPUModel, H, walkers::Integer; seed::UInt64=0x9E3779B97F4A7C15, offset::Integer=0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    start = address_words(model, starting_address(H))      # state.jl:26
    check(ccall((:gwz_walkers_create, LIB), Cint,
                (Ptr{Cvoid}, UInt64, UInt64, Ptr{UInt64}, UInt64, Ref{Ptr{Cvoid}}),
                model.ptr, walkers, offset, start, seed, r))
    return r[]
end

function _run!(res::GPUKineticVQMCResult, params, steps::Integer; keep::Bool, burn_in::Integer=0,
               kernel=GWZ_KERNEL_BITPARALLEL)
    p = Float64[only(params)]
    out = Ref{VqmcResult}()
    check(ccall((:gwz_vqmc_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, UInt64, UInt64, Cint, Cuint, Ref{VqmcResult}),
                res.walkers, p, steps, burn_in, kernel, keep ? GWZ_RUN_KEEP_ACC : Cuint(0), out))
    res.last = out[]
    res.steps = keep ? res.steps + steps : steps
    return res
end

"""
    kinetic_vqmc(H::HubbardReal1D, ansatz::GutzwillerAnsatz, params; samples=1e6, walkers, steps)

Drop-in for `Gutzwiller.kinetic_vqmc` (qmc.jl:129-146) running on the GPU.  `walkers` defaults to the
largest power of two leaving >= 1024 steps per walker (the reference default is thread-count derived,
qmc.jl:132); an explicit `walkers`/`steps`/`samples` is honoured exactly.
"""
function kinetic_vqmc(H::HubbardReal1D, ansatz, params;
                      samples=1e6,
                      walkers=min(2^20, prevpow(2, max(1, Int(samples) ÷ 1024))),
                      steps=round(Int, samples / walkers),
                      burn_in=0, seed=0x9E3779B97F4A7C15, ctx=default_context())
    model = GPUModel(H, ctx)
    w = _make_walkers(model, H, walkers; seed)
    res = GPUKineticVQMCResult(H, ansatz, model, w, Int(walkers), 0, VqmcResult(ntuple(_ -> 0, 13)...))
    finalizer(r -> ccall((:gwz_walkers_destroy, LIB), Cint, (Ptr{Cvoid},), r.walkers), res)
    return _run!(res, params, steps; keep=false, burn_in)
end

"`kinetic_vqmc!(res; steps)` (qmc.jl:108-111): continue from the walkers' current addresses."
kinetic_vqmc!(res::GPUKineticVQMCResult, params; steps=res.steps) = _run!(res, params, steps; keep=true)

val_and_grad(res::GPUKineticVQMCResult) = (res.last.val, SVector(res.last.grad))        # state.jl:137-149
val_err_and_grad(res::GPUKineticVQMCResult) = (res.last.val, res.last.err, SVector(res.last.grad))
Statistics_mean(res::GPUKineticVQMCResult) = res.last.val                                 # state.jl:287-292

# ---------------------------------------------------------------------------------------------
# KineticVQMC (wrapper.jl:45-121)
# ---------------------------------------------------------------------------------------------
struct GPUKineticVQMC{R}
    steps::Int
    walkers::Int
    result::R
end
function GPUKineticVQMC(H::HubbardReal1D, ansatz; samples=1e6,
                        walkers=min(2^20, prevpow(2, max(1, Int(samples) ÷ 1024))),
                        steps=round(Int, samples / walkers), kwargs...)
    res = kinetic_vqmc(H, ansatz, 0.0; walkers, steps=1, kwargs...)   # allocate walkers at starting_address
    return GPUKineticVQMC(Int(steps), Int(walkers), res)
end
function (ke::GPUKineticVQMC)(params; samples=ke.steps * ke.walkers, steps=samples ÷ ke.walkers)
    return _run!(ke.result, params, steps; keep=false).last.val                              # wrapper.jl:81-88
end
function val_and_grad(ke::GPUKineticVQMC, params; samples=ke.steps * ke.walkers, steps=samples ÷ ke.walkers)
    return val_and_grad(_run!(ke.result, params, steps; keep=false))                         # wrapper.jl:90-97
end
function val_err_and_grad(ke::GPUKineticVQMC, params; samples=ke.steps * ke.walkers, steps=samples ÷ ke.walkers)
    return val_err_and_grad(_run!(ke.result, params, steps; keep=false))                     # wrapper.jl:98-105
end
function (ke::GPUKineticVQMC)(F, G, params)                                                  # wrapper.jl:107-121
    res = _run!(ke.result, params, ke.steps; keep=false)
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
function (le::GPULocalEnergyEvaluator)(params)                                               # localenergy.jl:131-155
    p = Float64[only(params)]
    val = Ref{Cdouble}(0.0)
    check(ccall((:gwz_sweep_val, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}), le.sweep, p, val))
    return val[]
end
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
# amsgrad / gradient_descent on the two GPU objectives (amsgrad.jl:105-263): the loop runs inside
# the library; the trace comes back as plain vectors.  (Gutzwiller.amsgrad itself also works
# unchanged on these objects, since they implement val_and_grad / val_err_and_grad.)
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

function _descend(f, p0; use_amsgrad, maxiter=100, α=0.01, β1=0.1, β2=0.01, early_stop=true,
                  grad_tol=√eps(Float64), param_tol=√eps(Float64), val_tol=√eps(Float64), err_tol=0.0)
    rows = maxiter + 1
    bufs = [zeros(Cdouble, rows) for _ in 1:7]
    opts = GdOpts(maxiter, use_amsgrad, early_stop, 0, α, β1, β2, grad_tol, param_tol, val_tol, err_tol,
                  C_NULL, C_NULL, C_NULL)
    GC.@preserve bufs begin
        tr = GdTrace(0, 0, 0, 1, (pointer(b) for b in bufs)...)
        p = Float64[only(p0)]
        if f isa GPUKineticVQMC
            check(ccall((:gwz_amsgrad_vqmc, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cdouble}, Ref{GdOpts}, UInt64, Cint, Cuint, Ref{GdTrace}),
                        f.result.walkers, p, opts, f.steps, GWZ_KERNEL_BITPARALLEL, Cuint(0), tr))
        else
            check(ccall((:gwz_amsgrad_sweep, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cdouble}, Ref{GdOpts}, Ref{GdTrace}), f.sweep, p, opts, tr))
        end
        n = tr.iterations
        reason = ("iterations", "params", "gradient", "value", "errorbars")[tr.reason + 1]
        return (; iterations=Int(n), converged=tr.converged != 0, reason,
                param=bufs[1][1:n], value=bufs[2][1:n], error=bufs[3][1:n], gradient=bufs[4][1:n],
                first_moment=bufs[5][1:n], second_moment=bufs[6][1:n], param_delta=bufs[7][1:n])
    end
end
amsgrad(f::Union{GPUKineticVQMC,GPULocalEnergyEvaluator}, p0; kw...) = _descend(f, p0; use_amsgrad=1, kw...)
gradient_descent(f::Union{GPUKineticVQMC,GPULocalEnergyEvaluator}, p0; kw...) = _descend(f, p0; use_amsgrad=0, kw...)

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
    m = GPUModel(r[], ctx, num_particles(addr), num_modes(addr), cld(num_particles(addr) + num_modes(addr) - 1, 64))
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
