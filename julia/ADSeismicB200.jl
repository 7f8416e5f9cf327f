# ADSeismicB200.jl -- Julia binding of libadseis_b200.so (the C ABI in include/adseis.h).
#
# NOT EXECUTED IN THE BUILD IMAGE (no `julia` there): this file is the reference-side stub a maintainer adds; the
# same ABI is exercised end-to-end by the Python/ctypes mirror in adseismic.jl_b200/ and its tests.
#
# It keeps the public names of ADSeismic.jl's hot path (src/Struct.jl, src/Core.jl, src/Utils.jl) but evaluates
# eagerly on the GPU instead of building a TensorFlow graph:
#     AcousticPropagatorParams / AcousticSource / AcousticReceiver            (src/Struct.jl:82-133)
#     AcousticPropagatorSolver(param, src, c)  -> AcousticPropagator          (src/Core.jl:562)
#     SimulatedObservation!(ap, rcv)                                          (src/Core.jl:726)
#     ElasticPropagatorParams / ElasticSource / ElasticReceiver               (src/Struct.jl:4-62)
#     ElasticPropagatorSolver(param, src, ρ, λ, μ)                            (src/Core.jl:31)
#     AcousticPlan / set_points! / gradient!, compute_loss_and_grads_GPU      (src/Utils.jl:300-332, shot k on device k % n)
#     LBFGS!(loss_and_grad, x0)                                               (src/Optim.jl:135-193)
module ADSeismicB200

using Parameters
export AcousticPropagatorParams, AcousticSource, AcousticReceiver, AcousticPropagator,
       AcousticPropagatorSolver, SimulatedObservation!, acoustic_misfit_grad,
       ElasticPropagatorParams, ElasticSource, ElasticReceiver, ElasticPropagator,
       ElasticPropagatorSolver, elastic_misfit_grad,
       AcousticPlan, set_points!, set_model!, set_srcv!, set_obs!, gradient!, compute_loss_and_grads_GPU, LBFGS!

const libadseis = get(ENV, "ADSEIS_B200_LIB", joinpath(@__DIR__, "..", "adseismic.jl_b200", "libadseis_b200.so"))

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:adseis_last_error, libadseis), Cstring, ()))
    error("adseis error $rc: $msg")
end

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = -1)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:adseis_ctx_create, libadseis), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        c = new(h[])
        finalizer(x -> ccall((:adseis_ctx_destroy, libadseis), Cint, (Ptr{Cvoid},), x.handle), c)
        c
    end
end
const default_ctx = Ref{Union{Nothing,Context}}(nothing)
ctx() = (default_ctx[] === nothing && (default_ctx[] = Context()); default_ctx[])

# ---- C structs (field order = include/adseis.h) ------------------------------------------------------------------
struct CAcousticParams
    NX::Int64; NY::Int64; NSTEP::Int64
    DELTAX::Float64; DELTAY::Float64; DELTAT::Float64
    USE_PML_XMIN::Int32; USE_PML_XMAX::Int32; USE_PML_YMIN::Int32; USE_PML_YMAX::Int32
    NPOINTS_PML::Int64
    Rcoef::Float64; vp_ref::Float64
    mpi_convention::Int32; PropagatorKernel::Int32
end
struct CElasticParams
    NX::Int64; NY::Int64; NSTEP::Int64
    DELTAX::Float64; DELTAY::Float64; DELTAT::Float64
    f0::Float64; vp_ref::Float64
    USE_PML_XMIN::Int32; USE_PML_XMAX::Int32; USE_PML_YMIN::Int32; USE_PML_YMAX::Int32
    NPOINTS_PML::Int64
    NPOWER::Float64; K_MAX_PML::Float64; ALPHA_MAX_PML::Float64; Rcoef::Float64
    variant::Int32; reserved::Int32
end

# ---- public structs: same fields and defaults as src/Struct.jl ---------------------------------------------------
@with_kw mutable struct AcousticPropagatorParams
    NX::Int64 = 101; NY::Int64 = 641; NSTEP::Int64 = 4000
    DELTAX::Float64 = 10.; DELTAY::Float64 = 10.; DELTAT::Float64 = 1e-3
    USE_PML_XMIN::Bool = true; USE_PML_XMAX::Bool = true; USE_PML_YMIN::Bool = true; USE_PML_YMAX::Bool = true
    NPOINTS_PML::Int64 = 12; NPOWER::Int64 = 2
    Rcoef::Float64 = 0.001; vp_ref::Float64 = 1000.
    IT_DISPLAY::Int64 = 0
    PropagatorKernel::Int64 = 0      # as src/Struct.jl:120 (0: TF-op scheme, 1|2: custom-op scheme; slabs need 1)
    mpi_convention::Bool = false     # true: MPIAcousticPropagatorParams inputs
end
toC(p::AcousticPropagatorParams) = CAcousticParams(p.NX, p.NY, p.NSTEP, p.DELTAX, p.DELTAY, p.DELTAT,
    p.USE_PML_XMIN, p.USE_PML_XMAX, p.USE_PML_YMIN, p.USE_PML_YMAX, p.NPOINTS_PML, p.Rcoef, p.vp_ref,
    p.mpi_convention, p.PropagatorKernel)

mutable struct AcousticSource
    srci::Vector{Int64}; srcj::Vector{Int64}; srcv::Matrix{Float64}      # srcv: NSTEP(+1) x nsrc
end
mutable struct AcousticReceiver
    rcvi::Vector{Int64}; rcvj::Vector{Int64}; rcvv::Union{Matrix{Float64},Missing}
end
AcousticReceiver(rcvi, rcvj) = AcousticReceiver(rcvi, rcvj, missing)
mutable struct AcousticPropagator
    param::AcousticPropagatorParams; src::AcousticSource; c::Matrix{Float64}
end

# Julia arrays are column-major; the C ABI is row-major (i slow, j fast), as the reference's `Σx'[:]` / tf.reshape
# flattening (src/Core.jl:564-568).  permutedims converts.
rowmajor(A::AbstractMatrix) = collect(permutedims(A))          # (n1,n2) -> memory order i*n2+j
fromrowmajor(v::Vector{Float64}, n1, n2) = collect(permutedims(reshape(v, n2, n1)))

"AcousticPropagatorSolver(param, src, c): src/Core.jl:562-620 (c is the velocity on the padded (NX+2)x(NY+2) grid)"
AcousticPropagatorSolver(param::AcousticPropagatorParams, src::AcousticSource, c::Matrix{Float64}) =
    AcousticPropagator(param, src, c)

"SimulatedObservation!(ap, rcv): src/Core.jl:726-730; rcv.rcvv is (NSTEP+1) x nrcv"
function SimulatedObservation!(ap::AcousticPropagator, rcv::AcousticReceiver)
    p = ap.param; nsrc = length(ap.src.srci); nrcv = length(rcv.rcvi)
    srcv = rowmajor(ap.src.srcv); c = rowmajor(ap.c)
    out = Vector{Float64}(undef, (p.NSTEP + 1) * nrcv)
    check(ccall((:adseis_acoustic_forward, libadseis), Cint,
        (Ptr{Cvoid}, Ref{CAcousticParams}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Int64,
         Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
        ctx().handle, toC(p), c, nsrc, ap.src.srci, ap.src.srcj, srcv, size(ap.src.srcv, 1), nrcv, rcv.rcvi, rcv.rcvj,
        out, C_NULL))
    rcv.rcvv = fromrowmajor(out, p.NSTEP + 1, nrcv)
end

"loss = sum((rcvv - obs).^2) and its gradients (src/Utils.jl:308 + tf.gradients): returns (loss, grad_c, grad_srcv)"
function acoustic_misfit_grad(param::AcousticPropagatorParams, src::AcousticSource, c::Matrix{Float64},
                              rcv::AcousticReceiver, obs::Matrix{Float64})
    nsrc = length(src.srci); nrcv = length(rcv.rcvi)
    loss = Ref{Float64}(0.0)
    rcvv = Vector{Float64}(undef, (param.NSTEP + 1) * nrcv)
    gc = Vector{Float64}(undef, length(c)); gs = Vector{Float64}(undef, param.NSTEP * nsrc)
    check(ccall((:adseis_acoustic_misfit_grad, libadseis), Cint,
        (Ptr{Cvoid}, Ref{CAcousticParams}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Int64,
         Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx().handle, toC(param), rowmajor(c), nsrc, src.srci, src.srcj, rowmajor(src.srcv), size(src.srcv, 1), nrcv,
        rcv.rcvi, rcv.rcvj, rowmajor(obs), loss, rcvv, gc, gs))
    rcv.rcvv = fromrowmajor(rcvv, param.NSTEP + 1, nrcv)
    loss[], fromrowmajor(gc, size(c, 1), size(c, 2)), fromrowmajor(gs, param.NSTEP, nsrc)
end

# ---- device-resident plan: one per GPU, re-pointed per shot (include/adseis.h: adseis_acoustic_plan_*) ----------------
# The reference re-runs one Session graph per L-BFGS iteration and places shot k on device k % n_gpu
# (src/Optim.jl:135-193, src/Utils.jl:300-332).  Here the device state (history window, checkpoints, adjoint planes)
# lives in a plan that is created once per device; set_points! moves it to the next shot.
mutable struct AcousticPlan
    handle::Ptr{Cvoid}
    param::AcousticPropagatorParams
    nsrc::Int; nrcv::Int; shape::Tuple{Int,Int}
    function AcousticPlan(param::AcousticPropagatorParams, src::AcousticSource, rcv::AcousticReceiver; context = ctx(),
                          hist_bytes_budget::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:adseis_acoustic_plan_create, libadseis), Cint,
            (Ptr{Cvoid}, Ref{CAcousticParams}, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64},
             Csize_t, Ref{Ptr{Cvoid}}),
            context.handle, toC(param), C_NULL, length(src.srci), src.srci, src.srcj, length(rcv.rcvi), rcv.rcvi, rcv.rcvj,
            hist_bytes_budget, h))
        shape = param.mpi_convention ? (param.NX, param.NY) : (param.NX + 2, param.NY + 2)
        pl = new(h[], param, length(src.srci), length(rcv.rcvi), shape)
        finalizer(x -> ccall((:adseis_acoustic_plan_destroy, libadseis), Cint, (Ptr{Cvoid},), x.handle), pl)
        pl
    end
end
function set_points!(pl::AcousticPlan, src::AcousticSource, rcv::AcousticReceiver)
    check(ccall((:adseis_acoustic_plan_set_points, libadseis), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}),
        pl.handle, length(src.srci), src.srci, src.srcj, length(rcv.rcvi), rcv.rcvi, rcv.rcvj))
    pl.nsrc = length(src.srci); pl.nrcv = length(rcv.rcvi); pl
end
set_model!(pl::AcousticPlan, c::Matrix{Float64}) =
    check(ccall((:adseis_acoustic_plan_set_model, libadseis), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), pl.handle, rowmajor(c), 0))
set_srcv!(pl::AcousticPlan, srcv::Matrix{Float64}) =
    check(ccall((:adseis_acoustic_plan_set_srcv, libadseis), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint),
                pl.handle, rowmajor(srcv), size(srcv, 1), 0))
set_obs!(pl::AcousticPlan, obs::Matrix{Float64}) =
    check(ccall((:adseis_acoustic_plan_set_obs, libadseis), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), pl.handle, rowmajor(obs), 0))
"forward + reverse sweep; returns (loss, grad_c) -- ADSEIS_GET_LOSS = 2, ADSEIS_GET_GRAD_C = 3 (include/adseis.h:125-129)"
function gradient!(pl::AcousticPlan)
    check(ccall((:adseis_acoustic_plan_gradient, libadseis), Cint, (Ptr{Cvoid},), pl.handle))
    loss = Vector{Float64}(undef, 1); gc = Vector{Float64}(undef, prod(pl.shape))
    check(ccall((:adseis_acoustic_plan_get, libadseis), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), pl.handle, 2, loss, 0))
    check(ccall((:adseis_acoustic_plan_get, libadseis), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint), pl.handle, 3, gc, 0))
    loss[1], fromrowmajor(gc, pl.shape[1], pl.shape[2])
end

"""
compute_loss_and_grads_GPU(param, srcs, rcvs, Rs, c; contexts): src/Utils.jl:300-332 -- sum over shots of
sum((rcvv - Rs[k]).^2) and its gradient w.r.t. c.  Shot k (1-based) runs on contexts[k % n + 1 ...] exactly as the
reference deals shots to `/gpu:(k % n_gpu)`; one plan per device is created on first use and re-pointed per shot.
(One Julia task per device; the C calls of different contexts do not share state.)
"""
function compute_loss_and_grads_GPU(param::AcousticPropagatorParams, srcs::Vector{AcousticSource},
                                    rcvs::Vector{AcousticReceiver}, Rs::Vector{Matrix{Float64}}, c::Matrix{Float64};
                                    contexts::Vector{Context} = [ctx()], plans = Dict{Int,AcousticPlan}())
    n = length(contexts)
    partial = Vector{Tuple{Float64,Matrix{Float64}}}(undef, n)
    @sync for d in 1:n
        Threads.@spawn begin
            L = 0.0; G = zeros(size(c))
            for k in 1:length(srcs)
                k % n == d - 1 || continue                       # src/Utils.jl:326
                pl = get!(plans, d) do
                    AcousticPlan(param, srcs[k], rcvs[k]; context = contexts[d])
                end
                set_points!(pl, srcs[k], rcvs[k]); set_model!(pl, c); set_srcv!(pl, srcs[k].srcv); set_obs!(pl, Rs[k])
                l, g = gradient!(pl)
                L += l; G .+= g
            end
            partial[d] = (L, G)
        end
    end
    sum(first, partial), sum(last, partial)
end

"""
LBFGS!(loss_and_grad, x0; max_iter): the role of src/Optim.jl:135-193 (`LBFGS!(sess, loss, grads, vars)` around
Optim.jl's L-BFGS) without a TensorFlow session: `loss_and_grad(x) -> (loss, grad)` is evaluated eagerly.
"""
function LBFGS!(loss_and_grad::Function, x0::Array{Float64}; max_iter::Int = 15000, callback = nothing)
    Optim = Base.require(Base.PkgId(Base.UUID("429524aa-4258-5aef-a3af-852621145aeb"), "Optim"))
    losses = Float64[]
    fg!(F, G, x) = begin
        l, g = loss_and_grad(x)
        G === nothing || (G .= g)
        push!(losses, l)
        l
    end
    Base.invokelatest(Optim.optimize, Optim.only_fg!(fg!), x0, Optim.LBFGS(),
                      Optim.Options(iterations = max_iter, callback = callback === nothing ? (_ -> false) : callback))
    losses
end

# ---- elastic -----------------------------------------------------------------------------------------------------
@with_kw mutable struct ElasticPropagatorParams
    NX::Int64 = 101; NY::Int64 = 641; NSTEP::Int64 = 4000
    DELTAX::Float64 = 10.; DELTAY::Float64 = 10.; DELTAT::Float64 = 1e-3
    f0::Float64 = 5.; vp_ref::Float64 = 2000.
    USE_PML_XMIN::Bool = true; USE_PML_XMAX::Bool = true; USE_PML_YMIN::Bool = true; USE_PML_YMAX::Bool = true
    NPOINTS_PML::Int64 = 12; NPOWER::Float64 = 2.; K_MAX_PML::Float64 = 1.
    ALPHA_MAX_PML::Float64 = 2. * π * (f0 / 2.); Rcoef::Float64 = 0.001
    IT_DISPLAY::Int64 = 0
    variant::Int64 = 0               # 0: ElasticPropagatorSolver, 1: MPIElasticPropagatorSolver numerics
end
toC(p::ElasticPropagatorParams) = CElasticParams(p.NX, p.NY, p.NSTEP, p.DELTAX, p.DELTAY, p.DELTAT, p.f0, p.vp_ref,
    p.USE_PML_XMIN, p.USE_PML_XMAX, p.USE_PML_YMIN, p.USE_PML_YMAX, p.NPOINTS_PML, p.NPOWER, p.K_MAX_PML,
    p.ALPHA_MAX_PML, p.Rcoef, p.variant, 0)
mutable struct ElasticSource
    srci::Vector{Int64}; srcj::Vector{Int64}; srctype::Vector{Int64}; srcv::Matrix{Float64}
end
mutable struct ElasticReceiver
    rcvi::Vector{Int64}; rcvj::Vector{Int64}; rcvtype::Vector{Int64}; rcvv::Union{Matrix{Float64},Missing}
end
ElasticReceiver(rcvi, rcvj, rcvtype) = ElasticReceiver(rcvi, rcvj, rcvtype, missing)
mutable struct ElasticPropagator
    param::ElasticPropagatorParams; src::ElasticSource; ρ::Matrix{Float64}; λ::Matrix{Float64}; μ::Matrix{Float64}
end
ElasticPropagatorSolver(param::ElasticPropagatorParams, src::ElasticSource, ρ, λ, μ) =
    ElasticPropagator(param, src, ρ, λ, μ)

"SimulatedObservation!(ep, rcv): src/Core.jl:701-712; rcv.rcvv is nrcv x (NSTEP+1)"
function SimulatedObservation!(ep::ElasticPropagator, rcv::ElasticReceiver)
    p = ep.param; nsrc = length(ep.src.srci); nrcv = length(rcv.rcvi)
    out = Vector{Float64}(undef, nrcv * (p.NSTEP + 1))
    check(ccall((:adseis_elastic_forward, libadseis), Cint,
        (Ptr{Cvoid}, Ref{CElasticParams}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64},
         Ptr{Int64}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
        ctx().handle, toC(p), rowmajor(ep.ρ), rowmajor(ep.λ), rowmajor(ep.μ), nsrc, ep.src.srci, ep.src.srcj,
        ep.src.srctype, rowmajor(ep.src.srcv), size(ep.src.srcv, 1), nrcv, rcv.rcvi, rcv.rcvj, rcv.rcvtype, out, C_NULL))
    rcv.rcvv = fromrowmajor(out, nrcv, p.NSTEP + 1)
end

"returns (loss, grad_ρ, grad_λ, grad_μ, grad_srcv)"
function elastic_misfit_grad(p::ElasticPropagatorParams, src::ElasticSource, ρ, λ, μ, rcv::ElasticReceiver,
                             obs::Matrix{Float64})
    nsrc = length(src.srci); nrcv = length(rcv.rcvi); n = length(ρ)
    loss = Ref{Float64}(0.0)
    rcvv = Vector{Float64}(undef, nrcv * (p.NSTEP + 1))
    gr = Vector{Float64}(undef, n); gl = similar(gr); gm = similar(gr); gs = Vector{Float64}(undef, p.NSTEP * nsrc)
    check(ccall((:adseis_elastic_misfit_grad, libadseis), Cint,
        (Ptr{Cvoid}, Ref{CElasticParams}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64},
         Ptr{Int64}, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ref{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx().handle, toC(p), rowmajor(ρ), rowmajor(λ), rowmajor(μ), nsrc, src.srci, src.srcj, src.srctype,
        rowmajor(src.srcv), size(src.srcv, 1), nrcv, rcv.rcvi, rcv.rcvj, rcv.rcvtype, rowmajor(obs), loss, rcvv, gr, gl,
        gm, gs))
    rcv.rcvv = fromrowmajor(rcvv, nrcv, p.NSTEP + 1)
    s1, s2 = size(ρ)
    loss[], fromrowmajor(gr, s1, s2), fromrowmajor(gl, s1, s2), fromrowmajor(gm, s1, s2), fromrowmajor(gs, p.NSTEP, nsrc)
end

end # module
