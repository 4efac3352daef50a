# Drop-in replacement module for the hot path of baggepinnen/DifferentialDynamicProgramming.jl:
# same exports and call signatures, the sweeps run in libddp.so (hand-written CUDA for sm_100a)
# through `ccall`.  No CUDA.jl, no device code in Julia.
#
# STATUS: shipped as source.  Julia is not installed in the build container, so this file has not
# been executed there; the Python mirror (differentialdynamicprogramming.jl_b200/api.py) binds the
# same C ABI with the same semantics and IS exercised by the test-suite.  The struct layouts come from
# julia/ddp_structs.jl, which is GENERATED from the table the ctypes mirror uses (scripts/gen_julia_structs.py);
# tests/test_cpu_host.py checks sizeof AND offsetof of every field of that table against include/ddp.h compiled
# with gcc, and that the generated Julia file is up to date and lays its fields out at the same offsets.
#
# Replaces (reference file:line):
#   back_pass      src/backward_pass.jl:81-252       -> ddp_back_pass_f64   (12- and 15-argument methods)
#   back_pass_gps  src/backward_pass.jl:259-350      -> ddp_back_pass_gps_f64
#   boxQP          src/boxQP.jl:29-188               -> ddp_boxqp_f64
#   forward_pass   src/forward_pass.jl:9-33          -> ddp_forward_pass_f64
#   iLQG           src/iLQG.jl:143-341               -> ddp_ilqg_solve_f64
#   iLQGkl         src/iLQGkl.jl:25-252              -> ddp_ilqgkl_solve_f64
#
# Arrays are Julia's own column-major layout with an optional TRAILING batch dimension -- byte-identical to the
# device layout, so nothing is transposed on the way: cx (n,N[,B]), fx (n,n[,N][,B]), K (m,n,N[,B]) ...
#
# User callbacks: the reference takes arbitrary closures f / costfun / df.  Closures cannot run on a GPU and this
# package has NO CPU fallback, so those arguments must be one of the device model descriptors below
# (LinearModel, PendcartModel), passed in all three positions; anything else raises an error.
module DifferentialDynamicProgramming

using LinearAlgebra
export QPTrace, boxQP, demoQP, iLQG, iLQGkl, demo_linear, demo_linear_kl, demo_pendcart, GaussianPolicy   # the reference's list
export LinearModel, PendcartModel, SimpleLTVModel, back_pass, back_pass_gps, forward_pass, Engine

const libddp = get(ENV, "LIBDDP", joinpath(@__DIR__, "..", "differentialdynamicprogramming.jl_b200", "libddp.so"))

include("ddp_structs.jl")            # GENERATED mirrors of include/ddp.h

# ---- struct construction by field name (everything not named is zero / NULL) ------------------------------------
zero_of(::Type{Ptr{Cvoid}}) = C_NULL
zero_of(::Type{T}) where {T<:Number} = zero(T)
zero_of(::Type{NTuple{N,T}}) where {N,T} = ntuple(_ -> zero(T), N)
zero_of(::Type{T}) where {T} = mk(T)                                     # nested generated struct
tofield(::Type{Ptr{Cvoid}}, v) = Ptr{Cvoid}(v)
tofield(::Type{T}, v) where {T} = convert(T, v)
function mk(::Type{T}; kw...) where {T}
    vals = map(fieldnames(T), fieldtypes(T)) do f, ft
        haskey(kw, f) ? tofield(ft, kw[f]) : zero_of(ft)
    end
    T(vals...)
end
qpopts(; maxIter = 100, minGrad = 1e-8, minRelImprove = 1e-8, stepDec = 0.6, minStep = 1e-22, Armijo = 0.1) =      # boxQP.jl:29-36
    DdpBoxQPOpts(maxIter, minGrad, minRelImprove, stepDec, minStep, Armijo)

# ---- handle + device memory -------------------------------------------------------------------
mutable struct Engine
    h::Ptr{Cvoid}
    n::Int; m::Int; T::Int; B::Int
    function Engine(n, m, T, B = 1; device = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ddp_create, libddp), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Int64, UInt32), h, device, n, m, T, B, 0)
        rc == 0 || error("libddp: ", unsafe_string(ccall((:ddp_last_error, libddp), Cstring, (Ptr{Cvoid},), C_NULL)))
        e = new(h[], n, m, T, B)
        finalizer(e -> ccall((:ddp_destroy, libddp), Cint, (Ptr{Cvoid},), e.h), e)
        e
    end
end
lasterror(e::Engine) = unsafe_string(ccall((:ddp_last_error, libddp), Cstring, (Ptr{Cvoid},), e.h))
check(e::Engine, rc) = rc == 0 || error("libddp: ", lasterror(e))
sync(e::Engine) = check(e, ccall((:ddp_synchronize, libddp), Cint, (Ptr{Cvoid},), e.h))

# device buffers of one call: everything allocated through `Pool` is freed by `release!`
struct Pool
    e::Engine
    ptrs::Vector{Ptr{Cvoid}}
end
Pool(e::Engine) = Pool(e, Ptr{Cvoid}[])
function dmalloc(p::Pool, bytes)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(p.e, ccall((:ddp_malloc, libddp), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Csize_t), p.e.h, r, max(bytes, 8)))
    push!(p.ptrs, r[])
    r[]
end
function upload(p::Pool, a::AbstractArray)          # Julia arrays are already in the device layout (column-major, batch last)
    a = Array{Float64}(a)
    d = dmalloc(p, sizeof(a))
    GC.@preserve a check(p.e, ccall((:ddp_upload, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), p.e.h, d, pointer(a), sizeof(a)))
    d
end
function zeros_dev(p::Pool, bytes)
    d = dmalloc(p, bytes)
    check(p.e, ccall((:ddp_memset, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Csize_t), p.e.h, d, 0, bytes))
    d
end
function download!(e::Engine, a::Array, d)
    GC.@preserve a check(e, ccall((:ddp_download, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), e.h, pointer(a), d, sizeof(a)))
    a
end
release!(p::Pool) = (foreach(d -> ccall((:ddp_free, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), p.e.h, d), p.ptrs); empty!(p.ptrs); nothing)

# strided views (include/ddp.h: ddp_tensor): `nd` = number of dimensions of ONE block (2 for matrices, 1 for vectors, 3 for the
# second-order tensors); a following dimension of size T is time, one of size B after that (or instead) is the batch
function tensor(d, a::AbstractArray, nd::Int, T::Int, B::Int)
    isempty(a) && return DdpTensor(C_NULL, 0, 0)
    blk = prod(size(a)[1:nd])
    dims = size(a)[nd+1:end]
    st, sb = 0, 0
    if length(dims) >= 1 && dims[1] == T && T > 1
        st = blk
        sb = (length(dims) >= 2 && dims[2] == B && B > 1) ? blk * T : 0
    elseif length(dims) >= 1 && dims[1] == B && B > 1
        sb = blk
    end
    DdpTensor(Ptr{Cvoid}(d), sb, st)
end

# ---- GaussianPolicy (iLQG.jl:39-53) -----------------------------------------------------------
mutable struct GaussianPolicy{P}
    T::Int; n::Int; m::Int
    K::Array{P}; k::Array{P}; Σ::Array{P}; Σi::Array{P}           # (m,n,T[,B]), (m,T[,B]), (m,m,T[,B]), (m,m,T[,B])
end
GaussianPolicy(P::Type) = GaussianPolicy{P}(0, 0, 0, zeros(P, 0, 0, 0), zeros(P, 0, 0), zeros(P, 0, 0, 0), zeros(P, 0, 0, 0))
GaussianPolicy(P::Type, T, n, m) = GaussianPolicy{P}(T, n, m, zeros(P, m, n, T), zeros(P, m, T), cat([Matrix{P}(I, m, m) for t = 1:T]..., dims = 3),
                                                     cat([Matrix{P}(I, m, m) for t = 1:T]..., dims = 3))
Base.isempty(gp::GaussianPolicy) = gp.T == gp.n == gp.m == 0
Base.length(gp::GaussianPolicy) = gp.T

struct QPTrace                       # boxQP.jl:1-8 (kept for source compatibility; the device returns the final state only)
    x; value; search; clamped; nfactor
end

# ---- model descriptors: stand where the reference takes closures f / costfun / df ---------------
abstract type DeviceModel end
struct LinearModel <: DeviceModel      # x+ = A x + B u ; cost ½Σx'Qx + ½Σu'Ru   (demo_linear.jl:35-50); A (n,n[,T][,B]), B (n,m[,T][,B])
    A::Array{Float64}; B::Array{Float64}; Q::Matrix{Float64}; R::Matrix{Float64}
end
struct PendcartModel <: DeviceModel    # system_pendcart.jl:51-54, 83-106
    g::Float64; l::Float64; h::Float64; d::Float64; Q::Matrix{Float64}; R::Matrix{Float64}; goal::Vector{Float64}
end
PendcartModel() = PendcartModel(9.82, 0.35, 0.01, 0.99, Matrix(Diagonal([10.0, 1, 2, 1])), fill(1.0, 1, 1), [π, 0, 0, 0])
# what iLQGkl needs of the reference's `model` argument (LinearTimeVaryingModelsBase.SimpleLTVModel, demo_linear.jl:118 -- third party):
# fx = the state Jacobian df(model,x,u) returns, R1 = covariance(model,x,u)  (forward_pass.jl:38,42)
struct SimpleLTVModel
    fx::Array{Float64}; R1::Matrix{Float64}
end
needmodel(f, c, d) = (f isa DeviceModel && f === c && (d === nothing || f === d)) ||
    error("f / costfun / df must be ONE device model descriptor (LinearModel, PendcartModel) in every position: arbitrary Julia closures " *
          "cannot run on the GPU and this package has no CPU fallback")

function devmodel(p::Pool, model::DeviceModel, N, B)
    Q, R = model.Q, model.R
    flags = Int32(isdiag(Q) ? 1 : 0)
    if model isa LinearModel
        mk(DdpModel; kind = 1, A = tensor(upload(p, model.A), model.A, 2, N, B), Bm = tensor(upload(p, model.B), model.B, 2, N, B),
           Q = DdpTensor(upload(p, Q), 0, 0), R = DdpTensor(upload(p, R), 0, 0), flags = flags)
    else
        mk(DdpModel; kind = 2, Q = DdpTensor(upload(p, Q), 0, 0), R = DdpTensor(upload(p, R), 0, 0), goal = upload(p, model.goal),
           p = (model.g, model.l, model.h, model.d, 0.0, 0.0, 0.0, 0.0), terminal_cost = 1, flags = flags)
    end
end

# ---- back_pass   backward_pass.jl:162/179/217 (12 arguments) and :81/:132 (15 arguments) -------
back_pass(cx, cu, cxx, cxu, cuu, fx, fu, λ, regType, lims, x, u) = back_pass(cx, cu, cxx, cxu, cuu, fx, fu, [], [], [], λ, regType, lims, x, u)
function back_pass(cx, cu, cxx, cxu, cuu, fx, fu, fxx, fxu, fuu, λ, regType, lims, x, u)
    n, N = size(cx, 1), size(cx, 2); m = size(cu, 1); B = size(cx, 3)
    size(cu, 2) == N || error("size(cu) should be (m, N)")
    e = Engine(n, m, N, B); p = Pool(e)
    try
        K = zeros(m, n, N, B); k = zeros(m, N, B); Vx = zeros(n, N, B); Vxx = zeros(n, n, N, B); Quu = zeros(m, m, N, B); dV = zeros(2, B); dv = zeros(Int32, B)
        outs = (K, k, Vx, Vxx, Quu, dV, dv)
        d = map(a -> dmalloc(p, sizeof(a)), outs)
        uselims = !isempty(lims)
        a = mk(DdpBackPassArgs;
               cx = tensor(upload(p, cx), cx, 1, N, B), cu = tensor(upload(p, cu), cu, 1, N, B),
               cxx = tensor(upload(p, cxx), cxx, 2, N, B), cxu = tensor(upload(p, cxu), cxu, 2, N, B), cuu = tensor(upload(p, cuu), cuu, 2, N, B),
               fx = tensor(upload(p, fx), fx, 2, N, B), fu = tensor(upload(p, fu), fu, 2, N, B),
               fxx = isempty(fxx) ? DdpTensor(C_NULL, 0, 0) : tensor(upload(p, fxx), fxx, 3, N, B),
               fxu = isempty(fxu) ? DdpTensor(C_NULL, 0, 0) : tensor(upload(p, fxu), fxu, 3, N, B),
               fuu = isempty(fuu) ? DdpTensor(C_NULL, 0, 0) : tensor(upload(p, fuu), fuu, 3, N, B),
               lam = upload(p, fill(Float64(λ), B) .* ones(B)), reg_type = Int32(regType),
               lims = uselims ? upload(p, lims) : C_NULL,                       # (m,2) column-major = [lower; upper]
               u = uselims ? tensor(upload(p, u), u, 1, N, B) : DdpTensor(C_NULL, 0, 0),
               K = d[1], k = d[2], Vx = d[3], Vxx = d[4], Quu = d[5], dV = d[6], diverge = d[7], qp = qpopts())
        check(e, ccall((:ddp_back_pass_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpBackPassArgs}), e.h, a))
        sync(e)
        foreach((arr, dp) -> download!(e, arr, dp), outs, d)
        # field order of the reference's return (backward_pass.jl:251): Σ is never written there (quirk Q2)
        sq(a) = B == 1 ? dropdims(a, dims = ndims(a)) : a
        return (B == 1 ? Int(dv[1]) : Int.(dv)), GaussianPolicy{Float64}(N, n, m, sq(K), sq(k), Array{Float64}(undef, m, m, N), sq(Quu)), sq(Vx), sq(Vxx), sq(dV)
    finally
        release!(p)
    end
end

# ---- back_pass_gps(cx,cu,cxx,cxu,cuu,fx,fu,lims,x,u,kl_cost_terms)   backward_pass.jl:259 ----------
# kl_cost_terms = (traj_prev::GaussianPolicy, ηbracket): the five KL tensors of ∇kl (klutils.jl:8-23) are formed on the device from
# the previous policy, so the policy itself is passed instead of the pre-multiplied tensors.
function back_pass_gps(cx, cu, cxx, cxu, cuu, fx, fu, lims, x, u, kl_cost_terms)
    traj_prev, ηbracket = kl_cost_terms
    n, N = size(cx, 1), size(cx, 2); m = size(cu, 1); B = size(cx, 3)
    e = Engine(n, m, N, B); p = Pool(e)
    try
        K = zeros(m, n, N, B); k = zeros(m, N, B); Vx = zeros(n, N, B); Vxx = zeros(n, n, N, B); Quu = zeros(m, m, N, B); Quui = zeros(m, m, N, B)
        dV = zeros(2, B); dv = zeros(Int32, B)
        outs = (K, k, Vx, Vxx, Quu, Quui, dV, dv)
        d = map(a -> dmalloc(p, sizeof(a)), outs)
        uselims = !isempty(lims)
        η = ndims(ηbracket) == 1 ? fill(Float64(ηbracket[2]), B) : Float64.(vec(ηbracket[2, :]))
        a = mk(DdpBackPassArgs;
               cx = tensor(upload(p, cx), cx, 1, N, B), cu = tensor(upload(p, cu), cu, 1, N, B),
               cxx = tensor(upload(p, cxx), cxx, 2, N, B), cxu = tensor(upload(p, cxu), cxu, 2, N, B), cuu = tensor(upload(p, cuu), cuu, 2, N, B),
               fx = tensor(upload(p, fx), fx, 2, N, B), fu = tensor(upload(p, fu), fu, 2, N, B),
               lims = uselims ? upload(p, lims) : C_NULL, u = uselims ? tensor(upload(p, u), u, 1, N, B) : DdpTensor(C_NULL, 0, 0),
               K = d[1], k = d[2], Vx = d[3], Vxx = d[4], Quu = d[5], dV = d[7], diverge = d[8], qp = qpopts())
        g = mk(DdpGpsArgs; K_prev = tensor(upload(p, traj_prev.K), traj_prev.K, 2, N, B), k_prev = tensor(upload(p, traj_prev.k), traj_prev.k, 1, N, B),
               Sigi_prev = tensor(upload(p, traj_prev.Σi), traj_prev.Σi, 2, N, B), eta = upload(p, η), Quui = d[6])
        check(e, ccall((:ddp_back_pass_gps_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpBackPassArgs}, Ref{DdpGpsArgs}), e.h, a, g))
        sync(e)
        foreach((arr, dp) -> download!(e, arr, dp), outs, d)
        sq(a) = B == 1 ? dropdims(a, dims = ndims(a)) : a
        return (B == 1 ? Int(dv[1]) : Int.(dv)), GaussianPolicy{Float64}(N, n, m, sq(K), sq(k), sq(Quui), sq(Quu)), sq(Vx), sq(Vxx), sq(dV)   # :349
    finally
        release!(p)
    end
end

# ---- forward_pass(traj_new,x0,u,x,α,f,costfun,lims,diff)   forward_pass.jl:9 --------------------
function forward_pass(traj_new::GaussianPolicy, x0, u, x, α, f, costfun, lims, diff = -)
    needmodel(f, costfun, nothing)
    diff === (-) || error("only the default diff_fun (-) is supported on the device")
    m, N = size(u, 1), size(u, 2); B = size(u, 3); n = size(x0, 1)
    e = Engine(n, m, N, B); p = Pool(e)
    try
        md = devmodel(p, f, N, B)
        xnew = zeros(n, N, B); unew = zeros(m, N, B); cost = zeros(B)
        d = map(a -> dmalloc(p, sizeof(a)), (xnew, unew, cost))
        x0b = repeat(reshape(Array{Float64}(x0), n, :), 1, size(x0, 2) == B ? 1 : B)
        pol = !isempty(traj_new)
        a = mk(DdpForwardPassArgs;
               K = pol ? upload(p, traj_new.K) : C_NULL, k = pol ? upload(p, traj_new.k) : C_NULL,
               x0 = DdpTensor(upload(p, x0b), n, 0), x = pol ? tensor(upload(p, x), x, 1, N, B) : DdpTensor(C_NULL, 0, 0),
               u = tensor(upload(p, u), u, 1, N, B), alpha_scalar = Float64(α), u_scale = 1.0,
               lims = isempty(lims) ? C_NULL : upload(p, lims), xnew = d[1], unew = d[2], cost = d[3])
        check(e, ccall((:ddp_forward_pass_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpModel}, Ref{DdpForwardPassArgs}), e.h, md, a))
        sync(e)
        foreach((arr, dp) -> download!(e, arr, dp), (xnew, unew, cost), d)
        return B == 1 ? (xnew[:, :, 1], unew[:, :, 1], cost[1]) : (xnew, unew, cost)
    finally
        release!(p)
    end
end

# ---- boxQP(H,g,lower,upper,x0)   boxQP.jl:29 ---------------------------------------------------
function boxQP(H, g, lower, upper, x0::AbstractVector; maxIter = 100, minGrad = 1e-8, minRelImprove = 1e-8, stepDec = 0.6,
               minStep = 1e-22, Armijo = 0.1, print = 0)
    m = size(H, 1)
    m <= 16 || return boxQP_large(H, g, lower, upper, x0; maxIter = maxIter, minGrad = minGrad, minRelImprove = minRelImprove, stepDec = stepDec,
                                  minStep = minStep, Armijo = Armijo)
    e = Engine(m, m, 1, 1); p = Pool(e)
    try
        dH, dg, dl, du, dx0 = (upload(p, v) for v in (H, g, lower, upper, x0))
        x = zeros(m); res = Int32[0]; Hf = zeros(m, m); fm = UInt32[0]; nf = Int32[0]
        o = map(a -> dmalloc(p, sizeof(a)), (x, res, Hf, fm, nf))
        opts = qpopts(maxIter = maxIter, minGrad = minGrad, minRelImprove = minRelImprove, stepDec = stepDec, minStep = minStep, Armijo = Armijo)
        check(e, ccall((:ddp_boxqp_f64, libddp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
            Ref{DdpBoxQPOpts}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), e.h, 1, dH, dg, dl, du, dx0, opts, o...))
        sync(e)
        foreach((arr, dp) -> download!(e, arr, dp), (x, res, Hf, fm, nf), o)
        res[1] < 0 && throw(PosDefException(-1))                                   # where the reference's cholesky throws
        free = [((fm[1] >> (i - 1)) & 1) == 1 for i in 1:m]
        nfree = count(free)
        return x, Int(res[1]), Hf[1:nfree, 1:nfree], free, QPTrace[]
    finally
        release!(p)
    end
end

# large problems (demoQP's n = 500, boxQP.jl:190-199): ddp_boxqp_large_f64, one CTA per problem
function boxQP_large(H, g, lower, upper, x0; maxIter = 100, minGrad = 1e-8, minRelImprove = 1e-8, stepDec = 0.6, minStep = 1e-22, Armijo = 0.1)
    n = size(H, 1)
    e = Engine(1, 1, 1, 1); p = Pool(e)
    try
        dH, dg, dl, du, dx0 = (upload(p, v) for v in (H, g, lower, upper, x0))
        x = zeros(n); res = Int32[0]; Hf = zeros(n, n); free = zeros(UInt8, n); nf = Int32[0]
        o = map(a -> dmalloc(p, sizeof(a)), (x, res, Hf, free, nf))
        opts = qpopts(maxIter = maxIter, minGrad = minGrad, minRelImprove = minRelImprove, stepDec = stepDec, minStep = minStep, Armijo = Armijo)
        check(e, ccall((:ddp_boxqp_large_f64, libddp), Cint, (Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
            Ref{DdpBoxQPOpts}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), e.h, n, 1, dH, dg, dl, du, dx0, opts, o...))
        sync(e)
        foreach((arr, dp) -> download!(e, arr, dp), (x, res, Hf, free, nf), o)
        res[1] < 0 && throw(PosDefException(-1))
        fr = free .!= 0
        nfree = count(fr)
        return x, Int(res[1]), Hf[1:nfree, 1:nfree], fr, QPTrace[]
    finally
        release!(p)
    end
end

# ---- iLQG(f,costfun,df,x0,u0; kw...)   iLQG.jl:143 ---------------------------------------------
# x0 (n,) / (n,1[,B]): initial state; x0 (n,N[,B]) + `cost`: pre-rolled trajectory (iLQG.jl:193-197).  The returned trace is a Dict
# with the reference's MVHistory keys (:λ,:dλ,:cost,:α,:grad_norm,:improvement,:reduce_ratio => (iterations, values)), rebuilt from
# the per-iteration records the device keeps (ddp_ilqg_trace); B > 1 returns batched arrays and a Vector of such Dicts.
function iLQG(f, costfun, df, x0, u0;
              lims = [], α = exp10.(range(0, stop = -3, length = 11)), tol_fun = 1e-7, tol_grad = 1e-4, max_iter = 500,
              λ = 1.0, dλ = 1.0, λfactor = 1.6, λmax = 1e10, λmin = 1e-6, regType = 1, reduce_ratio_min = 0, diff_fun = -,
              cost = [], verbosity = 0, kwargs...)
    needmodel(f, costfun, df)
    diff_fun === (-) || error("only the default diff_fun (-) is supported on the device")
    model = f
    n = size(x0, 1); m, N = size(u0, 1), size(u0, 2); B = size(u0, 3)
    prerolled = size(x0, 2) == N && N > 1
    prerolled || size(x0, 2) == 1 || error("pre-rolled initial trajectory must be of correct length (size(x0,2) == N)")    # iLQG.jl:199
    prerolled && isempty(cost) && error("Initial trajectory supplied, initial cost must also be supplied")
    e = Engine(n, m, N, B); p = Pool(e)
    try
        md = devmodel(p, model, N, B)
        αt = ntuple(i -> i <= length(α) ? Float64(α[i]) : 0.0, 16)
        cap = 4 * max_iter + 64
        trace_d = dmalloc(p, cap * B * sizeof(DdpIlqgTrace))
        init = fill(DdpIlqgTrace(NaN, NaN, NaN, NaN, NaN, NaN, NaN, Int32(-1), Int32(0)), cap, B)       # the device keeps record (it, b) at it*B + b
        init = permutedims(init)                                                                         # Julia (B,cap) column-major == C [cap][B]
        GC.@preserve init check(e, ccall((:ddp_upload, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), e.h, trace_d, pointer(init), sizeof(init)))
        x_init, cost_init = C_NULL, C_NULL
        if prerolled
            x_init = upload(p, x0)
            c = Array{Float64}(cost)
            cost_init = upload(p, length(c) == B ? vec(c) : vec(sum(reshape(c, :, B), dims = 1)))
            x0b = reshape(Array{Float64}(x0), n, N, B)[:, 1, :]
        else
            x0b = repeat(reshape(Array{Float64}(x0), n, :), 1, size(reshape(Array{Float64}(x0), n, :), 2) == B ? 1 : B)
        end
        opts = mk(DdpIlqgOpts; n_alpha = Int32(length(α)), alpha = αt, tol_fun = tol_fun, tol_grad = tol_grad, max_iter = Int32(max_iter),
                  lam = λ, dlam = dλ, lam_factor = λfactor, lam_max = λmax, lam_min = λmin, reg_type = Int32(regType),
                  reduce_ratio_min = Float64(reduce_ratio_min), lims = isempty(lims) ? C_NULL : upload(p, lims),
                  x_init = x_init, cost_init = cost_init, trace = trace_d, trace_cap = Int32(cap))
        x = zeros(n, N, B); u = zeros(m, N, B); K = zeros(m, n, N, B); k = zeros(m, N, B); Vx = zeros(n, N, B); Vxx1 = zeros(n, n, B)
        st = Vector{DdpIlqgState}(undef, B)
        outs = (x, u, K, k, Vx, Vxx1)
        d = map(a -> zeros_dev(p, sizeof(a)), outs)
        dst = dmalloc(p, sizeof(st))
        nouter = Ref{Int32}(0)
        rc = ccall((:ddp_ilqg_solve_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpModel}, Ref{DdpIlqgOpts}, Ptr{Cvoid}, Ptr{Cvoid},
            Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Int32}),
            e.h, md, opts, upload(p, x0b), upload(p, u0), d..., dst, nouter)
        rc == -5 || check(e, rc)                       # DDP_ERR_INCOMPLETE still delivers every output (status 6 marks the stragglers)
        foreach((arr, dp) -> download!(e, arr, dp), outs, d)
        download!(e, st, dst)
        tr = Matrix{DdpIlqgTrace}(undef, B, cap)
        download!(e, tr, trace_d)
        traces = [ilqg_trace(tr[b, :], st[b], λ, dλ, nothing) for b in 1:B]
        if B == 1
            st[1].status == 4 && return nothing                                                     # iLQG.jl:205-210
            st[1].iter == 1 && error("Failure: no iterations completed, something is wrong.")       # iLQG.jl:335
            pol = GaussianPolicy{Float64}(N, n, m, K[:, :, :, 1], k[:, :, 1], zeros(m, m, 0), zeros(m, m, 0))
            return x[:, :, 1], u[:, :, 1], pol, Vx[:, :, 1], Vxx1[:, :, 1], st[1].cost, traces[1]
        end
        return x, u, GaussianPolicy{Float64}(N, n, m, K, k, zeros(m, m, 0), zeros(m, m, 0)), Vx, Vxx1, [s.cost for s in st], traces
    finally
        release!(p)
    end
end

# the reference's trace keys (iLQG.jl:176-177, 257, 325-330) from the device records: key => (iterations, values)
function ilqg_trace(rec::Vector{DdpIlqgTrace}, st::DdpIlqgState, λ0, dλ0, _)
    tr = Dict{Symbol,Tuple{Vector{Int},Vector{Float64}}}()
    put!(key, it, v) = (haskey(tr, key) || (tr[key] = (Int[], Float64[])); push!(tr[key][1], it); push!(tr[key][2], v))
    put!(:λ, 0, λ0); put!(:dλ, 0, dλ0)
    for (it, r) in enumerate(rec)
        it > st.iter && break
        isnan(r.grad_norm) || put!(:grad_norm, it, r.grad_norm)
        r.accepted < 0 && continue                                    # the reference broke out before "update trace" (:306-309, :319-322)
        put!(:λ, it, r.lam); put!(:dλ, it, r.dlam); put!(:α, it, r.alpha); put!(:improvement, it, r.improvement)
        put!(:cost, it, r.cost); put!(:reduce_ratio, it, r.reduce_ratio)
    end
    tr[:status] = ([Int(st.status)], [st.cost])
    tr
end

# ---- multi-GPU: one process per GPU, the batch sharded by contiguous ranges; the only exchange is the statistics all-reduce
comm_unique_id() = (id = zeros(UInt8, 128); ccall((:ddp_comm_unique_id, libddp), Cint, (Ptr{UInt8},), id) == 0 || error("libddp: NCCL not available"); id)
comm_init(e::Engine, nranks, rank, id::Vector{UInt8}) = check(e, ccall((:ddp_comm_init, libddp), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), e.h, nranks, rank, id))
allreduce_stats!(e::Engine, stats8_dev::Ptr{Cvoid}) = check(e, ccall((:ddp_comm_allreduce_stats_f64, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.h, stats8_dev))

# ---- iLQGkl(dynamics,costfun,derivs,x0,traj_prev,model; kw...)   iLQGkl.jl:25 --------------------
# Whole outer loop on the device (ddp_ilqgkl_solve_f64).  x0 is the pre-rolled trajectory (n,N[,B]) and `cost` its cost
# (iLQGkl.jl:63-70); traj_prev a GaussianPolicy with K (m,n,N[,B]), k, Σ, Σi; `model` a SimpleLTVModel (fx, R1).
# traj_prev.k is left untouched (the reference zeroes it during the call and restores it at :247).
function iLQGkl(dynamics, costfun, derivs, x0, traj_prev::GaussianPolicy, model::SimpleLTVModel;
                constrain_per_step = false, kl_step = 1, lims = [], max_iter = 50, cost = [], ηbracket = [1e-8, 1, 1e16], del0 = 0.0001,
                diff_fun = -, verbosity = 0, kwargs...)
    needmodel(dynamics, costfun, derivs)
    constrain_per_step && error("constrain_per_step = true is not supported (the reference's branch is broken on Julia >= 1.0: klutils.jl:150,195)")
    isempty(cost) && error("Initial trajectory supplied, initial cost must also be supplied")         # iLQGkl.jl:66-68
    dm = dynamics
    n, N = size(x0, 1), size(x0, 2); m = size(traj_prev.k, 1); B = size(x0, 3)
    size(traj_prev.k, 2) == N || error("pre-rolled initial trajectory must be of correct length (size(x0,2) == N)")
    e = Engine(n, m, N, B); p = Pool(e)
    try
        md = devmodel(p, dm, N, B)
        opts = mk(DdpIlqgklOpts; kl_step = Float64(kl_step), max_iter = Int32(max_iter), eta_bracket = (Float64(ηbracket[1]), Float64(ηbracket[2]), Float64(ηbracket[3])),
                  del0 = Float64(del0), max_eta_retries = Int32(200), lims = isempty(lims) ? C_NULL : upload(p, lims))
        c = Array{Float64}(cost)
        costs = length(c) == B ? vec(c) : vec(sum(reshape(c, :, B), dims = 1))
        xnew = zeros(n, N, B); unew = zeros(m, N, B); K = zeros(m, n, N, B); k = zeros(m, N, B); Σ = zeros(m, m, N, B); Σi = zeros(m, m, N, B)
        Vx = zeros(n, N, B); Vxx1 = zeros(n, n, B); cnew = zeros(B); st = Vector{DdpIlqgklState}(undef, B)
        outs = (xnew, unew, K, k, Σ, Σi, Vx, Vxx1, cnew, st)
        d = map(a -> zeros_dev(p, sizeof(a)), outs)
        a = mk(DdpIlqgklArgs; x = upload(p, x0), u = upload(p, traj_prev.k), cost = upload(p, costs),                 # u = traj_prev.k (iLQGkl.jl:47)
               K_prev = tensor(upload(p, traj_prev.K), traj_prev.K, 2, N, B), Sig_prev = tensor(upload(p, traj_prev.Σ), traj_prev.Σ, 2, N, B),
               Sigi_prev = tensor(upload(p, traj_prev.Σi), traj_prev.Σi, 2, N, B),
               fx_model = tensor(upload(p, model.fx), model.fx, 2, N, B), R1 = tensor(upload(p, model.R1), model.R1, 2, 1, B),
               xnew = d[1], unew = d[2], K = d[3], k = d[4], Sig = d[5], Sigi = d[6], Vx = d[7], Vxx1 = d[8], costnew = d[9], state = d[10])
        nouter = Ref{Int32}(0)
        check(e, ccall((:ddp_ilqgkl_solve_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpModel}, Ref{DdpIlqgklOpts}, Ref{DdpIlqgklArgs}, Ref{Int32}),
                       e.h, md, opts, a, nouter))
        foreach((arr, dp) -> download!(e, arr, dp), outs, d)
        trace = [Dict(:η => s.eta, :ηbracket => (s.eta_min, s.eta, s.eta_max), :divergence => s.divergence, :improvement => s.dcost,
                      :expected => s.expected, :cost => s.cost, :iter => Int(s.iter), :status => Int(s.status)) for s in st]
        if B == 1
            pol = GaussianPolicy{Float64}(N, n, m, K[:, :, :, 1], k[:, :, 1], Σ[:, :, :, 1], Σi[:, :, :, 1])
            return xnew[:, :, 1], unew[:, :, 1], pol, Vx[:, :, 1], Vxx1[:, :, 1], cnew[1], trace[1]
        end
        return xnew, unew, GaussianPolicy{Float64}(N, n, m, K, k, Σ, Σi), Vx, Vxx1, cnew, trace
    finally
        release!(p)
    end
end

# ---- the demos (exports of the reference): same problems, the device model descriptors in place of the closures ---------------
function demo_linear(; kwargs...)                      # demo_linear.jl:5-60
    T, n, m, h = 1000, 10, 2, 0.01
    G = randn(n, n)
    A = exp(h * (G - G'))                              # :14
    Bm = h * randn(n, m)
    model = LinearModel(A, Bm, h * Matrix{Float64}(I, n, n), 0.1h * Matrix{Float64}(I, m, m))      # Q = hI, R = 0.1hI (:18-19)
    x0, u0 = ones(n), 0.1 * randn(m, T)
    iLQG(model, model, model, x0, u0; kwargs...)
end

function demo_linear_kl(; kl_step = 1.0, kwargs...)    # demo_linear.jl:63-136: five KL-constrained outer iterations from the zero policy
    T, n, m, h = 1000, 10, 2, 0.01
    G = randn(n, n)
    A = exp(h * (G - G')); Bm = h * randn(n, m)
    Q, R = h * Matrix{Float64}(I, n, n), 0.1h * Matrix{Float64}(I, m, m)
    model = LinearModel(A, Bm, Q, R)
    x = zeros(n, T); x[:, 1] .= 1.0
    for t = 1:T-1; x[:, t+1] = A * x[:, t]; end        # rollout of u = 0 (:104-111 with traj.k = 0)
    traj = GaussianPolicy(Float64, T, n, m)
    ltv = SimpleLTVModel(A, 1e-4 * Matrix{Float64}(I, n, n))          # stand-in for covariance(model,x,u) of the third-party type
    local out
    for iter = 1:5
        u = traj.k
        cost0 = 0.5 * sum(x .* (Q * x)) + 0.5 * sum(u .* (R * u))
        out = iLQGkl(model, model, model, x, traj, ltv; cost = cost0, kl_step = kl_step, kwargs...)
        x, traj = out[1], out[3]
    end
    out
end

function demo_pendcart(; x0 = [π - 0.6, 0, 0, 0], goal = [π, 0, 0, 0], Q = Diagonal([10.0, 1, 2, 1]), R = 1.0, lims = 5.0 * [-1 1], T = 600, kwargs...)
    # system_pendcart.jl:42-212 without the LQR warm-up simulation (host-side setup, SURVEY section 2 #13): u0 = 0
    model = PendcartModel(9.82, 0.35, 0.01, 0.99, Matrix{Float64}(Q), fill(Float64(R), 1, 1), Float64.(goal))
    iLQG(model, model, model, Float64.(x0), zeros(1, T); lims = lims, regType = 2, α = exp10.(range(0.2, stop = -3, length = 6)), λmax = 1e15,
         tol_fun = 1e-8, tol_grad = 1e-8, max_iter = 1000, kwargs...)                                   # :197-206
end

function demoQP(; n = 500, kwargs...)                  # boxQP.jl:190-199
    g = randn(n); H = randn(n, n); H = H * H'
    lower = -ones(n); upper = ones(n)
    boxQP(H, g, lower, upper, randn(n); kwargs...)
end

end # module
