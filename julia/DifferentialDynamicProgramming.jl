# Drop-in replacement module for the hot path of baggepinnen/DifferentialDynamicProgramming.jl:
# same exports and call signatures, the sweeps run in libddp.so (hand-written CUDA for sm_100a)
# through `ccall`.  No CUDA.jl, no device code in Julia.
#
# STATUS: shipped as source.  Julia is not installed in the build container, so this file has not
# been executed there; the Python mirror (differentialdynamicprogramming.jl_b200/api.py) binds the
# same C ABI with the same semantics and IS exercised by the test-suite.  Struct layouts below
# mirror include/ddp.h one to one (checked against ctypes in tests/test_cpu_host.py).
#
# Replaces (reference file:line):
#   back_pass      src/backward_pass.jl:162-252      -> ddp_back_pass_f64
#   back_pass_gps  src/backward_pass.jl:259-350      -> ddp_back_pass_gps_f64
#   boxQP          src/boxQP.jl:29-188               -> ddp_boxqp_f64
#   forward_pass   src/forward_pass.jl:9-33          -> ddp_forward_pass_f64
#   iLQG           src/iLQG.jl:143-341               -> ddp_ilqg_solve_f64
module DifferentialDynamicProgramming

using LinearAlgebra
export iLQG, iLQGkl, boxQP, GaussianPolicy, LinearModel, PendcartModel, back_pass, forward_pass

const libddp = get(ENV, "LIBDDP", joinpath(@__DIR__, "..", "differentialdynamicprogramming.jl_b200", "libddp.so"))

# ---- mirrors of include/ddp.h ---------------------------------------------------------------
struct DdpTensor
    ptr::Ptr{Float64}
    stride_b::Int64
    stride_t::Int64
end
DdpTensor() = DdpTensor(C_NULL, 0, 0)

struct DdpBoxQPOpts
    max_iter::Int32
    min_grad::Float64
    min_rel_improve::Float64
    step_dec::Float64
    min_step::Float64
    armijo::Float64
end
DdpBoxQPOpts() = DdpBoxQPOpts(100, 1e-8, 1e-8, 0.6, 1e-22, 0.1)     # boxQP.jl:29-36

struct DdpBackPassArgs
    cx::DdpTensor; cu::DdpTensor; cxx::DdpTensor; cxu::DdpTensor; cuu::DdpTensor; fx::DdpTensor; fu::DdpTensor
    lambda::Ptr{Float64}
    reg_type::Int32
    lims::Ptr{Float64}
    u::DdpTensor
    active::Ptr{UInt8}
    diverge::Ptr{Int32}
    K::Ptr{Float64}; k::Ptr{Float64}; Vx::Ptr{Float64}; Vxx::Ptr{Float64}; Vxx1::Ptr{Float64}; Quu::Ptr{Float64}; dV::Ptr{Float64}
    qp::DdpBoxQPOpts
end

struct DdpModel
    kind::Int32
    A::DdpTensor; Bm::DdpTensor; Q::DdpTensor; R::DdpTensor
    goal::Ptr{Float64}
    p::NTuple{8,Float64}
    terminal_cost::Int32
    flags::Int32
end

struct DdpForwardPassArgs
    K::Ptr{Float64}; k::Ptr{Float64}
    x0::DdpTensor; x::DdpTensor; u::DdpTensor
    alpha::Ptr{Float64}; alpha_scalar::Float64; u_scale::Float64
    lims::Ptr{Float64}; active::Ptr{UInt8}
    xnew::Ptr{Float64}; unew::Ptr{Float64}; cost::Ptr{Float64}; cost_t::Ptr{Float64}; cx::Ptr{Float64}; cu::Ptr{Float64}
end

struct DdpIlqgOpts
    n_alpha::Int32
    alpha::NTuple{16,Float64}
    tol_fun::Float64; tol_grad::Float64
    max_iter::Int32
    lambda::Float64; dlambda::Float64; lambda_factor::Float64; lambda_max::Float64; lambda_min::Float64
    reg_type::Int32
    reduce_ratio_min::Float64
    lims::Ptr{Float64}
end

struct DdpIlqgState
    lambda::Float64; dlambda::Float64; cost::Float64; g_norm::Float64; last_dcost::Float64; last_alpha::Float64
    iter::Int32; accepted_iter::Int32; status::Int32; pad::Int32
end

struct DdpIlqgklOpts                      # ddp_ilqgkl_opts (defaults: iLQGkl.jl:25-42)
    kl_step::Float64
    max_iter::Int32
    eta_bracket::NTuple{3,Float64}
    del0::Float64
    max_eta_retries::Int32
    lims::Ptr{Float64}
end

struct DdpIlqgklState
    eta_min::Float64; eta::Float64; eta_max::Float64; del0::Float64; divergence::Float64; dcost::Float64; expected::Float64; cost::Float64
    iter::Int32; status::Int32; retries::Int32; pad::Int32
end

struct DdpIlqgklArgs
    x::Ptr{Float64}; u::Ptr{Float64}; cost::Ptr{Float64}
    K_prev::DdpTensor; Sig_prev::DdpTensor; Sigi_prev::DdpTensor; fx_model::DdpTensor; R1::DdpTensor
    xnew::Ptr{Float64}; unew::Ptr{Float64}; K::Ptr{Float64}; k::Ptr{Float64}; Sig::Ptr{Float64}; Sigi::Ptr{Float64}
    Vx::Ptr{Float64}; Vxx1::Ptr{Float64}; costnew::Ptr{Float64}; state::Ptr{DdpIlqgklState}
end

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

check(e::Engine, rc) = rc == 0 || error("libddp: ", unsafe_string(ccall((:ddp_last_error, libddp), Cstring, (Ptr{Cvoid},), e.h)))

function dmalloc(e::Engine, bytes)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(e, ccall((:ddp_malloc, libddp), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Csize_t), e.h, p, max(bytes, 8)))
    p[]
end
dfree(e::Engine, p) = ccall((:ddp_free, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.h, p)
function upload(e::Engine, a::Array)          # Julia arrays are already in the device layout (column-major, batch last)
    p = dmalloc(e, sizeof(a))
    GC.@preserve a check(e, ccall((:ddp_upload, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), e.h, p, pointer(a), sizeof(a)))
    p
end
function download!(e::Engine, a::Array, p)
    GC.@preserve a check(e, ccall((:ddp_download, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), e.h, pointer(a), p, sizeof(a)))
    a
end
# strided view: a 2-D matrix is time-invariant and shared; a trailing time (and batch) axis adds strides
tensor(p, a::AbstractArray, T, B) = DdpTensor(Ptr{Float64}(p),
    (ndims(a) >= 4 && size(a, 4) == B && B > 1) ? stride(a, 4) : 0,
    (ndims(a) >= 3 && size(a, 3) == T && T > 1) ? stride(a, 3) : 0)
vtensor(p, a::AbstractArray, B) = DdpTensor(Ptr{Float64}(p), (ndims(a) >= 3 && B > 1) ? stride(a, 3) : 0, stride(a, 2))

# ---- GaussianPolicy (iLQG.jl:39-53) -----------------------------------------------------------
mutable struct GaussianPolicy{P}
    T::Int; n::Int; m::Int
    K::Array{P,3}; k::Array{P,2}; Σ::Array{P,3}; Σi::Array{P,3}
end
GaussianPolicy(P) = GaussianPolicy(0, 0, 0, Array{P}(undef, 0, 0, 0), Array{P}(undef, 0, 0), Array{P}(undef, 0, 0, 0), Array{P}(undef, 0, 0, 0))
Base.isempty(gp::GaussianPolicy) = gp.T == gp.n == gp.m == 0
Base.length(gp::GaussianPolicy) = gp.T

# ---- model descriptors: stand where the reference takes closures f / costfun / df ---------------
abstract type DeviceModel end
struct LinearModel <: DeviceModel      # x+ = A x + B u ; cost ½Σx'Qx + ½Σu'Ru   (demo_linear.jl:35-50)
    A::Array{Float64}; B::Array{Float64}; Q::Matrix{Float64}; R::Matrix{Float64}
end
struct PendcartModel <: DeviceModel    # system_pendcart.jl:51-54, 83-106
    g::Float64; l::Float64; h::Float64; d::Float64; Q::Matrix{Float64}; R::Matrix{Float64}; goal::Vector{Float64}
end
PendcartModel() = PendcartModel(9.82, 0.35, 0.01, 0.99, Matrix(Diagonal([10.0, 1, 2, 1])), fill(1.0, 1, 1), [π, 0, 0, 0])

# ---- back_pass(cx,cu,cxx,cxu,cuu,fx,fu,λ,regType,lims,x,u)   backward_pass.jl:162/179/217 -------
function back_pass(cx, cu, cxx, cxu, cuu, fx, fu, λ, regType, lims, x, u)
    n, N = size(cx); m = size(cu, 1)
    e = Engine(n, m, N, 1)
    d = Dict(k => upload(e, Array{Float64}(v)) for (k, v) in pairs((; cx, cu, cxx, cxu, cuu, fx, fu, u)))
    λd = upload(e, [Float64(λ)])
    limsd = (isempty(lims) ? C_NULL : upload(e, Array{Float64}(lims)))          # (m,2) column-major = [lower; upper]
    K = zeros(m, n, N); k = zeros(m, N); Vx = zeros(n, N); Vxx = zeros(n, n, N); Quu = zeros(m, m, N); dV = zeros(2); dv = Int32[0]
    out = Dict(s => dmalloc(e, sizeof(a)) for (s, a) in pairs((; K, k, Vx, Vxx, Quu, dV, dv)))
    a = DdpBackPassArgs(vtensor(d[:cx], cx, 1), vtensor(d[:cu], cu, 1), tensor(d[:cxx], cxx, N, 1), tensor(d[:cxu], cxu, N, 1),
        tensor(d[:cuu], cuu, N, 1), tensor(d[:fx], fx, N, 1), tensor(d[:fu], fu, N, 1), Ptr{Float64}(λd), regType, Ptr{Float64}(limsd),
        vtensor(d[:u], u, 1), C_NULL, Ptr{Int32}(out[:dv]), Ptr{Float64}(out[:K]), Ptr{Float64}(out[:k]), Ptr{Float64}(out[:Vx]),
        Ptr{Float64}(out[:Vxx]), C_NULL, Ptr{Float64}(out[:Quu]), Ptr{Float64}(out[:dV]), DdpBoxQPOpts())
    check(e, ccall((:ddp_back_pass_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpBackPassArgs}), e.h, a))
    check(e, ccall((:ddp_synchronize, libddp), Cint, (Ptr{Cvoid},), e.h))
    for (s, arr) in pairs((; K, k, Vx, Vxx, Quu, dV, dv)); download!(e, arr, out[s]); end
    foreach(p -> dfree(e, p), values(d)); foreach(p -> dfree(e, p), values(out)); dfree(e, λd)
    # field order of the reference's return (backward_pass.jl:251): Σ is never written there (quirk Q2)
    return Int(dv[1]), GaussianPolicy(N, n, m, K, k, Array{Float64}(undef, m, m, N), Quu), Vx, Vxx, dV
end

# ---- boxQP(H,g,lower,upper,x0)   boxQP.jl:29 ---------------------------------------------------
function boxQP(H, g, lower, upper, x0::AbstractVector; maxIter = 100, minGrad = 1e-8, minRelImprove = 1e-8, stepDec = 0.6,
               minStep = 1e-22, Armijo = 0.1, print = 0)
    m = size(H, 1)
    e = Engine(m, m, 1, 1)
    dH, dg, dl, du, dx0 = (upload(e, Array{Float64}(v)) for v in (H, g, lower, upper, x0))
    x = zeros(m); res = Int32[0]; Hf = zeros(m, m); fm = UInt32[0]; nf = Int32[0]
    o = (dmalloc(e, 8m), dmalloc(e, 4), dmalloc(e, 8m * m), dmalloc(e, 4), dmalloc(e, 4))
    opts = DdpBoxQPOpts(maxIter, minGrad, minRelImprove, stepDec, minStep, Armijo)
    check(e, ccall((:ddp_boxqp_f64, libddp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
        Ref{DdpBoxQPOpts}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), e.h, 1, dH, dg, dl, du, dx0, opts, o...))
    check(e, ccall((:ddp_synchronize, libddp), Cint, (Ptr{Cvoid},), e.h))
    download!(e, x, o[1]); download!(e, res, o[2]); download!(e, Hf, o[3]); download!(e, fm, o[4]); download!(e, nf, o[5])
    res[1] < 0 && throw(PosDefException(-1))                                   # where the reference's cholesky throws
    free = [((fm[1] >> (i - 1)) & 1) == 1 for i in 1:m]
    nfree = count(!iszero, diag(Hf))
    return x, Int(res[1]), Hf[1:nfree, 1:nfree], free, Int(nf[1])
end

# ---- iLQG(f,costfun,df,x0,u0; kw...)   iLQG.jl:143 ---------------------------------------------
# `f`, `costfun`, `df` must all be the same DeviceModel: arbitrary closures cannot run on the GPU
# and there is no CPU fallback.
function iLQG(f::DeviceModel, costfun::DeviceModel, df::DeviceModel, x0, u0;
              lims = [], α = exp10.(range(0, stop = -3, length = 11)), tol_fun = 1e-7, tol_grad = 1e-4, max_iter = 500,
              λ = 1.0, dλ = 1.0, λfactor = 1.6, λmax = 1e10, λmin = 1e-6, regType = 1, reduce_ratio_min = 0, verbosity = 0, kwargs...)
    f === costfun === df || error("f, costfun and df must be one device model descriptor")
    model = f
    n = size(x0, 1); m, N = size(u0, 1), size(u0, 2); B = size(u0, 3)
    e = Engine(n, m, N, B)
    keep = Ptr{Cvoid}[]
    up(a) = (p = upload(e, Array{Float64}(a)); push!(keep, p); p)
    Q, R = model.Q, model.R
    qdiag = Int32(isdiag(Q) ? 1 : 0)
    md = if model isa LinearModel
        A, Bm = model.A, model.B
        DdpModel(1, tensor(up(A), A, N, B), tensor(up(Bm), Bm, N, B), DdpTensor(up(Q), 0, 0), DdpTensor(up(R), 0, 0), C_NULL,
                 ntuple(_ -> 0.0, 8), 0, qdiag)
    else
        DdpModel(2, DdpTensor(), DdpTensor(), DdpTensor(up(Q), 0, 0), DdpTensor(up(R), 0, 0), Ptr{Float64}(up(model.goal)),
                 (model.g, model.l, model.h, model.d, 0.0, 0.0, 0.0, 0.0), 1, qdiag)
    end
    limsd = isempty(lims) ? C_NULL : up(lims)
    αt = ntuple(i -> i <= length(α) ? Float64(α[i]) : 0.0, 16)
    opts = DdpIlqgOpts(length(α), αt, tol_fun, tol_grad, max_iter, λ, dλ, λfactor, λmax, λmin, regType, reduce_ratio_min, Ptr{Float64}(limsd))
    x0b = repeat(reshape(Array{Float64}(x0)[:, 1, :], n, :), 1, size(x0, 3) == B ? 1 : B)
    x = zeros(n, N, B); u = zeros(m, N, B); K = zeros(m, n, N, B); k = zeros(m, N, B); Vx = zeros(n, N, B); Vxx1 = zeros(n, n, B)
    st = Vector{DdpIlqgState}(undef, B)
    dptr = [dmalloc(e, sizeof(a)) for a in (x, u, K, k, Vx, Vxx1, st)]
    nouter = Ref{Int32}(0)
    check(e, ccall((:ddp_ilqg_solve_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpModel}, Ref{DdpIlqgOpts}, Ptr{Cvoid}, Ptr{Cvoid},
        Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Int32}),
        e.h, md, opts, up(x0b), up(u0), dptr..., nouter))
    for (a, p) in zip((x, u, K, k, Vx, Vxx1, st), dptr); download!(e, a, p); dfree(e, p); end
    foreach(p -> dfree(e, p), keep)
    if B == 1
        st[1].status == 4 && return nothing                                                     # iLQG.jl:205-210
        st[1].iter == 1 && error("Failure: no iterations completed, something is wrong.")       # iLQG.jl:335
        return x[:, :, 1], u[:, :, 1], GaussianPolicy(N, n, m, K[:, :, :, 1], k[:, :, 1], zeros(m, m, 0), zeros(m, m, 0)), Vx[:, :, 1], Vxx1[:, :, 1], st[1].cost, st
    end
    return x, u, (K, k), Vx, Vxx1, [s.cost for s in st], st
end

# ---- multi-GPU: one process per GPU, the batch sharded by contiguous ranges; the only exchange is the statistics all-reduce
comm_unique_id() = (id = zeros(UInt8, 128); ccall((:ddp_comm_unique_id, libddp), Cint, (Ptr{UInt8},), id) == 0 || error("libddp: NCCL not available"); id)
comm_init(e::Engine, nranks, rank, id::Vector{UInt8}) = check(e, ccall((:ddp_comm_init, libddp), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), e.h, nranks, rank, id))
allreduce_stats!(e::Engine, stats8_dev::Ptr{Cvoid}) = check(e, ccall((:ddp_comm_allreduce_stats_f64, libddp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.h, stats8_dev))

# ---- iLQGkl(dynamics,costfun,derivs,x0,traj_prev,model; kw...)   iLQGkl.jl:25 --------------------
# Whole outer loop on the device (ddp_ilqgkl_solve_f64).  `model` of the reference (LinearTimeVaryingModelsBase) is
# replaced by what it is used for: `fx_model` = df(model,x,u)[1] and `R1` = covariance(model,x,u) (forward_pass.jl:38,42).
# x0 is the pre-rolled trajectory (n,N[,B]); traj_prev = (K (m,n,N[,B]), k (m,N[,B]), Σ (m,m,N[,B]), Σi (m,m,N[,B])).
function iLQGkl(dynamics::DeviceModel, costfun::DeviceModel, derivs::DeviceModel, x0, traj_prev, fx_model, R1;
                cost = nothing, kl_step = 1.0, lims = [], max_iter = 50, ηbracket = [1e-8, 1.0, 1e16], del0 = 1e-4, kwargs...)
    dynamics === costfun === derivs || error("dynamics, costfun and derivs must be one device model descriptor")
    cost === nothing && error("Initial trajectory supplied, initial cost must also be supplied")         # iLQGkl.jl:66-68
    model = dynamics
    Kp, kp, Σp, Σip = traj_prev
    n, N = size(x0, 1), size(x0, 2); m = size(kp, 1); B = size(x0, 3)
    e = Engine(n, m, N, B)
    keep = Ptr{Cvoid}[]
    up(a) = (p = upload(e, Array{Float64}(a)); push!(keep, p); p)
    Q, R = model.Q, model.R
    qdiag = Int32(isdiag(Q) ? 1 : 0)
    md = if model isa LinearModel
        DdpModel(1, tensor(up(model.A), model.A, N, B), tensor(up(model.B), model.B, N, B), DdpTensor(up(Q), 0, 0), DdpTensor(up(R), 0, 0),
                 C_NULL, ntuple(_ -> 0.0, 8), 0, qdiag)
    else
        DdpModel(2, DdpTensor(), DdpTensor(), DdpTensor(up(Q), 0, 0), DdpTensor(up(R), 0, 0), Ptr{Float64}(up(model.goal)),
                 (model.g, model.l, model.h, model.d, 0.0, 0.0, 0.0, 0.0), 1, qdiag)
    end
    limsd = isempty(lims) ? C_NULL : up(lims)
    opts = DdpIlqgklOpts(kl_step, max_iter, (ηbracket[1], ηbracket[2], ηbracket[3]), del0, 200, Ptr{Float64}(limsd))
    costs = B == 1 ? [sum(cost)] : vec(sum(reshape(Array{Float64}(cost), :, B), dims = 1))
    xnew = zeros(n, N, B); unew = zeros(m, N, B); K = zeros(m, n, N, B); k = zeros(m, N, B); Σ = zeros(m, m, N, B); Σi = zeros(m, m, N, B)
    Vx = zeros(n, N, B); Vxx1 = zeros(n, n, B); cnew = zeros(B); st = Vector{DdpIlqgklState}(undef, B)
    outs = (xnew, unew, K, k, Σ, Σi, Vx, Vxx1, cnew, st)
    dptr = [dmalloc(e, sizeof(a)) for a in outs]
    args = DdpIlqgklArgs(up(x0), up(kp), up(costs),                                                       # u = traj_prev.k (iLQGkl.jl:47)
                         tensor(up(Kp), Kp, N, B), tensor(up(Σp), Σp, N, B), tensor(up(Σip), Σip, N, B),
                         tensor(up(fx_model), fx_model, N, B), tensor(up(R1), R1, 1, B), dptr...)
    nouter = Ref{Int32}(0)
    check(e, ccall((:ddp_ilqgkl_solve_f64, libddp), Cint, (Ptr{Cvoid}, Ref{DdpModel}, Ref{DdpIlqgklOpts}, Ref{DdpIlqgklArgs}, Ref{Int32}),
                   e.h, md, opts, args, nouter))
    for (a, p) in zip(outs, dptr); download!(e, a, p); dfree(e, p); end
    foreach(p -> dfree(e, p), keep)
    B == 1 && return xnew[:, :, 1], unew[:, :, 1], GaussianPolicy(N, n, m, K[:, :, :, 1], k[:, :, 1], Σ[:, :, :, 1], Σi[:, :, :, 1]),
                     Vx[:, :, 1], Vxx1[:, :, 1], cnew[1], st
    return xnew, unew, (K, k, Σ, Σi), Vx, Vxx1, cnew, st
end

iLQG(f, costfun, df, x0, u0; kwargs...) = error("iLQG: f/costfun/df must be a device model descriptor (LinearModel, PendcartModel); " *
                                                "arbitrary Julia closures cannot run on the GPU and this package has no CPU fallback")

end # module
