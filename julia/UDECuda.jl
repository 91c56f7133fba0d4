# julia/UDECuda.jl -- the reference-side binding of libude_cuda.so (include/ude_cuda.h).
#
# Drop next to src/Optimizers.jl of UniversalDiffEq.jl and `include("UDECuda.jl")` from src/UniversalDiffEq.jl.
# NOT EXECUTED in this repository's build image (no Julia there); it is kept mechanical -- plain `ccall`s and field copies --
# and mirrors, call for call, universaldiffeq.jl_b200/_capi.py (ctypes), which IS exercised by tests/test_gpu_*.py.
#
# What it replaces (reference file:line):
#   * the loss closures of src/LossFunctionConstructors.jl (`conditional_likelihood` :17-98 / :347-404, `multiple_shooting_loss`
#     :261-305 / :639-708, `shooting_loss` :196-232 / :545-601, `derivative_matching_loss` :175-192 / :500-540) and the
#     Zygote pullback through `OrdinaryDiffEq.solve(...; sensealg = ForwardDiffSensitivity())` (src/ProcessModels.jl:80-86,
#     src/MultiProcessModel.jl:81-87): one `ude_loss_grad` call;
#   * `Optimization.OptimizationFunction((x,p) -> loss(x), AutoZygote())` in `train!` (src/Optimizers.jl:473-478, :658-663):
#     `OptimizationFunction(f; grad = g!)` built by `cuda_objective` below;
#   * selected by the solver marker stored in `UDE.solvers.ode` (src/Models.jl:115-116, :148): `ode_solver = CudaRK4()` /
#     `CudaTsit5()` on any continuous constructor.
module UDECuda

using ComponentArrays
using LinearAlgebra

const lib = "libude_cuda"                       # on LD_LIBRARY_PATH (universaldiffeq.jl_b200/csrc/libude_cuda.so)

# ---- solver markers: `NODE(data; ode_solver = CudaRK4(4))` -----------------------------------------------------------------
struct CudaRK4;   substeps::Int; end
struct CudaTsit5; substeps::Int; end
CudaRK4() = CudaRK4(4)
CudaTsit5() = CudaTsit5(4)
is_cuda_solver(s) = s isa CudaRK4 || s isa CudaTsit5

# ---- enums of include/ude_cuda.h -------------------------------------------------------------------------------------------
const UDE_RK4, UDE_TSIT5, UDE_EULER = Int32(0), Int32(1), Int32(2)
const LOSS_STATE_SPACE, LOSS_MULTI_SHOOT, LOSS_SHOOT, LOSS_DERIV_MATCH = Int32(0), Int32(1), Int32(2), Int32(3)
const RHS_NODE, RHS_NODE_TIME, RHS_LV_UDE, RHS_LV_UDE_ABS, RHS_FISHERY, RHS_NN_DECAY, RHS_SERIES_GROWTH = Int32.(0:6)
const OPT_SPARSE_UHAT_IO, OPT_UKF_CAUSAL_INTERVAL = Int32(1), Int32(2)

# mirror of `ude_desc`; isbits, passed by reference
struct Desc
    abi_version::Int32; struct_bytes::Int32; device::Int32; dtype::Int32
    solver::Int32; substeps::Int32
    rhs_kind::Int32; d::Int32; n_cov::Int32; nn_in::Int32; nn_hidden::Int32; nn_out::Int32; n_scalars::Int32; has_series_param::Int32
    off_w1::Int64; off_b1::Int64; off_w2::Int64; off_b2::Int64
    off_scalar::NTuple{8,Int64}; off_series_param::Int64; n_pm::Int64
    time_scale::Float64; time_shift::Float64
    loss_kind::Int32; pred_length::Int32; proc_inv_dt::Int32; shoot_from_theta::Int32; shoot_first_scored::Int32; remove_ends::Int32
    preroll_eps::Float64
    A::NTuple{64,Float64}; B::NTuple{64,Float64}
    lanes_per_segment::Int32; hidden_per_lane::Int32; weights_in_smem::Int32; reserved::NTuple{5,Int32}
end

mutable struct Handle; ptr::Ptr{Cvoid}; end
check(h, rc) = (rc == 0 || rc == -7) ? rc :                 # -7 = UDE_NONFINITE: loss = Inf/NaN is still returned, like the reference
    error(unsafe_string(ccall((:ude_last_error, lib), Cstring, (Ptr{Cvoid},), h === nothing ? C_NULL : h.ptr)))

function create(desc::Desc)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(nothing, ccall((:ude_create, lib), Cint, (Ref{Desc}, Ref{Ptr{Cvoid}}), desc, out))
    h = Handle(out[]); finalizer(x -> ccall((:ude_destroy, lib), Cint, (Ptr{Cvoid},), x.ptr), h); h
end

# data :: d×N Matrix{Float64}, times :: Vector{Float64}, starts/lengths as returned by process_multi_data (1-based)
set_data!(h, data, times, starts, lengths, model_of_series = nothing) = check(h, ccall((:ude_set_data, lib), Cint,
    (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Float64}),
    h.ptr, length(starts), Int64.(starts .- 1), Int64.(lengths),
    model_of_series === nothing ? C_NULL : Int64.(model_of_series .- 1), length(times), times, data))
set_model_weights!(h, obs, proc, shoot, reg) = check(h, ccall((:ude_set_model_weights, lib), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h.ptr, obs, proc, shoot, reg))
set_covariates!(h, knot_off, knot_t, knot_x) = check(h, ccall((:ude_set_covariates, lib), Cint,
    (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}), h.ptr, knot_off, knot_t, knot_x))
set_targets!(h, uhat, dudt) = check(h, ccall((:ude_set_targets, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h.ptr, uhat, dudt))
set_skip!(h, mask::Vector{UInt8}) = check(h, ccall((:ude_set_skip, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h.ptr, mask))
set_option!(h, option, value) = check(h, ccall((:ude_set_option, lib), Cint, (Ptr{Cvoid}, Int32, Int64), h.ptr, option, value))

# loss(θ) and ∇loss(θ): θ is the flat ComponentVector data, getdata(parameters) (src/helpers.jl:3-11)
function loss_grad!(G, h, θ::Vector{Float64})
    L = Ref{Float64}(0.0)
    check(h, ccall((:ude_loss_grad, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Float64}, Ptr{Float64}),
                   h.ptr, θ, length(θ), L, G === nothing ? C_NULL : G))
    L[]
end
forward(h, θ, d, N) = (out = zeros(d, N); check(h, ccall((:ude_forward, lib), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), h.ptr, θ, length(θ), out)); out)
adam!(h, θ, η, iters) = (tr = zeros(iters); check(h, ccall((:ude_adam, lib), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Int32, Ptr{Float64}), h.ptr, θ, length(θ), η, iters, tr)); tr)

# ---- desc_from: everything the reference's loss constructors read, copied into a Desc --------------------------------------
pad64(M) = ntuple(i -> i <= length(M) ? Float64(M[i]) : 0.0, 64)          # column-major d×d into the 8×8 slot
pad8(v) = ntuple(i -> i <= length(v) ? Int64(v[i]) : Int64(0), 8)
# errors_to_matrix (LFC.jl:4-15): scalar -> e*I, vector -> diag(e), matrix -> itself
function errors_to_matrix(e, d)
    e isa Number && return Matrix{Float64}(e * I(d))
    ndims(e) == 1 && return Matrix{Float64}(Diagonal(e))
    return Matrix{Float64}(e)
end

"""
    desc_from(UDE, loss_function, loss_options; rhs_kind = RHS_NODE, scalar_names = (), time_scale = 1.0, time_shift = 0.0, device = 0)

`UDE` is a `UDE` or `MultiUDE` built by NODE / MultiNODE (rhs_kind = RHS_NODE) or by CustomDerivatives with a right-hand
side of the catalog (pass its `rhs_kind` and, in parameter order, the names of its scalar parameters).  `loss_function` and
`loss_options` are `train!`'s (src/Optimizers.jl:352-471).  Offsets follow the ComponentArray order
`process_model = (NN = (layer_1 = (weight, bias), layer_2 = (weight, bias)), scalars...)` (SURVEY App. B).
"""
function desc_from(UDE, loss_function::String, loss_options::NamedTuple; rhs_kind = RHS_NODE, scalar_names = (),
                   series_param = nothing, time_scale = 1.0, time_shift = 0.0, device = 0)
    d = size(UDE.data, 1)
    pm = UDE.parameters.process_model
    W1, W2 = pm.NN.layer_1.weight, pm.NN.layer_2.weight
    H, din = size(W1); dout = size(W2, 1)
    multi = hasproperty(UDE, :series_column_name)
    n_cov = UDE.X === nothing || UDE.X == 0 ? 0 : length(multi ? UDE.X(0.0, 1) : UDE.X(0.0))
    off_w1 = 0; off_b1 = off_w1 + H * din; off_w2 = off_b1 + H; off_b2 = off_w2 + dout * H
    o = off_b2 + dout
    offs = Int64[]
    for _ in scalar_names; push!(offs, o); o += 1; end
    off_sp = -1
    if series_param !== nothing
        off_sp = o; o += length(getproperty(pm, series_param))
    end
    solver = UDE.solvers.ode
    A = Matrix{Float64}(I(d)); B = Matrix{Float64}(I(d))
    loss_kind = LOSS_STATE_SPACE; pl = 5; inv_dt = 0; from_theta = 1; first_scored = 2; remove_ends = 0; eps = 0.0
    if loss_function == "conditional likelihood"                       # src/Optimizers.jl:365-390
        A = inv(errors_to_matrix(get(loss_options, :observation_error, 0.025), d))
        B = inv(errors_to_matrix(get(loss_options, :process_error, 0.025), d))
    elseif loss_function == "multiple shooting"                        # :444-455
        loss_kind = LOSS_MULTI_SHOOT; pl = get(loss_options, :pred_length, 5)
        eps = multi ? 0.0 : 1e-6                                       # LFC.jl:268 vs :646
    elseif loss_function == "shooting"                                 # :456-465
        loss_kind = LOSS_SHOOT; from_theta = multi ? 0 : 1             # LFC.jl:563: the multi variant starts from the data
    elseif loss_function == "derivative matching"                      # :424-443
        loss_kind = LOSS_DERIV_MATCH; remove_ends = get(loss_options, :remove_ends, 0)
    elseif loss_function == "default"                                  # UDE.loss_function, src/helpers.jl:18-44
        inv_dt = 1
    else
        error("loss_function \"$loss_function\" has no libude_cuda path (marginal likelihood: ude_ukf_loss_grad)")
    end
    Desc(1, sizeof(Desc), device, 0,
         solver isa CudaTsit5 ? UDE_TSIT5 : UDE_RK4, solver.substeps,
         rhs_kind, d, n_cov, din, H, dout, length(scalar_names), series_param === nothing ? 0 : 1,
         off_w1, off_b1, off_w2, off_b2, pad8(offs), off_sp, o, time_scale, time_shift,
         loss_kind, pl, inv_dt, from_theta, first_scored, remove_ends, eps,
         pad64(vec(A)), pad64(vec(B)),                 # d x d column-major in the first d*d entries
         0, 0, 0, ntuple(_ -> Int32(0), 5))
end

"""
    cuda_objective(UDE, loss_function, loss_options, regularization_weight; kw...) -> (f, g!, handle)

The two closures `train!` hands to `Optimization.OptimizationFunction(f; grad = g!)` instead of
`OptimizationFunction((x,p) -> loss(x), AutoZygote())` (src/Optimizers.jl:475-478, :660-663).
"""
function cuda_objective(UDE, loss_function, loss_options, regularization_weight; starts = nothing, lengths = nothing, kw...)
    h = create(desc_from(UDE, loss_function, loss_options; kw...))
    N = size(UDE.data, 2)
    starts === nothing && (starts = [1]; lengths = [N])
    set_data!(h, Matrix{Float64}(UDE.data), Vector{Float64}(UDE.times), starts, lengths)
    reg = [regularization_weight * UDE.process_regularization.reg_parameters.weight]   # the product, as LFC.jl:48 does
    set_model_weights!(h, C_NULL, C_NULL, C_NULL, reg)
    f = (x, p) -> loss_grad!(nothing, h, collect(getdata(x)))
    g! = (G, x, p) -> (loss_grad!(getdata(G), h, collect(getdata(x))); G)
    return f, g!, h
end

# ---- unscented-Kalman-filter marginal likelihood (src/UnscentedKalmanFilter.jl:76-87, :59-72; LFC.jl:103-127, :427-476) --------------
# mirror of `ude_ukf_opts` (include/ude_cuda.h)
struct UkfOpts
    alpha::Float64; beta::Float64; kappa::Float64; reg::Float64
    Peta::NTuple{64,Float64}
end
UkfOpts(Pη::AbstractMatrix; α = 1e-3, β = 2.0, κ = 0.0, reg = 0.0) = UkfOpts(α, β, κ, reg, pad64(Matrix{Float64}(Pη)))

# F: d×d factor of the process noise (Pν = F F'), or d×d×n_models on a many-model handle (folds × ensemble members);
# returns (loss per model, ∇θ, ∇F).  `train!(…; loss_function = "marginal likelihood")` optimises (process_model, F) with these.
function ukf_loss_grad(h, opts::UkfOpts, θ::Vector{Float64}, F::Array{Float64}, n_models = 1)
    L = zeros(n_models); Gθ = zeros(length(θ)); GF = zeros(size(F))
    check(h, ccall((:ude_ukf_loss_grad, lib), Cint,
                   (Ptr{Cvoid}, Ref{UkfOpts}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   h.ptr, Ref(opts), θ, length(θ), F, L, Gθ, GF))
    L, Gθ, GF
end
# ukf_smoothing: filtered states d×N and covariances d×d×N
function ukf_filter(h, opts::UkfOpts, θ::Vector{Float64}, F::Array{Float64}, d, N)
    xs = zeros(d, N); Ps = zeros(d, d, N)
    check(h, ccall((:ude_ukf_filter, lib), Cint,
                   (Ptr{Cvoid}, Ref{UkfOpts}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   h.ptr, Ref(opts), θ, length(θ), F, xs, Ps))
    xs, Ps
end

# ---- derivative-matching pre-step (LFC.jl:161-171, :510-522): RegularizationSmooth(u, t, d; alg = :gcv_svd) for all rows at once --
# U: rows×n (row-major on the C side: pass permutedims of an n×rows Julia matrix), Vt (r×n), s (r), M (n×n) from the time grid
function reg_smooth(Ut::Matrix{Float64}, Vt_t::Matrix{Float64}, s::Vector{Float64}, Mt::Matrix{Float64}; device = 0, λlo = 1e-4, λhi = 1e3)
    n, rows = size(Ut); r = length(s)
    u = zeros(n, rows); du = zeros(n, rows); λ = zeros(rows)
    rc = ccall((:ude_reg_smooth, lib), Cint,
               (Int32, Int32, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
               device, n, r, rows, Vt_t, s, Mt, Ut, λlo, λhi, u, du, λ)
    rc == 0 || error("ude_reg_smooth failed with status $rc")
    u, du, λ
end

end # module

# ---- the hook in train! (src/Optimizers.jl:473-478), verbatim --------------------------------------------------------------
#
#   if UDECuda.is_cuda_solver(UDE.solvers.ode)
#       f, g!, h = UDECuda.cuda_objective(UDE, loss_function, loss_options, regularization_weight)
#       optf = Optimization.OptimizationFunction(f; grad = g!)
#   else
#       optf = Optimization.OptimizationFunction((x, p) -> loss(x), Optimization.AutoZygote())
#   end
#   optprob = Optimization.OptimizationProblem(optf, params)
#   sol = Optimization.solve(optprob, OptimizationOptimisers.Adam(step_size); callback, maxiters)
