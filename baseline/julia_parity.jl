# baseline/julia_parity.jl -- pins the CPU oracle (and through it the CUDA path) to the REFERENCE, where Julia exists.
#
# NOT EXECUTED IN THIS REPOSITORY'S BUILD IMAGE (no julia binary, no depot, no network): parity stays "unpinned" until someone
# runs this once.  Usage, from the repository root, on a machine with Julia >= 1.10 and UniversalDiffEq.jl instantiated:
#
#     julia --project=/path/to/UniversalDiffEq.jl baseline/julia_parity.jl tests/golden/julia_cases tests/golden/julia_parity.json
#
# and commit tests/golden/julia_parity.json; tests/test_julia_parity.py consumes it when present (and skips otherwise).
#
# What it does, per case directory written by tests/golden/make_julia_inputs.py (identical inputs on both sides):
#   * builds the REFERENCE model object (NODE / NODE with covariates / MultiNODE with covariates: src/Models.jl:746-840,
#     src/MultipleTimeSeries.jl:241-290), i.e. the reference's own process_data / process_multi_data, interpolate_covariates,
#     Lux network, parameter ComponentArray and right-hand side;
#   * overwrites its parameters with the case's theta (uhat, layer_1.weight, layer_1.bias, layer_2.weight, layer_2.bias);
#   * evaluates the loss named in meta.txt with the reference's formulas (conditional likelihood LFC.jl:17-53, :347-404;
#     multiple shooting :261-305, :639-708; shooting :196-232, :545-601) but with a FIXED step sequence
#     (solve(...; adaptive = false, dt = interval / substeps), RK4() or Tsit5()) -- the discretisation the CUDA kernels and the
#     oracle use (DESIGN.md 1) -- through the reference's process_model.rhs closure and OrdinaryDiffEq's own steppers;
#   * differentiates that loss w.r.t. the flat theta with ForwardDiff (the derivative of the same discrete map that
#     Zygote + ForwardDiffSensitivity return for a fixed step sequence, SURVEY App. C.2);
#   * additionally evaluates the reference's OWN loss closure (adaptive Tsit5, abstol = reltol = 1e-6) where the package
#     exports one for the loss, so that the fixed-step value can be seen to sit inside the reference's solver tolerance.
using UniversalDiffEq, DataFrames, CSV, OrdinaryDiffEq, ComponentArrays, ForwardDiff, LinearAlgebra, Printf

readmeta(path) = Dict(strip(split(l, "=")[1]) => strip(split(l, "=")[2]) for l in eachline(path) if occursin("=", l))
readvec(path) = [parse(Float64, l) for l in eachline(path) if !isempty(strip(l))]

function build_model(dir, meta)
    data = CSV.read(joinpath(dir, "data.csv"), DataFrame)
    H = parse(Int, meta["hidden"])
    kind = meta["model"]
    if kind == "NODE"
        return NODE(data; hidden_units = H, seed = 1)
    elseif kind == "NODE_X"
        X = CSV.read(joinpath(dir, "covariates.csv"), DataFrame)
        return NODE(data, X; hidden_units = H, seed = 1)
    elseif kind == "MultiNODE_X"
        X = CSV.read(joinpath(dir, "covariates.csv"), DataFrame)
        return MultiNODE(data, X; hidden_units = H, seed = 1)
    end
    error("unknown model kind $kind")
end

# theta layout on the wire (SURVEY App. B): [uhat(:) | W1(:) | b1 | W2(:) | b2], all column-major
function set_theta!(p, theta, d, N, H, din)
    o = 0
    p.uhat .= reshape(theta[o+1:o+d*N], d, N); o += d * N
    p.process_model.NN.layer_1.weight .= reshape(theta[o+1:o+H*din], H, din); o += H * din
    p.process_model.NN.layer_1.bias .= theta[o+1:o+H]; o += H
    p.process_model.NN.layer_2.weight .= reshape(theta[o+1:o+d*H], d, H); o += d * H
    p.process_model.NN.layer_2.bias .= theta[o+1:o+d]; o += d
    @assert o == length(theta)
    return p
end

function flat_theta(p)
    vcat(vec(p.uhat), vec(p.process_model.NN.layer_1.weight), vec(p.process_model.NN.layer_1.bias),
         vec(p.process_model.NN.layer_2.weight), vec(p.process_model.NN.layer_2.bias))
end

# one observation interval with `S` fixed steps of the chosen stepper, through the reference's right-hand side
function make_flow(model, multi, has_cov, solver, S)
    pmod = model.process_model
    function rhs(u, pm, t, i)
        if multi
            X = pmod.covariates(t, i)
            return pmod.right_hand_side(u, i, X, t, pm)
        elseif has_cov
            return pmod.right_hand_side(u, pmod.covariates(t), pm, t)
        else
            return pmod.rhs(u, pm, t)
        end
    end
    alg = solver == "tsit5" ? Tsit5() : RK4()
    function flow(u, t0, t1, pm, i)
        if t1 == t0
            return u
        end
        prob = ODEProblem((uu, pp, tt) -> rhs(uu, pp, tt, i), u, (t0, t1), pm)
        sol = solve(prob, alg; adaptive = false, dt = (t1 - t0) / S, save_everystep = false, save_start = false)
        return sol.u[end]
    end
    return flow
end

function fixed_step_loss(model, meta, multi, has_cov)
    S = parse(Int, meta["substeps"]); pl = parse(Int, meta["pred_length"]); lossname = meta["loss"]
    flow = make_flow(model, multi, has_cov, meta["solver"], S)
    data, times = model.data, model.times
    d, N = size(data)
    if multi
        _, _, _, _, _, _, _, _, starts, lengths, _, _ = UniversalDiffEq.process_multi_data(model.data_frame, model.time_column_name, model.series_column_name)
    else
        starts, lengths = [1], [N]
    end
    function loss(p)
        L = zero(eltype(p))
        pm = p.process_model
        for s in eachindex(starts)
            cols = starts[s]:(starts[s] + lengths[s] - 1)
            tt = times[cols]; y = data[:, cols]; uh = p.uhat[:, cols]; n = length(cols)
            if lossname == "conditional likelihood"
                w_o = 1 / parse(Float64, meta["observation_error"]); w_p = 1 / parse(Float64, meta["process_error"])
                for t in 1:n
                    L += w_o * sum((y[:, t] .- uh[:, t]) .^ 2)            # errors_to_matrix(scalar) = e * I, inverse 1/e
                end
                for t in 2:n
                    L += w_p * sum((uh[:, t] .- flow(uh[:, t-1], tt[t-1], tt[t], pm, s)) .^ 2)
                end
            elseif lossname == "multiple shooting"
                for t in 1:pl:n
                    last = (n - t) >= pl ? t + pl : n
                    u = uh[:, t]
                    if !multi
                        u = flow(u, tt[t] - 1e-6, tt[t], pm, s)            # single-series variant starts 1e-6 early (LFC.jl:268)
                    end
                    for i in t:last
                        L += sum((y[:, i] .- u) .^ 2) / (d * n)
                        if i < last
                            u = flow(u, tt[i], tt[i+1], pm, s)
                        end
                    end
                end
            elseif lossname == "shooting"
                u = multi ? y[:, 1] : uh[:, 1]
                for i in 1:n
                    if i >= 3                                              # the first predicted point is skipped (LFC.jl:220-222, :572-574)
                        L += sum((y[:, i] .- u) .^ 2) * (multi ? 1.0 : parse(Float64, meta["shoot_weight"]))
                    end
                    if i < n
                        u = flow(u, tt[i], tt[i+1], pm, s)
                    end
                end
            else
                error("loss $lossname not covered")
            end
        end
        return L
    end
    return loss
end

function reference_adaptive_loss(model, meta, multi)
    # the package's own closures (adaptive Tsit5 at 1e-6): the names are internal, so failures are reported, not fatal
    try
        if meta["loss"] == "multiple shooting"
            f, _, _ = multi ? UniversalDiffEq.multiple_shooting_loss(model, 0.0, parse(Int, meta["pred_length"])) :
                              UniversalDiffEq.multiple_shooting_loss(model, 0.0, parse(Int, meta["pred_length"]))
            return f(model.parameters)
        elseif meta["loss"] == "conditional likelihood"
            e_o = parse(Float64, meta["observation_error"]); e_p = parse(Float64, meta["process_error"])
            f, _, _ = UniversalDiffEq.conditional_likelihood(model, 0.0, e_o, e_p)
            return f(model.parameters)
        end
    catch err
        @warn "reference closure not evaluated" err
    end
    return NaN
end

function main(indir, outfile)
    open(outfile, "w") do io
        println(io, "{")
        dirs = sort(filter(d -> isdir(joinpath(indir, d)), readdir(indir)))
        for (k, name) in enumerate(dirs)
            dir = joinpath(indir, name)
            meta = readmeta(joinpath(dir, "meta.txt"))
            model = build_model(dir, meta)
            multi = meta["model"] == "MultiNODE_X"; has_cov = meta["model"] != "NODE"
            d, N = size(model.data)
            H = parse(Int, meta["hidden"]); din = parse(Int, meta["nn_in"])
            theta = readvec(joinpath(dir, "theta.txt"))
            p = set_theta!(copy(model.parameters), theta, d, N, H, din)
            model.parameters = p
            loss = fixed_step_loss(model, meta, multi, has_cov)
            L = loss(p)
            ax = getaxes(p)
            g = ForwardDiff.gradient(v -> loss(ComponentArray(v, ax)), collect(p))
            gflat = flat_theta(ComponentArray(g, ax))
            Lad = reference_adaptive_loss(model, meta, multi)
            @printf(io, "  \"%s\": {\"loss\": %.17g, \"reference_adaptive_loss\": %s, \"grad\": [%s]}%s\n", name, L,
                    isnan(Lad) ? "null" : @sprintf("%.17g", Lad), join((@sprintf("%.17g", x) for x in gflat), ", "),
                    k < length(dirs) ? "," : "")
            @printf("%-40s loss %.12g (adaptive reference closure: %s)\n", name, L, string(Lad))
        end
        println(io, "}")
    end
end

main(ARGS[1], ARGS[2])
