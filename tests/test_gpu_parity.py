"""Parity of the CUDA hot path (through the C ABI, host buffers) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): loss and gradient within 1e-10 relative in FP64.  The oracle and the kernels share
the discretisation (fixed-step RK4 / Tsit5, `substeps` per observation interval) but no code; they differ in
summation order, FMA contraction, the tanh implementation (libm vs ude_tanh_tab, <= 2 ulp apart) and the covariate
interpolation (division vs precomputed reciprocal), all O(1e-16) effects.
"""
import numpy as np
import pytest

import cases
from oracle import oracle

pytestmark = pytest.mark.gpu

TOL = 1e-10   # relative, FP64 (north_star)


def _check(prob, theta, ude, **kw):
    L_ref, g_ref = oracle.loss_grad(prob, theta)
    with ude.Engine(prob, **kw) as eng:
        L, g = eng.loss_grad(theta)
        L2, _ = eng.loss_grad(theta, want_grad=False)
        assert eng.steps_per_eval() == prob.count_steps() == oracle.count_steps(prob)
    assert np.all(np.isfinite(L)) and np.all(np.isfinite(g))
    assert np.max(np.abs(L - L_ref) / np.abs(L_ref)) < TOL, (L, L_ref)
    assert np.max(np.abs(L2 - L_ref) / np.abs(L_ref)) < TOL, (L2, L_ref)
    # gradient: element-wise per leaf group (uhat, W1, b1, W2, b2, scalars, per-series vector) -- a norm over the whole
    # vector would let a small leaf be wrong; entries that are structurally zero must be exactly zero
    worst = cases.grad_mismatch(prob, g, g_ref, rtol=TOL)
    assert worst[0] <= 1.0, worst
    assert np.all(g[g_ref == 0.0] == 0.0)


@pytest.mark.parametrize("solver", [0, 1], ids=["rk4", "tsit5"])
@pytest.mark.parametrize("loss", cases.LOSSES)
@pytest.mark.parametrize("rhs", list(cases.RHS_SPECS))
def test_parity_all_kinds(ude, rhs, loss, solver):
    prob, theta = cases.make_case(rhs, loss, solver, skip=(loss in ("cond_lik", "mse")))
    _check(prob, theta, ude)


@pytest.mark.parametrize("loss", ["cond_lik", "mse", "multi_shoot", "shoot_theta", "deriv_match"])
def test_parity_many_models(ude, loss):
    """CV folds x ensemble members: independent theta vectors in one handle (src/CrossValidation.jl:17-45)."""
    prob, theta = cases.make_case("node_time_d3", loss, 0, lengths=(9, 8, 7, 6, 5, 4, 11, 3), n_models=8)
    _check(prob, theta, ude)
    prob, theta = cases.make_case("node_d2", loss, 1, lengths=(9, 3, 7, 6, 5, 4), n_models=3)
    _check(prob, theta, ude)


@pytest.mark.parametrize("g_hpl", [(4, 3), (16, 2), (4, 4)])      # (4, 4) = DMMA kernel, hidden padded to 16
@pytest.mark.parametrize("loss", ["multi_shoot", "multi_shoot_preroll", "cond_lik", "shoot_data", "deriv_match"])
def test_parity_group_shapes_node2(ude, g_hpl, loss):
    for solver in (0, 1):
        prob, theta = cases.make_case("node_d2", loss, solver, hidden=10, lengths=(23, 11, 6, 31, 5, 1, 2, 16), skip=(loss == "cond_lik"))
        _check(prob, theta, ude, lanes_per_segment=g_hpl[0], hidden_per_lane=g_hpl[1])


@pytest.mark.parametrize("g_hpl,hidden,ws", [((4, 3), 12, 1), ((16, 2), 32, 1), ((8, 4), 32, 1), ((16, 2), 20, 1), ((8, 4), 7, 1),
                                             ((16, 2), 32, 2), ((8, 4), 32, 2), ((8, 4), 19, 2),
                                             ((4, 8), 32, 0), ((4, 8), 27, 0), ((4, 8), 8, 0)])   # (4, 8) = the FP64 tensor-core (DMMA) kernel
@pytest.mark.parametrize("loss", ["multi_shoot", "mse", "deriv_match"])
def test_parity_group_shapes_c3(ude, g_hpl, hidden, ws, loss):
    """Every (lanes per segment, hidden units per lane, weight storage) specialisation of the C3 shape."""
    for solver in (0, 1):
        prob, theta = cases.make_case("node_d4x1", loss, solver, hidden=hidden, lengths=(30, 30, 7, 12, 1, 30, 26, 3), substeps=4)
        _check(prob, theta, ude, lanes_per_segment=g_hpl[0], hidden_per_lane=g_hpl[1], weights_in_smem=ws)
    if loss == "mse" and g_hpl != (4, 8):    # several models per handle: the per-warp weight records are reloaded at model boundaries
        prob, theta = cases.make_case("node_d4x1", loss, 0, hidden=hidden, lengths=(9, 8, 7, 6, 5, 4), n_models=3)
        _check(prob, theta, ude, lanes_per_segment=g_hpl[0], hidden_per_lane=g_hpl[1], weights_in_smem=ws)


@pytest.mark.parametrize("kernel", ["mma2_h32_nt2", "mma2_h32_nt1"])
@pytest.mark.parametrize("loss", ["multi_shoot", "multi_shoot_preroll", "shoot_theta", "shoot_data", "cond_lik", "mse"])
def test_parity_mma2_family(ude, kernel, loss, monkeypatch):
    """Round-2 tensor-core kernels (ude_kernel_mma2.cuh): 8 / 16 segments per warp, per-interval register prefetch of
    times / data / covariate knots, knot-by-knot covariate advance.  Ragged series, 1-point series, trailing one-point
    blocks, covariate knots that do not cover the span (flat extrapolation both sides), hidden widths below the padded 32."""
    monkeypatch.setenv("UDE_KERNEL", kernel)
    for solver in (0, 1):
        for hidden, lengths, substeps in ((32, (30, 30, 7, 12, 1, 30, 26, 3, 30, 30, 11, 2, 30, 17, 30, 30, 5, 30), 4),
                                          (27, (9, 1, 2, 64, 3), 3), (8, (16,) * 5 + (1,) * 3 + (6,) * 30, 1)):
            prob, theta = cases.make_case("node_d4x1", loss, solver, hidden=hidden, lengths=lengths, substeps=substeps,
                                          skip=(loss in ("cond_lik", "mse")))
            with ude.Engine(prob) as eng:
                assert kernel in eng.kernel_name()
            _check(prob, theta, ude)


@pytest.mark.parametrize("kernel", ["mma2_h32_nt2", "mma2_h32_nt1"])
def test_parity_mma2_dense_knots_and_forward(ude, kernel, monkeypatch):
    """More covariate knots than RK stages per interval (several knot crossings inside one stage advance) and the
    forward-only trajectories of the new kernels."""
    monkeypatch.setenv("UDE_KERNEL", kernel)
    rng = np.random.default_rng(5)
    for loss in ("multi_shoot", "shoot_theta", "cond_lik"):
        prob, theta = cases.make_case("node_d4x1", loss, 0, hidden=32, lengths=(12, 7, 30, 1, 9), substeps=2, skip=False)
        off, kt, kx = [0], [], []
        for s in range(prob.n_series):
            t = prob.times[prob.starts[s]:prob.starts[s] + prob.lengths[s]]
            k = np.sort(rng.uniform(t[0] - 0.05, t[-1] + 0.05, size=40 * max(len(t), 2)))
            kt.append(k); kx.append(rng.normal(size=k.shape[0])); off.append(off[-1] + k.shape[0])
        prob.knot_off = np.array(off); prob.knot_t = np.concatenate(kt); prob.knot_x = np.concatenate(kx)
        prob.finalize()
        _check(prob, theta, ude)
        ref = oracle.forward(prob, theta)
        with ude.Engine(prob) as eng:
            out = eng.forward(theta)
        assert cases.relerr(out, ref) < TOL, loss


@pytest.mark.parametrize("kernel", ["mma_h128", "mma_h128_biasmma", "mma_h128_b2", "g32h4"])
def test_parity_wide_node(ude, kernel, monkeypatch):
    """d = 8, hidden 128 (BASELINE config C5): the DMMA kernel with the hidden layer split over four warps, and the
    lane-per-hidden-unit kernel, against the oracle."""
    monkeypatch.setenv("UDE_KERNEL", kernel)
    prob, theta = cases.make_case("node_d8", "cond_lik", 0, hidden=128, lengths=(2,) * 9 + (5,), substeps=4)
    _check(prob, theta, ude)
    prob, theta = cases.make_case("node_d8", "multi_shoot", 1, hidden=100, lengths=(7, 6, 3, 11, 2, 9, 5, 5, 4), substeps=2)
    _check(prob, theta, ude)
    prob, theta = cases.make_case("node_d8", "deriv_match", 0, hidden=128, lengths=(7, 6, 3, 11))
    _check(prob, theta, ude)
    prob, theta = cases.make_case("node_d8", "mse", 1, hidden=30, lengths=(7, 6, 3, 11, 2), skip=True)
    _check(prob, theta, ude)


@pytest.mark.parametrize("name,kw", [("c1", {}), ("c2", dict(n_series=40)), ("c3", dict(n_series=25)),
                                     ("c3", dict(n_series=25, loss="mse")), ("c4", dict(n_members=3, n_folds=4, n_pts=30)),
                                     ("c5", dict(n_segments=70))])
def test_parity_baseline_configs_reduced(ude, name, kw):
    """The five BASELINE.json configs at sizes the oracle finishes in seconds."""
    for solver in (0, 1):
        prob, theta = ude.synthetic.CONFIGS[name](solver=solver, **kw)
        rng = np.random.default_rng(7)
        theta = theta + 0.01 * rng.normal(size=theta.shape)
        _check(prob, theta, ude)


def test_sparse_uhat_io_zero_copy(ude):
    """UDE_OPT_SPARSE_UHAT_IO: with page-locked host buffers the kernels read the block-start columns of uhat from, and
    write their gradient to, the caller's memory; nothing else of uhat crosses PCIe and the structurally-zero entries of
    the (pre-zeroed) gradient buffer are left alone.  Dense fallbacks: state-space losses, pageable buffers."""
    import torch
    for rhs, hidden in (("node_d4x1", 32), ("node_d2", 10), ("lv_ude", 10)):
        for loss in ("multi_shoot", "multi_shoot_preroll", "shoot_theta", "shoot_data", "cond_lik"):
            prob, theta = cases.make_case(rhs, loss, 0, hidden=hidden, lengths=(30, 7, 12, 1, 26, 3, 5, 6, 11))
            L_ref, g_ref = oracle.loss_grad(prob, theta)
            n = prob.n_theta
            h_theta = torch.from_numpy(theta.copy()).pin_memory()
            h_grad = torch.zeros(n, dtype=torch.float64).pin_memory()
            h_loss = torch.zeros(1, dtype=torch.float64).pin_memory()
            with ude.Engine(prob) as eng:
                eng.set_option(ude.OPT_SPARSE_UHAT_IO, 1)
                for rep in range(2):        # the second call reuses the buffer: entries the library does not write stay 0
                    eng.loss_grad_ptr(h_theta.data_ptr(), n, h_loss.data_ptr(), h_grad.data_ptr())
                    g = h_grad.numpy()
                    assert abs(h_loss[0].item() - L_ref[0]) < TOL * abs(L_ref[0]), (rhs, loss)
                    assert cases.grad_mismatch(prob, g, g_ref, rtol=TOL)[0] <= 1.0, (rhs, loss, rep)
                    assert np.all(g[g_ref == 0.0] == 0.0)
                h2d, d2h = eng.last_io_bytes()
                if loss == "cond_lik":
                    assert h2d == 8 * n
                else:
                    cols = 0 if loss == "shoot_data" else prob.count_segments()
                    assert h2d == 8 * (prob.d * cols + prob.n_pm) and d2h == h2d + 8, (rhs, loss, h2d)
                # pageable buffers: same option, dense copies, same numbers
                L, g2 = eng.loss_grad(theta)
                assert cases.grad_mismatch(prob, g2, g_ref, rtol=TOL)[0] <= 1.0
                assert eng.last_io_bytes()[0] == 8 * n


def test_forward_states(ude):
    """shooting_states / multiple_shooting_states / one-step predictions (LFC.jl:237-258, :310-341; ModelTesting.jl:181-195)."""
    for loss in ("multi_shoot", "multi_shoot_preroll", "shoot_theta", "shoot_data", "cond_lik"):
        prob, theta = cases.make_case("node_d2x1", loss, 0, lengths=(11, 1, 6, 16), skip=False)
        ref = oracle.forward(prob, theta)
        with ude.Engine(prob) as eng:
            out = eng.forward(theta)
        assert cases.relerr(out, ref) < TOL, loss


def test_repeatable_bitwise(ude):
    """Single-model reductions have a fixed order: two evaluations give bit-identical process-model gradients."""
    prob, theta = ude.synthetic.config_c3(n_series=300)
    with ude.Engine(prob) as eng:
        L1, g1 = eng.loss_grad(theta)
        L2, g2 = eng.loss_grad(theta)
    pmb = prob.pm_base(0)
    assert L1[0] == L2[0]
    assert np.array_equal(g1[pmb:], g2[pmb:])
    assert np.array_equal(g1[:pmb], g2[:pmb])     # <= 2 atomic contributions per uhat entry: order-independent


def test_error_paths(ude):
    prob, theta = cases.make_case("node_d2", "cond_lik", 0)
    with ude.Engine(prob) as eng:
        with pytest.raises(ude.UdeError):
            eng.loss_grad(theta[:-1])
    bad, _ = cases.make_case("node_d2", "cond_lik", 0, hidden=300)       # no compiled kernel that wide for d = 2
    with pytest.raises(ude.UdeError) as ei:
        with ude.Engine(bad) as eng:
            eng.loss_grad(np.zeros(bad.n_theta))
    assert ei.value.code == -2
    th = theta.copy(); th[prob.pm_base(0)] = np.inf
    with ude.Engine(prob) as eng:
        L, g = eng.loss_grad(th)
    assert not np.isfinite(L[0])


def test_adam_matches_host_loop(ude):
    """Device-resident Adam == the same update applied on the host with oracle gradients."""
    prob, theta = cases.make_case("lv_ude", "cond_lik", 0, lengths=(12, 9))
    with ude.Engine(prob) as eng:
        th_dev, trace = eng.adam(theta, 0.01, 5)
    th = theta.copy(); m = np.zeros_like(th); v = np.zeros_like(th)
    for it in range(1, 6):
        L, g = oracle.loss_grad(prob, th)
        assert abs(trace[it - 1, 0] - L[0]) / abs(L[0]) < 1e-9
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        th = th - 0.01 * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
    assert cases.relerr(th_dev, th) < 1e-9


@pytest.mark.parametrize("loss", ["multi_shoot", "cond_lik", "mse", "shoot_theta"])
@pytest.mark.parametrize("chunks", [2, 3, 7])
def test_pipelined_host_path(ude, loss, chunks, monkeypatch):
    """ude_loss_grad overlaps H2D(uhat chunk k+1) / kernel(chunk k) / D2H(grad chunk k-1); chunk borders fall inside
    series here, so columns shared by two chunks exercise the upload/download bookkeeping."""
    monkeypatch.setenv("UDE_HOST_CHUNKS", str(chunks))
    monkeypatch.setenv("UDE_HOST_CHUNK_ROUNDS", "0")      # tiny problem: keep the requested number of chunks (no rounding to whole grid rounds)
    for rhs, kw in (("node_d4x1", dict(hidden=32)), ("lv_ude", {})):
        prob, theta = cases.make_case(rhs, loss, 0, lengths=(13, 2, 30, 7, 1, 22, 9, 16, 5, 11, 3, 8), skip=(loss == "cond_lik"), **kw)
        _check(prob, theta, ude)


def test_pipelined_host_path_whole_rounds(ude, monkeypatch):
    """default chunk plan (whole rounds of the persistent grid per chunk) on a problem big enough to be split: same result as
    the resident path and as one chunk."""
    prob, theta = ude.synthetic.config_c3(n_series=20_000)
    theta = theta + 0.01 * np.random.default_rng(3).normal(size=theta.shape)
    res = []
    for k in ("12", "1"):
        monkeypatch.setenv("UDE_HOST_CHUNKS", k)
        with ude.Engine(prob) as eng:
            res.append(eng.loss_grad(theta))
    assert abs(res[0][0][0] - res[1][0][0]) < 1e-12 * abs(res[1][0][0])
    assert cases.relerr(res[0][1], res[1][1]) < 1e-12


@pytest.mark.parametrize("loss", ["cond_lik", "multi_shoot", "shoot_theta", "deriv_match"])
@pytest.mark.parametrize("chunks", [2, 3, 5])
def test_pipelined_host_path_many_models(ude, loss, chunks, monkeypatch):
    """Many-model handles (folds x members, src/CrossValidation.jl:17-45): ude_loss_grad moves theta up and gradient + losses
    down in runs of whole models while the neighbouring run is on the SMs; every run is finalised behind its own kernel."""
    monkeypatch.setenv("UDE_HOST_CHUNKS", str(chunks))
    monkeypatch.setenv("UDE_HOST_CHUNK_ROUNDS", "0")
    prob, theta = cases.make_case("node_time_d3", loss, 0, lengths=(9, 8, 7, 6, 5, 4, 11, 3, 2, 12, 6, 6, 3, 20), n_models=7, skip=(loss == "cond_lik"))
    _check(prob, theta, ude)
    prob, theta = cases.make_case("node_d4x1", loss, 1, hidden=32, lengths=(9, 3, 7, 6, 5, 4), n_models=3)
    _check(prob, theta, ude)


def test_pipelined_host_path_many_models_whole_rounds(ude, monkeypatch):
    """default plan (chunk borders at model boundaries just below whole rounds of the persistent grid) on a C4-shaped handle
    big enough to be split: per-model losses and gradients equal the single-chunk evaluation."""
    prob, theta = ude.synthetic.config_c4(n_members=300, n_folds=10)
    res = []
    for k in ("4", "1"):
        monkeypatch.setenv("UDE_HOST_CHUNKS", k)
        with ude.Engine(prob) as eng:
            res.append(eng.loss_grad(theta))
    assert res[0][0].shape == res[1][0].shape and np.max(np.abs(res[0][0] - res[1][0]) / np.abs(res[1][0])) < 1e-12
    assert cases.relerr(res[0][1], res[1][1]) < 1e-12


@pytest.mark.parametrize("rhs,hidden", [("node_d1", 32), ("node_d2", 29), ("node_d3", 32), ("node_d1x1", 20), ("node_d2x1", 32),
                                        ("node_d3x1", 17), ("node_d4", 32), ("node_d4x1", 128), ("node_d4", 100), ("node_d2", 128)])
def test_parity_dmma_shapes(ude, rhs, hidden, monkeypatch):
    """The FP64 tensor-core kernels for every instantiated NODE shape (hidden width not a multiple of 8 included)."""
    monkeypatch.setenv("UDE_KERNEL", "mma")
    for loss, solver in (("multi_shoot", 0), ("cond_lik", 1), ("deriv_match", 0), ("shoot_theta", 1)):
        prob, theta = cases.make_case(rhs, loss, solver, hidden=hidden, lengths=(9, 1, 14, 2, 6, 11, 3, 8, 5), skip=(loss == "cond_lik"))
        with ude.Engine(prob) as eng:
            assert "mma" in eng.kernel_name()
        _check(prob, theta, ude)


def test_long_shooting_series(ude):
    """One long trajectory per series (shooting loss): thousands of RK steps per segment, the reverse-sweep stash is sized
    by the longest segment and the persistent grid shrinks to fit it."""
    for rhs, kw in (("node_d2", dict(hidden=10)), ("node_d4x1", dict(hidden=32)), ("lv_ude", {})):
        prob, theta = cases.make_case(rhs, "shoot_theta", 0, lengths=(1500, 700, 64, 3), substeps=2, **kw)
        # tame the dynamics so a 1500-point open-loop trajectory stays bounded
        theta[prob.pm_base(0):] *= 0.2
        _check(prob, theta, ude)


def test_adam_cuda_graph_equals_plain_loop(ude, monkeypatch):
    """ude_adam replays one captured iteration as a CUDA graph; bit-identical to the launch-by-launch loop."""
    import time
    prob, theta = ude.synthetic.config_c1()
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("UDE_ADAM_GRAPH", flag)
        with ude.Engine(prob) as eng:
            eng.adam(theta, 0.025, 3)                       # warm-up (plan, module load)
            t0 = time.perf_counter()
            th, tr = eng.adam(theta, 0.025, 200)
            res[flag] = (th, tr, time.perf_counter() - t0, eng.adam_used_graph())
    assert res["1"][3] and not res["0"][3]
    assert np.array_equal(res["1"][0], res["0"][0]) and np.array_equal(res["1"][1], res["0"][1])
    assert res["1"][1][-1, 0] < res["1"][1][0, 0]
    print("C1 (59 segments), 200 Adam iterations: graph %.2f ms, plain loop %.2f ms" % (1e3 * res["1"][2], 1e3 * res["0"][2]))


TOL32 = 1e-4   # relative, optional FP32 mode (north_star)


@pytest.mark.parametrize("rhs,hidden,loss", [("node_d2", 10, "multi_shoot"), ("node_d2", 10, "cond_lik"), ("node_d4x1", 10, "multi_shoot"),
                                             ("node_d4x1", 32, "multi_shoot"), ("node_d4x1", 32, "mse"), ("node_d4x1", 32, "deriv_match"),
                                             ("node_time_d3", 10, "cond_lik"), ("lv_ude_abs", 10, "cond_lik"), ("lv_ude", 10, "shoot_theta"),
                                             ("node_d8", 128, "cond_lik"), ("node_d8", 20, "multi_shoot")])
def test_parity_fp32_mode(ude, rhs, hidden, loss):
    """FP32 arithmetic (times, covariates, loss terms, seeds and reductions in FP64) against the FP64 oracle: 1e-4."""
    for solver in (0, 1):
        prob, theta = cases.make_case(rhs, loss, solver, hidden=hidden, lengths=(14, 1, 9, 2, 30, 6, 11), skip=(loss == "cond_lik"))
        L_ref, g_ref = oracle.loss_grad(prob, theta)
        with ude.Engine(prob, dtype=1) as eng:
            assert eng.kernel_name().startswith("f32_")
            L, g = eng.loss_grad(theta)
        assert np.max(np.abs(L - L_ref) / np.abs(L_ref)) < TOL32, (L, L_ref)
        assert cases.relerr(g, g_ref) < TOL32, cases.relerr(g, g_ref)


@pytest.mark.parametrize("loss", ["cond_lik", "mse", "multi_shoot", "multi_shoot_preroll", "shoot_theta", "shoot_data"])
def test_parity_fp32_tensor_core_kernel(ude, loss):
    """FP32 mode on the wide NODE (C5 shape, d = 8): the 3xTF32 tensor-core kernel (csrc/ude_kernel_tf32.cuh, 16 segments per
    block, hidden layer split over four warps) against the FP64 oracle at the FP32-mode bar, per gradient leaf; ragged series,
    hidden widths below the padded 128, both solvers, and the forward-only trajectories."""
    for solver in (0, 1):
        for hidden, lengths, substeps in ((128, (14, 1, 9, 2, 30, 6, 11, 3, 3, 5, 17, 8, 2, 2, 4, 6, 7, 9, 1, 12), 2), (100, (5, 4, 3), 3), (9, (6,) * 37, 1)):
            prob, theta = cases.make_case("node_d8", loss, solver, hidden=hidden, lengths=lengths, substeps=substeps,
                                          skip=(loss in ("cond_lik", "mse")))
            L_ref, g_ref = oracle.loss_grad(prob, theta)
            with ude.Engine(prob, dtype=1) as eng:
                assert "tf32" in eng.kernel_name()
                L, g = eng.loss_grad(theta)
                out = eng.forward(theta)
            assert np.max(np.abs(L - L_ref) / np.abs(L_ref)) < TOL32, (L, L_ref)
            assert cases.grad_mismatch(prob, g, g_ref, rtol=TOL32, floor=1e-1)[0] <= 1.0, cases.grad_mismatch(prob, g, g_ref, rtol=TOL32, floor=1e-1)
            assert np.all(g[g_ref == 0.0] == 0.0)
            assert cases.relerr(out, oracle.forward(prob, theta)) < TOL32


@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_full_size_configs_sampled_and_additive(ude, name):
    """BASELINE.json's configurations at FULL size (C2 10^4 series, C3 10^5 series, C4 10^4 models, C5 10^6 segments),
    where the oracle cannot evaluate the whole problem in test time.  Size-independent properties instead:
      * a series' uhat gradient depends on that series and the shared parameters only, so contiguous blocks of series cut
        out of the full problem (sharding.shard_problem) and evaluated by the ORACLE must reproduce those slices of the
        full-size GPU gradient (1e-10); for the many-model C4 the blocks are whole models: loss and every gradient entry;
      * loss and shared-parameter gradient are additive over a partition of the series: GPU(first half) + GPU(second
        half) = GPU(all), the regulariser counted once (1e-11);
      * the evaluation is bitwise repeatable (single-model handles; many-model handles accumulate per-model sums with
        FP64 atomics and repeat to rounding)."""
    prob, theta = ude.synthetic.CONFIGS[name]()
    theta = theta + 0.01 * np.random.default_rng(11).normal(size=theta.shape)
    with ude.Engine(prob) as eng:
        L, g = eng.loss_grad(theta)
        L2, g2 = eng.loss_grad(theta)
    if prob.n_models == 1:
        assert np.array_equal(L, L2) and np.array_equal(g, g2)
    else:   # many models: per-model sums are FP64 atomics across the warps that share a model -- order-dependent rounding only
        assert cases.relerr(L, L2) < 1e-13 and cases.relerr(g, g2) < 1e-13
    assert np.all(np.isfinite(L)) and np.all(np.isfinite(g))
    rng = np.random.default_rng(5)
    if prob.n_models > 1:
        world = prob.n_models // 4
        for r in [0, world - 1] + list(rng.integers(1, world - 1, 3)):
            sh = ude.sharding.shard_problem(prob, theta, int(r), world)
            Lo, go = oracle.loss_grad(sh.problem, sh.theta)
            m0, m1 = sh.models
            assert np.max(np.abs(L[m0:m1] - Lo) / np.abs(Lo)) < TOL
            lo, hi = int(prob.theta_base[m0]), (int(prob.theta_base[m1]) if m1 < prob.n_models else prob.n_theta)
            assert cases.relerr(g[lo:hi], go) < TOL
        return
    world = prob.n_series // 64
    n_u = prob.d * prob.n_cols
    for r in [0, world - 1] + list(rng.integers(1, world - 1, 3)):
        sh = ude.sharding.shard_problem(prob, theta, int(r), world)
        _, go = oracle.loss_grad(sh.problem, sh.theta)
        glo, ghi, llo = sh.uhat_slices[0]
        assert cases.relerr(g[glo:ghi], go[llo:llo + (ghi - glo)]) < TOL, (name, r)
    parts = []
    for r in range(2):
        sh = ude.sharding.shard_problem(prob, theta, r, 2)
        with ude.Engine(sh.problem) as eng:
            Lr, gr = eng.loss_grad(sh.theta)
        parts.append((sh, Lr, gr))
    assert abs(parts[0][1][0] + parts[1][1][0] - L[0]) < 1e-11 * abs(L[0])
    pm_sum = sum(gr[sh.uhat_slices[0][1] - sh.uhat_slices[0][0]:][:sh.pm_shared] for sh, _, gr in parts)
    assert cases.relerr(pm_sum, g[n_u:n_u + parts[0][0].pm_shared]) < 1e-11
    for sh, _, gr in parts:
        glo, ghi, llo = sh.uhat_slices[0]
        assert cases.relerr(gr[llo:llo + (ghi - glo)], g[glo:ghi]) < 1e-12


@pytest.mark.parametrize("loss", ["cond_lik", "mse"])
@pytest.mark.parametrize("rhs", list(cases.RHS_SPECS))
def test_parity_euler_state_space(ude, rhs, loss):
    """UDE_EULER (one RHS evaluation per interval): the solver behind the spline-gradient-matching loss (LFC.jl:888-931)."""
    prob, theta = cases.make_case(rhs, loss, 2, substeps=1, skip=(loss == "cond_lik"))
    _check(prob, theta, ude)
    prob, theta = cases.make_case(rhs, loss, 2, substeps=3, hidden=32 if rhs == "node_d4x1" else 10)
    _check(prob, theta, ude)


def test_independent_handles_on_host_threads(ude):
    """SURVEY 8b threading contract: independent handles are usable concurrently from different host threads (the reference
    calls training! from Threads.@threads fold tasks, src/CrossValidation.jl:30)."""
    import threading
    specs = [("node_d2", "multi_shoot", 10), ("node_d4x1", "cond_lik", 32), ("lv_ude_abs", "mse", 10), ("node_time_d3", "shoot_theta", 10)]
    probs = [cases.make_case(r, l, 0, hidden=h, lengths=(9, 4, 12, 7, 3, 10), seed=i) for i, (r, l, h) in enumerate(specs)]
    refs = [oracle.loss_grad(p, th) for p, th in probs]
    out, errs = [None] * len(probs), []

    def work(i):
        try:
            p, th = probs[i]
            with ude.Engine(p) as eng:
                for _ in range(20):
                    out[i] = eng.loss_grad(th)
        except Exception as ex:      # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(probs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for (L, g), (Lr, gr) in zip(out, refs):
        assert abs(L[0] - Lr[0]) < TOL * abs(Lr[0]) and cases.relerr(g, gr) < TOL


def test_error_paths_of_the_wider_api(ude):
    """UNSUPPORTED / INVALID / STATE answers instead of wrong numbers: Euler with a trajectory loss, the UKF loss on a many-model
    handle, a forecast chain on a handle whose series are not single intervals, peer attach without export, FP32 + tensor-core
    request."""
    prob, theta = cases.make_case("node_d2", "multi_shoot", 2)                      # UDE_EULER is state-space only
    with pytest.raises(ude.UdeError) as ei:
        with ude.Engine(prob) as eng:
            eng.loss_grad(theta)
    assert ei.value.code == -2
    many, th_many = cases.make_case("node_d2", "cond_lik", 0, lengths=(5, 4, 6, 3), n_models=2)
    with ude.Engine(many) as eng:
        Lm, _, _ = eng.ukf_loss_grad(th_many, 0.3 * np.eye(2), 0.1 * np.eye(2), alpha=1.0)     # many-model UKF is supported (round 2)
        assert Lm.shape == (2,) and np.all(np.isfinite(Lm))
        with pytest.raises(ude.UdeError):
            eng.forecast_chain(th_many, np.ones(2), np.ones(2), np.zeros(2), np.ones(2) * 0.5, 0.25, 0.1)
        with pytest.raises(ude.UdeError):
            eng.peer_attach([b"\0" * 128, b"\0" * 128], 2, 0)
        with pytest.raises(ude.UdeError):
            eng.peer_export(1)
    sp, th_sp = cases.make_case("series_growth_d1", "cond_lik", 0, hidden=4, lengths=(5, 4))      # per-series parameter vector: no UKF path
    with ude.Engine(sp) as eng:
        with pytest.raises(ude.UdeError) as ei:
            eng.ukf_loss_grad(th_sp, 0.3 * np.eye(1), 0.1 * np.eye(1))
        assert ei.value.code == -2
    one, th_one = cases.make_case("node_d2", "cond_lik", 0, lengths=(1,))           # a single point: nothing to filter
    with ude.Engine(one) as eng:
        with pytest.raises(ude.UdeError):
            eng.ukf_loss_grad(th_one, 0.3 * np.eye(2), 0.1 * np.eye(2))
        L, g = eng.loss_grad(th_one)                                                # ...but the ordinary loss is defined (observation term)
        assert np.isfinite(L[0])
    wide, th_w = cases.make_case("node_d8", "cond_lik", 0, hidden=128, lengths=(4, 3))
    with ude.Engine(wide, dtype=1) as eng:                                          # FP32 mode never picks a DMMA kernel
        assert eng.kernel_name().startswith("f32_")


def test_no_device_memory_leak_over_handle_lifetimes(ude):
    """create -> evaluate (loss+grad, forward, Adam, UKF with its auxiliary handle, forecast chain) -> destroy, many times: the
    free device memory reported by the driver must come back."""
    import torch
    prob, theta = cases.make_case("node_d4x1", "multi_shoot", 0, hidden=32, lengths=(9, 4, 12, 7, 3, 10))
    sprob, stheta = cases.make_case("node_d2", "cond_lik", 0, hidden=5, lengths=(7, 4))

    def cycle():
        with ude.Engine(prob) as eng:
            eng.loss_grad(theta); eng.forward(theta); eng.adam(theta, 1e-3, 3)
        with ude.Engine(sprob) as eng:
            eng.ukf_loss_grad(stheta, 0.3 * np.eye(2), 0.1 * np.eye(2), alpha=1.0)
            eng.ukf_filter(stheta, 0.3 * np.eye(2), 0.1 * np.eye(2), alpha=1.0)

    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(40):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 * 2 ** 20, (free0, free1)      # the driver's own pools may keep a little


@pytest.mark.parametrize("rhs,hidden,n_models", [("lv_ude", 20, 1), ("node_time_d3", 24, 3), ("fishery", 32, 1), ("lv_ude_abs", 16, 2),
                                                 ("series_growth_d1", 16, 1), ("node_d4", 100, 2), ("node_d2x2", 32, 1)])
def test_catalog_shapes_compiled_on_demand(ude, rhs, hidden, n_models):
    """Hidden widths / models per handle without a pre-compiled instance (the parametric UDEs above 12 hidden units, NODE shapes
    above 12 units on a many-model handle): Engine compiles the same hand-written trait for that geometry on first use
    (user_rhs.ensure_catalog_kernels) instead of answering UNSUPPORTED; parity bar unchanged.  tools/fuzz_parity.py draws 800
    random (right-hand side, loss, solver, width, lengths, models) cases through the same path: profiles/fuzz_r02.log."""
    for loss, solver in (("cond_lik", 0), ("multi_shoot", 0), ("deriv_match", 0)):
        prob, theta = cases.make_case(rhs, loss, solver, hidden=hidden, lengths=(6, 3, 1, 5, 4, 2), n_models=n_models, skip=(loss == "cond_lik"))
        with ude.Engine(prob) as eng:
            assert eng.kernel_name().endswith("_jit")
        _check(prob, theta, ude)
