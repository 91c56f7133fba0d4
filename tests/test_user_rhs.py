"""User-written right-hand sides as expressions (universaldiffeq.jl_b200/user_rhs.py): code generation, the NumPy twin, and that
the generated trait compiles into a plugin of libude_cuda.so whose kernels land in the library's registry.  No GPU needed: nvcc
cross-compiles and the plugin's static initialiser only appends to a host-side table."""
import os

import numpy as np
import pytest

import cases
import user_rhs_cases as urc
from oracle import ude_oracle_np as onp


def test_generated_trait_and_twin(ude):
    rhs = urc.lv_twin(ude)
    assert rhs.kind >= 1000 and rhs.kind == urc.lv_twin(ude).kind            # the id is a hash of the generated source
    assert "du[0] = p_r*u1 - y1;" in rhs.cuda_trait and "gsc[2] += kb1*y1;" in rhs.cuda_trait
    assert "NEED_Y = true" in rhs.cuda_trait
    u, y, sc = np.array([0.7, 1.3]), np.array([0.4]), [0.1, 0.2, 0.3]
    assert np.allclose(rhs.du_numpy(u, [], y, sc), [0.1 * 0.7 - 0.4, 0.3 * 0.4 - 0.2 * 1.3])
    ep = urc.epidemic(ude)
    assert ep.nn_in == 3 + 1 + 1 and ep.has_time_input and "sgn1(" in ep.cuda_trait and "fabs(" in ep.cuda_trait
    gr = urc.growth_twin(ude)
    assert gr.series_param == "r" and "SERIES = true" in gr.cuda_trait and "gseries += kb0*u1*y1;" in gr.cuda_trait
    assert np.allclose(gr.du_numpy([2.0], [], [0.5], [], 0.3), [0.3 * 2.0 * 0.5])
    with pytest.raises(ValueError):
        ude.expression_rhs(d=1, nn_out=1, du=["q*u1"])                        # unknown symbol
    with pytest.raises(ValueError):
        ude.expression_rhs(d=1, nn_out=1, scalars=("u1",), du=["u1"])         # shadows a state
    with pytest.raises(ValueError):
        ude.expression_rhs(d=2, nn_out=1, du=["u1"])                          # one expression per state
    with pytest.raises(ValueError):
        ude.expression_rhs(d=9, nn_out=1, du=["y1"] * 9)                      # UDE_MAX_D


@pytest.mark.parametrize("loss", ["cond_lik", "multi_shoot", "deriv_match"])
def test_twin_of_a_catalog_entry_matches_the_oracles(ude, loss):
    """the LV UDE written as expressions: the NumPy twin evaluates the user's du; same loss as the catalog kind in the NumPy
    twin AND in the C oracle"""
    from oracle import oracle as c_oracle
    rhs = urc.lv_twin(ude)
    prob_u, theta = urc.make(ude, "user_lv", rhs, (1.0, 0.5, 0.5), loss, 0, hidden=6, lengths=(6, 1, 4))
    prob_c, theta_c = cases.make_case("lv_ude", loss, 0, hidden=6, lengths=(6, 1, 4))
    assert np.array_equal(theta, theta_c)
    Lu = onp.loss(prob_u, theta_c)
    Lc = onp.loss(prob_c, theta_c)
    L_c, _ = c_oracle.loss_grad(prob_c, theta_c)
    assert abs(Lu[0] - Lc[0]) <= 1e-14 * abs(Lc[0]) and abs(Lu[0] - L_c[0]) <= 1e-12 * abs(L_c[0])


def test_plugin_compiles_and_registers(ude):
    """nvcc turns the generated trait into the fused kernels (both loss families + derivative matching, with and without the
    reverse sweep) and the plugin loads; a second request is served from the cache"""
    rhs = urc.epidemic(ude)
    so = ude.user_rhs.ensure_kernels(rhs, 6, 0)
    assert os.path.exists(so) and so in ude.user_rhs._loaded
    assert ude.user_rhs.ensure_kernels(rhs, 6, 0) == so
    src = open(so[:-3] + ".cu").read()
    assert f"register_one<F64, RhsUser{rhs.kind}, 2, 3, 2, 0, LK_TRAJ, false>" in src and f"KIND = {rhs.kind}" in src


def test_one_hot_series_input(ude):
    """MultiTimeSeriesNetwork semantics (src/NeuralNetworkConstructors.jl:69-82): NN(vcat(x, distance * onehot(i))) -- here as
    constant covariates, one per series; the right-hand side the oracle evaluates must be exactly that"""
    model, nn, df = urc.one_hot_model(ude, m=3, distance=0.5)
    X = model.X_data_frame
    assert list(X.columns[2:]) == ["series_is_s00", "series_is_s01", "series_is_s02"] and X.shape == (3, 5)
    p = ude.build_problem(model, "conditional likelihood")
    assert p.n_cov == 3 and p.nn_in == 5
    W1, b1, W2, b2 = (np.asarray(nn[k]) for k in ("layer_1.weight", "layer_1.bias", "layer_2.weight", "layer_2.bias"))
    u = np.array([0.7, 1.1])
    for i in range(3):
        oh = np.zeros(3); oh[i] = 0.5
        ref = W2 @ np.tanh(W1 @ np.concatenate([u, oh]) + b1) + b2
        for t in (-3.0, 0.4, 50.0):                   # constant in time, before and after the single knot
            assert np.allclose(onp.rhs(p, model.process_model, i, i, u, t), ref, rtol=1e-14, atol=0)
